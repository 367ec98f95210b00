import os, sys
sys.path.insert(0, '/root/repo')
sys.argv=['x','none']
exec(open('/root/repo/tools/sweep_gemm.py').read().split("SHAPES = [")[0])
for shp in [(64,256,1,80),(256,1024,1,20),(64,64,3,80),(128,128,3,160),(1024,256,1,20)]:
    for st in (True, False):
        fn, fl = fwd(*shp, stats=st)
        us = timed(fn)
        print(shp, "stats" if st else "nostats", "%.1f us %.0f TF" % (us, fl/us/1e6), flush=True)
