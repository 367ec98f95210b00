"""(design-time tool, not part of the product build: compile tools/probe/probe.cu into its own library first --
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC tools/probe/probe.cu
open-solution-mapping-challenge_b200/csrc/host_common.cu -o gpurun_out/libprobe.so -- and point L.lib at it.)
Runs tools/probe/probe.cu over row shifts / stride offsets / base-offset modes and reports which descriptor forms read the
expected rows of a swizzled TMA tile."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mcb200
from mcb200 import _lib as L

dev = torch.device("cuda:0")
L.lib.mcb_debug_umma_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
for rowb in (128, 64):
    k = rowb // 2
    R = 256
    r = torch.arange(R).view(R, 1).float()
    c = torch.arange(k).view(1, k).float()
    G = torch.where((torch.arange(k) % 2 == 0).view(1, k), r.expand(R, k), c.expand(R, k)).contiguous()
    a = G.to(dev, torch.bfloat16).contiguous()
    ident = torch.eye(k, device=dev, dtype=torch.bfloat16).contiguous()
    for sbo_rows in (8, 10):
        for shift in (0, 1, 2, 3, 5, 8, 11, 21):
            res = []
            for mode in (0, 1):
                out = torch.full((128, k), -1.0, device=dev)
                L.fcall("mcb_debug_umma_probe", a.data_ptr(), ident.data_ptr(), out.data_ptr(), R, rowb, shift, sbo_rows * rowb, mode)
                torch.cuda.synchronize()
                m = torch.arange(128)
                rows = shift + (m // 8) * sbo_rows + m % 8
                exp = G[rows]
                ok = torch.equal(out.cpu(), exp)
                nbad = int((out.cpu() != exp).sum())
                res.append("mode%d:%s(%d bad)" % (mode, "OK" if ok else "BAD", nbad))
            print("rowb %3d sbo_rows %2d shift %2d  %s" % (rowb, sbo_rows, shift, "  ".join(res)), flush=True)
