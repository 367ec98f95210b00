"""one conv GEMM launch of a chosen shape, for `ncu --set full`:  python tools/ncu_one.py fwd 64 256 1 80 [stats]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mcb200
from mcb200 import ops

kind, cin, cout, k, hw = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
N = int(os.environ.get("BATCH", "32"))
dev = torch.device("cuda:0")
BF = torch.bfloat16
mk = lambda *s: (torch.randn(*s, device=dev) * 0.1).to(BF)
if kind == "fwd":
    x, w = mk(N, hw, hw, cin), mk(k * k, cout, cin)
    y = torch.empty(N, hw, hw, cout, dtype=BF, device=dev)
    st = torch.zeros(2 * cout, device=dev)
    fn = lambda: ops.conv_fwd(x, w, k, 1, stats=st, out=y)
elif kind == "dgradF":
    dy, w = mk(N, hw, hw, cout), mk(k * k, cout, cin)
    dx = torch.empty(N, hw, hw, cin, dtype=BF, device=dev)
    z = mk(N, hw, hw, cin)
    one, zero = torch.ones(cin, device=dev), torch.zeros(cin, device=dev)
    db, dg = torch.zeros(cin, device=dev), torch.zeros(cin, device=dev)
    fn = lambda: ops.conv_dgrad(dy, w, k, 1, (hw, hw), out=dx, bn_reduce=(z, zero, one, one, zero, db, dg))
elif kind == "wgrad":
    dy, x = mk(N, hw, hw, cout), mk(N, hw, hw, cin)
    dw = torch.zeros(k * k, cout, cin, device=dev)
    fn = lambda: ops.conv_wgrad(dy, x, dw, k, 1)
for _ in range(3):
    fn()
torch.cuda.synchronize()
