"""Run a few device-resident MSMs (for ncu launch lists / captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g
pkg = g.load_package()
from oracle.py import stark
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 20
c = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = pkg.Context(0)
dev = torch.device("cuda:0")
n = 1 << logn
pts = []
cur = stark.G
for _ in range(256):
    pts.append(cur); cur = stark.add(cur, stark.G)
base = torch.frombuffer(bytearray(b"".join(map(stark.point_to_bytes64, pts))), dtype=torch.uint8).to(dev)
bases = base.repeat(n // 256).contiguous()
torch.manual_seed(1)
scal = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev)
scal[:, 31] &= 0x07
out = torch.zeros(64, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
for _ in range(reps):
    ctx.msm_g1_device(bases.data_ptr(), scal.data_ptr(), n, out.data_ptr(), c)
    ctx.sync()
print("done", bytes(out.cpu().numpy().tobytes()).hex()[:32])
