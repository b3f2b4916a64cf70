"""One warm-up + one measured sequential shuffle_and_remask + verify_shuffle of the headline deck (and, with
`msm`, one 2^20 MSM) -- the command the ncu launch lists under profiles/ are taken from:
    ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
        python scripts/one_step.py [m n] [msm]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as g
import bench
pkg = g.load_package()
args = [a for a in sys.argv[1:] if a != "msm"]
m, n = (int(args[0]), int(args[1])) if len(args) >= 2 else (128, 512)
ctx = pkg.Context(0)
inst = bench.make_instance(ctx, m, n, 1)
ctx.set_params(m, n, inst["enc_g"], inst["ck_g"], inst["ck_h"], inst["ghat"])
for it in range(2):
    torch.cuda.synchronize()
    deck2, proof = ctx.shuffle_and_remask(inst["pk"], inst["deck"], inst["perm"], inst["rho"], inst["rand"])
    assert ctx.verify_shuffle(inst["pk"], inst["deck"], deck2, proof) == 0
if "msm" in sys.argv:
    dev = torch.device("cuda:0")
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    print(bench.msm_microbench(ctx, torch, dev, stream, 20, pkg))
ctx.close()
