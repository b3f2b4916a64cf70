"""BLS12-377 section of bench.py alone (no CPU leg), for occupancy / tuning probes:
   MP_ACC_MINBLOCKS=3|4|5 python scripts/bls12_377_probe.py [logn] [cpu]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as g
import bench
pkg = g.load_package()
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 18
cpu = len(sys.argv) > 2 and sys.argv[2] == "cpu"
r = bench.bls12_377_bench(pkg, torch, torch.device("cuda:0"), logn, cpu)
print(json.dumps({"minblocks": os.environ.get("MP_ACC_MINBLOCKS", "default"), "logn": logn,
                  "msm_ms": r["msm"]["ms"], "acc_adds_per_s": r["msm"]["accumulate_adds_per_s"], "acc_ms": r["msm"]["accumulate_ms_avg"],
                  "ct_ms": r["ct_msm"]["ms"], "pedersen_ms": r["pedersen"]["ms"], "mb": r["microbench"],
                  "shapes": r["reference_benchmark_shape"]["shapes"]}))
