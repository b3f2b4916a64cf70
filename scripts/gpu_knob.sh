#!/bin/bash
mkdir -p gpurun_out
for mb in 4 5 6; do
  MP_ACC_MINBLOCKS=$mb python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/bench_mb$mb.json 2> gpurun_out/bench_mb$mb.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_mb$mb.json"))
print("mb=$mb ms/step %.2f  acc adds/s %.3e  acc share %.3f  msm ms %.3f  msm acc adds/s %.3e" % (d["ms_per_step"], d["roofline"]["ec_adds_per_s"], d["roofline"]["share_of_step"], d["msm"]["ms"], d["msm"]["accumulate_adds_per_s"]))
PY
done
