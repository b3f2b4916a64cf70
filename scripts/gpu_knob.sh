#!/bin/bash
mkdir -p gpurun_out
for pers in 1 0; do
  MP_ACC_PERSISTENT=$pers timeout 200 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --batch52 0 --pipeline-decks 0 > gpurun_out/bench_p$pers.json 2> gpurun_out/bench_p$pers.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_p$pers.json") if l.startswith("{")][-1])
print("persistent=$pers ms/step %.2f  acc adds/s %.3e  msm ms %.3f  msm acc adds/s %.3e  prove %.1f" % (d["ms_per_step"], d["roofline"]["ec_adds_per_s"], d["msm"]["ms"], d["msm"]["accumulate_adds_per_s"], d["split"]["prove_ms"]))
PY
done
