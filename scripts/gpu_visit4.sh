#!/bin/bash
mkdir -p gpurun_out/ncu
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['split'], d['pipelined']['proofs_per_s'], d['msm']['ms'], d['batch52']['proofs_per_s'], d['sigma']['proofs_per_s'])
PY
cat > /tmp/sig_prof.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
import bench
ctx = pkg.Context(0)
inst = bench.make_instance(ctx, 4, 13, seed=1)
ctx.set_params(4, 13, inst["enc_g"], inst["ck_g"], inst["ck_h"], inst["ghat"])
print(bench.sigma_bench(pkg, ctx, 65536, False))
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_lincomb" -s 6 -c 6 -f -o /tmp/prof_lincomb python /tmp/sig_prof.py > /tmp/ncu_lincomb.log 2>&1
tail -2 /tmp/ncu_lincomb.log | cut -c 1-200
ncu -i /tmp/prof_lincomb.ncu-rep --page raw --csv > gpurun_out/ncu/k_lincomb_raw.csv 2>/dev/null
ncu -i /tmp/prof_lincomb.ncu-rep --page details --csv > gpurun_out/ncu/k_lincomb_details.csv 2>/dev/null
ls -la gpurun_out/ncu | tail -3
