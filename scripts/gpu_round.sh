#!/bin/bash
# One GPU visit: sanitizer on a small proof, parity suite, smoke, microbench, bench, launch list.
set -x
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --print-limit 3 python scripts/repro.py 5 3 2>&1 | grep -v "Host Frame\|^=========         in " | head -40
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_shuffle_2p16.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --msm-logn 16 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c 1-300
