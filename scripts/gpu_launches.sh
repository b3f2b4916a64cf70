#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_shuffle_2p16.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch52 0 --msm-logn 16 > gpurun_out/ncu_bench.log 2>&1
