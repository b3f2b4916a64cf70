#!/bin/bash
# quick GPU visit: parity suite + bench (no profiler)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3200 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
