#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_accumulate<.int.2>" -s 2 -c 2 \
    -f -o gpurun_out/prof_accumulate python bench.py --steps 1 --warmup 1 --no-cpu-baseline --msm-logn 16 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log | cut -c 1-200
ls -la gpurun_out/*.ncu-rep
