#!/bin/bash
# GPU visit: host hash probe, parity suite, smoke, bench, launch list, per-kernel ncu captures
mkdir -p gpurun_out
lscpu | grep -i "model name\|^CPU(s)\|MHz" > gpurun_out/lscpu.txt; grep -m1 flags /proc/cpuinfo | tr ' ' '\n' | grep -i "avx512f\|avx512vl\|avx2" | tr '\n' ' ' >> gpurun_out/lscpu.txt
(cd scripts/hashbench && g++ -O2 -std=c++17 -o /tmp/hbp hb_product.cpp && /tmp/hbp | tail -2 && MP_BLAKE2S_NO_AVX512=1 /tmp/hbp | tail -1) 2>&1 | tee gpurun_out/hashbench.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; head -c 1500 gpurun_out/bench.json; echo; tail -5 gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_shuffle_2p16.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch52 0 --pipeline-decks 0 --msm-logn 16 > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c 1-300
bash scripts/gpu_ncu_kernels.sh
