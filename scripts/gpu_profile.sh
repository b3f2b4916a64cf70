#!/bin/bash
# parity + bench + launch list + one full ncu capture of the dominant kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python scripts/microbench.py > /dev/null 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 700 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_shuffle_2p16.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch52 0 --pipeline-decks 0 --msm-logn 20 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_accumulate<.int.2" -s 1 -c 2 \
    -f -o gpurun_out/prof_accumulate python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch52 0 --pipeline-decks 0 --msm-logn 16 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c 1-200
ls -la gpurun_out/*.ncu-rep
