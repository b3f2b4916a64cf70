#!/bin/bash
# GPU visit: sanitizer on small proofs (both diagonal forms), parity suite, smoke, bench, launch list
mkdir -p gpurun_out
lscpu | grep -i "model name\|^CPU(s)\|flags" | cut -c 1-400 > gpurun_out/lscpu.txt
for k in 1 0; do
MP_SMALL_DECK_MAX=0 MP_DIAG_KARATSUBA=$k timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python scripts/repro.py 5 3 2>&1 | grep -v "Host Frame\|^=========         in " | head -30
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
MP_DIAG_KARATSUBA=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch52 0 --pipeline-decks 0 --msm-logn 16 > gpurun_out/bench_schoolbook.json 2> gpurun_out/bench_schoolbook.err; head -c 600 gpurun_out/bench_schoolbook.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_shuffle_2p16.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch52 0 --pipeline-decks 0 --msm-logn 16 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c 1-300
