#!/bin/bash
# Full visit: parity suite, smoke, microbench, bench, reference arm, launch list, ncu full capture of the dominant kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python scripts/microbench.py > /dev/null 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; head -c 500 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 700 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_shuffle_2p16.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch52 0 --pipeline-decks 0 --msm-logn 20 > gpurun_out/ncu_bench.log 2>&1
mkdir -p gpurun_out/ncu
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch52 0 --pipeline-decks 0 --msm-logn 16"
for spec in "k_accumulate_2 k_accumulate<.int.2 0" "k_reduce_seg k_reduce_seg\\( 3"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/ncu/prof_$1 $CMD > /tmp/ncu_$1.log 2>&1
  ncu -i gpurun_out/ncu/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/prof_$1.ncu-rep --page details --csv > gpurun_out/ncu/$1_details.csv 2>/dev/null
  tail -1 /tmp/ncu_$1.log | cut -c 1-160
done
ls -la gpurun_out/ncu | tail -8
