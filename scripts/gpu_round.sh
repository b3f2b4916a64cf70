#!/bin/bash
# One full GPU visit: parity suite, smoke, microbenchmarks, bench (+ reference arm), launch list, and
# `ncu --set full` captures of the dominant kernel and of the newest kernels.  Everything lands in
# gpurun_out/; copy what should be judged into profiles/.
mkdir -p gpurun_out/ncu
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python scripts/microbench.py > /dev/null 2>&1
timeout 120 python scripts/microbench_bls12_377.py > /dev/null 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; head -c 400 gpurun_out/bench.json; echo; tail -3 gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 500 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_shuffle_2p16.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch52 0 --sigma-cards 0 --bls12-377-logn 0 --msm-logn 20 > gpurun_out/ncu_bench.log 2>&1
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch52 0 --bls12-377-logn 0 --msm-logn 16"
cap() {  # name regex skip count [command]
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c $4 \
      -f -o gpurun_out/ncu/prof_$1 ${5:-$CMD} > /tmp/ncu_$1.log 2>&1
  ncu -i gpurun_out/ncu/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/prof_$1.ncu-rep --page details --csv > gpurun_out/ncu/$1_details.csv 2>/dev/null
  tail -1 /tmp/ncu_$1.log | cut -c 1-160
}
cap k_accumulate_2 "mp::k_accumulate<.int.2" 0 1
cap k_decompress "k_decompress" 1 1
cap k_lincomb "k_lincomb" 6 6
cap k_accumulate_bls12_377 "mp_bls12_377::k_accumulate" 2 1 "python scripts/bls12_377_probe.py 18"
rm -f gpurun_out/ncu/prof_k_accumulate_bls12_377.ncu-rep gpurun_out/ncu/prof_k_decompress.ncu-rep gpurun_out/ncu/prof_k_lincomb.ncu-rep
ls -la gpurun_out/ncu | tail -6
