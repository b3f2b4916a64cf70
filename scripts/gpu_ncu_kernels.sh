#!/bin/bash
# Round 2: the launch list of ONE sequential 2^16-card prove + verify (+ the 2^20 MSM of BASELINE config 5) and one
# `ncu --set full` capture of the largest launch of each kernel that matters (ordinals from that launch list).
# Raw / details pages are exported to CSV on the box; only the report of the dominant kernel travels back.
#     gpurun --timeout 1500 -- 'bash scripts/gpu_ncu_kernels.sh'
# Copy what should be judged from gpurun_out/ncu into profiles/r02_ncu/ (scripts/ncu_table.py makes the table).
mkdir -p gpurun_out/ncu
CMD="python scripts/one_step.py 128 512 msm"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/ncu/launches_one_step.csv $CMD > /tmp/one_step.log 2>&1
python scripts/launch_summary.py gpurun_out/ncu/launches_one_step.csv 40 > gpurun_out/ncu/launches_one_step.summary.txt
cap() {  # name regex skip
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 \
      -f -o /tmp/prof_$1 $CMD > /tmp/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu/$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details --csv > gpurun_out/ncu/$1_details.csv 2>/dev/null
  tail -1 /tmp/ncu_$1.log | cut -c 1-160
}
# skip counts: launches of that kernel BEFORE the one wanted, in the order of the launch list (first iteration of
# one_step.py is the warm-up: the prover's leaf-job launch is the first k_accumulate<2,*> of an iteration)
cap k_accumulate_2 "k_accumulate<.int.2" 2     # leaf jobs of the second (measured) prove
cap k_reduce_seg "k_reduce_seg<" 7             # leaf jobs
cap k_accumulate_1 "k_accumulate<.int.1" 12    # 2^20 MSM
cap k_fold "k_fold\\(" 16                      # 2^20 MSM: quad-cooperative Horner fold
cap k_reduce_group_quad "k_reduce_group_quad" 20
cap k_stitch "k_stitch\\(" 16
cap k_kara_points "k_kara_points\\(" 1
cap k_remask "k_remask\\(" 1
cp /tmp/prof_k_accumulate_2.ncu-rep gpurun_out/ncu/ 2>/dev/null; ls -la gpurun_out/ncu | head -40
