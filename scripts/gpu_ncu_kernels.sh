#!/bin/bash
# One `ncu --set full` capture of the largest launch of every kernel of a 2^16-card prove+verify
# (ordinals taken from the launch list of the same command, gpurun_out/launches_shuffle_2p16.csv).
# Raw pages are exported to CSV on the box; only the report of the dominant kernel travels back.
mkdir -p gpurun_out/ncu
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --batch52 0 --msm-logn 16"
cap() {  # name regex skip
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 \
      -f -o /tmp/prof_$1 $CMD > /tmp/ncu_$1.log 2>&1
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > gpurun_out/ncu/$1_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$1.ncu-rep --page details --csv > gpurun_out/ncu/$1_details.csv 2>/dev/null
  tail -1 /tmp/ncu_$1.log | cut -c 1-160
}
cap k_accumulate_2 "k_accumulate<.int.2" 0   # 9.770 ms in the launch list
cap k_reduce_seg "k_reduce_seg\\(" 3   # 4.677 ms in the launch list
cap k_kara_points "k_kara_points\\(" 0   # 1.760 ms in the launch list
cap k_batch_to_affine "k_batch_to_affine\\(" 0   # 0.754 ms in the launch list
cap k_fold "k_fold\\(" 3   # 0.999 ms in the launch list
cap k_stitch "k_stitch\\(" 7   # 0.741 ms in the launch list
cap k_accumulate_1 "k_accumulate<.int.1" 3   # 1.056 ms in the launch list
cap k_reduce_win "k_reduce_win\\(" 2   # 0.644 ms in the launch list
cap k_remask "k_remask\\(" 0   # 1.021 ms in the launch list
cap k_scatter "k_scatter\\(" 3   # 0.264 ms in the launch list
cap k_count "k_count\\(" 3   # 0.113 ms in the launch list
cap k_kara_combine "k_kara_combine\\(" 0   # 0.154 ms in the launch list
cp /tmp/prof_k_accumulate_2.ncu-rep gpurun_out/ncu/ 2>/dev/null; ls -la gpurun_out/ncu | head -50
