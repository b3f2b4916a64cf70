#!/bin/bash
# compute-sanitizer pass over the hot path plus the variant matrix of the two kernels that
# misbehaved on the 12-limb build in round 1.  Run on the GPU box:
#     gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
# Output: gpurun_out/sanitize/*.log and a one-line-per-run summary in gpurun_out/sanitize/SUMMARY.txt
OUT=gpurun_out/sanitize
mkdir -p $OUT
SUM=$OUT/SUMMARY.txt
: > $SUM
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # label, env assignment, repro name
  env $2 timeout 300 python scripts/repro_12limb.py $3 > $OUT/$1.log 2>&1
  echo "$1 ($2 $3): $(tail -1 $OUT/$1.log)" >> $SUM
}
# --- variant matrix (no sanitizer): which formulation reproduces, which does not
run win_default       X=0              win
run win_mode0_orig    MP_WIN_BLOCK=1   win
run win_mode1_syncwarp MP_WIN_BLOCK=2  win
run win_mode2_uniform MP_WIN_BLOCK=3   win
run win_mode3_memxch  MP_WIN_BLOCK=4   win
run table_default     X=0              table
run table_mode0_orig  MP_TABLE_TRICK=1 table
run table_mode1_gmem  MP_TABLE_TRICK=2 table
run table_mode2_inline MP_TABLE_TRICK=3 table
# --- sanitizer tools on the two original formulations
for tool in synccheck racecheck memcheck initcheck; do
  MP_WIN_BLOCK=1 timeout 400 $CS --tool $tool --print-limit 20 python scripts/repro_12limb.py win > $OUT/win_orig_$tool.log 2>&1
  echo "win_orig $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/win_orig_$tool.log | tail -1) / $(grep -E '^(PASS|FAIL)' $OUT/win_orig_$tool.log | tail -1)" >> $SUM
  MP_TABLE_TRICK=1 timeout 400 $CS --tool $tool --print-limit 20 python scripts/repro_12limb.py table > $OUT/table_orig_$tool.log 2>&1
  echo "table_orig $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/table_orig_$tool.log | tail -1) / $(grep -E '^(PASS|FAIL)' $OUT/table_orig_$tool.log | tail -1)" >> $SUM
done
# --- the product path of both curves under synccheck + memcheck: one small shuffle prove/verify and the MSM tests
for tool in synccheck memcheck; do
  timeout 500 $CS --tool $tool --print-limit 20 python -m pytest tests/test_gpu_shuffle.py -m gpu -x -q -k "golden and host_scalars or negative_case" > $OUT/shuffle_$tool.log 2>&1
  echo "shuffle tests $tool: $(grep -E 'ERROR SUMMARY' $OUT/shuffle_$tool.log | tail -1) / $(tail -1 $OUT/shuffle_$tool.log)" >> $SUM
done
cat $SUM
