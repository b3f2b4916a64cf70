#!/bin/bash
# compute-sanitizer pass over the hot path plus the root-cause harnesses of the two kernels that
# misbehaved on the 12-limb build in round 1 (profiles/r02_rootcause/).  Run on the GPU box:
#     gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
# Output: gpurun_out/sanitize/*.log and a one-line-per-run summary in gpurun_out/sanitize/SUMMARY.txt
OUT=gpurun_out/sanitize
mkdir -p $OUT
SUM=$OUT/SUMMARY.txt
: > $SUM
CS=/usr/local/cuda/bin/compute-sanitizer
# --- root-cause harnesses (scripts/repro/*.cu, build lines in their headers): round 1's kernels over by-pointer /
#     by-reference helpers (the defect) and over by-value helpers (the fix), plus the kernels msm.cu ships now
for b in table_stark table_377 win_stark win_377; do
  timeout 120 scripts/repro/repro_$b > $OUT/repro_$b.log 2>&1
  echo "repro_$b: $(tail -1 $OUT/repro_$b.log)" >> $SUM
done
# --- sanitizer tools on the 12-limb harnesses (by-pointer forms included: the defect is not a race / sync / memory error)
for tool in synccheck racecheck memcheck initcheck; do
  for b in table_377 win_377; do
    timeout 400 $CS --tool $tool --print-limit 20 scripts/repro/repro_$b > $OUT/${b}_$tool.log 2>&1
    echo "repro_$b $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/${b}_$tool.log | tail -1)" >> $SUM
  done
done
# --- the 12-limb product path end to end (known answers)
for w in win table; do
  timeout 300 python scripts/repro_12limb.py $w > $OUT/product_$w.log 2>&1
  echo "product path $w: $(tail -1 $OUT/product_$w.log)" >> $SUM
done
# --- the product path of both curves under synccheck + memcheck: one small shuffle prove/verify and the MSM tests
for tool in synccheck memcheck; do
  timeout 500 $CS --tool $tool --print-limit 20 python -m pytest tests/test_gpu_shuffle.py -m gpu -x -q -k "golden and host_scalars or negative_case" > $OUT/shuffle_$tool.log 2>&1
  echo "shuffle tests $tool: $(grep -E 'ERROR SUMMARY' $OUT/shuffle_$tool.log | tail -1) / $(tail -1 $OUT/shuffle_$tool.log)" >> $SUM
done
# --- the MSM pipeline of both curves (quad-cooperative fold / combine kernels use quad-masked shuffles) and the sigma kernel
for tool in synccheck racecheck; do
  timeout 500 $CS --tool $tool --print-limit 20 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_bls12_377.py -m gpu -x -q -k "msm_small or msm_golden or msm_jobs or ct_msm" > $OUT/msm_$tool.log 2>&1
  echo "msm tests (both curves) $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/msm_$tool.log | tail -1) / $(tail -1 $OUT/msm_$tool.log)" >> $SUM
done
timeout 500 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_sigma.py -m gpu -x -q -k "golden or vs_oracle or malformed" > $OUT/sigma_memcheck.log 2>&1
echo "sigma tests memcheck: $(grep -E 'ERROR SUMMARY' $OUT/sigma_memcheck.log | tail -1) / $(grep -E 'passed|failed' $OUT/sigma_memcheck.log | tail -1)" >> $SUM
# --- second curve: sigma protocols, wire-format deserialisation (square roots + G1 membership), chunked batch prover
for tool in memcheck synccheck; do
  timeout 500 $CS --tool $tool --print-limit 20 python -m pytest tests/test_gpu_bls12_377.py -m gpu -x -q -k "sigma or wire or golden" > $OUT/bls_sigma_wire_$tool.log 2>&1
  echo "BLS12-377 sigma / wire / prover tests $tool: $(grep -E 'ERROR SUMMARY' $OUT/bls_sigma_wire_$tool.log | tail -1) / $(grep -E 'passed|failed' $OUT/bls_sigma_wire_$tool.log | tail -1)" >> $SUM
done
timeout 500 $CS --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_shuffle.py -m gpu -x -q -k "lanes or shared_statement or split_over" > $OUT/batch_memcheck.log 2>&1
echo "batch drivers (chunked, shared hashes) memcheck: $(grep -E 'ERROR SUMMARY' $OUT/batch_memcheck.log | tail -1) / $(grep -E 'passed|failed' $OUT/batch_memcheck.log | tail -1)" >> $SUM
cat $SUM
