#!/bin/sh
# compute-sanitizer evidence: memcheck + racecheck + synccheck over small end-to-end cases and the GEMM matrix.
# Output: gpurun_out/sanitize_*.log (summaries are copied into profiles/ by hand).
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  # synccheck: the default barrier table overflows on the 32 mbarriers per CTA of the Jacobi kernel
  compute-sanitizer --tool $tool $( [ $tool = synccheck ] && echo --num-cuda-barriers 65536 ) python tools/sanitize_case.py > $OUT/sanitize_case_$tool.log 2>&1
  echo "case $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|SYNCCHECK|errors' $OUT/sanitize_case_$tool.log | tail -1)"
done
for persist in 0 2; do
  for tool in memcheck racecheck; do
    FFGP_CHECK_SMALL=1 FFGP_PERSIST=$persist compute-sanitizer --tool $tool python tools/check_gemm.py > $OUT/sanitize_gemm_p${persist}_$tool.log 2>&1
    echo "gemm persist=$persist $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitize_gemm_p${persist}_$tool.log | tail -1) / $(grep -c 'max rel err' $OUT/sanitize_gemm_p${persist}_$tool.log) cases, $(grep -c MISMATCH $OUT/sanitize_gemm_p${persist}_$tool.log) mismatches"
  done
done
