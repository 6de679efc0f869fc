#!/bin/sh
# memcheck + racecheck + synccheck over tools/sanitize_case.py only (the GEMM matrix is covered by tools/sanitize.sh)
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  compute-sanitizer --tool $tool $( [ $tool = synccheck ] && echo --num-cuda-barriers 65536 ) python tools/sanitize_case.py > $OUT/sanitize_case_$tool.log 2>&1
  echo "case $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitize_case_$tool.log | tail -1)"
  grep -E "^(batched|dense|hogp|eigh|match)" $OUT/sanitize_case_$tool.log | tr '\n' ';'; echo
done
