#!/bin/bash
# usage (under gpurun): bash tools/sanitize_r2b.sh <outdir>
out=${1:-gpurun_out/sanitize_r2b}; mkdir -p $out
for tool in memcheck racecheck synccheck; do
  log=$out/${tool}_r2b.log
  timeout 700 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_r2b.py > $log 2>&1
  echo "$tool fused lpt+wide+smaa: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) | $(grep -c checksum $log) workloads finished"
done | tee $out/summary.txt
