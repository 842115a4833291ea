#!/bin/bash
# usage (under gpurun): bash tools/sanitize.sh <outdir>
# compute-sanitizer evidence for the shared-memory staging (TMA + mbarrier) and the cooperative drain's job pool:
# racecheck, memcheck and synccheck on a tiny canvas (the whole frame is drain phase) and on a 1/8-scale mixed1024 frame, both builds.
out=${1:-gpurun_out/sanitize}; mkdir -p $out
for tool in racecheck memcheck synccheck; do
  for build in strict fused; do
    for wl in "mixed1024_4k 0.02" "mixed1024_4k 0.125" "default1080 0.1"; do
      set -- $wl
      log=$out/${tool}_${build}_$1_$2.log
      timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/run_once.py --workload $1 --scale $2 --build $build --reps 1 > $log 2>&1
      echo "$tool $build $1 x$2: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
    done
  done
done | tee $out/summary.txt
