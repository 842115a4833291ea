#!/bin/bash
# usage (under gpurun): bash tools/gpu_headless.sh <outfile> — the reference's unchanged main.cpp frame loop (animated default scene) on librtb200.so
out=${1:-gpurun_out/headless_loop.log}
B=raytracing-opengl_b200/host/build/rt_headless
RT_FRAMES=200 timeout 60 $B > /dev/null 2>&1      # warm the box (clocks, page cache)
for cfg in "1280 720" "1920 1080" "3840 2160"; do
  set -- $cfg
  for strict in 1 0; do for smaa in 1 0; do
    echo "== ${1}x${2} RT_STRICT=$strict RT_SMAA=$smaa"
    RT_FRAMES=601 RT_WIDTH=$1 RT_HEIGHT=$2 RT_STRICT=$strict RT_SMAA=$smaa timeout 120 $B 2>&1 | grep -E "frame loop"
  done; done
done > $out 2>&1
