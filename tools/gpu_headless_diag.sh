#!/bin/bash
# development: is the run-to-run spread of the headless frame loop on the GPU (kernel ms) or on the host, and is it one stall or every frame?
B=raytracing-opengl_b200/host/build/rt_headless
for rep in 1 2 3 4 5 6 7 8 9 10; do
  RT_FRAMES=601 RT_WIDTH=1280 RT_HEIGHT=720 RT_STRICT=0 RT_SMAA=1 timeout 120 $B 2>&1 | grep "frame loop" | tr '\n' ' '; echo
done
