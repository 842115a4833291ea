#!/bin/bash
# Development: parity suite + the textured default scene (quad kernel) bench line + the unchanged main.cpp frame loop.
tag=${1:-rX}
out=gpurun_out/$tag; mkdir -p $out
( time timeout 700 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 ) > $out/pytest.log 2>&1
timeout 300 python bench.py --workload default1080 --steps 5 --warmup 3 --no-extras --cpu-seconds 4 > $out/bench_default1080_n1.json 2> $out/bench_default1080.err
( RT_FRAMES=301 timeout 120 raytracing-opengl_b200/host/build/rt_headless; RT_FRAMES=301 RT_WIDTH=1920 RT_HEIGHT=1080 timeout 120 raytracing-opengl_b200/host/build/rt_headless ) > $out/headless_loop.log 2>&1
tail -4 $out/pytest.log; cut -c1-200 $out/bench_default1080_n1.json; grep -E "frame [0-2]:|frame loop" $out/headless_loop.log
