#!/bin/bash
# usage (under gpurun): bash tools/gpu_evidence.sh <tag>   -> gpurun_out/<tag>/...
# One gpurun call for the round's evidence: parity suite, one full ncu capture (-> summary + profiles/traffic.json), the bench line,
# the ncu launch list of the same command, the other BASELINE configs, the unchanged main.cpp frame loop.
tag=${1:-rX}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv > $out/smi.txt 2>&1
( time timeout 700 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 ) > $out/pytest.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:persistent_kernel -c 1 -f -o $out/prof_mixed4k_strict python tools/run_once.py --reps 1 > $out/ncu_full.log 2>&1
python tools/ncu_summary.py $out/prof_mixed4k_strict.ncu-rep 40 > $out/ncu_mixed1024_4k_strict.txt 2>> $out/ncu_full.log
python tools/ncu_traffic.py $out/prof_mixed4k_strict.ncu-rep mixed1024_4k_strict "profiles/r1b_ncu_mixed1024_4k_strict.txt" $out/traffic.json >> $out/ncu_full.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_bench_n1.csv python bench.py --steps 2 --warmup 1 --no-extras --cpu-seconds 1 > $out/launches_bench.log 2>&1
for w in default1080 spheres4k tori1080; do
  timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-extras --cpu-seconds 4 > $out/bench_${w}_n1.json 2> $out/bench_${w}.err
done
( RT_FRAMES=301 timeout 120 raytracing-opengl_b200/host/build/rt_headless; RT_FRAMES=301 RT_WIDTH=1920 RT_HEIGHT=1080 timeout 120 raytracing-opengl_b200/host/build/rt_headless ) > $out/headless_loop.log 2>&1
tail -4 $out/pytest.log; cut -c1-300 $out/bench_n1.json; tail -3 $out/headless_loop.log
