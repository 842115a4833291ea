#!/bin/bash
# Development: one gpurun call = GPU parity suite + A/B timings of library variants + bench line + ncu launch list + one full ncu capture.
# usage (on the GPU box, from the repo root): bash tools/gpu_round.sh <tag> [variant ...]
tag=${1:-rX}; shift
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv > $out/smi.txt 2>&1
( time timeout 700 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 ) > $out/pytest.log 2>&1
for v in "$@"; do
  RTB200_LIB=$PWD/build/variants/$v/librtb200.so timeout 300 python tools/variant_check.py --noparity >> $out/variants.jsonl 2>> $out/variants.err
done
timeout 600 python bench.py --steps 5 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_bench_n1.csv python bench.py --steps 2 --warmup 1 --no-extras > $out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:persistent_kernel -c 1 -f -o $out/prof_mixed4k_strict python tools/run_once.py --reps 1 > $out/ncu_full.log 2>&1
python tools/ncu_summary.py $out/prof_mixed4k_strict.ncu-rep 40 > $out/ncu_mixed1024_4k_strict.txt 2>> $out/ncu_full.log
ls -la $out > $out/ls.txt
tail -5 $out/pytest.log; cat $out/variants.jsonl; cat $out/bench_n1.json | cut -c1-400
