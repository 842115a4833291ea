#!/bin/bash
# Development: GPU parity suite + A/B timings of library variants + one full ncu capture (no bench line).
tag=${1:-rX}; shift
out=gpurun_out/$tag; mkdir -p $out
( time timeout 700 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 ) > $out/pytest.log 2>&1
for v in "$@"; do
  RTB200_LIB=$PWD/build/variants/$v/librtb200.so timeout 300 python tools/variant_check.py --noparity >> $out/variants.jsonl 2>> $out/variants.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:persistent_kernel -c 1 -f -o $out/prof_mixed4k_strict python tools/run_once.py --reps 1 > $out/ncu_full.log 2>&1
python tools/ncu_summary.py $out/prof_mixed4k_strict.ncu-rep 40 > $out/ncu_mixed1024_4k_strict.txt 2>> $out/ncu_full.log
tail -5 $out/pytest.log; cat $out/variants.jsonl
