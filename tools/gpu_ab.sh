#!/bin/bash
# Development (under gpurun): GPU parity suite + A/B timings of the library variants built by tools/build_variant.sh.
# usage: bash tools/gpu_ab.sh <tag> <variant> [variant ...]   -> gpurun_out/<tag>/{pytest.log,variants.jsonl}
tag=${1:-rX}; shift
out=gpurun_out/$tag; mkdir -p $out
( time timeout 700 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 ) > $out/pytest.log 2>&1
for v in "$@"; do
  RTB200_LIB=$PWD/build/variants/$v/librtb200.so timeout 300 python tests/dev/variant_check.py --noparity >> $out/variants.jsonl 2>> $out/variants.err
done
tail -5 $out/pytest.log; cat $out/variants.jsonl
