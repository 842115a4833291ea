#!/bin/bash
# Development (under gpurun): SMAA parity tests, timings (pass 2 over every pixel / over the compacted edge pixels), per-kernel launch list, one ncu --set full capture.
# Library variants built with tools/build_variant.sh (e.g. -DSMAA_EDGES_INLINE=1) can be compared by setting RTB200_LIB.
tag=${1:-tX}; o=gpurun_out/$tag; mkdir -p $o
(timeout 600 python -m pytest tests/test_smaa.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8) > $o/pytest_smaa.log 2>&1
python tools/smaa_probe.py mixed1024_4k default1080 > $o/smaa_probe.jsonl 2> $o/smaa_probe.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:smaa --csv --log-file $o/smaa_launches.csv python tools/smaa_probe.py mixed1024_4k --once > /dev/null 2> $o/smaa_ncu.err
ncu --set full --clock-control none --import-source on -k regex:smaa -o $o/smaa_full python tools/smaa_probe.py mixed1024_4k --once > /dev/null 2>> $o/smaa_ncu.err
