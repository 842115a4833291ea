#!/bin/bash
# Development (under gpurun): SMAA parity tests, A/B of the search-speculation variants and block shapes, ncu captures.
tag=${1:-tX}; o=gpurun_out/$tag; mkdir -p $o
(timeout 600 python -m pytest tests/test_smaa.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8) > $o/pytest_smaa.log 2>&1
python tools/smaa_probe.py mixed1024_4k default1080 > $o/smaa_probe.jsonl 2> $o/smaa_probe.err
for v in smaa_spec1 smaa_spec2 smaa_spec8; do echo "# $v" >> $o/smaa_probe.jsonl; RTB200_LIB=$PWD/build/variants/$v/librtb200.so python tools/smaa_probe.py mixed1024_4k >> $o/smaa_probe.jsonl 2>> $o/smaa_probe.err; done
for b in 32x4 32x16 64x4 128x2 256x1 128x1; do echo "# block $b" >> $o/smaa_probe.jsonl; RTB_SMAA_BLOCK=$b python tools/smaa_probe.py mixed1024_4k >> $o/smaa_probe.jsonl 2>> $o/smaa_probe.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:smaa --csv --log-file $o/smaa_launches.csv python tools/smaa_probe.py mixed1024_4k --once > /dev/null 2> $o/smaa_ncu.err
ncu --set full --clock-control none --import-source on -k regex:smaa -o $o/smaa_full python tools/smaa_probe.py mixed1024_4k --once > /dev/null 2>> $o/smaa_ncu.err
