#!/bin/bash
# Development helper: build librtb200 with extra -D flags into build/variants/<name>/librtb200.so
# usage: tools/build_variant.sh <name> [extra nvcc flags...]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
name=$1; shift
dst=$ROOT/build/variants/$name
mkdir -p $dst/pkg/csrc $dst/include
cp $ROOT/raytracing-opengl_b200/csrc/*.cu $ROOT/raytracing-opengl_b200/csrc/*.cuh $ROOT/raytracing-opengl_b200/csrc/*.h $ROOT/raytracing-opengl_b200/csrc/Makefile $dst/pkg/csrc/
cp $ROOT/include/*.h $dst/include/
make -C $dst/pkg/csrc -j4 EXTRA="$*" >/dev/null 2>$dst/make.err || { cat $dst/make.err; exit 1; }
cp $dst/pkg/librtb200.so $dst/librtb200.so
grep -A2 "persistent_kernelILb0ELi640" $dst/pkg/csrc/ptxas_strict.log | grep -E "Used|spill" | tr '\n' ' '; echo " <- $name strict persistent"
