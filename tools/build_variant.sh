#!/bin/bash
# Kernel A/B experiments: build liblq.so with extra -D knobs into build/variants/liblq_<name>.so (the in-tree library
# is untouched); run with LQ_LIBRARY=build/variants/liblq_<name>.so.
#   tools/build_variant.sh u8mb3 "-DLQ_COL_U1=8 -DLQ_COL_FWD_MINBLK=3"
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d /tmp/lqvar_XXXX)
mkdir -p $tmp/pkg/csrc $tmp/include $root/build/variants
cp $root/learned_quantization_b200/csrc/*.cu $root/learned_quantization_b200/csrc/*.cuh $root/learned_quantization_b200/csrc/Makefile $tmp/pkg/csrc/
cp $root/include/*.h $tmp/include/
make -s -C $tmp/pkg/csrc -j8 EXTRA="$*" > $tmp/build.log 2>&1 || { tail -20 $tmp/build.log; exit 1; }
cp $tmp/pkg/csrc/liblq.so $root/build/variants/liblq_$name.so
rm -rf $tmp
echo built build/variants/liblq_$name.so "($*)"
