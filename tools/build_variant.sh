#!/bin/bash
# Build a variant of libcausarray_b200.so with extra nvcc flags into tools/variants/ (kernel ablations).
#   tools/build_variant.sh NAME "-DCA_NACC=2"
set -e
name=$1; extra=$2
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=/tmp/ca_variant_$name
rm -rf $tmp && mkdir -p $tmp/causarray_b200 $tmp/include
cp -r $root/causarray_b200/csrc $tmp/causarray_b200/csrc
cp $root/include/*.h $tmp/include/
rm -f $tmp/causarray_b200/csrc/*.o
make -s -C $tmp/causarray_b200/csrc -j16 EXTRA="$extra" LIB=$root/tools/variants/lib_$name.so > $tmp/build.log 2>&1 || (tail -20 $tmp/build.log; exit 1)
echo built $root/tools/variants/lib_$name.so
