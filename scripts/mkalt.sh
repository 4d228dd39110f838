#!/bin/bash
# Build an A/B variant of the library into build/alt_<name>/ : scripts/mkalt.sh <name> -DFOO=1 ...
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/alt_$name
make -s -C multicore-hw2_b200/csrc -j8 OUT=$PWD/build/alt_$name OBJ=$PWD/build/alt_$name/obj EXTRA="$*" 2>&1 | grep -v "^nvcc\|^mkdir" | tail -3
[ -x build/alt_$name/nn_bench ] && echo "built $name"
