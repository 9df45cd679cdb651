#!/bin/bash
# development aid: build a variant of the library with extra -D flags for ONE translation unit
# usage: tools/build_variant.sh <name> <tu (e.g. gj_islands_ga)> <flags...>   -> greyjack-solver-rust_b200/build/variants/lib_<name>.so
set -e
cd "$(dirname "$0")/../greyjack-solver-rust_b200"
name=$1; tu=$2; shift 2
mkdir -p build/variants build/var_$name
NV="/usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --expt-relaxed-constexpr"
$NV "$@" -c -o build/var_$name/$tu.o csrc/$tu.cu 2> build/var_$name/$tu.log
objs=$(ls build/*.o | grep -v "/$tu.o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/lib_$name.so $objs build/var_$name/$tu.o
grep -A2 "Function properties" build/var_$name/$tu.log | grep -B1 -A1 "spill" | grep -v "^--" | paste - - - | grep -i "planned_vrp\|k_vrp_chains" | sed 's/ptxas info    ://g' | cut -c1-400
