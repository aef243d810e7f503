#!/bin/bash
# Dev tool: builds libact_b200 variants (compile-time knobs) into tools/bin/ for A/B timing on the GPU box.
# usage: tools/build_variants.sh name1:"-DFOO=1 -DBAR=2" name2:"..."
set -e
cd "$(dirname "$0")/../anonymous-credit-tokens_b200/csrc"
OUT=../../tools/bin
mkdir -p $OUT
[ -f cbor_host.o ] || g++ -O2 -fPIC -fvisibility=hidden -std=c++17 -c -o cbor_host.o cbor_host.cpp
pids=()
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  (
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xptxas -v $flags -c -o $OUT/$name.o act_engine.cu 2> $OUT/$name.ptxas.log
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libact_$name.so $OUT/$name.o cbor_host.o -Xlinker --exclude-libs,ALL
    rm -f $OUT/$name.o
    echo "$name [$flags]: $(grep -A3 'Compiling entry function .*spend_range' $OUT/$name.ptxas.log | grep -E 'Used' | sed 's/ptxas info    : //') $(grep -A3 'Compiling entry function .*spend_range' $OUT/$name.ptxas.log | grep -E 'spill' )"
  ) &
  pids+=($!)
  if [ ${#pids[@]} -ge 8 ]; then wait ${pids[0]}; pids=("${pids[@]:1}"); fi
done
wait
