#!/bin/bash
# Development aid: build named variants of libb200sync (extra -D flags) into build/variants/
#   scripts/variants.sh name1 "-DFOO -DBAR=3" name2 "..." ...
# then time them on the GPU box with scripts/variant_timing.py.
set -e
cd "$(dirname "$0")/../gr4_packet_modem_b200/csrc"
mkdir -p ../../build/variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -std=c++20 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off \
       -Xptxas -v $flags -shared -o ../../build/variants/lib_$name.so api.cu multi.cpp correlator.cu peaks.cu frontend.cu symbol_filter.cu cfc.cu costas.cu stimulus.cu \
       2> ../../build/variants/build_$name.log
  echo "$name: $(grep -A2 'correlate_kernel' ../../build/variants/build_$name.log | grep -E 'spill|Used' | tr '\n' ' ')"
done
