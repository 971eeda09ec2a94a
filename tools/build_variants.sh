#!/bin/bash
# Builds kernel variants of libtxasm.so into tianxin_b200/csrc/_build/variants/ for A/B timing on the GPU box:
#   tools/build_variants.sh name1="-DFLAG=1 ..." name2="..."      then   TXASM_LIB=.../libtxasm_name1.so python tools/quick_bench.py ...
set -e
cd "$(dirname "$0")/../tianxin_b200/csrc"
mkdir -p _build/variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  d=_build/variants/$name; mkdir -p $d
  for f in txasm_capi setup_kernels fill_atomic fill_rowgather bc_halo; do cp _build/$f.o $d/; done
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ \
      --expt-relaxed-constexpr -Xptxas -v $flags -c fill_rowtile.cu -o $d/fill_rowtile.o 2> $d/ptxas.log
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o _build/variants/libtxasm_$name.so $d/*.o -ccbin /usr/bin/g++ -ldl
  echo "$name: $(grep -A2 'k_fill_rowtileILi256ELi416ELb1ELb1' $d/ptxas.log | grep -E 'Used|spill' | tr '\n' ' ')"
done
