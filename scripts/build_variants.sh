#!/bin/bash
# Build variants of one translation unit HERE (nvcc cross-compiles) and link each into its own library under
# dnascent_b200/lib/variants/, so that one gpurun call can time them all without compiling on the GPU box.
# usage: scripts/build_variants.sh <unit.cu> <tag>="<nvcc flags>" [<tag>="<flags>" ...]
set -e
cd "$(dirname "$0")/.."
unit=$1; shift
base=$(basename "$unit" .cu)
make -s -C dnascent_b200/csrc > /dev/null
mkdir -p dnascent_b200/lib/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
for spec in "$@"; do
  tag=${spec%%=*}; flags=${spec#*=}
  obj=dnascent_b200/lib/variants/${base}_${tag}.o
  nvcc $ARCH -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-fopenmp,-fvisibility=hidden,-O2 \
       -Xptxas -v $flags -c dnascent_b200/csrc/$unit -o $obj 2> dnascent_b200/lib/variants/${base}_${tag}.ptxas.log
  others=$(ls dnascent_b200/lib/obj/*.o | grep -v "/${base}.o")
  nvcc $ARCH -shared -Xcompiler -fopenmp -o dnascent_b200/lib/variants/libdnascent_b200_${tag}.so $obj $others -lcudart_static -lgomp -ldl -lrt -lpthread
  echo "$tag: $flags: $(grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' dnascent_b200/lib/variants/${base}_${tag}.ptxas.log | tr '\n' ' ' | cut -c1-200)"
done
