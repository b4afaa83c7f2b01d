#!/bin/bash
# Builds libjaicov_b200.so (sm_100a only) in-tree.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC ${JAICOV_NVCC_EXTRA}"
mkdir -p _obj
pids=()
for f in prep_kernels assembly dense_sigma dense_kernels ozaki stage_kernels structured propagate dlt dist api; do
  if [ ! -f _obj/$f.o ] || [ csrc/$f.cu -nt _obj/$f.o ] || [ -n "$(find csrc include ../include -newer _obj/$f.o \( -name '*.h' -o -name '*.hpp' -o -name '*.cuh' \) 2>/dev/null | head -1)" ]; then
    rm -f _obj/$f.o          # a failed compile must not leave a stale object for the link step
    $NVCC $FLAGS -c csrc/$f.cu -o _obj/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done   # set -e: the first failed compile stops the build
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libjaicov_b200.so _obj/prep_kernels.o _obj/assembly.o _obj/dense_sigma.o _obj/dense_kernels.o _obj/ozaki.o _obj/stage_kernels.o _obj/structured.o _obj/propagate.o _obj/dlt.o _obj/dist.o _obj/api.o -ldl -lcudart
# native host mirror of the reference's Java API (header-only jaicov_host.hpp) behind flat C entry points
g++ -std=c++17 -O2 -Wall -Wextra -fPIC -shared -o libjaicov_host.so host/host_capi.cpp -L. -ljaicov_b200 -Wl,-rpath,'$ORIGIN'
echo built $(pwd)/libjaicov_b200.so $(pwd)/libjaicov_host.so
