#!/bin/bash
# Build the current source tree as build/variants/libws_<name>.so (for same-box A/B runs, tools/gpu_ab.sh).
name=$1; shift
mkdir -p build/variants
cd warpsense_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-O2 -shared \
  -ccbin /usr/bin/g++ "$@" -o ../../build/variants/libws_$name.so capi.cu update_tsdf.cu registration.cu map_ops.cu preprocess.cu voxelgrid.cu hdf5_export.cu hdf5_import.cu && echo built $name
