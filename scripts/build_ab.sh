#!/bin/bash
# DEV: build variants of the product library into build/ab/ (they travel to the GPU box with the snapshot)
#   scripts/build_ab.sh NAME "-DFOO=1 -DBAR=2"
cd "$(dirname "$0")/.." || exit 1
C=criteria3d_b200/csrc
mkdir -p build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared $2 \
    -I include -I $C -o build/ab/libsf3d_$1.so $C/sf3d_kernels.cu $C/sf3d_engine.cpp $C/sf3d_capi.cpp $C/sf3d_shim.cpp $C/sf3d_gis.cu && echo "built build/ab/libsf3d_$1.so"
