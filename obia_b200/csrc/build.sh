#!/bin/bash
# Build libobia_b200.so (sm_100a only) in-tree.  Usage: obia_b200/csrc/build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
mkdir -p ../_lib
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
     -Xcompiler -fPIC -shared "$@" \
     core.cu preprocess.cu kmeans.cu slic.cu slic_fast.cu connectivity.cu zonal.cu texture.cu legacy_rng.cu rasterize.cu quickshift.cu tiled.cu \
     -o ../_lib/libobia_b200.so
