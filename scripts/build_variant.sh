#!/bin/bash
# build_variant.sh NAME FILE.cu "-DFLAGS": a copy of the library with one source recompiled with extra flags
# (obia_b200/_lib/variants/lib_NAME.so; select it with OBIA_B200_LIB).  Needs an up-to-date main build.
set -e
cd "$(dirname "$0")/../obia_b200/csrc"
name=$1; src=$2; flags=$3
mkdir -p ../_lib/variants /tmp/obia_var
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $flags -c $src -o /tmp/obia_var/$name.o
objs=$(ls ../_lib/obj/*.o | grep -v "/${src%.cu}\.")
nvcc -gencode arch=compute_100a,code=sm_100a -shared $objs /tmp/obia_var/$name.o -o ../_lib/variants/lib_$name.so
