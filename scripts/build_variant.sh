#!/bin/bash
# build_variant.sh <name> [extra nvcc flags...]: tuning builds of the engine into variants/lib_<name>.so (selected with BN254_B200_LIB)
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC "$@" -o variants/lib_$name.so bn254_b200/csrc/bn254_b200.cu
