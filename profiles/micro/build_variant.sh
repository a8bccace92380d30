#!/bin/bash
# build a tuning variant of libsmc.so:  profiles/micro/build_variant.sh NAME -DSMC_K5_OCC=3 ...
# -> profiles/micro/variants/libsmc_NAME.so (git-ignored; use with SMC_LIB_PATH=...)
set -e
name=$1; shift
cd "$(dirname "$0")/../.."
mkdir -p profiles/micro/variants
src=particle_filters_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o profiles/micro/variants/libsmc_$name.so $src/sir_kernels.cu $src/ledh_kernels.cu $src/multinomial.cu
echo profiles/micro/variants/libsmc_$name.so
