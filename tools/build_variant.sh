#!/bin/bash
# Tuning build of the product library with extra -D flags: bash tools/build_variant.sh <name> [-DM3D_...=...]...
# -> build_variants/libm3dreg_<name>.so (use with M3DREG_LIB_PATH; never loaded by default).
name=$1; shift
mkdir -p build_variants
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xcompiler -ffp-contract=off -shared "$@" \
    -o build_variants/libm3dreg_${name}.so mandala-mapping_b200/csrc/m3dreg.cu
