#!/bin/bash
# Build a variant of the CUDA library for A/B runs: bash profiles/build_variant.sh NAME -DFLAG=1 ...
# -> arboris-python_b200/arboris_b200/lib/variants/libarboris_b200_NAME.so  (select with ARB_B200_LIB)
NAME=$1; shift
D=arboris-python_b200/arboris_b200/lib/variants
mkdir -p $D
cd arboris-python_b200/csrc && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -shared -diag-suppress 128 "$@" -o ../arboris_b200/lib/variants/libarboris_b200_$NAME.so arb_api.cu arb_fused.cu
