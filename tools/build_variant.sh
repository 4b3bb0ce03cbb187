#!/bin/bash
# usage: tools/build_variant.sh NAME [-DFLAG ...]  ->  tools/variants/libafr_NAME.so (tuning builds, AFR_LIB_PATH)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/variants
C=aliasfree-diffusion-models-pytorch_b200/csrc
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -cudart shared --threads 0 "$@" \
  -o tools/variants/libafr_$name.so $C/afr_api.cu $C/afr_generic.cu $C/afr_n3.cu $C/afr_stripn.cu $C/afr_rotate.cu $C/afr_small.cu $C/afr_actdown.cu $C/afr_norm.cu $C/afr_nhwc_resample.cu
