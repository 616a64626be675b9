#!/bin/bash
# A differently tuned build of the library for A/B runs: bash tools/build_variant.sh <name> -DISX_...=.. ...
# -> build/lib_<name>.so (used through ISX_LIB_PATH, see tools/ab_variants.sh)
name=$1; shift
cd "$(dirname "$0")/../instance_stixels_b200/csrc" && mkdir -p ../../build && \
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -ftz=true -fmad=false -prec-div=true -prec-sqrt=true \
  -Xcompiler -fPIC -Xptxas -v "$@" -shared -o ../../build/lib_${name}.so \
  context.cu host_model.cu join.cu tables.cu dp.cu emit.cu group.cu road.cu ingest.cu raster.cu pool.cu
