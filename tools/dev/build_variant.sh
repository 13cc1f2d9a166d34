#!/bin/bash
# build a variant of libosmr_b200.so with extra -D flags into tools/dev/variants/<name>.so (experiments; git-ignored)
# usage: tools/dev/build_variant.sh name -DOSMR_RASTER_MIN_BLOCKS=24 ...
set -e
name=$1; shift
cd "$(dirname "$0")/../.."
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo -Xcompiler -fPIC -shared -cudart static \
  -I include -I osm_renderer_b200/csrc "$@" -o tools/dev/variants/$name.so osm_renderer_b200/csrc/osmr_capi.cu
echo built tools/dev/variants/$name.so
