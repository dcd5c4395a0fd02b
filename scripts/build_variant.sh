#!/bin/bash
# Development aid: compile librt_cuda.so with extra nvcc flags into build_variants/<tag>/ (next to copies of the
# other libraries), for A/B runs on the GPU box:  build_variant.sh <tag> [nvcc flags...]   [SRC=<checkout>]
set -e
tag=${1:?usage: build_variant.sh <tag> [flags]}; shift
src=${SRC:-$PWD}
d=build_variants/$tag; mkdir -p $d
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -Xcompiler -fPIC,-ffp-contract=off -std=c++17 \
  -diag-suppress 186 -shared -I$src/include -I$src/raytracer-gamma_b200/csrc "$@" -o $d/librt_cuda.so $src/raytracer-gamma_b200/csrc/rt_shim.cu
cp raytracer-gamma_b200/librt_scene.so raytracer-gamma_b200/librt_cuda_multi.so raytracer-gamma_b200/rt_gamma $d/
echo "built $d ($*)"
