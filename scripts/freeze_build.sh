#!/bin/bash
# Development aid: copy the built libraries + host program into build_variants/<tag>/ so that a queued
# gpurun call (which snapshots the repo only when a GPU slot frees up) keeps using THAT build while the
# sources move on.  Use on the box with:  export RTG_LIB_DIR=$PWD/build_variants/<tag>
set -e
tag=${1:?usage: freeze_build.sh <tag>}
d=build_variants/$tag; mkdir -p $d
cp raytracer-gamma_b200/librt_cuda.so raytracer-gamma_b200/librt_scene.so raytracer-gamma_b200/librt_cuda_multi.so raytracer-gamma_b200/rt_gamma $d/
echo "frozen into $d"
