#!/bin/bash
set -u
O=gpurun_out/gf4; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
RTG_LIB_DIR=$PWD/build_variants/gfw timeout 600 python -m pytest tests -m gpu -x -q -k "golden or config1 or synthetic or rare_paths or strips_and or variants" > $O/pytest_gfw.txt 2>&1; echo "pytest gfw rc=$?"; tail -2 $O/pytest_gfw.txt
bash scripts/r2_ab.sh "fin gfx gfw fin gfx gfw" synth256 "synth1024 4K a2" "synth1024 8K a2" > $O/ab.txt 2>&1; cat $O/ab.txt
