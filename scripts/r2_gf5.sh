#!/bin/bash
set -u
O=gpurun_out/gf5; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
bash scripts/r2_ab.sh "cur3 hoist fin cur3 hoist" synth256 synth512 "synth1024 4K a2" "synth1024 8K a2" "accel synth1024" > $O/ab.txt 2>&1; cat $O/ab.txt
