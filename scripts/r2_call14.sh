#!/bin/bash
set -u
O=gpurun_out/call14; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
bash scripts/r2_ab.sh "l3 l3:lockstep=3 l3 l3:lockstep=3 l3:lockstep=1" synth256 "synth1024 4K a1" "accel synth1024" > $O/ab.txt 2>&1; cat $O/ab.txt
