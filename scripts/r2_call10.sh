#!/bin/bash
set -u
O=gpurun_out/call11; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
bash scripts/r2_ab.sh "cur slimA slimB cur slimA slimB" "synth1024 4K a1" "synth1024 4K a2" > $O/ab.txt 2>&1; cat $O/ab.txt
