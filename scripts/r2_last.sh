#!/bin/bash
# last check of the round: GPU tests and smoke on the final build
set -u
O=gpurun_out/last7; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -2
