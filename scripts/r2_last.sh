#!/bin/bash
# last check of the round: GPU tests, smoke, and the final build against the previous one
set -u
O=gpurun_out/last3; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -2
bash scripts/r2_ab.sh "cur4 fin2 cur4 fin2" synth256 "synth1024 8K a2" > $O/ab.txt 2>&1; cat $O/ab.txt
RTG_LIB_DIR=$PWD/build_variants/fin2 timeout 300 python scripts/tail_probe.py 4 2>&1 | cut -c1-150
