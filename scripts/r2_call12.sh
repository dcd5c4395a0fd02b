#!/bin/bash
set -u
O=gpurun_out/call12; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
bash scripts/r2_ab.sh "slimA unw head1 slimA unw" synth256 "synth1024 4K a1" "synth1024 4K a2" "accel synth1024" > $O/ab.txt 2>&1; cat $O/ab.txt
RTG_LIB_DIR=$PWD/build_variants/unw_pt timeout 300 python scripts/tail_probe.py 4 2>&1 | tee $O/tail.txt | cut -c1-220
