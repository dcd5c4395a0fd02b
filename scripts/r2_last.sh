#!/bin/bash
# last check of the round: GPU tests, smoke, the lockstep group choice on the sweep's middle
set -u
O=gpurun_out/last5; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -2
bash scripts/r2_ab.sh "auto auto:lockstep=1 auto:lockstep=3 auto:lockstep=2" synth256 synth512 synth768 "synth1024 4K a1" "accel synth1024" > $O/ab.txt 2>&1; cat $O/ab.txt
timeout 600 python scripts/sweep.py --gpus 1 --counts 128,256,384,512,640,768,896,1024 > $O/sweep_mid.jsonl 2>&1; grep "shared (TMA bulk)\"" $O/sweep_mid.jsonl | cut -c1-200
