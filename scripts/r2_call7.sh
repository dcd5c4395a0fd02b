#!/bin/bash
# GPU call 7: runtime lockstep switch + work order defaults + single-loop refill
set -u
O=gpurun_out/call7; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
bash scripts/r2_ab.sh "cur cur:lockstep=2 cur:lockstep=1 cur:order=2 cur:order=2,lockstep=2 inl g8 head1 cur" synth256 "synth1024 4K a1" "synth1024 4K a2" "accel synth1024" > $O/ab.txt 2>&1; cat $O/ab.txt
for o in "order=2" ""; do
  echo "== tail cur_pt [$o]"
  RTG_LIB_DIR=$PWD/build_variants/cur_pt RTG_OPTS=$o timeout 300 python scripts/tail_probe.py 4 2>&1 | tee -a $O/tail_$o.txt | cut -c1-220
done
