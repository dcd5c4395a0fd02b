#!/bin/bash
# GPU call 4: work order with the warp-cooperative pop; lockstep variants re-measured
set -u
O=gpurun_out/call5; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
bash scripts/r2_ab.sh "lpt:order=0 lpt g8 lock:order=0 lock head1 lpt" synth256 "synth1024 4K a1" "synth1024 4K a2" "accel synth1024" > $O/ab.txt 2>&1; cat $O/ab.txt
for o in "order=0" "order=1"; do
  echo "== tail lpt_pt [$o]"
  RTG_LIB_DIR=$PWD/build_variants/lpt_pt RTG_OPTS=$o timeout 300 python scripts/tail_probe.py 4 2>&1 | tee -a $O/tail_$o.txt | cut -c1-220
done
du -sh $O
