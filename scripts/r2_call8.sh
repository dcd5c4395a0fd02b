#!/bin/bash
# GPU call 8: guided sweep step + 256-item first-group chunks
set -u
O=gpurun_out/call8; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
bash scripts/r2_ab.sh "cur cur2 cur2:sweep_step=4 cur2:sweep_step=1 cur cur2" synth256 "synth1024 4K a1" "synth1024 4K a2" > $O/ab.txt 2>&1; cat $O/ab.txt
for o in "" "order=2"; do
  echo "== tail cur2_pt [$o]"
  RTG_LIB_DIR=$PWD/build_variants/cur2_pt RTG_OPTS=$o timeout 300 python scripts/tail_probe.py 4 2>&1 | tee -a $O/tail_$o.txt | cut -c1-220
done
