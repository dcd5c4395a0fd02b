#!/bin/bash
# First run of the v2 kernel (slot records in shared memory): GPU tests, quick perf (auto / forced local), tail probe
set -u
O=gpurun_out/r2_v2a; mkdir -p $O
export RTG_LIB_DIR=$PWD/build_variants/v2a
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -25 $O/pytest_gpu.txt
echo "== auto"; timeout 300 python scripts/quick_perf.py synth 2>&1 | tee $O/quick_auto.txt | grep -E "case|rror" | cut -c1-330
echo "== slot_mode=2 (local)"; RTG_OPTS=slot_mode=2 timeout 300 python scripts/quick_perf.py synth256 "synth1024 4K a1" 2>&1 | tee $O/quick_local.txt | grep -E "case|rror" | cut -c1-330
timeout 300 python scripts/tail_probe.py 4 > $O/tail_probe.txt 2>&1; cat $O/tail_probe.txt
