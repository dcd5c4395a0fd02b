#!/bin/bash
set -u
O=gpurun_out/call13; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
bash scripts/r2_ab.sh "unw res unw res" synth256 "synth1024 4K a1" "synth1024 4K a2" "accel synth1024" > $O/ab.txt 2>&1; cat $O/ab.txt
RTG_LIB_DIR=$PWD/build_variants/res_pt timeout 300 python scripts/tail_probe.py 4 2>&1 | tee $O/tail.txt | cut -c1-220
RTG_LIB_DIR=$PWD/build_variants/res_pt timeout 300 python scripts/quick_perf.py synth256 "synth1024 4K a2" 2>&1 | grep case | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['case'], d['ms'], 'phases(tail,setup,loop,resolve,advance)', d['phase_pct(refill+vote,setup,loop,resolve,advance)'], 'passes', d['passes_T/S2/S4/C'])"
