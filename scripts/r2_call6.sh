#!/bin/bash
# GPU call 6: work-order tuning (sweep step, deep threshold), phase shares under both orders
set -u
O=gpurun_out/call6; mkdir -p $O
bash scripts/r2_ab.sh "lpt:order=2 lpt lpt:sweep_step=1 lpt:sweep_step=2 lpt:sweep_step=8 lpt:deep_at=16 lpt:deep_at=24 lpt:deep_at=48 lpt:deep_at=16,sweep_step=1 lpt:order=2" synth256 "synth1024 4K a1" > $O/ab.txt 2>&1; cat $O/ab.txt
for o in "order=2" "order=1" "sweep_step=1"; do
  echo "== phases lpt_pt [$o]"
  RTG_LIB_DIR=$PWD/build_variants/lpt_pt RTG_OPTS=$o timeout 300 python scripts/quick_perf.py synth256 "synth1024 4K a1" 2>&1 | grep case | tee -a $O/phases_$o.txt | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['case'], d['ms'], 'phases(tail,setup,loop,resolve,advance)', d['phase_pct(refill+vote,setup,loop,resolve,advance)'], 'passes', d['passes_T/S2/S4/C'])"
done
for o in "order=2" "sweep_step=1" "sweep_step=1,deep_at=16" "sweep_step=2,deep_at=24"; do
  echo "== tail lpt_pt [$o]"
  RTG_LIB_DIR=$PWD/build_variants/lpt_pt RTG_OPTS=$o timeout 300 python scripts/tail_probe.py 4 2>&1 | tee -a $O/tail_$o.txt | cut -c1-220
done
