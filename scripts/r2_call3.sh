#!/bin/bash
# GPU call 3: work order (LPT buckets) + lockstep variants
set -u
O=gpurun_out/call3; mkdir -p $O
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_case.py > $O/sanitize.txt 2>&1; echo "sanitizer rc=$?"; tail -3 $O/sanitize.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.txt
for t in lock lock2 lock2w4; do
  RTG_LIB_DIR=$PWD/build_variants/$t timeout 600 python -m pytest tests -m gpu -x -q -k "golden or config1 or synthetic or rare_paths or strips_and" > $O/pytest_$t.txt 2>&1; echo "pytest $t rc=$?"; tail -2 $O/pytest_$t.txt
done
bash scripts/r2_ab.sh "lpt:order=0 lpt g8 lock lock1w4 lock2 lock2g8 lock2w4 head1 lpt" synth256 "synth1024 4K a1" "synth1024 4K a2" > $O/ab.txt 2>&1; cat $O/ab.txt
for o in "order=0" "order=1"; do
  echo "== tail lpt_pt [$o]"
  RTG_LIB_DIR=$PWD/build_variants/lpt_pt RTG_OPTS=$o timeout 300 python scripts/tail_probe.py 4 2>&1 | tee -a $O/tail_$o.txt | cut -c1-220
done
du -sh $O
