#!/bin/bash
# frames in global memory (64-byte chunks) vs local memory: tests, A/B, DRAM traffic of one launch
set -u
O=gpurun_out/gf2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
bash scripts/r2_ab.sh "fin gf gf2 fin gf gf2" synth256 "synth1024 4K a1" "synth1024 4K a2" "accel synth1024" > $O/ab.txt 2>&1; cat $O/ab.txt
RTG_LIB_DIR=$PWD/build_variants/gf2_pt timeout 300 python scripts/tail_probe.py 4 2>&1 | tee $O/tail.txt | cut -c1-220
export RTG_LIB_DIR=$PWD/build_variants/gf2
timeout 600 ncu --set full --clock-control none -k regex:trace_kernel -c 1 -f -o $O/ncu_c4k python scripts/profile_case.py 1024 3840 2160 1 8 1 > $O/ncu_c4k.log 2>&1; echo "ncu rc=$?"
ncu -i $O/ncu_c4k.ncu-rep --page raw --csv > $O/ncu_c4k.raw.csv 2>/dev/null; rm -f $O/ncu_c4k.ncu-rep
python scripts/ncu_metrics.py $O/ncu_c4k.raw.csv | grep -E "dram|time_dur|long_score|no_instr|hit_rate|issue_active"
