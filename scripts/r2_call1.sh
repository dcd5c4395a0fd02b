#!/bin/bash
# GPU call 1 of this session: tests on the in-tree build, A/B of frozen builds, phase timing, ncu full captures
set -u
O=gpurun_out/call1; mkdir -p $O
nvidia-smi -L > $O/box.txt; nvidia-smi --query-gpu=clocks.max.sm,clocks.sm,power.limit --format=csv >> $O/box.txt
{ echo "== OpenCL ICD probe"; ls -la /etc/OpenCL/vendors 2>&1; ls /usr/lib/x86_64-linux-gnu 2>/dev/null | grep -i -E "opencl|nvidia-opencl" ; ldconfig -p | grep -i opencl; which clinfo; } > $O/opencl_probe.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.txt
bash scripts/r2_ab.sh "head1 v2e c0s0 head1 c0s0" synth256 "synth1024 4K a1" "synth1024 4K a2" "accel synth1024" > $O/ab.txt 2>&1; cat $O/ab.txt
for t in c0s0_pt head1_pt; do
  echo "== phases $t"
  RTG_LIB_DIR=$PWD/build_variants/$t timeout 300 python scripts/quick_perf.py synth256 "synth1024 4K a1" "synth1024 4K a2" 2>&1 | grep case | tee -a $O/phases_$t.txt | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['case'], d['ms'], 'phases', d['phase_pct(refill+vote,setup,loop,resolve,advance)'], 'passes', d['passes_T/S2/S4/C'], 'served', d['served_T/S/C'])"
  RTG_LIB_DIR=$PWD/build_variants/$t timeout 300 python scripts/tail_probe.py 4 2>&1 | tee $O/tail_$t.txt | cut -c1-220
done
for t in c0s0 head1; do
  export RTG_LIB_DIR=$PWD/build_variants/$t
  for c in "c3 256 3840 2160 1 6" "c4k 1024 3840 2160 1 8"; do
    set -- $c; name=$1; shift
    timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/ncu_${t}_$name \
      python scripts/profile_case.py $@ 1 > $O/ncu_${t}_$name.log 2>&1; echo "ncu $t $name rc=$?"
    ncu -i $O/ncu_${t}_$name.ncu-rep --page raw --csv > $O/ncu_${t}_$name.raw.csv 2>/dev/null
    ncu -i $O/ncu_${t}_$name.ncu-rep --page source --csv --print-source sass > $O/ncu_${t}_$name.sass.csv 2>/dev/null
    ls -la $O/ncu_${t}_$name.*
    [ $(stat -c %s $O/ncu_${t}_$name.ncu-rep) -gt 12000000 ] && rm -f $O/ncu_${t}_$name.ncu-rep
  done
done
unset RTG_LIB_DIR
du -sh $O
