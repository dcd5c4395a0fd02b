#!/bin/bash
# GPU call 2: instruction-cache experiments (group sizes, lockstep passes)
set -u
O=gpurun_out/call2; mkdir -p $O
for t in lock lockg8; do
  RTG_LIB_DIR=$PWD/build_variants/$t timeout 600 python -m pytest tests -m gpu -x -q -k "golden or config1 or synthetic or rare_paths or strips_and" > $O/pytest_$t.txt 2>&1; echo "pytest $t rc=$?"; tail -3 $O/pytest_$t.txt
done
bash scripts/r2_ab.sh "c0s0 g8 g8s4 g4 lock lockg8 head1 head1_g8 c0s0" synth256 "synth1024 4K a1" "synth1024 4K a2" > $O/ab.txt 2>&1; cat $O/ab.txt
for t in lockg8 g8; do
  export RTG_LIB_DIR=$PWD/build_variants/$t
  for c in "c3 256 3840 2160 1 6"; do
    set -- $c; name=$1; shift
    timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/ncu_${t}_$name \
      python scripts/profile_case.py $@ 1 > $O/ncu_${t}_$name.log 2>&1; echo "ncu $t $name rc=$?"
    ncu -i $O/ncu_${t}_$name.ncu-rep --page raw --csv > $O/ncu_${t}_$name.raw.csv 2>/dev/null
    ncu -i $O/ncu_${t}_$name.ncu-rep --page source --csv --print-source sass > $O/ncu_${t}_$name.sass.csv 2>/dev/null
    rm -f $O/ncu_${t}_$name.ncu-rep
  done
done
unset RTG_LIB_DIR
du -sh $O
