#!/bin/bash
# ncu full captures (source-level) of a frozen build + the lockstep threshold: usage r2_final_prof.sh <tag>
set -u
tag=$1
O=gpurun_out/prof_$tag; mkdir -p $O
bash scripts/r2_ab.sh "$tag:lockstep=1 $tag:lockstep=2 $tag:lockstep=1 $tag:lockstep=2" synth256 synth512 synth768 "synth1024 4K a1" > $O/ab_lockstep_threshold.txt 2>&1; cat $O/ab_lockstep_threshold.txt
export RTG_LIB_DIR=$PWD/build_variants/$tag
for c in "c3 256 3840 2160 1 6" "c4k 1024 3840 2160 1 8"; do
  set -- $c; name=$1; shift
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/ncu_$name \
    python scripts/profile_case.py $@ 1 > $O/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i $O/ncu_$name.ncu-rep --page raw --csv > $O/ncu_$name.raw.csv 2>/dev/null
  ncu -i $O/ncu_$name.ncu-rep --page source --csv --print-source sass > $O/ncu_$name.sass.csv 2>/dev/null
  rm -f $O/ncu_$name.ncu-rep
done
RTG_OPTS=accel=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/ncu_accel4k \
    python scripts/profile_case.py 1024 3840 2160 1 8 1 > $O/ncu_accel4k.log 2>&1; echo "ncu accel rc=$?"
ncu -i $O/ncu_accel4k.ncu-rep --page raw --csv > $O/ncu_accel4k.raw.csv 2>/dev/null
ncu -i $O/ncu_accel4k.ncu-rep --page source --csv --print-source sass > $O/ncu_accel4k.sass.csv 2>/dev/null
rm -f $O/ncu_accel4k.ncu-rep
du -sh $O
