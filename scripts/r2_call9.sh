#!/bin/bash
# GPU call 9: cost of one record unpack+pack (doubled in variant dbl); ncu full captures of the current build
set -u
O=gpurun_out/call9; mkdir -p $O
bash scripts/r2_ab.sh "cur dbl cur dbl" synth256 "synth1024 4K a1" "synth1024 4K a2" > $O/ab.txt 2>&1; cat $O/ab.txt
export RTG_LIB_DIR=$PWD/build_variants/cur
for c in "c3 256 3840 2160 1 6" "c4k 1024 3840 2160 1 8" "c4 1024 7680 4320 2 8"; do
  set -- $c; name=$1; shift
  timeout 900 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/ncu_cur_$name \
    python scripts/profile_case.py $@ 1 > $O/ncu_cur_$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i $O/ncu_cur_$name.ncu-rep --page raw --csv > $O/ncu_cur_$name.raw.csv 2>/dev/null
  ncu -i $O/ncu_cur_$name.ncu-rep --page source --csv --print-source sass > $O/ncu_cur_$name.sass.csv 2>/dev/null
  ls -la $O/ncu_cur_$name.ncu-rep; [ $(stat -c %s $O/ncu_cur_$name.ncu-rep) -gt 12000000 ] && rm -f $O/ncu_cur_$name.ncu-rep
done
RTG_OPTS=accel=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/ncu_cur_accel4k \
    python scripts/profile_case.py 1024 3840 2160 1 8 1 > $O/ncu_cur_accel4k.log 2>&1; echo "ncu accel rc=$?"
ncu -i $O/ncu_cur_accel4k.ncu-rep --page raw --csv > $O/ncu_cur_accel4k.raw.csv 2>/dev/null
ncu -i $O/ncu_cur_accel4k.ncu-rep --page source --csv --print-source sass > $O/ncu_cur_accel4k.sass.csv 2>/dev/null
rm -f $O/ncu_cur_accel4k.ncu-rep
du -sh $O
