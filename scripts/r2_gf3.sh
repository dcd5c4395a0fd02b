#!/bin/bash
set -u
O=gpurun_out/gf3; mkdir -p $O
bash scripts/r2_ab.sh "fin gf gf2 fin gf gf2" "synth1024 8K a2" > $O/ab.txt 2>&1; cat $O/ab.txt
for t in gf fin; do
export RTG_LIB_DIR=$PWD/build_variants/$t
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:trace_kernel -c 1 --csv --log-file $O/ncu_c4_dram_$t.csv python scripts/profile_case.py 1024 7680 4320 2 8 1 > $O/ncu_c4_$t.log 2>&1; echo "ncu $t rc=$?"; grep -E "dram|duration" $O/ncu_c4_dram_$t.csv | cut -c1-300
done
