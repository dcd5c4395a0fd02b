#!/bin/bash
# ncu full captures (source-level) of a frozen build: usage r2_final_prof.sh <tag>   -> gpurun_out/prof_<tag>/
set -u
tag=$1
O=gpurun_out/prof_$tag; mkdir -p $O
export RTG_LIB_DIR=$PWD/build_variants/$tag
for c in "c3 256 3840 2160 1 6" "c4k 1024 3840 2160 1 8"; do
  set -- $c; name=$1; shift
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/ncu_$name \
    python scripts/profile_case.py $@ 1 > $O/ncu_$name.log 2>&1; echo "ncu $name rc=$?"
  ncu -i $O/ncu_$name.ncu-rep --page raw --csv > $O/ncu_$name.raw.csv 2>/dev/null
  ncu -i $O/ncu_$name.ncu-rep --page source --csv --print-source sass > $O/ncu_$name.sass.csv 2>/dev/null
  rm -f $O/ncu_$name.ncu-rep
done
# the bench frame itself (8K, 4 spp): full set without the source page (ncu saves and restores 2.6 GB per replay pass)
timeout 1200 ncu --set full --clock-control none -k regex:trace_kernel -c 1 -f -o $O/ncu_c4 python scripts/profile_case.py 1024 7680 4320 2 8 1 > $O/ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
ncu -i $O/ncu_c4.ncu-rep --page raw --csv > $O/ncu_c4.raw.csv 2>/dev/null; rm -f $O/ncu_c4.ncu-rep
du -sh $O
