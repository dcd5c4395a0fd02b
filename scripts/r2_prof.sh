#!/bin/bash
# ncu full captures (source-level) of a frozen build: usage r2_prof.sh <tag> [opts]   -> gpurun_out/prof_<tag>/
set -u
tag=$1; opts=${2:-}
O=gpurun_out/prof_$tag; mkdir -p $O
export RTG_LIB_DIR=$PWD/build_variants/$tag
export RTG_OPTS=$opts
python scripts/profile_case.py 256 3840 2160 1 6 2 > $O/c3_plain.log 2>&1 || { cat $O/c3_plain.log; exit 1; }
python scripts/profile_case.py 1024 3840 2160 1 8 2 > $O/c4k_plain.log 2>&1 || exit 1
cat $O/c3_plain.log $O/c4k_plain.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/prof_c3 \
    python scripts/profile_case.py 256 3840 2160 1 6 1 > $O/ncu_c3.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/prof_c4k \
    python scripts/profile_case.py 1024 3840 2160 1 8 1 > $O/ncu_c4k.log 2>&1
ls -la $O
