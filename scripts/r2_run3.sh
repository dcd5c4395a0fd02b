set -u
bash scripts/r2_ab.sh "head1 v2a v2a:slot_mode=2 v2b v2b:slot_mode=2" synth256 "synth1024 4K a1" "synth1024 4K a2"
O=gpurun_out/prof_v2a_local; mkdir -p $O
RTG_LIB_DIR=$PWD/build_variants/v2a RTG_OPTS=slot_mode=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/prof_c4k \
    python scripts/profile_case.py 1024 3840 2160 1 8 1 > $O/ncu_c4k.log 2>&1
O=gpurun_out/prof_v2b; mkdir -p $O
RTG_LIB_DIR=$PWD/build_variants/v2b timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/prof_c4k \
    python scripts/profile_case.py 1024 3840 2160 1 8 1 > $O/ncu_c4k.log 2>&1
