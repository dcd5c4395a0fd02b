set -u
O=gpurun_out/r2_v2c; mkdir -p $O
RTG_LIB_DIR=$PWD/build_variants/v2c timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.txt
bash scripts/r2_ab.sh "head1 v2c v2c:slot_mode=2" synth256 "synth1024 4K a1" "synth1024 4K a2" synth4096
O=gpurun_out/prof_v2c; mkdir -p $O
RTG_LIB_DIR=$PWD/build_variants/v2c timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/prof_c4k \
    python scripts/profile_case.py 1024 3840 2160 1 8 1 > $O/ncu_c4k.log 2>&1
RTG_LIB_DIR=$PWD/build_variants/v2c timeout 600 ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/prof_c3 \
    python scripts/profile_case.py 256 3840 2160 1 6 1 > $O/ncu_c3.log 2>&1
