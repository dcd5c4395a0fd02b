set -u
O=gpurun_out/r2_v2d; mkdir -p $O
RTG_LIB_DIR=$PWD/build_variants/v2d timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.txt
bash scripts/r2_ab.sh "v2d:rebalance=0,sparse_below=0 c0s1:rebalance=0,sparse_below=0 c0s0:rebalance=0,sparse_below=0 c1s0:rebalance=0,sparse_below=0 v2d v2d:sparse_below=0 v2d:rebalance=0" synth256 "synth1024 4K a1" "synth1024 4K a2" "accel synth1024"
for o in "rebalance=0,sparse_below=0" "sparse_below=0" "" "sparse_below=24"; do echo "== tail v2d [$o]"; RTG_LIB_DIR=$PWD/build_variants/v2d RTG_OPTS=$o timeout 300 python scripts/tail_probe.py 4 2>&1 | cut -c1-200; done
