#!/bin/bash
# Multi-GPU validation + numbers on N GPUs of one box: usage r2_multi.sh <tag> <N> [full]
set -u
tag=$1; N=$2; full=${3:-}
O=gpurun_out/r2_multi_n$N; mkdir -p $O
export RTG_LIB_DIR=$PWD/build_variants/$tag
nvidia-smi -L > $O/gpus.txt 2>&1; nvidia-smi topo -m >> $O/gpus.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -k "multi or strips" > $O/pytest_multi.txt 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_multi.txt
for n in $( [ -n "$full" ] && echo "2 4 8" || echo $N ); do
  [ $n -le $N ] || continue
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 5 --warmup 3 > $O/bench_n$n.json 2> $O/bench_n$n.err; echo "bench n=$n rc=$?"
  tail -c 600 $O/bench_n$n.err
  python - <<PY
import json
try:
    d = json.load(open("$O/bench_n$n.json"))
    print("n=$n value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "identical", d.get("multi_gpu_frame_identical_to_1gpu"),
          "native", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in (d.get("native_multi") or {}).items() if k in ("value", "ms_per_step", "frame_identical_to_1gpu", "error")})
    print("   per_rank", d["per_rank"]["rows"])
except Exception as e:
    print("bench n=$n: no JSON", e)
PY
done
timeout 600 $RTG_LIB_DIR/rt_gamma --gpus $N --spheres 1024 --width 7680 --height 4320 --alias 2 --depth 8 --frames 3 --out $O/frame.ppm > $O/rt_gamma_multi.txt 2>&1; echo "rt_gamma rc=$?"; tail -14 $O/rt_gamma_multi.txt; rm -f $O/frame.ppm
