#!/bin/bash
# 2-GPU validation of the multi-GPU paths: GPU tests (incl. the NCCL ones), bench at N=2
set -u
O=gpurun_out/multi2; mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n=2 rc=$?"; tail -c 400 $O/bench_n2.err
python - <<PY
import json
d = json.load(open("$O/bench_n2.json"))
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 1), "identical", d.get("multi_gpu_frame_identical_to_1gpu"),
      "native", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in (d.get("native_multi") or {}).items() if k in ("value", "ms_per_step", "frame_identical_to_1gpu", "error")})
print("per_rank", d["per_rank"]["rows"], "frac", round(d["roofline"]["frac"], 3))
PY
