#!/bin/bash
# last check of the round: GPU tests, smoke, and the bench lines of the final build
set -u
O=gpurun_out/last6; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > $O/bench_config4_n1.json 2> $O/bench.err; echo "bench rc=$?"
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --workload config3 > $O/bench_config3_n1.json 2> $O/bench3.err; echo "bench c3 rc=$?"
timeout 600 python scripts/sweep.py --gpus 1 > $O/sweep_config5_n1.jsonl 2>&1; echo "sweep rc=$?"
python - <<PY
import json
for f in ("bench_config4_n1", "bench_config3_n1"):
    d = [json.loads(l) for l in open("$O/" + f + ".json") if l.startswith("{")][-1]
    a = d.get("accelerated_mode") or {}
    print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "cpu", round(d["cpu_baseline"]["value"], 2), d["parity_sample"]["bit_exact"], "accel", a.get("ms_per_step"), a.get("speedup_vs_default"), a.get("rgb8_identical_to_default"))
PY
