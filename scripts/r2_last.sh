#!/bin/bash
set -u
O=gpurun_out/last; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
bash scripts/r2_ab.sh "cur4 dd cur4 dd" synth256 synth512 "synth1024 4K a2" "synth1024 8K a2" > $O/ab.txt 2>&1; cat $O/ab.txt
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > $O/bench_config4_n1.json 2> $O/bench.err; echo "bench rc=$?"
python - <<PY
import json
d = [json.loads(l) for l in open("$O/bench_config4_n1.json") if l.startswith("{")][-1]
print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["e2e"]["host_time_after_each_step_ms"], "frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"], "cpu", round(d["cpu_baseline"]["value"], 2), d["parity_sample"])
PY
python __graft_entry__.py smoke 2>&1 | tail -4
