#!/bin/bash
# last call of the round: GPU tests, smoke, the bench line and the reference arm on the final build
set -u
O=gpurun_out/last2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.txt
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > $O/bench_config4_n1.json 2> $O/bench.err; echo "bench rc=$?"
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --workload config3 > $O/bench_config3_n1.json 2> $O/bench3.err; echo "bench c3 rc=$?"
python - <<PY
import json
for f in ("bench_config4_n1", "bench_config3_n1"):
    d = [json.loads(l) for l in open("$O/" + f + ".json") if l.startswith("{")][-1]
    print(f, "value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"], "cpu", round(d["cpu_baseline"]["value"], 2), d["parity_sample"]["bit_exact"], d["clocks"])
PY
