#!/bin/bash
# 1-GPU validation of the round-2 host-side work: GPU tests, short bench (both arms), host program animation.
set -u
O=gpurun_out/r2_val1; mkdir -p $O
export RTG_LIB_DIR=$PWD/build_variants/head1
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu.txt
tail -15 $O/pytest_gpu.txt
python bench.py --steps 3 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
tail -c 1500 $O/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 --cpu-seconds 6 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
python -c "
import json
d=json.load(open('$O/bench_n1.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),round(d['e2e']['ms_per_step'],2),'frac',round(d['roofline']['frac'],3),'cpu',d['cpu_baseline']['value'],d['cpu_baseline']['cores'],'parity',d['parity_sample'])
r=json.load(open('$O/bench_ref.json')); print('ref',r['value'],r['cpu_baseline']['cores'],r['seconds_per_step'])
"
