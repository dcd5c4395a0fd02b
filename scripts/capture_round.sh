#!/bin/bash
# One gpurun call that produces everything scripts/collect_profiles.py files under profiles/<round>/.
# Every command runs once WITHOUT ncu first; numbers printed under ncu are never bench values.
set -u
O=gpurun_out
python bench.py > $O/bench_r1_final.json 2> $O/bench_r1_final.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r1_reference.json 2> $O/bench_r1_reference.err
for c in config1 config2 config3; do
  python bench.py --workload $c --steps 10 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err
done
python scripts/sweep.py > $O/sweep_r1.jsonl 2> $O/sweep_r1.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-accel > $O/bench_short.json 2> $O/bench_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r1.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-accel > $O/ncu_launches.log 2>&1
python scripts/profile_case.py 1024 7680 4320 2 8 1 > $O/profile_case.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:trace_kernel -c 1 -f -o $O/prof_r1_final \
    python scripts/profile_case.py 1024 7680 4320 2 8 1 > $O/ncu_full.log 2>&1
tail -1 $O/bench_r1_final.json | cut -c1-300
