#!/bin/bash
# Round-2 measurement campaign on N GPUs of one box: usage r2_campaign.sh <N> ["list of gpu counts to bench"] [nosweep] [c4only]
# -> gpurun_out/campaign_n<N>/ (bench lines for config 4 and config 3, config-5 sweep, host program, GPU tests)
set -u
N=$1; counts=${2:-$N}
O=gpurun_out/campaign_n$N; mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1; nvidia-smi topo -m >> $O/gpus.txt 2>&1
run_bench() {   # n workload tag
  local n=$1 wl=$2 out=$O/bench_$2_n$1.json
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --workload $wl > $out 2> $O/bench_$2_n$1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 5 --warmup 3 --workload $wl > $out 2> $O/bench_$2_n$1.err
  fi
  echo "bench $wl n=$n rc=$?"
  python - <<PY
import json
try:
    d = [json.loads(l) for l in open("$out") if l.startswith("{")][-1]
    nm = d.get("native_multi") or {}
    print("  value", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3),
          "identical", d.get("multi_gpu_frame_identical_to_1gpu"), "native", nm.get("value") and round(nm["value"], 1), nm.get("frame_identical_to_1gpu"), nm.get("error"))
    print("  per_rank", [r[0] for r in d["per_rank"]["rows"]], "fill", [r[2] for r in d["per_rank"]["rows"]][:2])
except Exception as e:
    print("  no JSON:", e)
PY
}
if [ $N -gt 1 ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q -k "multi or strips" > $O/pytest_multi.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_multi.txt
fi
for n in $counts; do run_bench $n config4; done
[ "${4:-}" = "c4only" ] || for n in $counts; do run_bench $n config3; done
if [ $N -eq 1 ]; then
  for wl in config1 config2; do run_bench 1 $wl; done
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "reference arm rc=$?"
fi
if [ $N -eq 1 ]; then
  # the ncu launch list of the bench command (cold-cache, serialised: shares, not absolute times)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-accel > $O/ncu_launches_bench.log 2>&1; echo "ncu launch list rc=$?"
  { echo "== OpenCL ICD probe"; ls -la /etc/OpenCL/vendors 2>&1; ls /usr/lib/x86_64-linux-gnu 2>/dev/null | grep -i -E "opencl"; ldconfig -p | grep -i opencl; which clinfo; echo "(end)"; } > $O/opencl_probe.txt 2>&1
fi
[ "${3:-}" = "nosweep" ] || timeout 900 python scripts/sweep.py --gpus $N > $O/sweep_config5_n$N.jsonl 2> $O/sweep_n$N.err; echo "sweep rc=$?"; tail -3 $O/sweep_config5_n$N.jsonl | cut -c1-260
if [ $N -gt 1 ]; then
  timeout 600 raytracer-gamma_b200/rt_gamma --gpus $N --spheres 1024 --width 7680 --height 4320 --alias 2 --depth 8 --frames 3 --out $O/frame.ppm > $O/rt_gamma_multi.txt 2>&1; echo "rt_gamma rc=$?"
  md5sum $O/frame.ppm >> $O/rt_gamma_multi.txt; tail -8 $O/rt_gamma_multi.txt; rm -f $O/frame.ppm
fi
du -sh $O
