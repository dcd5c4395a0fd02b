"""Development aid: copy the artefacts of the last gpurun capture set into profiles/<round>/ and
derive the summaries (ncu kernel summary, launch shares, traffic.json)."""
import collections, csv, json, shutil, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
G = ROOT / "gpurun_out"; P = ROOT / "profiles" / rnd; P.mkdir(parents=True, exist_ok=True)
for src, dst in [("bench_r1_final.json", "bench_n1.json"), ("bench_r1_reference.json", "bench_reference_arm.json"),
                 ("launches_r1.csv", "ncu_launches_bench.csv"), ("sweep_r1.jsonl", "sweep_config5.jsonl"),
                 ("bench_config1.json", "bench_config1.json"), ("bench_config2.json", "bench_config2.json"),
                 ("bench_config3.json", "bench_config3.json")]:
    if (G / src).exists():
        shutil.copy(G / src, P / dst)
rep = G / "prof_r1_final.ncu-rep"
raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_active.avg", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"]
out = {}
for h, u, v in zip(hdr, units, vals):
    if h in want or ("issue_stalled" in h and "per_issue_active" in h):
        out[h] = {"value": v, "unit": u}
out["_note"] = ("ncu --set full --clock-control none, one launch of trace_kernel on the bench workload "
                "(synth 1024 spheres, 7680x4320, alias 2, stack 8); per-launch values")
(P / "ncu_trace_kernel_full.json").write_text(json.dumps(out, indent=1) + "\n")
def nbytes(x):
    return float(x["value"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[x["unit"]]
traffic = nbytes(out["dram__bytes_read.sum"]) + nbytes(out["dram__bytes_write.sum"])
(ROOT / "profiles" / "traffic.json").write_text(json.dumps({
    "config4": traffic,
    "_note": "dram__bytes_read.sum + dram__bytes_write.sum of one trace_kernel launch on the bench workload "
             f"(profiles/{rnd}/ncu_trace_kernel_full.json).  It is the per-lane sample state (34-word slot records and "
             "call-stack frames of ~300k samples in flight) cycling through L2, not algorithmic bytes (2.1 GB of sample "
             "results); the kernel is FP32-pipe-bound, DRAM throughput stays a few per cent of peak."}, indent=1) + "\n")
rows = list(csv.DictReader(l for l in open(P / "ncu_launches_bench.csv") if l.startswith('"')))
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    k = r["Kernel Name"].split("(")[0]; agg[k][0] += 1; agg[k][1] += float(r["Metric Value"]) / 1e6
tot = sum(v[1] for v in agg.values())
lines = ["# ncu launch list of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-accel` (cold-cache, serialised: compare shares)",
         "", "kernel | launches | total ms | share", "---|---|---|---"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k} | {v[0]} | {v[1]:.3f} | {v[1] / tot * 100:.2f} %")
(P / "ncu_launches_summary.md").write_text("\n".join(lines) + "\n")
print("\n".join(lines))
for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
          "launch__registers_per_thread"):
    print(k, out.get(k))
print("traffic", traffic)
