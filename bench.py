#!/usr/bin/env python
"""bench.py — the trace loop's headline benchmark (BASELINE.json metric: Mrays/s and
frames/s, % of FP32 peak, next to the reference CPU path).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 2 --warmup 1      # the reference's CPU path

A "step" is one frame of the workload: trace kernel + device quantise (+ for N > 1 the
NCCL max all-reduce, RGB8 all-gather and strip assembly).  Default workload = BASELINE
config 4, the configuration the target is quoted on: synth(1024 spheres, 4 lights),
7680x4320, alias 2 (4 spp), stack depth 8.  N GPUs render interleaved 16-row strips of the
SAME frame (strong scaling: total work fixed).

One JSON line is printed by rank 0.  `value` is device-timed with the scene resident in
HBM; `e2e` goes through the C-ABI with host buffers (scene upload H2D + RGB8 readback D2H
inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as graft  # noqa: E402

WORKLOADS = {
    # name: (spheres, lights, W, H, alias, stack)
    "config1": (0, 2, 800, 600, 3.0, 6),
    "config2": (0, 2, 1920, 1080, 1.0, 4),
    "config3": (256, 4, 3840, 2160, 1.0, 6),
    "config4": (1024, 4, 7680, 4320, 2.0, 8),
}
ZOOM = -4.0
FLOP_PER_TEST = 17.0      # hoisted ray-sphere discriminant
FLOP_PER_CONTAIN = 8.0    # primaryContainer test
STRIP_ROWS = 4        # one 8x4-tile row per strip: 4320 rows split evenly over 2, 4 and 8 ranks


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config4", choices=sorted(WORKLOADS))
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--spheres", type=int, default=-1)
    ap.add_argument("--alias", type=float, default=0.0)
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-accel", action="store_true", help="skip the optional accelerated-mode leg")
    ap.add_argument("--phases", action="store_true", help="diagnostic: also time the parts of a step (stderr)")
    ap.add_argument("--verify", action="store_true",
                    help="N>1: rank 0 also renders the whole frame alone and checks the assembled frame is byte-identical")
    return ap.parse_args()


def workload(args):
    n, l, W, H, alias, S = WORKLOADS[args.workload]
    if args.spheres >= 0:
        n = args.spheres
    W = args.width or W
    H = args.height or H
    alias = args.alias or alias
    S = args.depth or S
    return n, l, W, H, alias, S


def scene_for(pkg, n, l):
    return pkg.default_scene() if n == 0 else pkg.synth_scene(n, l)


def describe(args, n, l, W, H, alias, S, gpus):
    return {
        "workload": f"{args.workload}: {'reference default scene (main.cpp:113-168)' if n == 0 else f'synth({n} spheres, {l} lights)'}"
                    f" {W}x{H} alias {alias:g} ({int(np.ceil(alias))**2} spp) stack depth {S}",
        "spheres": 3 if n == 0 else n, "lights": l, "width": W, "height": H, "alias": alias, "max_stack": S,
        "zoom": ZOOM,
        "parallelism": ("1 GPU" if gpus == 1 else f"{gpus} GPUs x interleaved {STRIP_ROWS}-row strips, NCCL max all-reduce + RGB8 all-gather"),
        "l2": "L2 flushed (256 MiB write) before every timed step; the 16 B/px float4 framebuffer also exceeds L2 at this size",
    }


# ---------------------------------------------------------------- clocks (NVML)
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting",
               0x10: "sync_boost"}

    def __init__(self, device_index: int, uuid: str | None):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
                except Exception:
                    h = None
            self.h = h or pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        if self.h is not None:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def report(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------- CPU legs
def sample_rows(H: int, count: int):
    """`count` rows spread evenly over the frame: r = begin + k*step."""
    count = max(1, min(count, H))
    step = max(1, H // count)
    begin = step // 2
    count = min(count, (H - 1 - begin) // step + 1)
    return begin, count, step


_COUNTER_CACHE: dict = {}


def cpu_leg(om, sph, lgt, W, H, alias, S, seconds: float, threads: int = 0):
    """Time the reference CPU implementation (oracle/_ref when present, else the C port) on
    a bounded sample of rows of the SAME workload.  -> dict, (rows, framebuffer)"""
    kind = "reference" if om.reference_available(S) else "port"
    orc = om.Oracle(kind)
    port = orc if kind == "port" else om.Oracle("port")
    cores = om.Oracle("port")._port.rt_oracle_threads() if threads <= 0 else threads
    # grow the sample until it costs about `seconds` of wall time (row costs vary a lot)
    count = cores
    for _ in range(6):
        rows = sample_rows(H, count)
        t0 = time.perf_counter()
        fb, ctr = orc.render(sph, lgt, W, H, ZOOM, alias, S, rows=rows, threads=threads)
        dt = time.perf_counter() - t0
        if dt >= 0.6 * seconds or rows[1] >= H or count >= H:
            break
        grow = min(8.0, seconds / max(dt, 1e-3))
        count = int(min(H, max(count + cores, count * grow)))
        count -= count % cores
    if kind == "reference":   # the reference has no counters: count the same rows with the port, untimed
        key = (W, H, alias, S, rows, len(sph), len(lgt))
        if key not in _COUNTER_CACHE:
            _COUNTER_CACHE[key] = port.render(sph, lgt, W, H, ZOOM, alias, S, rows=rows, threads=threads)[1]
        ctr = _COUNTER_CACHE[key]
    n = len(sph)
    rays = ctr["rays"]
    out = {
        "value": rays / dt / 1e6, "unit": "Mrays/s", "cores": int(cores), "kind": kind,
        "sample": f"{rows[1]} of {H} rows (every {rows[2]}th from row {rows[0]}) of the same frame, {rays} rays, {dt:.2f} s wall, "
                  f"OpenMP dynamic over rows, -O2 -ffp-contract=off",
        "seconds": dt, "rays": rays,
        "frames_per_s_extrapolated": 1.0 / (dt * H / rows[1]),
        "gflops_17": (FLOP_PER_TEST * ctr["sphere_tests"] + FLOP_PER_CONTAIN * ctr["contain_tests"]) / dt / 1e9,
        "sphere_tests": ctr["sphere_tests"], "spheres": n,
    }
    return out, (rows, fb)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    pkg = graft.load_package()
    om = graft.load_oracle()
    n, l, W, H, alias, S = workload(args)
    sph, lgt = scene_for(pkg, n, l)
    times, vals, last = [], [], None
    for i in range(args.warmup + args.steps):
        leg, _ = cpu_leg(om, sph, lgt, W, H, alias, S, args.cpu_seconds)
        if i >= args.warmup:
            times.append(leg["seconds"])
            vals.append(leg["value"])
        last = leg
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(times) * 1e3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": describe(args, n, l, W, H, alias, S, args.gpus),
        "cpu_baseline": {k: last[k] for k in ("value", "unit", "cores", "kind", "sample")} | {"value": value},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_s": last["frames_per_s_extrapolated"],
        "note": "each step = the reference CPU render of a bounded row sample of the workload's frame; "
                "frames_per_s extrapolates the sample to the full frame",
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------- our arm
class CudaArray:
    """Expose a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int, typestr: str = "|u1", itemsize: int = 1):
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr,
                                         "data": (ptr, False), "version": 2}


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 through torch.distributed.run (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the trace loop has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pkg = graft.load_package()
    if not pkg.LIB_PATH.exists():
        pkg.build()
    n, l, W, H, alias, S = workload(args)
    sph, lgt = scene_for(pkg, n, l)
    G = world

    r = pkg.Renderer(local_rank)
    stream = torch.cuda.Stream(device=dev)
    r.set_stream(stream.cuda_stream)
    import importlib
    par = importlib.import_module(pkg.__name__ + ".parallel")
    my_rows = par.shard_rows(H, STRIP_ROWS, rank, G) if G > 1 else np.arange(H)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    r.upload_scene(sph, lgt)

    def render():
        if G > 1:
            r.render_strips(W, H, ZOOM, alias, S, STRIP_ROWS, rank, G)
        else:
            r.render(W, H, ZOOM, alias, S)

    # first frame: allocate, learn the device pointers
    with torch.cuda.stream(stream):
        render()
        r.quantise(0.0)
    r.synchronize()
    rgb_local = torch.as_tensor(CudaArray(r.device_ptr("rgb8"), len(my_rows) * W * 3), device=dev)
    max_bits = torch.as_tensor(CudaArray(r.device_ptr("max"), 4, "<i4", 4), device=dev)
    if G > 1:
        xchg = par.StripExchange(dist, torch, H, W, STRIP_ROWS, rank, G, dev)
        frame = torch.empty(H * W * 3, dtype=torch.uint8, device=dev)
    host_frame = torch.empty(H * W * 3, dtype=torch.uint8).pin_memory() if rank == 0 else None

    launches_per_step = 0

    phase_ev = []

    def mark():
        if args.phases:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            phase_ev.append(e)

    def step_device():
        """One frame, everything on the device.  Returns the number of OUR kernels launched."""
        mark()
        render()
        mark()
        k = 1
        if G > 1:
            # global normalisation (algebra.h:68-91): max over shards; non-negative floats order as ints
            xchg.reduce_max(max_bits)
        mark()
        r.quantise(0.0)
        k += 1
        mark()
        if G > 1:
            gathered = xchg.gather(rgb_local)
            mark()
            r.assemble_rgb8(gathered.data_ptr(), frame.data_ptr(), W, H, STRIP_ROWS, G, xchg.pitch)
            k += 1
        mark()
        return k

    def barrier():
        if G > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            launches_per_step = step_device()
        barrier()
        uuid = None
        try:
            uuid = str(torch.cuda.get_device_properties(dev).uuid)
        except Exception:
            pass
        # NVML queries take driver locks that can delay kernel launches: one sampler (rank 0's GPU), 5 Hz
        clocks = ClockSampler(local_rank, uuid if rank == 0 else None)
        if rank != 0:
            clocks.h = None
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        kernel_ms = []
        clocks.start()
        wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.fill_(i & 0xFF)                 # evict L2 between timed iterations (untimed)
            ev[i][0].record(stream)
            step_device()
            ev[i][1].record(stream)
        barrier()
        wall = time.perf_counter() - wall0
        clocks.stop()
    if args.phases:
        per = 6 if G > 1 else 4
        last = phase_ev[-per * args.steps:]
        names = ["render", "reduce_max", "quantise", "all_gather", "assemble"] if G > 1 else ["render", "-", "quantise"]
        acc = [0.0] * (per - 1)
        for i in range(args.steps):
            for j in range(per - 1):
                acc[j] += last[i * per + j].elapsed_time(last[i * per + j + 1])
        print(f"[rank {rank}] phases ms/step: " + ", ".join(f"{n} {a / args.steps:.3f}" for n, a in zip(names, acc)), file=sys.stderr)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    st = r.stats()
    kernel_ms = st["kernel_ms"]                   # trace kernel of the last timed step (CUDA events in the shim)

    # whole-job aggregates: rays over all ranks, time = max over ranks
    agg = torch.tensor([st["rays"], st["rays"] - st["null_rays"], st["contain_tests"], st["shadow_rays"],
                        st["samples"], st["exact_tests"], st["lane_iters"], st["active_lane_iters"]],
                       dtype=torch.float64, device=dev)
    tmax = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=dev)
    if G > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    per_rank = torch.tensor([kernel_ms, float(st["rays"]), st["active_lane_iters"] / max(1, st["lane_iters"]),
                             total_ms / args.steps], dtype=torch.float64, device=dev)
    if G > 1:
        allr = [torch.zeros_like(per_rank) for _ in range(G)]
        dist.all_gather(allr, per_rank)
        per_rank_list = [[round(float(x), 4) for x in t] for t in allr]
    else:
        per_rank_list = [[round(float(x), 4) for x in per_rank]]
    rays, live_rays, contain_tests = float(agg[0]), float(agg[1]), float(agg[2])
    total_ms_max, kernel_ms_max = float(tmax[0]), float(tmax[1])
    ms_per_step = total_ms_max / args.steps
    nsph = len(sph)
    flops_frame = FLOP_PER_TEST * live_rays * nsph + FLOP_PER_CONTAIN * contain_tests

    # ---------------- end to end through the C-ABI with host buffers
    e2e = None
    if not args.no_e2e:
        host_rgb = np.empty((len(my_rows), W, 3), np.uint8)
        with torch.cuda.stream(stream):
            def step_e2e():
                r.upload_scene(sph, lgt)                      # H2D: the scene from host arrays
                if G == 1:
                    r.render(W, H, ZOOM, alias, S)
                    r.readback_rgb8(0.0, host_rgb)            # device quantise + D2H of 3 B/px
                else:
                    step_device_after_upload()
            def step_device_after_upload():
                step_device()
                if rank == 0:
                    host_frame.copy_(frame, non_blocking=True)   # D2H of the assembled frame
                    stream.synchronize()
            step_e2e()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step_e2e()
            barrier()
            e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if G > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
        e2e = {"value": rays * args.steps / e2e_s / 1e6, "unit": "Mrays/s",
               "h2d_bytes_per_step": int(sph.nbytes + lgt.nbytes) * G, "d2h_bytes_per_step": int(W * H * 3),
               "frames_per_s": args.steps / e2e_s, "ms_per_step": e2e_s / args.steps * 1e3,
               "path": "rt_cuda_upload_scene (host AoS) -> rt_cuda_render -> rt_cuda_readback_rgb8 (host buffer)"
                       if G == 1 else
                       "rt_cuda_upload_scene -> rt_cuda_render_strips -> NCCL all-reduce(max) -> rt_cuda_quantise -> "
                       "NCCL all-gather -> rt_cuda_assemble_rgb8 -> D2H to pinned host on rank 0"}

    # ---------------- optional accelerated mode: same workload, cluster filter on
    accel = None
    if not args.no_accel and nsph >= 768:
        ka = max(1, min(3, args.steps))
        with torch.cuda.stream(stream):
            step_device()
            stream.synchronize()
            base_rgb = rgb_local.clone()
            r.set_option("accel", 1)
            step_device()
            barrier()
            eva = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(ka)]
            for i in range(ka):
                flush.fill_(i & 0xFF)
                eva[i][0].record(stream)
                step_device()
                eva[i][1].record(stream)
            barrier()
            st_a = r.stats()
            same = bool(torch.equal(base_rgb, rgb_local))
            r.set_option("accel", 0)
        ta = torch.tensor([sum(a.elapsed_time(b) for a, b in eva), st_a["kernel_ms"], 0.0 if same else 1.0],
                          dtype=torch.float64, device=dev)
        if G > 1:
            dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        accel = {"value": rays * ka / (float(ta[0]) * 1e-3) / 1e6, "unit": "Mrays/s", "steps": ka,
                 "ms_per_step": float(ta[0]) / ka, "trace_kernel_ms_max_over_ranks": float(ta[1]),
                 "speedup_vs_default": ms_per_step / (float(ta[0]) / ka),
                 "rgb8_identical_to_default": float(ta[2]) == 0.0,
                 "clusters": st_a["clusters"], "filter_tests_rank0": st_a["filter_tests"], "exact_tests_rank0": st_a["exact_tests"],
                 "note": "option accel=1: two-level cluster filter, bit-identical frame; not the headline "
                         "(the roofline above is the brute-force kernel's)"}

    # ---------------- roofline of the dominant kernel (rank 0's trace kernel)
    roofline = cpu = parity = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        sm_max = float(peaks.get("sm_max_mhz") or clocks.max_mhz or 1965.0)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        peak_nominal = sms * 128 * 2 * sm_max * 1e6 / 1e12
        ffma = max(r.ffma_peak(8192) for _ in range(3))
        flops_rank0 = FLOP_PER_TEST * (st["rays"] - st["null_rays"]) * nsph + FLOP_PER_CONTAIN * st["contain_tests"]
        achieved = flops_rank0 / (st["kernel_ms"] * 1e-3) / 1e12
        traffic = None
        tfile = ROOT / "profiles" / "traffic.json"
        if tfile.exists():
            try:
                traffic = json.loads(tfile.read_text()).get(args.workload)
            except Exception:
                traffic = None
        roofline = {
            "bound": "fp32_fma", "kernel": "rtg::trace_kernel", "achieved": achieved, "peak": peak_nominal,
            "unit": "TFLOP/s", "frac": achieved / peak_nominal,
            "peak_source": f"derived: {sms} SMs x 128 FP32 lanes x 2 x {sm_max:g} MHz (MEASURED_PEAKS.json sm_max_mhz; "
                           "it holds no FP32 figure); this path uses no tensor cores and is not HBM-bound",
            "peak_ffma_measured": ffma, "frac_of_measured_ffma": achieved / ffma,
            "algorithmic_flops_per_launch": flops_rank0,
            "flop_model": "17 flop x ray-sphere tests executed ((rays - zero-direction rays) x spheres) + 8 flop x "
                          "primaryContainer tests as the reference counts them (SURVEY.md 8d)",
            "kernel_ms": st["kernel_ms"], "traffic": traffic,
            "executed_filter_tests": st["filter_tests"], "exact_tests": st["exact_tests"],
            # FP32 operations the filter loops really issued (packed FFMA2/FADD2 count 2 lanes x 2 / x 1):
            # trace pass 30, shadow pass 39 (4 rays) or 23 (2 rays), contain pass 14 flop per lane per sphere
            "executed_tflops": 32.0 * st["sph_padded"] * (30.0 * st["passes_trace"] + 39.0 * st["passes_shadow4"]
                                                          + 23.0 * st["passes_shadow2"] + 14.0 * st["passes_contain"])
                               / (st["kernel_ms"] * 1e-3) / 1e12,
            "passes": {"trace": st["passes_trace"], "shadow4": st["passes_shadow4"], "shadow2": st["passes_shadow2"],
                       "contain": st["passes_contain"]},
            "lane_utilisation": (st["active_lane_iters"] / st["lane_iters"]) if st["lane_iters"] else None,
            "engine": {1: "persistent multi-slot kernel", 2: "wavefront (filter + shade kernels)"}.get(st["engine"]),
            "launch": {"grid": st["grid"], "block": st["block"], "smem_bytes": st["smem_bytes"],
                       "staging": {1: "__constant__", 2: "shared (TMA bulk)"}.get(st["staging"])},
        }
        if G == 1 and not args.no_cpu_baseline:
            om = graft.load_oracle()
            cpu, (rows, ref_fb) = cpu_leg(om, sph, lgt, W, H, alias, S, args.cpu_seconds)
            # the same rows on the GPU: a free parity check of the benchmark workload itself
            r.render_strips(W, H, ZOOM, alias, S, 1, rows[0] % rows[2], rows[2])
            got, _ = r.readback()
            got = got[(rows[0] // rows[2]):][: rows[1]]
            rep = om.compare(ref_fb, got)
            parity = {"rows": rows[1], "bit_exact": rep["bit_exact"], "nan_masks_equal": rep["nan_masks_equal"],
                      "within_1lsb_frac": rep["within_1lsb_frac"], "max_lsb_diff": rep["max_lsb_diff"]}
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "frames_per_s_extrapolated",
                                       "gflops_17")}

    identical = None
    if args.verify and G > 1:
        with torch.cuda.stream(stream):
            step_device()
            stream.synchronize()
            if rank == 0:
                multi = frame.cpu().numpy().reshape(H, W, 3)
                r.render(W, H, ZOOM, alias, S)
                single = r.readback_rgb8(0.0)
                identical = bool(np.array_equal(multi, single))
        barrier()

    if rank == 0:
        line = {
            "metric": "Mrays/s", "value": rays * args.steps / (total_ms_max * 1e-3) / 1e6, "unit": "Mrays/s",
            "n_gpus": G, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": describe(args, n, l, W, H, alias, S, G),
            "frames_per_s": 1e3 / ms_per_step,
            "rays_per_frame": rays, "tflops_17": flops_frame * args.steps / (total_ms_max * 1e-3) / 1e12,
            "trace_kernel_ms_max_over_ranks": kernel_ms_max,
            "per_rank": {"columns": ["trace_kernel_ms", "rays", "lane_utilisation", "step_ms"], "rows": per_rank_list},
            "wall_s_timed_region": wall,
            "clocks": clocks.report(),
            "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "parity_sample": parity,
            "accelerated_mode": accel,
        }
        if identical is not None:
            line["multi_gpu_frame_identical_to_1gpu"] = identical
        print(json.dumps(line))
    r.close()
    if G > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
