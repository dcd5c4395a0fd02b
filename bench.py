#!/usr/bin/env python
"""bench.py — the trace loop's headline benchmark (BASELINE.json metric: Mrays/s and
frames/s, % of FP32 peak, next to the reference CPU path).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 2 --warmup 1      # the reference's CPU path

A "step" is one frame of the workload: trace kernel + device quantise (+ for N > 1 the
NCCL max all-reduce, RGB8 all-gather and strip assembly, all inside the C-ABI library
librt_cuda_multi.so).  Default workload = BASELINE config 4, the configuration the target
is quoted on: synth(1024 spheres, 4 lights), 7680x4320, alias 2 (4 spp), stack depth 8.
N GPUs render interleaved 4-row strips of the SAME frame (strong scaling: total work fixed).

One JSON line is printed by rank 0.  `value` is device-timed with the scene resident in
HBM; `e2e` goes through the C-ABI with host buffers (scene upload H2D + RGB8 readback D2H
inside the timed region).  At N > 1 the line also carries
  multi_gpu_frame_identical_to_1gpu   rank 0 renders the frame alone (untimed) and compares bytes
  native_multi                        the same frame through ONE process driving all N GPUs
                                      (rt_cuda_multi_init / ncclCommInitAll — what `rt_gamma --gpus N` runs)
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
    ap.add_argument("--no-native", action="store_true", help="N>1: skip the single-process (ncclCommInitAll) leg")
    ap.add_argument("--verify", action="store_true", help="accepted for compatibility: the identity check is always on")
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
        "parallelism": ("1 GPU" if gpus == 1 else f"{gpus} GPUs x interleaved {STRIP_ROWS}-row strips, NCCL max all-reduce + RGB8 all-gather "
                        "inside librt_cuda_multi.so (one process per GPU)"),
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


def host_threads() -> int:
    """The cores this process may run on.  torchrun exports OMP_NUM_THREADS=1 to its workers, so the
    OpenMP default is useless there: the CPU legs always pass this count explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def calibrate_rows(orc, sph, lgt, W, H, alias, S, seconds: float, threads: int):
    """Grow a row sample of the workload's frame until one render of it costs about `seconds`."""
    count = threads
    rows = sample_rows(H, count)
    for _ in range(6):
        rows = sample_rows(H, count)
        t0 = time.perf_counter()
        orc.render(sph, lgt, W, H, ZOOM, alias, S, rows=rows, threads=threads)
        dt = time.perf_counter() - t0
        if dt >= 0.6 * seconds or rows[1] >= H or count >= H:
            break
        grow = min(8.0, seconds / max(dt, 1e-3))
        count = int(min(H, max(count + threads, count * grow)))
        count -= count % threads
    return rows


def cpu_arm(om, sph, lgt, W, H, alias, S, seconds: float, steps: int, warmup: int):
    """The reference's CPU implementation (oracle/_ref when present, else the C port) on a bounded, FIXED
    sample of rows of the workload's frame, all host threads: calibrate the sample once, then time
    `steps` renders of it.  -> (dict, (rows, framebuffer))"""
    kind = "reference" if om.reference_available(S) else "port"
    orc = om.Oracle(kind)
    threads = host_threads()
    rows = calibrate_rows(orc, sph, lgt, W, H, alias, S, seconds, threads)
    times, fb = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        fb, ctr = orc.render(sph, lgt, W, H, ZOOM, alias, S, rows=rows, threads=threads)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    if kind == "reference":   # the reference has no counters: count the same rows with the port, untimed
        ctr = om.Oracle("port").render(sph, lgt, W, H, ZOOM, alias, S, rows=rows, threads=threads)[1]
    dt = float(np.median(times))
    rays = ctr["rays"]
    out = {
        "value": rays / dt / 1e6, "unit": "Mrays/s", "cores": int(threads), "kind": kind,
        "sample": f"{rows[1]} of {H} rows (every {rows[2]}th from row {rows[0]}) of the same frame, {rays} rays, "
                  f"median {dt:.2f} s of {len(times)} timed renders, OpenMP dynamic over rows on {threads} threads, -O2 -ffp-contract=off",
        "seconds": dt, "seconds_all": [round(t, 4) for t in times], "rays": rays,
        "frames_per_s_extrapolated": 1.0 / (dt * H / rows[1]),
        "gflops_17": (FLOP_PER_TEST * ctr["sphere_tests"] + FLOP_PER_CONTAIN * ctr["contain_tests"]) / dt / 1e9,
        "sphere_tests": ctr["sphere_tests"], "spheres": len(sph),
    }
    return out, (rows, fb)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    pkg = graft.load_package()      # scene builders only (librt_scene.so, host code): no CUDA library is loaded here
    om = graft.load_oracle()
    n, l, W, H, alias, S = workload(args)
    sph, lgt = scene_for(pkg, n, l)
    leg, _ = cpu_arm(om, sph, lgt, W, H, alias, S, args.cpu_seconds, args.steps, args.warmup)
    value = leg["value"]
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": leg["seconds"] * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": describe(args, n, l, W, H, alias, S, args.gpus),
        "cpu_baseline": {k: leg[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_s": leg["frames_per_s_extrapolated"], "seconds_per_step": leg["seconds_all"],
        "note": "each step = the reference CPU render of one fixed, bounded row sample of the workload's frame "
                "(calibrated once, untimed); value = rays of the sample / median step time; frames_per_s "
                "extrapolates the sample to the full frame",
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
    ctl = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        ctl = dist.new_group(backend="gloo")        # host-side barriers that keep the GPUs idle

    pkg = graft.load_package()
    if not pkg.LIB_PATH.exists():
        pkg.build()
    n, l, W, H, alias, S = workload(args)
    sph, lgt = scene_for(pkg, n, l)
    G = world
    frame_bytes = W * H * 3

    stream = torch.cuda.Stream(device=dev)
    if G > 1:
        # one process per GPU; the exchange (NCCL max all-reduce, RGB8 all-gather) runs inside the C-ABI library
        box = [pkg.multi_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        r = pkg.MultiRenderer(rank=rank, world=G, uid=box[0], device=local_rank)
        r.set_stream(stream.cuda_stream)
    else:
        r = pkg.Renderer(local_rank)
        r.set_stream(stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    r.upload_scene(sph, lgt)

    def step_device():
        """One frame, everything on the device.  Returns the number of OUR kernels launched."""
        if G > 1:
            r.render(W, H, ZOOM, alias, S, STRIP_ROWS)      # trace + combine + quantise + assemble (+ 2 NCCL kernels)
            return 4 if alias > 1.0 else 3
        r.render(W, H, ZOOM, alias, S)
        r.quantise(0.0)
        return 3 if alias > 1.0 else 2

    def stats():
        return r.stats(0) if G > 1 else r.stats()

    def barrier():
        if G > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    with torch.cuda.stream(stream):
        launches_per_step = step_device()        # first frame: allocations
        for _ in range(args.warmup):
            step_device()
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
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    st = stats()
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

    # ---------------- end to end through the C-ABI with host buffers: every step uploads the scene from host
    # arrays (H2D) and brings the quantised frame back into pinned host memory (D2H, asynchronous: the copy of
    # frame k overlaps the render of frame k+1; everything has landed before the clock stops)
    e2e = None
    if not args.no_e2e:
        host = [pkg.HostBuffer(frame_bytes), pkg.HostBuffer(frame_bytes)] if rank == 0 else None
        with torch.cuda.stream(stream):
            tickets = [None, None]

            def step_e2e(k):
                r.upload_scene(sph, lgt)                              # H2D: the scene from host arrays
                if G == 1:
                    r.render(W, H, ZOOM, alias, S)
                    tickets[k & 1] = r.readback_rgb8_async(host[k & 1], 0.0)    # device quantise + D2H of 3 B/px
                    if k > 0:
                        r.readback_wait(tickets[(k - 1) & 1])
                else:
                    r.render(W, H, ZOOM, alias, S, STRIP_ROWS)
                    if rank == 0:
                        r.readback_wait(0)                            # frame k-1 has landed (one copy in flight)
                        r.readback_rgb8_async(host[k & 1], 0)         # D2H of the assembled frame

            def finish_e2e(k):
                if G == 1:
                    r.readback_wait(tickets[(k - 1) & 1])
                elif rank == 0:
                    r.readback_wait(0)
                r.synchronize()

            step_e2e(0)          # two untimed frames: both readback tickets (device + pinned buffers) exist before the clock starts
            step_e2e(1)
            finish_e2e(2)
            barrier()
            t0 = time.perf_counter()
            marks = []
            for k in range(args.steps):
                step_e2e(k)
                marks.append(time.perf_counter() - t0)
            finish_e2e(args.steps)
            barrier()
            e2e_s = time.perf_counter() - t0
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        if G > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
        e2e = {"value": rays * args.steps / e2e_s / 1e6, "unit": "Mrays/s",
               "h2d_bytes_per_step": int(sph.nbytes + lgt.nbytes) * G, "d2h_bytes_per_step": int(frame_bytes),
               "frames_per_s": args.steps / e2e_s, "ms_per_step": e2e_s / args.steps * 1e3,
               "host_time_after_each_step_ms": [round(m * 1e3, 2) for m in marks],
               "path": "rt_cuda_upload_scene (host AoS) -> rt_cuda_render -> rt_cuda_readback_rgb8_async (pinned host "
                       "double buffer; the copy of frame k overlaps the render of frame k+1) -> rt_cuda_readback_wait"
                       if G == 1 else
                       "rt_cuda_multi_upload_scene -> rt_cuda_multi_render (strips, NCCL all-reduce(max), quantise, NCCL "
                       "all-gather, assemble: inside librt_cuda_multi.so) -> rt_cuda_multi_readback_rgb8_async to pinned host on rank 0"}

    # ---------------- optional accelerated mode: same workload, cluster filter on
    accel = None
    if not args.no_accel and nsph >= 768:
        ka = max(1, min(3, args.steps))
        with torch.cuda.stream(stream):
            step_device()
            stream.synchronize()
            base_rgb = r.readback_rgb8(0)[0].copy() if G > 1 else r.readback_rgb8(0.0).copy()
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
            st_a = stats()
            accel_rgb = r.readback_rgb8(0)[0] if G > 1 else r.readback_rgb8(0.0)
            same = bool(np.array_equal(base_rgb, accel_rgb))
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

    # ---------------- N > 1: the assembled frame against the frame rank 0 renders alone (always on), and the
    # same workload through ONE process driving all N GPUs (the native path of `rt_gamma --gpus N`)
    identical = native = None
    if G > 1:
        with torch.cuda.stream(stream):
            step_device()
            stream.synchronize()
        single = None
        if rank == 0:
            multi_frame = r.readback_rgb8(0)[0].copy()
            with pkg.Renderer(local_rank) as one:
                one.upload_scene(sph, lgt)
                one.render(W, H, ZOOM, alias, S)
                single = one.readback_rgb8(0.0)
            identical = bool(np.array_equal(multi_frame, single))
            del multi_frame
        dist.barrier(group=ctl)
        if not args.no_native:
            if rank == 0:
                try:
                    native = native_leg(pkg, G, sph, lgt, W, H, alias, S, args, rays, single)
                except Exception as e:          # reported, never hidden
                    native = {"error": f"{type(e).__name__}: {e}"}
            dist.barrier(group=ctl)             # the other ranks wait on the CPU: their GPUs stay idle for rank 0
        del single

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
        with pkg.Renderer(local_rank) as probe:
            ffma = max(probe.ffma_peak(8192) for _ in range(3))
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
            "engine": "persistent multi-slot kernel",
            "launch": {"grid": st["grid"], "block": st["block"], "smem_bytes": st["smem_bytes"],
                       "staging": {1: "constant bank (launch parameter)", 2: "shared (TMA bulk)"}.get(st["staging"])},
        }
        if G == 1 and not args.no_cpu_baseline:
            om = graft.load_oracle()
            cpu, (rows, ref_fb) = cpu_arm(om, sph, lgt, W, H, alias, S, args.cpu_seconds, 1, 0)
            # the same rows on the GPU: a free parity check of the benchmark workload itself
            r.render_strips(W, H, ZOOM, alias, S, 1, rows[0] % rows[2], rows[2])
            got, _ = r.readback()
            got = got[(rows[0] // rows[2]):][: rows[1]]
            rep = om.compare(ref_fb, got)
            parity = {"rows": rows[1], "bit_exact": rep["bit_exact"], "nan_masks_equal": rep["nan_masks_equal"],
                      "within_1lsb_frac": rep["within_1lsb_frac"], "max_lsb_diff": rep["max_lsb_diff"]}
            cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "frames_per_s_extrapolated",
                                       "gflops_17")}

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
        if G > 1:
            line["multi_gpu_frame_identical_to_1gpu"] = identical
            line["native_multi"] = native
        print(json.dumps(line))
    r.close()
    if G > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def native_leg(pkg, G, sph, lgt, W, H, alias, S, args, rays, single_frame):
    """The workload through ONE process driving all G GPUs of the box: rt_cuda_multi_init (ncclCommInitAll,
    one stream per device), no Python in the data path.  Device-timed per step with CUDA events on every
    device's stream (the library's), max over devices; L2 flushed on every device between steps."""
    with pkg.MultiRenderer(gpus=G) as m:
        m.upload_scene(sph, lgt)
        for _ in range(max(1, args.warmup)):
            m.render(W, H, ZOOM, alias, S, STRIP_ROWS)
        m.synchronize()
        ms = []
        for _ in range(args.steps):
            m.flush_l2()
            m.synchronize()
            m.render(W, H, ZOOM, alias, S, STRIP_ROWS)
            m.synchronize()
            ms.append(max(m.step_ms(g) for g in range(G)))
        frame, _ = m.readback_rgb8(0)
        kern = [m.stats(g)["kernel_ms"] for g in range(G)]
        # end to end from host buffers, frames pipelined (upload H2D every step, async D2H into pinned memory)
        host = [pkg.HostBuffer(W * H * 3), pkg.HostBuffer(W * H * 3)]
        t0 = time.perf_counter()
        for k in range(args.steps):
            m.upload_scene(sph, lgt)
            m.render(W, H, ZOOM, alias, S, STRIP_ROWS)
            m.readback_wait(0)
            m.readback_rgb8_async(host[k & 1], 0)
        m.readback_wait(0)
        m.synchronize()
        e2e_s = time.perf_counter() - t0
        same_e2e = bool(np.array_equal(host[(args.steps - 1) & 1].array.reshape(H, W, 3), frame))
        for h in host:
            h.free()
    total = float(sum(ms))
    return {"value": rays * args.steps / (total * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_step": total / args.steps,
            "frames_per_s": args.steps * 1e3 / total, "steps": args.steps,
            "trace_kernel_ms_per_device": [round(k, 3) for k in kern],
            "frame_identical_to_1gpu": bool(np.array_equal(frame, single_frame)),
            "e2e": {"value": rays * args.steps / e2e_s / 1e6, "unit": "Mrays/s", "frames_per_s": args.steps / e2e_s,
                    "ms_per_step": e2e_s / args.steps * 1e3, "frame_identical": same_e2e},
            "path": "one process, rt_cuda_multi_init(G) = ncclCommInitAll + one stream per device; timed with CUDA events "
                    "on each device's stream, max over devices"}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
