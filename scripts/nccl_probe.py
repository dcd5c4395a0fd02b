"""Development aid: latency of the max all-reduce that follows the trace kernel (the shape of a multi-GPU step)."""
import os, sys, torch, torch.distributed as dist
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as graft
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
pkg = graft.load_package()
sph, lgt = pkg.synth_scene(1024, 4)
x = torch.zeros(1, dtype=torch.int32, device=dev)
stream = torch.cuda.Stream(device=dev)
r = pkg.Renderer(lr); r.set_stream(stream.cuda_stream); r.upload_scene(sph, lgt)
W, H = 3840, 2160
def run(reps, label, work, spp_alias=1.0):
    with torch.cuda.stream(stream):
        for _ in range(2):
            work(); dist.all_reduce(x, op=dist.ReduceOp.MAX)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        evs = []
        for _ in range(reps):
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record(stream); work(); b.record(stream)
            dist.all_reduce(x, op=dist.ReduceOp.MAX); c.record(stream)
            evs.append((a, b, c))
        torch.cuda.synchronize()
    busy_ms = sorted(a.elapsed_time(b) for a, b, c in evs)[reps // 2]
    red = sorted(b.elapsed_time(c) for a, b, c in evs)
    print(f"rank {rank} {label}: work {busy_ms:.2f} ms, all_reduce after it median {red[reps//2]*1e3:.0f} us min {red[0]*1e3:.0f} max {red[-1]*1e3:.0f}", flush=True)
run(6, "sleep kernel", lambda: torch.cuda._sleep(int(5e7)))
run(6, "render 1 spp (trace kernel only)", lambda: r.render_strips(W, H, -4.0, 1.0, 8, 4, rank, world))
run(6, "render 4 spp (trace + combine)", lambda: r.render_strips(W, H, -4.0, 2.0, 8, 4, rank, world))
run(6, "render 4 spp, 1/8 of the rows", lambda: r.render_strips(W, H, -4.0, 2.0, 8, 4, rank, 8))
dist.barrier(); r.close(); dist.destroy_process_group()
