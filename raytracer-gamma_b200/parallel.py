"""Row-strip sharding of one frame over the GPUs of a box.

The reference is single-device (one cl_command_queue, main.cpp:229); pixels are
independent (raytrace_kernel.cl:884-972 touches only dst[gid]) and the scene is tiny, so
the frame shards with no data-path exchange.  Rank g renders the strips
{k : k mod G == g} of `strip_rows` rows (interleaved for load balance).  One exchange
step per frame remains because the image is normalised by its GLOBAL maximum
(algebra.h:68-91): a MAX all-reduce of one word, then an all-gather of the quantised
RGB8 strips and a de-interleave (rt_cuda_assemble_rgb8 on the GPU).

Everything here is host-side plumbing over torch.distributed (NCCL on GPUs, gloo in the
CPU tests); the rendering itself is the C-ABI library.
"""
from __future__ import annotations

import numpy as np


def shard_rows(height: int, strip_rows: int, rank: int, world: int) -> np.ndarray:
    """Global row numbers owned by `rank`, in storage (increasing) order."""
    rows = np.arange(height)
    return rows[(rows // strip_rows) % world == rank]


def shard_layout(height: int, width: int, strip_rows: int, world: int) -> dict:
    counts = [len(shard_rows(height, strip_rows, g, world)) for g in range(world)]
    max_rows = max(counts)
    pitch = ((max_rows * width * 3 + 15) // 16) * 16      # bytes per shard block, 16-byte aligned
    return {"rows": counts, "max_rows": max_rows, "pitch": pitch}


def local_row_of(row: int, strip_rows: int, world: int) -> tuple[int, int]:
    """(owning rank, row index inside that rank's packed buffer) of a global row —
    the same arithmetic as assemble_rgb8_kernel in csrc/rt_kernels.cuh."""
    strip = row // strip_rows
    return strip % world, (strip // world) * strip_rows + (row - strip * strip_rows)


def assemble_host(gathered: np.ndarray, height: int, width: int, strip_rows: int, world: int,
                  pitch: int) -> np.ndarray:
    """NumPy statement of the strip de-interleave (for tests and for hosts without a GPU step)."""
    out = np.empty((height, width, 3), np.uint8)
    flat = np.ascontiguousarray(gathered).reshape(-1)
    row_bytes = width * 3
    for row in range(height):
        g, lr = local_row_of(row, strip_rows, world)
        off = g * pitch + lr * row_bytes
        out[row] = flat[off:off + row_bytes].reshape(width, 3)
    return out


class StripExchange:
    """The per-frame exchange: max all-reduce + RGB8 all-gather, on whatever backend the
    process group uses.  Tensors are torch tensors on the group's device."""

    def __init__(self, dist, torch, height, width, strip_rows, rank, world, device):
        self.dist, self.torch = dist, torch
        self.H, self.W, self.strip_rows, self.rank, self.world = height, width, strip_rows, rank, world
        lay = shard_layout(height, width, strip_rows, world)
        self.pitch, self.my_rows = lay["pitch"], lay["rows"][rank]
        self.send = torch.zeros(self.pitch, dtype=torch.uint8, device=device)
        self.gathered = torch.empty(world * self.pitch, dtype=torch.uint8, device=device)

    def reduce_max(self, max_bits_i32):
        """In-place MAX of the per-shard maxima.  The maxima are non-negative floats, whose
        IEEE bit patterns order like int32, so the reduce runs on the raw bits."""
        self.dist.all_reduce(max_bits_i32, op=self.dist.ReduceOp.MAX)
        return max_bits_i32

    def gather(self, rgb_local_u8):
        n = self.my_rows * self.W * 3
        self.send[:n].copy_(rgb_local_u8[:n], non_blocking=True)
        self.dist.all_gather_into_tensor(self.gathered, self.send)
        return self.gathered
