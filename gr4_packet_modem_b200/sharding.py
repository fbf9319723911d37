"""Host-side logic of the time-sharded multi-GPU mode.

The capture is cut on the reference's own FFT-block grid (block b starts at sample b*stride), each
rank gets a contiguous range of blocks plus `halo` extra blocks of input on each side.  The only
cross-rank dependency is the search offset of the peak walk at each shard boundary; it is resolved by
composing the (T+1)-entry chain tables the ranks exchange (b200sync_sd_shard_phase1/2)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Shard:
    rank: int
    first_block: int   # first FFT block this rank decides
    n_blocks: int
    total_blocks: int
    first_sample: int  # absolute index of the first input sample the rank must hold (halo included)
    n_samples: int     # number of input samples the rank must hold


def total_blocks(total_samples: int, fft_size: int, stride: int) -> int:
    """Blocks the reference loop runs: `for (j = 0; j + fft_size <= n; j += stride)`
    (PM/syncword_detection.hpp:238)."""
    return (total_samples - fft_size) // stride + 1 if total_samples >= fft_size else 0


def plan_shards(total_samples: int, world: int, fft_size: int, stride: int, time_threshold: int) -> list[Shard]:
    tb = total_blocks(total_samples, fft_size, stride)
    halo = (time_threshold + stride) // stride  # blocks of metric context needed on each side
    shards = []
    for r in range(world):
        fb = r * tb // world
        nb = (r + 1) * tb // world - fb
        cb0, cb1 = max(0, fb - halo), min(tb, fb + nb + halo)
        s0 = cb0 * stride
        shards.append(Shard(r, fb, nb, tb, s0, (cb1 - 1) * stride + fft_size - s0 if cb1 > cb0 else 0))
    return shards


def entry_offsets(tables: list[np.ndarray]) -> list[int]:
    """Entry search offset of every shard: offset 0 at the stream start, then each shard's table maps
    its entry offset to the next shard's."""
    j, out = 0, []
    for t in tables:
        out.append(j)
        j = int(t[j])
    return out


def gather_entry_offset(table: np.ndarray, rank: int, world: int, group=None, device=None) -> int:
    """all_gather the ranks' chain tables (T+1 small integers each — the path's only exchange) with
    torch.distributed and return this rank's entry offset.  The tables are host arrays and stay host bytes: with
    `device=None` (the default, and what bench.py does) the tensors are CPU tensors and travel over the gloo
    backend of the process group — nothing goes through a GPU or NCCL."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(table).astype(np.int32))
    if device is not None:
        t = t.to(device)
    allt = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(allt, t, group=group)
    return entry_offsets([a.cpu().numpy() for a in allt])[rank]
