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


class SharedTableExchange:
    """The same exchange for ranks on ONE box without any network stack: every rank writes its (T+1)-entry table into
    its slot of a POSIX shared-memory segment and bumps its sequence counter; a rank spins until the ranks in front of
    it have published the step and composes their tables.  ~10 us instead of the ~300 us of a loopback all_gather; the
    tables never leave host memory (north_star: detections and tables are gathered on the host, no NCCL).
    Tables are double-buffered by step parity and a rank does not overwrite a slot before every later rank has
    finished reading the step that used it, so ranks may run ahead of each other freely.
    Rank 0 creates the segment, the others retry until it exists."""

    def __init__(self, name: str, rank: int, world: int, table_len: int):
        import time
        from multiprocessing import shared_memory

        self.rank, self.world, self.n = rank, world, table_len
        tab_bytes = (4 * table_len + 63) // 64 * 64
        self._stride = 64 + 2 * tab_bytes                   # [seq int64 | done int64 | pad] + two tables
        size = self._stride * world
        if rank == 0:
            try:
                old = shared_memory.SharedMemory(name=name)
                old.close()
                old.unlink()
            except FileNotFoundError:
                pass
            self._shm = shared_memory.SharedMemory(name=name, create=True, size=size)
            self._shm.buf[:size] = bytes(size)
        else:
            deadline = time.time() + 60
            while True:
                try:
                    self._shm = shared_memory.SharedMemory(name=name)
                    try:  # Python < 3.13 registers attached segments with the resource tracker too: only rank 0 owns it
                        from multiprocessing import resource_tracker

                        resource_tracker.unregister(self._shm._name, "shared_memory")
                    except Exception:
                        pass
                    if self._shm.size >= size:
                        break
                    self._shm.close()
                except FileNotFoundError:
                    pass
                if time.time() > deadline:
                    raise TimeoutError(f"shared table segment {name} did not appear")
                time.sleep(0.01)
        buf = self._shm.buf
        self._seq = [np.ndarray((1,), np.int64, buf, r * self._stride) for r in range(world)]
        self._done = [np.ndarray((1,), np.int64, buf, r * self._stride + 8) for r in range(world)]
        self._tab = [[np.ndarray((table_len,), np.int32, buf, r * self._stride + 64 + par * tab_bytes) for par in (0, 1)]
                     for r in range(world)]
        self._step = 0

    def entry_offset(self, table: np.ndarray) -> int:
        """Publish this rank's table for the next step, wait for the tables of the ranks in front, return this
        rank's entry offset (0 at the stream start, then through every earlier shard's table)."""
        self._step += 1
        s, par = self._step, self._step & 1
        for r in range(self.rank + 1, self.world):     # the slot was last used by step s-2: its readers must be done
            d = self._done[r]
            while d[0] < s - 2:
                pass
        self._tab[self.rank][par][:] = table
        self._seq[self.rank][0] = s                    # x86 keeps the two stores in order
        j = 0
        for r in range(self.rank):
            q = self._seq[r]
            while q[0] < s:
                pass
            j = int(self._tab[r][par][j])
        self._done[self.rank][0] = s
        return j

    def close(self):
        self._seq = self._done = self._tab = None
        try:
            self._shm.close()
            if self.rank == 0:
                self._shm.unlink()
        except Exception:
            pass
