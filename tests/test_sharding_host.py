"""CPU tests of the multi-GPU host logic: shard planning, chain-table composition, and the
world_size-2 exchange over torch.distributed (gloo)."""
import os
import socket

import numpy as np
import pytest

from gr4_packet_modem_b200.sharding import entry_offsets, plan_shards, total_blocks


def piece_tables(cand: np.ndarray, lo: int, hi: int, T: int) -> np.ndarray:
    """numpy reference of the per-range chain table: entry offset j in [0,T] at `lo`
    -> exit offset at `hi`, for the candidate bitmap `cand`."""
    Fr = T + 1
    table = np.arange(Fr)
    pos = lo
    while pos < hi:
        L = min(Fr, hi - pos)
        bits = cand[pos:pos + L]
        nxt = np.full(L + 1, -1)
        for a in range(L - 1, -1, -1):
            nxt[a] = a if bits[a] else nxt[a + 1]
        new = np.empty(Fr, dtype=np.int64)
        for j in range(Fr):
            x = table[j]
            if x >= L:
                new[j] = x - L
            else:
                a = nxt[x]
                new[j] = a + Fr - L if a >= 0 else 0
        table = new
        pos += L
    return table


def sequential_examined(cand: np.ndarray, hi: int, T: int) -> list[int]:
    r, out = 0, []
    while True:
        nz = np.flatnonzero(cand[r:hi])
        if nz.size == 0:
            return out
        p = r + int(nz[0])
        out.append(p)
        r = p + T + 1


def walk_range(cand, lo, hi, T, j):
    """examined peaks of one shard given its entry offset"""
    r, out = lo + j, []
    while r < hi:
        nz = np.flatnonzero(cand[r:hi])
        if nz.size == 0:
            break
        p = r + int(nz[0])
        out.append(p)
        r = p + T + 1
    return out


def test_plan_covers_all_blocks():
    for total, world in [(1 << 20, 1), (1 << 20, 3), (123457, 8), (4000, 2)]:
        sh = plan_shards(total, world, 2048, 1752, 768)
        tb = total_blocks(total, 2048, 1752)
        assert sum(s.n_blocks for s in sh) == tb
        assert sh[0].first_block == 0 and sh[-1].first_block + sh[-1].n_blocks == tb
        for s in sh:
            assert s.first_sample >= 0 and s.first_sample + s.n_samples <= total
            # halo: one block of 1752 covers T = 768 of context on each side
            assert s.first_sample <= max(0, s.first_block - 1) * 1752


@pytest.mark.parametrize("T,density", [(5, 0.3), (20, 0.02), (20, 1.0), (63, 0.0)])
def test_table_composition_equals_sequential_walk(T, density):
    rng = np.random.default_rng(T)
    n = 5000
    cand = rng.random(n) < density
    truth = sequential_examined(cand, n, T)
    for cuts in ([0, n], [0, 1234, n], [0, 7, 8, 2500, 2501, 4999, n]):
        tables = [piece_tables(cand, a, b, T) for a, b in zip(cuts[:-1], cuts[1:])]
        offs = entry_offsets(tables)
        got = []
        for (a, b), j in zip(zip(cuts[:-1], cuts[1:]), offs):
            got += walk_range(cand, a, b, T, j)
        assert got == truth


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from gr4_packet_modem_b200.sharding import gather_entry_offset

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    T = 31
    cand = rng.random(4000) < 0.05
    cuts = [0, 2100, 4000]
    table = piece_tables(cand, cuts[rank], cuts[rank + 1], T).astype(np.uint16)
    j = gather_entry_offset(table, rank, world)
    q.put((rank, j, walk_range(cand, cuts[rank], cuts[rank + 1], T, j)))
    dist.destroy_process_group()


def test_gloo_world2_exchange():
    """Two processes, gloo backend: each builds its shard's table, all_gathers, composes its entry
    offset; the union of the shard walks equals the single-process walk."""
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(1)
    cand = rng.random(4000) < 0.05
    assert res[0][1] == 0
    assert res[0][2] + res[1][2] == sequential_examined(cand, 4000, 31)


def _xchg_worker(rank, world, name, q):
    import time

    from gr4_packet_modem_b200.sharding import SharedTableExchange, entry_offsets

    ex = SharedTableExchange(name, rank, world, 769)
    rng = np.random.default_rng(5)
    for step in range(200):
        tabs = [rng.integers(0, 769, 769).astype(np.uint16) for _ in range(world)]
        if rank == world - 1 and step % 7 == 0:
            time.sleep(0.001)  # a slow reader: earlier ranks run ahead
        j = ex.entry_offset(tabs[rank])
        if j != entry_offsets(tabs)[rank]:
            q.put((rank, step, "mismatch"))
            return
    time.sleep(0.1)
    ex.close()
    q.put((rank, 200, "ok"))


def test_shared_memory_table_exchange_three_ranks():
    """The same-box exchange bench.py uses instead of a collective: tables as host bytes in a shared segment,
    double-buffered, ranks free to run ahead — every rank must get the composition of the shards before it."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    name = f"b200sync_test_{os.getpid()}"
    ps = [ctx.Process(target=_xchg_worker, args=(r, 3, name, q)) for r in range(3)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(timeout=120)
    res = sorted(q.get(timeout=5) for _ in range(3))
    assert [r[2] for r in res] == ["ok"] * 3, res
