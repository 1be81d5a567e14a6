"""world_size-2 `gloo` test of the cross-rank best-sample reduction (ambersim_b200/parallel.py):
sharding ranges, first-minimum / NaN / tie-break semantics, identical winners on every rank."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ambersim_b200.parallel import first_min_index, merge_best, shard_range
from tests._philox import philox4x32


def test_shard_ranges_cover_everything():
    for S in (1, 7, 100, 4096, 1_000_003):
        for R in (1, 2, 4, 8):
            r = [shard_range(S, k, R) for k in range(R)]
            assert r[0][0] == 0 and r[-1][1] == S
            assert all(r[k][1] == r[k + 1][0] for k in range(R - 1))


def test_first_min_semantics():
    nan = float("nan")
    costs = torch.tensor([[3.0, 1.0, nan, 2.0], [2.0, 1.0, 5.0, nan], [2.0, 4.0, nan, 2.0]])
    idx = torch.tensor([[5, 9, 40, 7], [3, 2, 1, 30], [1, 0, 20, 2]])
    # col0: min 2.0 tie -> idx 1 (rank 2); col1: tie at 1.0 -> idx 2 (rank 1); col2: NaN, lowest idx 20 (rank 2);
    # col3: NaN wins (rank 1)
    assert first_min_index(costs, idx).tolist() == [2, 1, 2, 1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, N, nx, nu = 3, 4, 5, 2
    g = torch.Generator().manual_seed(100 + rank)
    cost = torch.tensor([[1.0, 5.0, float("nan")], [0.5, 5.0, 2.0]])[rank]
    # column 1 ties on cost with global ids above 2^24 that differ by one: exact only if the id travels as int32 bits
    idx = torch.tensor([[3, 16777218, 8], [40, 16777217, 41]], dtype=torch.int32)[rank]
    xs, us = torch.randn((B, N + 1, nx), generator=g), torch.randn((B, N, nu), generator=g)
    out = merge_best(cost, idx, xs, us)
    q.put((rank, [t.numpy() for t in out], xs.numpy(), us.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_merge_best_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    (r0, o0, xs0, us0), (r1, o1, xs1, us1) = res
    for a, b in zip(o0, o1):
        assert np.array_equal(a, b, equal_nan=True)  # identical winners on both ranks
    xs, us, idx, cost = o0
    assert idx.tolist() == [40, 16777217, 8]  # lower cost; tie -> lower global index (not representable in float32); NaN counts as minimal
    assert np.array_equal(xs[0], xs1[0]) and np.array_equal(xs[1], xs1[1]) and np.array_equal(xs[2], xs0[2])
    assert np.array_equal(us[0], us1[0]) and np.array_equal(us[1], us1[1])


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 pin the numpy re-statement of the engine's
    generator (the GPU tests then compare device noise against it)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))]
    for ctr, key, exp in kat:
        assert tuple(int(x) for x in philox4x32(*ctr, *key)) == exp
