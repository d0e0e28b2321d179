"""CPU: the N>1 path — LPT sharding and the variable-length waveform all-gather — on 2 gloo ranks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from auralis_b200.parallel import gather_waveforms, lpt_partition, run_sharded


def test_lpt_partition_balances_and_is_deterministic():
    costs = [100, 90, 10, 10, 10, 80, 5, 5]
    p = lpt_partition(costs, 3)
    assert sorted(i for lst in p for i in lst) == list(range(8))
    loads = [sum(costs[i] for i in lst) for lst in p]
    assert max(loads) - min(loads) <= 10
    assert p == lpt_partition(costs, 3)
    assert lpt_partition([1, 2, 3], 1) == [[0, 1, 2]]
    assert lpt_partition([], 2) == [[], []]


def _wave(i):
    rng = np.random.RandomState(i)
    return rng.randn(100 + 37 * i).astype(np.float32)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 7
        costs = [100 + 37 * i for i in range(n)]
        seen = []

        def synth(idx):
            seen.extend(idx)
            return {i: _wave(i) for i in idx}
        out = run_sharded(list(range(n)), costs, synth)
        ok = all(np.array_equal(out[i], _wave(i)) for i in range(n))
        # ragged edge: one rank has nothing to contribute
        out2 = gather_waveforms({0: _wave(0)} if rank == 0 else {}, 1)
        ok = ok and np.array_equal(out2[0], _wave(0))
        q.put((rank, ok, sorted(seen)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gather():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    shares = {r: s for r, _, s in res}
    assert sorted(shares[0] + shares[1]) == list(range(7)) and shares[0] and shares[1]
