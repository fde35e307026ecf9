"""The multi-GPU path on CPU: two gloo ranks, each owning a block of agents
(bidinet_b200/sharding.py), against one process owning all of them.  The oracle stands in
for the engine; what is under test is the host logic bench.py uses at N > 1 -- agent
blocks, seeds, per-rank input frames, the max-over-ranks reduction -- and the property it
relies on: agents never interact, so sharding changes nothing."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bidinet_b200 import config as cf  # noqa: E402
from bidinet_b200 import sharding  # noqa: E402

AGENTS_PER_RANK, STEPS = 3, 4


def _run_block(first, count):
    from oracle.pyoracle import OracleHierarchy
    cfg, ld = cf.config_c2()
    frames = sharding.base_frames(STEPS, count, first)
    out = np.zeros((count, 64), dtype=np.float32)
    for a in range(count):
        h = OracleHierarchy(cfg, ld, sharding.SEED_BASE + first + a)
        h.run_runner(frames[:, a, :], STEPS, sharding.FB_IDX, sharding.FB_IDX)
        out[a] = h.get_predictions()
    return out


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert sharding.rank_info() == (rank, world, rank)
    first, count = sharding.agent_block(rank, AGENTS_PER_RANK)
    assert sharding.seed_base_for(rank, AGENTS_PER_RANK) == sharding.SEED_BASE + first
    pred = _run_block(first, count)
    slowest = sharding.max_over_ranks(1.0 + rank, dist)   # rank r "took" 1+r seconds
    dist.barrier()
    q.put((rank, pred, slowest))
    dist.destroy_process_group()


def test_two_ranks_equal_one_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=300) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = _run_block(0, 2 * AGENTS_PER_RANK)
    sharded = np.concatenate([g[1] for g in got])
    assert np.array_equal(whole, sharded)
    assert [g[2] for g in got] == [2.0, 2.0]          # every rank sees the slowest rank's time
    assert not np.array_equal(whole[0], whole[AGENTS_PER_RANK])  # different agents do differ
    assert sharding.whole_job_rate(AGENTS_PER_RANK, 2, STEPS, 2.0) == 2 * AGENTS_PER_RANK * STEPS / 2.0


def test_frames_depend_on_global_agent_index_only():
    a = sharding.base_frames(5, 4, 0)
    b = sharding.base_frames(5, 2, 2)
    assert np.array_equal(a[:, 2:], b)
    assert np.all(a[:, :, sharding.FB_IDX] == 0) and a.min() >= 0 and a.max() <= 1
