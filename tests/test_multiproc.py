"""world-size-2 `gloo` tests (CPU) of the multi-process plumbing bench.py uses under torchrun: batch sharding
with no data-path collective, barrier-bracketed timing, MAX-over-ranks reduction, rank-0 aggregation."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import bench
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # every rank owns its own batch shard: different seeds, no exchange of activations
        lo, hi = bench.shard_range(10, world, rank)
        ms = bench.reduce_max_ms(10.0 + 5.0 * rank, torch.device('cpu'), world)      # slowest rank defines the step
        tokens = bench.aggregate_tokens(per_rank_batch=hi - lo, world=world, uniform=False)
        out[rank] = (lo, hi, ms, tokens)
    finally:
        dist.destroy_process_group()


def test_world2_sharding_and_max_reduction():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][:2] == (0, 5) and out[1][:2] == (5, 10)          # disjoint, covering shards
    assert out[0][2] == out[1][2] == 15.0                          # MAX over ranks on both
    assert out[0][3] == out[1][3] == 10 * 784                      # whole-job tokens per step


def test_shard_range_covers_everything():
    import bench
    for total in (1, 7, 8, 1000):
        for world in (1, 2, 3, 8):
            spans = [bench.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
