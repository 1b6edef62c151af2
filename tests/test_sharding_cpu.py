"""world_size-2 gloo test of the N>1 host logic of bench.py on CPU: image sharding without any
data-path collective, barrier + max-over-ranks timing, rank-0-only reporting."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, ROOT)
    import bench
    w, r, l = bench.dist_setup(world)
    assert (w, r) == (world, rank) and dist.get_backend() == "gloo"
    start, count = bench.shard_images(513, world, rank)
    bench.barrier(world)
    t = bench.max_over_ranks(0.5 + rank, world, torch.device("cpu"))
    # the shards tile the global batch exactly, with no overlap
    owned = torch.zeros(513)
    owned[start:start + count] = 1
    dist.all_reduce(owned)
    q.put((rank, start, count, t, bool((owned == 1).all())))
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    world, port = 2, 29517
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1:3] == (0, 257) and res[1][1:3] == (257, 256)
    assert all(abs(r[3] - 1.5) < 1e-9 for r in res)      # max over ranks, seen by every rank
    assert all(r[4] for r in res)


def test_flops_model_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.flops_per_image(197) / 1e9 == pytest.approx(35.60, abs=0.02)   # dense ViT-B
    assert bench.flops_per_image(99) / 1e9 == pytest.approx(24.50, abs=0.02)    # r = 0.5
