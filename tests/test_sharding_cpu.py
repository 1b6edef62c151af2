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
    # fine-tune step: two dense forwards + their backward (data gradients ~ forward GEMMs,
    # attention backward 2.5x its forward): a little under 4x the dense forward
    ft = bench.finetune_flops_per_image(16) / 1e9
    assert 3.9 * 35.6 < ft < 4.1 * 35.6


def _arena_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.join(ROOT, "dynamic-tuning_b200"))
    from dyt_b200.ddp import GradArena
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    lin[0].bias.requires_grad = False                      # frozen tensors stay out of the arena
    arena = GradArena(lin.parameters())
    assert len(arena.params) == 3 and arena.nbytes == (36 + 16 + 4) * 4
    x = torch.full((4, 7), float(rank + 1))
    for step in range(2):
        arena.zero()
        lin(x).sum().backward()                            # autograd accumulates into the arena views
        assert all(p.grad.data_ptr() == arena.flat.data_ptr() + 4 * o
                   for p, o in zip(arena.params, arena.offsets))
        local = [p.grad.clone() for p in arena.params]
        arena.all_reduce_mean()
    q.put((rank, [g.tolist() for g in local], [p.grad.tolist() for p in arena.params]))
    dist.destroy_process_group()


def test_two_rank_gradient_arena_allreduce():
    """BASELINE configs[2] exchange step: one flat arena, one all-reduce, mean
    over ranks; equals the average of the per-rank gradients."""
    world, port = 2, 29531
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_arena_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, a0), (_, l1, a1) = res
    assert a0 == a1                                         # both ranks hold the same averaged grads
    for g0, g1, avg in zip(l0, l1, a0):
        want = (torch.tensor(g0) + torch.tensor(g1)) / 2
        assert torch.allclose(torch.tensor(avg), want, rtol=1e-6, atol=1e-7)
