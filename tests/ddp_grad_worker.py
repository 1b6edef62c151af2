"""torchrun worker of tests/test_ddp_gpu.py (2 GPUs): SURVEY.md section 8e's invariant of the
fine-tune path -- the gradients of ONE GPU on the full batch (loss = mean over the shards of the
per-shard loss: the reference's keep-rate term is evaluated per DDP replica, models/losses.py:69-72)
equal the all-reduced (averaged) gradients of N GPUs that each hold one shard.  Same weights, same
Gumbel draws and dropout multipliers (sliced per shard)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "dynamic-tuning_b200"), os.path.join(ROOT, "oracle")]


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from dyt_b200 import synthetic, train
    from dyt_b200.ddp import GradArena, trainable_parameters
    from dyt_b200.finetune import finetune_loss

    torch.manual_seed(0)
    model = synthetic.build_vit_b16(dev, flavour="train", ffn_num=16, scalar="1.0", seed=0)
    params = trainable_parameters(model)
    model.train()
    arena = GradArena(params)
    per = 4
    B = per * world
    g = torch.Generator().manual_seed(1)
    img = torch.randn(B, 3, 224, 224, generator=g).to(dev)
    tgt = torch.randint(0, 100, (B,), generator=g).to(dev)
    L = len(model.blocks)
    bott = model.blocks[0].adaptmlp.down_proj.out_features
    noises = [(-torch.empty(B, 196, 1).exponential_(generator=g).log().half().float().to(dev),
               -torch.empty(B, 196, 1).exponential_(generator=g).log().half().float().to(dev))
              for _ in range(2 * L)]
    drops = [((torch.rand(B, 197, bott, generator=g) >= 0.1).float() / 0.9).half().to(dev)
             for _ in range(2 * L)]

    def grads(sl):
        """scaled-loss gradients of the image slice sl, per-shard loss averaged over the shards in it"""
        arena.zero()
        n_sh = (sl.stop - sl.start) // per
        ns = [(a[sl], b[sl]) for a, b in noises]
        ds = [d[sl] for d in drops]
        with train.fixed_randomness(ns, ds):
            with torch.autocast("cuda", dtype=torch.float16):
                out_s, ts = model(img[sl])
                out_t, _ = model(img[sl], complete_model=True)
        loss = 0.0
        for i in range(n_sh):
            q = slice(i * per, (i + 1) * per)
            loss = loss + finetune_loss(out_s[q].float(), ts["token_select"][q].float(), out_t[q].float(),
                                        tgt[sl][q]) / n_sh
        (loss * 128.0).backward()
        return loss.detach()

    # data-parallel: every rank its own shard, one all-reduce of the arena
    loss_dp = grads(slice(rank * per, (rank + 1) * per))
    arena.all_reduce_mean()
    g_dp = arena.flat.clone()
    # one GPU, full batch
    loss_full = grads(slice(0, B))
    g_full = arena.flat.clone()
    torch.cuda.synchronize()
    rel = ((g_dp - g_full).norm() / g_full.norm()).item()
    mx = ((g_dp - g_full).abs().max() / g_full.abs().max()).item()
    t = torch.tensor([loss_dp.item()], device=dev)
    dist.all_reduce(t)
    ok = rel <= 2e-3 and mx <= 5e-3 and abs(t.item() / world - loss_full.item()) <= 1e-3 * abs(loss_full.item())
    print(f"rank {rank}: rel_l2 {rel:.2e} max {mx:.2e} loss_dp_mean {t.item() / world:.5f} "
          f"loss_full {loss_full.item():.5f} nonzero_grad {float(g_full.abs().max()):.3e} ok {ok}", flush=True)
    assert float(g_full.abs().max()) > 0
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
