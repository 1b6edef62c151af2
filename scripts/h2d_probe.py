"""GPU probe: host -> device bandwidth of the e2e loop's 154 MB image batch from pinned memory: one
cudaMemcpyAsync against the same bytes split over 2 / 4 streams."""
import time, torch
dev = torch.device("cuda:0")
host = torch.randn(256, 3, 224, 224).pin_memory()
dst = torch.empty_like(host, device=dev)
nbytes = host.numel() * 4
def run(parts, n=10):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    hs, ds = host.view(-1).chunk(parts), dst.view(-1).chunk(parts)
    def once():
        for s, h, d in zip(streams, hs, ds):
            with torch.cuda.stream(s):
                d.copy_(h, non_blocking=True)
    once(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): once()
    torch.cuda.synchronize()
    return nbytes * n / (time.perf_counter() - t0) / 1e9
for rnd in range(2):
    print("  ".join(f"{p} stream(s): {run(p):.1f} GB/s" for p in (1, 2, 4)), flush=True)
