"""Developer probe: per-rank pinned-host <-> device copy bandwidth, alone and with every rank copying at once.
Run under torchrun (or alone).  Explains the e2e / resident ratio of bench.py at N > 1."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    t = torch.zeros(1, device="cuda")
    dist.all_reduce(t)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for mb in (16, 128):
    n = mb * (1 << 20) // 8
    h = torch.ones(n, dtype=torch.float64).pin_memory()
    h2 = torch.ones(n, dtype=torch.float64).pin_memory()
    d = torch.zeros(n, dtype=torch.float64, device="cuda")
    d2 = torch.zeros(n, dtype=torch.float64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(kind, reps=10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if kind in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d.copy_(h, non_blocking=True)
            if kind in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    for kind in ("h2d", "d2h", "both"):
        run(kind, 2)
        # alone: ranks take turns
        alone = None
        for r in range(world):
            barrier()
            if r == rank:
                alone = run(kind)
            barrier()
        barrier()
        together = run(kind)
        barrier()
        gbs = lambda s: mb / 1024 / s  # noqa: E731
        print(f"rank {rank} {mb:4d} MiB {kind:5s}: alone {alone * 1e3:7.3f} ms ({gbs(alone):5.1f} GiB/s per direction)   "
              f"all {world} ranks at once {together * 1e3:7.3f} ms ({gbs(together):5.1f} GiB/s)", flush=True)
if world > 1:
    dist.destroy_process_group()
