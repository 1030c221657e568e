"""Developer probe: aggregate pinned host <-> device copy bandwidth with all ranks copying at once (torchrun)."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = 2 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s2 = torch.cuda.Stream()
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True)),
                 ("both", None)):
    for rep in range(2):
        if world > 1: dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if fn: fn()
        else:
            d.copy_(h, non_blocking=True)
            with torch.cuda.stream(s2):
                h2 = h  # same buffer direction d2h on another stream
                torch.empty(0)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        dt = time.perf_counter() - t0
    if fn:
        t = torch.tensor([n / dt / 1e9], device="cuda")
        if world > 1: dist.all_reduce(t)
        if rank == 0: print(f"{name}: {n/dt/1e9:.1f} GB/s on rank 0, {t.item():.1f} GB/s aggregate over {world} ranks", flush=True)
if world > 1: dist.destroy_process_group()
