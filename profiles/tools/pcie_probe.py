"""PCIe probe (context for bench.py's `e2e`): pinned-memory copy bandwidth H2D alone, D2H alone and both at once, at the e2e step's
transfer size (4 MiB) and at 64 MiB.  Alone:
    python profiles/tools/pcie_probe.py
With every GPU of the box copying AT THE SAME TIME (one process per GPU, a barrier before each timed section) — the box's
aggregate host ceiling, against which the 8-GPU end-to-end number has to be read:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 profiles/tools/pcie_probe.py
"""
import os
import time

import torch

rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def barrier():
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()


def run(nbytes, reps=50):
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d(); d2h()

    def wall(fn):
        fn(); barrier()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t) / reps
        if dist is not None:       # the slowest rank: all ranks copy during the same window
            t_ = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            dt = float(t_.item())
        return dt

    t1, t2, t3 = wall(h2d), wall(d2h), wall(both)
    if rank == 0:
        gb = lambda t: nbytes / t / 1e9
        print(f"{world} GPU(s) at once, {nbytes >> 20:4d} MiB per copy: H2D {gb(t1):6.1f} GB/s   D2H {gb(t2):6.1f} GB/s   both at once {gb(t3):6.1f} GB/s per direction "
              f"per GPU (slowest rank); aggregate both-at-once {world * gb(t3):6.1f} GB/s per direction", flush=True)


for n in (4 << 20, 64 << 20):
    run(n)
if dist is not None:
    dist.destroy_process_group()
