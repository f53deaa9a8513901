"""PCIe probe (context for bench.py's `e2e`): pinned-memory copy bandwidth H2D alone, D2H alone and both at once,
at the e2e step's transfer size (4 MiB) and at 64 MiB.  python profiles/tools/pcie_probe.py"""
import torch

def run(nbytes, reps=50):
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d_out = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def timed(fn):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        s1.synchronize(); s2.synchronize()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e-3
    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    def both():
        h2d(); d2h()
    import time
    def wall(fn):
        fn(); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) / reps
    t1, t2, t3 = wall(h2d), wall(d2h), wall(both)
    print(f"{nbytes >> 20:4d} MiB: H2D {nbytes / t1 / 1e9:6.1f} GB/s   D2H {nbytes / t2 / 1e9:6.1f} GB/s   both at once {nbytes / t3 / 1e9:6.1f} GB/s per direction")

for n in (4 << 20, 64 << 20):
    run(n)
