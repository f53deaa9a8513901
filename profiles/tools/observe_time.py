"""Times k_observe for int8 / fp32 on 2^20 games.  python profiles/tools/observe_time.py [players]"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import rl_6_nimmt_b200
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = 1 << 20
envs = [BatchedSechsNimmtEnv(B, P, seed=5 + i).reset() for i in range(3)]
for e in envs:
    for t in range(3):
        e.step(e.random_actions().clone())
for dt in (torch.int8, torch.float32):
    outs = [torch.empty((B, P, 47), dtype=dt, device="cuda") for _ in range(3)]
    for e, o in zip(envs, outs):
        e.observe(out=o)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for r in range(10):
        for e, o in zip(envs, outs):
            e.observe(out=o)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 30
    nbytes = (12 * P + 24 + 47 * P * outs[0].element_size()) * B
    print(f"k_observe<{P}> {dt}: {ms * 1e3:.1f} us, {nbytes / ms / 1e6:.0f} GB/s")
    del outs
