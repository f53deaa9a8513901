"""CUDA-event time of k_mcs_rollouts (BASELINE configs[2] shape): 256 four-player opening roots x 10 cards x 2000 rollouts per launch,
and one 4,096-root x 10,000-rollout decision batch.    python profiles/tools/mcs_time.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import rl_6_nimmt_b200
from rl_6_nimmt_b200 import rollouts as R
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv
P = 4


def roots_of(n, seed):
    obs = BatchedSechsNimmtEnv(n, P, seed=seed).reset().observe(dtype=torch.int8).cpu().numpy()
    return torch.as_tensor(np.stack([R.pack_root([[int(c) for c in row if c >= 0] for row in o[0, -24:].reshape(4, 6)], [int(c) for c in o[0, :10]],
                           [c for c in range(104) if c not in set(o[0, :10].tolist()) | set(o[0, -24:].tolist())], P) for o in obs])).cuda()


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for r in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


r256, r4096 = roots_of(256, 5), roots_of(4096, 6)
stats = torch.zeros((256, 10, 3), dtype=torch.int64, device="cuda")
ms = timed(lambda: R.mcs_rollouts(r256, P, 2000, seed=1, out=stats), 5)
print(f"256 roots x 10 x 2000: {ms:.3f} ms = {256 * 10 * 2000 / ms * 1e3:.3e} rollouts/s")
ms = timed(lambda: R.mcs_rollouts(r4096, P, 10000, seed=2), 2)
print(f"4096 roots x 10 x 10000: {ms:.3f} ms = {4096 * 10 * 10000 / ms * 1e3:.3e} rollouts/s")
ms = timed(lambda: R.mcs_rollouts(r4096[:1], P, 10000, seed=3), 20)
print(f"1 root x 10 x 10000: {ms * 1e3:.1f} us")
