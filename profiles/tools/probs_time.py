"""Times k_policy_probs alone (2^20 decisions, CUDA events): python profiles/tools/probs_time.py [label]
The library is the product one, or NIMMT_B200_LIB (experiment builds)."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import rl_6_nimmt_b200
from rl_6_nimmt_b200 import policy as PL
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv

torch.manual_seed(0)
blob = PL.pack_weights(PL.PolicyNet())
env = BatchedSechsNimmtEnv(1 << 18, 4, seed=5).reset()
obs = env.observe(dtype=torch.int8).reshape(-1, 47).contiguous()
for _ in range(3):
    PL.policy_probs(obs, blob)
torch.cuda.synchronize()
ts = []
for rep in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        PL.policy_probs(obs, blob)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) / 10)
ms = sorted(ts)[len(ts) // 2]
D = obs.shape[0]
print(f"{sys.argv[1] if len(sys.argv) > 1 else 'k_policy_probs'}: {ms:.4f} ms per {D} decisions = {D / ms * 1e3:.3e} decisions/s, "
      f"{D * 10 * 29800 / ms * 1e3 / 1e12:.0f} TFLOP/s (un-padded)")
