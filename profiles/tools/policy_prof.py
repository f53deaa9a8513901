import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import rl_6_nimmt_b200
from rl_6_nimmt_b200 import policy as PL, rollouts as R
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv
P = 4
torch.manual_seed(0)
blob = PL.pack_weights(PL.PolicyNet())
env = BatchedSechsNimmtEnv(1 << 18, P, seed=5).reset()
obs = env.observe(dtype=torch.int8).reshape(-1, 47).contiguous()          # 2^20 decisions
for _ in range(2):
    PL.policy_probs(obs, blob)
o = obs.reshape(-1, P, 47)[:256].cpu().numpy()
roots = np.stack([R.pack_root([[int(c) for c in row if c >= 0] for row in x[0, -24:].reshape(4, 6)], [int(c) for c in x[0, :10]],
                  [c for c in range(104) if c not in set(x[0, :10].tolist()) | set(x[0, -24:].tolist())], P) for x in o])
for _ in range(2):
    R.policy_rollouts(roots, P, blob, 50, seed=1)
obs8 = torch.empty((1 << 18, P, 47), dtype=torch.int8, device="cuda")
obs32 = torch.empty((1 << 18, P, 47), dtype=torch.float32, device="cuda")
for _ in range(2):
    env.observe(out=obs8); env.observe(out=obs32)
torch.cuda.synchronize()
print("ok")
