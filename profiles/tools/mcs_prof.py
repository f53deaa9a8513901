import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import rl_6_nimmt_b200
from rl_6_nimmt_b200 import rollouts as R
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv
P=4
obs0 = BatchedSechsNimmtEnv(256, P, seed=5).reset().observe(dtype=torch.int8).cpu().numpy()
roots = np.stack([R.pack_root([[int(c) for c in row if c >= 0] for row in o[0, -24:].reshape(4, 6)], [int(c) for c in o[0, :10]],
                  [c for c in range(104) if c not in set(o[0, :10].tolist()) | set(o[0, -24:].tolist())], P) for o in obs0])
roots_d = torch.as_tensor(roots).cuda()
stats = torch.zeros((256, 10, 3), dtype=torch.int64, device="cuda")
for r in range(4):
    R.mcs_rollouts(roots_d, P, 2000, seed=r, out=stats)
torch.cuda.synchronize()
print("ok", int(stats[:, :, 2].sum()))
