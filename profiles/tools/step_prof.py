"""A few k_step_smem<P> launches on 2^20 games for ncu.  python profiles/tools/step_prof.py [players]"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import rl_6_nimmt_b200
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4
envs = [BatchedSechsNimmtEnv(1 << 20, P, seed=5 + i).reset() for i in range(4)]
for t in range(3):
    for env in envs:
        env.step(env.random_actions().clone())
torch.cuda.synchronize()
print("ok")
