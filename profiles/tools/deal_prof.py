"""A few k_deal<P> launches on 2^20 games for ncu.  python profiles/tools/deal_prof.py [players]"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import rl_6_nimmt_b200
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4
env = BatchedSechsNimmtEnv(1 << 20, P, seed=5)
for _ in range(4):
    env.reset()
torch.cuda.synchronize()
print("ok")
