"""One launch of each auxiliary env kernel on 2^20 four-player games, for ncu.  python profiles/tools/misc_prof.py"""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import rl_6_nimmt_b200
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv
from rl_6_nimmt_b200.play import BatchedGameSession, MCSSeat, RandomSeat
P = 4
env = BatchedSechsNimmtEnv(1 << 20, P, seed=5)
for _ in range(2):
    env.reset()
    a = env.random_actions().clone()
    env.step(a)
    env.observe(dtype=torch.int8)
    env.observe(dtype=torch.float32)
    env.scores()
    env.step_random()
sess = BatchedGameSession([MCSSeat(rollouts_per_card=4), RandomSeat(), RandomSeat(), RandomSeat()], 1 << 16, seed=1)
sess.play_games()
torch.cuda.synchronize()
print("ok")
