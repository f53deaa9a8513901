import os, sys, statistics
sys.path.insert(0, os.getcwd())
import torch
import rl_6_nimmt_b200
from rl_6_nimmt_b200.play import BatchedGameSession, PolicySeat
from rl_6_nimmt_b200.policy import PolicyNet

def event_ms(fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)

import time
for name, kw in (("full", dict(mc_max=200, learn=True)), ("no learning", dict(mc_max=200, learn=False)), ("mc_max=1, learning", dict(mc_max=1, learn=True)), ("mc_max=1, no learning", dict(mc_max=1, learn=False))):
    torch.manual_seed(0)
    net = PolicyNet()
    session = BatchedGameSession([PolicySeat(net, puct=True, **kw) for _ in range(4)], 256, device="cuda", seed=11)
    session.play_games()
    ms = statistics.median(event_ms(session.play_games) for _ in range(3))
    t0 = time.perf_counter(); session.play_games(); t1 = time.perf_counter()   # host time to ENQUEUE one iteration (no sync inside?)
    torch.cuda.synchronize()
    print(f"{name}: {ms:.1f} ms per iteration (host enqueue {1e3 * (t1 - t0):.1f} ms)")
