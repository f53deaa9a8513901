"""Self-play iteration time (four learning PUCT seats on one net) against the number of games:
    python profiles/tools/selfplay_time.py [games ...]"""
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


for games in [int(x) for x in sys.argv[1:]] or [128, 192, 222, 256, 384, 444]:
    torch.manual_seed(0)
    net = PolicyNet()
    session = BatchedGameSession([PolicySeat(net, mc_max=200, puct=True, learn=True) for _ in range(4)], games, device="cuda", seed=11)
    session.play_games()
    ms = statistics.median(event_ms(session.play_games) for _ in range(3))
    print(f"{games} games: {ms:.1f} ms per iteration = {games / ms * 1e3:.0f} games/s, {4 * ((games + 2) // 3)} search CTAs per turn")
