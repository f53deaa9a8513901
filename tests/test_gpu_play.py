"""GPU tests of the batched game session (many games x mixed agents in lock-step)."""
import numpy as np
import pytest
import torch

import rl_6_nimmt_b200  # noqa: F401
from rl_6_nimmt_b200 import _native as N
from rl_6_nimmt_b200 import rollouts as R
from rl_6_nimmt_b200.agents import MCSAgent
from rl_6_nimmt_b200.play import BatchedGameSession, MCSSeat, PolicySeat, RandomSeat
from rl_6_nimmt_b200 import policy as PL

pytestmark = pytest.mark.gpu


def test_device_bookkeeping_equals_the_host_agent():
    """k_mc_roots / k_mc_choose against BaseMCAgent's own host code on the same trajectories: available-card
    memory (stale semantics included), root bytes and the decision rule, turn by turn."""
    B, P, seat = 64, 3, 1
    sess = BatchedGameSession([RandomSeat(), MCSSeat(), RandomSeat()], B, seed=4)
    env, lib = sess.env, N.lib()
    env.reset()
    agents = [MCSAgent(seed=1) for _ in range(B)]
    avail = torch.zeros((B, 16), dtype=torch.uint8, device="cuda")
    roots = torch.zeros((B, 64), dtype=torch.uint8, device="cuda")
    for turn in range(10):
        obs = env.observe(dtype=torch.int64).cpu().numpy()
        N.check(lib.nimmt_mc_roots(N.ptr(env.state), N.ptr(avail), N.ptr(roots), B, P, seat, int(turn == 0), 0), "roots")
        got = roots.cpu().numpy()
        for b in range(B):
            st = obs[b, seat].astype(np.float32)
            legal = [int(c) for c in obs[b, seat, :10] if c >= 0]
            ag = agents[b]
            if len(legal) == ag.handsize:
                ag._initialize_game(st)
            ag._memorize_cards(st, legal)
            want = R.pack_root_from_state(st, legal, ag.available_cards)
            assert (got[b] == want).all(), (turn, b)
        # decision rule on synthetic stats (ties, unvisited cards) vs the host rule
        n = 10 - turn
        rng = np.random.RandomState(turn)
        stats = np.zeros((B, 10, 3), np.int64)
        stats[:, :n, 2] = rng.randint(0, 4, size=(B, n))
        stats[:, :n, 0] = -rng.randint(0, 6, size=(B, n)) * stats[:, :n, 2]
        stats[stats[:, :n, 2].sum(axis=1) == 0, 0, 2] = 1
        acts = torch.zeros((B, P), dtype=torch.uint8, device="cuda")
        N.check(lib.nimmt_mc_choose(N.ptr(env.state), N.ptr(torch.from_numpy(stats).cuda()), N.ptr(acts), B, P, seat, 0), "choose")
        acts = acts.cpu().numpy()
        for b in range(B):
            legal = [int(c) for c in obs[b, seat, :10] if c >= 0]
            want = legal[0] if n == 1 else R.choose_from_stats(legal, stats[b, :n])[0]
            assert acts[b, seat] == want, (turn, b)
        env.step_random()


def test_mcs_beats_random_like_the_readme_says():
    """4096 four-player games, MCSAgent(mc_max=200) at seat 0 against three DrunkHamsters.  The reference's
    tournament table (README.md:32-38) has MCS at -8.06 mean score and Random at -13.49."""
    sess = BatchedGameSession([MCSSeat(mc_per_card=10, mc_max=200), RandomSeat(), RandomSeat(), RandomSeat()], 4096, seed=1)
    totals = sess.play_games().float()
    mean = totals.mean(dim=0).cpu().numpy()
    assert mean[0] > -9.5 and mean[1:].max() < -11.0, mean          # MCS clearly ahead of every random seat
    assert abs(mean[1:].mean() + 13.5) < 1.5, mean                   # random seats near the README's -13.49
    wins = (totals.argmax(dim=1) == 0).float().mean().item()
    assert wins > 0.35, wins                                         # README: MCS win fraction 0.40


def test_alpha05_seat_runs_in_a_batch():
    torch.manual_seed(0)
    sess = BatchedGameSession([PolicySeat(PL.PolicyNet(), mc_max=40), RandomSeat(), MCSSeat(mc_max=40)], 96, seed=2)
    totals = sess.play_games()
    assert totals.shape == (96, 3) and int(totals.max()) <= 0
    assert (sess.env.scores().int() == -totals).all()
