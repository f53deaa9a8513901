"""GPU tests of the rollout kernel and the MCSAgent drop-in, through the C ABI."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN

import rl_6_nimmt_b200  # noqa: F401
from rl_6_nimmt_b200 import rollouts as R
from rl_6_nimmt_b200.agents import DrunkHamster, MCSAgent
from rl_6_nimmt_b200.env import SechsNimmtEnv

pytestmark = pytest.mark.gpu


def _golden():
    return json.load(open(os.path.join(GOLDEN, "mcs_exact.json")))


@pytest.mark.parametrize("key", ["C", "D", "E"])
def test_kernel_vs_exact_enumeration(key):
    """Tolerance: |mean - exact| < 4.5 sigma / sqrt(N) with sigma from the exact law (N = 2e6)."""
    m = _golden()[key]
    root = R.pack_root(m["board"], m["own"], m["available"], m["P"])
    N = 2_000_000
    stats = R.mcs_rollouts(root[None], m["P"], N, seed=12)[0].cpu().numpy()
    for i, a in enumerate(sorted(m["own"])):
        ex = m["exact"][str(a)]
        var = ex["sumsq"] / ex["count"] - ex["mean"] ** 2
        s, ss, n = (int(x) for x in stats[i])
        assert n == N
        assert abs(s / n - ex["mean"]) < 4.5 * np.sqrt(var / n) + 1e-12, (key, a, s / n, ex["mean"])
        assert abs(ss / n - ex["sumsq"] / ex["count"]) < 0.02 * max(1.0, ex["sumsq"] / ex["count"])
    assert (stats[len(m["own"]):] == 0).all()


@pytest.mark.parametrize("key", ["MC4", "MC3"])
def test_kernel_vs_reference_law(key):
    c = _golden()[key]
    root = R.pack_root_from_state(np.array(c["state"]), c["legal"], c["available"])
    N = 1_000_000
    got = R.mcs_rollouts(root[None], c["P"], N, seed=5)[0].cpu().numpy()
    board = [[int(x) for x in row if x >= 0] for row in np.array(c["state"][-24:]).reshape(4, 6)]
    want = oracle.mcs_rollouts(c["P"], board, c["legal"], c["available"], 1_000_000, seed=17)
    for i, a in enumerate(c["legal"]):
        s, ss, n = (int(x) for x in got[i])
        mean, var = s / n, ss / n - (s / n) ** 2
        ws, wss, wn = (int(x) for x in want[i])
        wmean, wvar = ws / wn, wss / wn - (ws / wn) ** 2
        z = (mean - wmean) / np.sqrt(var / n + wvar / wn)
        assert abs(z) < 4.5, (key, a, mean, wmean, z)
        r = c["reference_mc"][str(a)]
        z = (mean - r["mean"]) / np.sqrt(var / n + r["var"] / r["count"])
        assert abs(z) < 4.5, (key, a, mean, r["mean"], z)


def test_kernel_equals_host_build_of_the_same_code():
    """The CUDA threads and the CPU build of rollout.cuh produce identical integers (same ids, same RNG)."""
    from host_sim import mcs
    m = _golden()["E"]
    root = R.pack_root(m["board"], m["own"], m["available"], m["P"])
    got = R.mcs_rollouts(root[None], m["P"], 3000, seed=77)[0].cpu().numpy()
    assert (got == mcs(m["P"], root.tobytes(), 3000, seed=77)).all()


def test_striping_invariance_and_batching():
    g = _golden()
    roots = np.stack([R.pack_root(g[k]["board"], g[k]["own"], g[k]["available"], g[k]["P"]) for k in ("C", "E")])
    whole = R.mcs_rollouts(roots, 2, 10_001, seed=3).cpu()
    for world in (2, 4, 8):
        parts = sum(R.mcs_rollouts(roots, 2, 10_001, seed=3, rank=r, world=world).cpu() for r in range(world))
        assert torch.equal(parts, whole)
    assert whole[0, :2, 2].tolist() == [10_001] * 2 and whole[1, :3, 2].tolist() == [10_001] * 3
    # root D has 3 players: skipped (stats zero) when the call says 2
    mixed = np.stack([roots[0], R.pack_root(g["D"]["board"], g["D"]["own"], g["D"]["available"], 3)])
    out = R.mcs_rollouts(mixed, 2, 100, seed=3).cpu()
    assert out[1].abs().sum() == 0 and out[0, 0, 2] == 100


def test_mcs_agent_dropin_plays_and_prefers_the_better_card():
    m = _golden()["C"]
    agent = MCSAgent(mc_max=200, rollouts_per_card=200_000, seed=1)
    agent.num_players = 2
    agent.available_cards = list(m["available"])
    state = np.full(47, -1, np.int64)
    state[:2] = m["own"]; state[10] = 2
    for r, cards in enumerate(m["board"]):
        state[23 + 6 * r: 23 + 6 * r + len(cards)] = cards
    action, info = agent(torch.tensor(state, dtype=torch.float), legal_actions=list(m["own"]))
    assert action == 25 and "log_prob" in info          # E[25] = -2.68 > E[43] = -5.00
    means = agent.last_stats[:, 0] / agent.last_stats[:, 2]
    assert abs(means[0] - m["exact"]["25"]["mean"]) < 0.03 and abs(means[1] - m["exact"]["43"]["mean"]) < 0.03
    # a whole game vs random agents through the drop-in env, reference-style loop (play.py:23-75)
    np.random.seed(3)
    env = SechsNimmtEnv(3, verbose=False)
    agents = [MCSAgent(mc_max=200, seed=5), DrunkHamster(), DrunkHamster()]
    states, legal = env.reset()
    done, turns = False, 0
    while not done:
        acts = [int(ag(torch.tensor(s, dtype=torch.float), legal_actions=l)[0]) for ag, s, l in zip(agents, states, legal)]
        (states, legal), rew, done, _ = env.step(acts)
        turns += 1
    assert turns == 10
    # stale card memory: only cards seen on the board at decision time were removed (mcts.py:66-73)
    assert len(agents[0].available_cards) >= 104 - 10 - 4 - 9 * 3


def test_tie_and_nan_semantics():
    # strict '>' in ascending card order: first maximum wins; unvisited (count 0) never wins
    stats = np.array([[-30, 0, 10], [-30, 0, 10], [0, 0, 0]], np.int64)
    assert R.choose_from_stats([5, 9, 50], stats)[0] == 5
    stats = np.array([[0, 0, 0], [-40, 0, 10], [-30, 0, 10]], np.int64)
    assert R.choose_from_stats([5, 9, 50], stats)[0] == 50
