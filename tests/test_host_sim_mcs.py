"""CPU check of the product's rollout logic (rollout.cuh, the code each CUDA thread runs) against
exact enumeration through the reference env and against the reference-law oracle.  This is where
the 'lazy opponents' equivalence is tested without a GPU."""
import json
import os

import numpy as np
import pytest

import oracle
from conftest import GOLDEN
from host_sim import mcs

import rl_6_nimmt_b200  # noqa: F401  (import shim)
from rl_6_nimmt_b200 import rollouts as R


def _golden():
    return json.load(open(os.path.join(GOLDEN, "mcs_exact.json")))


@pytest.mark.parametrize("key", ["C", "D", "E"])
def test_rollouts_vs_exact_enumeration(key):
    m = _golden()[key]
    root = R.pack_root(m["board"], m["own"], m["available"], m["P"])
    N = 150_000
    stats = mcs(m["P"], root.tobytes(), N, seed=2024)
    for i, a in enumerate(sorted(m["own"])):
        ex = m["exact"][str(a)]
        var = ex["sumsq"] / ex["count"] - ex["mean"] ** 2
        s, ss, n = stats[i]
        assert n == N
        assert abs(s / n - ex["mean"]) < 4.5 * np.sqrt(var / n) + 1e-12, (key, a, s / n, ex["mean"])
        # second moment too: the whole outcome law must match, not just its mean
        ex_m2 = ex["sumsq"] / ex["count"]
        assert abs(ss / n - ex_m2) < 0.05 * max(1.0, ex_m2), (key, a)
    assert (stats[len(m["own"]):] == 0).all()


@pytest.mark.parametrize("key", ["MC4", "MC3"])
def test_rollouts_vs_reference_law_oracle(key):
    """Opening and mid-game 3/4-player roots: z-test against the oracle that deals opponent hands
    the reference's way, and against the reference agent's own Monte-Carlo numbers."""
    c = _golden()[key]
    root = R.pack_root_from_state(np.array(c["state"]), c["legal"], c["available"])
    N = 60_000
    got = mcs(c["P"], root.tobytes(), N, seed=7)
    board = [[int(x) for x in row if x >= 0] for row in np.array(c["state"][-24:]).reshape(4, 6)]
    want = oracle.mcs_rollouts(c["P"], board, c["legal"], c["available"], 400_000, seed=11)
    for i, a in enumerate(c["legal"]):
        s, ss, n = got[i]
        mean, var = s / n, ss / n - (s / n) ** 2
        ws, wss, wn = want[i]
        wmean, wvar = ws / wn, wss / wn - (ws / wn) ** 2
        z = (mean - wmean) / np.sqrt(var / n + wvar / wn)
        assert abs(z) < 4.5, (key, a, mean, wmean, z)
        assert abs(var - wvar) < 0.1 * wvar, (key, a, var, wvar)
        r = c["reference_mc"][str(a)]
        z = (mean - r["mean"]) / np.sqrt(var / n + r["var"] / r["count"])
        assert abs(z) < 4.5, (key, a, mean, r["mean"], z)


def test_striping_is_world_size_invariant():
    """Same seed => identical integer tables whether 1, 2, 4 or 8 ranks play the stripes."""
    m = _golden()["D"]
    root = R.pack_root(m["board"], m["own"], m["available"], m["P"]).tobytes()
    whole = mcs(m["P"], root, 5000, seed=3)
    for world in (2, 4, 8):
        parts = sum(mcs(m["P"], root, 5000, seed=3, rank=r, world=world) for r in range(world))
        assert (parts == whole).all()
    assert (mcs(m["P"], root, 5000, seed=4) != whole).any()


def test_root_rejects_short_pool():
    root = R.pack_root([[1], [2], [3], [4]], [10, 20, 30], [40, 41, 42], 3)  # needs 6 unseen cards
    import ctypes
    from host_sim import lib, _p
    stats = np.zeros((10, 3), np.int64)
    buf = (ctypes.c_uint8 * 64).from_buffer_copy(root.tobytes())
    assert lib().sim_mcs(3, buf, ctypes.c_int64(10), ctypes.c_uint64(1), 0, 1, _p(stats)) == -2


def test_puct_rule_matches_reference_cases():
    """puct.cuh (the code the rollout kernel runs at the root) against PUCTAgent._compute_pucts /
    _normalize_q vectors generated from the reference, including <10 outcomes => (0,-10,-5), the
    median rule, and the all-equal 0/0 => NaN => first card case."""
    from host_sim import puct
    cases = json.load(open(os.path.join(GOLDEN, "puct_cases.json")))
    n_nan = 0
    for c in cases:
        outcomes = {int(a): o for a, o in c["outcomes"].items()}
        choice, pucts = puct(c["legal"], outcomes, c["probs"])
        for got, want in zip(pucts, c["pucts"]):
            if want is None:
                assert np.isnan(got)
                n_nan += 1
            else:
                assert abs(got - want) < 1e-12, (got, want)
        assert choice == c["choice"]
    assert n_nan > 0


def test_fisher_yates_resolved_without_swapping():
    """rollout.cuh::fisher_yates_source (the policy rollouts deal the opponents with it, one thread per draw) against a
    shuffle that really swaps."""
    from host_sim import fisher_yates_sources
    rng = np.random.RandomState(8)
    for _ in range(300):
        n_avail = int(rng.randint(1, 95))
        steps = int(rng.randint(1, min(n_avail, 90) + 1))
        targets = np.array([i + rng.randint(n_avail - i) for i in range(steps)], np.uint8)
        deck = list(range(n_avail))
        want = []
        for i, k in enumerate(targets):
            deck[i], deck[k] = deck[k], deck[i]
            want.append(deck[i])
        assert fisher_yates_sources(targets).tolist() == want
