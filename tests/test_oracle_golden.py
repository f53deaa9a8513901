"""Pins the CPU oracle (oracle/) against fixtures produced by the unmodified reference.

CPU-only.  If these fail the oracle is wrong and no GPU parity claim means anything.
"""
import json
import os

import numpy as np
import pytest

import oracle
from oracle import policy_oracle as po

from conftest import GOLDEN


def _board6(board):
    out = -np.ones((4, 6), np.int8)
    for r, cards in enumerate(board):
        out[r, : len(cards)] = cards
    return out


def _hands10(hands):
    out = -np.ones((len(hands), 10), np.int8)
    for p, h in enumerate(hands):
        out[p, : len(h)] = h
    return out


@pytest.fixture(scope="module")
def kat():
    return json.load(open(os.path.join(GOLDEN, "kat.json")))


def test_card_values(kat):
    # env.py:224-239; SURVEY.md §3.5 table
    v = oracle.card_values()
    assert v.tolist() == kat["card_values"]
    assert v.sum() == 171
    assert {int(k): int(c) for k, c in zip(*np.unique(v, return_counts=True))} == {1: 76, 2: 9, 3: 10, 5: 8, 7: 1}


def test_notebook_games_replay():
    """The five games recorded in experiments/simple_tournament.ipynb replay exactly."""
    games = json.load(open(os.path.join(GOLDEN, "notebook_games.json")))
    assert len(games) == 5
    finals = []
    for g in games:
        s0 = g["snapshots"][0]
        P = g["num_players"]
        acts = np.array(g["actions"], np.int8)[None]
        out = oracle.replay(P, _board6(s0["board"])[None], _hands10(s0["hands"])[None], acts)
        assert not out["illegal"].any()
        for t in range(10):
            snap = g["snapshots"][t + 1]
            assert (out["boards"][0, t] == _board6(snap["board"])).all(), t
            assert (out["hands"][0, t] == _hands10(snap["hands"])).all(), t
            assert out["scores"][0, t].tolist() == snap["scores"], t
        assert out["done"][0].tolist() == [0] * 9 + [1]
        finals.append(out["scores"][0, -1].tolist())
    # README.md:48-51 — Merle 10 16 11 3 4, Alpha0.5 1 3 14 8 6
    assert finals == [[10, 1], [16, 3], [11, 14], [3, 8], [4, 6]]


@pytest.mark.parametrize("P", range(2, 11))
def test_env_traces(P):
    """120 random-play games per P: rewards, done, hands, boards, scores, observations."""
    z = np.load(os.path.join(GOLDEN, "env_traces.npz"))
    g = lambda k: z[f"p{P}_{k}"]
    rows0 = oracle.rows_from_singletons(g("deal_rows"))
    out = oracle.replay(P, rows0, g("deal_hands"), g("actions"))
    assert not out["illegal"].any()
    assert (out["rewards"] == g("rewards")).all()
    assert (out["done"] == g("done")).all()
    assert (out["hands"] == g("hands")[:, 1:]).all()
    assert (out["boards"] == g("boards")[:, 1:]).all()
    assert (out["scores"] == g("scores")[:, 1:]).all()
    assert (out["obs"] == g("obs")[:, 1:]).all()
    out_ns = oracle.replay(P, rows0, g("deal_hands"), g("actions"), include_summaries=False)
    assert (out_ns["obs"] == g("obs_ns")[:, 1:]).all()
    # observation right after reset: replay zero steps is not expressible; check via a 1-step illegal probe
    bad = g("actions")[:, :1].copy()
    bad[:, 0, 0] = np.where(g("deal_hands")[:, 0, 0] == 0, 103, 0)  # card 0 or 103, whichever is not held ... usually
    held = (g("deal_hands")[:, 0, :] == bad[:, 0, 0][:, None]).any(axis=1)
    probe = oracle.replay(P, rows0, g("deal_hands"), bad)
    sel = ~held
    assert probe["illegal"][sel, 0].all()
    assert (probe["obs"][sel, 0] == g("obs")[sel, 0]).all()  # untouched state == reset observation
    assert (probe["rewards"][sel] == 0).all()


def test_kat_a(kat):
    a = kat["A"]
    hands = np.array(a["hands"], np.int8)
    rows = np.array([r[0] for r in a["rows"]], np.int8)
    acts = np.zeros((1, 10, 4), np.int8)
    cur = [list(h) for h in a["hands"]]
    # index policy a_p(t) = hand_p[(t (p+1)) mod len]
    for t in range(10):
        for p in range(4):
            c = cur[p][(t * (p + 1)) % len(cur[p])]
            acts[0, t, p] = c
        for p in range(4):
            cur[p].remove(int(acts[0, t, p]))
    out = oracle.replay(4, oracle.rows_from_singletons(rows[None]), hands[None], acts)
    assert out["rewards"][0].tolist() == a["rewards"]
    assert (-out["scores"][0, -1]).tolist() == a["totals"] == [-17, -13, -8, -7]
    final = [[int(c) for c in row if c >= 0] for row in out["boards"][0, -1]]
    assert final == a["final_rows"]


@pytest.mark.parametrize("P", [2, 4, 10])
def test_kat_b(kat, P):
    deals = np.load(os.path.join(GOLDEN, f"katb_deals_p{P}.npy"))
    hands = deals[:, : 10 * P].reshape(-1, P, 10)
    rows = deals[:, 10 * P:]
    n = len(deals)
    acts = np.zeros((n, 10, P), np.int8)
    cur = hands.copy().astype(np.int16)  # 127 = played sentinel keeps sort order
    for t in range(10):
        srt = np.sort(cur, axis=2)
        ln = 10 - t
        for p in range(P):
            acts[:, t, p] = srt[:, p, (t * (p + 1)) % ln]
        cur = np.where(cur == acts[:, t, :, None], 127, cur)
    out = oracle.replay(P, oracle.rows_from_singletons(rows), hands, acts, want_obs=False)
    assert not out["illegal"].any()
    assert out["rewards"].astype(np.int64).sum(axis=(0, 1)).tolist() == kat["B"][str(P)]["score_sums"]
    assert int(np.count_nonzero(out["rewards"])) == kat["B"][str(P)]["events"]


def test_edge_cases(kat):
    for case in kat["edge"]:
        P = len(case["hands"])
        acts = np.array([s["actions"] for s in case["steps"]], np.int8)[None]
        out = oracle.replay(P, _board6(case["board"])[None], _hands10(case["hands"])[None], acts)
        for t, s in enumerate(case["steps"]):
            assert out["rewards"][0, t].tolist() == s["rewards"], case["name"]
            assert (out["boards"][0, t] == _board6(s["board"])).all(), case["name"]
            assert (out["hands"][0, t] == _hands10(s["hands"])).all(), case["name"]
            assert out["scores"][0, t].tolist() == s["scores"], case["name"]
            assert bool(out["done"][0, t]) == s["done"], case["name"]
    names = {c["name"]: c for c in kat["edge"]}
    assert names["undercut_tie_lowest_index"]["steps"][0]["board"][1] == [3]
    assert names["sixth_card"]["steps"][0]["rewards"] == [-sum(kat["card_values"][c] for c in [10, 11, 12, 13, 14]), 0]


def test_illegal_move_leaves_state(kat):
    ill = kat["illegal"]
    assert ill["raised"] and ill["short_asserts"] and ill["p11_asserts"] and ill["p10_uses_all_cards"]
    out = oracle.replay(2, _board6([[10], [20], [30], [40]])[None], _hands10([[1, 2], [3, 4]])[None],
                        np.array([[[1, 5]]], np.int8))
    assert out["illegal"][0, 0] == 1
    assert [[int(c) for c in r if c >= 0] for r in out["boards"][0, 0]] == ill["board_after"]
    assert [[int(c) for c in h if c >= 0] for h in out["hands"][0, 0]] == ill["hands_after"]


@pytest.mark.parametrize("key", ["C", "D", "E"])
def test_mcs_oracle_vs_exact(key):
    """C-oracle rollouts (reference rollout law) agree with exact enumeration within 4.5 sigma."""
    m = json.load(open(os.path.join(GOLDEN, "mcs_exact.json")))[key]
    N = 400_000
    stats = oracle.mcs_rollouts(m["P"], m["board"], m["own"], m["available"], N, seed=99)
    assert stats[:, 2].sum() == N
    for i, a in enumerate(m["own"]):
        ex = m["exact"][str(a)]
        mean_exact = ex["mean"]
        var_exact = ex["sumsq"] / ex["count"] - mean_exact ** 2
        s, ss, n = stats[i]
        assert abs(s / n - mean_exact) < 4.5 * np.sqrt(var_exact / n) + 1e-12, (key, a, s / n, mean_exact)
        # uniform first move => counts ~ N / n_own
        assert abs(n - N / len(m["own"])) < 5 * np.sqrt(N)


def test_mcs_oracle_vs_reference_mc():
    m = json.load(open(os.path.join(GOLDEN, "mcs_exact.json")))
    for key in ("MC4", "MC3"):
        c = m[key]
        board = np.array(c["state"][-24:]).reshape(4, 6)
        board = [[int(x) for x in row if x >= 0] for row in board]
        stats = oracle.mcs_rollouts(c["P"], board, c["legal"], c["available"], 300_000, seed=5)
        for i, a in enumerate(c["legal"]):
            r = c["reference_mc"][str(a)]
            s, ss, n = stats[i]
            mean, var = s / n, ss / n - (s / n) ** 2
            z = (mean - r["mean"]) / np.sqrt(var / n + r["var"] / r["count"])
            assert abs(z) < 4.5, (key, a, mean, r["mean"], z)


def test_policy_oracle():
    z = np.load(os.path.join(GOLDEN, "policy_vectors.npz"))
    w = po.weights_from_golden(z)
    norm = po.normalize(z["rows_in"])
    np.testing.assert_allclose(norm, z["rows_norm"], rtol=0, atol=1e-6)
    scale, shift = po.normalization_affine()
    np.testing.assert_allclose(z["rows_in"] * scale + shift, z["rows_norm"], atol=1e-5)
    logits = po.mlp_logits(z["rows_norm"], w)
    np.testing.assert_allclose(logits, z["logits"], rtol=1e-4, atol=1e-5)
    off = 0
    for n in z["seg"]:
        p = po.policy_probs(z["rows_in"][off:off + n], w)
        np.testing.assert_allclose(p, z["probs"][off:off + n], rtol=1e-4, atol=1e-6)
        off += n
    assert off == len(z["rows_in"])


def test_puct_oracle():
    cases = json.load(open(os.path.join(GOLDEN, "puct_cases.json")))
    n_nan = 0
    for c in cases:
        outcomes = {int(a): o for a, o in c["outcomes"].items()}
        assert list(po.normalize_q(outcomes)) == c["norm"]
        p = po.pucts(c["legal"], outcomes, np.array(c["probs"], np.float32))
        for got, want in zip(p, c["pucts"]):
            if want is None:
                assert np.isnan(got)
                n_nan += 1
            else:
                assert abs(got - want) < 1e-9
        assert po.puct_choice(p) == c["choice"]
    assert n_nan > 0  # the 0/0 case (all outcomes equal) is covered


def test_policy_rollout_oracle_vs_reference():
    """The fp32 C restatement of PolicyMCSAgent rollouts agrees with the unmodified reference (sharpened
    policy, 4-player mid-game root): z-test per first card, and first-card frequencies ~ root policy."""
    z = np.load(os.path.join(GOLDEN, "policy_rollouts.npz"))
    w = {"w1": z["w_actor_latent_net_0_weight"], "b1": z["w_actor_latent_net_0_bias"], "w2": z["w_actor_latent_net_2_weight"],
         "b2": z["w_actor_latent_net_2_bias"], "w3": z["w_actor_head_nets_0_0_weight"], "b3": z["w_actor_head_nets_0_0_bias"]}
    board = [[int(c) for c in row if c >= 0] for row in z["state"][-24:].reshape(4, 6)]
    legal, avail = z["legal"].tolist(), z["available"].tolist()
    N = 25_000
    stats = oracle.policy_rollouts(4, board, legal, avail, N, w, seed=3)
    assert stats[:, 2].sum() == N
    for i, a in enumerate(legal):
        s, ss, n = (int(x) for x in stats[i])
        mean, var = s / n, ss / n - (s / n) ** 2
        zscore = (mean - z["mean"][i]) / np.sqrt(var / n + z["var"][i] / z["count"][i])
        assert abs(zscore) < 4.5, (a, mean, z["mean"][i], zscore)
        assert abs(n / N - z["root_probs"][i]) < 5 * np.sqrt(z["root_probs"][i] / N) + 1e-3
    # the policy matters: uniform-random rollouts from the same root give different values for some card
    uni = oracle.mcs_rollouts(4, board, legal, avail, N, seed=4)
    diffs = [abs(stats[i, 0] / stats[i, 2] - uni[i, 0] / uni[i, 2]) for i in range(len(legal))]
    assert max(diffs) > 0.15, diffs
