"""GPU parity tests of the environment kernels, through the C ABI, against the oracle and the
reference-generated goldens.  Bit-exact: hands, rows, rewards, scores, done, observations."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN

import rl_6_nimmt_b200  # noqa: F401
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv, InvalidMoveException, SechsNimmtEnv

pytestmark = pytest.mark.gpu


def _decode(env):
    """hands [B,P,10], boards [B,4,6], scores [B,P] of the device state, via observe/scores."""
    obs = env.observe(dtype=torch.int8).cpu().numpy()
    return obs[:, :, :10], obs[:, 0, -24:].reshape(-1, 4, 6), env.scores().cpu().numpy().astype(np.int16), obs


def _replay_gpu(P, rows0, hands0, actions, include_summaries=True):
    n, T = actions.shape[:2]
    env = BatchedSechsNimmtEnv(n, P, include_summaries=include_summaries)
    env.reset_to(rows0, hands0)
    out = dict(rewards=np.zeros((n, T, P), np.int8), done=np.zeros((n, T), np.uint8), illegal=np.zeros((n, T), np.uint8),
               hands=np.zeros((n, T, P, 10), np.int8), boards=np.zeros((n, T, 4, 6), np.int8), scores=np.zeros((n, T, P), np.int16),
               obs=np.zeros((n, T, P, env.obs_len), np.int8))
    first_obs = env.observe(dtype=torch.int8).cpu().numpy()
    for t in range(T):
        a = torch.as_tensor(actions[:, t].astype(np.uint8)).cuda()
        rew, done = env.step(a)
        out["rewards"][:, t], out["done"][:, t], out["illegal"][:, t] = rew.cpu().numpy(), done.cpu().numpy(), env.illegal.cpu().numpy()
        out["hands"][:, t], out["boards"][:, t], out["scores"][:, t], out["obs"][:, t] = _decode(env)
    return out, first_obs


@pytest.mark.parametrize("P", range(2, 11))
def test_golden_traces(P):
    z = np.load(os.path.join(GOLDEN, "env_traces.npz"))
    g = lambda k: z[f"p{P}_{k}"]
    out, first = _replay_gpu(P, oracle.rows_from_singletons(g("deal_rows")), g("deal_hands"), g("actions"))
    assert not out["illegal"].any()
    assert (first == g("obs")[:, 0]).all()
    for k in ("rewards", "done"):
        assert (out[k] == g(k)).all(), k
    for k in ("hands", "boards", "scores", "obs"):
        assert (out[k] == g(k)[:, 1:]).all(), k
    out_ns, first_ns = _replay_gpu(P, oracle.rows_from_singletons(g("deal_rows")), g("deal_hands"), g("actions"), include_summaries=False)
    assert (first_ns == g("obs_ns")[:, 0]).all() and (out_ns["obs"] == g("obs_ns")[:, 1:]).all()


@pytest.mark.parametrize("P", range(1, 11))
def test_device_rng_games_vs_oracle(P):
    """20k games per P dealt and played with the device RNG, replayed through the oracle: every
    byte of every turn must agree (200k games over all P)."""
    n = 20_000
    env = BatchedSechsNimmtEnv(n, P, seed=100 + P)
    env.reset()
    hands0, boards0, scores0, _ = _decode(env)
    assert (scores0 == 0).all()
    allc = np.concatenate([hands0.reshape(n, -1), boards0[:, :, 0]], axis=1)
    srt = np.sort(allc, axis=1)
    assert (srt[:, 1:] != srt[:, :-1]).all() and srt.min() >= 0 and srt.max() <= 103
    acts = np.zeros((n, 10, P), np.int8)
    got = dict(rewards=np.zeros((n, 10, P), np.int8), done=np.zeros((n, 10), np.uint8))
    snaps = []
    for t in range(10):
        a = env.random_actions().clone()
        acts[:, t] = a.cpu().numpy().astype(np.int8)
        rew, done = env.step(a)
        got["rewards"][:, t], got["done"][:, t] = rew.cpu().numpy(), done.cpu().numpy()
        assert not env.illegal.any()
        snaps.append(_decode(env))
    want = oracle.replay(P, boards0, hands0, acts)
    assert not want["illegal"].any()
    assert (got["rewards"] == want["rewards"]).all() and (got["done"] == want["done"]).all()
    for t, (h, b, s, o) in enumerate(snaps):
        assert (h == want["hands"][:, t]).all() and (b == want["boards"][:, t]).all() and (s == want["scores"][:, t]).all()
        assert (o == want["obs"][:, t]).all()
    assert want["done"][:, -1].all() and not want["done"][:, :-1].any()


@pytest.mark.parametrize("P", [1, 2, 3, 4, 7, 10])
def test_step_random_equals_random_actions_then_step(P):
    """The fused random-play step (k_step_smem<P,true> + the plain kernel on the ragged tail) against k_random_actions
    followed by a step: same cards, rewards, done flags and state bytes every turn, with and without recording the cards;
    one more step after the game is over changes nothing."""
    n = 4096 + 37  # ragged tail block
    a_env, b_env, c_env = (BatchedSechsNimmtEnv(n, P, seed=9) for _ in range(3))
    a_env.reset(); b_env.reset(); c_env.reset()
    assert torch.equal(a_env.state, b_env.state)
    for t in range(11):
        acts = a_env.random_actions()
        ra, da = a_env.step(acts)
        rb, db = b_env.step_random(record_actions=True)
        rc, dc = c_env.step_random(record_actions=False)
        assert torch.equal(acts, b_env._actions) and torch.equal(ra, rb) and torch.equal(da, db), t
        assert torch.equal(ra, rc) and torch.equal(da, dc), t
        assert torch.equal(a_env.state, b_env.state) and torch.equal(a_env.state, c_env.state), t
        if t == 9:
            final = a_env.state.clone()
    assert bool(da.all()) and bool((acts == 255).all()) and torch.equal(a_env.state, final)


@pytest.mark.parametrize("dtype", [torch.int8, torch.int16, torch.float32, torch.int64])
def test_observe_dtypes(dtype):
    for P, n in ((3, 1000), (10, 777)):
        env = BatchedSechsNimmtEnv(n, P, seed=1).reset()
        env.step_random()
        ref = env.observe(dtype=torch.int8)
        n_legal = torch.zeros((n, P), dtype=torch.uint8, device="cuda")
        got = env.observe(dtype=dtype, n_legal=n_legal)
        assert got.dtype == dtype and torch.equal(got.to(torch.int8), ref)
        assert torch.equal(n_legal.long(), (ref[:, :, :10] >= 0).sum(dim=2))


def test_kat_a_through_the_dropin_env():
    """np.random.seed(0); SechsNimmtEnv(4).reset() deals what the reference deals (KAT-A)."""
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))["A"]
    np.random.seed(0)
    env = SechsNimmtEnv(4, verbose=False)
    states, legal = env.reset()
    assert env._hands == kat["hands"] and env._board == kat["rows"]
    assert all(s.dtype == np.int64 and s.shape == (47,) for s in states) and legal == kat["hands"]
    rewards = []
    for t in range(10):
        a = [env._hands[p][(t * (p + 1)) % len(env._hands[p])] for p in range(4)]
        (states, legal), rew, done, info = env.step(a)
        assert rew.dtype == np.int32 and isinstance(done, bool) and info == {}
        rewards.append(rew.tolist())
    assert rewards == kat["rewards"] and done
    assert (-env._scores).tolist() == kat["totals"] and env._board == kat["final_rows"]


def test_notebook_games_through_the_dropin_env():
    games = json.load(open(os.path.join(GOLDEN, "notebook_games.json")))
    for g in games:
        env = SechsNimmtEnv(g["num_players"], verbose=False)
        s0 = g["snapshots"][0]
        env.reset_to([list(r) for r in s0["board"]], [list(h) for h in s0["hands"]])
        for t, a in enumerate(g["actions"]):
            _, rew, done, _ = env.step(a)
            snap = g["snapshots"][t + 1]
            assert env._board == snap["board"] and env._hands == snap["hands"] and env._scores.tolist() == snap["scores"]
        assert done


def test_edge_cases_and_illegal_moves():
    kat = json.load(open(os.path.join(GOLDEN, "kat.json")))
    for case in kat["edge"]:
        env = SechsNimmtEnv(len(case["hands"]), verbose=False)
        env.reset_to([list(r) for r in case["board"]], [list(h) for h in case["hands"]])
        for s in case["steps"]:
            _, rew, done, _ = env.step(s["actions"])
            assert rew.tolist() == s["rewards"] and done == s["done"], case["name"]
            assert env._board == s["board"] and env._hands == s["hands"] and env._scores.tolist() == s["scores"], case["name"]
    ill = kat["illegal"]
    env = SechsNimmtEnv(2, verbose=False)
    env.reset_to([[10], [20], [30], [40]], [[1, 2], [3, 4]])
    with pytest.raises(InvalidMoveException) as e:
        env.step([1, 5])
    assert str(e.value) == ill["message"]
    assert env._board == ill["board_after"] and env._hands == ill["hands_after"]   # nothing mutated
    with pytest.raises(AssertionError):
        env.step([1])
    with pytest.raises(AssertionError):
        SechsNimmtEnv(11)
    # step after the game is over raises (all hands empty)
    env.step([1, 3]); env.step([2, 4])
    with pytest.raises(InvalidMoveException):
        env.step([1, 3])
    # batched: flagged games untouched, others advance
    b = BatchedSechsNimmtEnv(3, 2)
    board = np.tile(oracle.rows_from_singletons(np.array([[10, 20, 30, 40]]))[0], (3, 1, 1))
    hands = -np.ones((3, 2, 10), np.int8); hands[:, 0, :2] = [1, 2]; hands[:, 1, :2] = [3, 4]
    b.reset_to(board, hands)
    before = b.state.clone()
    rew, done = b.step(torch.tensor([[1, 3], [1, 5], [200, 3]], dtype=torch.uint8).cuda())
    assert b.illegal.tolist() == [0, 1, 1] and rew[1:].abs().sum() == 0
    obs = b.observe(dtype=torch.int8).cpu().numpy()
    assert obs[0, 0, :2].tolist() == [2, -1] and obs[1, 0, :2].tolist() == [1, 2] and obs[2, 1, :2].tolist() == [3, 4]
    with pytest.raises(ValueError):
        b.reset_to(board, np.where(hands == 3, 1, hands))  # duplicate card


def test_deal_distribution_and_geometry_invariance():
    """chi-square uniformity of dealt cards per seat / row and game0 invariance."""
    n, P = 200_000, 4
    env = BatchedSechsNimmtEnv(n, P, seed=42).reset()
    hands, boards, _, _ = _decode(env)
    for cards in (hands[:, 0].reshape(-1), hands[:, 3].reshape(-1), boards[:, 0, 0], boards[:, 3, 0]):
        counts = np.bincount(cards.astype(np.int64), minlength=104)
        expect = len(cards) / 104
        chi2 = ((counts - expect) ** 2 / expect).sum()
        assert chi2 < 180, chi2  # 103 dof: mean 103, sd 14.4; 180 is > 5 sd
    # splitting the batch over two "ranks" deals the same games
    lo = BatchedSechsNimmtEnv(n // 2, P, seed=42, game0=0).reset()
    hi = BatchedSechsNimmtEnv(n // 2, P, seed=42, game0=n // 2).reset()
    h_lo, b_lo, _, _ = _decode(lo); h_hi, b_hi, _, _ = _decode(hi)
    assert (np.concatenate([h_lo, h_hi]) == hands).all() and (np.concatenate([b_lo, b_hi]) == boards).all()
    # uniform random action index
    acts = env.random_actions().cpu().numpy()
    idx = (hands < acts[:, :, None].astype(np.int16)).sum(axis=2).reshape(-1)
    counts = np.bincount(idx, minlength=10)
    chi2 = ((counts - len(idx) / 10) ** 2 / (len(idx) / 10)).sum()
    assert chi2 < 40, chi2  # 9 dof


def test_full_size_properties_1m_games():
    """BASELINE config 2 size (2^20 four-player games): size-independent invariants after a full game."""
    n, P = 1 << 20, 4
    env = BatchedSechsNimmtEnv(n, P, seed=7).reset()
    total = torch.zeros((n, P), dtype=torch.int32, device="cuda")
    vals = torch.tensor([oracle.card_values()[c] for c in range(104)], dtype=torch.int32, device="cuda")
    for t in range(10):
        rew, done = env.step_random()
        total += rew.int()
        assert bool(done.all()) == (t == 9)
    scores = env.scores().int()
    assert torch.equal(scores, -total)                      # cumulative score == minus the summed rewards
    obs = env.observe(dtype=torch.int8)
    assert bool((obs[:, :, :10] == -1).all())               # all hands empty
    board = obs[:, 0, -24:].long()
    on_board = torch.where(board >= 0, vals[board.clamp(min=0)], torch.zeros_like(board, dtype=torch.int32)).sum(dim=1)
    assert bool((on_board == obs[:, 0, 19:23].sum(dim=1)).all())   # row sums in the observation match the cards
    # bull heads are conserved: 44 dealt cards = taken + still on the board ... per game, via the deal
    env2 = BatchedSechsNimmtEnv(n, P, seed=7).reset()
    obs0 = env2.observe(dtype=torch.int8)
    dealt = torch.cat([obs0[:, :, :10].reshape(n, -1), obs0[:, 0, -24:]], dim=1).long()
    dealt_val = torch.where(dealt >= 0, vals[dealt.clamp(min=0)], torch.zeros_like(dealt, dtype=torch.int32)).sum(dim=1)
    assert torch.equal(dealt_val, scores.sum(dim=1) + on_board)
    assert bool(((obs[:, 0, 11:15] >= 1) & (obs[:, 0, 11:15] <= 5)).all())


def test_step_host_matches_step_with_byte_and_bit_done():
    """The end-to-end path (pinned host buffers, bench.py's `e2e`): same rewards and done as step(), with done
    delivered either as one byte or as one bit per game (nimmt_pack_flags), incl. a ragged batch size."""
    for B in (4096, 1000):
        P = 3
        a_env = BatchedSechsNimmtEnv(B, P, seed=21).reset()
        b_env = BatchedSechsNimmtEnv(B, P, seed=21).reset()
        h_act = torch.empty((B, P), dtype=torch.uint8).pin_memory()
        h_rew = torch.empty((B, P), dtype=torch.int8).pin_memory()
        h_done = torch.empty((B,), dtype=torch.uint8).pin_memory()
        h_bits = torch.empty(((B + 31) // 32,), dtype=torch.int32).pin_memory()
        for t in range(10):
            acts = a_env.random_actions().clone()
            rew, done = a_env.step(acts)
            h_act.copy_(acts)
            torch.cuda.synchronize()
            if t % 3 == 0:
                b_env.step_host(h_act, h_rew, h_done)
                torch.cuda.synchronize()
                got_done = h_done.numpy()
            elif t % 3 == 1:
                out, o_rew, o_bits = b_env.host_out_buffer()
                b_env.step_host(h_act, out)
                torch.cuda.synchronize()
                h_rew.copy_(o_rew)
                got_done = np.unpackbits(o_bits.numpy().view(np.uint8), bitorder="little")[:B]
            else:
                b_env.step_host(h_act, h_rew, h_bits)
                torch.cuda.synchronize()
                got_done = np.unpackbits(h_bits.numpy().view(np.uint8), bitorder="little")[:B]
                assert not np.unpackbits(h_bits.numpy().view(np.uint8), bitorder="little")[B:].any()
            assert (h_rew.numpy() == rew.cpu().numpy()).all() and (got_done == done.cpu().numpy()).all(), (B, t)
        assert got_done.all() and torch.equal(a_env.state, b_env.state)


@pytest.mark.parametrize("P", (2, 4, 7, 10))
def test_reset_to_midgame_positions_from_the_golden_traces(P):
    """reset_to with short hands (the MC agents' roots, agents/mcts.py:108-114): every golden game is re-entered at turns
    3, 7 and 9 — hands of 7, 3 and 1 cards, listed in DESCENDING order to show that the stored record does not depend on
    the caller's order — and played on; rewards, done, hands, boards and observations must match the reference's
    trace from there on (scores restart at zero, env.py:58)."""
    z = np.load(os.path.join(GOLDEN, "env_traces.npz"))
    g = lambda k: z[f"p{P}_{k}"]
    for t0 in (3, 7, 9):
        hands0 = g("hands")[:, t0].copy()                       # state after t0 steps: [n, P, 10], -1 padded
        shuffled = -np.ones_like(hands0)
        for idx in np.ndindex(hands0.shape[:2]):
            cards = hands0[idx][hands0[idx] >= 0][::-1]
            shuffled[idx][: len(cards)] = cards
        out, first = _replay_gpu(P, g("boards")[:, t0], shuffled, g("actions")[:, t0:])
        assert not out["illegal"].any()
        assert (first == g("obs")[:, t0]).all()
        for k in ("rewards", "done"):
            assert (out[k] == g(k)[:, t0:]).all(), (k, t0)
        for k in ("hands", "boards", "obs"):
            assert (out[k] == g(k)[:, t0 + 1:]).all(), (k, t0)
        base = g("scores")[:, t0][:, None, :]
        assert (out["scores"] == g("scores")[:, t0 + 1:] - base).all()


def test_batches_on_four_streams_equal_one_stream():
    """The ABI promises re-entrant, enqueue-only calls that are safe on distinct state buffers (threading): four
    batches dealt and played concurrently, each on its own stream (the shape of bench.py's four-stream figure), end in exactly
    the states, rewards and flags that the same four batches reach one after the other on one stream."""
    n, P = (1 << 16) + 32 * 7, 4

    def play(streams):
        envs = [BatchedSechsNimmtEnv(n, P, seed=300 + b, game0=b * n) for b in range(4)]
        outs = []
        for b, env in enumerate(envs):
            st = streams[b] if streams else torch.cuda.current_stream()
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                env.reset()
                rews, dones = [], []
                for t in range(10):
                    rew, done = env.step(env.random_actions())
                    rews.append(rew.clone())
                    dones.append(done.clone())
                outs.append((torch.stack(rews), torch.stack(dones), env.state.clone(), env.illegal.clone()))
        torch.cuda.synchronize()
        return outs

    serial = play(None)
    for rep in range(2):
        conc = play([torch.cuda.Stream() for _ in range(4)])
        for b in range(4):
            for x, y in zip(serial[b], conc[b]):
                assert torch.equal(x, y), (rep, b)
    assert all(bool(o[1][-1].all()) and not bool(o[3].any()) for o in serial)
