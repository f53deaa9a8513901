"""GPU parity tests added in round 2 ("Close the parity gaps"), all through the C ABI:

  a21  the DEVICE PUCT root rule (puct.cuh::puct_choose_lanes, the code k_policy_rollouts runs) on the 40 reference cases
  a14  BaseMCAgent's card memory — host agent AND k_mc_roots — against the set the unmodified reference agent held after
       watching the same four turns (tests/golden/mcs_exact.json["MC3"], agents/mcts.py:62-73)
  f3   MaskedPolicySeat probabilities on the device against tests/golden/masked_policy.npz
  C2   the full BASELINE configs[1] size, 2^20 four-player games, every byte of every turn against the oracle
"""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from conftest import GOLDEN

import rl_6_nimmt_b200  # noqa: F401
from rl_6_nimmt_b200 import _native as N
from rl_6_nimmt_b200 import policy as PL
from rl_6_nimmt_b200 import rollouts as R
from rl_6_nimmt_b200.agents import MCSAgent
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv, SechsNimmtEnv

pytestmark = pytest.mark.gpu


def _puct_device(cases, c_puct=2.0):
    offsets, idx, out, probs, n_legal = [0], [], [], np.zeros((len(cases), 10), np.float32), []
    for d, c in enumerate(cases):
        outcomes = {int(a): o for a, o in c["outcomes"].items()}
        # the reference appends outcomes per card; within one card the order is the order of play, and the rule depends on the
        # multiset only (counts, sums, min, max, median), so any interleaving gives the same answer
        for i, a in enumerate(c["legal"]):
            for o in outcomes[a]:
                idx.append(i)
                out.append(int(o))
        offsets.append(len(idx))
        probs[d, : len(c["legal"])] = c["probs"]
        n_legal.append(len(c["legal"]))
    dev = "cuda"
    t = lambda a, dt: torch.as_tensor(np.asarray(a, dt)).to(dev)
    offsets_d, idx_d, out_d = t(offsets, np.int32), t(idx if idx else [0], np.int32), t(out if out else [0], np.int32)
    probs_d, n_d = torch.as_tensor(probs).to(dev), t(n_legal, np.int32)
    pucts = torch.zeros((len(cases), 10), dtype=torch.float64, device=dev)
    choice = torch.full((len(cases),), -1, dtype=torch.int32, device=dev)
    N.check(N.lib().nimmt_puct_choose(N.ptr(offsets_d), N.ptr(idx_d), N.ptr(out_d), N.ptr(probs_d), N.ptr(n_d), len(cases), c_puct,
                                      N.ptr(pucts), N.ptr(choice), N.current_stream()), "nimmt_puct_choose")
    return pucts.cpu().numpy(), choice.cpu().numpy()


def test_device_puct_rule_matches_the_reference_cases():
    """a21: PUCTAgent._compute_pucts / _normalize_q / the strict-'>' choice (agents/mcts.py:276-315) as the DEVICE evaluates
    them — fp64 on the GPU, one card per lane, shuffles, the shared-memory histogram median — against the 40 vectors generated
    from the unmodified reference: every PUCT value to 1e-12, NaN where the reference has NaN, the choice exactly."""
    cases = json.load(open(os.path.join(GOLDEN, "puct_cases.json")))
    pucts, choice = _puct_device(cases)
    n_nan = n_prior = n_median = 0
    for d, c in enumerate(cases):
        for a, want in enumerate(c["pucts"]):
            if want is None:
                assert np.isnan(pucts[d, a]), (d, a, pucts[d, a])
                n_nan += 1
            else:
                assert abs(pucts[d, a] - want) < 1e-12, (d, a, pucts[d, a], want)
        assert (pucts[d, len(c["legal"]):] == 0).all()
        assert choice[d] == c["choice"], (d, choice[d], c["choice"])
        total = sum(len(o) for o in c["outcomes"].values())
        n_prior += total < 10
        n_median += total >= 10
    assert n_nan > 0 and n_prior > 0 and n_median > 0      # all three regimes of _normalize_q are in the fixture


def test_device_puct_rule_random_cases_against_the_numpy_restatement():
    """More cases than the fixture holds (odd/even outcome counts, ties, unvisited cards, every hand size), against
    oracle/policy_oracle.py's restatement of the same lines — which the CPU suite pins to the 40 reference cases."""
    from oracle import policy_oracle as PO
    rng = np.random.RandomState(5)
    cases = []
    for _ in range(400):
        n = int(rng.randint(1, 11))
        legal = sorted(rng.choice(104, n, replace=False).tolist())
        total = int(rng.choice([0, 1, 5, 9, 10, 11, 40, 199]))
        outcomes = {a: [] for a in legal}
        spread = int(rng.choice([0, 3, 30]))
        for _ in range(total):
            outcomes[legal[int(rng.randint(n))]].append(-float(rng.randint(0, spread + 1)))
        p = rng.dirichlet(np.ones(n)).astype(np.float32)
        cases.append({"legal": legal, "outcomes": {str(a): o for a, o in outcomes.items()}, "probs": p.tolist()})
    pucts, choice = _puct_device(cases)
    for d, c in enumerate(cases):
        outcomes = {int(a): o for a, o in c["outcomes"].items()}
        with np.errstate(all="ignore"):
            want = PO.pucts(c["legal"], outcomes, np.asarray(c["probs"], np.float32), 2.0)
        want_choice = PO.puct_choice(want)
        got = pucts[d, : len(want)]
        assert np.array_equal(np.isnan(got), np.isnan(want)), (d, got, want)
        ok = ~np.isnan(want)
        assert np.abs(got[ok] - want[ok]).max(initial=0.0) < 1e-12, (d, got, want)
        assert choice[d] == want_choice, (d, choice[d], want_choice)


def test_card_memory_matches_the_reference_agent_after_watching():
    """a14: the reference agent watched four turns of a 3-player game dealt from np.random.seed(7) and then held exactly
    mcs_exact.json["MC3"]["agent_available_after_watching"] (stale memory: cards played and swept within one step were never
    seen, agents/mcts.py:62-73).  Replays the same trajectory through (i) the drop-in env + the host BaseMCAgent and (ii) the
    batched env + k_mc_roots, and compares both with the reference's set and with the reference's observation."""
    m = json.load(open(os.path.join(GOLDEN, "mcs_exact.json")))["MC3"]
    P = m["P"]
    # (i) drop-in env seeded like the fixture generator (tests/golden/make_golden.py:396-415)
    np.random.seed(7)
    perm = np.arange(104, dtype=np.int32)
    np.random.shuffle(perm)
    np.random.seed(7)
    env = SechsNimmtEnv(P, verbose=False)
    states, legal = env.reset()
    agent = MCSAgent(mc_max=1, mc_per_card=1)
    # (ii) the same deal in the batched engine (two copies, to exercise more than one game per launch)
    benv = BatchedSechsNimmtEnv(2, P).reset_from_perm(np.stack([perm, perm]).astype(np.uint8))
    avail = torch.zeros((2, 16), dtype=torch.uint8, device="cuda")
    roots = torch.zeros((2, 64), dtype=torch.uint8, device="cuda")
    lib = N.lib()

    def device_memory(turn):
        N.check(lib.nimmt_mc_roots(N.ptr(benv.state), N.ptr(avail), N.ptr(roots), 2, P, 0, int(turn == 0), N.current_stream()), "nimmt_mc_roots")
        words = avail.cpu().numpy().view(np.uint32)
        return [sorted(c for c in range(104) if (int(w[c >> 5]) >> (c & 31)) & 1) for w in words]

    for t in range(4):
        st = torch.tensor(states[0], dtype=torch.float)
        if len(legal[0]) == agent.handsize:
            agent._initialize_game(st)
        agent._memorize_cards(st, list(map(int, legal[0])))
        dev_sets = device_memory(t)
        assert dev_sets[0] == dev_sets[1] == sorted(map(int, agent.available_cards)), t
        a = [int(l[(3 * t + p) % len(l)]) for p, l in enumerate(legal)]
        (states, legal), _, _, _ = env.step(a)
        benv.step(torch.tensor([a, a], dtype=torch.uint8, device="cuda"), check=True)
    assert list(map(int, states[0])) == m["state"] and list(map(int, legal[0])) == m["legal"]      # the env reproduced the reference's game
    agent._memorize_cards(torch.tensor(states[0], dtype=torch.float), list(map(int, legal[0])))
    assert sorted(map(int, agent.available_cards)) == m["agent_available_after_watching"]
    dev_sets = device_memory(4)
    assert dev_sets[0] == dev_sets[1] == m["agent_available_after_watching"]
    # the stale set differs from what a fresh agent would believe at this state (the fixture's "available"): the test has teeth
    assert m["agent_available_after_watching"] != m["available"]
    # and the root the kernel would search is the one the host agent would pack
    want = R.pack_root_from_state(np.asarray(states[0], np.float32), list(map(int, legal[0])), agent.available_cards)
    assert (roots.cpu().numpy()[0] == want).all()


def test_masked_policy_probabilities_on_device():
    """f3: MaskedReinforceAgent.forward (agents/policy.py:45-60) — the 47 -> 100 -> 100 -> 104 net, normalisation without the
    action feature, softmax over the cards in hand — evaluated on the GPU for every golden state, against the reference's own
    probabilities (tests/golden/masked_policy.npz).  fp32 GEMMs: tolerance 2e-5 absolute."""
    from torch import nn
    z = np.load(os.path.join(GOLDEN, "masked_policy.npz"))

    class Net(nn.Module):                                 # the reference's MultiHeadedMLP(47, (100, 100), (104,)) parameter tree
        def __init__(self):
            super().__init__()
            self.latent_net = nn.Sequential(nn.Linear(47, 100), nn.ReLU(), nn.Linear(100, 100), nn.ReLU())
            self.head_nets = nn.ModuleList([nn.Sequential(nn.Linear(100, 104))])

        def forward(self, x):
            h = self.latent_net(x)
            return [head(h) for head in self.head_nets]

    net = Net()
    net.load_state_dict({k[len("w_actor_"):].replace("latent_net_0_", "latent_net.0.").replace("latent_net_2_", "latent_net.2.")
                         .replace("head_nets_0_0_", "head_nets.0.0."): torch.from_numpy(z[k]) for k in z.files if k.startswith("w_actor_")})
    net = net.cuda()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for obs in (torch.from_numpy(z["states"]).to(torch.float32).cuda(), torch.from_numpy(z["states"]).cuda()):   # float and int8 observations
            with torch.no_grad():
                probs = PL.masked_card_probs(net, obs).cpu().numpy()
            assert probs.shape == z["probs"].shape
            assert np.abs(probs - z["probs"]).max() < 2e-5, np.abs(probs - z["probs"]).max()
            assert ((probs > 0).sum(axis=1) == z["n_legal"]).all()
            assert np.allclose(probs.sum(axis=1), 1.0, atol=1e-5)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    # the same net on the tcgen05 tile (nimmt_masked_probs: bf16 operands, fp32 accumulate): against the reference's fp32
    # probabilities to 2e-3, on the golden states and on 5,000 states of real games (three tiles and a ragged one, every hand size)
    blob = PL.pack_masked_weights(net)
    assert blob is not None
    obs8 = torch.from_numpy(z["states"]).cuda()
    probs, logits = PL.masked_probs(obs8, blob, want_logits=True)
    assert np.abs(probs.cpu().numpy() - z["probs"]).max() < 2e-3, np.abs(probs.cpu().numpy() - z["probs"]).max()
    assert ((probs > 0).sum(dim=1).cpu().numpy() == z["n_legal"]).all()
    env = BatchedSechsNimmtEnv(500, 4, seed=3).reset()
    states = []
    for t in range(10):
        states.append(env.observe(dtype=torch.int8)[:, t % 4].clone())
        env.step_random()
    states = torch.cat(states)[:4999].contiguous()
    with torch.no_grad():
        want = PL.masked_card_probs(net, states)
    got = PL.masked_probs(states, blob)
    assert float((got - want).abs().max()) < 2e-3, float((got - want).abs().max())
    assert bool(((got > 0) == (states[:, :10] >= 0)).all()) and torch.allclose(got.sum(dim=1), torch.ones(4999, device="cuda"), atol=1e-5)


def test_full_size_replay_against_the_oracle_1m_games():
    """BASELINE configs[1] at its full size: 2^20 four-player games dealt and played with the device RNG (k_deal,
    k_random_actions, k_step_smem), then replayed through the C oracle — rewards, done, hands, boards and scores of every game
    at every turn compared bit for bit (the observation kernel is what reads the device state back, so it is covered too)."""
    n, P = 1 << 20, 4
    env = BatchedSechsNimmtEnv(n, P, seed=20261018).reset()
    obs0 = env.observe(dtype=torch.int8).cpu().numpy()
    hands0, board0 = obs0[:, :, :10].copy(), obs0[:, 0, -24:].reshape(n, 4, 6).copy()
    acts = np.zeros((n, 10, P), np.int8)
    rewards = np.zeros((n, 10, P), np.int8)
    done = np.zeros((n, 10), np.uint8)
    hands = np.zeros((n, 10, P, 10), np.int8)
    boards = np.zeros((n, 10, 4, 6), np.int8)
    scores = np.zeros((n, 10, P), np.int16)
    for t in range(10):
        a = env.random_actions().clone()
        rew, dn = env.step(a)
        acts[:, t] = a.cpu().numpy().view(np.int8)
        rewards[:, t], done[:, t] = rew.cpu().numpy(), dn.cpu().numpy()
        obs = env.observe(dtype=torch.int8).cpu().numpy()
        hands[:, t], boards[:, t] = obs[:, :, :10], obs[:, 0, -24:].reshape(n, 4, 6)
        scores[:, t] = env.scores().cpu().numpy()
        assert not bool(env.illegal.any())
    want = oracle.replay(P, board0, hands0, acts, want_obs=False)
    assert not want["illegal"].any()
    for k, got in (("rewards", rewards), ("done", done), ("hands", hands), ("boards", boards), ("scores", scores)):
        assert np.array_equal(got, want[k]), k


@pytest.mark.parametrize("P", [2, 4, 10])
def test_free_row_choice_on_device(P):
    """§8f row 4: the optional row_choice="agent" mode (the reference's TODO, env.py:156) — nimmt_step_choice through
    k_step_tiles<P, false, true> (whole tiles) and k_step (ragged tail) against the oracle's list-based extension: random row
    choices, every byte of every turn; invalid choices reject the step and leave the game untouched."""
    n, T = 4096 + 19, 10
    env = BatchedSechsNimmtEnv(n, P, seed=77 + P).reset()
    obs0 = env.observe(dtype=torch.int8).cpu().numpy()
    hands0, board0 = obs0[:, :, :10].copy(), obs0[:, 0, -24:].reshape(n, 4, 6).copy()
    gen = torch.Generator(device="cuda").manual_seed(P)
    # a rejected step first: every 50th game names row 4, every 77th row 255; those games stay as dealt, the others move on
    a = env.random_actions().clone()
    r = torch.randint(0, 4, (n, P), generator=gen, device="cuda", dtype=torch.uint8)
    r[::50, 0] = 4
    r[::77, P - 1] = 255
    rew, dn = env.step(a, rows=r)
    one = oracle.replay(P, board0, hands0, a.cpu().numpy().view(np.int8)[:, None], want_obs=False, row_choice=r.cpu().numpy().view(np.int8)[:, None])
    bad = np.zeros(n, bool)
    bad[::50] = True
    bad[::77] = True
    assert (env.illegal.cpu().numpy() == bad).all() and (one["illegal"][:, 0] == bad).all()
    obs = env.observe(dtype=torch.int8).cpu().numpy()
    assert np.array_equal(rew.cpu().numpy(), one["rewards"][:, 0]) and np.array_equal(obs[:, :, :10], one["hands"][:, 0])
    assert np.array_equal(obs[:, 0, -24:].reshape(n, 4, 6), one["boards"][:, 0])
    assert np.array_equal(obs[bad][:, :, :10], hands0[bad])                              # untouched
    # then whole games with valid random choices
    env.reset()
    obs0 = env.observe(dtype=torch.int8).cpu().numpy()
    hands0, board0 = obs0[:, :, :10].copy(), obs0[:, 0, -24:].reshape(n, 4, 6).copy()
    acts, rows = np.zeros((n, T, P), np.int8), np.zeros((n, T, P), np.int8)
    out = dict(rewards=np.zeros((n, T, P), np.int8), done=np.zeros((n, T), np.uint8), illegal=np.zeros((n, T), np.uint8),
               hands=np.zeros((n, T, P, 10), np.int8), boards=np.zeros((n, T, 4, 6), np.int8), scores=np.zeros((n, T, P), np.int16))
    for t in range(T):
        a = env.random_actions().clone()
        r = torch.randint(0, 4, (n, P), generator=gen, device="cuda", dtype=torch.uint8)
        rew, dn = env.step(a, rows=r)
        acts[:, t], rows[:, t] = a.cpu().numpy().view(np.int8), r.cpu().numpy().view(np.int8)
        out["rewards"][:, t], out["done"][:, t], out["illegal"][:, t] = rew.cpu().numpy(), dn.cpu().numpy(), env.illegal.cpu().numpy()
        obs = env.observe(dtype=torch.int8).cpu().numpy()
        out["hands"][:, t], out["boards"][:, t], out["scores"][:, t] = obs[:, :, :10], obs[:, 0, -24:].reshape(n, 4, 6), env.scores().cpu().numpy()
    want = oracle.replay(P, board0, hands0, acts, want_obs=False, row_choice=rows)
    plain = oracle.replay(P, board0, hands0, acts, want_obs=False)
    assert (want["rewards"] != plain["rewards"]).any() and not want["illegal"].any()
    for k in ("rewards", "done", "illegal", "hands", "boards", "scores"):
        assert np.array_equal(out[k], want[k]), k


def test_free_row_choice_dropin():
    """The B = 1 drop-in in agent mode (nimmt_step1 with rows): same game as the oracle, InvalidMoveException on a bad row."""
    from rl_6_nimmt_b200.env import InvalidMoveException
    np.random.seed(3)
    env = SechsNimmtEnv(3, verbose=False, row_choice="agent")
    states, legal = env.reset()
    board0 = np.array([[r + [-1] * (6 - len(r)) for r in env._board]], np.int8)
    hands0 = np.array([[h + [-1] * (10 - len(h)) for h in env._hands]], np.int8)
    acts, rows, rewards = [], [], []
    rng = np.random.RandomState(1)
    with pytest.raises(InvalidMoveException):
        env.step([l[0] for l in legal], rows=[0, 4, 0])
    assert env._hands == [[int(c) for c in h if c >= 0] for h in hands0[0]]          # untouched
    for t in range(10):
        a = [l[rng.randint(len(l))] for l in legal]
        r = rng.randint(0, 4, 3).tolist()
        (states, legal), rew, done, _ = env.step(a, rows=r)
        acts.append(a); rows.append(r); rewards.append(rew.tolist())
    assert done
    want = oracle.replay(3, board0, hands0, np.array([acts], np.int8), row_choice=np.array([rows], np.int8))
    assert np.array_equal(np.array([rewards], np.int8), want["rewards"])
    assert np.array_equal(np.stack(states).astype(np.int8), want["obs"][0, -1])
    with pytest.raises(AssertionError):
        SechsNimmtEnv(3, verbose=False).step([0, 1, 2], rows=[0, 0, 0])             # rows need row_choice="agent"


def _elo_reference(scores, ratings, agents, k):
    """Plain restatement of multi_elo.calc_elo as Tournament._compute_elos calls it (tournament.py:157-164): pairwise Elo,
    K = k / (n - 1), S by place, all seats updated from the ratings before the game; games in order."""
    ratings = ratings.copy()
    hist = np.zeros(scores.shape, np.float64)
    for b in range(len(scores)):
        n = scores.shape[1]
        old = [ratings[agents[b, p]] for p in range(n)]
        for i in range(n):
            d = 0.0
            for j in range(n):
                if i != j:
                    s = 1.0 if scores[b, i] > scores[b, j] else (0.5 if scores[b, i] == scores[b, j] else 0.0)
                    d += s - 1.0 / (1.0 + 10.0 ** ((old[j] - old[i]) / 400.0))
            ratings[agents[b, i]] = old[i] + k / (n - 1) * d
            hist[b, i] = ratings[agents[b, i]]
    return ratings, hist


def test_elo_scan_on_device():
    """§8f row 4: Elo on the device (nimmt_elo_scan), sequential over the games like Tournament.score_game.  Known answers: two
    equal players, k = 32 -> +-16; a draw between equals changes nothing; then random multi-player games with ties and a
    seat -> agent map against the plain restatement above (1e-9), and the batched session's statistics."""
    from rl_6_nimmt_b200 import stats as S
    r = S.elo_scan(torch.tensor([[-3, -10]], device="cuda"))
    assert np.allclose(r.cpu().numpy(), [1616.0, 1584.0])
    r = S.elo_scan(torch.tensor([[-5, -5, -5]], device="cuda"))
    assert np.allclose(r.cpu().numpy(), [1600.0] * 3)
    rng = np.random.RandomState(0)
    for P, A in ((2, 2), (4, 7), (10, 12)):
        B = 300
        scores = -rng.randint(0, 12, size=(B, P)).astype(np.int32)          # small range: plenty of ties
        agents = np.stack([rng.choice(A, P, replace=False) for _ in range(B)]).astype(np.int32)
        start = 1600.0 + 50.0 * rng.randn(A)
        got, hist = S.elo_scan(torch.as_tensor(scores).cuda(), ratings=torch.as_tensor(start.copy()).cuda(), agents=torch.as_tensor(agents).cuda(),
                               k=24.0, want_history=True)
        want, want_hist = _elo_reference(scores, start, agents, 24.0)
        assert np.abs(got.cpu().numpy() - want).max() < 1e-9 and np.abs(hist.cpu().numpy() - want_hist).max() < 1e-9
        assert abs(got.sum().item() - start.sum()) < 1e-6                  # pairwise Elo is zero-sum
    from rl_6_nimmt_b200.play import BatchedGameSession, MCSSeat, RandomSeat
    sess = BatchedGameSession([MCSSeat(mc_max=50), RandomSeat(), RandomSeat()], 512, seed=2)
    sess.play_games()
    elo = sess.statistics()["elo"].cpu().numpy()
    assert elo[0] > 1600 > max(elo[1], elo[2])                              # the searching seat gains rating against two DrunkHamsters


@pytest.mark.parametrize("P,n", [(4, 2048), (2, 96), (10, 1024), (4, 48), (7, 33 * 32)])
def test_multi_turn_launch_equals_single_steps(P, n):
    """nimmt_step_many / nimmt_step_random_many (the tile stays in shared memory for T turns) against T single launches: every
    output byte and the final state identical; T = 1, 3, 7 and the whole game; whole tiles and a ragged batch (n = 48: one
    launch per turn); an illegal turn in the middle leaves its games untouched and the later turns see the unchanged hands."""
    ref = BatchedSechsNimmtEnv(n, P, seed=9).reset()
    tape = torch.empty((10, n, P), dtype=torch.uint8, device="cuda")
    want_rew = torch.empty((10, n, P), dtype=torch.int8, device="cuda")
    want_done = torch.empty((10, n), dtype=torch.uint8, device="cuda")
    for t in range(10):
        ref.random_actions(out=tape[t])
        rew, dn = ref.step(tape[t])
        want_rew[t], want_done[t] = rew, dn
    want_obs = ref.observe(dtype=torch.int8)
    for split in ((10,), (3, 7), (1, 2, 3, 4)):
        env = BatchedSechsNimmtEnv(n, P, seed=9).reset()
        t0 = 0
        for T in split:
            rew, dn, ill = env.step_many(tape[t0:t0 + T])
            assert torch.equal(rew, want_rew[t0:t0 + T]) and torch.equal(dn, want_done[t0:t0 + T]) and not bool(ill.any())
            t0 += T
        assert torch.equal(env.observe(dtype=torch.int8), want_obs) and torch.equal(env.scores(), ref.scores())
    # an illegal turn (turn 2 replayed twice: its cards are gone) is rejected per game and does not disturb what follows
    env = BatchedSechsNimmtEnv(n, P, seed=9).reset()
    seq = torch.stack([tape[0], tape[1], tape[2], tape[2], tape[3]])
    rew, dn, ill = env.step_many(seq)
    assert bool(ill[3].all()) and not bool(ill[[0, 1, 2, 4]].any()) and bool((rew[3] == 0).all())
    assert torch.equal(rew[4], want_rew[3]) and torch.equal(rew[2], want_rew[2])
    # fused random play: the same cards, rewards and final state as step_random called turn by turn
    a = BatchedSechsNimmtEnv(n, P, seed=13, game0=777).reset()
    b = BatchedSechsNimmtEnv(n, P, seed=13, game0=777).reset()
    rew_b, done_b, acts_b = b.step_random_many(4, record_actions=True)
    for t in range(4):
        rew, dn = a.step_random(record_actions=True)
        assert torch.equal(rew, rew_b[t]) and torch.equal(dn, done_b[t]) and torch.equal(a._actions, acts_b[t])
    rew_b, done_b, _ = b.step_random_many(6)
    for t in range(6):
        rew, dn = a.step_random()
        assert torch.equal(rew, rew_b[t]) and torch.equal(dn, done_b[t])
    assert bool(done_b[5].all()) and torch.equal(a.observe(dtype=torch.int8), b.observe(dtype=torch.int8))


@pytest.mark.parametrize("P,n", [(4, 4096), (2, 64), (10, 1024), (5, 96)])
def test_packed_transfer_format_equals_the_byte_format(P, n):
    """nimmt_step_packed (4-bit hand slots in, one bit record per game out — the PCIe-frugal form of the end-to-end path) against
    nimmt_step on the same games: rewards, done, illegal and the state after every turn; a replayed turn (cards already played)
    and out-of-range slots are rejected per game."""
    a = BatchedSechsNimmtEnv(n, P, seed=21).reset()
    b = BatchedSechsNimmtEnv(n, P, seed=21).reset()
    dealt = a.observe(dtype=torch.int8)[:, :, :10].clone()
    assert a.packed_sizes() == ((P + 1) // 2, (5 * P + 2 + 7) // 8)
    for t in range(10):
        cards = a.random_actions().clone()
        rew, dn = a.step(cards)
        packed = BatchedSechsNimmtEnv.pack_slots(cards, dealt)
        if t == 4:
            # slot 15 for player 0 of every third game: those games are rejected and untouched, the others step; the second call
            # with the right slots steps exactly the rejected ones (the others' cards are gone: rejected in turn)
            bad = packed.clone()
            bad[::3, 0] |= 0x0F
            r1, d1, i1 = b.unpack_results(b.step_packed(bad))
            b.turn -= 1
            want_bad = torch.zeros(n, dtype=torch.bool, device="cuda")
            want_bad[::3] = True
            assert torch.equal(i1, want_bad) and bool((r1[want_bad] == 0).all())
            r2, d2, i2 = b.unpack_results(b.step_packed(packed))
            assert torch.equal(i2, ~want_bad) and bool((r2[~want_bad] == 0).all())
            rew_p, done_p = torch.where(want_bad.unsqueeze(1), r2, r1), torch.where(want_bad, d2, d1)
        else:
            rew_p, done_p, ill_p = b.unpack_results(b.step_packed(packed))
            assert not bool(ill_p.any())
        assert torch.equal(rew_p, rew) and torch.equal(done_p, dn.bool())
        assert torch.equal(a.observe(dtype=torch.int8), b.observe(dtype=torch.int8)) and torch.equal(a.scores(), b.scores())
    assert bool(done_p.all())
    # host buffers: one H2D + one D2H per step
    c = BatchedSechsNimmtEnv(n, P, seed=21).reset()
    ab, rb = c.packed_sizes()
    h_in, h_out = torch.empty((n, ab), dtype=torch.uint8).pin_memory(), torch.empty((n, rb), dtype=torch.uint8).pin_memory()
    ref = BatchedSechsNimmtEnv(n, P, seed=21).reset()
    cards = ref.random_actions().clone()
    rew, dn = ref.step(cards)
    h_in.copy_(BatchedSechsNimmtEnv.pack_slots(cards, dealt).cpu())
    c.step_host_packed(h_in, h_out)
    torch.cuda.synchronize()
    rew_p, done_p, ill_p = c.unpack_results(h_out)
    assert torch.equal(rew_p, rew.cpu()) and not bool(ill_p.any())
