"""GPU tests of the tcgen05 policy-net kernel (Alpha0.5 leaf evaluation) through the C ABI."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import policy_oracle as po

import rl_6_nimmt_b200  # noqa: F401
from rl_6_nimmt_b200 import policy as PL

pytestmark = pytest.mark.gpu


def _golden():
    z = np.load(os.path.join(GOLDEN, "policy_vectors.npz"))
    net = PL.PolicyNet()
    net.load_state_dict({k[len("w_actor_"):].replace("latent_net_0_", "latent_net.0.").replace("latent_net_2_", "latent_net.2.")
                         .replace("head_nets_0_0_", "head_nets.0.0."): torch.from_numpy(z[k]) for k in z.files if k.startswith("w_actor_")})
    return z, net


def _decisions(z):
    """Golden rows are [card | obs47] grouped per decision: recover (obs47, n_legal) per decision."""
    obs, off = [], 0
    for n in z["seg"]:
        rows = z["rows_in"][off:off + n]
        assert (rows[:, 1:] == rows[0, 1:]).all() and (rows[:, 0] == rows[0, 1:1 + n]).all()
        obs.append(rows[0, 1:])
        off += n
    return np.array(obs, np.int8)


def test_probs_match_reference_and_bf16_emulation():
    z, net = _golden()
    w = po.weights_from_golden(z)
    obs = _decisions(z)                                     # [264, 47]
    blob = PL.pack_weights(net)
    probs, logits = PL.policy_probs(torch.from_numpy(obs).cuda(), blob, want_logits=True)
    probs, logits = probs.cpu().numpy(), logits.cpu().numpy()
    off = 0
    worst_ref = worst_emu = worst_logit = 0.0
    for d, n in enumerate(z["seg"]):
        got = probs[d, :n]
        assert (probs[d, n:] == 0).all() and abs(got.sum() - 1.0) < 1e-5
        worst_ref = max(worst_ref, np.abs(got - z["probs"][off:off + n]).max())
        emu_l = po.policy_logits_bf16(z["rows_in"][off:off + n], w)
        worst_logit = max(worst_logit, np.abs(logits[d, :n] - emu_l).max())
        worst_emu = max(worst_emu, np.abs(got - po.softmax(emu_l)).max())
        off += n
    # vs the kernel's own arithmetic (bf16 operands, fp32 accumulate) emulated in numpy: accumulation order only
    assert worst_logit < 2e-4 and worst_emu < 5e-5, (worst_logit, worst_emu)
    # vs the fp32 reference (torch, unmodified reference code): stated tolerance for bf16 operands
    assert worst_ref < 1e-3, worst_ref


def test_ragged_batches_and_tile_boundaries():
    z, net = _golden()
    obs = torch.from_numpy(_decisions(z)).cuda()
    blob = PL.pack_weights(net)
    full = PL.policy_probs(obs, blob)
    for D in (1, 11, 12, 13, 25, 263):
        part = PL.policy_probs(obs[:D].clone(), blob)
        assert torch.equal(part, full[:D]), D
    big = obs.repeat(40, 1)                                  # 10,560 decisions: many tiles per CTA
    out = PL.policy_probs(big, blob)
    assert torch.equal(out, full.repeat(40, 1))
    # enough tiles per CTA (2, 3, 4, 7, 15, 60 on 148 SMs) that the producers' six-stage operand ring and the three tensor-memory
    # slots are reused many times, with odd and even tile counts per CTA and a ragged last tile
    for reps, cut in ((13, 0), (20, 5), (27, 0), (47, 11), (100, 7), (400, 1)):
        big = obs.repeat(reps, 1)
        D = big.shape[0] - cut
        out, logits = PL.policy_probs(big[:D].clone(), blob, want_logits=True)
        assert torch.equal(out, full.repeat(reps, 1)[:D]), (reps, cut)
        assert torch.isfinite(logits).all()


def test_torch_module_is_state_dict_compatible():
    z, net = _golden()
    rows = torch.from_numpy(z["rows_norm"])
    (logit,) = net(rows)
    np.testing.assert_allclose(logit.detach().numpy().reshape(-1), z["logits"], rtol=1e-4, atol=1e-5)


# ---------------------------------------------------------------------------------------------------
# policy rollouts (PolicyMCSAgent / PUCTAgent searches on chip)
# ---------------------------------------------------------------------------------------------------
import oracle  # noqa: E402
from rl_6_nimmt_b200 import _native as N  # noqa: E402
from rl_6_nimmt_b200 import rollouts as R  # noqa: E402


def _rollout_golden():
    z = np.load(os.path.join(GOLDEN, "policy_rollouts.npz"))
    net = PL.PolicyNet()
    net.load_state_dict({k[len("w_actor_"):].replace("latent_net_0_", "latent_net.0.").replace("latent_net_2_", "latent_net.2.")
                         .replace("head_nets_0_0_", "head_nets.0.0."): torch.from_numpy(z[k]) for k in z.files if k.startswith("w_actor_")})
    w = {"w1": z["w_actor_latent_net_0_weight"], "b1": z["w_actor_latent_net_0_bias"], "w2": z["w_actor_latent_net_2_weight"],
         "b2": z["w_actor_latent_net_2_bias"], "w3": z["w_actor_head_nets_0_0_weight"], "b3": z["w_actor_head_nets_0_0_bias"]}
    return z, net, w


def test_policy_rollouts_vs_reference_and_oracle():
    """Sharpened policy, 4-player mid-game root.  Stratified root (equal budget per card): per-card means
    z-tested (|z| < 4.5) against the unmodified reference's PolicyMCSAgent rollouts and against the fp32
    C oracle; root policy within 1e-3 of the reference's."""
    z, net, w = _rollout_golden()
    legal, avail = z["legal"].tolist(), z["available"].tolist()
    root = R.pack_root_from_state(z["state"], legal, avail)
    blob = PL.pack_weights(net)
    n_mc = 6 * 20_000
    stats, probs = R.policy_rollouts(root[None], 4, blob, n_mc, root_rule=N.ROOT_STRATIFIED, seed=7)
    stats, probs = stats[0].cpu().numpy(), probs[0].cpu().numpy()
    assert np.abs(probs[:6] - z["root_probs"]).max() < 1e-3 and (probs[6:] == 0).all()
    board = [[int(c) for c in row if c >= 0] for row in z["state"][-24:].reshape(4, 6)]
    want = oracle.policy_rollouts(4, board, legal, avail, 30_000, w, seed=5)
    for i, a in enumerate(legal):
        s, ss, n = (int(x) for x in stats[i])
        assert n == 20_000
        mean, var = s / n, ss / n - (s / n) ** 2
        zr = (mean - z["mean"][i]) / np.sqrt(var / n + z["var"][i] / z["count"][i])
        ws, wss, wn = (int(x) for x in want[i])
        wmean, wvar = ws / wn, wss / wn - (ws / wn) ** 2
        zo = (mean - wmean) / np.sqrt(var / n + wvar / wn)
        assert abs(zr) < 4.5 and abs(zo) < 4.5, (a, mean, z["mean"][i], wmean, zr, zo)
        assert abs(var - wvar) < 0.15 * wvar
    assert (stats[6:] == 0).all()


def test_policy_root_rule_follows_the_policy():
    z, net, w = _rollout_golden()
    legal = z["legal"].tolist()
    root = R.pack_root_from_state(z["state"], legal, z["available"].tolist())
    blob = PL.pack_weights(net)
    n_mc = 60_000
    stats, probs = R.policy_rollouts(root[None], 4, blob, n_mc, root_rule=N.ROOT_POLICY, seed=11)
    counts = stats[0, :6, 2].cpu().numpy()
    assert counts.sum() == n_mc
    assert np.abs(counts / n_mc - z["root_probs"]).max() < 5 * np.sqrt(0.2 / n_mc) + 1e-3


def test_puct_search_batch():
    """256 trees at once (BASELINE configs[3] shape): every tree spends exactly n_mc rollouts, visits
    concentrate on the better cards, and equal roots with different tree ids give different searches."""
    z, net, w = _rollout_golden()
    legal = z["legal"].tolist()
    root = R.pack_root_from_state(z["state"], legal, z["available"].tolist())
    blob = PL.pack_weights(net)
    roots = np.repeat(root[None], 256, axis=0)
    stats, probs = R.policy_rollouts(roots, 4, blob, 200, c_puct=2.0, root_rule=N.ROOT_PUCT, seed=3)
    stats = stats.cpu().numpy()
    assert (stats[:, :6, 2].sum(axis=1) == 200).all() and (stats[:, 6:] == 0).all()
    assert len({tuple(s[:6, 2]) for s in stats}) > 100                  # different seeds per tree
    visits = stats[:, :6, 2].mean(axis=0)
    # cards 82 / 87 (indices 4, 5) are clearly best at this root (reference: -7.0 vs -9 .. -10)
    assert visits[4] + visits[5] > visits[:4].sum() * 0.6, visits
    means = stats[:, :6, 0].sum(axis=0) / stats[:, :6, 2].sum(axis=0)
    assert np.argmax(means) in (4, 5)
    # PUCT prior phase: with fewer than 10 outcomes q_hat uses (0,-10,-5); first visit goes to the highest prior
    s1, _ = R.policy_rollouts(roots[:3], 4, blob, 1, root_rule=N.ROOT_PUCT, seed=3)
    assert (s1[:, int(np.argmax(z["root_probs"])), 2] == 1).all()


def test_concurrent_searches_on_four_streams_equal_serial_ones():
    """A self-play turn launches the four seats' searches on four streams: 4 x 86 CTAs of k_policy_rollouts share the SMs three to
    one (each CTA holds 128 + 32 tensor-memory columns, allocated in two steps).  A search depends on (seed, tree) only, so the
    tables of concurrent launches must equal, bit for bit, those of the same launches one after the other."""
    z, net, w = _rollout_golden()
    root = R.pack_root_from_state(z["state"], z["legal"].tolist(), z["available"].tolist())
    blob = PL.pack_weights(net)
    roots = torch.as_tensor(np.repeat(root[None], 256, axis=0)).cuda()
    serial = [R.policy_rollouts(roots, 4, blob, 120, c_puct=2.0, root_rule=N.ROOT_PUCT, seed=50 + i)[0].clone() for i in range(4)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(4)]
    for rep in range(3):
        got = []
        for i, st in enumerate(streams):
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                got.append(R.policy_rollouts(roots, 4, blob, 120, c_puct=2.0, root_rule=N.ROOT_PUCT, seed=50 + i)[0])
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        for i in range(4):
            assert torch.equal(got[i], serial[i]), (rep, i)


def test_alpha05_agent_dropin_plays_and_learns():
    from rl_6_nimmt_b200.agents import DrunkHamster, PUCTAgent
    from rl_6_nimmt_b200.env import SechsNimmtEnv
    torch.manual_seed(0)
    np.random.seed(1)
    agent = PUCTAgent(mc_max=60, seed=2)
    agent.train()
    before = [p.detach().clone() for p in agent.parameters()]
    agents = [agent, DrunkHamster(), DrunkHamster()]
    env = SechsNimmtEnv(3, verbose=False)
    states, legal = env.reset()
    done, rewards = False, np.zeros(3, np.int32)
    while not done:                                   # the reference's GameSession loop (play.py:23-75)
        acts, infos = [], []
        for ag, s, l in zip(agents, states, legal):
            a, info = ag(torch.tensor(s, dtype=torch.float), legal_actions=l)
            acts.append(int(a)); infos.append(info)
        (nstates, nlegal), nrew, done, _ = env.step(acts)
        for ag, a, s, ns, r, nr, info, l, nl in zip(agents, acts, states, nstates, rewards, nrew, infos, legal, nlegal):
            ag.learn(state=s, legal_actions=list(l), reward=r, action=a, done=done, next_state=ns, next_legal_actions=list(nl),
                     next_reward=nr, num_episode=0, episode_end=done, **info)
        states, legal, rewards = nstates, nlegal, nrew
    assert agent.last_stats[:, 2].sum() >= 20        # last search (2 cards): min(mc_max, 10 * 2!) = 20 rollouts
    assert any(not torch.equal(b, p.detach()) for b, p in zip(before, agent.parameters()))   # one Adam step happened
    key = agent._packed_key
    agent._weights()
    assert agent._packed_key != key                    # weights were repacked after the update


@pytest.mark.parametrize("P", (1, 2, 3, 5, 6, 10))
def test_policy_rollouts_other_table_sizes_vs_oracle(P):
    """The kernel packs floor(12 / P) trees into a CTA and maps rows (decision, slot) three decisions per warp; this runs
    every packing (12, 6, 4, 2, 2, 1 trees per CTA) on mid-game roots and z-tests the per-card outcome means against the
    fp32 C oracle of the reference's PolicyMCSAgent rollouts (sharpened policy).  Also: roots of different hand sizes in
    one launch, and more roots than fit one CTA."""
    from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv
    z, net, w = _rollout_golden()
    blob = PL.pack_weights(net)
    n_roots = 7
    env = BatchedSechsNimmtEnv(n_roots, P, seed=40 + P).reset()
    roots, metas = [], []
    # play a different number of turns per game by stepping everything and snapshotting game g after 2 + (g % 4) turns
    snaps = {}
    for t in range(6):
        obs = env.observe(dtype=torch.int8).cpu().numpy()
        for g in range(n_roots):
            if t == 2 + (g % 4):
                snaps[g] = obs[g].copy()
        env.step(env.random_actions().clone())
    for g in range(n_roots):
        o = snaps[g]
        board = [[int(c) for c in row if c >= 0] for row in o[0, -24:].reshape(4, 6)]
        own = [int(c) for c in o[0, :10] if c >= 0]
        seen = set(own) | {c for row in board for c in row}
        avail = [c for c in range(104) if c not in seen]            # a fresh agent's memory: everything not visible
        roots.append(R.pack_root(board, own, avail, P))
        metas.append((board, own, avail))
    n_per_card = 1500
    stats, _ = R.policy_rollouts(np.stack(roots), P, blob, 8 * n_per_card, root_rule=N.ROOT_STRATIFIED, seed=9)   # hands hold <= 8 cards
    stats = stats.cpu().numpy()
    for g, (board, own, avail) in enumerate(metas):
        n = len(own)
        got = stats[g]
        want = oracle.policy_rollouts(P, board, own, avail, n * (1200 if P <= 5 else 500), w, seed=3 + g)
        for i in range(n):
            s, ss, cnt = (int(x) for x in got[i])
            ws, wss, wn = (int(x) for x in want[i])
            # every root plays its own budget: n_mc is per launch, so smaller hands get at least n_per_card per card too
            assert cnt >= n_per_card and wn > 0
            mean, var = s / cnt, max(ss / cnt - (s / cnt) ** 2, 1e-9)
            wmean, wvar = ws / wn, max(wss / wn - (ws / wn) ** 2, 1e-9)
            zscore = (mean - wmean) / np.sqrt(var / cnt + wvar / wn)
            assert abs(zscore) < 5.0, (P, g, i, mean, wmean, zscore)
        assert (got[n:] == 0).all()
        if g >= 1:
            break                                                    # two roots per table size keep the oracle's CPU time short
