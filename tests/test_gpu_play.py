"""GPU tests of the batched game session (many games x mixed agents in lock-step)."""
import numpy as np
import pytest
import torch

import rl_6_nimmt_b200  # noqa: F401
from rl_6_nimmt_b200 import _native as N
from rl_6_nimmt_b200 import rollouts as R
from rl_6_nimmt_b200.agents import MCSAgent
from rl_6_nimmt_b200.play import BatchedGameSession, MaskedPolicySeat, MCSSeat, PolicySeat, RandomSeat, ReinforceSeat
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


def test_alpha05_selfplay_loop_learns_on_device():
    """four PUCT seats sharing one net play 64 games in lock-step, then one batched imitation step
    (agents/mcts.py:230-261) runs on the device and the searches of the next games use the re-packed weights."""
    from rl_6_nimmt_b200 import policy as PL
    from rl_6_nimmt_b200 import train as T
    torch.manual_seed(0)
    net = PL.PolicyNet()
    seats = [PolicySeat(net, mc_max=40, puct=True, learn=True) for _ in range(4)]
    session = BatchedGameSession(seats, 64, seed=5)
    before = [p.detach().clone() for p in net.parameters()]
    blob0 = seats[0].weights.clone()
    totals = session.play_games()
    assert totals.shape == (64, 4) and int(totals.max()) <= 0
    assert len(session.losses) == 1
    loss = float(session.losses[0])
    assert 5.0 < loss < 25.0, loss                     # nine informative turns, ~log(n) each under a fresh policy
    assert any(not torch.equal(b, p.detach()) for b, p in zip(before, net.parameters()))
    assert all(s.weights is seats[0].weights for s in seats) and not torch.equal(blob0, seats[0].weights)
    session.play_games()
    assert len(session.losses) == 2 and np.isfinite(float(session.losses[1]))
    # the torch training path and the tcgen05 inference path are the same policy: probabilities of the chosen cards agree
    env = session.env.reset(seed=9)
    obs = env.observe(dtype=torch.int8)[:, 0].contiguous()
    slot = torch.zeros(obs.shape[0], dtype=torch.int64, device=obs.device)
    with torch.no_grad():
        want = T.imitation_log_probs(net, obs, slot).exp()
    got = PL.policy_probs(obs, seats[0].weights)[:, 0]
    assert float((want - got).abs().max()) < 1e-3


def test_imitation_reduces_the_loss_on_fixed_decisions():
    """Repeated imitation steps on one batch of (state, chosen card) pairs drive the loss down (Adam defaults)."""
    from rl_6_nimmt_b200 import policy as PL
    from rl_6_nimmt_b200 import train as T
    from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv
    torch.manual_seed(1)
    net = PL.PolicyNet().cuda()
    opt = torch.optim.Adam(net.parameters())
    env = BatchedSechsNimmtEnv(2048, 4, seed=3).reset()
    obs = env.observe(dtype=torch.int8)[:, 0].contiguous()
    slot = torch.full((2048,), 9, dtype=torch.int64, device=obs.device)      # always the highest card
    first = float(T.imitation_step(net, opt, obs, slot, episodes=2048))
    for _ in range(60):
        last = float(T.imitation_step(net, opt, obs, slot, episodes=2048))
    assert abs(first - np.log(10)) < 0.3 and last < 0.5 * first, (first, last)


def test_reinforce_seat_samples_from_the_policy():
    """a BatchedReinforceAgent-style seat (agents/policy.py:137-156) plays from the net's softmax.
    All games are forced to one root, so the empirical frequency of the first cards must match k_policy_probs."""
    from rl_6_nimmt_b200 import policy as PL
    torch.manual_seed(3)
    net = PL.PolicyNet()
    with torch.no_grad():
        net.head_nets[0][0].weight *= 60.0              # a policy that is far from uniform
    B = 20000
    sess = BatchedGameSession([ReinforceSeat(net), RandomSeat(), RandomSeat()], B, seed=2)
    env = sess.env.reset(seed=5)
    obs0 = env.observe(dtype=torch.int8)
    board = obs0[0, 0, -24:].reshape(4, 6).cpu().numpy()
    hands = obs0[0, :, :10].cpu().numpy()
    env.reset_to(np.repeat(board[None], B, 0), np.repeat(hands[None], B, 0))
    obs = env.observe(dtype=torch.int8)[:, 0].contiguous()
    probs = PL.policy_probs(obs, sess.seats[0].weights)[0].cpu().numpy()
    assert probs.max() - probs.min() > 0.03             # the test has power (noise: 5 sigma = 0.018)
    slot = torch.multinomial(PL.policy_probs(obs, sess.seats[0].weights), 1, generator=sess._generator).squeeze(1)
    freq = np.bincount(slot.cpu().numpy(), minlength=10) / B
    assert np.abs(freq - probs).max() < 5 * np.sqrt(0.25 / B)
    totals = sess.play_games()                          # a full session runs and every card played was legal
    assert totals.shape == (B, 3) and int(env.illegal.sum()) == 0
    greedy = BatchedGameSession([ReinforceSeat(net, greedy=True), RandomSeat()], 256, seed=1).play_games()
    assert greedy.shape == (256, 2)


def test_session_statistics_on_device():
    """BatchedGameSession.statistics(): the tournament's per-agent numbers (mean score, relative position, win rate) over
    all games so far; an MCS seat must beat two random seats on every one of them."""
    sess = BatchedGameSession([MCSSeat(mc_max=100), RandomSeat(), RandomSeat()], 4096, seed=3)
    sess.play_games()
    sess.play_games()
    st = {k: v.cpu().numpy() for k, v in sess.statistics().items()}
    assert st["mean_score"].shape == (3,) and abs(st["win_rate"].sum() - 1.0) < 1e-9
    assert st["mean_score"][0] > st["mean_score"][1:].max() + 2.0
    assert st["mean_relative_position"][0] > 0.6 > st["mean_relative_position"][1:].max()
    assert st["win_rate"][0] > 0.45


def test_masked_policy_seat_plays_legal_cards():
    """A 47 -> 104 state-only net (MaskedReinforceAgent / DQN shape) at the table, sampled and greedy: every card played is
    in the player's hand and the session finishes."""
    from torch import nn

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.latent_net = nn.Sequential(nn.Linear(47, 100), nn.ReLU(), nn.Linear(100, 100), nn.ReLU())
            self.head_nets = nn.ModuleList([nn.Sequential(nn.Linear(100, 104))])

        def forward(self, x):
            h = self.latent_net(x)
            return [head(h) for head in self.head_nets]

    torch.manual_seed(2)
    net = Net()
    for greedy in (False, True):
        sess = BatchedGameSession([MaskedPolicySeat(net, greedy=greedy), RandomSeat(), MCSSeat(mc_max=20)], 2048, seed=6)
        totals = sess.play_games()
        assert totals.shape == (2048, 3) and int(sess.env.illegal.sum()) == 0 and bool(sess.env.done.all())
