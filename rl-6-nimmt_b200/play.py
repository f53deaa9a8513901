"""Batched game sessions: B games × mixed agents in lock-step on the GPU.

The batched counterpart of ``rl_6_nimmt/play.py:GameSession.play_game`` (:23-75): reset, then ten
turns of [every seat chooses a card → env.step], returning per-game score arrays in the format of
``session.results`` (negative Hornochsen totals per seat, play.py:69-74).  Seats are described by
small spec objects instead of per-game Python agents:

    RandomSeat()                                   DrunkHamster (agents/random.py)
    MCSSeat(mc_per_card=10, mc_max=100)            MCSAgent (agents/mcts.py:180-188)
    PolicySeat(net, mc_max=100, puct=True)         PUCTAgent / PolicyMCSAgent (agents/mcts.py:191-323)
    ReinforceSeat(net)                             BatchedReinforceAgent.forward (agents/policy.py:137-156), inference:
                                                   the same 48-100-100-1 net, a card sampled from its softmax, no search
    MaskedPolicySeat(net)                          MaskedReinforceAgent.forward (agents/policy.py:45-60): a 47 -> 104 net,
                                                   softmax over the cards in hand (greedy=True: a DQN's argmax)

Everything — card memory, root construction, rollouts, the decision rule, the env step — runs in
kernels; the host only sequences launches.  ``PolicySeat(..., learn=True)`` closes the Alpha0.5 self-play loop
(SURVEY.md §8f row 2): the seat's observations and chosen cards of every turn stay on the device and, when the
games end, one batched imitation step (train.py; agents/mcts.py:230-261) updates the net and the packed weights.
Seats that share one ``net`` object share one optimizer and take one step on all their decisions.
"""
import math

import torch

from . import _native as N
from . import policy as PL
from . import rollouts as R
from . import stats as S
from . import train as T
from .env import BatchedSechsNimmtEnv


class RandomSeat:
    pass


class MCSSeat:
    def __init__(self, mc_per_card=10, mc_max=100, rollouts_per_card=None):
        self.mc_per_card, self.mc_max, self.rollouts_per_card = mc_per_card, mc_max, rollouts_per_card

    def n_mc(self, n_cards):            # BaseMCAgent._compute_n_mc (agents/mcts.py:105-106)
        return min(self.mc_max, self.mc_per_card * math.factorial(n_cards))

    def per_card(self, n_cards):
        if self.rollouts_per_card is not None:
            return int(self.rollouts_per_card)
        return max(1, -(-self.n_mc(n_cards) // n_cards))


class PolicySeat(MCSSeat):
    def __init__(self, net, mc_per_card=10, mc_max=100, puct=True, c_puct=2.0, learn=False):
        super().__init__(mc_per_card, mc_max)
        self.net = net
        self.weights = PL.pack_weights(net)
        self.root_rule = N.ROOT_PUCT if puct else N.ROOT_POLICY
        self.c_puct = c_puct
        self.learn = bool(learn)


class ReinforceSeat:
    """A model-free policy-gradient opponent at the table: one k_policy_probs launch per turn
    and a categorical draw per game; ``greedy=True`` plays the most probable card instead.  ``net`` is anything with
    the parameter names of the reference's ``MultiHeadedMLP(48, (100, 100), (1, ...))``: BatchedReinforceAgent's
    ``actor`` (agents/policy.py:134) or ACER's ``actor_critic`` (agents/actor_critic.py:44-46; its first head is the
    policy logit, the value head is ignored)."""

    def __init__(self, net, greedy=False):
        self.net, self.weights, self.greedy = net, PL.pack_weights(net), bool(greedy)


class MaskedPolicySeat:
    """A state-only net with one output per card at the table (MaskedReinforceAgent.forward, agents/policy.py:45-60; a DQN's
    Q-values with ``greedy=True``): normalise the observation, one forward pass of ``net`` for all games, restrict to the cards
    in hand, softmax, sample or take the argmax.  Nets of the reference's shape, MultiHeadedMLP(47, (100, 100), (104,)), run on
    the tcgen05 tile (nimmt_masked_probs, csrc/masked_policy.cu); any other ``net`` is evaluated by PyTorch on the device."""

    def __init__(self, net, greedy=False):
        self.net, self.greedy = net, bool(greedy)
        self.weights = PL.pack_masked_weights(net)


class BatchedGameSession:
    """``data_parallel=True`` (under torch.distributed, one process per GPU): every rank plays its own ``num_games`` games and
    all ranks train ONE net — the learning nets are broadcast from rank 0 at construction and their gradients are averaged
    over the ranks at every step (train.allreduce_gradients), so the replicas never diverge.  Give every rank its own ``seed``."""

    def __init__(self, seats, num_games, device=None, seed=0, data_parallel=False):
        self.seats = list(seats)
        self.data_parallel = bool(data_parallel)
        self.env = BatchedSechsNimmtEnv(num_games, len(self.seats), device=device, seed=seed)
        self.lib, self.seed = N.lib(), int(seed)
        B, dev = num_games, self.env.device
        # every searching seat has its own buffers and its own stream: the seats' searches of one turn are independent
        # (each reads the state and writes its own column of the action matrix), and a latency-bound search kernel
        # (k_policy_rollouts: one CTA per few trees, four warps) leaves most of an SM idle for the next seat's CTAs
        mc = [p for p, s in enumerate(self.seats) if isinstance(s, MCSSeat)]
        self._roots = {p: torch.zeros((B, R.ROOT_BYTES), dtype=torch.uint8, device=dev) for p in mc}
        self._stats = {p: torch.zeros((B, R.MAX_ACTIONS, 3), dtype=torch.int64, device=dev) for p in mc}
        self._available = {p: torch.zeros((B, 16), dtype=torch.uint8, device=dev) for p in mc}
        self._streams = {p: torch.cuda.Stream(device=dev) for p in mc}
        self._generator = torch.Generator(device=dev)
        self._generator.manual_seed(self.seed)
        self.results = []   # one int32 [B, P] tensor of (negative) totals per play_games() call
        self.losses = []    # one 0-d device tensor per (play_games() call, learning net): mean per-episode imitation loss
        self.games = 0
        # learning seats grouped by net: one Adam (agents/base.py:29-33 defaults) per distinct net
        self._learners = {}
        for p, s in enumerate(self.seats):
            if isinstance(s, PolicySeat) and s.learn:
                s.net.to(dev)
                if id(s.net) not in self._learners and self.data_parallel:
                    T.broadcast_parameters(s.net)
                group = self._learners.setdefault(id(s.net), {"net": s.net, "seats": [], "optimizer": torch.optim.Adam(s.net.parameters())})
                group["seats"].append(p)
        if self.data_parallel:
            for group in self._learners.values():
                packed = PL.pack_weights(group["net"], device=dev)
                for p in group["seats"]:
                    self.seats[p].weights = packed

    def _stream(self):
        return torch.cuda.current_stream(self.env.device).cuda_stream

    def statistics(self):
        """Per-seat mean score, mean relative position and win rate over every game played so far (the reference
        tournament's per-agent statistics, tournament.py:139-155, 240-256; stats.py)."""
        return S.summary(torch.cat(self.results))

    def play_games(self):
        env, B, P = self.env, self.env.num_games, self.env.num_players
        env.reset()
        totals = torch.zeros((B, P), dtype=torch.int32, device=env.device)
        learning = sorted(p for g in self._learners.values() for p in g["seats"])
        seen_obs = {p: [] for p in learning}       # per turn int8 [B,47]
        seen_slot = {p: [] for p in learning}      # per turn [B]: hand slot of the chosen card
        for turn in range(10):
            n_cards = 10 - turn
            actions = env.random_actions()                       # every seat; MC seats are overwritten below
            obs8 = None
            for p, seat in enumerate(self.seats):
                if isinstance(seat, RandomSeat):
                    continue
                if isinstance(seat, (ReinforceSeat, MaskedPolicySeat)):
                    if obs8 is None:
                        obs8 = env.observe(dtype=torch.int8)
                    mine = obs8[:, p].contiguous()
                    if isinstance(seat, MaskedPolicySeat):
                        if seat.weights is not None:
                            probs = PL.masked_probs(mine, seat.weights.to(env.device))
                        else:
                            with torch.no_grad():
                                probs = PL.masked_card_probs(seat.net.to(env.device), mine)
                    else:
                        probs = PL.policy_probs(mine, seat.weights)                # [B,10] by hand slot, 0 for empty slots
                    slot = probs.argmax(dim=1, keepdim=True) if seat.greedy else torch.multinomial(probs, 1, generator=self._generator)
                    actions[:, p] = mine[:, :10].gather(1, slot).squeeze(1).to(torch.uint8)
                    continue
                main, side = torch.cuda.current_stream(env.device), self._streams[p]
                side.wait_stream(main)                               # the state and the action matrix are ready
                with torch.cuda.device(env.device), torch.cuda.stream(side):
                    roots, stats = self._roots[p], self._stats[p]
                    N.check(self.lib.nimmt_mc_roots(N.ptr(env.state), N.ptr(self._available[p]), N.ptr(roots), B, P, p,
                                                    int(turn == 0), side.cuda_stream), "nimmt_mc_roots")
                    stats.zero_()
                    seed = (self.seed * 1000003 + self.games * 131 + turn * 17 + p) & (2**64 - 1)
                    if n_cards > 1:                                # a single card is played without search (agents/mcts.py:52-53)
                        if isinstance(seat, PolicySeat):
                            found, _ = R.policy_rollouts(roots, P, seat.weights, seat.n_mc(n_cards), c_puct=seat.c_puct,
                                                         root_rule=seat.root_rule, seed=seed, device=env.device)
                            stats.copy_(found)
                        else:
                            R.mcs_rollouts(roots, P, seat.per_card(n_cards), seed=seed, out=stats, device=env.device)
                    N.check(self.lib.nimmt_mc_choose(N.ptr(env.state), N.ptr(stats), N.ptr(actions), B, P, p, side.cuda_stream),
                            "nimmt_mc_choose")
            for side in self._streams.values():                      # join: every seat has written its column
                torch.cuda.current_stream(env.device).wait_stream(side)
            if learning:
                obs = env.observe(dtype=torch.int8)
                for p in learning:
                    seen_obs[p].append(obs[:, p].clone())
                    seen_slot[p].append((obs[:, p, :10] == actions[:, p].to(torch.int8).unsqueeze(1)).to(torch.uint8).argmax(dim=1))
            rewards, done = env.step(actions)
            totals += rewards.to(torch.int32)
        for group in self._learners.values():   # PolicyMCSAgent.learn at episode end, all episodes of all the net's seats at once
            obs = torch.cat([o for p in group["seats"] for o in seen_obs[p]])
            slot = torch.cat([c for p in group["seats"] for c in seen_slot[p]])
            self.losses.append(T.imitation_step(group["net"], group["optimizer"], obs, slot, episodes=B * len(group["seats"]),
                                                data_parallel=self.data_parallel))
            packed = PL.pack_weights(group["net"], device=env.device)
            for p in group["seats"]:
                self.seats[p].weights = packed
        self.results.append(totals)
        self.games += B
        return totals
