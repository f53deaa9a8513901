"""Batched game sessions: B games × mixed agents in lock-step on the GPU.

The batched counterpart of ``rl_6_nimmt/play.py:GameSession.play_game`` (:23-75): reset, then ten
turns of [every seat chooses a card → env.step], returning per-game score arrays in the format of
``session.results`` (negative Hornochsen totals per seat, play.py:69-74).  Seats are described by
small spec objects instead of per-game Python agents:

    RandomSeat()                                   DrunkHamster (agents/random.py)
    MCSSeat(mc_per_card=10, mc_max=100)            MCSAgent (agents/mcts.py:180-188)
    PolicySeat(net, mc_max=100, puct=True)         PUCTAgent / PolicyMCSAgent (agents/mcts.py:191-323), inference only

Everything — card memory, root construction, rollouts, the decision rule, the env step — runs in
kernels; the host only sequences launches.  Learning is not part of this harness.
"""
import math

import torch

from . import _native as N
from . import policy as PL
from . import rollouts as R
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
    def __init__(self, net, mc_per_card=10, mc_max=100, puct=True, c_puct=2.0):
        super().__init__(mc_per_card, mc_max)
        self.weights = PL.pack_weights(net)
        self.root_rule = N.ROOT_PUCT if puct else N.ROOT_POLICY
        self.c_puct = c_puct


class BatchedGameSession:
    def __init__(self, seats, num_games, device=None, seed=0):
        self.seats = list(seats)
        self.env = BatchedSechsNimmtEnv(num_games, len(self.seats), device=device, seed=seed)
        self.lib, self.seed = N.lib(), int(seed)
        B, dev = num_games, self.env.device
        self._roots = torch.zeros((B, R.ROOT_BYTES), dtype=torch.uint8, device=dev)
        self._stats = torch.zeros((B, R.MAX_ACTIONS, 3), dtype=torch.int64, device=dev)
        self._available = {p: torch.zeros((B, 16), dtype=torch.uint8, device=dev) for p, s in enumerate(self.seats) if not isinstance(s, RandomSeat)}
        self.results = []   # one int32 [B, P] tensor of (negative) totals per play_games() call
        self.games = 0

    def _stream(self):
        return torch.cuda.current_stream(self.env.device).cuda_stream

    def play_games(self):
        env, B, P = self.env, self.env.num_games, self.env.num_players
        env.reset()
        totals = torch.zeros((B, P), dtype=torch.int32, device=env.device)
        for turn in range(10):
            n_cards = 10 - turn
            actions = env.random_actions()                       # every seat; MC seats are overwritten below
            for p, seat in enumerate(self.seats):
                if isinstance(seat, RandomSeat):
                    continue
                with torch.cuda.device(env.device):
                    N.check(self.lib.nimmt_mc_roots(N.ptr(env.state), N.ptr(self._available[p]), N.ptr(self._roots), B, P, p,
                                                    int(turn == 0), self._stream()), "nimmt_mc_roots")
                    self._stats.zero_()
                    seed = (self.seed * 1000003 + self.games * 131 + turn * 17 + p) & (2**64 - 1)
                    if n_cards > 1:                                # a single card is played without search (agents/mcts.py:52-53)
                        if isinstance(seat, PolicySeat):
                            stats, _ = R.policy_rollouts(self._roots, P, seat.weights, seat.n_mc(n_cards), c_puct=seat.c_puct,
                                                         root_rule=seat.root_rule, seed=seed, device=env.device)
                            self._stats.copy_(stats)
                        else:
                            R.mcs_rollouts(self._roots, P, seat.per_card(n_cards), seed=seed, out=self._stats, device=env.device)
                    N.check(self.lib.nimmt_mc_choose(N.ptr(env.state), N.ptr(self._stats), N.ptr(actions), B, P, p, self._stream()),
                            "nimmt_mc_choose")
            rewards, done = env.step(actions)
            totals += rewards.to(torch.int32)
        self.results.append(totals)
        self.games += B
        return totals
