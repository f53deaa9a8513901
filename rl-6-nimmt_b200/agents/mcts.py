"""Monte-Carlo search agents backed by the rollout kernels.

Drop-ins for rl_6_nimmt/agents/mcts.py: same constructor keywords, same ``forward`` protocol, same
card memory (including its staleness), same decision rule.  Only ``_mcts`` is
replaced: instead of n_mc Python playouts it launches one kernel that plays R rollouts for every
legal card and reads back a 240-byte table.
"""
import logging
import math

import numpy as np
import torch

from .. import _native as N
from .. import policy as PL
from .. import rollouts as R
from .base import Agent

logger = logging.getLogger(__name__)


class BaseMCAgent(Agent):
    def __init__(self, handsize=10, num_rows=4, num_cards=104, threshold=6, mc_per_card=10, mc_max=100,
                 include_summaries=True, rollouts_per_card=None, seed=None, shard_over_ranks=False, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if (handsize, num_rows, num_cards, threshold) != (10, 4, 104, 6):
            raise NotImplementedError("kernels are built for handsize=10, num_rows=4, num_cards=104, threshold=6")
        self.num_players = None  # inferred from the first state of a game (mcts.py:31, 87-89)
        self.handsize, self.num_rows, self.num_cards, self.threshold = handsize, num_rows, num_cards, threshold
        self.mc_per_card, self.mc_max = mc_per_card, mc_max
        self.include_summaries = include_summaries
        self.available_cards = []
        # extensions (not in the reference): explicit per-card budget, RNG seed, multi-GPU sharding
        self.rollouts_per_card = rollouts_per_card
        self.shard_over_ranks = shard_over_ranks
        self._seed = np.random.randint(0, 2**31 - 1) if seed is None else int(seed)
        self._decisions = 0
        self.last_stats = None  # int64 [n,3] of the last decision, for inspection

    # -- protocol ------------------------------------------------------------------------------
    def forward(self, state, legal_actions, *args, **kwargs):
        n = len(legal_actions)
        if n == self.handsize:                       # mcts.py:47-48
            self._initialize_game(state)
        self._memorize_cards(state, legal_actions)   # mcts.py:49
        if n == 1:                                   # mcts.py:52-53
            return legal_actions[0], {"log_prob": torch.tensor(0.0).to(self.device, self.dtype)}
        return self._mcts(legal_actions, state)

    def learn(self, *args, **kwargs):
        raise NotImplementedError

    # -- card memory (mcts.py:62-89), host side, kept verbatim in behaviour ---------------------
    def _initialize_game(self, state):
        self.available_cards = list(range(self.num_cards))
        self.num_players = self._num_players_from_state(state)

    def _memorize_cards(self, state, legal_actions):
        seen = set(int(c) for c in legal_actions) | set(self._board_from_state(state, flatten=True))
        self.available_cards = [c for c in self.available_cards if c not in seen]

    def _board_from_state(self, state, flatten=True):
        grid = np.asarray(state.detach().cpu() if isinstance(state, torch.Tensor) else state)[-self.num_rows * self.threshold:]
        rows = [[int(c) for c in row if c >= 0.0] for row in grid.reshape(self.num_rows, self.threshold)]
        return [c for row in rows for c in row] if flatten else rows

    @staticmethod
    def _num_players_from_state(state):
        return int(state[10])

    # -- search ----------------------------------------------------------------------------------
    def _compute_n_mc(self, n_actions):
        return min(self.mc_max, self.mc_per_card * math.factorial(n_actions))   # mcts.py:105-106

    def _rollouts_per_card(self, n_actions):
        """The reference spreads n_mc rollouts over the cards at random (mcts.py:140-145); the kernel
        gives every card the same share, rounded up, so the total is never below the reference's."""
        if self.rollouts_per_card is not None:
            return int(self.rollouts_per_card)
        return max(1, -(-self._compute_n_mc(n_actions) // n_actions))

    def _mcts(self, legal_actions, state):
        legal_actions = [int(a) for a in legal_actions]
        root = R.pack_root_from_state(state, legal_actions, self.available_cards)
        per_card = self._rollouts_per_card(len(legal_actions))
        seed = (self._seed * 0x9E3779B1 + self._decisions) & (2**64 - 1)
        self._decisions += 1
        run = R.sharded_mcs_rollouts if self.shard_over_ranks else R.mcs_rollouts
        stats = run(root[None], self.num_players, per_card, seed=seed)[0].cpu().numpy()
        self.last_stats = stats[: len(legal_actions)]
        action, means = R.choose_from_stats(legal_actions, self.last_stats)
        if logger.isEnabledFor(logging.DEBUG):
            logger.debug("AlphaAlmostZero thoughts:")
            for a, m, row in zip(legal_actions, means, self.last_stats):
                logger.debug(f"  {'x' if a == action else ' '} {a + 1:>3d}: p = 1.00, n = {int(row[2]):>3d}, E[r] = {m:>5.1f}")
        return action, {"log_prob": torch.tensor(0.0).to(self.device, self.dtype)}

    def _choose_action_from_outcomes(self, outcomes, log_probs=None):
        """mcts.py:156-165 on explicit outcome lists (kept for callers that hold such dicts)."""
        best_action, best_mean = list(outcomes.keys())[0], -float("inf")
        for action, outcome in outcomes.items():
            mean = float(np.mean(outcome)) if len(outcome) else float("nan")
            if mean > best_mean:
                best_action, best_mean = action, mean
        info = {"log_prob": log_probs[best_action][0] if log_probs and log_probs.get(best_action) else torch.tensor(0.0)}
        return best_action, info


class MCSAgent(BaseMCAgent):
    """Monte-Carlo search with uniformly random playouts by all players (mcts.py:180-188)."""

    def learn(self, *args, **kwargs):
        pass


class PolicyMCSAgent(BaseMCAgent):
    """Monte-Carlo search whose playouts follow a learnable policy (agents/mcts.py:191-261).

    The policy net keeps the reference's parameter names (``actor.latent_net.0.weight`` ...), so state
    dicts are interchangeable.  Searches run in k_policy_rollouts (tcgen05 policy net + env dynamics on
    chip); the imitation update of ``learn`` stays in PyTorch autograd, as in the reference.
    """

    root_rule = N.ROOT_POLICY

    def __init__(self, hidden_sizes=(100, 100), activation=None, r_factor=0.1, **kwargs):
        super().__init__(**kwargs)
        self.r_factor = r_factor
        self.actor = PL.PolicyNet(self.state_length + 1, hidden_sizes)
        self._packed, self._packed_key = None, None
        self.last_root_probs = None

    def _weights(self):
        key = tuple(int(p._version) for p in self.actor.parameters())   # bumped by every optimizer step / load
        if self._packed is None or key != self._packed_key:
            self._packed, self._packed_key = PL.pack_weights(self.actor), key
        return self._packed

    def _compute_policy(self, legal_actions, state):
        return PL.torch_policy(self.actor, state, legal_actions)

    def _search_kwargs(self):
        return {}

    def _mcts(self, legal_actions, state):
        legal_actions = [int(a) for a in legal_actions]
        root = R.pack_root_from_state(state, legal_actions, self.available_cards)
        n_mc = self._compute_n_mc(len(legal_actions)) if self.rollouts_per_card is None else self.rollouts_per_card * len(legal_actions)
        seed = (self._seed * 0x9E3779B1 + self._decisions) & (2**64 - 1)
        self._decisions += 1
        stats, probs = R.policy_rollouts(root[None], self.num_players, self._weights(), n_mc, root_rule=self.root_rule, seed=seed,
                                         **self._search_kwargs())
        self.last_stats = stats[0].cpu().numpy()[: len(legal_actions)]
        self.last_root_probs = probs[0].cpu().numpy()[: len(legal_actions)]
        action, means = R.choose_from_stats(legal_actions, self.last_stats)          # mcts.py:156-165
        # info["log_prob"] = log pi(chosen card | root) (mcts.py:165, 215, 293), with autograd for learn()
        log_prob = torch.log(self._compute_policy(legal_actions, state)[legal_actions.index(action)])
        if logger.isEnabledFor(logging.DEBUG):
            logger.debug("AlphaAlmostZero thoughts:")
            for a, m, row, p in zip(legal_actions, means, self.last_stats, self.last_root_probs):
                logger.debug(f"  {'x' if a == action else ' '} {a + 1:>3d}: p = {p:.2f}, n = {int(row[2]):>3d}, E[r] = {m:>5.1f}")
        return action, {"log_prob": log_prob}

    def learn(self, state, reward, action, done, next_state, next_reward, episode_end, num_episode, legal_actions, *args, **kwargs):
        self.history.store(log_prob=kwargs["log_prob"], reward=reward * self.r_factor)      # mcts.py:232
        if not episode_end or not self.training:
            return 0.0
        loss = self._train()
        self.history.clear()
        return loss

    def _train(self):
        log_probs = torch.stack([lp for lp in self.history.rollout()["log_prob"] if lp.requires_grad] or [torch.zeros((), requires_grad=True)], dim=0)
        loss = -torch.sum(log_probs)           # imitate the (deterministic) search choice, mcts.py:246-249
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss.item()


class PUCTAgent(PolicyMCSAgent):
    """"Alpha0.5" (agents/mcts.py:264-323): PUCT over the root statistics, policy playouts below."""

    root_rule = N.ROOT_PUCT

    def __init__(self, c_puct=2.0, temperature=None, **kwargs):
        super().__init__(**kwargs)
        self.c_puct = c_puct
        self.temperature = temperature

    def _search_kwargs(self):
        return {"c_puct": self.c_puct}

    def _mcts(self, legal_actions, state):
        if self.temperature is not None and self.temperature > 1.0e-12:
            raise NotImplementedError                 # as mcts.py:318-323
        return super()._mcts(legal_actions, state)
