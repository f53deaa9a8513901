"""Alpha0.5 training step, batched over many self-play episodes.

The reference trains ``PolicyMCSAgent`` / ``PUCTAgent`` by imitation of the search
(``agents/mcts.py:230-261``): per episode ``loss = -sum_t log pi(a_t | s_t)`` where ``a_t`` is the card the
Monte-Carlo search chose and ``pi`` the policy net's softmax over the legal cards (``:209-228``); a turn with
a single legal card contributes the constant 0 (``:52-53``); one Adam step per episode (``agents/base.py:29-33``).

Here the same loss is evaluated for all decisions of all episodes in one pass: the decisions arrive as the int8
observations ``k_observe`` wrote and the hand-slot index of the chosen card, rows ``[card | obs47]`` are built with
one gather, and the three small GEMMs plus the masked log-softmax run in PyTorch autograd on the device (plain
library GEMMs; the searches that produce the targets run in ``k_policy_rollouts``).  With ``episodes = E`` the
batch loss is the MEAN over episodes of the reference's per-episode loss, so one batched Adam step has the
scale of one reference step; ``E = 1`` reproduces the reference's update exactly (tests/test_train.py).
"""
import torch

from .policy import normalize_rows


def decision_rows(obs):
    """obs [D,47] (any dtype) -> (rows float32 [D,10,48], legal bool [D,10]); row (d, s) = [obs[d,s] | obs[d]]
    (``agents/mcts.py:219-225``); slot s is legal when it holds a card (``env.py:209-210``)."""
    obs = obs.to(torch.float32)
    cards = obs[:, :10]
    rows = torch.cat((cards.unsqueeze(2), obs.unsqueeze(1).expand(-1, 10, -1)), dim=2)
    return rows, cards >= 0


def imitation_log_probs(net, obs, chosen_slot):
    """log pi(card in hand slot ``chosen_slot`` | state) for every decision, with autograd through ``net``.
    Decisions with one legal card give exactly 0 without a gradient, as the reference's n == 1 shortcut does."""
    rows, legal = decision_rows(obs)
    (logits,) = net(normalize_rows(rows.reshape(-1, 48)))
    logits = logits.reshape(-1, 10).masked_fill(~legal, float("-inf"))
    logp = torch.log_softmax(logits, dim=1).gather(1, chosen_slot.to(torch.int64).unsqueeze(1)).squeeze(1)
    return torch.where(legal.sum(dim=1) > 1, logp, torch.zeros_like(logp))


def imitation_step(net, optimizer, obs, chosen_slot, episodes=1):
    """One Adam step on ``-sum log pi / episodes`` (``_train`` + ``_gradient_step``, agents/mcts.py:245-261).
    Returns the loss as a 0-d tensor on the device (no host sync)."""
    loss = -imitation_log_probs(net, obs, chosen_slot).sum() / float(episodes)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss.detach()
