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


def _world(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


def broadcast_parameters(net, src=0, group=None):
    """Data-parallel self-play starts from ONE net: rank ``src``'s parameters go to every rank (one flat buffer, one
    broadcast).  No-op without a process group."""
    import torch.distributed as dist
    if _world(group) == 1:
        return
    params = [p.data for p in net.parameters()]
    flat = torch.cat([p.reshape(-1) for p in params])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    for p in params:
        p.copy_(flat[off:off + p.numel()].view_as(p))
        off += p.numel()


def allreduce_gradients(net, group=None):
    """The data-parallel exchange of the self-play loop: the 15,101 gradient entries of the policy net
    as ONE flat buffer, summed over the ranks (NCCL over NVLink: a 60 KB message, latency-bound) and divided by the world
    size, so that every rank applies the gradient of the mean loss over ALL ranks' episodes and the replicas stay
    identical.  No-op without a process group."""
    import torch.distributed as dist
    world = _world(group)
    if world == 1:
        return
    grads = [p.grad for p in net.parameters() if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= world
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def imitation_step(net, optimizer, obs, chosen_slot, episodes=1, data_parallel=False, group=None):
    """One Adam step on ``-sum log pi / episodes`` (``_train`` + ``_gradient_step``, agents/mcts.py:245-261).
    ``data_parallel``: every rank of the process group holds a replica of ``net`` and its own episodes; the gradients are
    averaged over the ranks before the step (allreduce_gradients), i.e. all GPUs train ONE net on the mean loss of all
    their episodes.  Returns this rank's loss as a 0-d tensor on the device (no host sync)."""
    loss = -imitation_log_probs(net, obs, chosen_slot).sum() / float(episodes)
    optimizer.zero_grad()
    loss.backward()
    if data_parallel:
        allreduce_gradients(net, group)
    optimizer.step()
    return loss.detach()
