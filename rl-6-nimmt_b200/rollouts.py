"""Host side of the Monte-Carlo rollout kernels: root packing, launches, the one collective.

Reference: BaseMCAgent._mcts / _draw_env / _deal_hands / _play_out (agents/mcts.py:91-154).
The kernels never see Python objects: a decision is a 64-byte ``nimmt_root`` (include/nimmt_b200.h).
"""
import numpy as np
import torch

from . import _native as N

MAX_ACTIONS = 10
ROOT_BYTES = 64


def cards_to_mask(cards):
    """104-bit card set as 4 little-endian uint32 words."""
    words = [0, 0, 0, 0]
    for c in cards:
        c = int(c)
        if not 0 <= c < 104:
            raise ValueError(f"card {c} out of range")
        words[c >> 5] |= 1 << (c & 31)
    return words


def pack_root(board, own_hand, available_cards, num_players):
    """Builds one nimmt_root image (uint8[64]).

    board: list of 4 lists of cards (oldest first), as BaseMCAgent._board_from_state(flatten=False)
    returns (mcts.py:75-85); own_hand: legal_actions; available_cards: BaseMCAgent.available_cards.
    """
    root = np.zeros(ROOT_BYTES, np.uint8)
    w = root.view(np.uint32)
    w[0:4] = cards_to_mask(own_hand)
    w[4:8] = cards_to_mask(available_cards)
    rows = np.full((4, 6), 255, np.uint8)
    assert len(board) == 4
    for r, cards in enumerate(board):
        assert 1 <= len(cards) <= 5, "a row holds 1..5 cards between placements (env.py:133,170)"
        rows[r, : len(cards)] = cards
    root[32:56] = rows.reshape(-1)
    root[56] = num_players
    return root


def pack_root_from_state(state, legal_actions, available_cards):
    """Root from an observation vector (env.py:174-212 layout) as the agents receive it."""
    st = np.asarray(state.detach().cpu() if isinstance(state, torch.Tensor) else state)
    grid = st[-24:].reshape(4, 6)
    board = [[int(c) for c in row if c >= 0] for row in grid]
    return pack_root(board, legal_actions, available_cards, int(st[10]))


def allreduce_stats(stats, group=None):
    """The path's only collective: integer SUM of the [D,10,3] (sum, sumsq, count)
    table over the ranks that each played a stripe of the rollouts.  Integer addition is order
    independent, so every world size yields the same table.  No-op without torch.distributed."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def mcs_rollouts(roots, num_players, rollouts_per_action, seed=0, rank=0, world=1, out=None, device=None):
    """Launches nimmt_mcs_rollouts.  roots: uint8 [D,64] (numpy or tensor).  Returns the device
    tensor int64 [D,10,3] of this rank's stripe (not yet reduced over ranks)."""
    if not torch.cuda.is_available():
        raise N.NimmtNativeError("mcs_rollouts needs a CUDA device; there is no CPU fallback")
    lib = N.lib()
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    if not isinstance(roots, torch.Tensor):
        roots = torch.as_tensor(np.ascontiguousarray(roots, dtype=np.uint8))
    roots = roots.to(device).contiguous()
    assert roots.dtype == torch.uint8 and roots.dim() == 2 and roots.shape[1] == ROOT_BYTES
    D = roots.shape[0]
    if out is None:
        out = torch.zeros((D, MAX_ACTIONS, 3), dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        N.check(lib.nimmt_mcs_rollouts(N.ptr(roots), D, int(num_players), int(rollouts_per_action), int(seed) & (2**64 - 1),
                                       int(rank), int(world), N.ptr(out), torch.cuda.current_stream(device).cuda_stream),
                "nimmt_mcs_rollouts")
    return out


def policy_rollouts(roots, num_players, weights, n_mc, c_puct=2.0, root_rule=N.ROOT_PUCT, seed=0, device=None):
    """Launches nimmt_policy_rollouts (PolicyMCSAgent / PUCTAgent searches, one per root).
    Returns (stats int64 [D,10,3], root_probs float32 [D,10]) on the device."""
    if not torch.cuda.is_available():
        raise N.NimmtNativeError("policy_rollouts needs a CUDA device; there is no CPU fallback")
    lib = N.lib()
    device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
    if not isinstance(roots, torch.Tensor):
        roots = torch.as_tensor(np.ascontiguousarray(roots, dtype=np.uint8))
    roots = roots.to(device).contiguous()
    assert roots.dtype == torch.uint8 and roots.dim() == 2 and roots.shape[1] == ROOT_BYTES
    D = roots.shape[0]
    stats = torch.zeros((D, MAX_ACTIONS, 3), dtype=torch.int64, device=device)
    probs = torch.zeros((D, MAX_ACTIONS), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        N.check(lib.nimmt_policy_rollouts(N.ptr(roots), D, int(num_players), N.ptr(weights), int(n_mc), float(c_puct), int(root_rule),
                                          int(seed) & (2**64 - 1), N.ptr(stats), N.ptr(probs),
                                          torch.cuda.current_stream(device).cuda_stream), "nimmt_policy_rollouts")
    return stats, probs


# Below this many rollouts in a decision batch, striping them over the ranks loses: the collective is pure latency (an
# NCCL all-reduce of 240 bytes costs ~30 us on NVLink, measured in SCALE_r01: one root x 10 cards x 10,000 rollouts took
# 36 us on one GPU and 58-66 us on 2-8), while the kernel is one wave of threads whatever the count (10^5 rollouts occupy a
# third of one B200's resident threads).  Every rank then plays the WHOLE batch itself: same roots, same seed, same integers
# on every rank — no communication at all, the latency of one GPU, and still the table the sharded path would produce.
MIN_ROLLOUTS_TO_SHARD = 1_000_000


def sharded_mcs_rollouts(roots, num_players, rollouts_per_action, seed=0, group=None, device=None, min_rollouts_to_shard=None):
    """One decision batch on all ranks of the process group.  Large batches are striped over the ranks (rollout ids
    ``id mod world``) and summed with the path's only collective; small ones (fewer than ``min_rollouts_to_shard``
    rollouts in total, default MIN_ROLLOUTS_TO_SHARD) are played redundantly by every rank, which is faster than any
    exchange and bit-identical (integer sums of the same rollout ids)."""
    import torch.distributed as dist
    rank, world = 0, 1
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    threshold = MIN_ROLLOUTS_TO_SHARD if min_rollouts_to_shard is None else int(min_rollouts_to_shard)
    total = int(len(roots)) * MAX_ACTIONS * int(rollouts_per_action)
    if world == 1 or total < threshold:
        return mcs_rollouts(roots, num_players, rollouts_per_action, seed, 0, 1, device=device)
    stats = mcs_rollouts(roots, num_players, rollouts_per_action, seed, rank, world, device=device)
    return allreduce_stats(stats, group)


def choose_from_stats(legal_actions, stats):
    """BaseMCAgent._choose_action_from_outcomes (mcts.py:156-165) on (sum, sumsq, count) rows:
    argmax of the per-card mean, strict '>' scanning cards in ascending order (first maximum wins);
    a card with no rollouts has mean NaN and can never win."""
    best_action, best_mean = legal_actions[0], -float("inf")
    means = []
    for i, action in enumerate(legal_actions):
        s, _, n = (int(x) for x in stats[i])
        mean = s / n if n > 0 else float("nan")
        means.append(mean)
        if mean > best_mean:
            best_action, best_mean = action, mean
    return best_action, means
