"""Per-game position statistics of a batch of finished games — the batched form of the reference tournament's
bookkeeping (``rl_6_nimmt/tournament.py``), for the evaluation harness of SURVEY.md §8f row 1.

``scores`` are the session results: int [B, P] negative Hornochsen totals (``play.py:69-74``).  Everything is a handful
of comparisons on the device; no host loop over games.  Pinned by ``tests/golden/position_stats.json`` (generated from the
unmodified reference).  Elo (``multi_elo``, an unpinned third-party package absent from the reference tree) is not built.
"""
import torch


def absolute_positions(scores):
    """``Tournament._compute_absolute_positions`` (tournament.py:240-247): tie-averaged rank, best player first.
    (The reference's docstring says 0 = best; what its code returns — and this reproduces — is 1 for a sole winner.)"""
    s = scores.unsqueeze(2)             # [B, P, 1]: the player
    o = scores.unsqueeze(1)             # [B, 1, P]: everybody
    better = (o > s).sum(dim=2)         # searchsorted(sorted(-scores), -score - 0.5)
    at_least = (o >= s).sum(dim=2)      # searchsorted(sorted(-scores), -score + 0.5)
    return 0.5 * (better + 1.0 + at_least).to(torch.float32)


def relative_positions(scores):
    """``Tournament._compute_relative_positions`` (tournament.py:249-256): 1 = best, 0 = worst, ties averaged."""
    s = scores.unsqueeze(2)
    o = scores.unsqueeze(1)
    at_most = (o <= s).sum(dim=2)       # searchsorted(sorted(scores), score + 0.5)
    worse = (o < s).sum(dim=2)          # searchsorted(sorted(scores), score - 0.5)
    pos = 0.5 * (at_most + 1.0 + worse).to(torch.float32)
    return (pos - 1.0) / (scores.shape[1] - 1)


def winners(scores):
    """``np.argmax(scores)`` (tournament.py:141): the first seat with the best score."""
    return scores.argmax(dim=1)


def summary(scores):
    """What ``Tournament.score_game`` / ``baseline_eval`` accumulate per agent (tournament.py:139-155, 182-195), averaged
    over the batch: mean score, mean relative position and win rate per seat, as float64 tensors [P] on the device."""
    P = scores.shape[1]
    wins = torch.nn.functional.one_hot(winners(scores), P).to(torch.float64).mean(dim=0)
    return {"mean_score": scores.to(torch.float64).mean(dim=0), "mean_relative_position": relative_positions(scores).to(torch.float64).mean(dim=0),
            "win_rate": wins}
