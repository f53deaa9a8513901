"""Per-game position statistics of a batch of finished games — the batched form of the reference tournament's
bookkeeping (``rl_6_nimmt/tournament.py``), for the evaluation harness of SURVEY.md §8f row 1.

``scores`` are the session results: int [B, P] negative Hornochsen totals (``play.py:69-74``).  Everything is a handful
of comparisons on the device; no host loop over games.  Pinned by ``tests/golden/position_stats.json`` (generated from the
unmodified reference).  Elo (``Tournament._compute_elos``, tournament.py:157-164) runs on the device too (``elo_scan``); the
``multi_elo`` package it calls is third-party and absent from the reference tree, so that one formula is stated, unit-tested
against a plain restatement, and "parity unpinned".
"""
import torch

from . import _native as N

ELO_INITIAL, ELO_K = 1600.0, 32.0     # Tournament.__init__ defaults (tournament.py:11)


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
    out = {"mean_score": scores.to(torch.float64).mean(dim=0), "mean_relative_position": relative_positions(scores).to(torch.float64).mean(dim=0),
           "win_rate": wins}
    if scores.is_cuda:
        out["elo"] = elo_scan(scores)     # every seat from 1600, k = 32, the games in the order they were played
    return out


def elo_scan(scores, ratings=None, agents=None, k=ELO_K, want_history=False):
    """``Tournament._compute_elos`` (tournament.py:157-164) applied to B finished games in order, on the device
    (nimmt_elo_scan; formula in include/nimmt_b200.h).  scores: int [B,P] session results.  ratings: float64 [A] device tensor,
    updated in place (default: ``ELO_INITIAL`` for P slots); agents: int32 [B,P] rating slot of every seat (default seat p =
    slot p).  Returns ratings, or (ratings, history [B,P]) with ``want_history``."""
    if not scores.is_cuda:
        raise N.NimmtNativeError("elo_scan needs device tensors; there is no CPU fallback")
    B, P = scores.shape
    scores = scores.to(torch.int32).contiguous()
    if ratings is None:
        ratings = torch.full((P,), ELO_INITIAL, dtype=torch.float64, device=scores.device)
    assert ratings.dtype == torch.float64 and ratings.is_cuda and ratings.is_contiguous()
    if agents is not None:
        agents = agents.to(device=scores.device, dtype=torch.int32).contiguous()
        assert agents.shape == (B, P)
    history = torch.empty((B, P), dtype=torch.float64, device=scores.device) if want_history else None
    with torch.cuda.device(scores.device):
        N.check(N.lib().nimmt_elo_scan(N.ptr(scores), N.ptr(agents), N.ptr(ratings), B, P, float(k), N.ptr(history),
                                       torch.cuda.current_stream(scores.device).cuda_stream), "nimmt_elo_scan")
    return (ratings, history) if want_history else ratings
