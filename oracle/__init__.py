"""CPU oracle for the 6 nimmt! hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product (rl-6-nimmt_b200/) never does; it has no CPU path.

Parity status: pinned against fixtures generated from the unmodified reference
(tests/golden/make_golden.py); see tests/test_oracle_golden.py.

  * nimmt_oracle.c  — C restatement of rl_6_nimmt/env.py dynamics + observations and of the
                      MCSAgent rollout loop of rl_6_nimmt/agents/mcts.py.
  * policy_oracle.py — numpy fp32 restatement of SechsNimmtStateNormalization +
                      MultiHeadedMLP + softmax and the PUCT root rule (float64, as numpy does
                      in the reference).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libnimmt_oracle.so")
_lib = None


def build(force=False):
    """Compiles oracle/nimmt_oracle.c with gcc (seconds)."""
    src = os.path.join(_HERE, "nimmt_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/libnimmt_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        i8p = ctypes.POINTER(ctypes.c_int8)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        i16p = ctypes.POINTER(ctypes.c_int16)
        i32p = ctypes.POINTER(ctypes.c_int)
        i64p = ctypes.POINTER(ctypes.c_int64)
        L.oracle_card_value.argtypes = [ctypes.c_int]
        L.oracle_replay.argtypes = [ctypes.c_int] * 4 + [i8p, i8p, i8p, i8p, u8p, u8p, i8p, i8p, i16p, i8p]
        L.oracle_replay_choice.argtypes = [ctypes.c_int] * 4 + [i8p, i8p, i8p, i8p, i8p, u8p, u8p, i8p, i8p, i16p, i8p]
        L.oracle_mcs_rollouts.argtypes = [ctypes.c_int, i32p, i32p, ctypes.c_int, i32p, ctypes.c_int,
                                          ctypes.c_int64, ctypes.c_uint64, ctypes.c_int, i64p]
        L.oracle_bench_env.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_uint64, i64p]
        L.oracle_bench_env.restype = ctypes.c_int64
        L.oracle_bench_mcs.argtypes = [ctypes.c_int, i32p, i32p, ctypes.c_int, i32p, ctypes.c_int,
                                       ctypes.c_int64, ctypes.c_int, ctypes.c_uint64, i64p]
        f32p = ctypes.POINTER(ctypes.c_float)
        L.oracle_policy_rollouts.argtypes = [ctypes.c_int, i32p, i32p, ctypes.c_int, i32p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64,
                                             ctypes.c_int, f32p, f32p, f32p, f32p, f32p, ctypes.c_float, i64p]
        L.oracle_policy_probs.argtypes = [ctypes.c_void_p, i32p, i32p, ctypes.c_int, f32p]
        L.oracle_deal_from_perm.argtypes = [ctypes.c_void_p, ctypes.c_int, i32p]
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def card_values():
    L = lib()
    return np.array([L.oracle_card_value(c) for c in range(104)], dtype=np.int32)


def deal_from_perm(perm, num_players):
    """env.py:99-112: returns (hands [P,10] ascending, rows [4]) for one shuffled deck."""
    perm = np.asarray(perm)
    hands = np.sort(perm[: 10 * num_players].reshape(num_players, 10), axis=1)
    rows = perm[[103, 102, 101, 100]]
    return hands.astype(np.int8), rows.astype(np.int8)


def replay(num_players, rows0, hands0, actions, include_summaries=True, want_obs=True, row_choice=None):
    """Replays games through the C oracle.  ``row_choice`` (int8 [n,T,P], values 0..3): the optional free-row-choice mode —
    the row a player takes if their card undercuts every row (env.py:156 TODO); None = the reference's lowest-penalty rule.

    rows0 [n,4,6] int8 (-1 padded), hands0 [n,P,10] int8 (-1 padded), actions [n,T,P] int8.
    Returns dict(rewards [n,T,P] i8, done [n,T] u8, illegal [n,T] u8, hands [n,T,P,10] i8,
    boards [n,T,4,6] i8, scores [n,T,P] i16, obs [n,T,P,L] i8) — state AFTER each step.
    An illegal step leaves the game untouched (env.py:68-69) and is flagged.
    """
    P = num_players
    rows0 = np.ascontiguousarray(rows0, dtype=np.int8)
    hands0 = np.ascontiguousarray(hands0, dtype=np.int8)
    actions = np.ascontiguousarray(actions, dtype=np.int8)
    n, T = actions.shape[0], actions.shape[1]
    assert rows0.shape == (n, 4, 6) and hands0.shape == (n, P, 10) and actions.shape == (n, T, P)
    L_obs = 47 if include_summaries else 35
    out = dict(
        rewards=np.zeros((n, T, P), np.int8), done=np.zeros((n, T), np.uint8), illegal=np.zeros((n, T), np.uint8),
        hands=np.zeros((n, T, P, 10), np.int8), boards=np.zeros((n, T, 4, 6), np.int8),
        scores=np.zeros((n, T, P), np.int16), obs=np.zeros((n, T, P, L_obs), np.int8) if want_obs else None,
    )
    if row_choice is not None:
        row_choice = np.ascontiguousarray(row_choice, dtype=np.int8)
        assert row_choice.shape == (n, T, P)
    rc = lib().oracle_replay_choice(
        P, n, T, int(include_summaries), _p(rows0, ctypes.c_int8), _p(hands0, ctypes.c_int8), _p(actions, ctypes.c_int8),
        _p(row_choice, ctypes.c_int8) if row_choice is not None else None,
        _p(out["rewards"], ctypes.c_int8), _p(out["done"], ctypes.c_uint8), _p(out["illegal"], ctypes.c_uint8),
        _p(out["hands"], ctypes.c_int8), _p(out["boards"], ctypes.c_int8), _p(out["scores"], ctypes.c_int16),
        _p(out["obs"], ctypes.c_int8) if want_obs else None,
    )
    if rc:
        raise ValueError(f"oracle_replay failed ({rc})")
    return out


def rows_from_singletons(row_cards):
    """[n,4] first cards -> [n,4,6] -1 padded boards."""
    row_cards = np.asarray(row_cards, dtype=np.int8)
    out = -np.ones(row_cards.shape + (6,), np.int8)
    out[..., 0] = row_cards
    return out


def mcs_rollouts(num_players, board, own, available, n_rollouts, seed=0, first_action=-1):
    """Reference-style MCS rollouts (agents/mcts.py:91-154). Returns int64 [len(own),3] = (sum, sumsq, count)."""
    rows = -np.ones((4, 6), np.int32)
    for r, cards in enumerate(board):
        rows[r, : len(cards)] = cards
    own = np.ascontiguousarray(own, dtype=np.int32)
    available = np.ascontiguousarray(available, dtype=np.int32)
    stats = np.zeros((len(own), 3), np.int64)
    rc = lib().oracle_mcs_rollouts(num_players, _p(rows, ctypes.c_int), _p(own, ctypes.c_int), len(own),
                                   _p(available, ctypes.c_int), len(available), int(n_rollouts), int(seed),
                                   int(first_action), _p(stats, ctypes.c_int64))
    if rc:
        raise ValueError(f"oracle_mcs_rollouts failed ({rc})")
    return stats


def bench_env(num_players, games_per_thread, n_threads, seed=1):
    cs = ctypes.c_int64(0)
    return int(lib().oracle_bench_env(num_players, int(games_per_thread), int(n_threads), int(seed), ctypes.byref(cs)))


def bench_mcs(num_players, board, own, available, rollouts_per_thread, n_threads, seed=1):
    rows = -np.ones((4, 6), np.int32)
    for r, cards in enumerate(board):
        rows[r, : len(cards)] = cards
    own = np.ascontiguousarray(own, dtype=np.int32)
    available = np.ascontiguousarray(available, dtype=np.int32)
    stats = np.zeros((len(own), 3), np.int64)
    rc = lib().oracle_bench_mcs(num_players, _p(rows, ctypes.c_int), _p(own, ctypes.c_int), len(own),
                                _p(available, ctypes.c_int), len(available), int(rollouts_per_thread), int(n_threads),
                                int(seed), _p(stats, ctypes.c_int64))
    if rc:
        raise ValueError(f"oracle_bench_mcs failed ({rc})")
    return stats


def policy_rollouts(num_players, board, own, available, n_rollouts, weights, seed=0, first_action=-1):
    """Reference-style PolicyMCSAgent rollouts in fp32 (agents/mcts.py:91-154, 209-228).
    weights: dict w1 [100,48], b1, w2 [100,100], b2, w3 [1,100] or [100], b3.  Returns int64 [len(own),3]."""
    rows = -np.ones((4, 6), np.int32)
    for r, cards in enumerate(board):
        rows[r, : len(cards)] = cards
    own = np.ascontiguousarray(own, dtype=np.int32)
    available = np.ascontiguousarray(available, dtype=np.int32)
    w = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in weights.items()}
    stats = np.zeros((len(own), 3), np.int64)
    rc = lib().oracle_policy_rollouts(num_players, _p(rows, ctypes.c_int), _p(own, ctypes.c_int), len(own), _p(available, ctypes.c_int),
                                      len(available), int(n_rollouts), int(seed), int(first_action),
                                      _p(w["w1"], ctypes.c_float), _p(w["b1"], ctypes.c_float), _p(w["w2"], ctypes.c_float),
                                      _p(w["b2"], ctypes.c_float), _p(w["w3"].reshape(-1), ctypes.c_float), float(w["b3"].reshape(-1)[0]),
                                      _p(stats, ctypes.c_int64))
    if rc:
        raise ValueError(f"oracle_policy_rollouts failed ({rc})")
    return stats
