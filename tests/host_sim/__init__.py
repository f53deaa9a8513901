"""ctypes binding of the TEST-ONLY host build of the product's per-game logic (host_sim.cu)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libhost_sim.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "host_sim.cu")
        csrc = os.path.join(_HERE, "..", "..", "rl-6-nimmt_b200", "csrc")
        newest = max(os.path.getmtime(p) for p in [src] + [os.path.join(csrc, f) for f in ("game.cuh", "handrec.cuh", "step.cuh", "step_tile.cuh", "rollout.cuh", "puct.cuh")])
        if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < newest:
            subprocess.check_call(["nvcc", "-O1", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                                   "-Wno-deprecated-gpu-targets", "-o", _LIB, src])
        _lib = ctypes.CDLL(_LIB)
        _lib.sim_select.restype = ctypes.c_uint
        _lib.sim_select.argtypes = [ctypes.c_uint32] * 5
    return _lib


def set_form(form):
    """0: the per-game logic on 104-bit card sets (Game<P>); 1: on the stored hand records (GameRec<P>, handrec.cuh);
    2: in place in 32-game tile records through step_tile.cuh::step_lane, the per-lane code of k_step_tiles;
    3: the same with the compact transfer format (4-bit hand slots in, bit-packed results out; the slots are relative to hands0,
    so hands0 must be the hands as dealt, or whatever reset_to was given)."""
    lib().sim_set_form(int(form))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def replay(P, rows0, hands0, actions, row_choice=None):
    """row_choice (int8 [n,T,P]): the free-row-choice mode, as oracle.replay."""
    rows0 = np.ascontiguousarray(rows0, np.int8)
    hands0 = np.ascontiguousarray(hands0, np.int8)
    actions = np.ascontiguousarray(actions, np.int8)
    n, T = actions.shape[:2]
    out = dict(rewards=np.zeros((n, T, P), np.int8), done=np.zeros((n, T), np.uint8), illegal=np.zeros((n, T), np.uint8),
               hands=np.zeros((n, T, P, 10), np.int8), boards=np.zeros((n, T, 4, 6), np.int8), scores=np.zeros((n, T, P), np.int16))
    if row_choice is not None:
        row_choice = np.ascontiguousarray(row_choice, np.int8)
        assert row_choice.shape == actions.shape
    lib().sim_set_row_choice(_p(row_choice) if row_choice is not None else None)
    try:
        rc = lib().sim_replay(P, n, T, _p(rows0), _p(hands0), _p(actions), _p(out["rewards"]), _p(out["done"]), _p(out["illegal"]),
                              _p(out["hands"]), _p(out["boards"]), _p(out["scores"]))
    finally:
        lib().sim_set_row_choice(None)
    assert rc == 0
    return out


def play_random_tiles(P, rows0, hands0, turns, seed, game0=0):
    """Random-vs-random play through the fused per-lane step (step_lane<P, true>).  Returns replay()'s dict plus the cards drawn."""
    rows0 = np.ascontiguousarray(rows0, np.int8)
    hands0 = np.ascontiguousarray(hands0, np.int8)
    n, T = len(rows0), turns
    out = dict(actions=np.zeros((n, T, P), np.int8), rewards=np.zeros((n, T, P), np.int8), done=np.zeros((n, T), np.uint8),
               illegal=np.zeros((n, T), np.uint8), hands=np.zeros((n, T, P, 10), np.int8), boards=np.zeros((n, T, 4, 6), np.int8),
               scores=np.zeros((n, T, P), np.int16))
    rc = lib().sim_play_random_tiles(P, n, T, _p(rows0), _p(hands0), ctypes.c_uint64(seed), ctypes.c_uint64(game0), _p(out["actions"]),
                                     _p(out["rewards"]), _p(out["done"]), _p(out["illegal"]), _p(out["hands"]), _p(out["boards"]), _p(out["scores"]))
    assert rc == 0
    return out


def deal(P, n, seed, game0=0):
    hands = np.zeros((n, P, 10), np.int8)
    boards = np.zeros((n, 4, 6), np.int8)
    rc = lib().sim_deal(P, n, ctypes.c_uint64(seed), ctypes.c_uint64(game0), _p(hands), _p(boards))
    assert rc == 0
    return hands, boards


def random_actions(P, rows0, hands0, seed, turn, game0=0):
    rows0 = np.ascontiguousarray(rows0, np.int8)
    hands0 = np.ascontiguousarray(hands0, np.int8)
    n = len(rows0)
    act = np.zeros((n, P), np.uint8)
    rc = lib().sim_random_actions(P, n, _p(rows0), _p(hands0), ctypes.c_uint64(seed), ctypes.c_uint64(game0), ctypes.c_uint32(turn), _p(act))
    assert rc == 0
    return act


def mcs(P, root_bytes, rollouts, seed, rank=0, world=1):
    """root_bytes: 64-byte nimmt_root image. Returns int64 [10,3] (sum, sumsq, count) per first-card rank."""
    buf = (ctypes.c_uint8 * 64).from_buffer_copy(root_bytes)
    stats = np.zeros((10, 3), np.int64)
    rc = lib().sim_mcs(P, buf, ctypes.c_int64(rollouts), ctypes.c_uint64(seed), rank, world, _p(stats))
    assert rc == 0, rc
    return stats


def puct(legal, outcomes, probs, c_puct=2.0):
    """outcomes: dict card -> list of outcomes. Returns (choice index, pucts float64[n])."""
    idx, out = [], []
    for i, a in enumerate(legal):
        for o in outcomes[a]:
            idx.append(i)
            out.append(int(o))
    idx, out = np.array(idx, np.int32), np.array(out, np.int32)
    probs = np.ascontiguousarray(probs, np.float32)
    pucts = np.zeros(len(legal), np.float64)
    choice = lib().sim_puct(len(legal), len(out), _p(idx), _p(out), _p(probs), ctypes.c_float(c_puct), _p(pucts))
    return choice, pucts


def fisher_yates_sources(targets):
    """targets[i] = swap partner of step i (>= i).  Returns the original position each step's output comes from."""
    t = np.zeros(90, np.uint8)
    t[: len(targets)] = targets
    out = np.zeros(len(targets), np.int32)
    lib().sim_fisher_yates_sources(_p(t), len(targets), _p(out))
    return out


def handrec(cards, played, score):
    """handrec.cuh on one hand: cards ascending (<= 10), `played` = bit mask of slots already empty."""
    cards = np.ascontiguousarray(cards, np.uint8)
    out = dict(find=np.zeros(256, np.int32), take_ok=np.zeros(256, np.uint8), take_meta=np.zeros(256, np.uint32), slot_card=np.zeros(10, np.uint8),
               count=np.zeros(1, np.int32), mask=np.zeros(4, np.uint32), select=np.zeros(10, np.uint8))
    buf = np.zeros(10, np.uint8)
    buf[: len(cards)] = cards
    lib().sim_handrec(_p(buf), len(cards), ctypes.c_uint32(played), ctypes.c_uint32(score), _p(out["find"]), _p(out["take_ok"]), _p(out["take_meta"]),
                      _p(out["slot_card"]), _p(out["count"]), _p(out["mask"]), _p(out["select"]))
    return out
