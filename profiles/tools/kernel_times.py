#!/usr/bin/env python
"""CUDA-event times of the env kernels, one JSON line per (players, games):

    python profiles/tools/kernel_times.py [--players 4,10] [--log2 20] [--what step,fused,deal,actions,observe]

Every kernel is timed inside a CUDA graph of back-to-back launches that rotate over enough independent batches to exceed
the 126 MB L2 (the same method as bench.py's roofline block).  NIMMT_B200_LIB selects a tuning variant of the library."""
import argparse
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rl_6_nimmt_b200  # noqa: E402,F401
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv  # noqa: E402


def graph_ms(fn, launches, reps=5):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b) / launches)
    return statistics.median(times)


def run(P, B, what):
    state_bytes = (12 * P + 24) * B
    nsets = max(1, min(4, -(-3 * 126_000_000 // state_bytes)))
    envs = [BatchedSechsNimmtEnv(B, P, seed=11 + s, game0=s * B) for s in range(nsets)]
    tapes = [torch.empty((10, B, P), dtype=torch.uint8, device="cuda") for _ in range(nsets)]
    for e, tp in zip(envs, tapes):
        e.reset(seed=e.seed)
        for t in range(10):
            e.random_actions(out=tp[t])
            e.step(tp[t])
        assert not bool(e.illegal.any()) and bool(e.done.all())
    out = {"players": P, "games": B, "batches": nsets, "lib": os.environ.get("NIMMT_B200_LIB", "default")}

    def redeal():
        for e in envs:
            e.reset(seed=e.seed)

    if "step" in what:
        redeal()

        def steps():
            for t in range(10):
                for e, tp in zip(envs, tapes):
                    e.step(tp[t])
        # a graph replay plays the ten turns of every batch once: re-deal (untimed) before each replay
        redeal(); steps(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        redeal()
        with torch.cuda.graph(g):
            steps()
        ts = []
        for _ in range(5):
            redeal()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / (10 * nsets))
        assert all(not bool(e.illegal.any()) and bool(e.done.all()) for e in envs)
        ms = statistics.median(ts)
        moved = 12 * P + 24 + P + 4 * P + 24 + P + 2
        out.update(k_step_us=1e3 * ms, step_env_steps_per_s=B / (ms * 1e-3), step_moved_GBs=moved * B / (ms * 1e-3) / 1e9,
                   step_canonical_GBs=(36 * P + 49) * B / (ms * 1e-3) / 1e9)
    if "many" in what:
        # the same ten turns as ONE launch per batch (nimmt_step_many): the state stays in shared memory for all of them
        stacked = [tp.clone() for tp in tapes]
        rew = [torch.empty((10, B, P), dtype=torch.int8, device="cuda") for _ in range(nsets)]
        dn = [torch.empty((10, B), dtype=torch.uint8, device="cuda") for _ in range(nsets)]
        il = [torch.empty((10, B), dtype=torch.uint8, device="cuda") for _ in range(nsets)]

        def many():
            for e, tp, r, d, i in zip(envs, stacked, rew, dn, il):
                e.step_many(tp, r, d, i)
        redeal(); many(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        redeal()
        with torch.cuda.graph(g):
            many()
        ts = []
        for _ in range(5):
            redeal()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) / (10 * nsets))
        assert all(not bool(i.any()) and bool(d[9].all()) for i, d in zip(il, dn))
        ms = statistics.median(ts)
        out.update(step_many10_us_per_step=1e3 * ms, step_many10_env_steps_per_s=B / (ms * 1e-3))

        def fused_many():
            for e in envs:
                e.reset(seed=e.seed)
                e.step_random_many(10, rew[0], dn[0])
        out["fused_many10_game_us_per_batch"] = 1e3 * graph_ms(fused_many, 1) / nsets      # deal + ten fused turns, one launch
    if "fused" in what:
        def fused():
            for e in envs:
                e.reset(seed=e.seed)
            for t in range(10):
                for e in envs:
                    e.step_random()
        ms_all = graph_ms(fused, 1)
        out["fused_game_us_per_batch"] = 1e3 * ms_all / nsets       # deal + 10 fused steps of one batch
    if "deal" in what:
        ms = graph_ms(lambda: [e.reset(seed=e.seed) for e in envs for _ in range(3)], 3 * nsets)
        out.update(k_deal_us=1e3 * ms, deal_GBs=(12 * P + 24) * B / (ms * 1e-3) / 1e9)
    if "actions" in what:
        redeal()
        ms = graph_ms(lambda: [e.random_actions(out=tp[0], turn=0) for e, tp in zip(envs, tapes) for _ in range(3)], 3 * nsets)
        out.update(k_random_actions_us=1e3 * ms, actions_GBs=(12 * P + P) * B / (ms * 1e-3) / 1e9)
    if "observe" in what and B * P * 47 * nsets < 40e9:
        redeal()
        obs = [torch.empty((B, P, 47), dtype=torch.int8, device="cuda") for _ in range(nsets)]
        ms = graph_ms(lambda: [e.observe(out=o) for e, o in zip(envs, obs) for _ in range(3)], 3 * nsets)
        out.update(k_observe_i8_us=1e3 * ms, observe_i8_GBs=(47 * P + 12 * P + 24) * B / (ms * 1e-3) / 1e9)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--players", default="4")
    ap.add_argument("--log2", default="20")
    ap.add_argument("--what", default="step,fused,deal,actions,observe")
    a = ap.parse_args()
    for P in [int(x) for x in a.players.split(",")]:
        for lg in [int(x) for x in a.log2.split(",")]:
            run(P, 1 << lg, a.what.split(","))
