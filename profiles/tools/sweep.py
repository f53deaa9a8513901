#!/usr/bin/env python
"""Per-kernel throughput sweep (BASELINE configs[4]: 10-player max-table sweep, plus P = 2, 4).

For every (P, B): CUDA-graph a full 10-turn game of one batch (k_deal + 10 x [k_random_actions, k_step]) and
time each kernel family separately with events around graphs of back-to-back launches, rotating over enough
independent batches to exceed the 126 MB L2.  Prints one JSON line per (P, B).

    python profiles/tools/sweep.py [--max-log2 26]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rl_6_nimmt_b200  # noqa: E402,F401
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv  # noqa: E402

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def graph_time(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run(P, B):
    state_bytes = (12 * P + 24) * B
    nsets = max(1, min(4, -(-3 * 126_000_000 // state_bytes)))
    envs = [BatchedSechsNimmtEnv(B, P, seed=11 + s, game0=s * B) for s in range(nsets)]
    tapes = [torch.empty((10, B, P), dtype=torch.uint8, device="cuda") for _ in range(nsets)]
    for e, tp in zip(envs, tapes):
        e.reset(seed=e.seed)
        for t in range(10):
            e.random_actions(out=tp[t])
            e.step(tp[t])
    out = {"players": P, "games": B, "batches": nsets}
    for e in envs:
        e.reset(seed=e.seed)
    torch.cuda.synchronize()

    def steps():
        for t in range(10):
            for e, tp in zip(envs, tapes):
                e.step(tp[t])
    # deals are replayed outside the timed graph so the tapes stay legal
    def deals():
        for e in envs:
            e.reset(seed=e.seed)
    ms_deal = graph_time(deals) / nsets
    deals(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    steps(); deals(); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        steps()
    tot = 0.0
    for _ in range(3):
        deals()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    ms_step = tot / 3 / (10 * nsets)
    assert all(int(e.illegal.any()) == 0 for e in envs)
    deals()
    ms_ra = graph_time(lambda: [e.random_actions(out=tp[0], turn=0) for e, tp in zip(envs, tapes)]) / nsets
    out["k_step_ms"], out["k_deal_ms"], out["k_random_actions_ms"] = ms_step, ms_deal, ms_ra
    obs8 = None
    if B * P * 47 <= 40e9:   # the int8 observations of 2^28 ten-player games alone would be 126 GB
        obs8 = torch.empty((B, P, 47), dtype=torch.int8, device="cuda")
        ms_obs8 = graph_time(lambda: [e.observe(out=obs8) for e in envs]) / nsets
        out["k_observe_i8_ms"] = ms_obs8
        out["observe_i8_GBs"] = (12 * P + 24 + 47 * P) * B / (ms_obs8 * 1e-3) / 1e9
    if B * P * 47 * 4 <= 24e9:
        obs32 = torch.empty((B, P, 47), dtype=torch.float32, device="cuda")
        out["k_observe_f32_ms"] = graph_time(lambda: [e.observe(out=obs32) for e in envs]) / nsets
        out["observe_f32_GBs"] = (12 * P + 24 + 188 * P) * B / (out["k_observe_f32_ms"] * 1e-3) / 1e9
        del obs32
    alg = 36 * P + 49
    out["env_steps_per_sec_step_only"] = B / (ms_step * 1e-3)
    out["env_steps_per_sec_with_redeal"] = B / ((ms_step + ms_deal / 10) * 1e-3)
    out["step_algorithmic_GBs"] = alg * B / (ms_step * 1e-3) / 1e9
    out["step_frac_of_measured_hbm_peak"] = out["step_algorithmic_GBs"] / PEAK
    out["deal_GBs"] = (12 * P + 24) * B / (ms_deal * 1e-3) / 1e9
    del envs, tapes, obs8
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-log2", type=int, default=26)
    ap.add_argument("--min-log2", type=int, default=20)
    ap.add_argument("--players", type=int, nargs="*", default=[10, 4, 2])
    args = ap.parse_args()
    for P in args.players:
        for lg in (20, 22, 24, 26, 28):
            if lg < args.min_log2 or lg > args.max_log2 or (12 * P + 24 + 20 * P) * (1 << lg) * min(4, max(1, 400_000_000 // ((12 * P + 24) << lg) + 1)) > 120e9:
                continue
            print(json.dumps(run(P, 1 << lg)), flush=True)
