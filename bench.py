#!/usr/bin/env python
"""bench.py — headline benchmark of the 6 nimmt! hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): env steps/s — one env step = one simultaneous turn of one game
(P card placements).  Workload at N = 1: BASELINE.json configs[1], 2^20 concurrent 4-player
games with uniformly random legal actions.  One bench "step" = one pass of env.step (k_step_tiles) over
one batch of 2^20 games whose uniformly random legal actions are already resident in HBM (recorded,
untimed, by playing the same deals once with k_random_actions), plus the re-deal of the batch
(k_deal) when its 10-turn games are over.  Four independent batches (4 x 86 MB of state + I/O,
> the 126 MB L2) are visited round-robin so no step finds its state in L2; their games are staggered so
that every window of steps holds the steady-state share of re-deals.  The K timed steps are CUDA-graph
replays, repeated 7 times (median).  The action generator, the fused random-play kernel, step + observe, the
B = 1 drop-in, the MCS and Alpha0.5 kernels (each with its roofline) and the ten-player sweep of
BASELINE configs[4] are timed separately and reported in the same line.

Printed on rank 0, one JSON line: value = whole-job env steps/s with everything resident in HBM;
e2e = the same through BatchedSechsNimmtEnv.step_host with the actions coming from pinned host
memory and rewards/done going back every step; roofline = k_step's algorithmic bytes (193 B per
4-player step) over its CUDA-event time against the measured HBM copy peak;
cpu_baseline = the C oracle port on the host cores, for context.

--impl reference times the CPU implementation of the same workload (the oracle port: the Python
reference cannot travel to the GPU box) on all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NUM_PLAYERS = 4
GAMES = 1 << 20
NSETS = 4
METRIC = "env_steps_per_sec"
UNIT = "env steps/s"
WORKLOAD = "batched env.step: 2^20 concurrent 4-player games, random legal actions (BASELINE configs[1])"


def bytes_per_step(P):
    """Algorithmic HBM bytes of one env step, SURVEY.md §8(d): read S + P actions, write S + P rewards + done,
    with the canonical state S = 17 P + 24."""
    return 36 * P + 49


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "sm_max_mhz": 1965.0}, "fallback"


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU implementation of the same workload (oracle port, all host threads)
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import oracle
    oracle.build()
    cores = host_cores()
    # one bench step = up to 2^20 env steps = 104,858 ten-turn games (incl. deal and the observation
    # rebuild the reference performs inside every step, env.py:73), bounded so that the whole
    # --steps/--warmup run takes about two minutes of wall clock
    t0 = time.perf_counter()
    oracle.bench_env(NUM_PLAYERS, 2000, cores, seed=7)
    games_per_sec = 2000 * cores / (time.perf_counter() - t0)
    budget_s = 120.0 / max(args.steps + args.warmup, 1)
    games_per_step = int(min((GAMES + 9) // 10, max(cores * 100, games_per_sec * budget_s)))
    per_thread = (games_per_step + cores - 1) // cores
    for _ in range(args.warmup):
        oracle.bench_env(NUM_PLAYERS, per_thread, cores, seed=1)
    t0 = time.perf_counter()
    done = 0
    for i in range(args.steps):
        done += oracle.bench_env(NUM_PLAYERS, per_thread, cores, seed=100 + i)
    dt = time.perf_counter() - t0
    value = done / dt
    sample = f"{args.steps} x {per_thread * cores} random-vs-random 4-player games (deal + 10 steps + 4 observations/step) on {cores} threads, {cpu_model()}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "players": NUM_PLAYERS, "games_per_step": per_thread * cores,
                   "note": "CPU port of the reference algorithm (oracle/nimmt_oracle.c); the Python reference itself cannot travel to the GPU box"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append((time.perf_counter(), parts))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        inside = [p for t, p in self.samples if t_begin <= t <= t_end] or [p for _, p in self.samples[-3:]]
        sm = [float(p[0]) for p in inside if p[0].replace(".", "").isdigit()]
        mx = [float(p[1]) for p in inside if p[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for p in inside for n, v in zip(names, p[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(inside)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def moved_bytes_per_step(P):
    """HBM bytes the shipped layout really moves per env step: read the tile record (12 P + 24) and P action
    bytes, write the mutable block (4 P + 24), P reward bytes, the done and the illegal flag."""
    return (12 * P + 24) + P + (4 * P + 24) + P + 2


def pack_roots(R, np, obs, P):
    return np.stack([R.pack_root([[int(c) for c in row if c >= 0] for row in o[0, -24:].reshape(4, 6)], [int(c) for c in o[0, :10]],
                                 [c for c in range(104) if c not in set(o[0, :10].tolist()) | set(o[0, -24:].tolist())], P) for o in obs])


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import rl_6_nimmt_b200  # noqa: F401
    from rl_6_nimmt_b200 import rollouts as R
    from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv, SechsNimmtEnv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        if world == 1:
            return vals if len(vals) > 1 else vals[0]
        t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out = [float(x) for x in t]
        return out if len(out) > 1 else out[0]

    def event_ms(fn):
        """CUDA-event time of fn() on the current stream, bracketed by barrier + synchronize, max over ranks."""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        barrier()
        return max_over_ranks(a.elapsed_time(b))

    peaks, peak_kind = measured_peaks()
    P, B, K, W = NUM_PLAYERS, GAMES, args.steps, args.warmup
    # games are independent: rank r owns global games [r * NSETS * B, (r + 1) * NSETS * B) — no data-path collective
    envs = [BatchedSechsNimmtEnv(B, P, seed=1234, game0=(rank * NSETS + s) * B) for s in range(NSETS)]
    # Synthetic input, resident in HBM before the timed region: for every batch the ten uniformly random
    # legal action tensors of its (deterministic) deal, recorded by playing the deal once.
    tapes = [torch.empty((10, B, P), dtype=torch.uint8, device=dev) for _ in range(NSETS)]
    for env, tape in zip(envs, tapes):
        env.reset(seed=env.seed)
        for t in range(10):
            env.random_actions(out=tape[t])
            env.step(tape[t])
        env.turn = 10
    launches = 0

    def one_step(i):
        nonlocal launches
        env, tape = envs[i % NSETS], tapes[i % NSETS]
        if env.turn == 10:
            env.reset(seed=env.seed)   # same seed => same deal => the recorded actions stay legal
            launches += 1
        env.step(tape[env.turn])
        launches += 1

    CYCLE = NSETS * 10   # steps after which every batch has played one full game and sits at turn 10 again

    def capture(n_steps, first=0):
        """A CUDA graph of steps first .. first + n_steps - 1 of the cycle (the loop is launch-bound from Python: a step is
        ~25 us of GPU work); returns (graph, launches in it)."""
        nonlocal launches
        g = torch.cuda.CUDAGraph()
        before = launches
        with torch.cuda.graph(g):
            for i in range(first, first + n_steps):
                one_step(i)
        return g, launches - before

    for i in range(max(W, 3)):          # warm-up steps, eager
        one_step(i)
    for i in range((-max(W, 3)) % CYCLE):   # untimed: bring every batch back to a game boundary
        one_step(max(W, 3) + i)
    # Stagger the batches' games (untimed): batch s starts STAGGER[s] turns into its game, so the four re-deals of a cycle fall on
    # steps 0, 13, 22 and 35 instead of 0, 1, 2, 3 and ANY window of 20 steps holds the steady-state share of re-deals (one per ten
    # steps of a batch: two) — aligned games would put four re-deals into the first 20 steps of a cycle and none into the next 20.
    STAGGER = (0, 7, 5, 2)
    for s_, env in enumerate(envs):
        if STAGGER[s_ % len(STAGGER)]:
            env.reset(seed=env.seed)
            for t in range(STAGGER[s_ % len(STAGGER)]):
                env.step(tapes[s_][t])
    barrier()
    # EXACTLY K steps are timed, always as graph replays: K // 40 replays of the whole 40-step cycle (4 re-deals) plus one
    # graph holding the K % 40 remaining steps; an untimed third graph finishes the cycle so that every repetition starts
    # at a game boundary.  The timed region is repeated REPS times and the median is reported.
    g_cycle, n_cycle = capture(CYCLE)
    rem = K % CYCLE
    g_rem, n_rem = capture(rem) if rem else (None, 0)
    g_fin, _ = capture(CYCLE - rem, first=rem) if rem else (None, 0)
    g_cycle.replay()
    barrier()
    REPS = 7

    def timed_region():
        for _ in range(K // CYCLE):
            g_cycle.replay()
        if g_rem is not None:
            g_rem.replay()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    t_begin = time.perf_counter()
    region_ms = []
    for _ in range(REPS):
        region_ms.append(event_ms(timed_region))
        if g_fin is not None:
            g_fin.replay()
    # keep the GPU under the same load until the clock sampler has a few samples inside the measured window
    while time.perf_counter() - t_begin < 0.8:
        g_cycle.replay()
        torch.cuda.synchronize()
    t_end = time.perf_counter()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    assert sum(int(e.illegal.any()) for e in envs) == 0, "random legal actions were rejected"
    elapsed_ms = statistics.median(region_ms)
    timed_launches = (K // CYCLE) * n_cycle + n_rem
    value = world * B * K / (elapsed_ms * 1e-3)

    # ---- the dominant kernel on its own: the same cycle split into two graphs, the 4 re-deals and the
    # 40 k_step launches, with CUDA events around the latter.  (Events around every single launch add
    # ~5 us each to a ~25 us kernel, and eager launches from Python cannot keep the queue full.) ----
    for env in envs:
        env.turn = 10
    g_deal, g_steps = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_deal):
        for env in envs:
            env.reset(seed=env.seed)
    with torch.cuda.graph(g_steps):
        for i in range(CYCLE):
            envs[i % NSETS].step(tapes[i % NSETS][i // NSETS])
    g_deal.replay(); g_steps.replay()
    kstep_runs, kdeal_runs = [], []
    for _ in range(7):
        kdeal_runs.append(event_ms(g_deal.replay) / NSETS)
        kstep_runs.append(event_ms(g_steps.replay) / CYCLE)
    assert sum(int(e.illegal.any()) for e in envs) == 0
    kstep_ms, kdeal_ms = statistics.median(kstep_runs), statistics.median(kdeal_runs)

    # ---- also: the same 40-step cycle with the re-deals issued on a side stream, so that the deal of one batch (instruction-bound)
    # runs under the other batches' steps (memory-bound).  Every batch's own order deal -> ten steps is kept. ----
    side = torch.cuda.Stream(device=dev)
    g_side = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_side):
        cap = torch.cuda.current_stream(dev)
        side.wait_stream(cap)
        dealt = []
        with torch.cuda.stream(side):
            for env in envs:
                env.reset(seed=env.seed)
                ev = torch.cuda.Event()
                ev.record(side)
                dealt.append(ev)
        for i in range(CYCLE):
            if i < NSETS:
                cap.wait_event(dealt[i])
            envs[i % NSETS].step(tapes[i % NSETS][i // NSETS])
        cap.wait_stream(side)
    g_side.replay()
    side_ms = statistics.median(event_ms(g_side.replay) for _ in range(7)) / CYCLE
    assert sum(int(e.illegal.any()) for e in envs) == 0

    # ---- also: every batch's whole game (deal -> ten steps) on its OWN stream, the four chains forked off and joined back inside one
    # graph: consecutive launches of `value`'s single stream belong to different batches and do not depend on each other, so here
    # one kernel's ramp and drain run under its neighbours and the deals run under the steps ----
    chains = [torch.cuda.Stream(device=dev) for _ in range(NSETS)]
    g_four = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_four):
        cap = torch.cuda.current_stream(dev)
        for b, st in enumerate(chains):
            st.wait_stream(cap)
            with torch.cuda.stream(st):
                envs[b].reset(seed=envs[b].seed)
                for t in range(CYCLE // NSETS):
                    envs[b].step(tapes[b][t])
        for st in chains:
            cap.wait_stream(st)
    g_four.replay()
    four_ms = statistics.median(event_ms(g_four.replay) for _ in range(7)) / CYCLE
    assert sum(int(e.illegal.any()) for e in envs) == 0

    # ---- step + observe ("report step-only and step+observe separately"; the reference rebuilds all P observations
    # inside every step, env.py:73): the same cycle with k_observe after every step, int8 and fp32 observations ----
    step_obs = {}
    for name, dt in (("i8", torch.int8), ("f32", torch.float32)):
        obs_buf = [torch.empty((B, P, 47), dtype=dt, device=dev) for _ in range(NSETS)]
        for env in envs:
            env.turn = 10
        g_so = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_so):
            for i in range(CYCLE):
                env = envs[i % NSETS]
                if env.turn == 10:
                    env.reset(seed=env.seed)
                env.step(tapes[i % NSETS][env.turn])
                env.observe(out=obs_buf[i % NSETS])
        g_so.replay()
        ms = statistics.median(event_ms(g_so.replay) for _ in range(5)) / CYCLE
        bytes_alg = bytes_per_step(P) + 47 * P * obs_buf[0].element_size() + (12 * P + 24)
        step_obs[name] = {"env_steps_per_sec": world * B / (ms * 1e-3), "ms_per_step": ms, "hbm_frac": bytes_alg * B / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                          "bytes_per_step": bytes_alg}
        del obs_buf, g_so
    for env in envs:
        env.turn = 10

    # ---- end to end: host buffers in, host buffers out, every step ---------------------------------
    # Legal action sequences are recorded once (untimed) into pinned host memory by playing each
    # batch's deal with the device RNG; the timed loop re-deals the same games and feeds them back.
    streams = [torch.cuda.Stream(device=dev) for _ in range(NSETS)]
    h_actions = [torch.empty((10, B, P), dtype=torch.uint8).pin_memory() for _ in range(NSETS)]
    h_out = [env.host_out_buffer() for env in envs]   # (pinned buffer, rewards view, done-bits view): one D2H copy per step
    h_done = [o[2] for o in h_out]
    for s, env in enumerate(envs):
        env.reset(seed=99 + s)
        for t in range(10):
            a = env.random_actions()
            h_actions[s][t].copy_(a)
            env.step(a)
    torch.cuda.synchronize()
    K2 = max(CYCLE, min(K - K % CYCLE, 2000))

    def e2e_cycle():
        """One full game of every batch through the host-buffer API, each batch on its own stream:
        deal, then ten times (actions H2D -> k_step -> rewards + done D2H)."""
        for s, env in enumerate(envs):
            with torch.cuda.stream(streams[s]):
                env.reset(seed=99 + s)
                for t in range(10):
                    env.step_host(h_actions[s][t], h_out[s][0])

    e2e_cycle()  # warm-up, eager
    barrier()
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2):   # fork the four chains off the capture stream, join them back
        cap = torch.cuda.current_stream(dev)
        for st in streams:
            st.wait_stream(cap)
        e2e_cycle()
        for st in streams:
            cap.wait_stream(st)
    g2.replay()
    # PCIe on a shared host is noisy (other tenants' traffic): K2 steps are timed five times and the MEDIAN is reported,
    # with all runs in the JSON line
    e2e_runs = [event_ms(lambda: [g2.replay() for _ in range(K2 // CYCLE)]) for _ in range(5)]
    assert all(int(e.illegal.any()) == 0 for e in envs) and all(bool((d == -1).all()) for d in h_done), "e2e replay diverged"
    e2e_ms = statistics.median(e2e_runs)
    e2e_value = world * B * K2 / (e2e_ms * 1e-3)

    # ---- the same end-to-end loop in the packed transfer format (nimmt_step_packed): 4-bit hand slots in, one bit record per game
    # out — 2 + 3 bytes per 4-player game across PCIe instead of 4 + 4.1 ----
    ab, rb = envs[0].packed_sizes()
    h_slots = [torch.empty((10, B, ab), dtype=torch.uint8).pin_memory() for _ in range(NSETS)]
    h_packed = [torch.empty((B, rb), dtype=torch.uint8).pin_memory() for _ in range(NSETS)]
    for s_, env in enumerate(envs):
        env.reset(seed=99 + s_)
        dealt = env.observe(dtype=torch.int8)[:, :, :10].clone()
        for t in range(10):
            h_slots[s_][t].copy_(BatchedSechsNimmtEnv.pack_slots(h_actions[s_][t].to(dev), dealt))
        del dealt
    torch.cuda.synchronize()

    def e2e_packed_cycle():
        for s_, env in enumerate(envs):
            with torch.cuda.stream(streams[s_]):
                env.reset(seed=99 + s_)
                for t in range(10):
                    env.step_host_packed(h_slots[s_][t], h_packed[s_])

    e2e_packed_cycle()
    barrier()
    g3 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g3):
        cap = torch.cuda.current_stream(dev)
        for st in streams:
            st.wait_stream(cap)
        e2e_packed_cycle()
        for st in streams:
            cap.wait_stream(st)
    g3.replay()
    e2e_packed_runs = [event_ms(lambda: [g3.replay() for _ in range(K2 // CYCLE)]) for _ in range(5)]
    for s_, env in enumerate(envs):
        _, dn_, il_ = env.unpack_results(h_packed[s_])
        assert bool(dn_.all()) and not bool(il_.any()), "packed e2e replay diverged"
    e2e_packed_value = world * B * K2 / (statistics.median(e2e_packed_runs) * 1e-3)

    # ---- also: the action generator alone, the fused random-play kernel, the B = 1 drop-in ----
    def graph_ms(fn, launches_in_graph, reps=5):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        g.replay()
        return statistics.median(event_ms(g.replay) for _ in range(reps)) / launches_in_graph

    for env in envs:
        env.reset(seed=env.seed)
    ra_ms = graph_ms(lambda: [envs[i % NSETS].random_actions(out=tapes[i % NSETS][0], turn=0) for i in range(3 * NSETS)], 3 * NSETS)

    def fused_cycle():
        for env in envs:
            env.reset(seed=env.seed)
        for t in range(10):
            for env in envs:
                env.step_random()
    fused_ms = graph_ms(fused_cycle, CYCLE)     # per env step, re-deals included

    facade_us = None
    if rank == 0:
        # the B = 1 drop-in (SechsNimmtEnv: reference signatures, numpy in and out), random legal play, wall clock incl. every sync
        np.random.seed(0)
        env1 = SechsNimmtEnv(P, verbose=False, device=dev)
        n_steps, t0 = 0, None
        for game in range(6):
            states, legal = env1.reset()
            if game == 1:
                torch.cuda.synchronize(); t0 = time.perf_counter(); n_steps = 0
            done = False
            while not done:
                (states, legal), _, done, _ = env1.step([l[np.random.randint(len(l))] for l in legal])
                n_steps += 1
        facade_us = 1e6 * (time.perf_counter() - t0) / n_steps

    # ---- secondary metric: MCS rollouts/s (BASELINE configs[2] shape: 4 players, 10 candidate cards) ---
    obs0 = BatchedSechsNimmtEnv(256, P, seed=5, game0=rank * 256).reset().observe(dtype=torch.int8).cpu().numpy()
    roots = pack_roots(R, np, obs0, P)
    roots_d = torch.as_tensor(roots).to(dev)
    per_action = 2000
    stats = torch.zeros((256, 10, 3), dtype=torch.int64, device=dev)
    for _ in range(3):
        R.mcs_rollouts(roots_d, P, per_action, seed=1, out=stats)
    reps = 5
    mcs_ms = statistics.median(event_ms(lambda: [R.mcs_rollouts(roots_d, P, per_action, seed=2 + r, out=stats) for r in range(reps)]) for _ in range(3))
    mcs_value = world * reps * 256 * 10 * per_action / (mcs_ms * 1e-3)
    mcs_per_gpu = mcs_value / world
    int_peak = 148 * 128 * peaks.get("sm_max_mhz", 1965.0) * 1e6          # thread-instructions/s at the maximum SM clock
    mcs_roofline = {"bound": "int-issue", "achieved": mcs_per_gpu * 40 * 64, "peak": int_peak, "unit": "thread-inst/s",
                    "frac": mcs_per_gpu * 40 * 64 / int_peak, "traffic": None, "kernel": "k_mcs_rollouts<4>",
                    "how": "SURVEY.md 8d: rollouts/s x n P placements (40) x the 64-instruction algorithmic budget per placement, over 148 SMs x 128 lanes x the maximum SM clock; per GPU"}

    # BASELINE configs[2] itself: ONE decision batch (the same roots on every rank), 10,000 rollouts per candidate card,
    # rollout ids striped over the ranks, then the path's only collective (int64 [D,10,3] all-reduce over NCCL / NVLink)
    shard_obs = BatchedSechsNimmtEnv(4096, P, seed=6, game0=0).reset().observe(dtype=torch.int8).cpu().numpy()
    shard_roots = torch.as_tensor(pack_roots(R, np, shard_obs, P)).to(dev)
    sharded = {}
    table = None
    for D, reps_d in ((1, 20), (4096, 2)):
        R.sharded_mcs_rollouts(shard_roots[:D], P, 10_000, seed=1, device=dev)

        def run_d():
            nonlocal table
            for r in range(reps_d):
                table = R.sharded_mcs_rollouts(shard_roots[:D], P, 10_000, seed=2 + r, device=dev)
        ms_d = statistics.median(event_ms(run_d) for _ in range(3)) / reps_d
        assert int(table[:, :, 2].sum()) == D * 10 * 10_000       # every rollout of every candidate was played exactly once
        sharded[f"D{D}"] = {"ms_per_decision_batch": ms_d, "rollouts_per_sec": D * 10 * 10_000 / (ms_d * 1e-3)}

    # ---- Alpha0.5 (BASELINE configs[3]): 256 PUCT searches per GPU, 200 rollouts each, policy net on tcgen05 ----
    from rl_6_nimmt_b200 import policy as PL
    torch.manual_seed(0)
    blob = PL.pack_weights(PL.PolicyNet(), device=dev)
    R.policy_rollouts(roots_d, P, blob, 200, seed=1)
    puct_ms = statistics.median(event_ms(lambda: R.policy_rollouts(roots_d, P, blob, 200, seed=2 + r)) for r in range(3))
    # the same kernel with every slot of the chip taken: three trees per CTA, three CTAs per SM (256 trees are 86 CTAs on 148 SMs)
    sat_trees = 3 * 3 * torch.cuda.get_device_properties(dev).multi_processor_count
    R.policy_rollouts(shard_roots[:sat_trees], P, blob, 200, seed=1)
    sat_ms = statistics.median(event_ms(lambda: R.policy_rollouts(shard_roots[:sat_trees], P, blob, 200, seed=2 + r)) for r in range(3))
    # batched leaf evaluation: the policy for every seat of 2^18 games (2^20 decisions, ~8.9e6 rows of 48 features)
    obs_all = envs[0].reset(seed=77).observe(dtype=torch.int8).reshape(-1, 47)[: 1 << 20].contiguous()
    PL.policy_probs(obs_all, blob)
    def leaf_run():
        for _ in range(5):   # back to back: one launch alone carries ~5 % of launch latency
            PL.policy_probs(obs_all, blob)
    leaf_ms = statistics.median(event_ms(leaf_run) for _ in range(3)) / 5
    # the whole self-play loop of configs[3]: 256 games, four PUCT seats sharing one net (mc_max = 200, run.py), every
    # search on chip, one batched imitation step per iteration
    from rl_6_nimmt_b200.play import BatchedGameSession, PolicySeat
    torch.manual_seed(0)
    sp_net = PL.PolicyNet()
    session = BatchedGameSession([PolicySeat(sp_net, mc_max=200, puct=True, learn=True) for _ in range(P)], 256, device=dev, seed=7)
    session.play_games()
    selfplay_ms = statistics.median(event_ms(session.play_games) for _ in range(2))
    rows_per_rollout = P * sum(range(1, 11))           # 220 policy rows in a full 4-player rollout
    flop_per_row = 2 * (48 * 100 + 100 * 100 + 100)    # un-padded, SURVEY.md §8d
    search_tflops = 256 * 200 * rows_per_rollout * flop_per_row / (puct_ms * 1e-3) / 1e12          # per GPU
    leaf_tflops = float((obs_all[:, :10] >= 0).sum()) * flop_per_row / (leaf_ms * 1e-3) / 1e12      # per GPU
    alpha = {
        "puct_rollouts_per_sec": world * 256 * 200 / (puct_ms * 1e-3), "ms_per_256_decisions": puct_ms,
        "config": "256 PUCT searches per GPU (4-player opening roots, 10 legal cards), 200 sequential rollouts each, random-init policy net (torch.manual_seed(0))",
        "policy_tflops_in_search": world * search_tflops,
        "selfplay_games_per_sec": world * 256 / (selfplay_ms * 1e-3), "selfplay_ms_per_256_games": selfplay_ms,
        "selfplay_config": "256 four-player games per GPU, every seat a PUCT agent (mc_max 200) on one shared net, 36 search launches + one batched imitation step (Adam) per iteration",
        "puct_saturated": {"trees_per_gpu": sat_trees, "ms": sat_ms, "rollouts_per_sec": world * sat_trees * 200 / (sat_ms * 1e-3),
                           "policy_tflops_per_gpu": sat_trees * 200 * rows_per_rollout * flop_per_row / (sat_ms * 1e-3) / 1e12,
                           "note": "k_policy_rollouts with all 3 x 148 CTA slots of the chip filled (1,332 searches per GPU): the throughput the kernel has for batches beyond configs[3]'s 256 trees, whose 86 CTAs measure the latency of a search"},
        "leaf_eval_decisions_per_sec": world * (1 << 20) / (leaf_ms * 1e-3),
        "leaf_eval_tflops": world * leaf_tflops,
        "roofline": {"bound": "tensor", "achieved": search_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": search_tflops / peaks["bf16_tflops"],
                     "traffic": None, "kernel": "k_policy_rollouts<4>",
                     "how": "un-padded FLOPs (29,800 per policy row x 220 rows per 4-player rollout) over the CUDA-event time of one launch, per GPU; a search is latency-bound (sequential rollouts), see DESIGN.md 5.4",
                     "leaf_eval": {"kernel": "k_policy_probs", "achieved": leaf_tflops, "frac": leaf_tflops / peaks["bf16_tflops"]}},
    }

    # ---- BASELINE configs[4]: ten-player max-table sweep, the games split evenly over the ranks (weak per-rank timing, max over ranks) ----
    sweep = []
    free_bytes = torch.cuda.mem_get_info(dev)[0]
    del envs, tapes, h_actions, h_out, h_slots, h_packed, streams, g_cycle, g_rem, g_fin, g_deal, g_steps, g_side, g2, g3, session
    torch.cuda.empty_cache()
    for lg in args.sweep_log2:
        total_games = 1 << lg
        Bs = total_games // world
        need = Bs * (12 * 10 + 24 + 10 + 10 + 2) * 1.05
        if need > 0.8 * torch.cuda.mem_get_info(dev)[0]:
            sweep.append({"games": total_games, "skipped": "does not fit one GPU's share of HBM"})
            continue
        nsets = max(1, min(4, -(-3 * 126_000_000 // (Bs * 144))))       # rotate batches until the L2 cannot hold them
        es = [BatchedSechsNimmtEnv(Bs, 10, seed=21 + s, game0=(rank * nsets + s) * Bs) for s in range(nsets)]
        acts = [torch.empty((Bs, 10), dtype=torch.uint8, device=dev) for _ in range(nsets)]
        step_ms = deal_ms = 0.0
        for rep in range(2):    # the first pass warms up
            deal_ms = event_ms(lambda: [e.reset(seed=e.seed) for e in es]) / nsets
            step_ms = 0.0
            for t in range(10):
                for e, a in zip(es, acts):
                    e.random_actions(out=a)
                step_ms += event_ms(lambda: [e.step(a) for e, a in zip(es, acts)]) / nsets
        assert all(bool(e.done.all()) and not bool(e.illegal.any()) for e in es)
        step_ms /= 10
        sweep.append({"games": total_games, "games_per_gpu": Bs, "batches_per_gpu": nsets, "k_step_ms": step_ms, "k_deal_ms": deal_ms,
                      "env_steps_per_sec": world * Bs / (step_ms * 1e-3),
                      "frac_canonical": bytes_per_step(10) * Bs / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                      "frac_moved": moved_bytes_per_step(10) * Bs / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]})
        del es, acts
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    achieved = bytes_per_step(P) * B / (kstep_ms * 1e-3) / 1e9
    moved = moved_bytes_per_step(P) * B / (kstep_ms * 1e-3) / 1e9
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic, traffic_note = tj.get("k_step_p4_dram_bytes_per_launch"), tj.get("note_short")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "players": P, "games_per_gpu_per_step": B,
                   "l2": f"{NSETS} independent batches visited round-robin ({NSETS} x {(12 * P + 24 + 2 * P + 2) * B / 1e6:.0f} MB > 126 MB L2), no explicit flush",
                   "step": "k_step over recorded uniformly random legal actions resident in HBM, + k_deal every 10th visit of a batch; the K timed steps are CUDA-graph replays (whole 40-step cycles + one graph of the K % 40 remainder); the four batches' games are staggered (0, 7, 5, 2 turns in) so that every 20-step window holds the steady-state share of re-deals (two)", "seed": 1234},
        "timed_region_ms": elapsed_ms, "timed_region_runs_ms": region_ms, "timed_region_reported": f"median of {REPS} repetitions of exactly {K} steps",
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": traffic, "traffic_note": traffic_note, "kernel": "k_step_tiles<4>", "kernel_ms": kstep_ms,
                     "frac_canonical": achieved / peaks["hbm_gbs"], "frac_moved": moved / peaks["hbm_gbs"], "achieved_moved": moved,
                     "moved_bytes_per_launch": moved_bytes_per_step(P) * B,
                     "how": "CUDA events around a graph of 40 back-to-back k_step launches (4 batches x 10 turns), median of 7; `frac` = SURVEY 8d's canonical 193 B/step, `frac_moved` = the 122 B/step the shipped layout really moves",
                     "algorithmic_bytes_per_launch": bytes_per_step(P) * B, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * P, "d2h_bytes_per_step": (B * P + 15) // 16 * 16 + 4 * ((B + 31) // 32), "steps": K2,
                "runs": [world * B * K2 / (m * 1e-3) for m in e2e_runs], "reported": "median of 5 runs of `steps` steps",
                "api": "BatchedSechsNimmtEnv.step_host (pinned host actions in; rewards int8 [B,P] + done as one bit per game out in one copy), 4 batches on 4 streams, 40-step cycles replayed as a CUDA graph",
                "packed": {"value": e2e_packed_value, "unit": UNIT, "h2d_bytes_per_step": B * ab, "d2h_bytes_per_step": B * rb,
                           "runs": [world * B * K2 / (m * 1e-3) for m in e2e_packed_runs],
                           "api": "BatchedSechsNimmtEnv.step_host_packed: 4-bit hand slots in, one bit record per game (5 bits of bull heads per player, done, illegal) out; same loop, same graph structure"}},
        "gpu_launches": timed_launches,
        "clocks": clocks,
        "also": {"k_random_actions_ms": ra_ms, "k_deal_ms": kdeal_ms,
                 "env_steps_per_sec_redeal_on_side_stream": world * B / (side_ms * 1e-3),
                 "env_steps_per_sec_batches_on_four_streams": world * B / (four_ms * 1e-3),
                 "batches_on_four_streams_note": "the same 40 steps + 4 re-deals with every batch's game (deal -> ten steps) on its own stream, forked and joined inside one graph: up to four 2^20-game kernels are in flight at once, so this is NOT configs[1]'s one batch at a time — it shows what the ramp and drain of a 23 us launch and the serial deals cost `value`",
                 "redeal_on_side_stream_note": "the timed 40-step cycle with the 4 re-deals issued on a second stream (each batch still deal -> ten steps in order): the instruction-bound deal runs under the other batches' memory-bound steps; `value` keeps everything on one stream",
                 "fused_random_play_env_steps_per_sec": world * B / (fused_ms * 1e-3),
                 "fused_note": "k_step_tiles<4,true>: actions drawn in-kernel (DrunkHamster for every seat), + k_deal every 10th visit; max over ranks",
                 "step_plus_observe_i8": step_obs["i8"], "step_plus_observe_f32": step_obs["f32"],
                 "step_plus_observe_note": "k_step + k_observe of all P seats after every step (+ k_deal every 10th), the reference's own step semantics (env.py:73); hbm_frac = algorithmic bytes (193 + 47 P sizeof + 12 P + 24) over the cycle time",
                 "b1_facade_us_per_step": facade_us,
                 "b1_facade_note": "SechsNimmtEnv (the B = 1 drop-in with the reference's signatures) playing random legal cards, wall clock per env.step incl. observation rebuild and host syncs; the Python reference takes ~130 us per 4-player step (BASELINE.md)"},
        "alpha05": alpha,
        "mcs": {"metric": "mcs_rollouts_per_sec", "value": mcs_value, "sharded_decision_10k_per_card": sharded, "unit": "rollouts/s",
                "config": "256 four-player opening roots x 10 candidate cards x 2000 rollouts per launch, 5 launches", "roofline": mcs_roofline},
        "sweep_p10": sweep,
    }
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        oracle.build()
        cores = host_cores()
        probe_games = 2000
        t0 = time.perf_counter()
        oracle.bench_env(P, probe_games, cores, seed=3)
        rate = probe_games * cores * 10 / (time.perf_counter() - t0)
        per_thread = max(1000, int(rate * 12 / 10 / cores))  # ~12 s of CPU work
        t0 = time.perf_counter()
        n = oracle.bench_env(P, per_thread, cores, seed=4)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{per_thread * cores} random-vs-random 4-player games ({n} env steps incl. deal and per-step observations) in {dt:.1f} s on {cores} threads, {cpu_model()}; C port of the reference algorithm (oracle/nimmt_oracle.c)"}
        # the MCS agent's rollout loop (agents/mcts.py:91-154) on the same host cores, same root shape as the GPU number
        o = obs0[0]
        board = [[int(c) for c in row if c >= 0] for row in o[0, -24:].reshape(4, 6)]
        own = [int(c) for c in o[0, :10]]
        avail = [c for c in range(104) if c not in set(own) | set(sum(board, []))]
        t0 = time.perf_counter()
        oracle.bench_mcs(P, board, own, avail, 2000, cores, seed=5)
        rate = 2000 * cores / (time.perf_counter() - t0)
        per_thread = max(2000, int(rate * 8 / cores))   # ~8 s of CPU work
        t0 = time.perf_counter()
        st = oracle.bench_mcs(P, board, own, avail, per_thread, cores, seed=6)
        dt = time.perf_counter() - t0
        line["mcs"]["cpu_baseline"] = {"value": int(st[:, 2].sum()) / dt, "unit": "rollouts/s", "cores": cores, "kind": "port",
                                       "sample": f"{int(st[:, 2].sum())} reference-law rollouts of a 4-player opening root (determinise + 10 turns, agents/mcts.py:108-154) in {dt:.1f} s on {cores} threads; C port (oracle/nimmt_oracle.c)"}
        # Alpha0.5's rollout loop (PolicyMCSAgent: every move sampled from softmax(policy net), agents/mcts.py:129-154, 209-228) on the
        # same host cores: the fp32 C port, one call per thread (ctypes releases the GIL), same net and root as the GPU number
        from concurrent.futures import ThreadPoolExecutor
        torch.manual_seed(0)
        sd = {k: v.detach().cpu().numpy() for k, v in PL.PolicyNet().state_dict().items()}
        wts = {"w1": sd["latent_net.0.weight"], "b1": sd["latent_net.0.bias"], "w2": sd["latent_net.2.weight"], "b2": sd["latent_net.2.bias"],
               "w3": sd["head_nets.0.0.weight"], "b3": sd["head_nets.0.0.bias"]}
        t0 = time.perf_counter()
        oracle.policy_rollouts(P, board, own, avail, 100, wts, seed=1)
        per_thread = max(100, int(100 / (time.perf_counter() - t0) * 6))   # ~6 s of CPU work per thread
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as pool:
            done_rollouts = sum(int(st[:, 2].sum()) for st in pool.map(lambda i: oracle.policy_rollouts(P, board, own, avail, per_thread, wts, seed=10 + i), range(cores)))
        dt = time.perf_counter() - t0
        line["alpha05"]["cpu_baseline"] = {"value": done_rollouts / dt, "unit": "rollouts/s", "cores": cores, "kind": "port",
                                           "sample": f"{done_rollouts} policy-driven rollouts of a 4-player opening root (220 fp32 evaluations of the 48-100-100-1 net each; the root move sampled from the policy, PUCT differs in that move only) in {dt:.1f} s on {cores} threads; C port (oracle/nimmt_oracle.c)"}
    else:
        line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep-log2", type=lambda v: [int(x) for x in v.split(",") if x], default=[20, 24, 28],
                    help="total games of the ten-player sweep (BASELINE configs[4]), log2, split evenly over the GPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        # keep the CPU arm within minutes: one step is ~0.25 s of 8-thread work
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
