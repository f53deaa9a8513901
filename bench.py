#!/usr/bin/env python
"""bench.py — headline benchmark of the 6 nimmt! hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Metric (BASELINE.json): env steps/s — one env step = one simultaneous turn of one game
(P card placements).  Workload at N = 1: BASELINE.json configs[1], 2^20 concurrent 4-player
games with uniformly random legal actions.  One bench "step" = one pass of env.step (k_step) over
one batch of 2^20 games whose uniformly random legal actions are already resident in HBM (recorded,
untimed, by playing the same deals once with k_random_actions), plus the re-deal of the batch
(k_deal) when its 10-turn games are over.  Four independent batches (4 x 101 MB of state + I/O,
> the 126 MB L2) are visited round-robin so no step finds its state in L2.  The action generator
and the fused random-play kernel are timed separately and reported under "also".

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
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import rl_6_nimmt_b200  # noqa: F401
    from rl_6_nimmt_b200 import rollouts as R
    from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv

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

    def one_step(i, ev=None):
        nonlocal launches
        env, tape = envs[i % NSETS], tapes[i % NSETS]
        if env.turn == 10:
            env.reset(seed=env.seed)   # same seed => same deal => the recorded actions stay legal
            launches += 1
        if ev is not None:
            ev[0].record()
        env.step(tape[env.turn])
        if ev is not None:
            ev[1].record()
        launches += 1

    CYCLE = NSETS * 10   # steps after which every batch has played one full game and sits at turn 10 again

    for i in range(max(W, 3)):
        one_step(i)
    for i in range((-max(W, 3)) % CYCLE):   # untimed: bring every batch back to a game boundary
        one_step(max(W, 3) + i)
    barrier()
    # The inner loop is launch-bound from Python (a k_step launch is ~35 us of GPU time), so one
    # cycle of 40 steps + 4 re-deals is captured once into a CUDA graph and replayed.
    graph = torch.cuda.CUDAGraph()
    launches = 0
    with torch.cuda.graph(graph):
        for i in range(CYCLE):
            one_step(i)
    launches_per_cycle = launches
    for env in envs:
        env.turn = 10
    graph.replay()   # warm the instantiated graph
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches = 0
    barrier()
    t_begin = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K // CYCLE):
        graph.replay()
    for env in envs:
        env.turn = 10
    for i in range(K % CYCLE):   # the remainder, eagerly: exactly K steps are timed
        one_step(i)
    e1.record()
    barrier()
    t_end = time.perf_counter()
    elapsed_ms = e0.elapsed_time(e1)
    timed_launches = launches + (K // CYCLE) * launches_per_cycle
    illegal = sum(int(e.illegal.any()) for e in envs)
    assert illegal == 0, "random legal actions were rejected"
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    value = world * B * K / (elapsed_ms * 1e-3)

    # ---- the dominant kernel on its own: the same cycle split into two graphs, the 4 re-deals and the
    # 40 k_step launches, with CUDA events around the latter.  (Events around every single launch add
    # ~5 us each to a ~34 us kernel, and eager launches from Python cannot keep the queue full.) ----
    for i in range((-(K % CYCLE)) % CYCLE):
        one_step((K % CYCLE) + i)
    barrier()
    g_deal, g_steps = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_deal):
        for env in envs:
            env.reset(seed=env.seed)
    with torch.cuda.graph(g_steps):
        for i in range(CYCLE):
            env = envs[i % NSETS]
            env.step(tapes[i % NSETS][i // NSETS])
    reps = 5
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    g_deal.replay(); g_steps.replay()
    barrier()
    for a, b_ in pairs:
        g_deal.replay()
        a.record()
        g_steps.replay()
        b_.record()
    barrier()
    assert sum(int(e.illegal.any()) for e in envs) == 0
    for env in envs:
        env.turn = 10
    kstep_ms = statistics.mean(a.elapsed_time(b_) for a, b_ in pairs) / CYCLE

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
    K2 = max(NSETS * 10, min(K, 2000))
    K2 -= K2 % (NSETS * 10)

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
    # PCIe on a shared host is noisy (other tenants' traffic): K2 steps are timed three times and the MEDIAN is reported,
    # with all three in the JSON line
    e2e_runs = []
    for _ in range(3):
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(K2 // CYCLE):
            g2.replay()
        f1.record()
        barrier()
        ms = f0.elapsed_time(f1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        e2e_runs.append(ms)
    assert all(int(e.illegal.any()) == 0 for e in envs) and all(bool((d == -1).all()) for d in h_done), "e2e replay diverged"
    e2e_ms = statistics.median(e2e_runs)
    e2e_value = world * B * K2 / (e2e_ms * 1e-3)

    # ---- also: the action generator alone and the fused random-play kernel (same batches, same rotation) ----
    def timed(fn, n):
        for i in range(NSETS):
            fn(i)
        barrier()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(int(4e7))   # hold the GPU while the host enqueues: measures kernels, not launch latency
        a.record()
        for i in range(n):
            fn(i)
        b_.record()
        barrier()
        return a.elapsed_time(b_) / n

    for env in envs:
        env.reset(seed=env.seed)
    ra_ms = timed(lambda i: envs[i % NSETS].random_actions(out=tapes[i % NSETS][0], turn=0), 200)

    def fused(i):
        env = envs[i % NSETS]
        if env.turn == 10:
            env.reset()
        env.step_random()
    fused_ms = timed(fused, 400)
    deal_ms = timed(lambda i: envs[i % NSETS].reset(), 100)

    # ---- secondary metric: MCS rollouts/s (BASELINE configs[2] shape: 4 players, 10 candidate cards) ---
    obs0 = BatchedSechsNimmtEnv(256, P, seed=5, game0=rank * 256).reset().observe(dtype=torch.int8).cpu().numpy()
    roots = np.stack([R.pack_root([[int(c) for c in row if c >= 0] for row in o[0, -24:].reshape(4, 6)], [int(c) for c in o[0, :10]],
                                  [c for c in range(104) if c not in set(o[0, :10].tolist()) | set(o[0, -24:].tolist())], P) for o in obs0])
    roots_d = torch.as_tensor(roots).to(dev)
    per_action = 2000
    stats = torch.zeros((256, 10, 3), dtype=torch.int64, device=dev)
    for _ in range(3):
        R.mcs_rollouts(roots_d, P, per_action, seed=1, out=stats)
    barrier()
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    m0.record()
    for r in range(reps):
        R.mcs_rollouts(roots_d, P, per_action, seed=2 + r, out=stats)
    m1.record()
    barrier()
    mcs_ms = m0.elapsed_time(m1)
    if world > 1:
        t = torch.tensor([mcs_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mcs_ms = float(t.item())
    mcs_value = world * reps * 256 * 10 * per_action / (mcs_ms * 1e-3)

    # BASELINE configs[2] itself: ONE decision batch (the same roots on every rank), 10,000 rollouts per candidate card,
    # rollout ids striped over the ranks, then the path's only collective (int64 [D,10,3] all-reduce over NCCL / NVLink)
    shard_obs = BatchedSechsNimmtEnv(4096, P, seed=6, game0=0).reset().observe(dtype=torch.int8).cpu().numpy()
    shard_roots = torch.as_tensor(np.stack([R.pack_root([[int(c) for c in row if c >= 0] for row in o[0, -24:].reshape(4, 6)], [int(c) for c in o[0, :10]],
                                  [c for c in range(104) if c not in set(o[0, :10].tolist()) | set(o[0, -24:].tolist())], P) for o in shard_obs])).to(dev)
    sharded = {}
    for D, reps_d in ((1, 20), (4096, 2)):
        R.sharded_mcs_rollouts(shard_roots[:D], P, 10_000, seed=1, device=dev)
        barrier()
        m0.record()
        for r in range(reps_d):
            table = R.sharded_mcs_rollouts(shard_roots[:D], P, 10_000, seed=2 + r, device=dev)
        m1.record()
        barrier()
        ms_d = m0.elapsed_time(m1) / reps_d
        if world > 1:
            t = torch.tensor([ms_d], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_d = float(t.item())
        assert int(table[:, :, 2].sum()) == D * 10 * 10_000       # every rollout of every candidate was played exactly once
        sharded[f"D{D}"] = {"ms_per_decision_batch": ms_d, "rollouts_per_sec": D * 10 * 10_000 / (ms_d * 1e-3)}

    # ---- Alpha0.5 (BASELINE configs[3]): 256 PUCT searches per GPU, 200 rollouts each, policy net on tcgen05 ----
    from rl_6_nimmt_b200 import policy as PL
    torch.manual_seed(0)
    blob = PL.pack_weights(PL.PolicyNet(), device=dev)
    R.policy_rollouts(roots_d, P, blob, 200, seed=1)
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for r in range(3):
        R.policy_rollouts(roots_d, P, blob, 200, seed=2 + r)
    a1.record()
    barrier()
    puct_ms = a0.elapsed_time(a1) / 3
    # batched leaf evaluation: the policy for every seat of 2^18 games (2^20 decisions, ~8.9e6 rows of 48 features)
    obs_all = envs[0].reset(seed=77).observe(dtype=torch.int8).reshape(-1, 47)[: 1 << 20].contiguous()
    PL.policy_probs(obs_all, blob)
    barrier()
    a0.record()
    for r in range(3):
        PL.policy_probs(obs_all, blob)
    a1.record()
    barrier()
    leaf_ms = a0.elapsed_time(a1) / 3
    if world > 1:
        t = torch.tensor([puct_ms, leaf_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        puct_ms, leaf_ms = float(t[0]), float(t[1])
    # the whole self-play loop of configs[3]: 256 games, four PUCT seats sharing one net (mc_max = 200, run.py), every
    # search on chip, one batched imitation step per iteration
    from rl_6_nimmt_b200.play import BatchedGameSession, PolicySeat
    torch.manual_seed(0)
    sp_net = PL.PolicyNet()
    session = BatchedGameSession([PolicySeat(sp_net, mc_max=200, puct=True, learn=True) for _ in range(P)], 256, device=dev, seed=7)
    session.play_games()
    barrier()
    a0.record()
    for r in range(2):
        session.play_games()
    a1.record()
    barrier()
    selfplay_ms = a0.elapsed_time(a1) / 2
    if world > 1:
        t = torch.tensor([selfplay_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        selfplay_ms = float(t.item())
    rows_per_rollout = P * sum(range(1, 11))           # 220 policy rows in a full 4-player rollout
    flop_per_row = 2 * (48 * 100 + 100 * 100 + 100)    # un-padded, SURVEY.md §8d
    alpha = {
        "puct_rollouts_per_sec": world * 256 * 200 / (puct_ms * 1e-3), "ms_per_256_decisions": puct_ms,
        "config": "256 PUCT searches per GPU (4-player opening roots, 10 legal cards), 200 sequential rollouts each, random-init policy net (torch.manual_seed(0))",
        "policy_tflops_in_search": world * 256 * 200 * rows_per_rollout * flop_per_row / (puct_ms * 1e-3) / 1e12,
        "selfplay_games_per_sec": world * 256 / (selfplay_ms * 1e-3), "selfplay_ms_per_256_games": selfplay_ms,
        "selfplay_config": "256 four-player games per GPU, every seat a PUCT agent (mc_max 200) on one shared net, 36 search launches + one batched imitation step (Adam) per iteration",
        "leaf_eval_decisions_per_sec": world * (1 << 20) / (leaf_ms * 1e-3),
        "leaf_eval_tflops": world * float((obs_all[:, :10] >= 0).sum()) * flop_per_row / (leaf_ms * 1e-3) / 1e12,
    }

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks, peak_kind = measured_peaks()
    achieved = bytes_per_step(P) * B / (kstep_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("k_step_p4_dram_bytes_per_launch")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "players": P, "games_per_gpu_per_step": B,
                   "l2": f"{NSETS} independent batches visited round-robin ({NSETS} x {(12 * P + 24 + 2 * P + 2) * B / 1e6:.0f} MB > 126 MB L2), no explicit flush",
                   "step": "k_step over recorded uniformly random legal actions resident in HBM, + k_deal every 10th visit of a batch; 40-step cycles replayed as a CUDA graph", "seed": 1234},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": traffic, "kernel": "k_step_smem<4>", "kernel_ms": kstep_ms,
                     "how": "CUDA events around a graph of 40 back-to-back k_step launches (4 batches x 10 turns), mean of 5",
                     "algorithmic_bytes_per_launch": bytes_per_step(P) * B, "peak_source": f"MEASURED_PEAKS.json ({peak_kind})"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * P, "d2h_bytes_per_step": int(h_out[0][0].numel()), "steps": K2,
                "runs": [world * B * K2 / (m * 1e-3) for m in e2e_runs], "reported": "median of 3 runs of `steps` steps",
                "api": "BatchedSechsNimmtEnv.step_host (pinned host actions in; rewards int8 [B,P] + done as one bit per game out in one copy), 4 batches on 4 streams, 40-step cycles replayed as a CUDA graph"},
        "gpu_launches": timed_launches,
        "clocks": clocks,
        "also": {"k_random_actions_ms": ra_ms, "k_deal_ms": deal_ms,
                 "fused_random_play_env_steps_per_sec": world * B / (fused_ms * 1e-3),
                 "fused_note": "k_step_smem<4,true>: actions drawn in-kernel (DrunkHamster for every seat), + k_deal every 10th visit; per-rank ms, not max-reduced"},
        "alpha05": alpha,
        "mcs": {"metric": "mcs_rollouts_per_sec", "value": mcs_value, "sharded_decision_10k_per_card": sharded, "unit": "rollouts/s",
                "config": "256 four-player opening roots x 10 candidate cards x 2000 rollouts per launch, 5 launches"},
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
    args = ap.parse_args()
    if args.impl == "reference":
        # keep the CPU arm within minutes: one step is ~0.25 s of 8-thread work
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
