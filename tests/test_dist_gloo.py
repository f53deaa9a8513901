"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: rollout striping + the path's one
collective (integer all-reduce of the [D,10,3] table) + the decision rule, and the weak-scaling
game-id partition of the batched env.  The per-rank compute is played by the host build of the
product's rollout code (tests/host_sim): the kernels themselves need a GPU."""
import json
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, root_bytes, P, rollouts, seed, legal, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import rl_6_nimmt_b200  # noqa: F401
    from rl_6_nimmt_b200 import rollouts as R
    from host_sim import mcs
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    stripe = torch.from_numpy(mcs(P, root_bytes, rollouts, seed, rank=rank, world=world))[None]   # [1,10,3]
    before = stripe.clone()
    total = R.allreduce_stats(stripe)          # the product's collective wrapper
    action, means = R.choose_from_stats(legal, total[0].numpy())
    # the product's front end (rollouts.sharded_mcs_rollouts) with the kernel launch replaced by the host build of the same
    # rollout code: a small batch is played redundantly by every rank (no collective), a large one is striped + all-reduced
    calls = []

    def host_rollouts(roots, num_players, rollouts_per_action, seed=0, rank=0, world=1, out=None, device=None):
        calls.append((rank, world))
        return torch.from_numpy(mcs(num_players, bytes(roots[0]), rollouts_per_action, seed, rank=rank, world=world))[None]
    R.mcs_rollouts = host_rollouts
    roots = np.frombuffer(root_bytes, np.uint8)[None]
    redundant = R.sharded_mcs_rollouts(roots, P, rollouts, seed)                              # 200,010 rollouts < MIN_ROLLOUTS_TO_SHARD
    striped = R.sharded_mcs_rollouts(roots, P, rollouts, seed, min_rollouts_to_shard=1)
    assert calls == [(0, 1), (rank, world)], calls
    out.put((rank, before.numpy(), total.numpy(), action, means, redundant.numpy(), striped.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_mcs_decision_world2():
    m = json.load(open(os.path.join(GOLDEN, "mcs_exact.json")))["C"]
    import rl_6_nimmt_b200  # noqa: F401
    from rl_6_nimmt_b200 import rollouts as R
    from host_sim import mcs
    root = R.pack_root(m["board"], m["own"], m["available"], m["P"]).tobytes()
    rollouts, seed, world = 20_001, 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, root, m["P"], rollouts, seed, sorted(m["own"]), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = mcs(m["P"], root, rollouts, seed)
    # stripes are disjoint and complete: counts add up, and the reduced table equals the single-rank table bit for bit
    assert res[0][1][0, :2, 2].tolist() == [10_001, 10_001] and res[1][1][0, :2, 2].tolist() == [10_000, 10_000]
    for rank, _, total, action, means, redundant, striped in res:
        assert (total[0] == whole).all()
        assert (redundant[0] == whole).all() and (striped[0] == whole).all()
        assert action == 25                                   # E[25] = -2.68 > E[43] = -5.00 (KAT-C)
        assert abs(means[0] - m["exact"]["25"]["mean"]) < 0.1 and abs(means[1] - m["exact"]["43"]["mean"]) < 0.1


def test_allreduce_is_noop_without_process_group():
    import rl_6_nimmt_b200  # noqa: F401
    from rl_6_nimmt_b200 import rollouts as R
    t = torch.arange(30, dtype=torch.int64).reshape(1, 10, 3)
    assert torch.equal(R.allreduce_stats(t.clone()), t)


def test_weak_scaling_game_partition():
    """bench.py gives rank r the global game ids [r*NSETS*B, (r+1)*NSETS*B): disjoint, contiguous, complete."""
    B, NSETS = 1 << 20, 4
    for world in (1, 2, 4, 8):
        ranges = [((r * NSETS + s) * B, (r * NSETS + s + 1) * B) for r in range(world) for s in range(NSETS)]
        ranges.sort()
        assert ranges[0][0] == 0 and ranges[-1][1] == world * NSETS * B
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
