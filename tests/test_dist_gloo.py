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


def _dp_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import rl_6_nimmt_b200  # noqa: F401
    from rl_6_nimmt_b200 import policy as PL
    from rl_6_nimmt_b200 import train as T
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.load(os.path.join(GOLDEN, "policy_train.npz"))
    obs, chosen = torch.from_numpy(z["obs"]), torch.from_numpy(z["chosen"])
    torch.manual_seed(100 + rank)               # the replicas START different: the broadcast must make them one net
    net = PL.PolicyNet()
    T.broadcast_parameters(net)
    opt = torch.optim.Adam(net.parameters())
    mine = slice(10 * rank, 10 * rank + 10)     # rank r trains on episode r
    (-T.imitation_log_probs(net, obs[mine], chosen[mine]).sum()).backward()
    T.allreduce_gradients(net)
    grads = [p.grad.detach().numpy().copy() for p in net.parameters()]
    for _ in range(2):
        T.imitation_step(net, opt, obs[mine], chosen[mine], episodes=1, data_parallel=True)
    out.put((rank, [p.detach().numpy().copy() for p in net.parameters()], grads))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_self_play_trains_one_net_world2():
    """two ranks, each with its own episodes, train ONE net — parameters broadcast from rank 0,
    the 15,101-entry gradient all-reduced (mean) before every Adam step.  The exchanged gradient is the gradient of the mean loss
    over both ranks' episodes (checked against a single process), and after two steps the replicas are still bit-identical.
    (Weights are not compared with the single process: where a gradient is rounding noise Adam's step is noise of size lr,
    see test_train.py.)"""
    import rl_6_nimmt_b200  # noqa: F401
    from rl_6_nimmt_b200 import policy as PL
    from rl_6_nimmt_b200 import train as T
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for a, b in zip(res[0][1], res[1][1]):
        assert np.array_equal(a, b)                                   # the replicas never diverged
    z = np.load(os.path.join(GOLDEN, "policy_train.npz"))
    obs, chosen = torch.from_numpy(z["obs"]), torch.from_numpy(z["chosen"])
    torch.manual_seed(100)                                            # rank 0's initial net
    net = PL.PolicyNet()
    (-T.imitation_log_probs(net, obs[:20], chosen[:20]).sum() / 2.0).backward()   # mean over the two episodes = mean of the ranks' losses
    for g0, g1, p in zip(res[0][2], res[1][2], net.parameters()):
        assert np.array_equal(g0, g1)
        np.testing.assert_allclose(g0, p.grad.detach().numpy(), rtol=1e-4, atol=2e-6)
