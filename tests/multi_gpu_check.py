"""Run under torchrun on N GPUs: checks that one MCS decision batch sharded over the ranks
(striped rollouts + NCCL all-reduce of the int64 [D,10,3] table) equals the unsharded table bit
for bit, and that split deals equal the unsplit deal.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rl_6_nimmt_b200  # noqa: E402,F401
from rl_6_nimmt_b200 import rollouts as R  # noqa: E402
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "mcs_exact.json")))
    roots = np.stack([R.pack_root(g[k]["board"], g[k]["own"], g[k]["available"], g[k]["P"]) for k in ("C", "E")])
    sharded = R.sharded_mcs_rollouts(roots, 2, 100_003, seed=9)          # stripe + all-reduce
    whole = R.mcs_rollouts(roots, 2, 100_003, seed=9)                    # every rank: the unsharded table
    assert torch.equal(sharded, whole), "sharded table differs from the unsharded one"
    assert sharded[0, :2, 2].tolist() == [100_003] * 2
    # a small batch is played redundantly by every rank instead (no collective): same table as the unsharded call
    small = R.sharded_mcs_rollouts(roots[:1], 2, 1000, seed=9)
    assert torch.equal(small, R.mcs_rollouts(roots[:1], 2, 1000, seed=9)) and int(small[0, 0, 2]) == 1000
    # weak-scaling deal partition: rank r deals games [r*n, (r+1)*n) of the same seed
    n = 4096
    mine = BatchedSechsNimmtEnv(n, 4, seed=3, game0=rank * n).reset().observe(dtype=torch.int8)
    full = BatchedSechsNimmtEnv(n * world, 4, seed=3).reset().observe(dtype=torch.int8)
    assert torch.equal(mine, full[rank * n:(rank + 1) * n])
    # data-parallel Alpha0.5 self-play: every rank plays its own games, all ranks train ONE net (gradient all-reduce over NCCL)
    from rl_6_nimmt_b200 import policy as PL
    from rl_6_nimmt_b200.play import BatchedGameSession, PolicySeat
    torch.manual_seed(1000 + rank)            # different initial nets: the session must broadcast rank 0's
    net = PL.PolicyNet()
    sess = BatchedGameSession([PolicySeat(net, mc_max=20, puct=True, learn=True) for _ in range(2)], 64, seed=50 + rank, data_parallel=True)
    for _ in range(2):
        sess.play_games()
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    everyone = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(everyone, flat)
    assert all(torch.equal(everyone[0], e) for e in everyone), "data-parallel replicas diverged"
    dist.barrier()
    if rank == 0:
        print(f"multi_gpu_check ok on {world} GPUs: sharded MCS table bit-identical; split deals identical; data-parallel self-play keeps one net")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
