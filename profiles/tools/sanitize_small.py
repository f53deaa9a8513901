#!/usr/bin/env python
"""Small batches of every hot kernel, for compute-sanitizer:

    compute-sanitizer --tool memcheck  python profiles/tools/sanitize_small.py
    compute-sanitizer --tool racecheck python profiles/tools/sanitize_small.py

Covers k_deal, k_random_actions, k_step_smem<4>/<10> (TMA pipeline, in-place shared-memory stepping), the ragged-tail
k_step, the fused random step, k_observe, k_mcs_rollouts<4>, k_policy_probs and k_policy_rollouts<4>; each result is
also checked against the oracle so that the run proves the sanitized launches did the real work."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import rl_6_nimmt_b200  # noqa: E402,F401
from rl_6_nimmt_b200 import policy as PL  # noqa: E402
from rl_6_nimmt_b200 import rollouts as R  # noqa: E402
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv  # noqa: E402


def env_case(n, P):
    env = BatchedSechsNimmtEnv(n, P, seed=3).reset()
    obs0 = env.observe(dtype=torch.int8).cpu().numpy()
    hands0, board0 = obs0[:, :, :10], obs0[:, 0, -24:].reshape(n, 4, 6)
    acts = np.zeros((n, 10, P), np.int8)
    rewards = np.zeros((n, 10, P), np.int8)
    for t in range(10):
        a = env.random_actions().clone()
        rew, _ = env.step(a)
        acts[:, t], rewards[:, t] = a.cpu().numpy().view(np.int8), rew.cpu().numpy()
    want = oracle.replay(P, board0, hands0, acts)
    obs = env.observe(dtype=torch.int8).cpu().numpy()
    assert (rewards == want["rewards"]).all() and (obs == want["obs"][:, -1]).all()
    env.reset()
    for t in range(10):
        env.step_random()
    assert bool(env.done.all())
    print(f"env ok: {n} games, P={P}", flush=True)


def main():
    torch.cuda.set_device(0)
    env_case(1024 + 7, 4)      # whole tiles through k_step_smem + a ragged tail through k_step
    env_case(512, 10)
    env_case(96, 2)
    P = 4
    env = BatchedSechsNimmtEnv(8, P, seed=5).reset()
    o = env.observe(dtype=torch.int8).cpu().numpy()
    roots = np.stack([R.pack_root([[int(c) for c in row if c >= 0] for row in x[0, -24:].reshape(4, 6)], [int(c) for c in x[0, :10]],
                                  [c for c in range(104) if c not in set(x[0, :10].tolist()) | set(x[0, -24:].tolist())], P) for x in o])
    stats = R.mcs_rollouts(roots, P, 512, seed=1).cpu().numpy()
    assert (stats[:, :, 2] == 512).all()
    print("mcs ok", flush=True)
    torch.manual_seed(0)
    blob = PL.pack_weights(PL.PolicyNet())
    st, probs = R.policy_rollouts(roots, P, blob, 12, seed=2)
    assert (st.cpu().numpy()[:, :, 2].sum(axis=1) == 12).all()
    pr = PL.policy_probs(torch.as_tensor(o[:, 0]).cuda().contiguous(), blob).cpu().numpy()
    assert np.allclose(pr.sum(axis=1), 1.0, atol=1e-4)
    torch.cuda.synchronize()
    print("policy ok", flush=True)


if __name__ == "__main__":
    main()
