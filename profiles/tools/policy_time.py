"""Times k_policy_rollouts (256 PUCT searches x 200 rollouts) and k_policy_probs (2^20 decisions) with CUDA events.
    python profiles/tools/policy_time.py [players]"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import rl_6_nimmt_b200
from rl_6_nimmt_b200 import policy as PL, rollouts as R
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
blob = PL.pack_weights(PL.PolicyNet())
env = BatchedSechsNimmtEnv(1 << 18, P, seed=5).reset()
obs = env.observe(dtype=torch.int8).reshape(-1, 47).contiguous()
o = obs.reshape(-1, P, 47)[:256].cpu().numpy()
roots = np.stack([R.pack_root([[int(c) for c in row if c >= 0] for row in x[0, -24:].reshape(4, 6)], [int(c) for c in x[0, :10]],
                  [c for c in range(104) if c not in set(x[0, :10].tolist()) | set(x[0, -24:].tolist())], P) for x in o])
roots_d = torch.as_tensor(roots).cuda()


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for r in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for name, mode in (("puct", 0), ("policy", 1)):
    ms = timed(lambda: R.policy_rollouts(roots_d, P, blob, 200, seed=3, root_rule=mode), 3)
    print(f"k_policy_rollouts<{P}> mode={name}: {ms:.3f} ms per 256 x 200 rollouts = {256 * 200 / ms * 1e3:.3e} rollouts/s, {ms * 1e3 / 200 / 10:.2f} us per turn")
ms = timed(lambda: PL.policy_probs(obs, blob), 5)
print(f"k_policy_probs: {ms:.3f} ms per {obs.shape[0]} decisions = {obs.shape[0] / ms * 1e3:.3e} decisions/s")
