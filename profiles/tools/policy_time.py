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

# the state-only 47-100-100-104 nets of the model-free agents on the same tile (k_masked_probs), against PyTorch on the device
from torch import nn


class _MaskedNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.latent_net = nn.Sequential(nn.Linear(47, 100), nn.ReLU(), nn.Linear(100, 100), nn.ReLU())
        self.head_nets = nn.ModuleList([nn.Sequential(nn.Linear(100, 104))])

    def forward(self, x):
        h = self.latent_net(x)
        return [head(h) for head in self.head_nets]


mnet = _MaskedNet().cuda()
mblob = PL.pack_masked_weights(mnet)
ms = timed(lambda: PL.masked_probs(obs, mblob), 5)
print(f"k_masked_probs: {ms:.3f} ms per {obs.shape[0]} decisions = {obs.shape[0] / ms * 1e3:.3e} decisions/s, "
      f"{obs.shape[0] * 2 * (47 * 100 + 100 * 100 + 100 * 104) / ms * 1e3 / 1e12:.0f} TFLOP/s (un-padded)")
with torch.no_grad():
    ms = timed(lambda: PL.masked_card_probs(mnet, obs), 3)
print(f"the same net in PyTorch (fp32 cuBLAS) on the device: {ms:.3f} ms")
