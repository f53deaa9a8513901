"""Alpha0.5 training step (train.py) against the unmodified reference's learn/_train (CPU; torch autograd).

Fixture tests/golden/policy_train.npz: three episodes of one reference PolicyMCSAgent — decisions, the
log-probabilities it stored, the loss of every episode and its weights after every Adam step."""
import os

import numpy as np
import torch

from conftest import GOLDEN

import rl_6_nimmt_b200  # noqa: F401
from rl_6_nimmt_b200 import policy as PL
from rl_6_nimmt_b200 import train as T


def _load(z, prefix):
    net = PL.PolicyNet()
    net.load_state_dict({k[len(prefix) + len("actor_"):].replace("latent_net_0_", "latent_net.0.").replace("latent_net_2_", "latent_net.2.")
                         .replace("head_nets_0_0_", "head_nets.0.0."): torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix + "actor_")})
    return net


def test_log_probs_and_episode_updates_match_the_reference():
    z = np.load(os.path.join(GOLDEN, "policy_train.npz"))
    net = _load(z, "w0_")
    opt = torch.optim.Adam(net.parameters())          # agents/base.py:29-33 with optim_kwargs=None
    obs, chosen = torch.from_numpy(z["obs"]), torch.from_numpy(z["chosen"])
    lp = T.imitation_log_probs(net, obs[:10], chosen[:10]).detach().numpy()
    np.testing.assert_allclose(lp, z["log_prob"][:10], rtol=1e-5, atol=1e-6)
    assert lp[9] == 0.0 and z["n_legal"][9] == 1      # the single last card: constant 0 (agents/mcts.py:52-53)
    # One Adam step per episode, as the reference does.  Adam turns a gradient g into lr * g / (|g| + 1e-8): where g is
    # rounding noise around a mathematical zero (the head bias under a softmax, units that are constant over a decision's
    # rows) the step is noise of size lr in the reference as well, so weights are compared where every gradient so far was
    # solid (|g| > 1e-5), and exact zeros (dead ReLU units) must stay untouched.  A noise-sized difference in one bias can flip
    # a ReLU in the next episode and change whole gradient columns (in the reference too, across BLAS builds), so after each
    # comparison the net continues from the reference's weights: every episode's step is checked on its own.
    solid = [torch.ones_like(p, dtype=torch.bool) for p in net.parameters()]
    for ep in range(3):
        before = [p.detach().clone() for p in net.parameters()]
        loss = T.imitation_step(net, opt, obs[10 * ep:10 * ep + 10], chosen[10 * ep:10 * ep + 10], episodes=1)
        assert abs(float(loss) - z["loss"][ep]) < 1e-3, (ep, float(loss), z["loss"][ep])
        want = _load(z, f"w{ep + 1}_")
        checked = 0
        for i, ((name, a), b) in enumerate(zip(net.named_parameters(), want.parameters())):
            g = a.grad
            solid[i] &= g.abs() > 1e-5
            if ep == 0:
                assert torch.equal(a.detach()[g == 0], before[i][g == 0]), name
            np.testing.assert_allclose(a.detach()[solid[i]].numpy(), b.detach()[solid[i]].numpy(), rtol=0, atol=5e-6, err_msg=f"episode {ep} {name}")
            checked += int(solid[i].sum())
            a.data.copy_(b.data)
        assert checked > 3000, checked                # most of the 15,101 parameters are live and compared


def test_batched_step_is_the_mean_of_the_episode_losses():
    z = np.load(os.path.join(GOLDEN, "policy_train.npz"))
    obs, chosen = torch.from_numpy(z["obs"]), torch.from_numpy(z["chosen"])
    net = _load(z, "w0_")
    per_episode = [float(-T.imitation_log_probs(net, obs[10 * e:10 * e + 10], chosen[10 * e:10 * e + 10]).sum()) for e in range(3)]
    grads = []
    for e in range(3):
        net.zero_grad()
        (-T.imitation_log_probs(net, obs[10 * e:10 * e + 10], chosen[10 * e:10 * e + 10]).sum()).backward()
        grads.append([p.grad.clone() for p in net.parameters()])
    net.zero_grad()
    loss = -T.imitation_log_probs(net, obs, chosen).sum() / 3
    loss.backward()
    assert abs(float(loss) - np.mean(per_episode)) < 1e-5
    for i, p in enumerate(net.parameters()):
        torch.testing.assert_close(p.grad, sum(g[i] for g in grads) / 3, rtol=1e-4, atol=1e-6)


def test_two_headed_actor_critic_packs_like_its_policy_head():
    """ACER's net is MultiHeadedMLP(48, (100, 100), head_sizes=(1, 1)) (agents/actor_critic.py:44-46): its first head is the
    policy logit, so a ReinforceSeat can seat it — pack_weights reads head_nets.0.0 and ignores the value head."""
    from torch import nn
    torch.manual_seed(4)
    one = PL.PolicyNet()
    two = PL.PolicyNet()
    two.head_nets.append(nn.Sequential(nn.Linear(100, 1)))          # the critic head
    two.load_state_dict({**one.state_dict(), **{k: v for k, v in two.state_dict().items() if k.startswith("head_nets.1.")}})
    assert torch.equal(PL.pack_weights(one, device="cpu"), PL.pack_weights(two, device="cpu"))


def test_masked_policy_probs_match_the_reference():
    """policy.masked_card_probs against MaskedReinforceAgent.forward of the unmodified reference (normalisation without the
    action feature, 47 -> 100 -> 100 -> 104 net, softmax over the legal cards): tests/golden/masked_policy.npz."""
    from torch import nn
    z = np.load(os.path.join(GOLDEN, "masked_policy.npz"))

    class Net(nn.Module):                                 # the reference's MultiHeadedMLP(47, (100, 100), (104,)) parameter tree
        def __init__(self):
            super().__init__()
            self.latent_net = nn.Sequential(nn.Linear(47, 100), nn.ReLU(), nn.Linear(100, 100), nn.ReLU())
            self.head_nets = nn.ModuleList([nn.Sequential(nn.Linear(100, 104))])

        def forward(self, x):
            h = self.latent_net(x)
            return [head(h) for head in self.head_nets]

    net = Net()
    net.load_state_dict({k[len("w_actor_"):].replace("latent_net_0_", "latent_net.0.").replace("latent_net_2_", "latent_net.2.")
                         .replace("head_nets_0_0_", "head_nets.0.0."): torch.from_numpy(z[k]) for k in z.files if k.startswith("w_actor_")})
    obs = torch.from_numpy(z["states"])
    np.testing.assert_allclose(PL.normalize_states(obs.to(torch.float32)).numpy(), z["norm"], rtol=1e-6, atol=1e-6)
    with torch.no_grad():
        probs = PL.masked_card_probs(net, obs).numpy()
    np.testing.assert_allclose(probs, z["probs"], rtol=1e-4, atol=1e-6)
    assert ((probs > 0).sum(axis=1) == z["n_legal"]).all()
