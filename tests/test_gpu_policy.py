"""GPU tests of the tcgen05 policy-net kernel (Alpha0.5 leaf evaluation) through the C ABI."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import policy_oracle as po

import rl_6_nimmt_b200  # noqa: F401
from rl_6_nimmt_b200 import policy as PL

pytestmark = pytest.mark.gpu


def _golden():
    z = np.load(os.path.join(GOLDEN, "policy_vectors.npz"))
    net = PL.PolicyNet()
    net.load_state_dict({k[len("w_actor_"):].replace("latent_net_0_", "latent_net.0.").replace("latent_net_2_", "latent_net.2.")
                         .replace("head_nets_0_0_", "head_nets.0.0."): torch.from_numpy(z[k]) for k in z.files if k.startswith("w_actor_")})
    return z, net


def _decisions(z):
    """Golden rows are [card | obs47] grouped per decision: recover (obs47, n_legal) per decision."""
    obs, off = [], 0
    for n in z["seg"]:
        rows = z["rows_in"][off:off + n]
        assert (rows[:, 1:] == rows[0, 1:]).all() and (rows[:, 0] == rows[0, 1:1 + n]).all()
        obs.append(rows[0, 1:])
        off += n
    return np.array(obs, np.int8)


def test_probs_match_reference_and_bf16_emulation():
    z, net = _golden()
    w = po.weights_from_golden(z)
    obs = _decisions(z)                                     # [264, 47]
    blob = PL.pack_weights(net)
    probs, logits = PL.policy_probs(torch.from_numpy(obs).cuda(), blob, want_logits=True)
    probs, logits = probs.cpu().numpy(), logits.cpu().numpy()
    off = 0
    worst_ref = worst_emu = worst_logit = 0.0
    for d, n in enumerate(z["seg"]):
        got = probs[d, :n]
        assert (probs[d, n:] == 0).all() and abs(got.sum() - 1.0) < 1e-5
        worst_ref = max(worst_ref, np.abs(got - z["probs"][off:off + n]).max())
        emu_l = po.policy_logits_bf16(z["rows_in"][off:off + n], w)
        worst_logit = max(worst_logit, np.abs(logits[d, :n] - emu_l).max())
        worst_emu = max(worst_emu, np.abs(got - po.softmax(emu_l)).max())
        off += n
    # vs the kernel's own arithmetic (bf16 operands, fp32 accumulate) emulated in numpy: accumulation order only
    assert worst_logit < 2e-4 and worst_emu < 5e-5, (worst_logit, worst_emu)
    # vs the fp32 reference (torch, unmodified reference code): stated tolerance for bf16 operands
    assert worst_ref < 1e-3, worst_ref


def test_ragged_batches_and_tile_boundaries():
    z, net = _golden()
    obs = torch.from_numpy(_decisions(z)).cuda()
    blob = PL.pack_weights(net)
    full = PL.policy_probs(obs, blob)
    for D in (1, 11, 12, 13, 25, 263):
        part = PL.policy_probs(obs[:D].clone(), blob)
        assert torch.equal(part, full[:D]), D
    big = obs.repeat(40, 1)                                  # 10,560 decisions: many tiles per CTA
    out = PL.policy_probs(big, blob)
    assert torch.equal(out, full.repeat(40, 1))


def test_torch_module_is_state_dict_compatible():
    z, net = _golden()
    rows = torch.from_numpy(z["rows_norm"])
    (logit,) = net(rows)
    np.testing.assert_allclose(logit.detach().numpy().reshape(-1), z["logits"], rtol=1e-4, atol=1e-5)
