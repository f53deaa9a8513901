"""Host side of the Alpha0.5 policy net (PolicyMCSAgent.actor, agents/mcts.py:194-228).

The network is the reference's: ``MultiHeadedMLP(48, (100, 100), (1,))`` with ReLU
(utils/nets.py:100-132) behind ``SechsNimmtStateNormalization(action=True)``
(utils/preprocessing.py:12-57).  It lives in PyTorch as an ``nn.Module`` with the reference's
parameter names (so its state_dict is interchangeable and autograd trains it); for inference the
parameters are packed once into a device blob and evaluated by the tcgen05 kernel.
"""
import numpy as np
import torch
from torch import nn

from . import _native as N


class PolicyNet(nn.Module):
    """Same parameter tree as the reference's ``actor`` (``latent_net.{0,2}``, ``head_nets.0.0``)."""

    def __init__(self, input_size=48, hidden_sizes=(100, 100)):
        super().__init__()
        if (input_size, tuple(hidden_sizes)) != (48, (100, 100)):
            raise NotImplementedError("the tcgen05 kernel is built for the reference's 48-100-100-1 policy net")
        self.latent_net = nn.Sequential(nn.Linear(48, 100), nn.ReLU(), nn.Linear(100, 100), nn.ReLU())
        self.head_nets = nn.ModuleList([nn.Sequential(nn.Linear(100, 1))])

    def forward(self, inputs):          # fp32 torch path: training only (agents/mcts.py:230-261)
        latent = self.latent_net(inputs)
        return [head(latent) for head in self.head_nets]


# (begin, end, min, max) of the segments of [action, obs47] (utils/preprocessing.py:21-47)
_SEGMENTS = ((0, 1, 0.0, 103.0), (1, 11, 0.0, 103.0), (11, 12, 0.0, 6.0), (12, 16, 1.0, 5.0), (16, 20, 0.0, 103.0),
             (20, 24, 1.0, 10.0), (24, 48, 0.0, 103.0))


def normalize_rows(rows):
    """SechsNimmtStateNormalization(action=True).forward (utils/preprocessing.py:12-57) in torch (training path)."""
    scale = torch.empty(48, dtype=rows.dtype, device=rows.device)
    shift = torch.empty(48, dtype=rows.dtype, device=rows.device)
    for a, b, lo, hi in _SEGMENTS:
        scale[a:b] = 2.0 / (hi - lo)
        shift[a:b] = -1.0 - 2.0 * lo / (hi - lo)
    return rows * scale + shift


def normalize_states(obs):
    """SechsNimmtStateNormalization(action=False).forward (utils/preprocessing.py:12-57): the 47-vector without a
    candidate card in front, as the state-only nets consume it (MaskedReinforceAgent, agents/policy.py:38)."""
    scale = torch.empty(47, dtype=obs.dtype, device=obs.device)
    shift = torch.empty(47, dtype=obs.dtype, device=obs.device)
    for a, b, lo, hi in _SEGMENTS[1:]:
        scale[a - 1:b - 1] = 2.0 / (hi - lo)
        shift[a - 1:b - 1] = -1.0 - 2.0 * lo / (hi - lo)
    return obs * scale + shift


def masked_card_probs(net, obs):
    """MaskedReinforceAgent.forward up to the sampling (agents/policy.py:45-50) for a batch: obs [D,47] (any dtype) ->
    probabilities float32 [D,10] over the hand slots (0 for empty slots).  ``net`` maps normalised states [D,47] to a
    list whose first entry holds one logit per card [D,104] (``MultiHeadedMLP(47, hidden, (104,))``); plain library GEMMs."""
    obs = obs.to(torch.float32)
    (logits,) = net(normalize_states(obs))[:1]
    hand = obs[:, :10].to(torch.int64)
    legal = hand >= 0
    picked = logits.gather(1, hand.clamp(min=0)).masked_fill(~legal, float("-inf"))
    return torch.softmax(picked, dim=1)


def pack_masked_weights(net, device=None):
    """Packs a ``MultiHeadedMLP(47, (100, 100), (104,))`` (MaskedReinforceAgent's actor; a DQN's Q-net has the same shape) into the
    device blob of nimmt_masked_probs.  Returns None if ``net`` does not have that parameter tree (the caller then evaluates it
    with PyTorch on the device)."""
    sd = {k: v.detach().to("cpu", torch.float32).contiguous().numpy() for k, v in net.state_dict().items()}
    keys = ("latent_net.0.weight", "latent_net.0.bias", "latent_net.2.weight", "latent_net.2.bias", "head_nets.0.0.weight", "head_nets.0.0.bias")
    if any(k not in sd for k in keys):
        return None
    w1, b1, w2, b2, w3, b3 = (sd[k] for k in keys)
    if w1.shape != (100, 47) or w2.shape != (100, 100) or w3.shape != (104, 100):
        return None
    lib = N.lib()
    blob = np.zeros(lib.nimmt_masked_weights_bytes(), np.uint8)
    ptr = lambda a: np.ascontiguousarray(a).ctypes.data
    N.check(lib.nimmt_masked_pack_weights(ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(w3), ptr(b3), blob.ctypes.data), "nimmt_masked_pack_weights")
    t = torch.from_numpy(blob)
    if device is not None or torch.cuda.is_available():
        t = t.to(device if device is not None else "cuda")
    return t


def masked_probs(obs, weights, want_logits=False):
    """MaskedReinforceAgent.forward up to the sampling on the tensor cores (nimmt_masked_probs): obs int8 [D,47] (device) ->
    probabilities float32 [D,10] over the hand slots (0 for empty slots)."""
    if not torch.cuda.is_available():
        raise N.NimmtNativeError("masked_probs needs a CUDA device; there is no CPU fallback")
    lib = N.lib()
    assert obs.dtype == torch.int8 and obs.is_cuda and obs.dim() == 2 and obs.shape[1] == 47
    obs = obs.contiguous()
    if obs.data_ptr() % 4:
        obs = obs.clone()
    D = obs.shape[0]
    probs = torch.empty((D, 10), dtype=torch.float32, device=obs.device)
    logits = torch.empty((D, 10), dtype=torch.float32, device=obs.device) if want_logits else None
    with torch.cuda.device(obs.device):
        N.check(lib.nimmt_masked_probs(N.ptr(obs), D, N.ptr(weights), N.ptr(probs), N.ptr(logits),
                                       torch.cuda.current_stream(obs.device).cuda_stream), "nimmt_masked_probs")
    return (probs, logits) if want_logits else probs


def torch_policy(net, state, legal_actions):
    """PolicyMCSAgent._compute_policy (agents/mcts.py:219-228) with autograd: probabilities over the legal cards."""
    state = torch.as_tensor(state, dtype=torch.float32).reshape(-1)
    cards = torch.tensor([float(a) for a in legal_actions], dtype=torch.float32).unsqueeze(1)
    rows = torch.cat((cards, state.unsqueeze(0).expand(len(legal_actions), -1)), dim=1)
    (logits,) = net(normalize_rows(rows))
    return torch.softmax(logits, dim=0).flatten()


def pack_weights(net, device=None):
    """Packs a PolicyNet (or anything with the same state_dict keys) into the device blob."""
    lib = N.lib()
    sd = {k: v.detach().to("cpu", torch.float32).contiguous().numpy() for k, v in net.state_dict().items()}
    w1, b1 = sd["latent_net.0.weight"], sd["latent_net.0.bias"]
    w2, b2 = sd["latent_net.2.weight"], sd["latent_net.2.bias"]
    w3, b3 = sd["head_nets.0.0.weight"].reshape(-1), float(sd["head_nets.0.0.bias"].reshape(-1)[0])
    assert w1.shape == (100, 48) and w2.shape == (100, 100) and w3.shape == (100,)
    blob = np.zeros(lib.nimmt_policy_weights_bytes(), np.uint8)
    ptr = lambda a: np.ascontiguousarray(a).ctypes.data
    N.check(lib.nimmt_policy_pack_weights(ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(w3), b3, blob.ctypes.data), "nimmt_policy_pack_weights")
    t = torch.from_numpy(blob)
    if device is not None or torch.cuda.is_available():
        t = t.to(device if device is not None else "cuda")
    return t


def policy_probs(obs, weights, want_logits=False):
    """PolicyMCSAgent._compute_policy for a batch: obs int8 [D,47] (device) -> probs float32 [D,10]
    by hand slot (0 for empty slots)."""
    if not torch.cuda.is_available():
        raise N.NimmtNativeError("policy_probs needs a CUDA device; there is no CPU fallback")
    lib = N.lib()
    assert obs.dtype == torch.int8 and obs.is_cuda and obs.dim() == 2 and obs.shape[1] == 47
    obs = obs.contiguous()
    if obs.data_ptr() % 4:   # a row slice of a larger tensor: the kernel reads with 32-bit loads from a 4-byte aligned base
        obs = obs.clone()
    D = obs.shape[0]
    probs = torch.empty((D, 10), dtype=torch.float32, device=obs.device)
    logits = torch.empty((D, 10), dtype=torch.float32, device=obs.device) if want_logits else None
    with torch.cuda.device(obs.device):
        N.check(lib.nimmt_policy_probs(N.ptr(obs), D, N.ptr(weights), N.ptr(probs), N.ptr(logits),
                                       torch.cuda.current_stream(obs.device).cuda_stream), "nimmt_policy_probs")
    return (probs, logits) if want_logits else probs
