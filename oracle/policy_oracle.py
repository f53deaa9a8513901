"""numpy restatement of the Alpha0.5 leaf evaluation and root rule — TEST INFRASTRUCTURE ONLY.

Follows (reference paths):
  rl_6_nimmt/utils/preprocessing.py:12-57   SechsNimmtStateNormalization(action=True)
  rl_6_nimmt/utils/nets.py:100-132          MultiHeadedMLP(48, (100, 100), (1,)), ReLU
  rl_6_nimmt/agents/mcts.py:219-228         _compute_policy: softmax over the legal-card rows
  rl_6_nimmt/agents/mcts.py:295-315         PUCTAgent._compute_pucts / _normalize_q
Pinned against tests/golden/policy_vectors.npz and puct_cases.json (generated from the
unmodified reference).
"""
import numpy as np

# (start, stop, min, max) of every segment of the 48-vector [action, obs47] (preprocessing.py:21-47)
SEGMENTS = (
    (0, 1, 0.0, 103.0),    # candidate card
    (1, 11, 0.0, 103.0),   # own hand
    (11, 12, 0.0, 6.0),    # number of players
    (12, 16, 1.0, 5.0),    # cards per row
    (16, 20, 0.0, 103.0),  # top card per row
    (20, 24, 1.0, 10.0),   # bull heads per row
    (24, 48, 0.0, 103.0),  # board 4x6
)


def normalization_affine():
    """Returns (scale[48], shift[48]) with normalised = x * scale + shift (preprocessing.py:56-57)."""
    scale = np.zeros(48, np.float64)
    shift = np.zeros(48, np.float64)
    for a, b, lo, hi in SEGMENTS:
        scale[a:b] = 2.0 / (hi - lo)
        shift[a:b] = -1.0 - 2.0 * lo / (hi - lo)
    return scale, shift


def normalize(rows):
    rows = np.asarray(rows, np.float32)
    out = np.empty_like(rows)
    for a, b, lo, hi in SEGMENTS:
        out[:, a:b] = np.float32(-1.0) + np.float32(2.0) * (rows[:, a:b] - np.float32(lo)) / np.float32(hi - lo)
    return out


def mlp_logits(norm_rows, w):
    """w: dict with w1[100,48], b1[100], w2[100,100], b2[100], w3[1,100], b3[1] (torch Linear layout)."""
    h = np.maximum(norm_rows @ w["w1"].T + w["b1"], 0.0)
    h = np.maximum(h @ w["w2"].T + w["b2"], 0.0)
    return (h @ w["w3"].T + w["b3"]).reshape(-1)


def softmax(x):
    e = np.exp(x - np.max(x))
    return e / e.sum()


def policy_probs(rows, w):
    """rows [n,48] raw (un-normalised) -> probs [n] over the n candidate cards."""
    return softmax(mlp_logits(normalize(rows), w).astype(np.float32))


def weights_from_golden(npz):
    return {
        "w1": npz["w_actor_latent_net_0_weight"], "b1": npz["w_actor_latent_net_0_bias"],
        "w2": npz["w_actor_latent_net_2_weight"], "b2": npz["w_actor_latent_net_2_bias"],
        "w3": npz["w_actor_head_nets_0_0_weight"], "b3": npz["w_actor_head_nets_0_0_bias"],
    }


def normalize_q(outcomes):
    """mcts.py:304-315: (max, min, 'mean'=median) of all outcomes, or (0,-10,-5) below 10 samples."""
    flat = [o for lst in outcomes.values() for o in lst]
    if len(flat) < 10:
        return 0.0, -10.0, -5.0
    return float(np.max(flat)), float(np.min(flat)), float(np.median(flat))


def pucts(legal, outcomes, probs, c_puct=2.0):
    """mcts.py:295-302.  0/0 -> NaN is kept (all outcomes equal), as in the reference."""
    n = np.array([len(outcomes[a]) for a in legal])
    n_total = n.sum()  # numpy int64 scalar, as sum(ndarray) is in the reference: keeps its dtype promotion
    mx, mn, md = normalize_q(outcomes)
    q = np.array([md if not outcomes[a] else np.mean(outcomes[a]) for a in legal], dtype=np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        q = np.clip((q - mn) / (mx - mn), 0.0, 1.0)
    return q + c_puct * np.asarray(probs) * (n_total + 1.0e-9) ** 0.5 / (1.0 + n)


def puct_choice(p):
    """mcts.py:286-293: strict '>' scan from -inf; NaN never wins => index 0."""
    best, choice = -float("inf"), 0
    for i, v in enumerate(p):
        if v > best:
            best, choice = v, i
    return choice


# ----------------------------------------------------------------------------------------------
# Emulation of what the tcgen05 kernel computes, for sharp kernel tests: normalisation folded
# into layer 1, bf16 operands, fp32 accumulation, biases carried through the GEMMs as two bf16 terms
# (hi + lo) against constant-1 inputs, hidden layer 1 rounded to bf16, layer 3 in fp32.
# The fp32 functions above remain the reference semantics; this one only explains the rounding.
# ----------------------------------------------------------------------------------------------
def _bf16(x):
    """Round-to-nearest-even to bfloat16, returned as float32."""
    import torch
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(torch.bfloat16).to(torch.float32).numpy()


def fold_normalization(w):
    """(W1', b1') with W1' x_raw + b1' == W1 normalize(x_raw) + b1 (preprocessing.py:56-57 is affine)."""
    scale, shift = normalization_affine()
    w1 = w["w1"].astype(np.float64)
    return (w1 * scale[None, :]).astype(np.float32), (w["b1"].astype(np.float64) + w1 @ shift).astype(np.float32)


def _bf16_hi_lo(b):
    hi = _bf16(b)
    return hi + _bf16(np.asarray(b, np.float32) - hi)


def _bf16_three_terms(b):
    b = np.asarray(b, np.float32)
    hi = _bf16(b)
    mid = _bf16(b - hi)
    return hi.astype(np.float64) + mid + _bf16(b - hi - mid)


def policy_logits_bf16(rows, w):
    w1f, b1f = fold_normalization(w)
    x = _bf16(rows)                      # raw features are small integers: exact
    h1 = np.maximum(x @ _bf16(w1f).T + _bf16_hi_lo(b1f), 0.0).astype(np.float32)
    a2 = _bf16(h1)
    w2b, b2b = _bf16(w["w2"]), _bf16_hi_lo(w["b2"])
    h2 = (a2 @ w2b.T + b2b).astype(np.float32)
    # relu(h) = (h + |h|) / 2: the linear half of the head rides in the GEMM as one more output unit whose weights are the
    # combined row sum_n (w3_n / 2) [W2 | b2]_n, carried as three bf16 terms; the threads add the |h| half in fp32
    w3h = (0.5 * w["w3"].reshape(-1).astype(np.float64))
    row = _bf16_three_terms((w3h @ w2b.astype(np.float64)).astype(np.float32))
    bias = _bf16_three_terms(np.array([(w3h * b2b.astype(np.float64)).sum()], np.float32))[0]
    linear = (a2 @ row + bias).astype(np.float32)
    return (linear + np.abs(h2) @ w3h.astype(np.float32) + np.float32(w["b3"].reshape(-1)[0])).astype(np.float32)


def policy_probs_bf16(rows, w):
    return softmax(policy_logits_bf16(rows, w))
