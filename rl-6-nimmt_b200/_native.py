"""ctypes binding of libnimmt_b200.so (include/nimmt_b200.h) — the only way into the CUDA code.

There is no CPU fallback: if the library is missing or a call fails, this module raises.
PyTorch is used by the callers for device memory and streams only; no torch type crosses
this boundary (pointers are passed as integers).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# NIMMT_B200_LIB selects another build of the same library (kernel-tuning variants, profiles/tools/build_variants.sh)
LIB_PATH = os.environ.get("NIMMT_B200_LIB") or os.path.join(_HERE, "lib", "libnimmt_b200.so")

ABI_VERSION = 3
OK, E_BADARG, E_ALIGN, E_CUDA, E_UNSUPPORTED = 0, -1, -2, -3, -4
DT_I8, DT_I16, DT_F32, DT_I64 = 0, 1, 2, 3
ROOT_PUCT, ROOT_POLICY, ROOT_STRATIFIED = 0, 1, 2

_ERRORS = {E_BADARG: "bad argument", E_ALIGN: "misaligned pointer", E_CUDA: "CUDA error", E_UNSUPPORTED: "unsupported"}


class NimmtNativeError(RuntimeError):
    pass


class Root(ctypes.Structure):
    """struct nimmt_root (64 bytes)."""
    _fields_ = [("own", ctypes.c_uint32 * 4), ("available", ctypes.c_uint32 * 4),
                ("rows", (ctypes.c_uint8 * 6) * 4), ("num_players", ctypes.c_uint8), ("pad", ctypes.c_uint8 * 7)]


_vp, _i64, _u64, _u32, _int = ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int

# name -> (restype, argtypes); must list every NIMMT_API symbol of include/nimmt_b200.h
SIGNATURES = {
    "nimmt_abi_version": (_int, []),
    "nimmt_last_cuda_error": (ctypes.c_char_p, []),
    "nimmt_state_bytes": (ctypes.c_size_t, [_i64, _int]),
    "nimmt_obs_len": (_int, [_int]),
    "nimmt_card_value": (_int, [_int]),
    "nimmt_deal": (_int, [_vp, _i64, _int, _u64, _u64, _vp]),
    "nimmt_deal_from_perm": (_int, [_vp, _vp, _i64, _int, _vp]),
    "nimmt_reset_to": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _vp]),
    "nimmt_step": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _vp]),
    "nimmt_step1": (_int, [_vp, _i64, _vp, _vp, _int, _int, _vp, _vp]),
    "nimmt_step_many": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _int, _vp]),
    "nimmt_step_random_many": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _u64, _u32, _u64, _int, _vp]),
    "nimmt_packed_bytes": (_int, [_int, _vp, _vp]),
    "nimmt_step_packed": (_int, [_vp, _vp, _vp, _i64, _int, _vp]),
    "nimmt_step_choice": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _vp]),
    "nimmt_pack_flags": (_int, [_vp, _vp, _i64, _vp]),
    "nimmt_observe": (_int, [_vp, _vp, _vp, _i64, _int, _int, _int, _vp]),
    "nimmt_scores": (_int, [_vp, _vp, _i64, _int, _vp]),
    "nimmt_random_actions": (_int, [_vp, _vp, _i64, _int, _u64, _u32, _u64, _vp]),
    "nimmt_step_random": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _u64, _u32, _u64, _vp]),
    "nimmt_mcs_rollouts": (_int, [_vp, _int, _int, _i64, _u64, _int, _int, _vp, _vp]),
    "nimmt_mc_roots": (_int, [_vp, _vp, _vp, _i64, _int, _int, _int, _vp]),
    "nimmt_mc_choose": (_int, [_vp, _vp, _vp, _i64, _int, _int, _vp]),
    "nimmt_elo_scan": (_int, [_vp, _vp, _vp, _i64, _int, ctypes.c_double, _vp, _vp]),
    "nimmt_policy_weights_bytes": (ctypes.c_size_t, []),
    "nimmt_policy_pack_weights": (_int, [_vp, _vp, _vp, _vp, _vp, ctypes.c_float, _vp]),
    "nimmt_policy_probs": (_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "nimmt_masked_weights_bytes": (ctypes.c_size_t, []),
    "nimmt_masked_pack_weights": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "nimmt_masked_probs": (_int, [_vp, _i64, _vp, _vp, _vp, _vp]),
    "nimmt_policy_rollouts": (_int, [_vp, _int, _int, _vp, _int, ctypes.c_float, _int, _u64, _vp, _vp, _vp]),
    "nimmt_puct_choose": (_int, [_vp, _vp, _vp, _vp, _vp, _int, ctypes.c_float, _vp, _vp, _vp]),
}

_lib = None


def lib():
    """Loads the shared library (once).  Raises if it has not been built: no fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NimmtNativeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C rl-6-nimmt_b200/csrc`). There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.nimmt_abi_version() != ABI_VERSION:
            raise NimmtNativeError(f"ABI mismatch: library {L.nimmt_abi_version()}, binding {ABI_VERSION}; rebuild")
        _lib = L
    return _lib


def check(rc, what):
    if rc != OK:
        detail = lib().nimmt_last_cuda_error().decode() if rc == E_CUDA else ""
        raise NimmtNativeError(f"{what} failed: {_ERRORS.get(rc, rc)} {detail}".strip())


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
