"""CPU-only: the C-ABI library loads and exports exactly what include/nimmt_b200.h declares.
No compute call is made (there is no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest

import rl_6_nimmt_b200  # noqa: F401
from rl_6_nimmt_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "nimmt_b200.h")).read()
    return sorted(set(re.findall(r"NIMMT_API[^;(]*?\b(nimmt_\w+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    declared = _declared()
    assert len(declared) >= 14
    lib = N.lib()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(N.SIGNATURES) == declared, "ctypes binding out of sync with the header"
    out = subprocess.run(["nm", "-D", "--defined-only", N.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and "nimmt_" in l.split()[-1])
    assert exported == declared, "library exports symbols the header does not declare (or vice versa)"


def test_host_only_entry_points():
    lib = N.lib()
    assert lib.nimmt_abi_version() == N.ABI_VERSION
    assert lib.nimmt_state_bytes(1 << 20, 4) == (12 * 4 + 24) << 20       # (12 P + 24) bytes per game, whole tiles of 32
    assert lib.nimmt_state_bytes(33, 2) == 2 * 32 * (12 * 2 + 24)
    assert lib.nimmt_state_bytes(10, 10) == 32 * (12 * 10 + 24)
    assert lib.nimmt_state_bytes(1, 0) == 0 and lib.nimmt_state_bytes(1, 11) == 0 and lib.nimmt_state_bytes(-1, 4) == 0
    assert lib.nimmt_obs_len(1) == 47 and lib.nimmt_obs_len(0) == 35
    vals = [lib.nimmt_card_value(c) for c in range(104)]
    assert sum(vals) == 171 and vals[54] == 7 and vals[10] == 5 and vals[9] == 3 and vals[4] == 2 and vals[0] == 1
    assert lib.nimmt_card_value(104) == -1 and lib.nimmt_card_value(-1) == -1


def test_argument_validation_needs_no_gpu():
    """Bad arguments are rejected before any CUDA call (status codes, never exceptions/aborts)."""
    lib = N.lib()
    assert lib.nimmt_step(None, None, None, None, None, 4, 4, None) == N.E_BADARG
    assert lib.nimmt_deal(ctypes.c_void_p(16), 4, 11, 0, 0, None) == N.E_BADARG
    assert lib.nimmt_deal(ctypes.c_void_p(8), 4, 4, 0, 0, None) == N.E_ALIGN
    assert lib.nimmt_deal(ctypes.c_void_p(16), -1, 4, 0, 0, None) == N.E_BADARG
    assert lib.nimmt_deal(ctypes.c_void_p(16), 0, 4, 0, 0, None) == N.OK  # empty batch is a no-op
    assert lib.nimmt_observe(ctypes.c_void_p(16), ctypes.c_void_p(16), None, 4, 4, 1, 9, None) == N.E_BADARG
    assert lib.nimmt_mcs_rollouts(ctypes.c_void_p(16), 1, 4, 10, 0, 2, 2, ctypes.c_void_p(16), None) == N.E_BADARG
    assert lib.nimmt_mcs_rollouts(ctypes.c_void_p(16), 0, 4, 10, 0, 0, 1, ctypes.c_void_p(16), None) == N.OK


def test_root_struct_layout():
    assert ctypes.sizeof(N.Root) == 64
    from rl_6_nimmt_b200 import rollouts as R
    img = R.pack_root([[1], [2, 3], [4], [5]], [10, 40, 100], list(range(50, 70)), 3)
    root = N.Root.from_buffer_copy(img.tobytes())
    assert root.own[0] == 1 << 10 and root.own[1] == 1 << 8 and root.own[3] == 1 << 4
    assert list(root.rows[1]) == [2, 3, 255, 255, 255, 255] and root.num_players == 3


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv
    with pytest.raises(N.NimmtNativeError):
        BatchedSechsNimmtEnv(4, 4)
    from rl_6_nimmt_b200 import rollouts as R
    with pytest.raises(N.NimmtNativeError):
        R.mcs_rollouts(R.pack_root([[1], [2], [3], [4]], [10, 20], list(range(30, 60)), 2)[None], 2, 10)
