"""Loader for the UNMODIFIED reference (coolo/rl-6-nimmt) — fixture generation only.

Used by ``make_golden.py`` in the build container, where ``/root/reference`` exists.
Nothing in ``tests/``, ``bench.py`` or ``__graft_entry__`` imports this at run time:
the GPU box has no ``/root/reference``; it sees only the committed fixtures.

Why a loader is needed: ``import rl_6_nimmt`` fails at HEAD because
``gym`` is not installed, ``rl_6_nimmt/__init__.py:4`` imports a module that does not
exist, and ``agents/__init__.py`` pulls in matplotlib / multi_elo.  We register a
duck-typed ``gym`` and empty parent packages whose ``__path__`` points at the
reference tree, then import the four modules on the hot path unmodified.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("NIMMT_REFERENCE", "/root/reference")


def _install_gym_stub():
    if "gym" in sys.modules:
        return
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Env:  # base class only; the reference never calls into it
        pass

    class Discrete:
        def __init__(self, n):
            self.n = n

    class Box:
        def __init__(self, low, high, shape, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

    gym.Env = Env
    spaces.Discrete = Discrete
    spaces.Box = Box
    gym.spaces = spaces
    sys.modules["gym"] = gym
    sys.modules["gym.spaces"] = spaces


def load_reference():
    """Returns a namespace with env, mcts, random, play modules of the reference."""
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_gym_stub()
    pkg_root = os.path.join(REF_ROOT, "rl_6_nimmt")
    for name, sub in (("rl_6_nimmt", ""), ("rl_6_nimmt.agents", "agents"), ("rl_6_nimmt.utils", "utils")):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            mod.__path__ = [os.path.join(pkg_root, sub) if sub else pkg_root]
            sys.modules[name] = mod
    ns = types.SimpleNamespace()
    ns.env = importlib.import_module("rl_6_nimmt.env")
    ns.mcts = importlib.import_module("rl_6_nimmt.agents.mcts")
    ns.random = importlib.import_module("rl_6_nimmt.agents.random")
    ns.play = importlib.import_module("rl_6_nimmt.play")
    ns.preprocessing = importlib.import_module("rl_6_nimmt.utils.preprocessing")
    ns.nets = importlib.import_module("rl_6_nimmt.utils.nets")
    return ns


def load_tournament():
    """rl_6_nimmt.tournament with an empty stand-in for the absent third-party `multi_elo` (tournament.py:5): only the
    position statistics (static methods, tournament.py:240-256) are used; Elo is out of scope and unpinned."""
    load_reference()
    if "multi_elo" not in sys.modules:
        sys.modules["multi_elo"] = types.ModuleType("multi_elo")
    return importlib.import_module("rl_6_nimmt.tournament")


def load_policy_agents():
    """rl_6_nimmt.agents.policy (MaskedReinforceAgent, BatchedReinforceAgent) with an empty stand-in for matplotlib, which
    utils/various.py:4 imports for plotting only."""
    load_reference()
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.lines"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].__path__ = []
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib.lines"].Line2D = object
    return importlib.import_module("rl_6_nimmt.agents.policy")
