"""rl-6-nimmt_b200 — B200-native (sm_100a) implementation of the hot path of coolo/rl-6-nimmt.

Import as ``rl_6_nimmt_b200`` (a root-level shim aliases the hyphenated directory name).

    from rl_6_nimmt_b200.env import SechsNimmtEnv, BatchedSechsNimmtEnv, InvalidMoveException
    from rl_6_nimmt_b200.agents import MCSAgent, DrunkHamster

Scope: SURVEY.md §8 — the environment dynamics (rl_6_nimmt/env.py) and the Monte-Carlo rollouts
(rl_6_nimmt/agents/mcts.py), nothing else.  All rules run in hand-written CUDA kernels behind the
C ABI of include/nimmt_b200.h; there is no CPU fallback.
"""
from . import _native  # noqa: F401
from .env import BatchedSechsNimmtEnv, InvalidMoveException, SechsNimmtEnv  # noqa: F401

__all__ = ["SechsNimmtEnv", "BatchedSechsNimmtEnv", "InvalidMoveException"]
