"""Accelerated agents (the ones on the hot path): DrunkHamster, MCSAgent, PolicyMCSAgent, PUCTAgent."""
from .base import Agent  # noqa: F401
from .mcts import BaseMCAgent, MCSAgent, PolicyMCSAgent, PUCTAgent  # noqa: F401
from .random import DrunkHamster  # noqa: F401

AGENTS = {"random": DrunkHamster, "mcs": MCSAgent, "pmcs": PolicyMCSAgent, "alpha0.5": PUCTAgent}
