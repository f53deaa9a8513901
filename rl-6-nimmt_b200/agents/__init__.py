"""Accelerated agents (the ones on the hot path): DrunkHamster, MCSAgent."""
from .base import Agent  # noqa: F401
from .mcts import BaseMCAgent, MCSAgent  # noqa: F401
from .random import DrunkHamster  # noqa: F401

AGENTS = {"random": DrunkHamster, "mcs": MCSAgent}
