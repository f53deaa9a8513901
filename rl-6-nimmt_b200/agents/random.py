"""DrunkHamster (rl_6_nimmt/agents/random.py:5-13): uniformly random legal card, host RNG as in the
reference.  The batched, on-device equivalent is BatchedSechsNimmtEnv.random_actions()."""
import numpy as np

from .base import Agent


class DrunkHamster(Agent):
    def forward(self, state, legal_actions, **kwargs):
        return np.random.choice(np.array(legal_actions, dtype=np.int32), size=1)[0], {}

    def learn(self, *args, **kwargs):
        return 0.0
