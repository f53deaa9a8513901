"""Agent protocol of the reference (rl_6_nimmt/agents/base.py:7-62), kept so that the accelerated
agents can sit wherever the reference's do (GameSession, Tournament)."""
import torch
from torch import nn

from ..env import SechsNimmtEnv


class History:
    """Per-episode key -> list store, the subset of rl_6_nimmt/utils/replay_buffer.py:206-272 the
    Monte-Carlo agents use (store / rollout / clear)."""

    def __init__(self, max_length=None, dtype=torch.float, device=torch.device("cpu")):
        self.max_length, self.dtype, self.device = max_length, dtype, device
        self.memories = None

    def store(self, **kwargs):
        if self.memories is None:
            self.memories = {key: [] for key in kwargs}
        for key, val in kwargs.items():
            self.memories[key].append(val)
            if self.max_length is not None and len(self.memories[key]) > self.max_length:
                self.memories[key].pop(0)

    def rollout(self):
        return self.memories

    def clear(self):
        self.memories = None

    def __len__(self):
        return 0 if not self.memories else len(next(iter(self.memories.values())))


class Agent(nn.Module):
    """Abstract agent: ``agent(state, legal_actions) -> (action, info)`` and ``agent.learn(...)``."""

    def __init__(self, env=None, gamma=0.99, optim_kwargs=None, history_length=None, dtype=torch.float, device=torch.device("cpu")):
        if env is None:
            env = SechsNimmtEnv(num_players=4)  # spaces only; no device work happens here
        self.gamma, self.device, self.dtype = gamma, device, dtype
        self.action_space = env.action_space
        self.state_length = env.observation_space.shape[0]
        self.num_actions = self.action_space.n
        self.history = History(max_length=history_length, dtype=dtype, device=device)
        self.optimizer = None
        self.optim_kwargs = optim_kwargs
        super().__init__()

    def train(self, mode=True):
        super().train(mode=mode)
        if mode and any(True for _ in self.parameters()):
            self.optimizer = torch.optim.Adam(params=self.parameters(), **(self.optim_kwargs or {}))
        return self

    def forward(self, state, legal_actions, *args, **kwargs):
        raise NotImplementedError

    def learn(self, state, reward, action, done, next_state, next_reward, episode_end, num_episode, legal_actions, *args, **kwargs):
        raise NotImplementedError
