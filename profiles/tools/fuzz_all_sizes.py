#!/usr/bin/env python
"""One-off differential fuzz, beyond what tests/ runs every time: for every table size P = 1..10, 2^18 + 13 games (whole tiles + a
ragged tail) dealt and played with the device RNG, single steps for odd seeds and one ten-turn launch for even ones, then replayed
through the C oracle — rewards, done, hands, boards and scores of every game at every turn, bit for bit.
    python profiles/tools/fuzz_all_sizes.py [seeds per P]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import rl_6_nimmt_b200  # noqa: E402,F401
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv  # noqa: E402

oracle.build()
seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = (1 << 18) + 13
total = 0
t_start = time.time()
for P in range(1, 11):
    for s in range(seeds):
        env = BatchedSechsNimmtEnv(n, P, seed=9000 + 100 * P + s).reset()
        obs0 = env.observe(dtype=torch.int8).cpu().numpy()
        hands0, board0 = obs0[:, :, :10].copy(), obs0[:, 0, -24:].reshape(n, 4, 6).copy()
        acts = np.zeros((n, 10, P), np.int8)
        rewards = np.zeros((n, 10, P), np.int8)
        done = np.zeros((n, 10), np.uint8)
        hands = np.zeros((n, 10, P, 10), np.int8)
        boards = np.zeros((n, 10, 4, 6), np.int8)
        scores = np.zeros((n, 10, P), np.int16)
        for t in range(10):
            a = env.random_actions().clone()
            rew, dn = env.step(a)
            acts[:, t] = a.cpu().numpy().view(np.int8)
            rewards[:, t], done[:, t] = rew.cpu().numpy(), dn.cpu().numpy()
            obs = env.observe(dtype=torch.int8).cpu().numpy()
            hands[:, t], boards[:, t] = obs[:, :, :10], obs[:, 0, -24:].reshape(n, 4, 6)
            scores[:, t] = env.scores().cpu().numpy()
            assert not bool(env.illegal.any())
        want = oracle.replay(P, board0, hands0, acts, want_obs=False)
        assert not want["illegal"].any()
        for k, got in (("rewards", rewards), ("done", done), ("hands", hands), ("boards", boards), ("scores", scores)):
            assert np.array_equal(got, want[k]), (P, s, k)
        # the whole tiles of the same games once more as ONE ten-turn launch from the same deal (a deal depends on (seed, game id)
        # only; the multi-turn launch takes batches that are a multiple of 16 games)
        n2 = n - 13
        env2 = BatchedSechsNimmtEnv(n2, P, seed=9000 + 100 * P + s).reset()
        tape = torch.as_tensor(acts[:n2].view(np.uint8).transpose(1, 0, 2).copy()).cuda()
        rew_m, done_m, ill_m = env2.step_many(tape)
        assert np.array_equal(rew_m.cpu().numpy().transpose(1, 0, 2), want["rewards"][:n2]) and not bool(ill_m.any()), (P, s, "step_many")
        assert np.array_equal(env2.scores().cpu().numpy(), want["scores"][:n2, -1]), (P, s, "step_many scores")
        total += n * 10
        print(f"P={P:2d} seed {s}: {n} games x 10 turns bit-exact (single steps and one ten-turn launch)", flush=True)
print(f"fuzz ok: {total} env steps compared with the oracle in {time.time() - t_start:.0f} s")
