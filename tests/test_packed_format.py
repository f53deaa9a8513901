"""CPU tests of the host side of the packed transfer format (nimmt_step_packed): slot packing and record unpacking are plain
tensor code and must agree with the bit layout the kernel writes (include/nimmt_b200.h; the kernel side is checked against the
oracle by tests/test_host_sim.py[tile-records-packed-io] and on the GPU by tests/test_gpu_parity_r2.py)."""
import numpy as np
import torch

import rl_6_nimmt_b200  # noqa: F401
from rl_6_nimmt_b200.env import BatchedSechsNimmtEnv, unpack_packed_results


def test_pack_slots_names_the_dealt_position():
    rng = np.random.RandomState(0)
    for P in (1, 2, 3, 4, 7, 10):
        B = 200
        dealt = np.stack([np.sort(rng.choice(104, 10 * P, replace=False).reshape(P, 10), axis=1) for _ in range(B)]).astype(np.int8)
        slot = rng.randint(0, 10, size=(B, P))
        cards = np.take_along_axis(dealt, slot[:, :, None], axis=2)[:, :, 0].copy()
        cards[::7, 0] = 104 % 128          # a card nobody holds (104 is not a card; as int8 it is still 104): slot 15
        packed = BatchedSechsNimmtEnv.pack_slots(torch.from_numpy(cards), torch.from_numpy(dealt)).numpy()
        assert packed.shape == (B, (P + 1) // 2) and packed.dtype == np.uint8
        for p in range(P):
            got = (packed[:, p >> 1] >> (4 * (p & 1))) & 15
            want = slot[:, p].copy()
            if p == 0:
                want[::7] = 15
            assert (got == want).all(), (P, p)
        if P % 2:
            assert ((packed[:, -1] >> 4) == 0).all()     # the unused high nibble of the last byte


def test_unpack_results_bit_layout():
    rng = np.random.RandomState(1)
    for P in (1, 2, 4, 5, 10):
        B = 300
        pen = rng.randint(0, 28, size=(B, P))
        done, ill = rng.randint(0, 2, size=B), rng.randint(0, 2, size=B)
        nb = (5 * P + 2 + 7) // 8
        packed = np.zeros((B, nb), np.uint8)
        for b in range(B):
            rec = sum(int(pen[b, p]) << (5 * p) for p in range(P)) | int(done[b]) << (5 * P) | int(ill[b]) << (5 * P + 1)
            packed[b] = [(rec >> (8 * i)) & 255 for i in range(nb)]
        rew, d, i = unpack_packed_results(torch.from_numpy(packed), P)
        assert (rew.numpy() == -pen).all() and (d.numpy() == done.astype(bool)).all() and (i.numpy() == ill.astype(bool)).all()
