"""Position statistics (stats.py) against the unmodified reference's Tournament static methods (CPU)."""
import json
import os

import numpy as np
import torch

from conftest import GOLDEN

import rl_6_nimmt_b200  # noqa: F401
from rl_6_nimmt_b200 import stats as S


def test_positions_and_winner_match_the_reference():
    cases = json.load(open(os.path.join(GOLDEN, "position_stats.json")))
    by_p = {}
    for c in cases:
        by_p.setdefault(len(c["scores"]), []).append(c)
    assert sorted(by_p) == list(range(2, 11))
    for P, cs in by_p.items():
        scores = torch.tensor([c["scores"] for c in cs], dtype=torch.int32)
        np.testing.assert_array_equal(S.absolute_positions(scores).numpy(), np.array([c["absolute"] for c in cs], np.float32))
        np.testing.assert_allclose(S.relative_positions(scores).numpy(), np.array([c["relative"] for c in cs], np.float32), rtol=0, atol=1e-6)
        assert S.winners(scores).tolist() == [c["winner"] for c in cs]
        out = S.summary(scores)
        assert abs(float(out["win_rate"].sum()) - 1.0) < 1e-12 and out["mean_score"].shape == (P,)
