"""GPU tier: raft::matrix::argmin -- the separate-pass row arg-min, pinned to the reference's own
known-answer test (cpp/tests/matrix/argmin.cu:71-75) and to numpy on ties / ragged / unaligned shapes."""
import numpy as np
import pytest
import torch

from raft_b200.matrix import argmin

pytestmark = pytest.mark.gpu


def test_reference_known_answer():
    # cpp/tests/matrix/argmin.cu:71-72: 3 x 4 matrix, expected {0, 3, 3}
    x = torch.tensor([0.1, 0.2, 0.3, 0.4, 0.4, 0.3, 0.2, 0.1, 0.2, 0.3, 0.5, 0.0], device="cuda").view(3, 4)
    assert argmin(x).cpu().tolist() == [0, 3, 3]


@pytest.mark.parametrize("shape", [(1, 1), (7, 3), (300, 257), (64, 2048), (33, 2049), (5, 100_003), (2000, 4096)])
def test_vs_numpy_with_ties(shape):
    rng = np.random.default_rng(sum(shape))
    x = rng.integers(0, 50, size=shape).astype(np.float32)      # many exact ties: smaller index must win
    got = argmin(torch.from_numpy(x).cuda()).cpu().numpy()
    assert (got == x.argmin(axis=1)).all()                       # numpy: first occurrence


def test_nan_and_inf_rows():
    x = torch.full((4, 37), float("inf"), device="cuda")
    x[1] = float("nan")
    x[2, 5] = float("nan"); x[2, 9] = 1.0
    x[3, 36] = -float("inf")
    assert argmin(x).cpu().tolist() == [0, 0, 9, 36]             # start value (0, +inf): nothing below it -> 0
