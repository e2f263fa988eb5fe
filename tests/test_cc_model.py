"""CPU: the data-parallel decomposition used by the CUDA connectivity kernels (cc_parallel_model)
reproduces the sequential reference algorithm exactly, including the BFS size cap, the
last-seen-neighbour merge target and the start_label=1 "label 0" re-scan corner case."""
import numpy as np
import pytest

import slic_oracle as so
from cc_parallel_model import enforce_connectivity_model


def _cases(seed, n):
    rng = np.random.RandomState(seed)
    for _ in range(n):
        H, W = int(rng.randint(3, 26)), int(rng.randint(3, 26))
        start_label = int(rng.randint(0, 2))
        if rng.rand() < 0.35:
            lab = rng.randint(0, rng.randint(1, 8), size=(H, W))
        else:
            by, bx = int(rng.randint(2, 9)), int(rng.randint(2, 9))
            yy, xx = np.mgrid[:H, :W]
            lab = (yy // by) * ((W + bx - 1) // bx) + xx // bx
            noise = rng.rand(H, W) < rng.choice([0.0, 0.05, 0.2])
            lab = np.where(noise, rng.randint(0, lab.max() + 1, size=(H, W)), lab)
        lab = lab + start_label
        if rng.rand() < 0.3:
            lab = np.where(rng.rand(H, W) < 0.2, start_label - 1, lab)
        if rng.rand() < 0.5:
            min_size = int(rng.randint(2, 8)); max_size = min_size * 6
        else:
            min_size = int(rng.randint(0, 12)); max_size = int(rng.randint(1, 40))
        yield lab, min_size, max_size, start_label


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_model_equals_sequential_oracle(seed):
    for lab, min_size, max_size, start_label in _cases(seed, 60):
        want = so.enforce_connectivity(lab, min_size, max_size, start_label)
        got, _ = enforce_connectivity_model(lab, min_size, max_size, start_label)
        np.testing.assert_array_equal(got, want)


def test_oracle_rejects_zero_max_size():
    with pytest.raises(ValueError):
        so.enforce_connectivity(np.ones((4, 4), int), 1, 0, 1)
