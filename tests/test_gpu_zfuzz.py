"""GPU parity on generated cases (cases.generated: jittered tanks with box + geometry fills and free-surface lids at random
heights, periodic channels; dr, dr_wall, jitter and fill seed drawn from the seed): the CUDA path through the C-ABI against
the CPU oracle, every stage bit-exact -- the same comparison as tests/test_gpu_parity.py, on inputs nobody tuned for.
(The oracle side of these cases is pinned against the reference's host build in tests/test_oracle.py.)"""
import pytest

import cases
from test_gpu_parity import _cmp_pipeline

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(8))
def test_generated_case_vs_oracle(seed, oracle, cuda_backend):
    c = cases.generated(seed)
    o = cases.run(oracle, c)
    g = cases.run(cuda_backend, c)
    _cmp_pipeline(o, g, c)
    assert g["nfluid"] > 0
