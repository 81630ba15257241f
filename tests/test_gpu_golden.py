"""The reference's golden vectors through the CUDA library (C ABI)."""
import pytest

import golden_cases as gc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(gc.ALL_CASES))
def test_cuda_golden(cuda, name):
    gc.ALL_CASES[name](cuda)
