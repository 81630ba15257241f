"""The reference's saved proptest regressions through the CUDA library."""
import pytest

import regression_cases as rc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,i", rc.all_cases())
def test_cuda_regression(cuda, kind, i):
    rc.run(cuda, kind, i)
