"""join_trace_test (operator/join.rs:1035-1113) on the CUDA library; see recursion_cases.py."""
import pytest

import recursion_cases as rc

pytestmark = pytest.mark.gpu


def test_cuda_join_trace_test(cuda):
    rc.run_join_trace_test(cuda)
