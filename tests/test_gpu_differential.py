"""The reference's differential property tests (operator/aggregate/mod.rs:706-876 incl. stream_aggregate
:362-376, operator/distinct.rs:892-1081, operator/join.rs:136-155) and the semijoin model tests on the CUDA path."""
import pytest

import differential_cases as dc
import semijoin_cases as sc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(dc.ALL_CASES))
def test_cuda_differential(cuda, name):
    dc.ALL_CASES[name](cuda)


@pytest.mark.parametrize("name", sorted(sc.ALL_CASES))
def test_cuda_semijoin(cuda, name):
    sc.ALL_CASES[name](cuda)
