"""SemiJoinStream::eval (operator/semijoin.rs:100-142) on the CPU oracle vs a pure-Python model."""
import pytest

import semijoin_cases as sc


@pytest.mark.parametrize("name", sorted(sc.ALL_CASES))
def test_oracle_semijoin(oracle, name):
    sc.ALL_CASES[name](oracle)
