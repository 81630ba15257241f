"""Incremental == non-incremental (the reference's differential property tests) on the CPU oracle."""
import pytest

import differential_cases as dc


@pytest.mark.parametrize("name", sorted(dc.ALL_CASES))
def test_oracle_differential(oracle, name):
    dc.ALL_CASES[name](oracle)
