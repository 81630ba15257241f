"""Pins the CPU oracle against the reference's golden vectors
and exercises the host-side Stream layer — no GPU needed."""
import pytest

import golden_cases as gc


@pytest.mark.parametrize("name", sorted(gc.ALL_CASES))
def test_oracle_golden(oracle, name):
    gc.ALL_CASES[name](oracle)


@pytest.mark.parametrize("name", sorted(gc.ORACLE_ONLY_CASES))
def test_oracle_only_golden(oracle, name):
    gc.ORACLE_ONLY_CASES[name](oracle)
