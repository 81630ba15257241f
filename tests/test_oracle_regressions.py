"""Pins the CPU oracle against the reference's saved proptest regressions
(consolidation.txt, consolidation/tests/proptests.txt,
merge_batcher/tests.txt, operator/distinct.txt)."""
import pytest

import regression_cases as rc


@pytest.mark.parametrize("kind,i", rc.all_cases())
def test_oracle_regression(oracle, kind, i):
    rc.run(oracle, kind, i)
