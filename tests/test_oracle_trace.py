"""The reference's trace / merger property tests (spine_fueled.rs:1282-1347,
trace/layers/test.rs:734-879, trace/mod.rs:371-396) on the CPU oracle — pins
the oracle's Spine, truncation and fuelled Merger against the TestBatch model."""
import pytest

import trace_cases as tc


@pytest.mark.parametrize("name", sorted(tc.ALL_CASES))
def test_oracle_trace(oracle, name):
    tc.ALL_CASES[name](oracle)


@pytest.mark.parametrize("seed", range(20, 28))
def test_oracle_indexed_spine_seeds(oracle, seed):
    tc.run_indexed_zset_spine(oracle, seed=seed)
    tc.run_zset_spine(oracle, seed=seed)
