"""filter_map_test (operator/filter_map.rs:737-910) on the oracle; see filter_map_cases.py."""
import pytest

import filter_map_cases as fc


@pytest.mark.parametrize("name", sorted(fc.CASES))
def test_oracle_filter_map(oracle, name):
    fc.run_filter_map_case(oracle, name)


@pytest.mark.parametrize("with_closure", [False, True], ids=["index", "index_with"])
def test_oracle_index(oracle, with_closure):
    """index_test / index_with_test (operator/index.rs:240-300)."""
    fc.run_index_test(oracle, with_closure)


def test_oracle_neg_plus_and_sum_are_zero(oracle):
    """zset_sum (operator/neg.rs:85-110, operator/sum.rs:138-200)."""
    fc.run_neg_plus_zero(oracle)
    fc.run_sum_zero(oracle)


def test_oracle_filter_map_circuit(oracle):
    """filter_map_test as one circuit over Stream.{filter,map,flat_map,map_index,flat_map_index}."""
    fc.run_filter_map_circuit(oracle)


def test_oracle_sum_circuit(oracle):
    fc.run_sum_circuit(oracle)


def test_oracle_input_zset(oracle):
    """zset_test_st (operator/input.rs:1058-1100)."""
    fc.run_input_zset_test(oracle)
