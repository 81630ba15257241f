"""join_trace_test (operator/join.rs:1035-1113) on the oracle; see recursion_cases.py."""
import recursion_cases as rc


def test_oracle_join_trace_test(oracle):
    rc.run_join_trace_test(oracle)


def test_oracle_propagate_test(oracle):
    rc.run_propagate_test(oracle)
