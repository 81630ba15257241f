"""Trace checkpoint / resume on the CPU oracle (file format shared with the CUDA library)."""
import pytest

import snapshot_cases as sc
from dbsp_b200 import Schema


@pytest.mark.parametrize("schema", [Schema("u"), Schema("u", "u"), Schema("ui", "iu")], ids=lambda s: f"{s.key}_{s.val}")
def test_oracle_snapshot_roundtrip(oracle, tmp_path, schema):
    sc.run_snapshot_roundtrip(oracle, tmp_path, schema)


def test_oracle_snapshot_errors(oracle, tmp_path):
    sc.run_snapshot_errors(oracle, tmp_path)
