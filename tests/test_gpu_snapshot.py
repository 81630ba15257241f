"""Trace checkpoint / resume on the CUDA library, and across the two libraries (same file format)."""
import numpy as np
import pytest

import snapshot_cases as sc
from dbsp_b200 import Schema, Spine
from parity_util import assert_batches_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("schema", [Schema("u"), Schema("u", "u"), Schema("ui", "iu")], ids=lambda s: f"{s.key}_{s.val}")
def test_cuda_snapshot_roundtrip(cuda, tmp_path, schema):
    sc.run_snapshot_roundtrip(cuda, tmp_path, schema)


def test_cuda_snapshot_errors(cuda, tmp_path):
    sc.run_snapshot_errors(cuda, tmp_path)


def test_snapshot_crosses_libraries(cuda, oracle, tmp_path):
    s = Schema("u", "uu")
    tc, _ = sc.build_spine(cuda, s, 5)
    to, _ = sc.build_spine(oracle, s, 5)
    pc, po = str(tmp_path / "cuda.bin"), str(tmp_path / "oracle.bin")
    tc.save(pc)
    to.save(po)
    assert_batches_equal(Spine.load(cuda, po, s).consolidate(), tc.consolidate(), "oracle snapshot loaded by CUDA")
    assert_batches_equal(Spine.load(oracle, pc, s).consolidate(), to.consolidate(), "CUDA snapshot loaded by the oracle")


def test_snapshot_large(cuda, tmp_path):
    rng = np.random.default_rng(1)
    s = Schema("u", "u")
    tr = Spine(cuda, s)
    for n in (3_000_000, 700_000, 900_000):
        tr.insert(cuda.batch_from_columns(s, [rng.integers(0, 1 << 22, n).astype(np.uint64), rng.integers(0, 1 << 30, n).astype(np.uint64)], rng.integers(-1, 3, n)))
    p = str(tmp_path / "big.bin")
    tr.save(p)
    back = Spine.load(cuda, p, s)
    assert back.stats() == tr.stats()
    assert_batches_equal(back.consolidate(), tr.consolidate(), "large snapshot")
