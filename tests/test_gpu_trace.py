"""The batch / trace boundary on the CUDA library: the reference's property
tests against the TestBatch model (trace_cases.py), then bit-exact parity with
the CPU oracle at sizes that span many merge tiles."""
import numpy as np
import pytest

import trace_cases as tc
from dbsp_b200 import Schema, Spine
from parity_util import assert_batches_equal

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(tc.ALL_CASES))
def test_cuda_trace(cuda, name):
    tc.ALL_CASES[name](cuda)


@pytest.mark.parametrize("seed", range(20, 24))
def test_cuda_indexed_spine_seeds(cuda, seed):
    tc.run_indexed_zset_spine(cuda, seed=seed)
    tc.run_zset_spine(cuda, seed=seed)


def _pair(rng, be_list, s, na, nb, domain):
    out = []
    cols_a = [rng.integers(0, domain, na).astype(np.int64) for _ in range(s.nl)]
    cols_b = [rng.integers(0, domain, nb).astype(np.int64) for _ in range(s.nl)]
    wa, wb = rng.integers(-2, 3, na), rng.integers(-2, 3, nb)
    for be in be_list:
        out.append((be.batch_from_columns(s, cols_a, wa), be.batch_from_columns(s, cols_b, wb)))
    return out


@pytest.mark.parametrize("schema", [Schema("u", "u"), Schema("i", "iu"), Schema("uu", "uuu")], ids=lambda s: f"{s.key}_{s.val}")
def test_merge_bounded_parity(cuda, oracle, schema):
    rng = np.random.default_rng(31)
    (ga, gb), (oa, ob) = _pair(rng, [cuda, oracle], schema, 120_000, 90_000, 400)
    for bound in (0, 1, 200, 399, 400, 1 << 40):
        vb = [bound] * schema.nv
        assert_batches_equal(cuda.merge(ga, gb, val_lower_bound=vb), oracle.merge(oa, ob, val_lower_bound=vb), f"bounded {bound}")


@pytest.mark.parametrize("schema", [Schema("u"), Schema("u", "u"), Schema("ui", "uiu")], ids=lambda s: f"{s.key}_{s.val}")
@pytest.mark.parametrize("fuel", [1000, 33_333, 1 << 40])
def test_merger_parity(cuda, oracle, schema, fuel):
    rng = np.random.default_rng(37)
    (ga, gb), (oa, ob) = _pair(rng, [cuda, oracle], schema, 100_000, 150_000, 300)
    want = oracle.merge(oa, ob)
    m = cuda.merger(ga, gb)
    calls = 0
    while m.work(fuel) <= 0:
        calls += 1
        assert calls < 1000
    assert_batches_equal(m.done(), want, f"merger fuel={fuel}")
    if schema.nv:
        vb = [150] * schema.nv
        m = cuda.merger(ga, gb, vb)
        while m.work(fuel) <= 0:
            pass
        assert_batches_equal(m.done(), oracle.merge(oa, ob, val_lower_bound=vb), "bounded merger")


def test_spine_truncation_parity(cuda, oracle):
    """Same inserts / key bounds / value bounds / exert on both spines: every
    consolidated snapshot identical (both apply the bounds at the same points)."""
    rng = np.random.default_rng(41)
    s = Schema("u", "uu")
    sc, so = Spine(cuda, s), Spine(oracle, s)
    for i in range(16):
        n = int(rng.integers(1, 30_000))
        cols = [rng.integers(0, 500, n).astype(np.uint64), rng.integers(i * 10, i * 10 + 200, n).astype(np.uint64),
                rng.integers(0, 4, n).astype(np.uint64)]
        w = rng.integers(-1, 3, n)
        sc.insert(cuda.batch_from_columns(s, cols, w))
        so.insert(oracle.batch_from_columns(s, cols, w))
        if i % 3 == 2:
            for sp in (sc, so):
                sp.truncate_keys_below([i * 5])
                sp.truncate_values_below([i * 10, 2])
        if i % 5 == 4:
            sc.exert(50_000)
            so.exert(50_000)
        assert_batches_equal(sc.consolidate(), so.consolidate(), f"spine step {i}")
