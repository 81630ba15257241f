"""Model-based tests of the batch / trace boundary, written once against the
backend API and run on the CPU oracle (`-m "not gpu"`) and on the CUDA library
(`-m gpu`).  They restate the reference's own property tests with a seeded
numpy RNG in place of proptest:

* `spine_fueled.rs:1282-1347` — `test_zset_spine`, `test_indexed_zset_spine`,
  `test_truncate_value_bounded_memory` against the `TestBatch` model
  (`trace/test_batch.rs:21-170, 700-836`: a BTreeMap of tuples with
  `lower_key_bound` / `lower_val_bound` = max of the bounds seen);
* `trace/layers/test.rs:734-879` — merge == map merge, `truncate_below` for
  every bound;
* `trace/mod.rs:371-396` — the fuelled Merger contract.
"""
import numpy as np

from dbsp_b200 import Schema, Spine


class Model:
    """TestBatch (trace/test_batch.rs): {(key.., val..): weight} + bounds."""

    def __init__(self, schema):
        self.s = schema
        self.data = {}
        self.kb = None
        self.vb = None

    def insert(self, rows):
        nk = self.s.nk
        for r in rows:
            t, w = tuple(r[:-1]), r[-1]
            if self.kb is not None and t[:nk] < self.kb:
                continue
            if self.vb is not None and t[nk:] < self.vb:
                continue
            self.data[t] = self.data.get(t, 0) + w
            if self.data[t] == 0:
                del self.data[t]

    def truncate_keys_below(self, kb):
        kb = tuple(kb)
        self.kb = kb if self.kb is None else max(self.kb, kb)
        nk = self.s.nk
        self.data = {t: w for t, w in self.data.items() if t[:nk] >= self.kb}

    def truncate_values_below(self, vb):
        vb = tuple(vb)
        self.vb = vb if self.vb is None else max(self.vb, vb)
        nk = self.s.nk
        self.data = {t: w for t, w in self.data.items() if t[nk:] >= self.vb}

    def rows(self):
        return [t + (w,) for t, w in sorted(self.data.items())]


def signed_rows(b):
    """rows() with i64 lanes read as signed (rows() already does) as tuples."""
    return [tuple(int(x) for x in r) for r in b.rows()]


def retain_vals(rows, nk, vb):
    return rows if vb is None else [r for r in rows if tuple(r[nk:-1]) >= tuple(vb)]


def rand_rows(rng, n, max_key, max_val, max_w, nk=1, nv=1):
    cols = [rng.integers(0, max_key, n) for _ in range(nk)] + [rng.integers(0, max_val, n) for _ in range(nv)]
    w = rng.integers(-max_w, max_w, n)
    return [tuple(int(c[i]) for c in cols) + (int(w[i]),) for i in range(n)]


# ---- spine_fueled.rs:1298-1319 test_zset_spine --------------------------------
def run_zset_spine(be, seed=1):
    rng = np.random.default_rng(seed)
    s = Schema("i")
    tr, ref = Spine(be, s), Model(s)
    for _ in range(int(rng.integers(1, 20))):
        rows = rand_rows(rng, int(rng.integers(0, 100)), 50, 1, 2, nk=1, nv=0)
        bound = int(rng.integers(0, 50))
        b = be.batch_from_rows(s, rows)
        one = Model(s)
        one.insert(rows)
        assert signed_rows(b) == one.rows()
        tr.insert(b)
        ref.insert(rows)
        assert signed_rows(tr.consolidate()) == ref.rows()
        tr.truncate_keys_below([bound])
        ref.truncate_keys_below([bound])
        assert signed_rows(tr.consolidate()) == ref.rows()


# ---- spine_fueled.rs:1321-1347 test_indexed_zset_spine ------------------------
def run_indexed_zset_spine(be, seed=2, schema=None, max_key=100, max_val=5):
    rng = np.random.default_rng(seed)
    s = schema or Schema("i", "i")
    tr, ref = Spine(be, s), Model(s)
    for _ in range(int(rng.integers(1, 20))):
        rows = rand_rows(rng, int(rng.integers(0, 500)), max_key, max_val, 2, nk=s.nk, nv=s.nv)
        kb = [int(rng.integers(0, max_key)) for _ in range(s.nk)]
        vb = [int(rng.integers(0, max_val)) for _ in range(s.nv)]
        tr.insert(be.batch_from_rows(s, rows))
        ref.insert(rows)
        # assert_trace_eq (test_batch.rs:146-168): compare above the value bound only
        assert retain_vals(signed_rows(tr.consolidate()), s.nk, ref.vb) == ref.rows()
        tr.truncate_keys_below(kb)
        ref.truncate_keys_below(kb)
        tr.truncate_values_below(vb)
        ref.truncate_values_below(vb)
        assert retain_vals(signed_rows(tr.consolidate()), s.nk, ref.vb) == ref.rows()


# ---- spine_fueled.rs:1282-1296 test_truncate_value_bounded_memory -------------
def run_truncate_value_bounded_memory(be, seed=3):
    """Monotone values (a sliding window of 20 per batch) + truncate_values_below:
    the trace must stay bounded because merges drop the values below the bound."""
    rng = np.random.default_rng(seed)
    s = Schema("i", "i")
    tr = Spine(be, s)
    for i in range(100):
        n = int(rng.integers(1, 500))
        rows = [(int(k), int(v), int(w)) for k, v, w in zip(rng.integers(0, 50, n), rng.integers(i * 20, i * 20 + 100, n),
                                                          rng.integers(1, 3, n))]
        tr.insert(be.batch_from_rows(s, rows))
        tr.truncate_values_below([i * 20])
        if i % 10 == 9:
            for _ in range(64):   # Trace::exert until reduced (the loop of Trace::consolidate, spine_fueled.rs:583-589):
                tr.exert(1 << 40)   # the merges apply the bound
                if tr.stats()[1] <= 1:
                    break
            n_tuples, _ = tr.stats()
            # live window: values in [i*20, i*20+100) for 50 keys
            assert n_tuples <= 50 * 200, n_tuples


# ---- trace/layers/test.rs:734-879 merge + truncate_below -----------------------
def run_merge_truncate(be, seed=4, schema=None):
    rng = np.random.default_rng(seed)
    s = schema or Schema("i", "i")
    for _ in range(5):
        left = rand_rows(rng, int(rng.integers(0, 400)), 30, 8, 3, nk=s.nk, nv=s.nv)
        right = rand_rows(rng, int(rng.integers(0, 400)), 30, 8, 3, nk=s.nk, nv=s.nv)
        a, b = be.batch_from_rows(s, left), be.batch_from_rows(s, right)
        m = Model(s)
        m.insert(left)
        m.insert(right)
        assert signed_rows(be.merge(a, b)) == m.rows()
        merged = be.merge(a, b)
        for bound in range(0, 31, 3):
            kb = [bound] * s.nk
            assert signed_rows(be.truncate_keys_below(merged, kb)) == [r for r in m.rows() if r[:s.nk] >= tuple(kb)]
        if s.nv:
            for bound in range(0, 9):
                vb = [bound] * s.nv
                got = signed_rows(be.merge(a, b, val_lower_bound=vb))
                assert got == [r for r in m.rows() if r[s.nk:-1] >= tuple(vb)], (bound, got)


# ---- trace/mod.rs:371-396 the fuelled Merger -----------------------------------
def run_merger_fuel(be, seed=5, schema=None, n=3000, fuels=(1, 7, 100, 1000, 1 << 30)):
    rng = np.random.default_rng(seed)
    s = schema or Schema("i", "i")
    left = rand_rows(rng, n, 200, 6, 3, nk=s.nk, nv=s.nv)
    right = rand_rows(rng, n // 2 + 1, 200, 6, 3, nk=s.nk, nv=s.nv)
    a, b = be.batch_from_rows(s, left), be.batch_from_rows(s, right)
    want = signed_rows(be.merge(a, b))
    for fuel in fuels:
        for vb in ([None] if not s.nv else [None, [3] * s.nv]):
            m = be.merger(a, b, vb)
            calls = 0
            while True:
                left_fuel = m.work(fuel)
                calls += 1
                assert calls < 10 * (len(a) + len(b)) + 10, "merger does not terminate"
                if left_fuel > 0:   # complete (trace/mod.rs:388-395)
                    break
            got = signed_rows(m.done())
            exp = want if vb is None else [r for r in want if r[s.nk:-1] >= tuple(vb)]
            assert got == exp, (fuel, vb)
    # done() before completion is an error, and the merger stays usable
    m = be.merger(a, b)
    if len(a) + len(b) > 2 and m.work(1) <= 0:
        try:
            m.done()
            raise AssertionError("done() on an incomplete merge must fail")
        except AssertionError:
            raise
        except Exception:
            pass
        while m.work(1 << 30) <= 0:
            pass
        assert signed_rows(m.done()) == want
    # empty inputs complete at once
    e = be.batch_empty(s)
    m = be.merger(e, e)
    assert m.work(1) > 0 and len(m.done()) == 0


def run_spine_exert(be, seed=6):
    """Trace::exert (spine_fueled.rs:561-581): effort moves merges along without changing what cursors see; with
    unbounded effort, repeated until the spine is reduced (the loop of Trace::consolidate, :583-589), one batch is left."""
    rng = np.random.default_rng(seed)
    s = Schema("u", "u")
    tr, ref = Spine(be, s), Model(s)
    for n in (4000, 1500, 600, 200, 70, 20, 5):
        rows = rand_rows(rng, n, 1 << 20, 1 << 20, 2)
        tr.insert(be.batch_from_rows(s, rows))
        ref.insert(rows)
    n0, nb0 = tr.stats()
    assert nb0 >= 2
    tr.exert(10)                   # a little effort: contents unchanged
    assert signed_rows(tr.consolidate()) == ref.rows()
    for _ in range(64):
        tr.exert(1 << 40)
        if tr.stats()[1] <= 1:
            break
    assert tr.stats()[1] == 1
    assert signed_rows(tr.consolidate()) == ref.rows()
    # inserting after compaction keeps working
    rows = rand_rows(rng, 100, 1 << 20, 1 << 20, 2)
    tr.insert(be.batch_from_rows(s, rows))
    ref.insert(rows)
    assert signed_rows(tr.consolidate()) == ref.rows()


def run_spine_schedule(be, seed=8):
    """The fuelled merge schedule (spine_fueled.rs:728-974): many inserts of mixed sizes — the number of batches a
    cursor has to visit stays logarithmic (at most two per layer), the contents always equal the model."""
    rng = np.random.default_rng(seed)
    s = Schema("u", "u")
    tr, ref = Spine(be, s), Model(s)
    total = 0
    for i in range(120):
        n = int(rng.choice([1, 3, 40, 700, 5000]))
        rows = rand_rows(rng, n, 1 << 12, 1 << 8, 2)
        tr.insert(be.batch_from_rows(s, rows))
        ref.insert(rows)
        total += n
        n_tuples, nb = tr.stats()
        assert nb <= 2 * (total.bit_length() + 2), (i, nb, total)
        if i % 7 == 0:
            assert signed_rows(tr.consolidate()) == ref.rows(), i
    assert signed_rows(tr.consolidate()) == ref.rows()


ALL_CASES = {
    "zset_spine": run_zset_spine,
    "indexed_zset_spine": run_indexed_zset_spine,
    "indexed_zset_spine_wide": lambda be: run_indexed_zset_spine(be, seed=12, schema=Schema("ui", "iu"), max_key=12, max_val=4),
    "truncate_value_bounded_memory": run_truncate_value_bounded_memory,
    "merge_truncate": run_merge_truncate,
    "merge_truncate_wide": lambda be: run_merge_truncate(be, seed=14, schema=Schema("iu", "ui")),
    "merge_truncate_zset": lambda be: run_merge_truncate(be, seed=15, schema=Schema("ii")),
    "merger_fuel": run_merger_fuel,
    "merger_fuel_zset": lambda be: run_merger_fuel(be, seed=16, schema=Schema("i")),
    "merger_fuel_wide": lambda be: run_merger_fuel(be, seed=17, schema=Schema("uu", "iuu"), n=1200, fuels=(13, 500, 1 << 30)),
    "spine_exert": run_spine_exert,
    "spine_schedule": run_spine_schedule,
}
