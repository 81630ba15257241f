"""Checkpoint / resume of a trace (dbsp_spine_save / dbsp_spine_load): a loaded spine shows cursors
the same contents, keeps the key / value bounds and continues the same merge schedule."""
import os

import numpy as np

from dbsp_b200 import Schema, Spine
from trace_cases import rand_rows, signed_rows


def build_spine(be, s, seed, n_inserts=23):
    rng = np.random.default_rng(seed)
    tr = Spine(be, s)
    rows_all = []
    for i in range(n_inserts):
        n = int(rng.choice([1, 7, 90, 800, 4000]))
        rows = rand_rows(rng, n, 1 << 10, 1 << 6, 3, nk=s.nk, nv=s.nv)
        tr.insert(be.batch_from_rows(s, rows))
        rows_all.append(rows)
        if i == 9:
            tr.truncate_keys_below([40] * s.nk)
        if i == 15 and s.nv:
            tr.truncate_values_below([5] * s.nv)
    return tr, rng


def run_snapshot_roundtrip(be, tmpdir, schema=Schema("u", "u"), seed=21):
    tr, rng = build_spine(be, schema, seed)
    path = os.path.join(str(tmpdir), f"spine_{be.name}_{schema.key}_{schema.val}.bin")
    tr.save(path)
    back = Spine.load(be, path, schema)
    assert back.stats() == tr.stats()
    assert signed_rows(back.consolidate()) == signed_rows(tr.consolidate())
    # both continue identically (same bounds, same schedule)
    for _ in range(9):
        rows = rand_rows(rng, int(rng.integers(1, 2000)), 1 << 10, 1 << 6, 3, nk=schema.nk, nv=schema.nv)
        for sp in (tr, back):
            sp.insert(be.batch_from_rows(schema, rows))
        assert back.stats() == tr.stats()
    assert signed_rows(back.consolidate()) == signed_rows(tr.consolidate())
    return path


def run_snapshot_errors(be, tmpdir):
    import pytest
    from dbsp_b200._capi import DbspError

    bad = os.path.join(str(tmpdir), "not_a_snapshot.bin")
    open(bad, "wb").write(b"hello world, this is not a spine")
    with pytest.raises(DbspError):
        Spine.load(be, bad, Schema("u"))
    with pytest.raises(DbspError):
        Spine.load(be, os.path.join(str(tmpdir), "missing.bin"), Schema("u"))
