"""The reference's literal transitive-closure test (`join_trace_test`, operator/join.rs:1035-1113): the reachability
relation after each of 8 edge deltas.  The reference computes it in a nested scope (`circuit.recursive`); nested
timestamps are out of scope here, so the fixed point is iterated by the host over the same operators —
paths' = distinct(edges + join(index(map(paths, (x,y)->(y,x))), index(edges), |via, from, to| (from, to))) — which pins
`Join::eval`, `map_index`, `plus` and `distinct` composed the way the reference composes them, against its literals."""
from dbsp_b200 import Proj, Schema, key, lval, rval, val

EDGE_DELTAS = [
    {(1, 2): 1}, {(2, 3): 1}, {(1, 3): 1}, {(3, 1): 1}, {(3, 1): -1}, {(1, 2): -1}, {(2, 4): 1, (4, 1): 1}, {(2, 3): -1, (3, 2): 1},
]
EXPECTED = [
    {(1, 2)},
    {(1, 2), (2, 3), (1, 3)},
    {(1, 2), (2, 3), (1, 3)},
    {(1, 1), (2, 2), (3, 3), (1, 2), (1, 3), (2, 3), (2, 1), (3, 1), (3, 2)},
    {(1, 2), (2, 3), (1, 3)},
    {(2, 3), (1, 3)},
    {(1, 3), (2, 3), (2, 4), (2, 1), (4, 1), (4, 3)},
    {(a, b) for a in (1, 2, 3, 4) for b in (1, 2, 3, 4)},
]
PAIRS = Schema("uu")
INDEXED = Schema("u", "u")


def closure(be, edges):
    """Least fixed point of paths = distinct(edges + paths_inverted ⋈ edges); `edges` is an OrdZSet<(from,to)>."""
    edges_indexed = be.reindex(edges, 1)                                   # index(): from -> to
    paths = be.batch_empty(PAIRS)
    for _ in range(64):
        inverted = be.map_index(be.reindex(paths, 1), Proj(INDEXED, [val(0), key(0)]))    # (x, y) -> y: x
        joined = be.join_batches(inverted, edges_indexed, Proj(PAIRS, [lval(0), rval(0)]))  # |via, from, to| (from, to)
        nxt = be.stream_distinct(be.merge(edges, joined))
        if nxt == paths:
            return paths
        paths = nxt
    raise AssertionError("no fixed point in 64 rounds")


def run_join_trace_test(be):
    edges = be.batch_empty(PAIRS)
    for delta, want in zip(EDGE_DELTAS, EXPECTED):
        d = be.batch_from_rows(PAIRS, [(a, b, w) for (a, b), w in delta.items()])
        edges = be.merge(edges, d)
        got = closure(be, edges)
        rows = got.rows()
        assert {(int(r[0]), int(r[1])) for r in rows} == want and all(int(r[-1]) == 1 for r in rows), (delta, rows)
