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


# ---- propagate_test (operator/join.rs:1155-1243): labels pushed along the edges of an acyclic graph, no distinct —
# the weight of Label(node, l) is the number of paths that carry l to the node; the test's outputs are the per-step
# deltas of that multiset.  Host-iterated fixed point of result = labels + join(index(result), index(edges)).
PROP_EDGES = [
    {(1, 2): 1, (1, 3): 1, (2, 4): 1, (3, 4): 1}, {(5, 7): 1, (6, 7): 1}, {(4, 5): 1, (4, 6): 1}, {(3, 8): 1, (8, 9): 1},
    {(2, 4): -1, (7, 10): 1}, {(3, 4): -1}, {(1, 4): 1}, {(9, 7): 1},
]
PROP_LABELS = [{(1, 0): 1}, {(4, 1): 1}, {}, {(1, 0): -1, (1, 2): 1}, {}, {(8, 3): 1}, {(4, 1): -1}, {}]
PROP_EXPECTED = [
    {(1, 0): 1, (2, 0): 1, (3, 0): 1, (4, 0): 2},
    {(4, 1): 1},
    {(5, 0): 2, (5, 1): 1, (6, 0): 2, (6, 1): 1, (7, 0): 4, (7, 1): 2},
    {(1, 0): -1, (1, 2): 1, (2, 0): -1, (2, 2): 1, (3, 0): -1, (3, 2): 1, (4, 0): -2, (4, 2): 2, (5, 0): -2, (5, 2): 2,
     (6, 0): -2, (6, 2): 2, (7, 0): -4, (7, 2): 4, (8, 2): 1, (9, 2): 1},
    {(4, 2): -1, (5, 2): -1, (6, 2): -1, (7, 2): -2, (10, 1): 2, (10, 2): 2},
    {(4, 2): -1, (5, 2): -1, (6, 2): -1, (7, 2): -2, (8, 3): 1, (9, 3): 1, (10, 2): -2},
    {(4, 1): -1, (4, 2): 1, (5, 1): -1, (5, 2): 1, (6, 1): -1, (6, 2): 1, (7, 1): -2, (7, 2): 2, (10, 1): -2, (10, 2): 2},
    {(7, 2): 1, (7, 3): 1, (10, 2): 1, (10, 3): 1},
]


def propagate(be, edges, labels):
    edges_indexed = be.reindex(edges, 1)                                   # index_with(|e| (e.0, e.1))
    result = labels
    for _ in range(64):
        computed = be.join_batches(be.reindex(result, 1), edges_indexed, Proj(PAIRS, [rval(0), lval(0)]))  # Label(to, label)
        nxt = be.merge(labels, computed)
        if nxt == result:
            return result
        result = nxt
    raise AssertionError("no fixed point in 64 rounds (cyclic graph?)")


def run_propagate_test(be):
    def zset(d):
        return be.batch_from_rows(PAIRS, [(a, b, w) for (a, b), w in d.items()])

    edges, labels, prev = be.batch_empty(PAIRS), be.batch_empty(PAIRS), be.batch_empty(PAIRS)
    for de, dl, want in zip(PROP_EDGES, PROP_LABELS, PROP_EXPECTED):
        edges, labels = be.merge(edges, zset(de)), be.merge(labels, zset(dl))
        cur = propagate(be, edges, labels)
        delta = be.merge(cur, be.neg(prev))
        got = {(int(r[0]), int(r[1])): int(r[-1]) for r in delta.rows()}
        assert got == want, (de, dl, got)
        prev = cur
