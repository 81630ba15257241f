"""Turns the reference's saved proptest regressions (shrunk failing inputs that
the reference re-runs before every property test) into a JSON fixture:

    python tests/golden/make_regressions.py   # needs /root/reference; writes proptest_regressions.json.gz

Sources (crates/dbsp/proptest-regressions/):
  trace/consolidation.txt, trace/consolidation/tests/proptests.txt   -> "consolidation" ((key,val),weight) lists and
                                                                          "pairs" (u32,u32) lists of the quicksort test
  trace/ord/merge_batcher/tests.txt                                  -> "batcher_batches" (lists of pushed batches) and
                                                                          "batcher_state" (MergeSorter queue chunks + a batch)
  operator/distinct.txt                                              -> "distinct" (rounds of indexed Z-set deltas)
Only inputs are stored; expected outputs are the model the reference's tests
compare against (a BTreeMap sum), recomputed by tests/regression_cases.py.
"""
import ast
import gzip
import json
import os
import re

REF = "/root/reference/crates/dbsp/proptest-regressions"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "proptest_regressions.json.gz")


def cases(path):
    with open(os.path.join(REF, path)) as f:
        for i, line in enumerate(l for l in f if l.startswith("cc ")):
            yield f"{path}#{i}", line.split("# shrinks to ", 1)[1].strip()


def flat(tuples):   # [((k, v), w), ...] -> [[k, v, w], ...]
    return [[k, v, w] for ((k, v), w) in tuples]


def main():
    out = {"consolidation": [], "pairs": [], "batcher_batches": [], "batcher_state": [], "distinct": []}
    for path in ("trace/consolidation.txt", "trace/consolidation/tests/proptests.txt"):
        for src, text in cases(path):
            if text.startswith("batch = "):
                out["consolidation"].append({"source": src, "batch": flat(ast.literal_eval(text[len("batch = "):]))})
            elif text.startswith("mut data = "):
                out["pairs"].append({"source": src, "data": [list(p) for p in ast.literal_eval(text[len("mut data = "):])]})
            else:
                raise ValueError(src)
    for src, text in cases("trace/ord/merge_batcher/tests.txt"):
        if text.startswith("batches = "):
            out["batcher_batches"].append({"source": src, "batches": [flat(b) for b in ast.literal_eval(text[len("batches = "):])]})
        else:
            m = re.match(r"mut merger = MergeSorter \{ queue: (.*), stash: \[\] \}(?:, mut batch = (.*))?$", text)
            queue = ast.literal_eval(m.group(1))   # lists of sorted chunks
            batch = ast.literal_eval(m.group(2)) if m.group(2) else []
            out["batcher_state"].append({"source": src, "queue": [[flat(chunk) for chunk in lst] for lst in queue], "batch": flat(batch)})
    for src, text in cases("operator/distinct.txt"):
        body = re.sub(r", workers = \d+$", "", text[len("inputs = "):])
        body = re.sub(r"OrdIndexedZSet \{ layer: OrderedLayer \{ keys: (\[[^\]]*\]), offs: (\[[^\]]*\]), "
                      r"vals: ColumnLayer \{ keys: (\[[^\]]*\]), diffs: (\[[^\]]*\]) \} \} \}",
                      r'{"keys": \1, "offs": \2, "vals": \3, "diffs": \4}', body)
        out["distinct"].append({"source": src, "rounds": ast.literal_eval(body)})
    with gzip.open(OUT, "wt") as f:
        json.dump(out, f, separators=(",", ":"))
    print({k: len(v) for k, v in out.items()}, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
