"""ctypes binding of csrc/nexmark_gen.cpp (libnexmark_gen.so, host code).

Distributions follow crates/nexmark/src/generator/*.rs; see the C++ header
comment.  One call generates the Person / Auction / Bid column tables of a
contiguous range of event ids.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_PKG, "libnexmark_gen.so")
SRC_PATH = os.path.join(_PKG, "csrc", "nexmark_gen.cpp")

BASE_TIME = 1436918400000
STATES = ["AZ", "CA", "ID", "OR", "WA", "WY"]  # generator/people.rs:18-25 (sorted)

PERSON_COLS = ["id", "name", "city", "state", "date_time"]
AUCTION_COLS = ["id", "seller", "category", "date_time", "expires"]
BID_COLS = ["auction", "bidder", "price", "date_time", "extra"]


def state_code(s: str) -> int:
    return STATES.index(s)


def build_generator():
    if not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(SRC_PATH):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", LIB_PATH, SRC_PATH])
    return LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build_generator()
        _lib = C.CDLL(LIB_PATH)
        _lib.nexmark_counts.argtypes = [C.c_uint64, C.c_uint64] + [C.POINTER(C.c_uint64)] * 3
        _lib.nexmark_counts.restype = None
        _lib.nexmark_generate.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int] + [C.c_void_p] * 15
        _lib.nexmark_generate.restype = None
        _lib.nexmark_generate_rate.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int] + [C.c_void_p] * 15
        _lib.nexmark_generate_rate.restype = None
    return _lib


class NexmarkGenerator:
    """Seeded generator; `tables(first, n)` returns dict(person=[cols..],
    auction=[cols..], bid=[cols..]) as uint64 numpy columns."""

    def __init__(self, seed: int = 0x7FC359184519C0AA, threads: int | None = None, first_event_rate: int = 0):
        """first_event_rate: NexmarkConfig::first_event_rate in events/s (crates/nexmark/src/config.rs:51); 0 = the
        reference default of 10 M/s, i.e. 10 000 events per millisecond of event time."""
        self.seed = seed
        self.rate = int(first_event_rate)
        self.threads = threads or min(os.cpu_count() or 1, 32)
        self.lib = _load()

    def counts(self, first: int, n: int):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.lib.nexmark_counts(first, n, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def tables(self, first: int, n: int, want=("person", "auction", "bid"), alloc=None):
        np_, na, nb = self.counts(first, n)
        alloc = alloc or (lambda k: np.empty(k, dtype=np.uint64))
        out = {
            "person": [alloc(np_) for _ in PERSON_COLS] if "person" in want else None,
            "auction": [alloc(na) for _ in AUCTION_COLS] if "auction" in want else None,
            "bid": [alloc(nb) for _ in BID_COLS] if "bid" in want else None,
        }
        ptrs = []
        for t, ncol in (("person", 5), ("auction", 5), ("bid", 5)):
            cols = out[t]
            ptrs += [c.ctypes.data for c in cols] if cols is not None else [None] * ncol
        self.lib.nexmark_generate_rate(self.seed, self.rate, first, n, self.threads, *ptrs)
        return out
