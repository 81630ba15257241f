#!/usr/bin/env python
"""bench.py — Nexmark events/s through the B200 Z-set hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A *step* is one `Circuit::step()` of the query over one batch of synthetic
Nexmark events (E events per rank per step).  The primary workload at N=1 is
BASELINE.json configs[1]: Nexmark q3 (person |x| auction incremental join),
100 M events (E = 5 M, W+K = 20 steps).  q4 and q7 ride along in `queries`
(same metric, same harness) unless --query pins one.

`value`   : events/s with the step's input columns already resident in HBM.
`e2e`     : the same steps through the C ABI with HOST (pinned) columns — H2D
            copies of every step's inputs and a D2H download of the step's
            output Z-set are inside the timed region.
`roofline`: the step's dominant kernel class, algorithmic bytes / device time
            from CUDA events the library records on its own stream
            (dbsp_ctx_profile), against MEASURED_PEAKS.json.
`--impl reference`: the CPU oracle (C++ restatement of the reference's
            algorithms; the Rust reference cannot be built here) on all host
            threads, N worker replicas with the reference's hash-shard/exchange
            scheme, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "nexmark_events_per_sec"
UNIT = "events/s"
QUERY_COLS = {   # columns each query actually reads (the others are not generated / copied)
    "q3": ("person", "auction"),
    "q4": ("auction", "bid"),
    "q7": ("bid",),
    "q0": ("bid",),
}
FULL_EVENTS = {"q3": 100_000_000, "q4": 100_000_000, "q7": 1_000_000_000, "q0": 1_000_000}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML, ~2 ms period)."""

    HW_SLOWDOWN, SW_THERMAL, HW_THERMAL, SW_POWER_CAP = 0x8, 0x20, 0x40, 0x4

    def __init__(self, gpu_index=0):
        self.idx, self.rows, self.stop, self.t = gpu_index, [], False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            # CUDA_VISIBLE_DEVICES may remap indices: NVML enumerates physical devices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv is not None:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.t is not None:
            self.t.join(timeout=1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [r[0] for r in self.rows]
        bits = 0
        for r in self.rows:
            bits |= r[1]
        names = [("hw_slowdown", self.HW_SLOWDOWN), ("hw_thermal_slowdown", self.HW_THERMAL),
                 ("sw_thermal_slowdown", self.SW_THERMAL), ("sw_power_cap", self.SW_POWER_CAP)]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.max_sm), "reasons": [n for n, b in names if bits & b],
                "samples": len(sm)}


def gen_steps(query, rank, world, n_steps, events_per_step, pinned):
    """Host column tables of n_steps steps for this rank.  Rank r's step s is
    the contiguous event range [(s*world + r) * E, +E) (round-robin input
    distribution at batch granularity, operator/input.rs:664-703)."""
    from dbsp_b200.nexmark import NexmarkGenerator

    gen = NexmarkGenerator(threads=max(1, (os.cpu_count() or 8) // max(world, 1)))
    want = QUERY_COLS[query]
    if pinned:
        import torch

        def alloc(k):
            return torch.empty(k, dtype=torch.int64).pin_memory().numpy().view(np.uint64)
    else:
        alloc = None
    steps = []
    for s in range(n_steps):
        first = (s * world + rank) * events_per_step
        t = gen.tables(first, events_per_step, want=want, alloc=alloc)
        steps.append(t)
    return steps


def empty_cols():
    return [np.empty(0, np.uint64) for _ in range(5)]


def feed_host(handles, t):
    for k in ("person", "auction", "bid"):
        handles[k].set(t[k] if t[k] is not None else empty_cols())


def run_b200(args, query, rank, world, comm, device, do_e2e=True):
    """Returns dict with value / e2e / profile for `query` on this rank."""
    import torch

    import dbsp_b200
    from dbsp_b200.nexmark import queries as nq
    from dbsp_b200.runtime import Runtime

    E, W, K = args.events_per_step, args.warmup, args.steps
    steps = gen_steps(query, rank, world, W + K, E, pinned=True)
    res = {}

    def build(be):
        c = dbsp_b200.RootCircuit(be, comm)
        inp, handles = nq.add_nexmark_input(c)
        out = nq.QUERIES[query](inp).output()
        build.tables = inp
        return c, handles, out

    def sync_all(be):
        be.sync()
        if comm is not None:
            torch.cuda.synchronize(device)
            comm.barrier()

    def max_over_ranks(x):
        if comm is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # ---- (1) device-resident inputs -------------------------------------------
    os.environ.setdefault("DBSP_POOL_RESERVE_GB", "24")   # pool growth (cudaMalloc of slabs) stays out of the timed region
    be = Runtime(device.index)
    ext = torch.cuda.ExternalStream(be.stream_ptr, device=device)
    dev_steps = []
    for t in steps:
        d = {}
        for k, cols in t.items():
            d[k] = None if cols is None else [torch.from_numpy(c.view(np.int64)).to(device, non_blocking=True) for c in cols]
        dev_steps.append(d)
    torch.cuda.synchronize(device)
    c, handles, out = build(be)

    def feed_dev(d):
        for k in ("person", "auction", "bid"):
            if d[k] is None:
                handles[k].set(empty_cols())
            else:
                handles[k].set_device([int(x.data_ptr()) for x in d[k]], int(d[k][0].numel()))

    for s in range(W):
        feed_dev(dev_steps[s])
        c.step()
    sync_all(be)
    be.stats(reset=True)
    be.profile(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(device.index) as clk:
        t0 = time.perf_counter()
        ev0.record(ext)
        for s in range(W, W + K):
            feed_dev(dev_steps[s])
            c.step()
        ev1.record(ext)
        sync_all(be)
        wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    ms = max_over_ranks(dev_ms)
    st = be.stats()
    prof = be.profile_read()
    be.profile(False)
    out_rows = len(out.value)
    res.update(value=world * E * K / (ms / 1e3), ms_per_step=ms / K, wall_ms_per_step=wall * 1e3 / K,
               gpu_launches=st["kernel_launches"], profile=prof, clocks=clk.summary(), last_step_out_rows=out_rows)
    del c, handles, out, dev_steps
    be.sync()

    # ---- (2) end to end: host columns in, output Z-set out -----------------------------
    if do_e2e:
        be2, ext2 = be, ext   # same context / stream / memory pool, fresh circuit state
        c, handles, out = build(be2)
        tabs = {"person": build.tables.person, "auction": build.tables.auction, "bid": build.tables.bid}
        masks = {k: tabs[k].table_mask() for k in tabs}

        def start_uploads(t):
            """H2D of one step's tables on the copy stream (only the columns the query reads)."""
            return {k: (be2.upload_begin(t[k], masks[k]) if (t[k] is not None and masks[k]) else None) for k in tabs}

        def feed_uploads(ups):
            for k in tabs:
                if ups[k] is None:
                    handles[k].set(empty_cols())
                else:
                    handles[k].set_upload(ups[k])

        for s in range(W):
            feed_uploads(start_uploads(steps[s]))
            c.step()
            out.value.download()
        sync_all(be2)
        be2.stats(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(ext2)
        nxt = start_uploads(steps[W])             # every step's H2D copy is inside the timed region ...
        for s in range(W, W + K):
            cur = nxt
            nxt = start_uploads(steps[s + 1]) if s + 1 < W + K else None   # ... and overlaps the previous step's kernels
            feed_uploads(cur)
            c.step()
            out.value.download()                  # D2H of the step's result Z-set
        e1.record(ext2)
        sync_all(be2)
        wall2 = time.perf_counter() - t0
        ms2 = max_over_ranks(max(e0.elapsed_time(e1), wall2 * 1e3))   # host copies are part of the path: take the larger clock
        st2 = be2.stats()
        res["e2e"] = {"value": world * E * K / (ms2 / 1e3), "unit": UNIT, "h2d_bytes_per_step": st2["h2d_bytes"] // K,
                      "d2h_bytes_per_step": st2["d2h_bytes"] // K, "ms_per_step": ms2 / K}
        del c, handles, out
    be.sync()
    import gc

    gc.collect()
    be.close()
    return res


ROOFLINE_NOTES = {
    "q3": "q3 reads only Person/Auction events (8% of the stream): a 5M-event step is ~60k-row batches, so every kernel "
          "is launch-latency bound and the dominant class (radix passes on 60k rows) sits far below the HBM roofline; "
          "see queries.q4.roofline and merge_sweep.roofline for the bandwidth-bound kernels",
}


def roofline_from_profile(prof, query=None):
    peak, how = peaks()
    if not prof:
        return None
    name = max(prof, key=lambda k: prof[k]["ms"])
    p = prof[name]
    total_ms = sum(v["ms"] for v in prof.values())
    achieved = p["alg_bytes"] / (p["ms"] / 1e3) / 1e9 if p["ms"] > 0 else 0.0
    return {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "launches": p["launches"], "avg_launch_us": 1e3 * p["ms"] / p["launches"],
            "alg_bytes_per_launch": p["alg_bytes"] / p["launches"], "share_of_kernel_time": p["ms"] / total_ms if total_ms else None,
            "peak_source": how, **({"note": ROOFLINE_NOTES[query]} if query in ROOFLINE_NOTES else {})}


def merge_sweep(device_index, rows=50_000_000, n_val_lanes=1):
    """BASELINE.json configs[4]: merge of two consolidated OrdIndexedZSet<u64,u64,i64>
    batches (Zipf-ish keys), algorithmic GB/s of the merge kernel vs the HBM peak."""
    import torch

    from dbsp_b200 import Schema
    from dbsp_b200.runtime import Runtime

    be = Runtime(device_index)
    dev = torch.device("cuda", device_index)
    g = torch.Generator(device=dev)
    g.manual_seed(0x7FC359184519C0AA & 0x7FFFFFFF)
    s = Schema("u", "u" * n_val_lanes)
    batches = []
    for _ in range(2):
        # Zipf(s=1) keys over a domain of rows/4: inverse-CDF of a log-uniform draw
        u = torch.rand(rows, generator=g, device=dev, dtype=torch.float64)
        dom = rows // 4
        keys = torch.exp(u * np.log(dom)).to(torch.int64).clamp_(1, dom)
        vals = [torch.randint(0, 1 << 40, (rows,), generator=g, device=dev, dtype=torch.int64) for _ in range(n_val_lanes)]
        w = torch.randint(0, 4, (rows,), generator=g, device=dev, dtype=torch.int64)
        w = torch.where(w >= 2, w - 1, w - 2)   # {-2,-1,1,2}
        torch.cuda.synchronize(dev)
        batches.append(be.batch_from_columns(s, [int(keys.data_ptr())] + [int(v.data_ptr()) for v in vals], int(w.data_ptr()), n=rows, on_device=True))
        be.sync()
        del u, keys, vals, w
    a, b = batches
    for _ in range(3):
        be.merge(a, b)
    be.profile(True)
    for _ in range(5):
        m = be.merge(a, b)
    prof = be.profile_read()
    be.profile(False)
    peak, how = peaks()
    p = prof["merge_tiles"]
    ach = p["alg_bytes"] / (p["ms"] / 1e3) / 1e9
    traffic = None
    try:   # dram__bytes_read+write of one launch from the committed ncu --set full capture (2 x 20 M rows), scaled by rows
        if n_val_lanes != 1:
            raise KeyError("capture is for 2-lane rows")
        prof_j = json.load(open(os.path.join(ROOT, "profiles", "r1_merge_tiles_ncu_full.json")))
        def nbytes(x):   # "960.49 Mbyte" -> bytes
            v, u = x.split()[:2]
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        traffic = (nbytes(prof_j["dram__bytes_read.sum"]) + nbytes(prof_j["dram__bytes_write.sum"])) * (len(a) + len(b)) / 40_000_000
    except Exception:
        pass
    return {"workload": f"merge 2 x OrdIndexedZSet<u64,{'(' + ','.join(['u64'] * n_val_lanes) + ')' if n_val_lanes > 1 else 'u64'},i64>, {len(a)}+{len(b)} rows -> {len(m)}", "rows_per_s": (len(a) + len(b)) / (p["ms"] / 5 / 1e3),
            "roofline": {"bound": "hbm", "kernel": "merge_tiles", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic, "traffic_source": "ncu dram__bytes_read.sum+dram__bytes_write.sum, profiles/r1_merge_tiles_ncu_full.json (2x20M-row launch) scaled by rows",
                         "alg_bytes_per_launch": p["alg_bytes"] / p["launches"], "avg_launch_us": 1e3 * p["ms"] / p["launches"], "peak_source": how},
            "inputs_larger_than_l2": True}


def run_reference(args, query, n_workers=None, budget_s=25.0):
    """The CPU arm: oracle circuit replicas on all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dbsp_b200
    from dbsp_b200.nexmark import NexmarkGenerator
    from dbsp_b200.nexmark import queries as nq
    from oracle_backend import OracleBackend
    from thread_workers import run_workers

    T = n_workers or min(os.cpu_count() or 1, 16)
    E, W, K = args.events_per_step, args.warmup, args.steps
    # bounded sample: CPU steps are sized so that W+K of them fit the budget
    cpu_E = args.cpu_events_per_step
    gen = NexmarkGenerator()
    want = QUERY_COLS[query]
    tables = [gen.tables(s * cpu_E, cpu_E, want=want) for s in range(W + K)]

    def worker(rank, comm):
        be = OracleBackend()
        c = dbsp_b200.RootCircuit(be, comm if T > 1 else None)
        inp, handles = nq.add_nexmark_input(c)
        out = nq.QUERIES[query](inp).output()
        times = []
        for s, t in enumerate(tables):
            mine = {k: (None if v is None else [col[rank::T] for col in v]) for k, v in t.items()}
            comm.barrier()
            t0 = time.perf_counter()
            feed_host(handles, mine)
            c.step()
            comm.barrier()
            times.append(time.perf_counter() - t0)
        return times

    all_times = run_workers(T, worker)
    per_step = np.max(np.array(all_times), axis=0)
    timed = per_step[W:]
    secs = float(timed.sum())
    return {"value": cpu_E * K / secs, "cores": T, "kind": "port", "ms_per_step": 1e3 * secs / K,
            "sample": f"{query}: {W}+{K} steps of {cpu_E} events on {T} oracle worker threads (hash-shard + in-process exchange)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--query", default=None, help="q3 | q4 | q7 (default: q3 primary + q4, q7 riding along)")
    ap.add_argument("--events-per-step", type=int, default=5_000_000)
    ap.add_argument("--cpu-events-per-step", type=int, default=400_000)
    ap.add_argument("--no-extras", action="store_true", help="skip q4/q7 and the merge sweep")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    primary = args.query or "q3"
    E, W, K = args.events_per_step, args.warmup, args.steps
    cfg = {"workload": f"Nexmark {primary}, {E} events/step/GPU x {K} timed steps (+{W} warm-up) = {world * E * (W + K)} events; "
                       f"BASELINE configs[1..3] full size {FULL_EVENTS[primary]}",
           "events_per_step_per_gpu": E, "query": primary, "l2": "inputs larger than L2 (each step streams fresh event columns); traces grow past L2",
           "parallelism": f"key-hash shard x{world}" if world > 1 else "single GPU"}

    if args.impl == "reference":
        if rank != 0:
            return
        r = run_reference(args, primary)
        line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
                "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
                "data": "synthetic", "impl": "reference", "config": cfg,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: the hot path has no CPU fallback (use --impl reference for the CPU arm)"}))
        sys.exit(2)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    comm = None
    if world > 1:
        import torch.distributed as dist

        from dbsp_b200.parallel import Comm

        dist.init_process_group("nccl", device_id=device)
        comm = Comm(device)

    res = run_b200(args, primary, rank, world, comm, device)
    line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64", "data": "synthetic", "config": cfg, "clocks": res["clocks"], "e2e": res.get("e2e"),
            "gpu_launches": res["gpu_launches"], "roofline": roofline_from_profile(res["profile"], primary),
            "kernel_profile": {k: {"launches": v["launches"], "ms": round(v["ms"], 3), "alg_GB": round(v["alg_bytes"] / 1e9, 4)} for k, v in res["profile"].items()}}
    if comm is not None:
        line["nvlink_bytes_sent_rank0"] = comm.bytes_sent

    if not args.no_extras and args.query is None:
        extras = {}
        for q in ("q4", "q7"):
            r = run_b200(args, q, rank, world, comm, device)
            extras[q] = {"value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"], "e2e": r.get("e2e"),
                         "gpu_launches": r["gpu_launches"], "roofline": roofline_from_profile(r["profile"]),
                         "kernel_profile": {k: {"launches": v["launches"], "ms": round(v["ms"], 3), "alg_GB": round(v["alg_bytes"] / 1e9, 4)} for k, v in r["profile"].items()},
                         "events": world * E * (W + K)}
        line["queries"] = extras
        if world == 1:
            line["merge_sweep"] = merge_sweep(local)

    if rank == 0 and world == 1:
        cb = run_reference(args, primary)
        line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": cb["cores"], "kind": cb["kind"], "sample": cb["sample"]}
        if "queries" in line:
            for q in ("q4", "q7"):
                c2 = run_reference(args, q)
                line["queries"][q]["cpu_baseline"] = {"value": c2["value"], "unit": UNIT, "cores": c2["cores"], "kind": c2["kind"], "sample": c2["sample"]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
