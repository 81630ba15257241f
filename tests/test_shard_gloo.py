"""N>1 host path on CPU: world_size-2 gloo, oracle backend.  The sharded
circuit's gathered output must equal the single-worker output (the property
the reference tests in join_test_mt / test_shard, join.rs:1019-1033,
communication/shard.rs:264-307)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


WORLDS = [2, 3]   # 2: receiver merges the peers' sorted segments; > 2: small deltas are consolidated in one go


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, query, n_events, step, outdir):
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import dbsp_b200
    from dbsp_b200.nexmark import NexmarkGenerator
    from dbsp_b200.parallel import Comm
    from oracle_backend import OracleBackend
    from parity_util import build_query, feed

    be = OracleBackend()
    comm = Comm()
    c, handles, _ = build_query(be, query, comm)
    # rebuild with a gather on the output
    from dbsp_b200 import RootCircuit
    from dbsp_b200.nexmark import queries as nq

    c = RootCircuit(be, comm)
    inp, handles = nq.add_nexmark_input(c)
    out = nq.QUERIES[query](inp).gather(0).output()
    gen = NexmarkGenerator()
    results = []
    for s0 in range(0, n_events, step):
        t = gen.tables(s0, step)
        mine = {k: [col[rank::world].copy() for col in v] for k, v in t.items()}   # round-robin input (input.rs:664-703)
        feed(handles, mine)
        c.step()
        if rank == 0:
            results.append(out.value.rows())
    if rank == 0:
        np.save(os.path.join(outdir, "sharded.npy"), np.array(results, dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", WORLDS)
@pytest.mark.parametrize("query", ["q3", "q4", "q7"])
def test_sharded_equals_single(tmp_path, oracle, query, world):
    from parity_util import build_query, feed
    from dbsp_b200.nexmark import NexmarkGenerator

    n_events, step = (120_000, 40_000) if query != "q7" else (400_000, 100_000)
    port = _free_port()
    mp.spawn(_worker, args=(world, port, query, n_events, step, str(tmp_path)), nprocs=world, join=True)
    sharded = np.load(os.path.join(tmp_path, "sharded.npy"), allow_pickle=True)
    c, handles, out = build_query(oracle, query)
    gen = NexmarkGenerator()
    for i, s0 in enumerate(range(0, n_events, step)):
        feed(handles, gen.tables(s0, step))
        c.step()
        assert out.value.rows() == list(sharded[i]), f"{query} step {i}"
