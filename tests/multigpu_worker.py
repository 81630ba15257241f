"""torchrun worker (one process per GPU): the in-library exchange (csrc/comm.cu) and the sharded Nexmark circuits
over it, each checked against a single-GPU run of the same inputs on rank 0."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import numpy as np
import torch
import torch.distributed as dist

import dbsp_b200
from dbsp_b200 import RootCircuit, Schema
from dbsp_b200.nexmark import NexmarkGenerator
from dbsp_b200.nexmark import queries as nq
from dbsp_b200.parallel import Comm
from dbsp_b200.runtime import Runtime
from parity_util import assert_batches_equal, build_query, feed


def rand_batch(be, schema, n, domain, seed):
    rng = np.random.default_rng(seed)
    cols = [rng.integers(-domain, domain, n) if t == "i" else rng.integers(0, domain, n).astype(np.uint64) for t in schema.lanes]
    return be.batch_from_columns(schema, cols, rng.integers(-2, 3, n))


def merge_all(be, bs):
    acc = bs[0]
    for b in bs[1:]:
        acc = be.merge(acc, b)
    return acc


def run_ops(be, comm, rank, world):
    """dbsp_allreduce_max_u64 / dbsp_shard / dbsp_shard2 / dbsp_gather on random batches."""
    assert comm.native is be, "the CUDA backend must use the in-library exchange"
    for k in range(5):
        assert comm.allreduce_max(1000 * k + 7 * rank) == 1000 * k + 7 * (world - 1)
    cases = [(Schema("u", "u"), 50_000, 1 << 14), (Schema("ui", "uu"), 1_500_000, 1 << 22), (Schema("u"), 3, 10),
             (Schema("uu", "uiu"), 120_000, 300)]
    for ci, (s, n, dom) in enumerate(cases):
        for round_ in range(3):
            # rank-dependent sizes, an empty batch on the last rank every other round
            mine_n = 0 if (rank == world - 1 and round_ == 1) else n // (rank + 1)
            mine = rand_batch(be, s, mine_n, dom, 1000 * ci + 10 * round_ + rank)
            sh = comm.shard(be, mine)
            got = comm.gather(be, sh, 0)
            if rank == 0:
                want = merge_all(be, [rand_batch(be, s, 0 if (r == world - 1 and round_ == 1) else n // (r + 1), dom, 1000 * ci + 10 * round_ + r)
                                      for r in range(world)])
                assert_batches_equal(got, want, f"gather(shard) case {ci} round {round_}")
            else:
                assert len(got) == 0
            # every row of my shard hashes to me: re-sharding it is the identity
            assert_batches_equal(comm.shard(be, sh), sh, "shard is idempotent")
    # two streams in one round == two rounds
    a = rand_batch(be, Schema("u", "uu"), 200_000 // (rank + 1), 5000, 77 + rank)
    b = rand_batch(be, Schema("u", "u"), 30_000 * (rank + 1), 5000, 99 + rank)
    a2, b2 = comm.shard_many(be, [a, b])
    assert_batches_equal(a2, comm.shard(be, a), "shard2 stream 0")
    assert_batches_equal(b2, comm.shard(be, b), "shard2 stream 1")
    if rank == 0:
        print("OPS_OK", "NVLink bytes sent by rank 0:", comm.bytes_sent)


def run_query(be, comm, rank, world, query):
    rate = 100_000 if query == "q7" else 0
    n_events, step = (600_000, 200_000) if query != "q7" else (2_400_000, 300_000)
    c = RootCircuit(be, comm)
    inp, handles = nq.add_nexmark_input(c)
    out = nq.QUERIES[query](inp).gather(0).output()
    gen = NexmarkGenerator(first_event_rate=rate)
    got = []
    for s0 in range(0, n_events, step):
        t = gen.tables(s0, step)
        feed(handles, {k: [col[rank::world].copy() for col in v] for k, v in t.items()})
        c.step()
        if rank == 0:
            got.append(out.value)
    if rank == 0:
        c1, h1, o1 = build_query(be, query)
        total = 0
        for i, s0 in enumerate(range(0, n_events, step)):
            feed(h1, gen.tables(s0, step))
            c1.step()
            assert_batches_equal(got[i], o1.value, f"{query}: sharded != single at step {i}")
            total += len(o1.value)
        assert total > 0
        print("SHARDED_OK", query, total, "rows; NVLink bytes sent by rank 0:", comm.bytes_sent)


def main():
    what = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    be = Runtime(local)
    comm = Comm(dev)
    comm.attach(be)
    if what == "ops":
        run_ops(be, comm, rank, world)
    else:
        run_query(be, comm, rank, world, what)
    be.sync()
    dist.barrier()
    comm.detach()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
