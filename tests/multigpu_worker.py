"""torchrun worker: sharded Nexmark query over NCCL vs a single-GPU run of the same events."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import torch
import torch.distributed as dist

import dbsp_b200
from dbsp_b200 import RootCircuit
from dbsp_b200.nexmark import NexmarkGenerator
from dbsp_b200.nexmark import queries as nq
from dbsp_b200.parallel import Comm
from dbsp_b200.runtime import Runtime
from parity_util import build_query, feed


def main():
    query = sys.argv[1]
    n_events, step = (600_000, 200_000) if query != "q7" else (1_200_000, 300_000)
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    be = Runtime(local)
    comm = Comm(dev)
    c = RootCircuit(be, comm)
    inp, handles = nq.add_nexmark_input(c)
    out = nq.QUERIES[query](inp).gather(0).output()
    gen = NexmarkGenerator()
    got = []
    for s0 in range(0, n_events, step):
        t = gen.tables(s0, step)
        feed(handles, {k: [col[rank::world].copy() for col in v] for k, v in t.items()})
        c.step()
        if rank == 0:
            got.append(out.value.rows())
    if rank == 0:
        be1 = Runtime(local)
        c1, h1, o1 = build_query(be1, query)
        for i, s0 in enumerate(range(0, n_events, step)):
            feed(h1, gen.tables(s0, step))
            c1.step()
            assert o1.value.rows() == got[i], f"{query}: sharded != single at step {i}"
        print("SHARDED_OK", query, sum(len(g) for g in got), "rows; NVLink bytes sent by rank 0:", comm.bytes_sent)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
