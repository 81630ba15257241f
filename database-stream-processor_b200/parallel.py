"""Multi-GPU plumbing: one process per GPU.

Mirrors the reference's only distribution scheme: N identical circuit replicas, inputs
re-partitioned by key hash in front of every keyed operator (operator/communication/shard.rs:36-162), one
exchange round per sharded stream per step (exchange.rs:36-44), gather for verification (gather.rs:41-103),
watermark all-reduce (watermark.rs:53-70).  The traces never move: only delta batches cross NVLink.

Product path (CUDA backend): the exchange lives INSIDE the library — `dbsp_shard / dbsp_shard2 / dbsp_gather /
dbsp_allreduce_max_u64` (csrc/comm.cu): the partition kernel scatters straight into the peers' receive slots over
NVLink, counts travel through peer-mapped flag words, the receiver merges in place.  `torch.distributed` is used
ONCE, at `attach()`, to all-gather the 128-byte region descriptors (the bootstrap any host can do).

Host path (`native` is None): the same protocol through `torch.distributed` collectives on flat tensors — used by
the CPU test-suite (gloo + the oracle backend, tests/test_shard_gloo.py) and as the `DBSP_EXCHANGE=torch` escape
hatch.  Wire format of that path: every rank sends to peer p one contiguous int64 segment
[lane 0 rows | lane 1 rows | ... | weights] of the rows with hash(key) % P == p — already sorted, so the receiver
only merges (shard.rs:136-144); a P-element count all-to-all precedes the payload all-to-all.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from .zset import Backend, Batch


class Comm:
    def __init__(self, device: torch.device | None = None, group=None):
        assert dist.is_initialized(), "torch.distributed must be initialised (torchrun)"
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        self.device = device or torch.device("cpu")
        self._bytes_sent = 0   # payload bytes that left this rank through the host path
        self.native = None     # CUDA backend whose context carries the in-library exchange

    # -- in-library exchange (csrc/comm.cu) ---------------------------------
    def attach(self, be: Backend, slot_bytes: int = 0):
        """Create the context's receive region, all-gather the region descriptors, map the peers."""
        if be.name != "cuda" or os.environ.get("DBSP_EXCHANGE") == "torch":
            return
        blob = be.comm_create(self.rank, self.world_size, slot_bytes)
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(self.device)
        allb = torch.empty(self.world_size * len(blob), dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(allb, mine, group=self.group)
        be.comm_connect(bytes(allb.cpu().numpy().tobytes()))
        dist.barrier(group=self.group)   # every rank has mapped every region before the first store
        self.native = be

    def detach(self):
        if self.native is not None:
            dist.barrier(group=self.group)   # no peer still writes into a region that is about to go
            self.native.comm_destroy()
            self.native = None

    @property
    def bytes_sent(self):
        if self.native is not None:
            return self.native.comm_info()[2]
        return self._bytes_sent

    # -- scalars -----------------------------------------------------------
    def allreduce_max(self, x: int) -> int:
        if self.native is not None:
            return self.native.allreduce_max(x)
        t = torch.tensor([int(x)], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    def allreduce_sum(self, x: float) -> float:
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return float(t.item())

    def barrier(self):
        dist.barrier(group=self.group)

    # -- batches -----------------------------------------------------------
    def _exchange(self, be: Backend, parts: list[Batch]) -> list[Batch]:
        """All-to-all of one batch per peer; returns the batches received."""
        return self._exchange_many(be, [parts])[0]

    def _exchange_many(self, be: Backend, streams: list[list[Batch]]) -> list[list[Batch]]:
        """One exchange round for several streams at once: streams[i][p] is the
        batch of stream i destined to peer p.  One count all-to-all (one count
        per stream and peer) and ONE payload all-to-all whose per-peer segment
        is the concatenation of the streams' [lane.. | weights] blocks — the
        fixed latency of an exchange is paid once per operator, not once per
        input (a join shards both of its inputs in the same step)."""
        P, nS = self.world_size, len(streams)
        schemas = [parts[0].schema for parts in streams]
        L1 = [s.nl + 1 for s in schemas]
        counts = [[len(streams[i][p]) for i in range(nS)] for p in range(P)]      # [peer][stream]
        send_counts = torch.tensor(counts, dtype=torch.int64, device=self.device).reshape(-1)
        recv_counts = torch.empty(P * nS, dtype=torch.int64, device=self.device)
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        rc = recv_counts.reshape(P, nS).tolist()                                  # [peer][stream]
        segs = []
        be.sync()   # the partition kernels ran on the library's stream; torch reads the columns on its own
        for p in range(P):
            for i in range(nS):
                if counts[p][i]:
                    cols, w = be.batch_flat_tensors(streams[i][p], synced=True)
                    segs.extend(cols)
                    segs.append(w)
        send = torch.cat(segs) if segs else torch.empty(0, dtype=torch.int64, device=self.device)
        in_split = [sum(counts[p][i] * L1[i] for i in range(nS)) for p in range(P)]
        out_split = [sum(rc[q][i] * L1[i] for i in range(nS)) for q in range(P)]
        recv = torch.empty(sum(out_split), dtype=torch.int64, device=self.device)
        dist.all_to_all_single(recv, send, output_split_sizes=out_split, input_split_sizes=in_split, group=self.group)
        self._bytes_sent += sum(in_split[p] for p in range(P) if p != self.rank) * 8
        if recv.is_cuda:
            torch.cuda.current_stream(recv.device).synchronize()   # once, before the library adopts the segments
        # offsets of stream i's block inside peer q's segment
        base, off = [], 0
        for q in range(P):
            row = []
            for i in range(nS):
                row.append(off)
                off += rc[q][i] * L1[i]
            base.append(row)
        out = []
        for i in range(nS):
            schema, l1 = schemas[i], L1[i]
            ns = [rc[q][i] for q in range(P)]
            total = sum(ns)
            if 0 < total <= self.SORT_THRESHOLD and P > 2:
                # Small deltas: P-1 merges cost more in launches and read-backs than one
                # consolidation of the concatenated segments (Batch::from_tuples).
                cols = [torch.cat([recv[base[q][i] + l * ns[q]: base[q][i] + (l + 1) * ns[q]] for q in range(P)])
                        for l in range(schema.nl)]
                w = torch.cat([recv[base[q][i] + schema.nl * ns[q]: base[q][i] + l1 * ns[q]] for q in range(P)])
                out.append([be.batch_from_device_tensors(schema, cols, w)])
                continue
            got = []
            for q in range(P):
                n, o = ns[q], base[q][i]
                cols = [recv[o + l * n: o + (l + 1) * n] for l in range(schema.nl)]
                got.append(be.batch_from_flat_tensors(schema, cols, recv[o + schema.nl * n: o + l1 * n], synced=True))
            out.append(got)
        be.sync()   # the adoption copies run on the backend's stream: finish them before `recv` goes back to torch's allocator
        return out

    SORT_THRESHOLD = 1 << 20

    @staticmethod
    def _merge_all(be: Backend, batches: list[Batch]) -> Batch:
        """Receiver side of shard(): insert the P batches into a spine and
        consolidate (shard.rs:136-144) — here a balanced merge tree."""
        while len(batches) > 1:
            nxt = [be.merge(batches[i], batches[i + 1]) for i in range(0, len(batches) - 1, 2)]
            if len(batches) % 2:
                nxt.append(batches[-1])
            batches = nxt
        return batches[0]

    def shard(self, be: Backend, b: Batch) -> Batch:
        """shard (communication/shard.rs:106-162)."""
        if self.native is be:
            return be.shard(b)
        parts = be.shard_partition(b, self.world_size)
        return self._merge_all(be, self._exchange(be, parts))

    def shard_many(self, be: Backend, batches: list[Batch]) -> list[Batch]:
        """shard() of several streams in one exchange round (the two inputs of a join)."""
        if self.native is be and len(batches) == 2:
            return list(be.shard2(batches[0], batches[1]))
        streams = [be.shard_partition(b, self.world_size) for b in batches]
        return [self._merge_all(be, got) for got in self._exchange_many(be, streams)]

    def gather(self, be: Backend, b: Batch, root: int = 0) -> Batch:
        """gather (communication/gather.rs:41-103): everything to `root`,
        empty batches elsewhere."""
        if self.native is be:
            return be.gather(b, root)
        empty = be.batch_empty(b.schema)
        parts = [b if p == root else empty for p in range(self.world_size)]
        return self._merge_all(be, self._exchange(be, parts))
