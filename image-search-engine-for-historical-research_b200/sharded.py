"""Row-sharded exact search across the GPUs of one box (one process per GPU, torch.distributed).

The database is split into contiguous row ranges, rank g owning rows ``[offsets[g], offsets[g+1])``
(SURVEY.md section 8e).  Every rank scores ALL queries against its shard and produces an exact
local top-k with GLOBAL ids (``id_offset``); the only exchange step is one all-gather of the
``[nq, k]`` (score, id) lists -- 84 KB per rank at nq=70, k=100 -- followed by a ``G*k -> k`` merge
kernel (xs_merge_candidates).  Rescoring needs no communication: a shard holds the fp32 rows of
its own candidates.

On NVLink boxes the collective library can be taken off the data path altogether: ``PeerExchange`` gives every
rank a mailbox in its own HBM (CUDA IPC), a push kernel stores the packed list into all mailboxes and raises
a flag, and the merge kernel itself waits for the world's flags (xs_exchange_*; ``ShardedSearcher(...,
exchange=...)``).  torch.distributed then only carries the 64-byte handles and the set-up barriers.

The local searcher and the merge are injectable so that the sharding / id-offset / gather logic
is covered by world_size-2 gloo tests on CPU, where the test passes the oracle's searcher and merge
in; the product wiring (`CudaShard`) is CUDA only and there is no automatic fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat


def shard_bounds(n_rows: int, world: int):
    """Contiguous, near-equal row ranges: ``bounds[g] .. bounds[g+1]``."""
    base, rem = divmod(int(n_rows), int(world))
    b = [0]
    for g in range(world):
        b.append(b[-1] + base + (1 if g < rem else 0))
    return b


class ShardedSearcher:
    """Glue between a per-rank local searcher and the process group.

    ``local_search(queries, k) -> packed`` : one 1-D uint8 tensor per rank holding the local exact
    top-k as ``[ids int64 (nq*k) | sims f32 (nq*k)]`` with GLOBAL ids -- packed so that the exchange
    is ONE all-gather; ``merge(packed_all, world, nq, k) -> (ids [nq,k], sims [nq,k])``.
    """

    def __init__(self, local_search, merge, group=None, exchange=None, exchange_pipelined=False, lane_stream=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.local_search = local_search
        self.merge = merge
        self.exchange = exchange                    # PeerExchange: push + waiting merge instead of the all-gather
        # Measured on 8 B200s: the push wins on the blocking (latency) form, 0.231 vs 0.239 ms per 70-query search;
        # the pipelined form is 2 % faster with NCCL's all-gather on its own stream -- so that is its default.
        self.exchange_pipelined = bool(exchange_pipelined)
        self._inflight = [None, None]
        self._gathered = {}
        # lane_stream(slot) -> a CUDA stream owned by the local searcher's lane `slot` (CudaShard with two lanes):
        # the pipelined form then runs slot 0 and slot 1 searches on two streams, so that the selection / rescoring
        # tail of one batch overlaps the database scan of the next.  None = everything on the caller's stream.
        self.lane_stream = lane_stream
        # callbacks may take a result-slot argument (the CUDA shard does: two result buffers alternate)
        import inspect
        self._ls_slot = "slot" in inspect.signature(local_search).parameters
        self._mg_slot = "slot" in inspect.signature(merge).parameters

    def search(self, queries, k: int):
        """Blocking form: local exact top-k, one all-gather, merge.  Every rank returns the full answer."""
        return self.search_async(queries, k, blocking=True).result()

    def search_async(self, queries, k: int, blocking: bool = False):
        """Throughput form: enqueue the local search and START the all-gather, return a handle.  The
        collective runs on NCCL's own stream, so the next batch's scan overlaps it; ``handle.result()``
        makes the current stream wait for the gather and enqueues the merge.  Two result slots alternate,
        so at most two searches may be in flight per searcher."""
        import torch
        nq = int(queries.shape[0])
        slot = self._slot = (getattr(self, "_slot", 1) + 1) & 1
        use_peer = self.exchange is not None and (blocking or self.exchange_pipelined)
        if self.exchange is not None:
            if self._inflight[slot] is not None:
                self._inflight[slot].result()       # the slot's previous merge must be enqueued before its next push
            self.exchange.before_local_search(slot)
        lane = self.lane_stream(slot) if self.lane_stream is not None else None
        if lane is not None and blocking:
            torch.cuda.current_stream(lane.device).wait_stream(lane)   # an uncollected pipelined search of this lane goes first
            lane = None
        if lane is None:
            return self._enqueue(queries, k, nq, slot, use_peer, blocking, None)
        lane.wait_stream(torch.cuda.current_stream(lane.device))     # the queries (and this slot's previous merge) come first
        with torch.cuda.stream(lane):
            pending = self._enqueue(queries, k, nq, slot, use_peer, blocking, lane)
        return pending

    def _enqueue(self, queries, k, nq, slot, use_peer, blocking, lane):
        import torch
        packed = self.local_search(queries, k, slot) if self._ls_slot else self.local_search(queries, k)
        if use_peer:
            self.exchange.push(packed, slot, overlap=not blocking)
            self._inflight[slot] = _PendingPeer(self, nq, k, slot)
            return self._inflight[slot]
        if self.world == 1:
            done = None
            if lane is not None:
                done = torch.cuda.Event()
                done.record(lane)
            return _Pending(self, None, packed, nq, k, slot, done)
        key = (nq, k, packed.device, slot)
        if key not in self._gathered:
            self._gathered[key] = torch.empty((self.world * packed.numel(),), dtype=torch.uint8, device=packed.device)
        packed_all = self._gathered[key]
        work = self.dist.all_gather_into_tensor(packed_all, packed, group=self.group, async_op=True)
        return _Pending(self, work, packed_all, nq, k, slot)


class _Pending:
    def __init__(self, owner, work, buf, nq, k, slot, done=None):
        self.owner, self.work, self.buf, self.nq, self.k, self.slot, self.done = owner, work, buf, nq, k, slot, done

    def result(self):
        if self.work is None:
            if self.done is not None:                       # searched on a lane stream: the caller's stream waits for it
                import torch
                torch.cuda.current_stream().wait_event(self.done)
            return unpack(self.buf, self.nq, self.k)
        self.work.wait()                                    # stream-level wait, the host does not block
        o = self.owner
        return o.merge(self.buf, o.world, self.nq, self.k, self.slot) if o._mg_slot else o.merge(self.buf, o.world, self.nq, self.k)


class _PendingPeer:
    def __init__(self, owner, nq, k, slot):
        self.owner, self.nq, self.k, self.slot, self.out = owner, nq, k, slot, None

    def result(self):
        if self.out is None:
            self.out = self.owner.exchange.merge(self.nq, self.k, self.slot)
            if self.owner._inflight[self.slot] is self:
                self.owner._inflight[self.slot] = None
        return self.out


class PeerExchange:
    """Mailboxes + handshake for the peer-memory exchange (xs_exchange_*).  ``part_bytes`` bounds one rank's
    packed result (``packed_bytes(nq, k)`` of the largest search).  Collective: every rank of ``group``
    constructs it at the same point.  CUDA only."""

    def __init__(self, device: int, part_bytes: int, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group, self.device = torch, dist, group, int(device)
        self.lib = nat.load()
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.part_bytes = (int(part_bytes) + 15) // 16 * 16
        self._h = C.c_void_p()
        self._out = {}
        self._side, self._pushed, self._busy = None, None, [False, False]
        handle = (C.c_ubyte * 64)()
        err = None
        try:
            nat.check(self.lib.xs_exchange_create(self.device, self.world, self.rank, self.part_bytes, C.byref(self._h), handle),
                      "xs_exchange_create")
        except Exception as e:                      # keep walking through the collectives below: peers are waiting in them
            err = e
        if self.world > 1:
            dev = torch.device("cuda", self.device)
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
            everyone = torch.empty((self.world * 64,), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(everyone, mine, group=group)
            if err is None:
                try:
                    nat.check(self.lib.xs_exchange_connect(self._h, everyone.cpu().numpy().tobytes()), "xs_exchange_connect")
                except Exception as e:
                    err = e
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)   # also the barrier: every mailbox zeroed and mapped
            if int(ok.item()) == 0 and err is None:
                err = RuntimeError("peer exchange could not be set up on another rank")
        if err is not None:
            if self._h:
                self.lib.xs_exchange_destroy(self._h)
                self._h = C.c_void_p()
            raise RuntimeError(f"peer exchange unavailable: {err}") from err

    def push(self, packed, slot: int, overlap: bool = False):
        """Enqueue the push of ``packed`` (this rank's result of the search just enqueued).  ``overlap``: run it
        on a side stream behind an event, so that the caller's stream goes straight on to the next search --
        nothing has to join it, the merge kernel waits for the arrival flags (this rank's own included)."""
        torch = self.torch
        cur = torch.cuda.current_stream(self.device)
        if overlap:
            if self._side is None:
                self._side = torch.cuda.Stream(self.device)
                self._pushed = [torch.cuda.Event(), torch.cuda.Event()]
            self._side.wait_stream(cur)
            stream = self._side.cuda_stream
        else:
            stream = cur.cuda_stream
        nat.check(self.lib.xs_exchange_push(self._h, C.c_void_p(packed.data_ptr()), int(packed.numel()), int(slot),
                                            C.c_void_p(stream) if stream else None), "xs_exchange_push")
        if overlap:
            self._pushed[slot].record(self._side)
            self._busy[slot] = True

    def before_local_search(self, slot: int):
        """The result buffer of ``slot`` is about to be overwritten: a side-stream push still reading it goes first."""
        if self._busy[slot]:
            self.torch.cuda.current_stream(self.device).wait_event(self._pushed[slot])
            self._busy[slot] = False

    def merge(self, nq: int, k: int, slot: int):
        torch = self.torch
        key = (nq, k, slot)
        if key not in self._out:
            dev = torch.device("cuda", self.device)
            self._out[key] = (torch.empty((nq, k), dtype=torch.int64, device=dev), torch.empty((nq, k), dtype=torch.float32, device=dev))
        out_i, out_s = self._out[key]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(self.lib.xs_exchange_merge(self._h, int(slot), int(nq), int(k), C.c_void_p(out_i.data_ptr()),
                                             C.c_void_p(out_s.data_ptr()), C.c_void_p(stream) if stream else None),
                  "xs_exchange_merge")
        return out_i, out_s

    def close(self):
        """Collective: drains this rank's stream work, waits for every rank, then unmaps and frees."""
        if self._h:
            self.torch.cuda.synchronize(self.device)
            if self.world > 1:
                self.dist.barrier(self.group)
            self.lib.xs_exchange_destroy(self._h)
            self._h = C.c_void_p()


def packed_bytes(nq: int, k: int) -> int:
    """Bytes of one rank's packed result, padded to 16 so that every part stays aligned after the gather."""
    return (nq * k * 12 + 15) // 16 * 16


def unpack(packed, nq: int, k: int):
    """Views (no copy) of one rank's packed result: ``(ids int64 [nq,k], sims f32 [nq,k])``."""
    import torch
    ids = packed[: nq * k * 8].view(torch.int64).view(nq, k)
    sims = packed[nq * k * 8: nq * k * 12].view(torch.float32).view(nq, k)
    return ids, sims


class CudaShard:
    """The CUDA local searcher + merge for one rank: wraps an ExactIndex built with ``id_offset``."""

    def __init__(self, index, device: int, lanes: int = 1):
        import torch
        self.torch = torch
        self.index = index
        self.device = device
        self.lib = nat.load()
        self._out = {}
        # lanes = 2: result slot 1 searches through a clone of the index (own workspaces, same database arrays) and
        # both slots get a stream of their own -- hand `lane_stream` to ShardedSearcher to overlap consecutive batches
        self.lanes = [index] + [index.clone() for _ in range(max(0, int(lanes) - 1))]
        self._streams = [torch.cuda.Stream(device) for _ in self.lanes] if len(self.lanes) > 1 else []

    def lane_stream(self, slot: int):
        return self._streams[slot % len(self._streams)] if self._streams else None

    def close(self):
        for ix in self.lanes[1:]:
            ix.close()
        self.lanes = self.lanes[:1]
        self._streams = []

    def _buffers(self, nq, k, slot=0):
        torch = self.torch
        key = (nq, k, slot)
        if key not in self._out:
            dev = torch.device("cuda", self.device)
            packed = torch.empty((packed_bytes(nq, k),), dtype=torch.uint8, device=dev)
            ids, sims = unpack(packed, nq, k)
            self._out[key] = (ids, sims, torch.zeros((nq,), dtype=torch.int32, device=dev), packed)
        return self._out[key]

    def local_search(self, queries, k, slot=0):
        """``queries``: fp32 row-major torch tensor on this rank's device.  Returns the packed result
        (written into result slot ``slot``; two slots alternate while an all-gather is in flight)."""
        torch = self.torch
        nq = int(queries.shape[0])
        ids, sims, status, packed = self._buffers(nq, k, slot)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.lanes[slot % len(self.lanes)].search_device(queries.data_ptr(), nq, k, ids.data_ptr(), sims.data_ptr(),
                                                        status_ptr=status.data_ptr(), stream=stream)
        return packed

    def uncertified(self, nq, k) -> int:
        """Number of queries of the last local searches (both slots) the bf16 pass could not certify."""
        return int(self._buffers(nq, k, 0)[2].sum().item()) + int(self._buffers(nq, k, 1)[2].sum().item())

    def merge(self, packed_all, world, nq, k, slot=0):
        torch = self.torch
        key = ("m", nq, k, slot)
        if key not in self._out:
            self._out[key] = (torch.empty((nq, k), dtype=torch.int64, device=packed_all.device),
                              torch.empty((nq, k), dtype=torch.float32, device=packed_all.device))
        out_i, out_s = self._out[key]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        stride = packed_bytes(nq, k)
        base = packed_all.data_ptr()
        nat.check(self.lib.xs_merge_candidates_strided(self.device, C.c_void_p(base), C.c_void_p(base + nq * k * 8), stride, stride,
                                                       int(world), int(nq), int(k), C.c_void_p(out_i.data_ptr()),
                                                       C.c_void_p(out_s.data_ptr()), C.c_void_p(stream) if stream else None),
                  "xs_merge_candidates_strided")
        return out_i, out_s

    def merge_lists(self, ids_all, sims_all, k):
        """Merge of two dense ``[G, nq, k]`` arrays (xs_merge_candidates) -- kept for callers that gather separately."""
        torch = self.torch
        g, nq, kk = ids_all.shape
        out_i = torch.empty((nq, k), dtype=torch.int64, device=ids_all.device)
        out_s = torch.empty((nq, k), dtype=torch.float32, device=ids_all.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(self.lib.xs_merge_candidates(self.device, C.c_void_p(ids_all.data_ptr()), C.c_void_p(sims_all.data_ptr()),
                                               int(g), int(nq), int(k), C.c_void_p(out_i.data_ptr()),
                                               C.c_void_p(out_s.data_ptr()), C.c_void_p(stream) if stream else None),
                  "xs_merge_candidates")
        return out_i, out_s
