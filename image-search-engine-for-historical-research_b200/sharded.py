"""Row-sharded exact search across the GPUs of one box (one process per GPU, torch.distributed).

The database is split into contiguous row ranges, rank g owning rows ``[offsets[g], offsets[g+1])``
(SURVEY.md section 8e).  Every rank scores ALL queries against its shard and produces an exact
local top-k with GLOBAL ids (``id_offset``); the only exchange step moves the ``[nq, k]`` (score, id)
lists plus one certificate word per query -- 84 KB per rank at nq=70, k=100 -- followed by a
``G*k -> k`` merge kernel.  Rescoring needs no communication: a shard holds the fp32 rows of its own
candidates.

Two exchange paths:

* ``PeerExchange`` (NVLink boxes): every rank owns a mailbox in its own HBM (CUDA IPC); the LAST kernel of the
  local search stores each query's results straight into all mailboxes and releases a per-query flag, and the
  merge kernel waits for the world's flags (xs_search_dev_push / xs_exchange_merge).  No collective, no packed
  local result, no extra launch on the sending side.  torch.distributed only carries the 64-byte handles and the
  set-up barriers.
* NCCL: ``all_gather_into_tensor`` of the packed result on NCCL's own stream + xs_merge_candidates_strided.

Certificates travel with the lists: the merged status word of a query is the OR over the shards, so every rank
sees the same verdict and the uncertified queries (crowded or duplicate-heavy data) are re-run collectively on
the exact fp32 path before a result is handed out -- on the blocking and on the pipelined form.

The local searcher and the merge are injectable so that the sharding / id-offset / gather / re-run logic
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


def packed_bytes(nq: int, k: int) -> int:
    """Bytes of one rank's packed result ``ids int64 [nq*k] | sims f32 [nq*k] | status int32 [nq]``, padded to 16
    so that every part stays aligned after the gather (== xs_exchange_part_bytes)."""
    return (nq * k * 12 + nq * 4 + 15) // 16 * 16


def unpack(packed, nq: int, k: int):
    """Views (no copy) of one rank's packed result: ``(ids int64 [nq,k], sims f32 [nq,k], status int32 [nq])``."""
    import torch
    ids = packed[: nq * k * 8].view(torch.int64).view(nq, k)
    sims = packed[nq * k * 8: nq * k * 12].view(torch.float32).view(nq, k)
    status = packed[nq * k * 12: nq * k * 12 + nq * 4].view(torch.int32)
    return ids, sims, status


class ShardedSearcher:
    """Glue between a per-rank local searcher and the process group.

    ``local_search(queries, k[, slot][, exact]) -> packed`` : one 1-D uint8 tensor per rank holding the local exact
    top-k (``packed_bytes`` layout) with GLOBAL ids -- packed so that the exchange is ONE all-gather;
    ``merge(packed_all, world, nq, k[, slot]) -> (ids [nq,k], sims [nq,k], status [nq])``.
    ``exchange``: a ``PeerExchange``; then ``local_push(queries, k, exchange, slot)`` replaces ``local_search``
    on the hot path (the search's emit step writes into the peers' mailboxes).
    """

    def __init__(self, local_search, merge, group=None, exchange=None, local_push=None, lane_stream=None, check=True):
        import inspect
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.local_search = local_search
        self.merge = merge
        self.exchange = exchange
        self.local_push = local_push
        if exchange is not None and local_push is None:
            raise ValueError("a peer exchange needs the local_push callback (CudaShard.local_push)")
        self.check = bool(check)                    # read the merged certificate words and re-run what they flag
        self.n_rerun = 0                            # queries re-run on the exact path so far
        self._inflight = [None, None]
        self._gathered = {}
        # lane_stream(slot) -> a CUDA stream owned by the local searcher's lane `slot` (CudaShard with two lanes):
        # the pipelined form then runs slot 0 and slot 1 searches on two streams, so that the selection / rescoring
        # tail of one batch overlaps the database scan of the next.  None = everything on the caller's stream.
        self.lane_stream = lane_stream
        p = inspect.signature(local_search).parameters
        self._ls_slot, self._ls_exact = "slot" in p, "exact" in p
        self._mg_slot = "slot" in inspect.signature(merge).parameters

    # -- public ---------------------------------------------------------------------------------
    def search(self, queries, k: int):
        """Blocking form: local exact top-k, one exchange, merge, certificate check.  Every rank returns the full answer."""
        return self.search_async(queries, k, blocking=True).result()

    def search_async(self, queries, k: int, blocking: bool = False):
        """Throughput form: enqueue the local search and START the exchange, return a handle; ``handle.result()``
        enqueues the merge (behind the gather) and checks the certificates.  Two result slots alternate, so at most
        two searches may be in flight per searcher."""
        import torch
        nq = int(queries.shape[0])
        slot = self._slot = (getattr(self, "_slot", 1) + 1) & 1
        if self._inflight[slot] is not None:
            self._inflight[slot].result()           # the slot's previous merge must be enqueued before it is reused
        lane = self.lane_stream(slot) if self.lane_stream is not None else None
        if lane is not None and blocking:
            torch.cuda.current_stream(lane.device).wait_stream(lane)   # an uncollected pipelined search of this lane goes first
            lane = None
        if lane is None:
            pending = self._enqueue(queries, k, nq, slot, None)
        else:
            lane.wait_stream(torch.cuda.current_stream(lane.device))   # the queries (and this slot's previous merge) come first
            with torch.cuda.stream(lane):
                pending = self._enqueue(queries, k, nq, slot, lane)
        self._inflight[slot] = pending
        return pending

    # -- internals ------------------------------------------------------------------------------
    def _local(self, queries, k, slot, exact=False):
        kw = {}
        if self._ls_slot:
            kw["slot"] = slot
        if exact:
            if not self._ls_exact:
                raise RuntimeError("the local searcher cannot re-run uncertified queries (no `exact` argument)")
            kw["exact"] = True
        return self.local_search(queries, k, **kw)

    def _merge(self, packed_all, world, nq, k, slot):
        out = self.merge(packed_all, world, nq, k, slot) if self._mg_slot else self.merge(packed_all, world, nq, k)
        return out if len(out) == 3 else (out[0], out[1], None)

    def _enqueue(self, queries, k, nq, slot, lane):
        import torch
        if self.exchange is not None:
            # `local_push` may hand back the merged outputs itself (xs_search_dev_exchange: the merge rides in the search's
            # last kernel); None = only the sending end ran and result() enqueues the merge
            merged = self.local_push(queries, k, self.exchange, slot)
            return _Pending(self, "peer", None, merged, queries, nq, k, slot, lane)
        packed = self._local(queries, k, slot)
        if self.world == 1:
            return _Pending(self, "local", None, packed, queries, nq, k, slot, lane)
        key = (nq, k, packed.device, slot)
        if key not in self._gathered:
            self._gathered[key] = torch.empty((self.world * packed.numel(),), dtype=torch.uint8, device=packed.device)
        packed_all = self._gathered[key]
        work = self.dist.all_gather_into_tensor(packed_all, packed, group=self.group, async_op=True)
        return _Pending(self, "gather", work, packed_all, queries, nq, k, slot, lane)

    def _gather_blocking(self, packed):
        import torch
        if self.world == 1:
            return packed
        out = torch.empty((self.world * packed.numel(),), dtype=torch.uint8, device=packed.device)
        self.dist.all_gather_into_tensor(out, packed, group=self.group)
        return out

    def _rerun(self, queries, k, rows, ids, sims):
        """Collective: every rank saw the same merged certificate words, so every rank re-runs the same queries on
        its shard's exact fp32 path, gathers and merges them (plain blocking collectives: this is the rare path)."""
        import torch
        sel = queries[torch.as_tensor(rows, device=queries.device)]
        packed = self._local(sel, k, slot=0, exact=True)
        packed_all = self._gather_blocking(packed.clone())
        i2, s2, _ = self._merge(packed_all, self.world, len(rows), k, 0)
        idx = torch.as_tensor(rows, device=ids.device)
        ids[idx] = i2
        sims[idx] = s2
        self.n_rerun += len(rows)


class _Pending:
    def __init__(self, owner, kind, work, buf, queries, nq, k, slot, lane):
        self.owner, self.kind, self.work, self.buf = owner, kind, work, buf
        self.queries, self.nq, self.k, self.slot, self.lane = queries, nq, k, slot, lane
        self.out = None

    def result(self):
        if self.out is not None:
            return self.out
        import torch
        o = self.owner
        if self.kind == "peer":
            # the merge goes onto the stream of the push: the only flags it can ever wait for are other GPUs'
            if self.buf is not None:                        # already merged by the search itself
                ids, sims, status = self.buf
                if self.lane is not None:
                    torch.cuda.current_stream(self.lane.device).wait_stream(self.lane)
            elif self.lane is not None:
                with torch.cuda.stream(self.lane):
                    ids, sims, status = o.exchange.merge(self.nq, self.k, self.slot)
                torch.cuda.current_stream(self.lane.device).wait_stream(self.lane)
            else:
                ids, sims, status = o.exchange.merge(self.nq, self.k, self.slot)
        elif self.kind == "local":
            if self.lane is not None:
                torch.cuda.current_stream(self.lane.device).wait_stream(self.lane)
            ids, sims, status = unpack(self.buf, self.nq, self.k)
        else:
            self.work.wait()                                # stream-level wait, the host does not block
            ids, sims, status = o._merge(self.buf, o.world, self.nq, self.k, self.slot)
        if o.check and status is not None:
            bad = np.flatnonzero(status.cpu().numpy())       # one small device->host read: the caller is about to consume the result
            if bad.size:
                o._rerun(self.queries, self.k, bad.tolist(), ids, sims)
        if o._inflight[self.slot] is self:
            o._inflight[self.slot] = None
        self.out = (ids, sims)
        return self.out


class PeerExchange:
    """Mailboxes + handshake for the peer-memory exchange (xs_exchange_*).  Sized for searches of at most
    ``max_queries`` queries and ``k_max`` results.  Collective: every rank of ``group`` constructs it at the same
    point (the handles travel through the group's backend: NCCL with CUDA tensors, gloo with CPU tensors -- the
    latter also lets two ranks share ONE device, which NCCL refuses).  CUDA only."""

    def __init__(self, device: int, max_queries: int, k_max: int, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group, self.device = torch, dist, group, int(device)
        self.lib = nat.load()
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.max_queries, self.k_max = int(max_queries), int(k_max)
        self._h = C.c_void_p()
        self._out = {}
        handle = (C.c_ubyte * 64)()
        err = None
        try:
            nat.check(self.lib.xs_exchange_create(self.device, self.world, self.rank, self.max_queries, self.k_max, C.byref(self._h), handle),
                      "xs_exchange_create")
        except Exception as e:                      # keep walking through the collectives below: peers are waiting in them
            err = e
        if self.world > 1:
            on_gpu = dist.get_backend(group) == "nccl"
            dev = torch.device("cuda", self.device) if on_gpu else torch.device("cpu")
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
            everyone = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(everyone, mine, group=group)
            if err is None:
                try:
                    nat.check(self.lib.xs_exchange_connect(self._h, torch.cat(everyone).cpu().numpy().tobytes()), "xs_exchange_connect")
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

    @property
    def handle(self):
        return self._h

    def push(self, packed, nq: int, k: int, slot: int):
        """Sending end as a kernel of its own: ``packed`` (``packed_bytes`` layout, this rank's result) goes into every
        mailbox on the current stream.  The hot path does not need it (see ``CudaShard.local_push``)."""
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        nat.check(self.lib.xs_exchange_push(self._h, C.c_void_p(packed.data_ptr()), int(nq), int(k), int(slot),
                                            C.c_void_p(stream) if stream else None), "xs_exchange_push")

    def outputs(self, nq: int, k: int, slot: int):
        """The merged-result buffers of ``slot`` (reused from step to step)."""
        torch = self.torch
        key = (nq, k, slot)
        if key not in self._out:
            dev = torch.device("cuda", self.device)
            self._out[key] = (torch.empty((nq, k), dtype=torch.int64, device=dev), torch.empty((nq, k), dtype=torch.float32, device=dev),
                              torch.empty((nq,), dtype=torch.int32, device=dev))
        return self._out[key]

    def merge(self, nq: int, k: int, slot: int):
        torch = self.torch
        out_i, out_s, out_st = self.outputs(nq, k, slot)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        nat.check(self.lib.xs_exchange_merge(self._h, int(slot), int(nq), int(k), C.c_void_p(out_i.data_ptr()),
                                             C.c_void_p(out_s.data_ptr()), C.c_void_p(out_st.data_ptr()),
                                             C.c_void_p(stream) if stream else None), "xs_exchange_merge")
        return out_i, out_s, out_st

    def close(self):
        """Collective: drains this rank's stream work, waits for every rank, then unmaps and frees."""
        if self._h:
            self.torch.cuda.synchronize(self.device)
            if self.world > 1:
                self.dist.barrier(self.group)
            self.lib.xs_exchange_destroy(self._h)
            self._h = C.c_void_p()


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
            packed = torch.zeros((packed_bytes(nq, k),), dtype=torch.uint8, device=dev)
            self._out[key] = unpack(packed, nq, k) + (packed,)
        return self._out[key]

    def local_search(self, queries, k, slot=0, exact=False):
        """``queries``: fp32 row-major torch tensor on this rank's device.  Returns the packed result (ids, sims and
        certificate words; written into result slot ``slot`` -- two slots alternate while an all-gather is in flight).
        ``exact``: the fp32 path, for queries the coarse pass could not certify."""
        torch = self.torch
        nq = int(queries.shape[0])
        queries = queries.contiguous()
        ids, sims, status, packed = self._buffers(nq, k, ("x", slot) if exact else slot)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        ix = self.lanes[slot % len(self.lanes)]
        if exact:
            ix.set_param("force_path", nat.PATH_EXACT)
        try:
            ix.search_device(queries.data_ptr(), nq, k, ids.data_ptr(), sims.data_ptr(), status_ptr=status.data_ptr(), stream=stream)
        finally:
            if exact:
                ix.set_param("force_path", nat.PATH_AUTO)
        return packed

    def local_push(self, queries, k, exchange, slot=0):
        """The local search with the exchange's sending end fused into its last kernel (xs_search_dev_push)."""
        nq = int(queries.shape[0])
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        self.lanes[slot % len(self.lanes)].search_device_push(queries.data_ptr(), nq, k, exchange.handle, slot, stream=stream)

    def local_exchange(self, queries, k, exchange, slot=0):
        """Search + push + merge in one call (xs_search_dev_exchange): returns the merged ``(ids, sims, status)``."""
        nq = int(queries.shape[0])
        out_i, out_s, out_st = exchange.outputs(nq, k, slot)
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        self.lanes[slot % len(self.lanes)].search_device_exchange(queries.data_ptr(), nq, k, exchange.handle, slot, out_i.data_ptr(),
                                                                  out_s.data_ptr(), out_st.data_ptr(), stream=stream)
        return out_i, out_s, out_st

    def merge(self, packed_all, world, nq, k, slot=0):
        torch = self.torch
        key = ("m", nq, k, slot)
        if key not in self._out:
            self._out[key] = (torch.empty((nq, k), dtype=torch.int64, device=packed_all.device),
                              torch.empty((nq, k), dtype=torch.float32, device=packed_all.device),
                              torch.empty((nq,), dtype=torch.int32, device=packed_all.device))
        out_i, out_s, out_st = self._out[key]
        stream = torch.cuda.current_stream(self.device).cuda_stream
        stride = packed_bytes(nq, k)
        base = packed_all.data_ptr()
        nat.check(self.lib.xs_merge_candidates_strided(self.device, C.c_void_p(base), C.c_void_p(base + nq * k * 8), C.c_void_p(base + nq * k * 12),
                                                       stride, stride, stride, int(world), int(nq), int(k), C.c_void_p(out_i.data_ptr()),
                                                       C.c_void_p(out_s.data_ptr()), C.c_void_p(out_st.data_ptr()),
                                                       C.c_void_p(stream) if stream else None),
                  "xs_merge_candidates_strided")
        return out_i, out_s, out_st

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


class PipelinedSearcher:
    """The hot loop of the sharded search behind ONE C call per step (xs_pipeline_*): local search with the exchange's
    sending end fused in, merge and the certificate read-back are enqueued natively on one of two lane streams;
    ``result()`` waits for the step's event.  Same interface as ``ShardedSearcher`` (``search`` / ``search_async`` ->
    handle with ``result()``), same certificate handling: queries the merged words flag are re-run collectively on the
    exact path through ``fallback`` (a ``ShardedSearcher`` over the same shard).  At eight GPUs a step is ~0.12 ms of
    device time -- less than the host work of driving it from Python call by call."""

    def __init__(self, index, device: int, nq_max: int, k_max: int, lanes: int = 2, exchange: "PeerExchange | None" = None,
                 fallback: "ShardedSearcher | None" = None):
        import torch
        self.torch, self.device, self.lib = torch, int(device), nat.load()
        self.index, self.exchange, self.fallback = index, exchange, fallback
        self.world = exchange.world if exchange is not None else 1
        self.nq_max, self.k_max = int(nq_max), int(k_max)
        self.n_rerun = 0
        self._h = C.c_void_p()
        nat.check(self.lib.xs_pipeline_create(index._h, exchange.handle if exchange is not None else None, self.nq_max, self.k_max,
                                              int(lanes), C.byref(self._h)), "xs_pipeline_create")
        self._pending = [None, None]
        self._next = 0
        self._views = {}
        self._flag_buf = (C.c_int32 * self.nq_max)()

    def close(self):
        if self._h:
            self.lib.xs_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def search(self, queries, k: int):
        return self.search_async(queries, k).result()

    def search_async(self, queries, k: int):
        nq = int(queries.shape[0])
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        nxt = self._next
        if self._pending[nxt] is not None:          # the slot about to be reused has to be collected first: do it for the caller
            self._pending[nxt].result()
        slot = C.c_int(0)
        nat.check(self.lib.xs_pipeline_submit(self._h, C.c_void_p(queries.data_ptr()), nq, int(k), C.c_void_p(stream) if stream else None,
                                              C.byref(slot)), "xs_pipeline_submit")
        h = _PipelineHandle(self, queries, nq, int(k), slot.value)
        self._pending[slot.value] = h
        self._next = slot.value ^ 1
        return h


class _PipelineHandle:
    def __init__(self, owner, queries, nq, k, slot):
        self.owner, self.queries, self.nq, self.k, self.slot = owner, queries, nq, k, slot
        self.collected, self.out = False, None

    def result(self):
        if self.out is not None:
            return self.out
        o = self.owner
        torch = o.torch
        stream = torch.cuda.current_stream(o.device).cuda_stream
        pi, ps, nf = C.c_void_p(), C.c_void_p(), C.c_int64(0)
        flagged = o._flag_buf
        nat.check(o.lib.xs_pipeline_collect(o._h, self.slot, C.c_void_p(stream) if stream else None, C.byref(pi), C.byref(ps), C.byref(nf), flagged),
                  "xs_pipeline_collect")
        self.collected = True
        key = (pi.value, ps.value, self.nq, self.k)
        views = o._views.get(key)
        if views is None:                           # the slot's result buffers do not move: wrap them once
            views = o._views[key] = (_wrap_device(torch, pi.value, (self.nq, self.k), torch.int64, o.device),
                                     _wrap_device(torch, ps.value, (self.nq, self.k), torch.float32, o.device))
        ids, sims = views
        if nf.value:
            if o.fallback is None:
                raise RuntimeError(f"{nf.value} queries could not be certified and no exact fallback was configured")
            ids, sims = ids.clone(), sims.clone()
            o.fallback._rerun(self.queries, self.k, list(flagged[: nf.value]), ids, sims)
            o.n_rerun += int(nf.value)
        if o._pending[self.slot] is self:
            o._pending[self.slot] = None
        self.out = (ids, sims)
        return self.out


def _wrap_device(torch, ptr: int, shape, dtype, device: int):
    """A torch view of device memory owned by the library (valid until the slot's next submit)."""
    n = 1
    for x in shape:
        n *= int(x)
    itemsize = torch.empty((), dtype=dtype).element_size()

    class _Mem:                                   # __cuda_array_interface__ carrier
        pass
    m = _Mem()
    m.__cuda_array_interface__ = {"shape": (n,), "typestr": {torch.int64: "<i8", torch.float32: "<f4", torch.int32: "<i4"}[dtype],
                                  "data": (int(ptr), False), "version": 3, "strides": None}
    del itemsize
    return torch.as_tensor(m, device=torch.device("cuda", device)).view(*shape)


def make_searcher(index, device: int, lanes: int = 1, exchange: "PeerExchange | None" = None, group=None, check=True,
                  pipeline: "tuple | None" = None):
    """The product wiring in one call: ``(CudaShard, searcher)`` for this rank's index.  ``pipeline=(nq_max, k_max)`` selects
    the native two-slot pipeline (``PipelinedSearcher``: one C call per step) -- the hot loop of the multi-GPU bench; the
    default is the Python-driven ``ShardedSearcher`` (also the NCCL path and the exact re-run path)."""
    shard = CudaShard(index, device, lanes=1 if pipeline else lanes)
    searcher = ShardedSearcher(shard.local_search, shard.merge, group=group, exchange=exchange,
                               local_push=shard.local_push if exchange is not None else None,
                               lane_stream=shard.lane_stream if (lanes > 1 and not pipeline) else None, check=check)
    if pipeline:
        return shard, PipelinedSearcher(index, device, pipeline[0], pipeline[1], lanes=lanes, exchange=exchange, fallback=searcher)
    return shard, searcher


def self_knn_rowsharded(searcher: ShardedSearcher, rows_local, bounds, k: int, rank: int, group=None, block: int = 8192):
    """The N x N self-kNN graph (``self.knn.search(self.features, n_trunc)``, src/utils/diffusion.py:67) over a database
    that is ROW-SHARDED across the ranks -- the variant for databases beyond one GPU.

    ``rows_local``: this rank's rows, fp32 ``[n_local, D]`` torch tensor on its device (what its index was built from);
    ``bounds``: ``shard_bounds(N, world)``.  The owners take turns: a block of ``block`` query rows is broadcast from
    its owner, every rank searches it against its own shard, the per-shard lists meet through ``searcher`` (peer
    exchange or all-gather + merge) and the owner keeps the merged block.  Two blocks are in flight, so the exchange
    of one overlaps the scan of the next.  Returns ``(sims [n_local, k], ids [n_local, k])`` for the LOCAL rows, a
    row's own id first (diffusion.py:108 relies on it)."""
    import torch
    dist = searcher.dist
    world = searcher.world
    n_local = int(rows_local.shape[0])
    dev = rows_local.device
    out_i = torch.empty((n_local, k), dtype=torch.int64, device=dev)
    out_s = torch.empty((n_local, k), dtype=torch.float32, device=dev)
    bufs = [torch.empty((block, rows_local.shape[1]), dtype=torch.float32, device=dev) for _ in range(3)] if world > 1 else []
    pending = []                                         # (handle, owner, b0, c)

    def collect(entry):
        h, owner, b0, c = entry
        ids, sims = h.result()
        if owner != rank:
            return
        lo = bounds[rank] + b0
        own = torch.arange(lo, lo + c, device=dev, dtype=torch.int64)
        ids, sims = ids.clone(), sims.clone()
        wrong = torch.nonzero(ids[:, 0] != own).flatten()
        if wrong.numel():                                # duplicated rows: an equal-score neighbour with a lower id came first
            for r in wrong.tolist():
                pos = torch.nonzero(ids[r] == own[r]).flatten()
                p = int(pos[0]) if pos.numel() else k - 1
                self_sim = sims[r, p] if pos.numel() else torch.dot(rows_local[b0 + r], rows_local[b0 + r])
                ids[r, 1:p + 1] = ids[r, 0:p].clone()
                sims[r, 1:p + 1] = sims[r, 0:p].clone()
                ids[r, 0], sims[r, 0] = own[r], self_sim
        out_i[b0:b0 + c] = ids
        out_s[b0:b0 + c] = sims

    step = 0
    for owner in range(world):
        n_o = bounds[owner + 1] - bounds[owner]
        for b0 in range(0, n_o, block):
            c = min(block, n_o - b0)
            if world > 1:
                blk = bufs[step % 3][:c]
                if owner == rank:
                    blk.copy_(rows_local[b0:b0 + c])
                dist.broadcast(blk, src=owner if group is None else dist.get_global_rank(group, owner), group=group)
            else:
                blk = rows_local[b0:b0 + c]
            if len(pending) == 2:                        # the buffer about to be reused two steps from now is free again
                collect(pending.pop(0))
            pending.append((searcher.search_async(blk, k), owner, b0, c))
            step += 1
    while pending:
        collect(pending.pop(0))
    return out_s, out_i
