"""Multi-GPU partitioning of the k-NN path: one process per GPU (torch.distributed).

The reference never shards retrieval (faiss-gpu is pinned to device 0,
src/dataset/embedding_rag_infer_dataset.py:218), so both modes are new:

  * window sharding (BASELINE cfgs 2-4): windows are independent (own panel slice, own index:
    src/dataset/rag_train_dataset.py:52-136), so rank g owns a contiguous window range and
    searches it with NO data-path collective; results live in disjoint slices.
  * row-sharded panel (BASELINE cfg 5): rank g holds panel rows [g*N/G, (g+1)*N/G) of every
    window, searches its rows with GLOBAL ids (id_offset), then ONE exchange of the per-rank
    candidates and an on-device k-way merge on the (distance, id) order; result == unsharded search by
    construction.  On the GPU box the exchange is `PeerExchange` - candidate keys stored straight into the
    owning rank's memory over NVLink (CUDA IPC peer memory) and merged by the same kernel, no collective
    library on the data path - with an NCCL route (`all_to_all_single`) for GPUs without peer access; the
    stateless helpers below use an all-gather (every rank gets the full result) or an all-to-all (the result
    stays sharded by query: 1/G of the traffic and of the merge work), NCCL on GPUs and gloo in the CPU tests.

`search_fn` / `merge_fn` are injected so the plumbing (offsets, gather order, shapes) is testable
on CPU with world_size 2 over gloo; on the GPU they are the CUDA index's search and
`topk_merge` (snv_topk_merge) — there is no CPU implementation in this package.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) range of `n_items` for `rank` (first n % world ranks get one more)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(int(n_items), world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def window_owner(window: int, n_windows: int, world: int) -> int:
    """Rank that owns `window` under shard_range."""
    base, rem = divmod(int(n_windows), world)
    cut = rem * (base + 1)
    if window < cut:
        return window // (base + 1)
    return rem + (window - cut) // max(base, 1)


def search_row_sharded(search_fn: Callable, merge_fn: Callable, queries, k: int, row_lo: int,
                       group=None, world: Optional[int] = None, distribute: str = "all"):
    """Row-sharded exact k-NN for one window batch.

    search_fn(queries, k, id_offset) -> (D, I) tensors [.., nq, k] over THIS rank's rows, ids
    already global; merge_fn(D_parts [G, nq, k], I_parts [G, nq, k], k) -> (D, I).
    Every rank passes the same `queries`.

    distribute="all": one all-gather of the per-rank (D, I); every rank merges and returns the full result.
    distribute="scatter": one all-to-all instead - rank r receives, from every rank, only the rows of the
    flattened query range shard_range(n_queries, G, r), merges those and returns (q_lo, q_hi, D, I) for its
    slice: 1/G of the bytes on the wire and 1/G of the merge work per rank (the merged result stays
    query-sharded, which is what a data-parallel consumer wants)."""
    import torch
    import torch.distributed as dist

    D, I = search_fn(queries, k, row_lo)
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if distribute not in ("all", "scatter"):
        raise ValueError("distribute must be 'all' or 'scatter'")
    if world == 1:
        if distribute == "scatter":  # one shard: its top-k already is the result
            D2, I2 = D.reshape(-1, k), I.reshape(-1, k)
            return 0, D2.shape[0], D2, I2
        return merge_fn(D.unsqueeze(0), I.unsqueeze(0), k)
    D = D.contiguous()
    I = I.contiguous()
    if distribute == "scatter":
        rank = dist.get_rank(group)
        D2, I2 = D.reshape(-1, k), I.reshape(-1, k)
        nqt = D2.shape[0]
        cuts = [shard_range(nqt, world, r) for r in range(world)]
        lo, hi = cuts[rank]
        Dg = torch.empty((world, hi - lo, k), dtype=D.dtype, device=D.device)
        Ig = torch.empty((world, hi - lo, k), dtype=I.dtype, device=I.device)
        dist.all_to_all(list(Dg.unbind(0)), [D2[a:b] for a, b in cuts], group=group)
        dist.all_to_all(list(Ig.unbind(0)), [I2[a:b] for a, b in cuts], group=group)
        Dm, Im = merge_fn(Dg, Ig, k)
        return lo, hi, Dm, Im
    Dg = torch.empty((world,) + tuple(D.shape), dtype=D.dtype, device=D.device)
    Ig = torch.empty((world,) + tuple(I.shape), dtype=I.dtype, device=I.device)
    # list-of-views form: same call on nccl (GPU box) and gloo (CPU tests)
    dist.all_gather(list(Dg.unbind(0)), D, group=group)
    dist.all_gather(list(Ig.unbind(0)), I, group=group)
    return merge_fn(Dg, Ig, k)


def search_window_sharded(search_fn: Callable, n_windows: int, world: int, rank: int):
    """Runs search_fn(w_lo, w_hi) on this rank's window range; returns (w_lo, w_hi, result).
    No collective: callers that need every window on every rank gather the results themselves."""
    lo, hi = shard_range(n_windows, world, rank)
    return lo, hi, (search_fn(lo, hi) if hi > lo else None)


class PeerExchange:
    """NVLink exchange of a row-sharded search, fused with its merge into one kernel launch per batch (snv_peer_*, see
    include/snvknn.h): every rank stores its candidate keys straight into the owning rank's receive slot through CUDA IPC
    peer memory, raises a flag there, waits for the flags of all sources and merges - no collective library on the data
    path.  torch.distributed only carries the one-off all-gather of the 64-byte IPC handles (and the agreement that every
    rank could map every peer).

    `PeerExchange.connect(device, slot_bytes, group)` is collective: every rank of `group` calls it with the same
    `slot_bytes`.  It returns None (on every rank) when some pair of GPUs has no peer access, so the caller can take the
    NCCL route instead."""

    def __init__(self, handle, device: int, rank: int, world: int, slot_bytes: int):
        self._h = handle
        self.device, self.rank, self.world, self.slot_bytes = int(device), int(rank), int(world), int(slot_bytes)

    @classmethod
    def _create(cls, device: int, rank: int, world: int, slot_bytes: int):
        import ctypes

        from . import _lib as L

        h = ctypes.c_void_p()
        raw = ctypes.create_string_buffer(64)
        L.check(L.lib().snv_peer_create(int(device), int(rank), int(world), int(slot_bytes), ctypes.byref(h), raw), "snv_peer_create")
        return cls(h, device, rank, world, slot_bytes), raw.raw

    @classmethod
    def connect(cls, device: int, slot_bytes: int, group=None) -> Optional["PeerExchange"]:
        import torch
        import torch.distributed as dist

        from . import _lib as L

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        me, raw = cls._create(device, rank, world, slot_bytes)
        # the handles travel as a CPU-side object gather (64 bytes per rank, once)
        handles = [None] * world
        dist.all_gather_object(handles, raw, group=group)
        rc = L.lib().snv_peer_open(me._h, b"".join(handles))
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int32, device=f"cuda:{device}")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            me.close()
            return None
        return me

    @classmethod
    def connect_local(cls, device: int, world: int, slot_bytes: int):
        """`world` exchange objects in THIS process on one device, wired by pointer (single-GPU test hook: run their
        exchanges on different streams)."""
        import ctypes

        from . import _lib as L

        peers = [cls._create(device, r, world, slot_bytes)[0] for r in range(world)]
        arr = (ctypes.c_void_p * world)(*[p._h for p in peers])
        for p in peers:
            L.check(L.lib().snv_peer_open_local(p._h, arr), "snv_peer_open_local")
        return peers

    def exchange(self, D, I, k_out: Optional[int] = None, out=None):
        """D int32 / I int64 [nw, nq, k] (this rank's candidates, global ids) -> (D, I) [nw, nq / world, k_out]: this
        rank's queries merged over all ranks.  Runs on the current CUDA stream; all calls of one object must use the
        same stream."""
        import torch

        from . import _lib as L
        from .index import _current_stream

        nw, nq, k = (int(x) for x in D.shape)
        k_out = int(k_out or k)
        if D.dtype != torch.int32 or I.dtype != torch.int64 or not D.is_contiguous() or not I.is_contiguous():
            raise ValueError("PeerExchange.exchange: D must be contiguous int32, I contiguous int64")
        if nq % self.world:
            raise ValueError("PeerExchange.exchange: the queries of a window must split evenly over the ranks")
        if nw * nq * k * 8 > self.slot_bytes:
            raise ValueError("PeerExchange.exchange: batch larger than the exchange slot")
        qg = nq // self.world
        if out is None:
            Do = torch.empty((nw, qg, k_out), dtype=torch.int32, device=D.device)
            Io = torch.empty((nw, qg, k_out), dtype=torch.int64, device=D.device)
        else:
            Do, Io = out
        L.check(L.lib().snv_peer_exchange(self._h, D.data_ptr(), I.data_ptr(), nw, nq, k, k_out, Do.data_ptr(), Io.data_ptr(),
                                          _current_stream(self.device)), "snv_peer_exchange")
        return Do, Io

    def push(self, D, I) -> int:
        """Phase 1 alone (pack + push + flag) on the current stream; returns the batch's epoch for `merge`."""
        import ctypes

        import torch

        from . import _lib as L
        from .index import _current_stream

        nw, nq, k = (int(x) for x in D.shape)
        if D.dtype != torch.int32 or I.dtype != torch.int64 or not D.is_contiguous() or not I.is_contiguous():
            raise ValueError("PeerExchange.push: D must be contiguous int32, I contiguous int64")
        if nq % self.world:
            raise ValueError("PeerExchange.push: the queries of a window must split evenly over the ranks")
        epoch = ctypes.c_uint64(0)
        L.check(L.lib().snv_peer_push(self._h, D.data_ptr(), I.data_ptr(), nw, nq, k, ctypes.byref(epoch), _current_stream(self.device)),
                "snv_peer_push")
        return int(epoch.value)

    def merge(self, epoch: int, nw: int, nq: int, k: int, out, k_out: Optional[int] = None) -> None:
        """Phase 2 alone (wait for every source's flag of `epoch` + merge into out=(D, I) [nw, nq / world, k_out])."""
        from . import _lib as L
        from .index import _current_stream

        Do, Io = out
        L.check(L.lib().snv_peer_merge(self._h, int(epoch), int(nw), int(nq), int(k), int(k_out or k), Do.data_ptr(), Io.data_ptr(),
                                       _current_stream(self.device)), "snv_peer_merge")

    def close(self) -> None:
        if self._h is not None:
            from . import _lib as L

            L.lib().snv_peer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RowShardedSearch:
    """Row-sharded exact k-NN over a multi-window index, pipelined: the panel rows of every window are split over the
    ranks (this rank holds rows [row_lo, row_lo + ntotal)); the merged result comes back sharded by QUERY - rank r
    gets queries shard_range(nq, G, r) of every window, D / I [nw, nq_r, k] - which is what a data-parallel consumer
    wants and moves 1/G of the bytes an all-gather would.

    Per call the windows are cut into `chunks` groups.  For each group: local scan with global ids on the caller's
    stream, then the exchange of the packed candidates (one int64 key = distance << 40 | id per neighbour instead of
    separate int32 / int64 arrays: 8 instead of 12 bytes on the wire) and the on-device k-way merge.
    transport="peer" (default): `PeerExchange` on the caller's stream - one fused kernel for sync=True, push / merge
    split around the next scan for sync=False (see `_search_peer`).  transport="nccl" (or no peer access): ONE
    all_to_all_single + merge on a side stream, overlapping the next group's scan.  The result equals the unsharded
    search by construction of the (distance, id) total order.

    `search_fn(queries [nw_c, nq, ..], k, w0) -> (D, I) [nw_c, nq, k]` (ids global) and `merge_fn(D [G, n, k], I [G, n, k], k)`
    are injectable so that the plumbing runs on CPU tensors over gloo in the tests; by default they are the CUDA
    index's search (`index.search(q, k, w0=w0, id_offset=row_lo)`) and `topk_merge`."""

    _ID_BITS = 40

    def __init__(self, index, row_lo: int, world: Optional[int] = None, group=None, merge_fn: Optional[Callable] = None,
                 search_fn: Optional[Callable] = None, chunks: Optional[int] = None, transport: Optional[str] = None):
        import os

        import torch.distributed as dist

        self.index = index
        self.row_lo = int(row_lo)
        self.group = group
        self.world = int(world) if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = dist.get_rank(group) if (self.world > 1 and dist.is_initialized()) else 0
        self._native = merge_fn is None
        if merge_fn is None:
            from .index import topk_merge as merge_fn  # noqa: PLC0415
        self.merge_fn = merge_fn
        self.search_fn = search_fn or (lambda q, k, w0: index.search(q, k, w0=w0, id_offset=self.row_lo))
        self.chunks = chunks
        self._side = None
        self._pending = None
        self._last = ""
        # exchange transport on CUDA: "peer" = the fused NVLink kernel (PeerExchange), "nccl" = pack + all_to_all_single +
        # merge; default peer, falling back to nccl when the GPUs have no peer access
        self.transport = (transport or os.environ.get("SNV_EXCHANGE", "peer")).lower()
        if self.transport not in ("peer", "nccl"):
            raise ValueError("transport must be 'peer' or 'nccl'")
        self._peer = None
        self._peer_failed = False
        self._scan_out = search_fn is None   # the default search can write into caller-owned buffers
        self._rings = {}
        self._merges = []   # NVLink route: (epoch, windows, queries, k, out) of pushed batches whose merge is not queued yet

    def describe(self) -> str:
        return self._last or "not run yet"

    @classmethod
    def pack_keys(cls, D, I):
        """(D int32 >= 0, I int64 global id or -1) -> one int64 key per neighbour; missing entries -> the largest key"""
        import torch

        key = (D.to(torch.int64) << cls._ID_BITS) | (I & ((1 << cls._ID_BITS) - 1))
        return torch.where(I < 0, torch.full_like(key, torch.iinfo(torch.int64).max), key)

    @classmethod
    def unpack_keys(cls, key):
        import torch

        missing = key == torch.iinfo(torch.int64).max
        D = torch.where(missing, torch.full_like(key, 0x7FFFFFFF), key >> cls._ID_BITS).to(torch.int32)
        I = torch.where(missing, torch.full_like(key, -1), key & ((1 << cls._ID_BITS) - 1))
        return D, I

    def _scan_buffer(self, device, wc: int, nq: int, k: int, ring: int = 2):
        """`ring` persistent (D, I) scan-result buffers per batch shape, used in turn."""
        import torch

        shape = (wc, nq, k)
        nring = ring
        ring = self._rings.setdefault((shape, nring), {"bufs": [], "n": 0})
        i = ring["n"] % nring
        ring["n"] += 1
        if len(ring["bufs"]) <= i:
            ring["bufs"].append({"D": torch.empty(shape, dtype=torch.int32, device=device),
                                 "I": torch.empty(shape, dtype=torch.int64, device=device), "busy": None})
        return ring["bufs"][i]

    def _peer_for(self, device, nw: int, nq: int, k: int):
        """The NVLink exchange object, (re)connected collectively when a batch needs a larger slot.  Every rank sees the
        same shapes, so every rank takes the same decision."""
        if self.transport != "peer" or self._peer_failed:
            return None
        need = nw * nq * k * 8
        if self._peer is None or self._peer.slot_bytes < need:
            import torch

            if self._peer is not None:
                self._flush_merges()
                torch.cuda.synchronize(device)
                self._peer.close()
            self._peer = PeerExchange.connect(device.index, need, self.group)
            if self._peer is None:
                self._peer_failed = True
        return self._peer

    def _flush_merges(self) -> None:
        """Queue the merges of the batches pushed so far (current stream)."""
        for (epoch, wc, nq, k, out) in self._merges:
            self._peer.merge(epoch, wc, nq, k, out)
        self._merges = []

    def _search_peer(self, queries, k: int, sync: bool, bounds, q_lo: int, q_hi: int):
        """NVLink route, everything on the caller's stream.  sync=True: scan, then the fused exchange (one kernel).
        sync=False: scan i, merge of batch i - 1, push of batch i - the merge of a batch is queued behind the NEXT scan, when
        every peer's push is a whole scan old, so none of its blocks waits for a late peer while holding an SM; `wait()`
        queues the last merge."""
        import torch

        nw, nq = int(queries.shape[0]), int(queries.shape[1])
        G = self.world
        dev = queries.device
        outD = torch.empty((nw, q_hi - q_lo, k), dtype=torch.int32, device=dev)
        outI = torch.empty((nw, q_hi - q_lo, k), dtype=torch.int64, device=dev)
        for (w0, w1) in bounds:
            wc = w1 - w0
            buf = self._scan_buffer(dev, wc, nq, k, ring=1)   # stream order alone protects it: the push reads it before the next scan
            D, I = self.index.search(queries[w0:w1], k, w0=w0, id_offset=self.row_lo, out=(buf["D"], buf["I"]))
            self._flush_merges()
            if sync and w1 == nw:
                self._peer.exchange(D, I, k, out=(outD[w0:w1], outI[w0:w1]))
            else:
                self._merges.append((self._peer.push(D, I), wc, nq, k, (outD[w0:w1], outI[w0:w1])))
        self._pending = None
        self._last = (f"{len(bounds)} window group(s) per call; NVLink peer memory (CUDA IPC), no collective library on the data path: "
                      f"int64 (distance << 40 | id) keys stored straight into the owning rank's receive slot + flag "
                      f"({nq // G} queries x k to each of {G} ranks), then wait + k-way merge - "
                      + ("one fused kernel" if sync else "push after scan i, merge queued behind scan i + 1 (no block waits for a late peer)"))
        return q_lo, q_hi, outD, outI

    def wait(self) -> None:
        """Make the caller's current stream wait for the exchange + merge of the last `search(..., sync=False)`."""
        import torch

        if self._merges:
            self._flush_merges()
        if self._pending is not None:
            torch.cuda.current_stream(self._pending[1]).wait_event(self._pending[0])
            self._pending = None

    def search(self, queries, k: int, sync: bool = True):
        """queries [nw, nq, ..] (the same on every rank) -> (q_lo, q_hi, D [nw, q_hi - q_lo, k], I [..]): this rank's
        queries of every window, merged over all row shards.

        sync=False (CUDA): the exchange + merge stay on the side stream and the call returns without making the caller's
        stream wait for them, so the NEXT batch's scan overlaps this batch's NVLink exchange (software pipelining across
        batches).  The returned D / I must not be read before `wait()` (or any later `search(..., sync=True)`)."""
        import torch
        import torch.distributed as dist

        nw, nq = int(queries.shape[0]), int(queries.shape[1])
        G = self.world
        if G == 1:
            D, I = self.search_fn(queries, k, 0)
            self._last = "single shard: no exchange"
            return 0, nq, D, I
        cuts = [shard_range(nq, G, r) for r in range(G)]
        q_lo, q_hi = cuts[self.rank]
        even = nq % G == 0
        # window groups: pipelining pays once a group still fills the machine for several rounds (small groups lose more to
        # the scan's tail than the overlap wins)
        n_chunks = self.chunks if self.chunks else (2 if nw >= 16 else 1)
        n_chunks = max(1, min(int(n_chunks), nw))
        bounds = [shard_range(nw, n_chunks, c) for c in range(n_chunks)]
        on_gpu = queries.is_cuda
        outD = outI = None
        main = side = None
        if (on_gpu and even and self._native and self._scan_out and int(k) <= 32
                and self._peer_for(queries.device, nw, nq, int(k)) is not None):
            return self._search_peer(queries, int(k), sync, bounds, q_lo, q_hi)
        if on_gpu:
            main = torch.cuda.current_stream(queries.device)
            if self._side is None:
                # high priority: the exchange of batch i goes in front of the scan CTAs of batch i + 1 when both are ready
                self._side = torch.cuda.Stream(queries.device, priority=-1)
            side = self._side
            side.wait_stream(main)
        for (w0, w1) in bounds:
            wc = w1 - w0
            ring = None
            if on_gpu and self._scan_out:
                # the scan writes into one of two persistent (D, I) buffers; a buffer is reused only after the exchange
                # that read it has finished (bounded pipeline depth, no allocator traffic per batch)
                ring = self._scan_buffer(queries.device, wc, nq, k)
                if ring["busy"] is not None:
                    main.wait_event(ring["busy"])
                D, I = self.index.search(queries[w0:w1], k, w0=w0, id_offset=self.row_lo, out=(ring["D"], ring["I"]))
            else:
                D, I = self.search_fn(queries[w0:w1], k, w0)
            if outD is None:
                outD = torch.empty((nw, q_hi - q_lo, k), dtype=D.dtype, device=D.device)
                outI = torch.empty((nw, q_hi - q_lo, k), dtype=torch.int64, device=D.device)
                if on_gpu:
                    outD.record_stream(side)
                    outI.record_stream(side)

            def exchange_and_merge(D=D, I=I, w0=w0, w1=w1, wc=wc):
                if on_gpu and even and D.dtype == torch.int32 and self._native:
                    # NCCL path: one pack kernel, one collective, one merge kernel writing the result slice
                    from . import _lib as L
                    from .index import _current_stream

                    qg = nq // G
                    dev = D.device.index
                    send = torch.empty((G, wc, qg, k), dtype=torch.int64, device=D.device)
                    L.check(L.lib().snv_exchange_pack(dev, D.data_ptr(), I.data_ptr(), wc, nq, int(k), G, send.data_ptr(),
                                                      _current_stream(dev)), "snv_exchange_pack")
                    recv = torch.empty_like(send)
                    dist.all_to_all_single(recv, send, group=self.group)
                    L.check(L.lib().snv_exchange_merge(dev, recv.data_ptr(), G, wc * qg, int(k), int(k), outD[w0:w1].data_ptr(),
                                                       outI[w0:w1].data_ptr(), _current_stream(dev)), "snv_exchange_merge")
                    return
                packed = D.dtype == torch.int32
                parts = [self.pack_keys(D, I)] if packed else [D, I]
                got = []
                for t in parts:
                    if even:
                        qg = nq // G
                        send = t.reshape(wc, G, qg, k).permute(1, 0, 2, 3).contiguous()   # [dest][window][query][k]
                        recv = torch.empty_like(send)
                        dist.all_to_all_single(recv, send, group=self.group)
                        got.append(recv.reshape(G, wc * qg, k))
                    else:
                        send = torch.cat([t[:, a:b].reshape(-1, k) for a, b in cuts], dim=0)
                        recv = torch.empty((G * wc * (q_hi - q_lo), k), dtype=t.dtype, device=t.device)
                        dist.all_to_all_single(recv, send, output_split_sizes=[wc * (q_hi - q_lo)] * G,
                                               input_split_sizes=[wc * (b - a) for a, b in cuts], group=self.group)
                        got.append(recv.reshape(G, wc * (q_hi - q_lo), k))
                Dg, Ig = self.unpack_keys(got[0]) if packed else (got[0], got[1])
                Dm, Im = self.merge_fn(Dg.contiguous(), Ig.contiguous(), k)
                outD[w0:w1] = Dm.reshape(wc, q_hi - q_lo, k)
                outI[w0:w1] = Im.reshape(wc, q_hi - q_lo, k)

            if on_gpu:
                ev = main.record_event()
                with torch.cuda.stream(side):
                    side.wait_event(ev)
                    exchange_and_merge()
                    if ring is not None:
                        ring["busy"] = side.record_event()
                if ring is None:
                    D.record_stream(side)
                    I.record_stream(side)
            else:
                exchange_and_merge()
        if on_gpu:
            if sync:
                main.wait_stream(side)
                self._pending = None
            else:
                self._pending = (side.record_event(), queries.device)
        self._last = (f"{n_chunks} window group(s) per call; per group one all_to_all_single of packed int64 (distance << 40 | id) keys "
                      f"({nq // G if even else q_hi - q_lo} queries x k from each of {G} ranks) + k-way merge on a side stream, overlapping the next group's scan")
        return q_lo, q_hi, outD, outI
