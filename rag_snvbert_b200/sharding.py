"""Multi-GPU partitioning of the k-NN path: one process per GPU (torch.distributed).

The reference never shards retrieval (faiss-gpu is pinned to device 0,
src/dataset/embedding_rag_infer_dataset.py:218), so both modes are new:

  * window sharding (BASELINE cfgs 2-4): windows are independent (own panel slice, own index:
    src/dataset/rag_train_dataset.py:52-136), so rank g owns a contiguous window range and
    searches it with NO data-path collective; results live in disjoint slices.
  * row-sharded panel (BASELINE cfg 5): rank g holds panel rows [g*N/G, (g+1)*N/G) of every
    window, searches its rows with GLOBAL ids (id_offset), then ONE exchange of the per-rank
    (D, I) [Q, k] (NCCL over NVLink on the GPU box, gloo in the CPU tests) - an all-gather (every rank
    gets the full result) or an all-to-all (the result stays sharded by query: 1/G of the traffic and of
    the merge work) - and an on-device k-way merge on the (distance, id) order.  Result == unsharded
    search by construction.

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


class RowShardedSearch:
    """Row-sharded exact k-NN over a multi-window index, pipelined: the panel rows of every window are split over the
    ranks (this rank holds rows [row_lo, row_lo + ntotal)); the merged result comes back sharded by QUERY - rank r
    gets queries shard_range(nq, G, r) of every window, D / I [nw, nq_r, k] - which is what a data-parallel consumer
    wants and moves 1/G of the bytes an all-gather would.

    Per call the windows are cut into `chunks` groups.  For each group: local scan with global ids on the caller's
    stream; then, on a side stream, ONE all_to_all_single of the packed candidates (one int64 key = distance << 40 | id
    per neighbour instead of separate int32 / int64 arrays: 8 instead of 12 bytes on the wire, one collective
    instead of two) and the on-device k-way merge - overlapping the next group's scan.  The result equals the
    unsharded search by construction of the (distance, id) total order.

    `search_fn(queries [nw_c, nq, ..], k, w0) -> (D, I) [nw_c, nq, k]` (ids global) and `merge_fn(D [G, n, k], I [G, n, k], k)`
    are injectable so that the plumbing runs on CPU tensors over gloo in the tests; by default they are the CUDA
    index's search (`index.search(q, k, w0=w0, id_offset=row_lo)`) and `topk_merge`."""

    _ID_BITS = 40

    def __init__(self, index, row_lo: int, world: Optional[int] = None, group=None, merge_fn: Optional[Callable] = None,
                 search_fn: Optional[Callable] = None, chunks: Optional[int] = None):
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

    def wait(self) -> None:
        """Make the caller's current stream wait for the exchange + merge of the last `search(..., sync=False)`."""
        import torch

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
        if on_gpu:
            main = torch.cuda.current_stream(queries.device)
            if self._side is None:
                self._side = torch.cuda.Stream(queries.device)
            side = self._side
            side.wait_stream(main)
        for (w0, w1) in bounds:
            D, I = self.search_fn(queries[w0:w1], k, w0)
            wc = w1 - w0
            if outD is None:
                outD = torch.empty((nw, q_hi - q_lo, k), dtype=D.dtype, device=D.device)
                outI = torch.empty((nw, q_hi - q_lo, k), dtype=torch.int64, device=D.device)

            def exchange_and_merge(D=D, I=I, w0=w0, w1=w1, wc=wc):
                if on_gpu and even and D.dtype == torch.int32 and self._native:
                    # CUDA fast path: one pack kernel, one collective, one merge kernel writing the result slice
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
                outD.record_stream(main)
                outI.record_stream(main)
        self._last = (f"{n_chunks} window group(s) per call; per group one all_to_all_single of packed int64 (distance << 40 | id) keys "
                      f"({nq // G if even else q_hi - q_lo} queries x k from each of {G} ranks) + k-way merge on a side stream, overlapping the next group's scan")
        return q_lo, q_hi, outD, outI
