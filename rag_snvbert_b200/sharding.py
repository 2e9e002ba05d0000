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
