"""Host logic of the tensor-core Hamming engine, checked without a GPU.

`snv_debug_hamming_plan` runs the planner of csrc/hamming_tc.cu (engine choice, row splits, tail split) and enumerates the
work items with the decoder the kernel uses.  Whatever the shape, the items must tile every (window, query tile, panel
tile) exactly once, pieces of one (window, query tile) must carry distinct piece numbers below their piece count, and
the partial keys of the pieces must fit the workspace the planner asked for."""
import os

import numpy as np
import pytest

BM = 128  # queries per tile (TMEM lanes)


@pytest.fixture(scope="module")
def L():
    from rag_snvbert_b200 import _lib

    _lib.lib()
    return _lib


@pytest.fixture
def force_engine():
    old = os.environ.get("SNV_HAMMING_ENGINE")

    def setter(name):
        if name is None:
            os.environ.pop("SNV_HAMMING_ENGINE", None)
        else:
            os.environ["SNV_HAMMING_ENGINE"] = name

    yield setter
    if old is None:
        os.environ.pop("SNV_HAMMING_ENGINE", None)
    else:
        os.environ["SNV_HAMMING_ENGINE"] = old


def _bn(engine):
    return 160 if engine == 5 else (240 if engine >= 3 else 256)


def _check_cover(plan, items, nw, nq, n, k):
    eng = plan["engine"]
    assert eng in (1, 2, 3, 4, 5)
    bn = _bn(eng)
    qtiles = -(-nq // BM)
    n_tiles = -(-n // bn)
    assert plan["qtiles"] == qtiles and plan["n_tiles"] == n_tiles
    assert plan["kt"] == (8 if k <= 8 else 32) and plan["kt"] >= k
    # every id of a piece must fit the id field of the 32-bit selection key
    assert items[:, 3].max() * bn <= (1 << plan["idx_bits"])
    assert (1 << (32 - plan["idx_bits"])) > 0 and plan["idx_bits"] >= 1
    cover = np.zeros((nw, qtiles, n_tiles), np.int32)
    pieces = {}
    for w, qt, t0, nt, piece, npieces, row_base, slot in items.tolist():
        assert 0 <= w < nw and nt >= 1 and 0 <= t0 and t0 + nt <= n_tiles
        assert 0 <= piece < npieces
        if qt >= qtiles:
            assert eng >= 4 and qt == qtiles and qtiles % 2 == 1  # the idle half of the last CTA pair
            continue
        cover[w, qt, t0 : t0 + nt] += 1
        key = (w, qt)
        pieces.setdefault(key, []).append((piece, npieces, row_base))
        if npieces > 1:
            # partial keys [row - row_base][pieces][kt] stay inside the workspace
            last_row = w * nq + min(nq, (qt + 1) * BM) - 1 - row_base
            assert last_row >= 0
            assert ((last_row + 1) * npieces) * plan["kt"] * 8 <= plan["workspace_kib"] * 1024
    assert (cover == 1).all(), "panel tiles covered %d..%d times" % (cover.min(), cover.max())
    for key, ps in pieces.items():
        ids = sorted(p[0] for p in ps)
        assert ids == list(range(len(ps))) and all(p[1] == len(ps) for p in ps), (key, ps)
        assert len({p[2] for p in ps}) == 1
    units = 74 if eng >= 4 else 148
    assert items[:, 7].max() < units


@pytest.mark.parametrize("eng", ["tc", "tc4", "tc4x2", "tc4x2ta"])
@pytest.mark.parametrize(
    "shape",
    [
        (1000, 2000, 5008, 1030, 8),  # BASELINE cfg 2
        (8, 10000, 200000, 1030, 32),  # cfg 5, one step of 8 windows
        (8, 10000, 25000, 1030, 32),  # cfg 5 row shard on 8 GPUs
        (1, 64, 40000, 256, 8),  # one item: rows are split
        (10, 2000, 5008, 1030, 8),  # 80 pair items on 74 pairs: tail split
        (1, 1, 1, 1, 1),
        (3, 129, 241, 33, 9),
        (149, 128, 512, 64, 8),
        (75, 256, 480, 100, 32),
    ],
)
def test_items_tile_the_search(L, force_engine, eng, shape):
    force_engine(eng)
    nw, nq, n, d, k = shape
    plan, items = L.debug_hamming_plan(nw, nq, n, d, k)
    assert plan["engine"] == {"tc": 1, "tc4": 3, "tc4x2": 4, "tc4x2ta": 5}[eng]
    _check_cover(plan, items, nw, nq, n, k)


def test_random_shapes(L, force_engine):
    rng = np.random.default_rng(2024)
    seen_split = seen_tail = 0
    for case in range(300):
        force_engine(["tc", "tc4", "tc4x2", "tc4x2ta"][case % 4])
        nw = int(rng.integers(1, 200))
        nq = int(rng.integers(1, 1200))
        n = int(rng.choice([rng.integers(1, 600), rng.integers(600, 30000), rng.integers(30000, 400000)]))
        d = int(rng.integers(1, 4095))
        k = int(rng.integers(1, 33))
        plan, items = L.debug_hamming_plan(nw, nq, n, d, k)
        _check_cover(plan, items, nw, nq, n, k)
        seen_split += plan["nsplit"] > 1
        seen_tail += plan["tail_split"] > 1
    assert seen_split > 10 and seen_tail > 10  # both mechanisms exercised


def test_cfg_shapes_pick_expected_plans(L, force_engine):
    """the plans the committed measurements were taken with"""
    force_engine(None)
    plan, items = L.debug_hamming_plan(1000, 2000, 5008, 1030, 8)
    assert plan["engine"] == 5 and plan["nsplit"] == 1 and plan["kblocks"] == 5 and plan["n_tiles"] == 32
    # a chunk of cfg 2 as api.cu cuts it: items a multiple of the 74 SM pairs -> no splits at all
    plan, items = L.debug_hamming_plan(37, 2000, 5008, 1030, 8)
    assert plan["nsplit"] == 1 and plan["tail_split"] == 0 and len(items) == 2 * 37 * 8
    # cfg 5 step on one GPU (k = 32)
    plan, items = L.debug_hamming_plan(8, 10000, 200000, 1030, 32)
    assert plan["engine"] == 5 and plan["kt"] == 32
    # unsplit items + tail split (measured 6.5 ms against 7.2 ms for two row splits, round 2)
    assert plan["nsplit"] == 1 and plan["tail_items"] == 24 and plan["tail_split"] == 3
    # its 8-GPU row shard: 8 windows x 40 tile pairs = 320 items = 4 rounds + 24 -> the 24 trailing items are cut in 3
    plan, items = L.debug_hamming_plan(8, 10000, 25000, 1030, 32)
    assert plan["engine"] == 5 and plan["kt"] == 32 and plan["nsplit"] == 1
    assert plan["tail_items"] == 24 and plan["tail_split"] == 3


def test_auto_engine_by_shape(L, force_engine):
    force_engine(None)
    eng = lambda *a: L.debug_hamming_plan(*a)[0]["engine"]
    assert eng(1, 300, 2000, 1030, 8) == 5
    assert eng(1, 300, 2000, 1500, 8) == 4  # wider than the tensor-memory operand budget: shared-memory query ring
    assert eng(1, 100, 2000, 1030, 8) == 3  # one query tile: nothing for the second CTA of a pair
    assert eng(1, 2, 5008, 1030, 8) == 0  # the reference's training-time call: popcount scan
    assert eng(1, 64, 600, 5000, 8) == 0  # d >= 4096
    assert eng(1, 64, 300, 1030, 8) == 0  # panel smaller than a useful tile
    assert eng(1, 64, 5008, 1030, 33) == 0  # k > 32
    force_engine("popc")
    assert eng(1000, 2000, 5008, 1030, 8) == 0
    # the TMEM-operand engine holds at most 5 k-blocks of a query tile (20 MMA slots of 64 sites, one of them the
    # column-index block of the list epilogue: 1216 sites); wider windows run engine 4
    force_engine("tc4x2ta")
    assert eng(10, 2000, 5008, 1216, 8) == 5
    assert eng(10, 2000, 5008, 1217, 8) == 4


def test_bad_arguments(L):
    with pytest.raises((ValueError, RuntimeError)):
        L.debug_hamming_plan(1, 1, 1, 0, 1)
    with pytest.raises((ValueError, RuntimeError)):
        L.debug_hamming_plan(1, 1, 1, 10, 0)


def _check_bounds(b, nw):
    assert b[0] == 0 and b[-1] == nw and (np.diff(b) > 0).all()


def test_chunk_bounds_cfg2(L, force_engine):
    """cfg 2 from host buffers: full chunks of 37 windows = 4 items per SM pair, a short first and last chunk"""
    force_engine(None)
    b = L.debug_hamming_chunks(1000, 2000, 5008, 1030, 8, host_io=True)
    _check_bounds(b, 1000)
    sizes = np.diff(b)
    assert sizes[0] == 4 and sizes[-1] == 4 and (sizes[1:-2] == 37).all() and sizes[-2] <= 37 + 4
    assert 20 <= len(sizes) <= 32
    # device-resident searches are one chunk
    assert L.debug_hamming_chunks(1000, 2000, 5008, 1030, 8, host_io=False).tolist() == [0, 1000]


def test_chunk_bounds_random(L, force_engine):
    rng = np.random.default_rng(7)
    for case in range(300):
        force_engine([None, "popc", "tc4", "tc4x2", "tc", "tc4x2ta"][case % 6])
        nw = int(rng.integers(1, 3000))
        nq = int(rng.integers(1, 5000))
        n = int(rng.integers(1, 20000))
        d = int(rng.integers(1, 4095))
        k = int(rng.integers(1, 33))
        b = L.debug_hamming_chunks(nw, nq, n, d, k, host_io=True)
        _check_bounds(b, nw)
        assert len(b) - 1 <= 64 + 3, (nw, nq, len(b))  # the pipeline stays a few dozen chunks whatever the shape
        plan, _ = L.debug_hamming_plan(nw, nq, n, d, k, cap=0)
        sizes = np.diff(b)
        if plan["engine"] and len(sizes) >= 8:
            # full-size chunks hold a whole number of items per SM (pair)
            per_w = -(-plan["qtiles"] // 2) if plan["engine"] >= 4 else plan["qtiles"]
            units = 74 if plan["engine"] >= 4 else 148
            full = sizes[1:-2]
            assert (full == full[0]).all()
            assert (full[0] * per_w) % units == 0 or full[0] * per_w >= 2 * units
