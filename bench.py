#!/usr/bin/env python
"""bench.py — BASELINE.json metric on BASELINE.json configs[1] (chr21-scale sweep), plus short secondary lines.

A "step" is one pass of the hot path over one batch of synthetic input: 1,000 genomic windows x 5,008 reference
haplotypes x 1,030 sites (bit-packed), 2,000 query haplotypes per window, exact top-8 by (Hamming distance, id).

  python bench.py [--gpus N] [--steps K] [--warmup W]      # the driver's call
  python bench.py --impl reference ...                      # the reference's CPU algorithm on the host cores
  python bench.py --workload cfg1|cfg2|cfg3|cfg4|cfg5 ...   # one configuration as the headline

With N GPUs (one process per GPU, torch.distributed over NCCL) the default is what BASELINE.json words: the SAME
1,000 windows window-sharded over the N GPUs ("scaling": "strong", no data-path collective); `weak` in the same line
is the other reading (every rank its own 1,000 windows).  `secondary` carries one short measured line each for cfg 5
(200,000-row panel row-sharded over the N GPUs, NCCL exchange of the per-shard top-32 + on-device merge), cfg 4
(float L2 on tcgen05) and cfg 1 (faiss.IndexFlatL2 add + search through the drop-in module), each with its own
roofline / e2e / cpu_baseline.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ref-haplotypes scanned/s at k=8 (window-queries/s x panel rows)"
UNIT = "ref-haplotypes/s"
ENGINE_NAMES = ["popcount", "tcgen05-fp8", "tcgen05-fp8-hbm", "tcgen05-fp4", "tcgen05-fp4-cta-pair", "tcgen05-fp4-cta-pair-tmemA"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=1000)
    ap.add_argument("--refs", type=int, default=5008)
    ap.add_argument("--sites", type=int, default=1030)
    ap.add_argument("--queries", type=int, default=2000)
    ap.add_argument("-k", type=int, default=8)
    ap.add_argument("--masked", action="store_true", help="cfg 3: per-query observed-site masks")
    ap.add_argument("--workload", default="all", choices=["all", "cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="all (default) = cfg2 headline + secondary cfg5 / cfg4 / cfg1; cfgN = that configuration alone")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="cfg2/3 with N GPUs: strong = the job's windows split over the ranks (BASELINE wording), weak = every rank its own")
    ap.add_argument("--dim", type=int, default=256, help="cfg4 embedding dimension")
    ap.add_argument("--precision", default="tf32x3", choices=["tf32", "tf32x3"], help="cfg4 cross-term precision")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU baseline sample budget (headline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    return ap.parse_args()


def workload_name(a, which=None):
    which = which or a.workload
    if which == "cfg5":
        return (f"cfg5 biobank-scale: {a.windows} of 500 windows x {a.refs} ref haplotypes x {a.sites} sites, "
                f"{a.queries} queries/window, k={a.k}")
    if which == "cfg4":
        return (f"cfg4 embedding-RAG retrieval: {a.refs} ref x {a.queries} query embeddings, dim {a.dim}, float L2 "
                f"(tcgen05 {a.precision} cross term), k={a.k}")
    if which == "cfg1":
        return (f"cfg1 faiss IndexFlatL2 build+search, 1 window: panel {a.refs} x {a.sites} float 0/1, {a.queries} queries, k={a.k}")
    return (f"cfg2 chr21-scale sweep: {a.windows} windows x {a.refs} ref haplotypes x {a.sites} sites, "
            f"{a.queries} queries/window, k={a.k}, bit-packed " + ("masked " if a.masked else "") + "Hamming")


def sub_args(a, **kw):
    b = argparse.Namespace(**vars(a))
    for k_, v in kw.items():
        setattr(b, k_, v)
    return b


def cfg_defaults(a, which):
    """The BASELINE.json shape of configuration `which`, keeping any size the caller set away from the cfg-2 defaults."""
    d = dict(windows=1000, refs=5008, sites=1030, queries=2000, k=8)
    cur = {n: getattr(a, n) for n in d}
    if which == "cfg5":
        new = dict(windows=8, refs=200000, queries=10000, k=32)
    elif which == "cfg4":
        new = dict(queries=4096)
    elif which == "cfg1":
        new = dict(windows=1, queries=1000, k=1)
    else:
        new = {}
    out = dict(cur)
    for n, v in new.items():
        if cur[n] == d[n]:
            out[n] = v
    return sub_args(a, workload=which, masked=(which == "cfg3") or (a.masked and which in ("cfg2", "cfg3")), **out)


# --------------------------------------------------------------------------- synthetic data
def gen_windows_device(torch, dev, seed, n_windows, n_rows, n_sites, founders_seed_base, chunk=25, w_first=0):
    """Mosaic-of-founders haplotypes generated on the device (SURVEY.md §8d hapgen): per window
    64 founders ~ Bernoulli(p_s), p_s ~ Beta(.25,.75); every haplotype copies a founder, switching
    with prob 1/200 per site, alleles flipped with prob 1e-3.  Returns packed uint32-as-int32
    [n_windows, n_rows, stride] (library pack kernel).  Seeds are keyed by the GLOBAL index of a chunk's first
    window (w_first + offset), so a window shard of a job holds exactly the job's windows."""
    from rag_snvbert_b200 import _lib
    from rag_snvbert_b200.index import pack_rows

    stride = _lib.packed_stride(n_sites)
    out = torch.empty((n_windows, n_rows, stride), dtype=torch.int32, device=dev)
    beta = torch.distributions.Beta(torch.tensor(0.25, device=dev), torch.tensor(0.75, device=dev))
    for w0 in range(0, n_windows, chunk):
        nw = min(chunk, n_windows - w0)
        gw = w_first + w0
        gf = torch.Generator(device=dev)
        gf.manual_seed(founders_seed_base + gw)  # founders shared by panel and queries of a window
        torch.manual_seed(founders_seed_base + gw)
        p = beta.sample((nw, 1, n_sites))
        F = (torch.rand((nw, 64, n_sites), device=dev, generator=gf) < p).to(torch.uint8)
        g = torch.Generator(device=dev)
        g.manual_seed(seed + gw)
        sw = torch.rand((nw, n_rows, n_sites), device=dev, generator=g) < (1.0 / 200)
        seg = torch.cumsum(sw.to(torch.int32), dim=2)  # segment id per site
        max_seg = int(seg.max().item()) + 1
        choice = torch.randint(0, 64, (nw, n_rows, max_seg), device=dev, generator=g)
        fid = torch.gather(choice, 2, seg.long())  # founder per site
        del sw, seg, choice
        hap = torch.gather(F, 1, fid)  # F[w, fid[w,r,s], s]
        del fid
        flip = torch.rand((nw, n_rows, n_sites), device=dev, generator=g) < 1e-3
        hap ^= flip.to(torch.uint8)
        del flip
        out[w0:w0 + nw] = pack_rows(hap.reshape(-1, n_sites), n_sites).reshape(nw, n_rows, stride)
        del hap
    return out


def gen_masks_device(torch, dev, seed, n_windows, nq, n_sites, chunk=50, w_first=0):
    from rag_snvbert_b200 import _lib
    from rag_snvbert_b200.index import pack_rows

    stride = _lib.packed_stride(n_sites)
    out = torch.empty((n_windows, nq, stride), dtype=torch.int32, device=dev)
    for w0 in range(0, n_windows, chunk):
        nw = min(chunk, n_windows - w0)
        g = torch.Generator(device=dev)
        g.manual_seed(seed + w_first + w0)
        rate = 0.1 + 0.8 * torch.rand((nw, nq, 1), device=dev, generator=g)
        obs = (torch.rand((nw, nq, n_sites), device=dev, generator=g) >= rate).to(torch.uint8)
        out[w0:w0 + nw] = pack_rows(obs.reshape(-1, n_sites), n_sites).reshape(nw, nq, stride)
    return out


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
            os.close(fd)
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
        self.f.close()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
                for nm, v in zip(names, parts[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            # median over the upper half = "under load" samples
            s = sorted(sm)
            out["sm_mhz"] = float(np.median(s[len(s) // 2:]))
            out["sm_max_mhz"] = float(max(mx))
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def sample_clocks_over(ctx, timed_seconds, step_fn):
    """The sampler (nvidia-smi every 100 ms) needs about a second under load: after a short timed region keep running the
    same step, untimed, until it has seen one."""
    torch = ctx.torch
    t0 = time.perf_counter()
    while timed_seconds + (time.perf_counter() - t0) < 1.2:
        step_fn()
        torch.cuda.synchronize()


# --------------------------------------------------------------------------- CPU arm
def host_threads() -> int:
    """Cores this process may run on (not OMP_NUM_THREADS: torchrun sets that to 1 for every rank)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_sample(a, seconds):
    """The reference's CPU algorithm for this path (faiss IndexBinaryFlat-style popcount scan +
    per-query heap, restated in oracle/snv_oracle.c: 'port'; faiss itself is absent from the
    image and the reference has no native code to compile) on a bounded sample of the workload,
    all host threads.  Returns (ref_haps_per_s, n_threads, sample description, seconds)."""
    from oracle import oracle as O
    from oracle import cbind

    # explicit thread count: torch.distributed.run exports OMP_NUM_THREADS=1 to every rank, which must not
    # handicap the baseline (the arm runs on rank 0 alone and may use every core the process is allowed on)
    threads = host_threads()
    s = (a.sites + 31) // 32
    stride = -(-s // 4) * 4
    # the sample is made of whole windows (a smaller slice under-reports the rate and the sample comes out short),
    # except for very large windows (cfg 5: 2 x 10^9 pairs each), whose queries are sub-sampled; the rate is per pair
    nq0 = int(min(a.queries, max(200, 4e8 // a.refs)))
    P = O.pack_bits_u32(O.hapgen(2000, a.refs, a.sites), stride)[None]
    Qfull = O.pack_bits_u32(O.hapgen(5000, nq0, a.sites, founder_seed=2000), stride)[None]
    M = None
    if a.masked:
        rng = np.random.default_rng(8000)
        rate = rng.uniform(0.1, 0.9, size=(nq0, 1))
        M = O.pack_bits_u32((rng.random((nq0, a.sites)) >= rate).astype(np.uint8), stride)[None]
    # calibrate on one window, then size the sample for ~`seconds`
    cbind.hamming_topk_packed(P, Qfull[:, :nq0], a.k, None if M is None else M[:, :nq0], words=s, n_threads=threads)
    t0 = time.perf_counter()
    cbind.hamming_topk_packed(P, Qfull[:, :nq0], a.k, None if M is None else M[:, :nq0], words=s, n_threads=threads)
    dt0 = max(time.perf_counter() - t0, 1e-6)
    rate0 = nq0 * a.refs / dt0
    n_win = int(max(1, min(a.windows, seconds * rate0 / (nq0 * a.refs))))
    Pn = np.ascontiguousarray(np.broadcast_to(P, (n_win,) + P.shape[1:]))
    Qn = np.ascontiguousarray(np.broadcast_to(Qfull, (n_win,) + Qfull.shape[1:]))
    Mn = None if M is None else np.ascontiguousarray(np.broadcast_to(M, (n_win,) + M.shape[1:]))
    t0 = time.perf_counter()
    cbind.hamming_topk_packed(Pn, Qn, a.k, Mn, words=s, n_threads=threads)
    dt = time.perf_counter() - t0
    val = n_win * nq0 * a.refs / dt
    sample = (f"{n_win} of {a.windows} windows x {nq0}" + ("" if nq0 == a.queries else f" of {a.queries}") +
              f" queries x {a.refs} refs, k={a.k}, "
              f"C popcount port (oracle/snv_oracle.c), {threads} OpenMP threads, {dt:.2f} s")
    return val, threads, sample, dt


def blas_l2_cpu(a, refs, q, what):
    """faiss's BLAS path restated with numpy/OpenBLAS (oracle.l2_topk_f32_blas), all host threads."""
    from oracle import oracle as O

    threads = host_threads()
    try:  # BLAS pools also obey torchrun's OMP_NUM_THREADS=1: set the pool size explicitly
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=threads)
    except Exception:
        import contextlib
        ctx = contextlib.nullcontext()
    with ctx:
        O.l2_topk_f32_blas(refs[:512], q[:256], a.k)
        t0 = time.perf_counter()
        O.l2_topk_f32_blas(refs, q, a.k)
        dt = time.perf_counter() - t0
    sample = (f"full step: {q.shape[0]} queries x {refs.shape[0]} refs x dim {refs.shape[1]}, k={a.k}, numpy/BLAS "
              f"|x|^2+|y|^2-2xy restatement of faiss ({what}oracle.l2_topk_f32_blas), {threads} host threads, {dt:.2f} s")
    return q.shape[0] * refs.shape[0] / dt, threads, sample, dt


def cfg4_inputs(a):
    refs_h = np.random.default_rng(4001).standard_normal((a.refs, a.dim)).astype(np.float32)
    q_h = np.random.default_rng(4002).standard_normal((a.queries, a.dim)).astype(np.float32)
    return refs_h, q_h


def cfg1_inputs(a):
    """cfg 1 arrays: float32 0/1 panel seed 1001, queries seed 1002 - plain numpy, no oracle import."""
    p = np.random.default_rng(1001).beta(0.25, 0.75, size=(1, a.sites))
    panel = (np.random.default_rng(1001).random((a.refs, a.sites)) < p).astype(np.float32)
    q = (np.random.default_rng(1002).random((a.queries, a.sites)) < p).astype(np.float32)
    return panel, q


def run_reference(a):
    """--impl reference: the reference's own CPU implementation of the path, on rank 0 alone."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    which = "cfg2" if a.workload in ("all", "cfg2", "cfg3") else a.workload
    a = cfg_defaults(a, "cfg3" if a.workload == "cfg3" else which)
    vals, times = [], []
    sample, threads = "", 1
    per_step = max(2.0, min(a.cpu_seconds, 120.0 / max(1, a.steps + a.warmup)))
    dtype = "u32-popcount"
    for i in range(a.warmup + a.steps):
        if which == "cfg4":
            refs_h, q_h = cfg4_inputs(a)
            v, threads, sample, dt = blas_l2_cpu(a, refs_h, q_h, "")
            dtype = "f32"
        elif which == "cfg1":
            panel, q = cfg1_inputs(a)
            v, threads, sample, dt = blas_l2_cpu(a, panel, q, "IndexFlatL2.search of batch_test_faiss_l2.py:110; add is a memcpy; ")
            dtype = "f32"
        else:
            v, threads, sample, dt = cpu_reference_sample(a, per_step)
        if i >= a.warmup:
            vals.append(v)
            times.append(dt)
    value = float(np.mean(vals))
    scaling = "strong" if (which == "cfg5" or (which == "cfg2" and a.scaling == "strong")) else "weak"
    line = {
        "impl": "reference", "metric": METRIC.replace("k=8", f"k={a.k}"), "value": value, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(np.mean(times) * 1e3),
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": dtype,
        "data": "synthetic", "config": {"workload": workload_name(a, which), "sample_per_step": sample},
        "window_queries_per_s": value / a.refs,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm plumbing
class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.peaks = {}
        try:
            self.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum_i64(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(x) for x in t.tolist()]

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def fp4_peak(ctx, engine):
    """Tensor-pipe peak for the narrow-float engines.  MEASURED_PEAKS.json holds a bf16 GEMM figure only; the fp4
    (kind::mxf4) rate measured on this pool with an MMA-only build of the scan kernel itself is recorded in
    profiles/tensor_peaks.json when available, else the nominal ratios (4x fp4, 2x fp8) of the measured bf16 peak."""
    bf16 = float(ctx.peaks.get("bf16_tflops", 1590.0))
    fp4 = engine >= 3
    try:
        tp = json.load(open(os.path.join(ROOT, "profiles", "tensor_peaks.json")))
        key = "mxf4_tflops" if fp4 else "f8f6f4_tflops"
        if tp.get(key):
            return float(tp[key]), f"measured: {tp.get('how', 'MMA-only build of the scan kernel')} (profiles/tensor_peaks.json)"
    except Exception:
        pass
    mult = 4.0 if fp4 else 2.0
    src = ("%g x measured bf16_tflops (%s rate, nominal ratio)" % (mult, "fp4" if fp4 else "fp8")
           if "bf16_tflops" in ctx.peaks else "%g x fallback 1590 TFLOP/s" % mult)
    return mult * bf16, src


def hamming_roofline(ctx, a, engine, kern_ms, step_ms, pairs, W, clocks, k_cap):
    S, N, Q = a.sites, a.refs, a.queries
    words = (S + 31) // 32
    scan_gbs = pairs * words * 4 / (kern_ms * 1e-3) / 1e9
    hbm_peak = float(ctx.peaks.get("hbm_gbs", 6650.0))
    sm_mhz = (clocks or {}).get("sm_mhz") or float(ctx.peaks.get("sm_max_mhz", 1965.0))
    traffic = None
    if engine in (1, 2, 3, 4, 5):
        fp4 = engine >= 3
        kname = "hamming_tc_kernel<K=%d,%s>" % (k_cap, {5: "fp4,cta-pair,tmemA", 4: "fp4,cta-pair", 3: "fp4"}.get(engine, "fp8"))
        peak, peak_src = fp4_peak(ctx, engine)
        achieved = 2.0 * pairs * S / (kern_ms * 1e-3) / 1e12
        k_per_mma = 64 if fp4 else 32
        n_tile = {5: 160, 4: 240, 3: 240}.get(engine, 256)
        mmas = -(-words * 32 // k_per_mma) + (1 if engine >= 4 else 0)  # + the column-index MMA of the list epilogue
        issued = 2.0 * W * (-(-Q // 128) * 128) * (-(-N // n_tile) * n_tile) * mmas * k_per_mma / (kern_ms * 1e-3) / 1e12
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            ent = tj.get(f"{kname}|W={W},N={N},Q={Q}")
            if ent and S == 1030:
                traffic = ent["bytes"]
        except Exception:
            pass
        roofline = {
            "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "traffic": traffic, "kernel": kname, "kernel_ms": kern_ms, "step_ms": step_ms,
            "peak_source": peak_src, "issued_tflops": issued,
            "frac_of_nominal_4x_bf16": achieved / (4.0 * float(ctx.peaks.get("bf16_tflops", 1590.0))) if fp4 else None,
            "note": ("algorithmic FLOP = 2 x pairs x sites; issued_tflops counts the tile and K padding the MMAs "
                     "really execute (incl. the column-index MMA). See DESIGN.md for the measured limiters"),
            "scan_equivalent": {"achieved": scan_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": scan_gbs / hbm_peak,
                                "note": "pairs x 132 B / kernel time against the measured HBM copy peak (SURVEY.md 8d reading)"},
        }
        dtype = ("fp4-e2m1" if fp4 else "fp8-e4m3") + " (exact 0/+-1 products, fp32 accumulate)"
    else:
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in ctx.peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        popc_per_pair = 16 if words == 33 else words  # CSA depth 2 leaves 16 POPC for 33 words
        kname = "hamming_topk_kernel<33,masked=%d,K=%d>" % (1 if a.masked else 0, k_cap)
        roofline = {
            "bound": "hbm", "achieved": scan_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": scan_gbs / hbm_peak,
            "traffic": traffic, "kernel": kname, "kernel_ms": kern_ms, "step_ms": step_ms, "peak_source": peak_src,
            "note": ("scan-equivalent bandwidth = pairs x 132 B / kernel time (SURVEY.md 8d): the panel tile is "
                     "served from shared memory/L2, so this may exceed 1.0; the binding unit is the integer pipes"),
            "int_pipe": {"pairs_per_s": pairs / (kern_ms * 1e-3), "alu_instr_per_pair": 68 if words == 33 else None,
                         "popc_per_pair": popc_per_pair,
                         "alu_frac_of_64_per_clk_sm": (pairs / (kern_ms * 1e-3)) * 68 / (148 * 64 * sm_mhz * 1e6) if words == 33 else None},
        }
        dtype = "u32-popcount"
    return roofline, dtype


# --------------------------------------------------------------------------- cfg 2 / 3: window-sharded Hamming sweep
def bench_windows(ctx, a, scaling, steps, warmup, want_e2e=True, want_clocks=True):
    """One measured pass of the cfg-2/3 sweep.  scaling 'strong': this rank owns windows shard_range(W, world, rank) of the
    job's W windows; 'weak': its own W windows.  No data-path collective either way.  Returns the line's pieces."""
    torch = ctx.torch
    from rag_snvbert_b200 import WindowedHammingIndex, _lib
    from rag_snvbert_b200.sharding import shard_range

    W, N, S, Q, k = a.windows, a.refs, a.sites, a.queries, a.k
    stride = _lib.packed_stride(S)
    if scaling == "strong":
        w_lo, w_hi = shard_range(W, ctx.world, ctx.rank)
        seed_off = 0
    else:
        w_lo, w_hi = 0, W
        seed_off = 100000 * ctx.rank
    Wr = w_hi - w_lo
    panel = gen_windows_device(torch, ctx.dev, 2000 + seed_off, Wr, N, S, 777 + seed_off, w_first=w_lo)
    queries = gen_windows_device(torch, ctx.dev, 5000 + seed_off, Wr, Q, S, 777 + seed_off, w_first=w_lo)
    masks = gen_masks_device(torch, ctx.dev, 8000 + seed_off, Wr, Q, S, w_first=w_lo) if a.masked else None
    index = WindowedHammingIndex(S, max(Wr, 1), ctx.local)
    index.add(panel)
    del panel
    torch.cuda.synchronize()

    def step():
        return index.search(queries, k, observed=masks)

    for _ in range(max(warmup, 3)):
        D, I = step()
    ctx.barrier()
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0 and want_clocks:
        sampler.start()
    _lib.profile_enable(True)
    launches0 = _lib.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.barrier()
    t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0e.record()
    for i in range(steps):
        ev[i][0].record()
        D, I = step()
        ev[i][1].record()
    t1e.record()
    ctx.barrier()
    launches = _lib.launch_count() - launches0
    total_ms = t0e.elapsed_time(t1e)
    kern_ms = _lib.profile_last_ms()  # the scan kernel alone, events on its own stream (last timed step)
    _lib.profile_enable(False)
    engine = _lib.last_hamming_engine()
    step_ms_events = float(np.mean([s.elapsed_time(e) for s, e in ev]))
    clocks = None
    if ctx.rank == 0 and want_clocks:
        sample_clocks_over(ctx, total_ms * 1e-3, step)
        clocks = sampler.stop()
        clocks["note"] = "sampled every 100 ms over the timed region and, when that is shorter than 1.2 s, over further identical steps run right after it"
    ctx.barrier()
    ms_per_step = ctx.allmax(total_ms) / steps
    pairs_rank = Wr * Q * N
    pairs_job = W * Q * N if scaling == "strong" else ctx.world * pairs_rank
    value = pairs_job / (ms_per_step * 1e-3)
    chk = ctx.allsum_i64([int(I.sum().item()), int(D.sum().item())])

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H in the timed region
    e2e = None
    if want_e2e and Wr > 0:
        def pinned(shape, dt):
            return torch.empty(shape, dtype=dt, pin_memory=True)

        # host rows in the dense packed layout (33 words = 132 bytes per 1030-site row; the device re-strides to 36)
        words = (S + 31) // 32
        hq = pinned((Wr, Q, words), torch.int32)
        hq.copy_(queries[:, :, :words])
        hm = None
        if masks is not None:
            hm = pinned((Wr, Q, words), torch.int32)
            hm.copy_(masks[:, :, :words])
        hq_np, hm_np = hq.numpy(), (None if hm is None else hm.numpy())
        hD = pinned((Wr, Q, k), torch.int32).numpy()   # caller-owned pinned result buffers
        hI = pinned((Wr, Q, k), torch.int64).numpy()
        for _ in range(2):
            Dh, Ih = index.search(hq_np, k, observed=hm_np, out=(hD, hI))
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            Dh, Ih = index.search(hq_np, k, observed=hm_np, out=(hD, hI))
        torch.cuda.synchronize()
        dt = ctx.allmax(time.perf_counter() - t0) / steps
        h2d = hq_np.nbytes + (0 if hm_np is None else hm_np.nbytes)
        d2h = Dh.nbytes + Ih.nbytes
        e2e = {"value": pairs_job / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": dt * 1e3, "host_copy_gbs_per_rank": (h2d + d2h) / dt / 1e9,
               "api": "WindowedHammingIndex.search(pinned numpy dense packed uint32 [W,Q,33], out=pinned (D int32, I int64)); "
                      "window chunks pipelined over 3 streams inside libsnvknn (H2D | scan | D2H); bytes are per rank"}
        assert np.array_equal(Ih, I.cpu().numpy()) and np.array_equal(Dh, D.cpu().numpy()), "host/device result mismatch"
        # compact wire format: int32 ids + uint16 distances (ids < panel rows, distances <= sites), same results
        if hasattr(index, "search_compact"):
            hD16 = pinned((Wr, Q, k), torch.int16).numpy().view(np.uint16)
            hI32 = pinned((Wr, Q, k), torch.int32).numpy()
            for _ in range(2):
                index.search_compact(hq_np, k, observed=hm_np, out=(hD16, hI32))
            ctx.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                index.search_compact(hq_np, k, observed=hm_np, out=(hD16, hI32))
            torch.cuda.synchronize()
            dtc = ctx.allmax(time.perf_counter() - t0) / steps
            assert np.array_equal(hI32.astype(np.int64), Ih) and np.array_equal(hD16.astype(np.int32), Dh), "compact result mismatch"
            d2hc = hD16.nbytes + hI32.nbytes
            e2e["compact"] = {"value": pairs_job / dtc, "ms_per_step": dtc * 1e3, "h2d_bytes_per_step": int(h2d),
                              "d2h_bytes_per_step": int(d2hc), "host_copy_gbs_per_rank": (h2d + d2hc) / dtc / 1e9,
                              "api": "WindowedHammingIndex.search_compact(..., out=pinned (D uint16, I int32)): same (D, I), 6 instead of 12 bytes per neighbour on the wire"}
    return {"value": value, "ms_per_step": ms_per_step, "e2e": e2e, "launches": int(launches), "engine": engine,
            "kern_ms": kern_ms, "step_ms": step_ms_events, "clocks": clocks, "pairs_rank": pairs_rank, "windows_rank": Wr,
            "checksum": chk}


def line_windows(ctx, a, r, scaling, steps, warmup):
    roofline, dtype = hamming_roofline(ctx, a, r["engine"], r["kern_ms"], r["step_ms"], r["pairs_rank"], r["windows_rank"], r["clocks"],
                                       8 if a.k <= 8 else 32)
    W, N, Q, S = a.windows, a.refs, a.queries, a.sites
    stride_b = -(-((S + 31) // 32) // 4) * 16
    return {
        "metric": METRIC.replace("k=8", f"k={a.k}"), "value": r["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": steps,
        "warmup": max(warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": workload_name(a, "cfg2"), "windows_per_gpu": r["windows_rank"], "engine": ENGINE_NAMES[r["engine"]],
                   "parallelism": (f"the job's {W} windows window-sharded over {ctx.world} GPU(s), no collective" if scaling == "strong"
                                   else f"every rank its own {W} windows x{ctx.world}, no collective"),
                   "l2_policy": f"inputs larger than L2 (packed panel {r['windows_rank'] * N * stride_b / 1e6:.0f} MB + queries "
                                f"{r['windows_rank'] * Q * stride_b / 1e6:.0f} MB per GPU per step; L2 is 126 MB)"},
        "window_queries_per_s": r["value"] / N, "e2e": r["e2e"], "gpu_launches": r["launches"], "roofline": roofline,
        "clocks": r["clocks"], "checksum": r["checksum"],
    }


# --------------------------------------------------------------------------- cfg 4: float L2 on tcgen05
def bench_cfg4(ctx, a, steps, warmup, want_cpu):
    torch = ctx.torch
    from rag_snvbert_b200 import WindowedL2Index, _lib

    N, Q, d, k = a.refs, a.queries, a.dim, a.k
    refs_h, q_h = cfg4_inputs(a)
    refs = torch.from_numpy(refs_h).to(ctx.dev)
    q = torch.from_numpy(q_h).to(ctx.dev)
    index = WindowedL2Index(d, 1, ctx.local, a.precision)
    index.add(refs)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.dev)  # > 126 MB L2
    for _ in range(max(warmup, 3)):
        D, I = index.search(q, k)
    ctx.barrier()
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0:
        sampler.start()
    _lib.profile_enable(True)
    launches0 = _lib.launch_count()
    step_ms, kern_ms = [], []
    for i in range(steps):
        flush.zero_()  # L2 flush between timed iterations (inputs are 9 MB)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        e0.record()
        D, I = index.search(q, k)
        e1.record()
        ctx.barrier()
        step_ms.append(e0.elapsed_time(e1))
        kern_ms.append(_lib.profile_last_ms())
    launches = _lib.launch_count() - launches0
    _lib.profile_enable(False)
    clocks = None
    if ctx.rank == 0:
        sample_clocks_over(ctx, float(np.sum(step_ms)) * 1e-3, lambda: index.search(q, k))
        clocks = sampler.stop()
    ms_per_step = ctx.allmax(float(np.sum(step_ms))) / steps
    value = ctx.world * Q * N / (ms_per_step * 1e-3)

    # the reference's own GPU path for this search (src/dataset/embedding_rag_dataset.py:397-402), warmed, same flush
    def cdist_topk():
        return torch.cdist(q, refs, p=2).topk(k, dim=1, largest=False)

    for _ in range(3):
        cdist_topk()
    torch.cuda.synchronize()
    cd = []
    for i in range(min(steps, 20)):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        cdist_topk()
        e1.record()
        torch.cuda.synchronize()
        cd.append(e0.elapsed_time(e1))

    hq = torch.empty((Q, d), dtype=torch.float32, pin_memory=True)
    hq.copy_(q)
    hq_np = hq.numpy()
    for _ in range(2):
        Dh, Ih = index.search(hq_np, k)
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        Dh, Ih = index.search(hq_np, k)
    torch.cuda.synchronize()
    dt = ctx.allmax((time.perf_counter() - t0) / steps)
    assert np.array_equal(Ih, I.cpu().numpy()), "host/device result mismatch"
    e2e = {"value": ctx.world * Q * N / dt, "unit": UNIT, "h2d_bytes_per_step": int(hq_np.nbytes),
           "d2h_bytes_per_step": int(Dh.nbytes + Ih.nbytes), "ms_per_step": dt * 1e3,
           "api": "WindowedL2Index.search(pinned numpy float32 [Q,d]) -> numpy (D float32, I int64)"}
    bf16 = float(ctx.peaks.get("bf16_tflops", 1590.0))
    peak = bf16 / 2.0  # kind::tf32 issues at half the bf16 rate
    km = float(np.mean(kern_ms))
    flops = 2.0 * Q * N * d
    achieved = flops / (km * 1e-3) / 1e12
    passes = 3 if a.precision == "tf32x3" else 1
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": None, "kernel": "l2_topk_kernel<8>", "kernel_ms": km, "step_ms": float(np.mean(step_ms)),
                "issued_tflops": achieved * passes, "issued_frac": achieved * passes / peak,
                "peak_source": ("measured bf16_tflops / 2 (tf32 rate)" if "bf16_tflops" in ctx.peaks else "fallback 1590/2"),
                "note": "achieved = algorithmic 2*Q*N*d FLOP / kernel time; tf32x3 issues 3x that on the tensor pipe (issued_*)"}
    cpu = None
    if want_cpu and ctx.rank == 0:
        v, threads, sample, _ = blas_l2_cpu(a, refs_h, q_h, "")
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    return {
        "metric": METRIC.replace("k=8", f"k={k}"), "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32",
        "data": "synthetic", "config": {"workload": workload_name(a, "cfg4"), "parallelism": f"replicas x{ctx.world}",
                                        "l2_policy": "256 MB buffer written between timed iterations (L2 flush)"},
        "window_queries_per_s": value / N, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        "torch_cdist_topk_ms": float(np.mean(cd)),
        "torch_cdist_topk_note": "torch.cdist(q, refs).topk(k) on the same GPU, warmed, same L2 flush: the reference's GPU path (embedding_rag_dataset.py:397-402)",
        "cpu_baseline": cpu, "clocks": clocks}


# --------------------------------------------------------------------------- cfg 1: faiss.IndexFlatL2 add + search (drop-in module)
def bench_cfg1(ctx, a, steps, warmup, want_cpu):
    """BASELINE configs[0] exactly as batch_test_faiss_l2.py:109-111 / build_ref_db_l2.py:89-90 code it: float32 0/1 rows from
    HOST numpy arrays, index = faiss.IndexFlatL2(d); index.add(panel); D, I = index.search(q, k) - through faiss_compat."""
    torch = ctx.torch
    import rag_snvbert_b200.faiss_compat as faiss
    from rag_snvbert_b200 import _lib

    panel, q = cfg1_inputs(a)
    k = a.k
    torch.cuda.synchronize()
    torch.cuda.empty_cache()  # the library allocates with cudaMalloc: give back what torch cached for the earlier workloads

    def step():
        index = faiss.IndexFlatL2(a.sites)
        index.add(panel)
        return index.search(q, k)

    for _ in range(max(warmup, 3)):
        D, I = step()
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    t0 = time.perf_counter()
    for _ in range(steps):
        D, I = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    launches = _lib.launch_count() - launches0
    # search alone (index resident), the timed region of batch_test_faiss_l2.py:109-111
    index = faiss.IndexFlatL2(a.sites)
    index.add(panel)
    for _ in range(3):
        index.search(q, k)
    t0 = time.perf_counter()
    for _ in range(steps):
        D2, I2 = index.search(q, k)
    torch.cuda.synchronize()
    dts = (time.perf_counter() - t0) / steps
    assert np.array_equal(I, I2) and np.array_equal(D, D2)
    pairs = a.queries * a.refs
    cpu = None
    if want_cpu and ctx.rank == 0:
        v, threads, sample, _ = blas_l2_cpu(a, panel, q, "IndexFlatL2.search of batch_test_faiss_l2.py:110; ")
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    e2e = {"value": pairs / dt, "unit": UNIT, "h2d_bytes_per_step": int(panel.nbytes + q.nbytes), "d2h_bytes_per_step": int(D.nbytes + I.nbytes),
           "ms_per_step": dt * 1e3, "api": "faiss_compat.IndexFlatL2(d).add(numpy float32 panel) + .search(numpy float32 q, k): host arrays in, numpy (D, I) out"}
    hbm = float(ctx.peaks.get("hbm_gbs", 6650.0))
    return {
        "metric": METRIC.replace("k=8", f"k={k}"), "value": pairs / dts, "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": dts * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "tf32 (exact on 0/1 rows)",
        "data": "synthetic", "config": {"workload": workload_name(a, "cfg1"), "search_path": index.last_search_path,
                                        "note": "value = search alone from host arrays with the index resident (the reference's timed region); e2e = add + search"},
        "window_queries_per_s": a.queries / dts, "e2e": e2e, "gpu_launches": int(launches), "cpu_baseline": cpu,
        "roofline": {"bound": "hbm", "note": "launch / PCIe-bound at this size (SURVEY 8d): 20.6 MB panel + 4.1 MB queries per call",
                     "achieved": (panel.nbytes + q.nbytes) / dt / 1e9, "peak": hbm, "unit": "GB/s",
                     "frac": (panel.nbytes + q.nbytes) / dt / 1e9 / hbm, "traffic": None}}


# --------------------------------------------------------------------------- cfg 5: row-sharded panel + NCCL merge
CFG5_BLOCK = 5000  # panel rows are generated in blocks keyed by their GLOBAL block id: the panel is the same for every GPU count


def bench_cfg5(ctx, a, steps, warmup, want_cpu):
    """Biobank-scale panel: N = 200,000 haplotypes x 1,030 sites, 10,000 queries per window, k = 32.
    The panel ROWS are sharded over the ranks (25,000 per GPU at 8 GPUs); every rank scans its rows
    for all queries with global ids, then one all-to-all of (D, I) and an on-device merge (result sharded by query).
    Strong scaling: the job (windows x N x Q) is fixed, `--windows` of the 500 are run per step."""
    torch = ctx.torch
    from rag_snvbert_b200 import WindowedHammingIndex, _lib
    from rag_snvbert_b200.sharding import RowShardedSearch, shard_range

    N, Q, k, W, S = a.refs, a.queries, a.k, a.windows, a.sites
    blk = CFG5_BLOCK if N % (CFG5_BLOCK * ctx.world) == 0 else max(1, N // ctx.world)
    lo, hi = shard_range(N, ctx.world, ctx.rank)
    queries = gen_windows_device(torch, ctx.dev, 5000, W, Q, S, 777, chunk=1)
    stride = _lib.packed_stride(S)
    panel = torch.empty((W, hi - lo, stride), dtype=torch.int32, device=ctx.dev)
    for b0 in range(lo, hi, blk):
        nb = min(blk, hi - b0)
        panel[:, b0 - lo:b0 - lo + nb] = gen_windows_device(torch, ctx.dev, 9000 + 7919 * (b0 // blk), W, nb, S, 777, chunk=1)
    index = WindowedHammingIndex(S, W, ctx.local)
    index.add(panel)
    del panel
    searcher = RowShardedSearch(index, lo, world=ctx.world, chunks=int(os.environ.get("SNV_CFG5_CHUNKS", "0")) or None)

    def step():
        # sync=False: the exchange + merge of this batch stay on the side stream and overlap the next batch's scan; the
        # timed region ends with searcher.wait() + a device synchronisation, so every batch is complete inside it
        return searcher.search(queries, k, sync=False)

    for _ in range(max(warmup, 3)):
        q_lo, q_hi, D, I = step()
    searcher.wait()
    ctx.barrier()
    sampler = ClockSampler(ctx.local)
    if ctx.rank == 0:
        sampler.start()
    _lib.profile_enable(True)
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    for _ in range(steps):
        q_lo, q_hi, D, I = step()
    searcher.wait()
    e1.record()
    ctx.barrier()
    kern_ms = _lib.profile_last_ms()
    launches = _lib.launch_count() - launches0
    _lib.profile_enable(False)
    total_ms = e0.elapsed_time(e1)
    clocks = None
    if ctx.rank == 0:
        sample_clocks_over(ctx, total_ms * 1e-3, lambda: index.search(queries, k, id_offset=lo))
        clocks = sampler.stop()
    ctx.barrier()
    ms_per_step = ctx.allmax(total_ms) / steps
    value = W * Q * N / (ms_per_step * 1e-3)
    # checksum of the whole (query-sharded) result: equal for every GPU count because the panel is
    chk = ctx.allsum_i64([int(I.sum().item()), int(D.sum().item())])
    # e2e: queries from pinned host memory every step, this rank's slice of the merged result back to the host
    hq = torch.empty(tuple(queries.shape), dtype=torch.int32, pin_memory=True)
    hq.copy_(queries)
    dq = torch.empty_like(queries)
    hD = torch.empty(tuple(D.shape), dtype=D.dtype, pin_memory=True)
    hI = torch.empty(tuple(I.shape), dtype=I.dtype, pin_memory=True)

    def step_e2e():
        dq.copy_(hq, non_blocking=True)
        _, _, D2, I2 = searcher.search(dq, k)
        hD.copy_(D2, non_blocking=True)
        hI.copy_(I2, non_blocking=True)

    for _ in range(2):
        step_e2e()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    torch.cuda.synchronize()
    dt = ctx.allmax((time.perf_counter() - t0) / steps)
    e2e = {"value": W * Q * N / dt, "unit": UNIT, "h2d_bytes_per_step": int(hq.numel() * 4),
           "d2h_bytes_per_step": int(hD.numel() * hD.element_size() + hI.numel() * 8), "ms_per_step": dt * 1e3,
           "api": "RowShardedSearch.search: pinned host queries -> device, local scan with global ids, NCCL exchange, merge, "
                  "this rank's query slice of (D, I) -> pinned host; bytes are per rank"}
    a_shard = sub_args(a, refs=hi - lo)
    roofline, dtype = hamming_roofline(ctx, a_shard, _lib.last_hamming_engine(), kern_ms, ms_per_step, W * Q * (hi - lo), W, clocks, 8 if k <= 8 else 32)
    roofline["note"] = "per-GPU local shard scan (rows_per_gpu): algorithmic 2 x pairs x sites FLOP / kernel time; " + roofline.get("note", "")
    cpu = None
    if want_cpu and ctx.rank == 0:
        v, threads, sample, _ = cpu_reference_sample(a, 4.0)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
    return {
        "metric": METRIC.replace("k=8", f"k={k}"), "value": value, "unit": UNIT, "n_gpus": ctx.world, "steps": steps,
        "warmup": max(warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": f"cfg5 biobank-scale: {W} of 500 windows x {N} ref haplotypes x {S} sites, {Q} queries/window, "
                               f"k={k}, panel row-sharded over {ctx.world} GPU(s) + NCCL exchange of the per-shard top-k + on-device merge (result sharded by query)",
                   "engine": ENGINE_NAMES[_lib.last_hamming_engine()], "rows_per_gpu": hi - lo,
                   "l2_policy": f"panel shard larger than L2 ({W * (hi - lo) * stride * 4 / 1e6:.0f} MB per GPU)",
                   "exchange": searcher.describe()},
        "window_queries_per_s": value / N, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
        "cpu_baseline": cpu, "checksum": chk}


# --------------------------------------------------------------------------- driver
def run_ours(a):
    ctx = Ctx()
    want_cpu = not a.no_cpu_baseline
    which = a.workload
    line = None
    if which in ("all", "cfg2", "cfg3"):
        a2 = cfg_defaults(a, "cfg3" if which == "cfg3" else "cfg2")
        scaling = a.scaling if ctx.world > 1 else "strong"
        r = bench_windows(ctx, a2, scaling, a.steps, a.warmup, want_e2e=not a.no_e2e)
        if ctx.rank == 0:
            line = line_windows(ctx, a2, r, scaling, a.steps, a.warmup)
        if ctx.world > 1 and which == "all" and not a.no_secondary:
            other = "weak" if scaling == "strong" else "strong"
            ro = bench_windows(ctx, a2, other, max(3, min(a.steps, 10)), a.warmup, want_e2e=not a.no_e2e, want_clocks=False)
            if ctx.rank == 0:
                line[other] = {"value": ro["value"], "ms_per_step": ro["ms_per_step"], "windows_per_gpu": ro["windows_rank"],
                               "e2e": ro["e2e"], "kernel_ms": ro["kern_ms"],
                               "note": ("every rank its own %d windows" % a2.windows) if other == "weak" else "the job's windows split over the ranks"}
        if ctx.rank == 0:
            line["cpu_baseline"] = None
            if want_cpu:
                v, threads, sample, _ = cpu_reference_sample(a2, a.cpu_seconds)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        if which == "all" and not a.no_secondary:
            ss, sw = max(3, min(a.steps, 10)), 3
            r5 = bench_cfg5(ctx, cfg_defaults(a, "cfg5"), ss, sw, want_cpu)
            r4 = bench_cfg4(ctx, cfg_defaults(a, "cfg4"), max(ss, 20), sw, want_cpu)
            r1 = bench_cfg1(ctx, cfg_defaults(a, "cfg1"), ss, sw, want_cpu) if ctx.rank == 0 else None
            if ctx.rank == 0:
                sec = {"cfg5": r5, "cfg4": r4, "cfg1": r1}
                line["secondary"] = sec
                line["gpu_launches_total"] = line["gpu_launches"] + sum(int(v.get("gpu_launches", 0)) for v in sec.values() if v)
    elif which == "cfg4":
        line = bench_cfg4(ctx, cfg_defaults(a, "cfg4"), a.steps, a.warmup, want_cpu)
    elif which == "cfg5":
        line = bench_cfg5(ctx, cfg_defaults(a, "cfg5"), a.steps, a.warmup, want_cpu)
    elif which == "cfg1":
        line = bench_cfg1(ctx, cfg_defaults(a, "cfg1"), a.steps, a.warmup, want_cpu) if ctx.rank == 0 else None
    if ctx.rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    ctx.barrier()
    ctx.close()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
