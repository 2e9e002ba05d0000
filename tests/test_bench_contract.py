"""CPU: the parts of the bench.py contract that do not need a GPU.

* `--impl reference` (the reference's CPU algorithm for the path, restated in oracle/snv_oracle.c, all host threads)
  prints ONE JSON line with the agreed keys; under a multi-rank launch only rank 0 works and prints.
* the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.pop("SNV_HAMMING_ENGINE", None)
    if env:
        e.update(env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True,
                          text=True, timeout=timeout)


def _json_lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-seconds", "2", "--windows", "40"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["metric"].startswith("ref-haplotypes scanned/s") and d["unit"] == "ref-haplotypes/s"
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["higher_is_better"] is True and d["scaling"] == "strong" and d["vs_baseline"] is None  # the job's windows, as BASELINE words it
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == d["unit"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_do_nothing():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"},
             timeout=120)
    assert r.returncode == 0 and _json_lines(r.stdout) == []


def test_product_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        return  # covered by the GPU run of bench.py itself
    r = _run(["--steps", "1", "--warmup", "0"], timeout=300)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
    assert _json_lines(r.stdout) == []


def test_reference_arm_under_torchrun_two_ranks():
    """the driver's own launch line for N > 1: rank 0 alone measures and prints, the other rank exits 0"""
    e = dict(os.environ)
    e.pop("SNV_HAMMING_ENGINE", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--cpu-seconds", "2", "--windows", "20"]
    r = subprocess.run(cmd, cwd=ROOT, env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2 and lines[0]["value"] > 0
    # torchrun exports OMP_NUM_THREADS=1 to its ranks: the baseline must still use every core the process may run on
    cores = len(os.sched_getaffinity(0))
    assert lines[0]["cpu_baseline"]["cores"] == cores
    assert f"{cores} OpenMP threads" in lines[0]["cpu_baseline"]["sample"]


def test_reference_arm_cfg5_runs_on_the_host():
    """cfg 5 (200k-row panel, k = 32): the CPU arm sub-samples the queries of a window and needs no GPU"""
    r = _run(["--workload", "cfg5", "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "2", "--refs", "20000"])
    assert r.returncode == 0, r.stderr[-2000:]
    (d,) = _json_lines(r.stdout)
    assert d["impl"] == "reference" and "k=32" in d["metric"] and d["scaling"] == "strong" and d["value"] > 0
    assert "cfg5" in d["config"]["workload"] and "20000 ref" in d["config"]["workload"]
