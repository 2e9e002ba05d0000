#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${TAG:-r16}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err; echo "bench rc=$?"
python - <<EOF2
import time, numpy as np, sys, torch
sys.path.insert(0, ".")
import bench
import rag_snvbert_b200.faiss_compat as faiss
dev = torch.device("cuda", 0)
N, Q, d = 5008, 20000, 1030
torch.manual_seed(0)
panel = (torch.rand(N, d, device=dev) < 0.3).float(); q = (torch.rand(Q, d, device=dev) < 0.3).float()
res = {}
for fp in (True, False):
    idx = faiss.IndexFlatL2(d, binary_fast_path=fp); idx.add(panel)
    for _ in range(3): D, I = idx.search(q, 8)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): D, I = idx.search(q, 8)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    res[fp] = (D.clone(), I.clone())
    print("IndexFlatL2 20000 x 5008 x 1030 0/1 float (CUDA tensors), k=8: fast_path=%s path=%s %.3f ms" % (fp, idx.last_search_path, (t1 - t0) / 10 * 1e3))
print("identical:", bool(torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])))
EOF2
