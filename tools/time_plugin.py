"""Where the plugin seam's host time goes (config 2): pageable vs pinned outputs, csr_array construction variants."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fempy_b200 import AssemblyPlan, Constitutive, Elements
from scipy.sparse import csr_array

wl = bench.make_workload("c2")
el = Elements.QuadElement2D(1)
plan = AssemblyPlan(wl["numNodes"])
plan.addBlock(*el.tables(), wl["conn"]["quad"])
plan.finalize()
mat = Constitutive.IsoPlaneStress(70e9, 0.3, 2700.0, 1.0).material()
indptr, indices = plan.pattern()

def t(label, fn, n=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    print(f"{label:60s} {1e3 * (time.perf_counter() - t0) / n:8.2f} ms", flush=True)
    return out

args = (wl["coords"], wl["states"], wl["thick"], mat, wl["bcDOF"], wl["bcVal"], True)
K, _ = t("assemble K only, fresh pageable np.empty output", lambda: plan.assemble(*args, None, wantR=False))
Kbuf = np.empty(plan.nnz)
t("assemble K only, reused pageable output", lambda: plan.assemble(*args, None, wantR=False, Kout=Kbuf))
import torch
Kpin = torch.empty(plan.nnz, dtype=torch.float64).pin_memory().numpy()
t("assemble K only, reused pinned output", lambda: plan.assemble(*args, None, wantR=False, Kout=Kpin))
t("assemble R only", lambda: plan.assemble(*args, wl["loads"], wantK=False))
M = t("csr_array((K, indices, indptr)) constructor", lambda: csr_array((K, indices, indptr), shape=(plan.numDOF, plan.numDOF)))
print("index dtype after constructor:", M.indices.dtype, "shares indices:", np.shares_memory(M.indices, indices))
t("template._with_data(K, copy=False)", lambda: M._with_data(K, copy=False))
t("np.empty(nnz) + first touch", lambda: np.empty(plan.nnz).fill(0.0))
t("convertBCDictToLists-style list building (1026 entries)", lambda: (list(wl["bcDOF"]), list(wl["bcVal"])))
