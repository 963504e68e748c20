"""Value parity AT THE BASELINE SIZES (BASELINE.json configs 2-5), pinned to the unmodified reference.

tests/golden/baseline_digests.npz holds digests (oracle/digests.py) of the systems the reference's own
FEMpyProblem._assembleMatrix / _assembleResidual (FEMpy/Problem.py:565-661) and computeFunction (:282-355) produce on the
bench.py workloads: whole meshes for configs 2-4, eight strips of 16 full-width element rows for config 5 (the reference
cannot hold that mesh).  CPU tests: the numpy oracle reproduces the digests (config 2 whole, strips of config 5).
GPU tests: the CUDA path through the C ABI reproduces them at full size -- pattern bit-exact (SHA-256 of indptr / indices),
K, R, von Mises within TOL = 1e-12 of the array's largest magnitude.
"""
import os

import numpy as np
import pytest

from util import MAT, ROOT, make_cm, oracle_blocks, ELEMENT_FOR

import bench
from fempy_b200 import AssemblyPlan
from oracle import digests as dg
from oracle import fem_oracle as fo

TOL = 1e-12
FIXTURE = os.path.join(ROOT, "tests", "golden", "baseline_digests.npz")


@pytest.fixture(scope="module")
def ref():
    return np.load(FIXTURE)


def compare(ref, got, prefix, tol=TOL):
    """every entry of the digest `got` (dict, same prefix) against the reference's"""
    keys = [k for k in ref.files if k.startswith(prefix + "_")]
    assert keys and set(keys) == set(got), (sorted(set(keys) ^ set(got)))
    r = {k[len(prefix) + 1 :]: ref[k] for k in keys}
    g = {k[len(prefix) + 1 :]: got[k] for k in keys}
    worst = {}
    for name in ("shape", "pattern_sha", "rows_global", "elems_global", "row_len", "cols_sha"):
        if name in r:
            assert np.array_equal(r[name], g[name]), f"{prefix}: {name} differs"
    kscale = float(r["K_max"]) if "K_max" in r else float(np.abs(r["K_row_abs"]).max())
    for name, scale in (("K_samples", kscale), ("K_max", kscale), ("Kv_rows", kscale), ("K_row_abs", kscale),
                        ("R_rows", float(np.abs(r["R_rows"]).max())), ("vm_elems", float(np.abs(r["vm_elems"]).max()))):  # fmt: skip
        if name in r:
            worst[name] = float(np.abs(r[name] - g[name]).max() / scale)
            assert worst[name] < tol, f"{prefix}: {name} off by {worst[name]:.2e} of its scale"
    for name in ("K_abs_sum", "R_abs_sum", "vm_sum"):  # sums of millions of terms: relative to the sum itself
        if name in r:
            worst[name] = float(abs(r[name] - g[name]) / abs(r[name]))
            assert worst[name] < tol, f"{prefix}: {name} off by {worst[name]:.2e}"
    if "R_sum" in r:  # cancels heavily: relative to the sum of magnitudes
        assert abs(r["R_sum"] - g["R_sum"]) < tol * float(r["R_abs_sum"]) * 10
    return worst


def oracle_system(elType, coords, conn, states, thick, bcDOF, bcVal, loads):
    blocks = oracle_blocks({elType: conn})
    K = fo.assemble_matrix(blocks, coords, states, thick, 0, MAT["E"], MAT["nu"], bcDOF, dropZeros=False)
    K.sort_indices()
    R = fo.assemble_residual(blocks, coords, states, thick, 0, MAT["E"], MAT["nu"], loads, bcDOF, bcVal)
    vm = fo.compute_function(blocks, coords, states, fo.VON_MISES, "mean", 0, MAT["E"], MAT["nu"], MAT["rho"])[0]
    return K, R, vm


def test_fixture_is_complete(ref):
    for cfg in ("c2", "c3", "c4"):
        assert f"{cfg}_pattern_sha" in ref.files and f"{cfg}_vm_elems" in ref.files
    assert sum(k.endswith("_cols_sha") for k in ref.files) == len(dg.C5_STRIPS)
    # sizes of SURVEY.md section 8: DOF and nnz of the four configurations
    assert tuple(ref["c2_shape"][:2]) == (526338, 9449476)
    assert tuple(ref["c3_shape"][:2]) == (2004002, 64032004)
    assert tuple(ref["c4_shape"][:2]) == (4999122, 69937124)


def test_oracle_reproduces_reference_digest_config2(ref):
    wl = bench.make_workload("c2")
    thick = dg.parity_inputs(wl)
    K, R, vm = oracle_system("quad", wl["coords"], wl["conn"]["quad"], wl["states"], thick, wl["bcDOF"], wl["bcVal"], wl["loads"])
    compare(ref, dg.digest(K, R, vm, "c2"), "c2")


@pytest.mark.parametrize("k", [0, 3, 7])
def test_oracle_reproduces_reference_digest_config5_strips(ref, k):
    n = bench.WORKLOADS["c5"][1]
    wl = _c5_workload()
    thick = dg.parity_inputs(wl)
    start, rows = dg.C5_STRIPS[k]
    s = dg.strip_of(wl, thick, start, rows)
    K, R, vm = oracle_system("quad", s["coords"], s["conn"], s["states"], s["thick"], s["bcDOF"], s["bcVal"], s["loads"])
    compare(ref, dg.strip_digest(K, R, vm, s, k, f"c5s{k}"), f"c5s{k}")
    assert n == 5000


_C5 = {}


def _c5_workload():
    if "wl" not in _C5:
        _C5["wl"] = bench.make_workload("c5")
    return _C5["wl"]


# ---- CUDA path at full size ------------------------------------------------------------------------------------------
def _plan_for(wl):
    elType = wl["elType"]
    plan = AssemblyPlan(wl["numNodes"])
    N_, dN, w = ELEMENT_FOR[elType]().tables()
    plan.addBlock(N_, dN, w, wl["conn"][elType])
    plan.finalize()
    return plan


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["c2", "c3", "c4"])
def test_cuda_path_reproduces_reference_digest(ref, cfg):
    import scipy.sparse as sp

    wl = bench.make_workload(cfg)
    thick = dg.parity_inputs(wl)
    plan = _plan_for(wl)
    mat = make_cm(0).material()
    Kd, R = plan.assemble(wl["coords"], wl["states"], thick, mat, wl["bcDOF"], wl["bcVal"], True, wl["loads"])
    indptr, indices = plan.pattern()
    K = sp.csr_array((Kd, indices, indptr), shape=(plan.numDOF, plan.numDOF))
    vm = plan.function(wl["coords"], wl["states"], mat, 1, "mean")
    worst = compare(ref, dg.digest(K, R, vm, cfg), cfg)
    print(cfg, "worst relative deviations from the reference:", {k: f"{v:.1e}" for k, v in worst.items()})
    plan.close()


@pytest.mark.gpu
def test_cuda_path_reproduces_reference_digest_config5_strips(ref):
    """config 5 (25 M elements, 900 M non-zeros) stays on the device: only the sampled rows of the eight strips come back"""
    import torch

    wl = _c5_workload()
    thick = dg.parity_inputs(wl)
    plan = _plan_for(wl)
    plan.setBCs(wl["bcDOF"], wl["bcVal"])
    mat = make_cm(0).material()
    dev = torch.device("cuda", 0)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
         (("coords", wl["coords"]), ("states", wl["states"]), ("thick", thick), ("loads", wl["loads"]))}  # fmt: skip
    d_K = torch.empty(plan.nnz, dtype=torch.float64, device=dev)
    d_R = torch.empty(plan.numDOF, dtype=torch.float64, device=dev)
    d_vm = torch.empty(plan.numElements, dtype=torch.float64, device=dev)
    d_indptr = torch.empty(plan.numDOF + 1, dtype=torch.int64, device=dev)
    d_indices = torch.empty(plan.nnz, dtype=torch.int64, device=dev)
    plan.setStream(torch.cuda.current_stream().cuda_stream)
    plan.assembleDev(d["coords"], d["states"], d["thick"], mat, True, d["loads"], d_K, d_R)
    plan.functionDev(d["coords"], d["states"], mat, 1, 2, d_vm)
    plan.patternDev(d_indptr, d_indices)
    torch.cuda.synchronize()
    assert plan.nnz == 900120004 and plan.numDOF == 50020002
    for k in range(len(dg.C5_STRIPS)):
        p = f"c5s{k}"
        rows = torch.from_numpy(ref[p + "_rows_global"]).to(dev)
        starts, ends = d_indptr[rows], d_indptr[rows + 1]
        lens = (ends - starts).cpu().numpy()
        idx = torch.cat([torch.arange(int(a), int(b), device=dev) for a, b in zip(starts.cpu().numpy(), ends.cpu().numpy())])
        cols, vals = d_indices[idx].cpu().numpy(), d_K[idx].cpu().numpy()
        seg = np.repeat(np.arange(lens.size), lens)
        kv = np.stack([np.bincount(seg, weights=vals * dg.probe_vector(j, cols), minlength=lens.size) for j in range(3)])
        got = {"rows_global": ref[p + "_rows_global"], "elems_global": ref[p + "_elems_global"], "row_len": lens.astype(np.int64),
               "cols_sha": np.array([dg.sha(cols)]), "Kv_rows": kv, "K_row_abs": np.bincount(seg, weights=np.abs(vals), minlength=lens.size),
               "R_rows": d_R[rows].cpu().numpy(), "vm_elems": d_vm[torch.from_numpy(ref[p + "_elems_global"]).to(dev)].cpu().numpy()}  # fmt: skip
        worst = compare(ref, {f"{p}_{key}": v for key, v in got.items()}, p)
        print(p, {key: f"{v:.1e}" for key, v in worst.items()})
    plan.close()
