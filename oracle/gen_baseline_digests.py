"""Digests of the BASELINE configurations 2-5 from the UNMODIFIED reference -> tests/golden/baseline_digests.npz.

TEST INFRASTRUCTURE.  Run once in the build container (needs /root/reference, ~10 GB of memory, a few minutes):
    python -m oracle.gen_baseline_digests
Every system is assembled by the reference's own FEMpyProblem._assembleMatrix / _assembleResidual (applyBCs=True,
FEMpy/Problem.py:565-661) and computeFunction("Von-Mises-Stress", "mean") (:282-355) on the bench.py workload of the
configuration (mesh, BCs, loads, states) with a random thickness per element (oracle/digests.py: parity_inputs):
  c2, c3, c4   whole mesh  -> digests.digest
  c5           8 strips of 16 full-width element rows (the whole mesh needs 51 GB of temporaries in the reference); rows of
               nodes inside a strip are complete, so they equal the rows of the whole mesh -> digests.strip_digest
"""
import contextlib
import io
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import digests as dg  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "baseline_digests.npz")


def reference_system(fp, elType, coords, conn, states, thick, bcDOF, loads):
    left = np.unique(bcDOF // 2)
    right = (loads.reshape(-1, 2)[:, 0] != 0).nonzero()[0]
    m = bench.MAT
    model, prob = rh.build_problem(fp, coords, {elType: conn}, 0, m["E"], m["nu"], m["rho"], m["t"], thickness=thick,
                                   bcs=[("Fixed", left, [0, 1], 0.0)] if left.size else None,
                                   loads=[("Load", right, [0], 1e4, False)] if right.size else None, states=states)
    K, R = rh.assemble(prob)
    with contextlib.redirect_stdout(io.StringIO()):
        vm = prob.computeFunction("Von-Mises-Stress", "mean")[elType]
    return K, R, vm


def main():
    fp = rh.import_reference()
    out = {}
    for cfg in ("c2", "c3", "c4"):
        t0 = time.time()
        wl = bench.make_workload(cfg)
        thick = dg.parity_inputs(wl)
        K, R, vm = reference_system(fp, wl["elType"], wl["coords"], wl["conn"][wl["elType"]], wl["states"], thick, wl["bcDOF"], wl["loads"])
        out.update(dg.digest(K, R, vm, cfg))
        print(cfg, K.shape, K.nnz, f"{time.time() - t0:.1f} s", flush=True)
        del K, R, vm, wl
    wl = bench.make_workload("c5")
    thick = dg.parity_inputs(wl)
    for k, (start, rows) in enumerate(dg.C5_STRIPS):
        t0 = time.time()
        s = dg.strip_of(wl, thick, start, rows)
        K, R, vm = reference_system(fp, "quad", s["coords"], s["conn"], s["states"], s["thick"], s["bcDOF"], s["loads"])
        out.update(dg.strip_digest(K, R, vm, s, k, f"c5s{k}"))
        print("c5 strip", k, start, rows, K.nnz, f"{time.time() - t0:.1f} s", flush=True)
    np.savez_compressed(OUT, **out)
    print(OUT, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
