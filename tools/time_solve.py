"""Device-resident solve (femb_assemble_solve) on a BASELINE config-2 style problem: iterations, time per iteration,
achieved bandwidth of the iteration, and the host direct solve the reference would run (scipy factorized) beside it.
usage: python tools/time_solve.py [nx] [--no-host]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fempy_b200 import Constitutive, meshes  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from hostmirror import FEMpyModel  # noqa: E402  (test-support stand-in for the reference model)

nx = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 512
coords, conn = meshes.quad4_grid(nx, nx)
model = FEMpyModel(Constitutive.IsoPlaneStress(70e9, 0.3, 2700.0, 1.0), nodeCoords=coords, connectivity=conn)
left, right = meshes.edge_nodes(coords, x=0.0), meshes.edge_nodes(coords, x=2.0)
model.addFixedBCToNodes(name="Fixed", nodeInds=[int(i) for i in left], dof=[0, 1], value=0.0)
prob = model.addProblem("scaling")
prob.addLoadToNodes(name="Load", nodeInds=[int(i) for i in right], dof=0, value=1e4)
prob.solveOnDevice(relTol=1e-2, maxIter=50)  # plan build + warm-up
prob.reset()
t0 = time.perf_counter()
info = prob.solveOnDevice(relTol=1e-10, maxIter=200000)
t1 = time.perf_counter()
uDev = prob.states.copy()
from fempy_b200.problem import planFor  # noqa: E402

nnz, numDOF = planFor(model).nnz, prob.numDOF
bytesPerIt = 8 * nnz + nnz + 9 * 8 * numDOF  # K values + node columns (4 B per 2x2 block) + 9 vector passes
print(f"quad {nx}x{nx}: {numDOF} DOF, nnz {nnz}; device solve (H2D inputs + assembly + PCG + D2H update) {1e3 * (t1 - t0):.1f} ms, "
      f"{info['iterations']} iterations, rel. residual {info['relativeResidual']:.2e}, "
      f"{1e6 * (t1 - t0) / max(info['iterations'], 1):.1f} us per iteration incl. everything "
      f"-> >= {bytesPerIt / ((t1 - t0) / max(info['iterations'], 1)) / 1e9:.0f} GB/s of algorithmic traffic")
if "--no-host" not in sys.argv:
    prob.reset()
    t0 = time.perf_counter()
    prob.solve()
    t1 = time.perf_counter()
    err = np.abs(prob.states - uDev).max() / np.abs(prob.states).max()
    print(f"host path (assembly on the GPU, D2H of K, csr_array, scipy factorized + solve): {1e3 * (t1 - t0):.1f} ms; "
          f"max |u_device - u_host| / max |u| = {err:.2e}")
