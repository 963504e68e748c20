"""TEST / BENCH INFRASTRUCTURE ONLY -- never imported by fempy_b200.

Process-parallel driver of the numpy oracle (oracle/fem_oracle.py) for bench.py's CPU legs (`cpu_baseline`,
`--impl reference`): every worker process assembles K (COO -> CSR, FEMpy/Problem.py:565-615) and R (:617-661) of its own
strip of element rows, so the leg uses all host cores the way the reference's Numba kernels would use their threads
(the reference's Python gathers and scipy conversion are serial; a process per strip is the generous reading).
"""
import time

import numpy as np


def strip_inputs(wl, startRow, rows, mat):
    """Sub-mesh of element rows [startRow, startRow + rows) of a structured workload, renumbered from node 0."""
    elType, n = wl["elType"], wl["n"]
    conn = wl["conn"][elType]
    if elType == "triangle":  # all lower triangles first, then all upper ones
        half = conn.shape[0] // 2
        per = half // n
        sub = np.vstack((conn[per * startRow : per * (startRow + rows)], conn[half + per * startRow : half + per * (startRow + rows)]))
    else:
        per = conn.shape[0] // n
        sub = conn[per * startRow : per * (startRow + rows)]
    n0, n1 = int(sub.min()), int(sub.max()) + 1
    keep = (wl["bcDOF"] >= 2 * n0) & (wl["bcDOF"] < 2 * n1)
    return dict(elType=elType, conn=sub - n0, coords=wl["coords"][n0:n1], states=wl["states"][n0:n1],
                thick=np.full(sub.shape[0], mat["t"]), bcDOF=wl["bcDOF"][keep] - 2 * n0, bcVal=wl["bcVal"][keep],
                loads=wl["loads"][2 * n0 : 2 * n1], E=mat["E"], nu=mat["nu"])  # fmt: skip


def assemble_strip(s):
    """One K + R assembly of a strip with the oracle's stages; returns (elements, seconds, start, end)."""
    import os
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import fem_oracle as fo
    from util import oracle_blocks

    blocks = oracle_blocks({s["elType"]: s["conn"]})
    t0 = time.perf_counter()
    K = fo.assemble_matrix(blocks, s["coords"], s["states"], s["thick"], 0, s["E"], s["nu"], s["bcDOF"])
    R = fo.assemble_residual(blocks, s["coords"], s["states"], s["thick"], 0, s["E"], s["nu"], s["loads"], s["bcDOF"], s["bcVal"])
    t1 = time.perf_counter()
    assert K.nnz > 0 and R.size == 2 * s["coords"].shape[0]
    return s["conn"].shape[0], t1 - t0, time.time() - (t1 - t0), time.time()


class StripPool:
    """`procs` worker processes (spawned: safe next to an initialised CUDA context), one strip each per round."""

    def __init__(self, procs):
        import multiprocessing as mp

        self.procs = procs
        self.pool = mp.get_context("spawn").Pool(procs) if procs > 1 else None

    def run(self, strips):
        """Assemble all strips concurrently; returns (total elements, wall seconds from first start to last end)."""
        if self.pool is None:
            out = [assemble_strip(s) for s in strips]
            return sum(o[0] for o in out), sum(o[1] for o in out)
        out = self.pool.map(assemble_strip, strips, chunksize=1)
        return sum(o[0] for o in out), max(o[3] for o in out) - min(o[2] for o in out)

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
