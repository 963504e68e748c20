"""Digests of a full-size assembly: what is stored for the BASELINE configurations instead of whole matrices.

TEST INFRASTRUCTURE -- never imported by fempy_b200.  Used by oracle/gen_baseline_digests.py (run on the UNMODIFIED
reference, FEMpy/Problem.py:565-661 and :282-355) and by the tests that compare the oracle and the CUDA path with the
committed fixture tests/golden/baseline_digests.npz.

A digest of (K, R, per-element von Mises mean) holds
  * the SHA-256 of the int64 indptr / indices arrays (pattern: bit-exact),
  * K.data at SAMPLES fixed positions of the sorted CSR, max |K|, sum |K|,
  * (K @ v_k)[rows] for three fixed pseudo-random vectors v_k and ROWS fixed rows (every entry of a sampled row enters
    with a random weight: one wrong value anywhere in the row shows), the rows' sums of |K_ij|,
  * R[rows], sum R, sum |R|,
  * the von Mises mean of ELEMS fixed elements and its sum over all elements.
All index samples come from numpy's PCG64 `default_rng(seed)` (stable by NEP 19), the vectors from a closed formula of
the global DOF index so that a strip of a mesh can evaluate its own part.
"""
import hashlib

import numpy as np

SAMPLES, ROWS, ELEMS = 16384, 8192, 8192


def probe_vector(k, dofIndex):
    """v_k at the global DOF indices `dofIndex` (int64 array): values in [-1, 1], no structure the mesh could align with."""
    g = np.asarray(dofIndex, dtype=np.float64)
    return np.sin(0.7548776662466927 * g + 1.3 * (k + 1)) * np.cos(0.5698402909980532 * g + 0.31 * k)


def sample_indices(numDOF, nnz, numElems):
    rng = np.random.default_rng(20261020)
    pos = np.unique(rng.integers(0, nnz, min(SAMPLES, nnz)))
    rows = np.unique(rng.integers(0, numDOF, min(ROWS, numDOF)))
    elems = np.unique(rng.integers(0, numElems, min(ELEMS, numElems)))
    return pos, rows, elems


def sha(a, dtype=np.int64):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=dtype).tobytes()).hexdigest()


def row_products(indptr, indices, data, rows, vectors):
    """(K @ v)[rows] and sum_j |K_ij| for the given rows, from CSR arrays (numpy; the rows are gathered, K is not copied)."""
    indptr = np.asarray(indptr)
    starts, ends = indptr[rows], indptr[rows + 1]
    lens = ends - starts
    seg = np.repeat(np.arange(rows.size), lens)
    idx = np.repeat(starts - np.concatenate(([0], np.cumsum(lens)[:-1])), lens) + np.arange(lens.sum())
    cols, vals = np.asarray(indices)[idx], np.asarray(data)[idx]
    out = [np.bincount(seg, weights=vals * v(cols), minlength=rows.size) for v in vectors]
    return np.stack(out), np.bincount(seg, weights=np.abs(vals), minlength=rows.size)


def digest(K, R, vm, prefix):
    """dict of arrays (keys prefixed) for one assembled system: scipy CSR `K` with sorted indices, residual `R`, per-element
    von Mises mean `vm` (or None)."""
    numDOF, nnz = K.shape[0], K.nnz
    numElems = vm.size if vm is not None else 1
    pos, rows, elems = sample_indices(numDOF, nnz, numElems)
    kv, rabs = row_products(K.indptr, K.indices, K.data, rows, [lambda c, k=k: probe_vector(k, c) for k in range(3)])
    d = {
        "shape": np.array([numDOF, nnz, numElems], dtype=np.int64),
        "pattern_sha": np.array([sha(K.indptr), sha(K.indices)]),
        "K_samples": K.data[pos], "K_max": np.abs(K.data).max(), "K_abs_sum": np.abs(K.data).sum(),
        "Kv_rows": kv, "K_row_abs": rabs, "R_rows": R[rows], "R_sum": R.sum(), "R_abs_sum": np.abs(R).sum(),
    }  # fmt: skip
    if vm is not None:
        d["vm_elems"] = vm[elems]
        d["vm_sum"] = vm.sum()
    return {f"{prefix}_{k}": np.asarray(v) for k, v in d.items()}


def parity_inputs(wl, seed=1):
    """The inputs the digests are taken on: a bench.py workload (mesh, BCs, loads, states = 1e-3 rng(0)) with a random
    thickness per element (mirrors Examples/LBracketExample.py:98-100) so that every element weighs differently."""
    rng = np.random.default_rng(seed)
    return rng.random(wl["numElems"]) * 1e-2 + 1e-3


# ---- strips of config 5 (the reference cannot hold the whole mesh: 51 GB of temporaries) --------------------------------
C5_STRIPS = [(0, 16), (613, 16), (1250, 16), (2499, 16), (3333, 16), (4100, 16), (4700, 16), (4984, 16)]  # (first element row, rows)
STRIP_ROWS, STRIP_ELEMS = 2048, 2048


def strip_of(wl, thick, startRow, rows):
    """Element rows [startRow, startRow + rows) of a structured Q4 bench.py workload as a mesh of its own (node ids shifted
    by n0 = startRow * (n + 1)): dict of inputs + `n0`, `e0` and `interior`, the local ids of the nodes whose rows are
    complete in the strip (strictly inside it, or on the boundary of the whole mesh)."""
    n = wl["n"]
    assert wl["elType"] == "quad"
    e0, e1 = n * startRow, n * (startRow + rows)
    conn = wl["conn"]["quad"][e0:e1]
    n0, n1 = startRow * (n + 1), (startRow + rows + 1) * (n + 1)
    keep = (wl["bcDOF"] >= 2 * n0) & (wl["bcDOF"] < 2 * n1)
    first = 0 if startRow == 0 else 1
    last = rows + 1 if startRow + rows == n else rows
    return dict(conn=conn - n0, coords=wl["coords"][n0:n1], states=wl["states"][n0:n1], thick=thick[e0:e1],
                bcDOF=wl["bcDOF"][keep] - 2 * n0, bcVal=wl["bcVal"][keep], loads=wl["loads"][2 * n0 : 2 * n1], n0=n0, e0=e0,
                interior=np.arange(first * (n + 1), last * (n + 1)))  # fmt: skip


def strip_samples(strip, k):
    """Fixed sample of complete DOF rows (local ids) and of elements (local ids) of strip number k."""
    rng = np.random.default_rng(977 + k)
    dof = np.stack((2 * strip["interior"], 2 * strip["interior"] + 1), axis=1).reshape(-1)
    rows = np.unique(rng.choice(dof, min(STRIP_ROWS, dof.size), replace=False))
    elems = np.unique(rng.integers(0, strip["conn"].shape[0], STRIP_ELEMS))
    return rows, elems


def strip_digest(K, R, vm, strip, k, prefix):
    """Digest of the complete rows of a strip system (K scipy CSR over the strip's local DOFs): columns are shifted to
    global DOF ids before they are hashed / multiplied, so the same numbers come out of the rows of the whole mesh."""
    rows, elems = strip_samples(strip, k)
    shift = 2 * strip["n0"]
    kv, rabs = row_products(K.indptr, K.indices, K.data, rows, [lambda c, j=j: probe_vector(j, c + shift) for j in range(3)])
    starts, ends = K.indptr[rows], K.indptr[rows + 1]
    cols = np.concatenate([K.indices[a:b] for a, b in zip(starts, ends)]) + shift
    d = {"rows_global": rows + shift, "elems_global": elems + strip["e0"], "row_len": (ends - starts).astype(np.int64),
         "cols_sha": np.array([sha(cols)]), "Kv_rows": kv, "K_row_abs": rabs, "R_rows": R[rows], "vm_elems": vm[elems]}  # fmt: skip
    return {f"{prefix}_{key}": np.asarray(v) for key, v in d.items()}
