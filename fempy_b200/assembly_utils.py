"""Mirror of the reference's assembly utilities (reference FEMpy/Utils/AssemblyUtils.py), same names
and semantics.  The two O(E) routines -- `localMatricesToCOOArrays` and `scatterLocalResiduals` --
run on the GPU through the C ABI; the dictionary converters and the O(nBC) helpers are host logic.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_f64, as_i64, check, dptr, iptr


def convertBCDictToLists(bcDict):
    """AssemblyUtils.py:97-137."""
    bcDOF, bcValues = [], []
    for key in bcDict:
        bcDOF += bcDict[key]["DOF"]
        bcValues += bcDict[key]["Value"]
    return bcDOF, bcValues


def convertLoadsDictToVector(loadsDict, numDOF):
    """AssemblyUtils.py:140-165 (fancy-index `+=`: duplicates inside one entry do not accumulate)."""
    loadForce = np.zeros(numDOF)
    for key in loadsDict:
        loadForce[loadsDict[key]["DOF"]] += loadsDict[key]["Value"]
    return loadForce


def applyBCsToMatrix(rows, cols, values, bcDOF):
    """AssemblyUtils.py:26-62 on COO triplets (host; the fused CUDA assembly applies BCs in-kernel)."""
    values[np.nonzero(np.isin(rows, bcDOF))[0]] = 0.0
    rows = np.append(rows, bcDOF)
    cols = np.append(cols, bcDOF)
    values = np.append(values, np.ones(len(bcDOF)))
    return rows, cols, values


def applyBCsToVector(vector, state, bcDOF, bcValues):
    """AssemblyUtils.py:65-94."""
    vector[bcDOF] = state[bcDOF] - bcValues
    return vector


def localMatricesToCOOArrays(localMats, localDOF, device=0):
    """AssemblyUtils.py:168-214: element-major, row-major COO with exactly-zero values dropped."""
    lib = _lib.load()
    dtype = np.asarray(localMats).dtype
    mats, dof = as_f64(localMats), as_i64(localDOF)
    numElements, nd = dof.shape
    if mats.shape != (numElements, nd, nd):
        raise ValueError("localMats must be numElements x n x n and localDOF numElements x n")
    cap = numElements * nd * nd
    rows, cols, vals = np.empty(cap, dtype=np.int64), np.empty(cap, dtype=np.int64), np.empty(cap)
    n = C.c_int64(0)
    check(lib.femb_local_matrices_to_coo(dptr(mats), iptr(dof), numElements, nd, device, iptr(rows), iptr(cols), dptr(vals), C.byref(n)))
    vals = vals[: n.value]
    if np.issubdtype(dtype, np.integer):
        vals = vals.astype(dtype)
    return rows[: n.value].copy(), cols[: n.value].copy(), vals.copy()


def scatterLocalResiduals(localResiduals, connectivity, globalResidual, device=0):
    """AssemblyUtils.py:217-238: globalResidual[node*s + st] += localResiduals[e, n, st], in place."""
    lib = _lib.load()
    res, conn = as_f64(localResiduals), as_i64(connectivity)
    numElements, nn, s = res.shape
    if globalResidual.dtype != np.float64 or not globalResidual.flags.c_contiguous:
        raise ValueError("globalResidual must be a contiguous float64 array")
    if globalResidual.size % s:
        raise ValueError("globalResidual length must be a multiple of numStates")
    check(lib.femb_scatter_local_residuals(dptr(res), iptr(conn), numElements, nn, s, globalResidual.size // s, device, dptr(globalResidual)))
