"""AssemblyPlan: Python handle on a `femb_plan` (mesh topology + element tables on the GPU).

Built once per model; owns the CSR pattern and the element scatter map and runs the fused assembly,
element-function and element-block kernels through the C ABI (include/femb200.h).
"""
import ctypes as C

import numpy as np

from . import _lib, _pinned
from ._lib import as_f64, as_i64, check, dptr, iptr

WANT_K, WANT_R = 1, 2
# "ksmin" is not a typo: the reference evaluates the KS *max* for both names (Element.py:261-265)
REDUCTIONS = {None: 0, "none": 0, "sum": 1, "mean": 2, "integrate": 3, "min": 4, "max": 5, "ksmax": 6, "ksmin": 6}


class AssemblyPlan:
    def __init__(self, numNodes, numDim=2, numStates=2, device=0):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        check(self._lib.femb_plan_create(int(numNodes), int(numDim), int(numStates), int(device), C.byref(self._h)))
        self.numNodes = int(numNodes)
        self.numDim = int(numDim)
        self.numStates = int(numStates)
        self.device = int(device)
        self.blocks = []  # (nodesPerElem, numQuad, numElems)
        self.finalized = False
        self._pattern = None

    # -- construction ---------------------------------------------------------------------------
    def addBlock(self, N, dNdxi, w, conn):
        """One element type: N (P,n), dNdxi (P,d,n), w (P,), conn (E,n) int64."""
        N, dNdxi, w, conn = as_f64(N), as_f64(dNdxi), as_f64(w), as_i64(conn)
        P, n = N.shape
        if dNdxi.shape != (P, self.numDim, n) or w.shape != (P,) or conn.ndim != 2 or conn.shape[1] != n:
            raise ValueError("inconsistent element table / connectivity shapes")
        check(self._lib.femb_plan_add_block(self._h, n, P, dptr(N), dptr(dNdxi), dptr(w), conn.shape[0], iptr(conn)))
        self.blocks.append((n, P, conn.shape[0]))

    def addBlockDev(self, N, dNdxi, w, d_conn, numElems):
        """addBlock with the (numElems, n) int64 connectivity already on the plan's device (pointer or object with
        .data_ptr()): device-side mesh ingest."""
        N, dNdxi, w = as_f64(N), as_f64(dNdxi), as_f64(w)
        P, n = N.shape
        if dNdxi.shape != (P, self.numDim, n) or w.shape != (P,):
            raise ValueError("inconsistent element table shapes")
        check(self._lib.femb_plan_add_block_dev(self._h, n, P, dptr(N), dptr(dNdxi), dptr(w), int(numElems), _p(d_conn)))
        self.blocks.append((n, P, int(numElems)))

    def finalize(self):
        check(self._lib.femb_plan_finalize(self._h))
        self.finalized = True
        return self

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.femb_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- queries --------------------------------------------------------------------------------
    @property
    def handle(self):
        return self._h

    @property
    def numDOF(self):
        return int(self._lib.femb_plan_num_dof(self._h))

    @property
    def numElements(self):
        return int(self._lib.femb_plan_num_elems(self._h))

    @property
    def numPoints(self):
        return sum(P * E for _, P, E in self.blocks)

    @property
    def nnz(self):
        return int(self._lib.femb_plan_nnz(self._h))

    @property
    def buildMs(self):
        return float(self._lib.femb_plan_build_ms(self._h))

    @property
    def lastLaunches(self):
        return int(self._lib.femb_plan_last_launches(self._h))

    def setStream(self, cudaStream):
        """Run on the caller's CUDA stream handle (int; 0 = legacy default stream); None = the plan's own stream."""
        if cudaStream is None:
            check(self._lib.femb_plan_set_stream(self._h, None, 0))
        else:
            check(self._lib.femb_plan_set_stream(self._h, C.c_void_p(int(cudaStream)), 1))

    def setOption(self, name, value):
        check(self._lib.femb_plan_set_option(self._h, name.encode(), int(value)))

    def info(self, name):
        return int(self._lib.femb_plan_get_info(self._h, name.encode()))

    def sync(self):
        check(self._lib.femb_plan_sync(self._h))

    def pattern(self):
        """(indptr, indices) int64 of the CSR stiffness pattern; cached."""
        if self._pattern is None:
            if self.nnz < 0:  # not finalized: let the library say so
                check(self._lib.femb_plan_pattern(self._h, None, None))
            indptr = np.empty(self.numDOF + 1, dtype=np.int64)
            indices = np.empty(self.nnz, dtype=np.int64)
            check(self._lib.femb_plan_pattern(self._h, iptr(indptr), iptr(indices)))
            self._pattern = (indptr, indices)
        return self._pattern

    # -- device-side model -------------------------------------------------------------------------
    def setGeometry(self, coords, thickness):
        """Keep node coordinates / per-element thickness on the device (femb_plan_set_geometry): the host-pointer calls
        below then take None for them and upload only the states (and loads).  None drops the resident copy."""
        c = as_f64(coords) if coords is not None else None
        t = as_f64(thickness) if thickness is not None else None
        if (c is not None and c.shape != (self.numNodes, self.numDim)) or (t is not None and t.shape != (self.numElements,)):
            raise ValueError("coords/thickness have the wrong shape for this plan")
        check(self._lib.femb_plan_set_geometry(self._h, dptr(c) if c is not None else None, dptr(t) if t is not None else None))

    # -- host-pointer compute calls -----------------------------------------------------------------
    def assemble(self, coords, states, thickness, material, bcDOF=None, bcValues=None, applyBCs=True, loads=None,
                 wantK=True, wantR=True, Kout=None, Rout=None):
        """Kout / Rout: optional preallocated float64 outputs (e.g. pinned host memory).  coords / thickness None: use
        the resident copies (setGeometry)."""
        coords = as_f64(coords) if coords is not None else None
        thickness = as_f64(thickness) if thickness is not None else None
        states = as_f64(states)
        if ((coords is not None and coords.shape != (self.numNodes, self.numDim)) or states.size != self.numDOF
                or (thickness is not None and thickness.shape != (self.numElements,))):
            raise ValueError("coords/states/thickness have the wrong shape for this plan")
        nBC = 0
        bcd = bcv = None
        if applyBCs and bcDOF is not None and len(bcDOF):
            bcd, bcv = as_i64(bcDOF), as_f64(bcValues)
            if bcd.shape != bcv.shape:
                raise ValueError("bcDOF and bcValues must have the same length")
            nBC = bcd.size
        lv = None
        if loads is not None:
            lv = as_f64(loads)
            if lv.shape != (self.numDOF,):
                raise ValueError("loads must have numDOF entries")
        # fresh result arrays come from the page-locked pool (_pinned.py): same value semantics, PCIe-speed copy back
        pinned = _pinned.pool(self._lib)
        K = (Kout if Kout is not None else pinned.empty(self.nnz)) if wantK else None
        R = (Rout if Rout is not None else pinned.empty(self.numDOF)) if wantR else None
        for out, n in ((K, self.nnz), (R, self.numDOF)):
            if out is not None and (out.dtype != np.float64 or out.shape != (n,) or not out.flags.c_contiguous):
                raise ValueError("output buffers must be contiguous float64 arrays of length nnz / numDOF")
        want = (WANT_K if wantK else 0) | (WANT_R if wantR else 0)
        check(
            self._lib.femb_assemble(
                self._h, dptr(coords) if coords is not None else None, dptr(states),
                dptr(thickness) if thickness is not None else None, C.byref(material),
                iptr(bcd) if nBC else None, dptr(bcv) if nBC else None, nBC, int(bool(applyBCs)),
                dptr(lv) if lv is not None else None, want,
                dptr(K) if wantK else None, dptr(R) if wantR else None,
            )
        )  # fmt: skip
        return K, R

    def function(self, coords, states, material, funcId, reduction):
        if isinstance(reduction, str) or reduction is None:
            key = reduction.lower() if isinstance(reduction, str) else None
            assert key in REDUCTIONS, "elementReductionType not valid"
            reduction = REDUCTIONS[key]
        coords, states = (as_f64(coords) if coords is not None else None), as_f64(states)
        out = np.empty(self.numPoints if reduction == 0 else self.numElements)
        check(self._lib.femb_function(self._h, dptr(coords) if coords is not None else None, dptr(states), C.byref(material),
                                      int(funcId), int(reduction), dptr(out)))
        return out

    def pointQuantities(self, coords, states, want=("Coord", "State", "StateGrad", "JacDet")):
        """Element._computeFunctionEvaluationQuantities on the device (femb_point_quantities) at the points the blocks were
        added with: dict of (numPoints, ...) arrays for the names in `want` out of Coord, State, StateGrad, Jac, JacDet,
        StateGradSens (single-block plans: (numPoints, numDim, nodesPerElement))."""
        coords, states = (as_f64(coords) if coords is not None else None), as_f64(states)
        P = self.numPoints
        shapes = {"Coord": (P, 2), "State": (P, 2), "StateGrad": (P, 2, 2), "Jac": (P, 2, 2), "JacDet": (P,)}
        if "StateGradSens" in want:
            if len(self.blocks) != 1:
                raise ValueError("StateGradSens needs a single-block plan")
            shapes["StateGradSens"] = (P, 2, self.blocks[0][0])
        out = {k: np.empty(shapes[k]) for k in want}
        args = [dptr(out[k]) if k in out else None for k in ("Coord", "State", "StateGrad", "Jac", "JacDet", "StateGradSens")]
        check(self._lib.femb_point_quantities(self._h, dptr(coords) if coords is not None else None, dptr(states), *args))
        return out

    def elementBlocks(self, coords, states, thickness, material, wantK=True, wantR=True):
        coords, states, thickness = as_f64(coords), as_f64(states), as_f64(thickness)
        s = self.numStates
        Ke = np.empty(sum(E * (s * n) ** 2 for n, _, E in self.blocks)) if wantK else None
        Re = np.empty(sum(E * n * s for n, _, E in self.blocks)) if wantR else None
        check(
            self._lib.femb_element_blocks(
                self._h, dptr(coords), dptr(states), dptr(thickness), C.byref(material),
                dptr(Ke) if wantK else None, dptr(Re) if wantR else None,
            )
        )  # fmt: skip
        return Ke, Re

    # -- device-pointer compute calls (torch tensors or raw pointers) ------------------------------
    def setBCs(self, bcDOF, bcValues):
        bcd, bcv = as_i64(bcDOF), as_f64(bcValues)
        check(self._lib.femb_plan_set_bcs(self._h, iptr(bcd) if bcd.size else None, dptr(bcv) if bcd.size else None, bcd.size))

    def assembleDev(self, d_coords, d_states, d_thickness, material, applyBCs, d_loads, d_K, d_R):
        """All arguments are device pointers (int) or objects with .data_ptr(); asynchronous."""
        want = (WANT_K if d_K is not None else 0) | (WANT_R if d_R is not None else 0)
        check(
            self._lib.femb_assemble_dev(
                self._h, _p(d_coords), _p(d_states), _p(d_thickness), C.byref(material), int(bool(applyBCs)),
                _p(d_loads), want, _p(d_K), _p(d_R),
            )
        )  # fmt: skip

    def functionDev(self, d_coords, d_states, material, funcId, reduction, d_out):
        check(self._lib.femb_function_dev(self._h, _p(d_coords), _p(d_states), C.byref(material), int(funcId), int(reduction), _p(d_out)))

    def solvePCG(self, K, rhs, relTol=1e-10, maxIter=20000):
        """K x = rhs on the GPU (femb_solve_pcg): K = the CSR values assembled with applyBCs and the plan's current BCs.
        Returns (x, iterations, relative residual)."""
        Kv, b = as_f64(K), as_f64(rhs)
        x = np.empty(self.numDOF)
        it, res = C.c_int32(0), C.c_double(0.0)
        check(self._lib.femb_solve_pcg(self._h, dptr(Kv), dptr(b), dptr(x), float(relTol), int(maxIter), C.byref(it), C.byref(res)))
        return x, it.value, res.value

    def assembleSolve(self, coords, states, thickness, material, bcDOF, bcValues, loads=None, relTol=1e-10, maxIter=20000):
        """One fused assembly (BCs applied) + update = K^-1 (-R) on the GPU (femb_assemble_solve); the CSR values stay on
        the device.  Returns (update, R, iterations, relative residual)."""
        c, q, t = (as_f64(coords) if coords is not None else None), as_f64(states), (as_f64(thickness) if thickness is not None else None)
        bcd, bcv = as_i64(bcDOF if bcDOF is not None else []), as_f64(bcValues if bcValues is not None else [])
        ld = as_f64(loads) if loads is not None else None
        update, R = np.empty(self.numDOF), np.empty(self.numDOF)
        it, res = C.c_int32(0), C.c_double(0.0)
        check(self._lib.femb_assemble_solve(
            self._h, dptr(c) if c is not None else None, dptr(q), dptr(t) if t is not None else None, C.byref(material),
            iptr(bcd) if bcd.size else None, dptr(bcv) if bcd.size else None,
            bcd.size, dptr(ld) if ld is not None else None, float(relTol), int(maxIter), dptr(update), dptr(R), C.byref(it), C.byref(res)))  # fmt: skip
        return update, R, it.value, res.value

    def solvePCGDev(self, d_K, d_rhs, d_x, relTol=1e-10, maxIter=20000):
        """Device pointers; synchronises the plan's stream every 25 iterations.  Returns (iterations, relative residual)."""
        it, res = C.c_int32(0), C.c_double(0.0)
        check(self._lib.femb_solve_pcg_dev(self._h, _p(d_K), _p(d_rhs), _p(d_x), float(relTol), int(maxIter), C.byref(it), C.byref(res)))
        return it.value, res.value

    def profile(self, enable=True):
        check(self._lib.femb_plan_profile(self._h, int(bool(enable))))

    def profileRead(self):
        """(summed kernel milliseconds, number of timed assembly calls) since the last read."""
        ms, n = C.c_double(0.0), C.c_int64(0)
        check(self._lib.femb_plan_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def patternDev(self, d_indptr, d_indices):
        check(self._lib.femb_plan_pattern_dev(self._h, _p(d_indptr), _p(d_indices)))


def _p(x):
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return C.c_void_p(int(x))
