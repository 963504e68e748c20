"""CPU oracle (numpy) for the FEMpy assembly hot path.  TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module.  The shipped package (``fempy_b200``) never does: its hot path is the
CUDA library and fails loudly without it.

This is a plain restatement of the reference's algorithm, stage by stage, with the same tensor
contractions the reference evaluates (general strain-sensitivity form, so the non-linear Green strain
follows the same code), vectorised over elements with numpy and processed in element chunks:

  stage                                   reference (relative to /root/reference)
  --------------------------------------  ---------------------------------------------------------
  J = dNdxi @ X_e, detJ, J^-1             FEMpy/Elements/Element.py:331-341,942-978; LinAlg/LinAlg.py:70-189
  G = J^-1 dNdxi,  u' = (G q_e)^T         FEMpy/Elements/Element.py:344-353,981-1048
  strains / strain sensitivities          FEMpy/Constitutive/StrainModels/ContinuumStrains.py:90-174,180-328
  D matrices, stress = eps @ D            FEMpy/Constitutive/StressModels/IsoStress.py:25-78,157-298
  weak residual  theta * S^T sigma        FEMpy/Constitutive/ConstitutiveModel.py:240-282,345-386
  weak Jacobian  theta * S^T D S          FEMpy/Constitutive/ConstitutiveModel.py:286-338,389-460
  r = G^T weakRes, dr/dq = G^T wJ G       FEMpy/Elements/Element.py:413-414,477-478,1051-1144
  integrate  sum_p (.) detJ_p w_p         FEMpy/Elements/Element.py:422-423,483-484,1147-1168
  local blocks -> COO (zero entries dropped)   FEMpy/Utils/AssemblyUtils.py:168-214
  residual scatter                        FEMpy/Utils/AssemblyUtils.py:217-238
  BCs on matrix / vector, loads           FEMpy/Utils/AssemblyUtils.py:26-165
  COO -> CSR (scipy, duplicates summed)   FEMpy/Problem.py:565-615
  residual driver                         FEMpy/Problem.py:617-661
  element functions + reductions          FEMpy/Elements/Element.py:163-273; Constitutive/IsoPlaneStress.py:187-227;
                                          Constitutive/StressModels/VonMises.py:26-89

Parity is PINNED: ``tests/test_oracle_vs_reference.py`` checks this module (a) against the reference's
own golden vectors (tests/testAssembleutils.py:31-81, tests/testElements.py:34-47,
tests/testProblem.py:67-143, tests/testModel.py:99-104) and (b) against outputs of the unmodified
reference run in the build container, stored under ``tests/golden/`` by ``oracle/gen_golden.py``.
"""
import numpy as np
from scipy.sparse import coo_array, csr_array

PLANE_STRESS, PLANE_STRAIN, SOLID_3D, BAR_1D = 0, 1, 2, 3

CHUNK = 32768


# ---------------------------------------------------------------------------------------------
# constitutive pieces
# ---------------------------------------------------------------------------------------------
def d_matrix(model, E, nu):
    """Stress-strain matrix (reference IsoStress.py:25-78 for 2-D, :100-150 for 3-D)."""
    if model == PLANE_STRESS:
        return E / (1.0 - nu**2) * np.array([[1.0, nu, 0.0], [nu, 1.0, 0.0], [0.0, 0.0, (1.0 - nu) / 2.0]])
    if model == PLANE_STRAIN:
        return (
            E
            / ((1.0 + nu) * (1.0 - 2.0 * nu))
            * np.array([[1.0 - nu, nu, 0.0], [nu, 1.0 - nu, 0.0], [0.0, 0.0, (1.0 - 2.0 * nu) / 2.0]])
        )
    if model == SOLID_3D:
        D = np.zeros((6, 6))
        D[:3, :3] = nu
        D[[0, 1, 2], [0, 1, 2]] = 1.0 - nu
        D[[3, 4, 5], [3, 4, 5]] = (1.0 - 2.0 * nu) / 2.0
        return E / ((1.0 + nu) * (1.0 - 2.0 * nu)) * D
    if model == BAR_1D:
        return np.array([[E]])
    raise ValueError(f"unknown constitutive model {model}")


def strains(uprime, nonlinear=False):
    """Engineering strains from u'[p, s, d] (reference ContinuumStrains.py:29-85,90-128,180-250)."""
    d = uprime.shape[-1]
    if d == 1:
        e = uprime[:, 0, 0:1].copy()
        if nonlinear:
            e += 0.5 * uprime[:, 0, 0:1] ** 2
        return e
    if d == 2:
        e = np.stack((uprime[:, 0, 0], uprime[:, 1, 1], uprime[:, 0, 1] + uprime[:, 1, 0]), axis=1)
        if nonlinear:
            e[:, 0] += 0.5 * (uprime[:, 0, 0] ** 2 + uprime[:, 1, 0] ** 2)
            e[:, 1] += 0.5 * (uprime[:, 0, 1] ** 2 + uprime[:, 1, 1] ** 2)
            e[:, 2] += uprime[:, 0, 0] * uprime[:, 0, 1] + uprime[:, 1, 0] * uprime[:, 1, 1]
        return e
    # 3-D: [e_xx, e_yy, e_zz, gamma_yz, gamma_xz, gamma_xy]
    u = uprime
    e = np.stack(
        (u[:, 0, 0], u[:, 1, 1], u[:, 2, 2], u[:, 1, 2] + u[:, 2, 1], u[:, 0, 2] + u[:, 2, 0], u[:, 0, 1] + u[:, 1, 0]),
        axis=1,
    )
    if nonlinear:
        for i in range(3):
            e[:, i] += 0.5 * (u[:, 0, i] ** 2 + u[:, 1, i] ** 2 + u[:, 2, i] ** 2)
        for k, (i, j) in zip((3, 4, 5), ((1, 2), (0, 2), (0, 1))):
            e[:, k] += u[:, 0, i] * u[:, 0, j] + u[:, 1, i] * u[:, 1, j] + u[:, 2, i] * u[:, 2, j]
    return e


def strain_sens(uprime, nonlinear=False):
    """d strain / d u' -> sens[p, e, s, d] (reference ContinuumStrains.py:55-85,131-174,253-328)."""
    P, _, d = uprime.shape
    if d == 1:
        S = np.ones((P, 1, 1, 1))
        if nonlinear:
            S[:, 0, 0, 0] += uprime[:, 0, 0]
        return S
    if d == 2:
        S = np.zeros((P, 3, 2, 2))
        S[:, 0, 0, 0] = 1.0
        S[:, 1, 1, 1] = 1.0
        S[:, 2, 0, 1] = 1.0
        S[:, 2, 1, 0] = 1.0
        if nonlinear:
            S[:, 0, 0, 0] += uprime[:, 0, 0]
            S[:, 0, 1, 0] += uprime[:, 1, 0]
            S[:, 1, 0, 1] += uprime[:, 0, 1]
            S[:, 1, 1, 1] += uprime[:, 1, 1]
            S[:, 2, 0, 0] += uprime[:, 0, 1]
            S[:, 2, 0, 1] += uprime[:, 0, 0]
            S[:, 2, 1, 0] += uprime[:, 1, 1]
            S[:, 2, 1, 1] += uprime[:, 1, 0]
        return S
    S = np.zeros((P, 6, 3, 3))
    for i in range(3):
        S[:, i, i, i] = 1.0
    for k, (i, j) in zip((3, 4, 5), ((1, 2), (0, 2), (0, 1))):
        S[:, k, i, j] = 1.0
        S[:, k, j, i] = 1.0
    if nonlinear:
        for i in range(3):
            for s in range(3):
                S[:, i, s, i] += uprime[:, s, i]
        for k, (i, j) in zip((3, 4, 5), ((1, 2), (0, 2), (0, 1))):
            for s in range(3):
                S[:, k, s, i] += uprime[:, s, j]
                S[:, k, s, j] += uprime[:, s, i]
    return S


def von_mises(stress, model, nu):
    """Von Mises stress (reference VonMises.py:26-89)."""
    if model == PLANE_STRESS:
        s11, s22, s12 = stress[:, 0], stress[:, 1], stress[:, 2]
        return np.sqrt(s11**2 - s11 * s22 + s22**2 + 3 * s12**2)
    if model == PLANE_STRAIN:
        s11, s22, s12 = stress[:, 0], stress[:, 1], stress[:, 2]
        s33 = nu * (s11 + s22)
        return np.sqrt(0.5 * ((s11 - s22) ** 2 + (s22 - s33) ** 2 + (s33 - s11) ** 2) + 3 * s12**2)
    if model == SOLID_3D:
        s = stress
        return np.sqrt(
            0.5 * ((s[:, 0] - s[:, 1]) ** 2 + (s[:, 1] - s[:, 2]) ** 2 + (s[:, 2] - s[:, 0]) ** 2)
            + 3 * (s[:, 3] ** 2 + s[:, 4] ** 2 + s[:, 5] ** 2)
        )
    raise ValueError("von Mises undefined for this model")


# ---------------------------------------------------------------------------------------------
# small-matrix algebra on stacks (reference LinAlg.py:54-189; determinants are SIGNED)
# ---------------------------------------------------------------------------------------------
def det_stack(A):
    d = A.shape[-1]
    if d == 1:
        return A[..., 0, 0].copy()
    if d == 2:
        return A[..., 0, 0] * A[..., 1, 1] - A[..., 0, 1] * A[..., 1, 0]
    return (
        A[..., 0, 0] * (A[..., 1, 1] * A[..., 2, 2] - A[..., 1, 2] * A[..., 2, 1])
        - A[..., 0, 1] * (A[..., 1, 0] * A[..., 2, 2] - A[..., 1, 2] * A[..., 2, 0])
        + A[..., 0, 2] * (A[..., 1, 0] * A[..., 2, 1] - A[..., 1, 1] * A[..., 2, 0])
    )


def inv_stack(A):
    d = A.shape[-1]
    inv = np.empty_like(A)
    idet = 1.0 / det_stack(A)
    if d == 1:
        inv[..., 0, 0] = idet
    elif d == 2:
        inv[..., 0, 0] = idet * A[..., 1, 1]
        inv[..., 1, 1] = idet * A[..., 0, 0]
        inv[..., 0, 1] = -idet * A[..., 0, 1]
        inv[..., 1, 0] = -idet * A[..., 1, 0]
    else:
        inv[..., 0, 0] = idet * (A[..., 1, 1] * A[..., 2, 2] - A[..., 1, 2] * A[..., 2, 1])
        inv[..., 0, 1] = -idet * (A[..., 0, 1] * A[..., 2, 2] - A[..., 0, 2] * A[..., 2, 1])
        inv[..., 0, 2] = idet * (A[..., 0, 1] * A[..., 1, 2] - A[..., 0, 2] * A[..., 1, 1])
        inv[..., 1, 0] = -idet * (A[..., 1, 0] * A[..., 2, 2] - A[..., 1, 2] * A[..., 2, 0])
        inv[..., 1, 1] = idet * (A[..., 0, 0] * A[..., 2, 2] - A[..., 0, 2] * A[..., 2, 0])
        inv[..., 1, 2] = -idet * (A[..., 0, 0] * A[..., 1, 2] - A[..., 0, 2] * A[..., 1, 0])
        inv[..., 2, 0] = idet * (A[..., 1, 0] * A[..., 2, 1] - A[..., 1, 1] * A[..., 2, 0])
        inv[..., 2, 1] = -idet * (A[..., 0, 0] * A[..., 2, 1] - A[..., 0, 1] * A[..., 2, 0])
        inv[..., 2, 2] = idet * (A[..., 0, 0] * A[..., 1, 1] - A[..., 0, 1] * A[..., 1, 0])
    return inv


# ---------------------------------------------------------------------------------------------
# per-element kernels
# ---------------------------------------------------------------------------------------------
def point_quantities(dNdxi, X, q):
    """J, detJ, G = J^-1 dNdxi and u' for element stacks X (E,n,d), q (E,n,s); all (E,P,...)."""
    J = np.einsum("pdn,enk->epdk", dNdxi, X)
    detJ = det_stack(J)
    G = np.einsum("epdk,pkn->epdn", inv_stack(J), dNdxi)
    uprime = np.einsum("epdn,ens->epsd", G, q)
    return detJ, G, uprime


def element_jacobians(dNdxi, w, X, q, theta, model, E, nu, nonlinear=False):
    """K_e for a stack of elements -> (E, n*s, n*s)  (reference Element.py:426-486)."""
    numEl, n, _ = X.shape
    s = q.shape[2]
    P = len(w)
    D = d_matrix(model, E, nu)
    out = np.empty((numEl, n * s, n * s))
    for lo in range(0, numEl, CHUNK):
        hi = min(lo + CHUNK, numEl)
        detJ, G, up = point_quantities(dNdxi, X[lo:hi], q[lo:hi])
        S = strain_sens(up.reshape(-1, *up.shape[2:]), nonlinear)
        scale = np.repeat(theta[lo:hi], P)
        # weak Jacobian  wj[p,d,s,S,D] = theta * S[o,s,d] D[o,e] S[e,S,D]
        wj = np.einsum("posd,oe,peSD,p->pdsSD", S, D, S, scale, optimize=True)
        Gf = G.reshape(-1, *G.shape[2:])
        jac = np.einsum("pdn,pdsSD,pDN->pnsNS", Gf, wj, Gf, optimize=True)
        jac = jac.reshape(hi - lo, P, n, s, n, s)
        out[lo:hi] = np.einsum("epnsNS,ep,p->ensNS", jac, detJ, w).reshape(hi - lo, n * s, n * s)
    return out


def element_residuals(dNdxi, w, X, q, theta, model, E, nu, nonlinear=False):
    """R_e for a stack of elements -> (E, n, s)  (reference Element.py:370-424)."""
    numEl, n, _ = X.shape
    s = q.shape[2]
    P = len(w)
    D = d_matrix(model, E, nu)
    out = np.empty((numEl, n, s))
    for lo in range(0, numEl, CHUNK):
        hi = min(lo + CHUNK, numEl)
        detJ, G, up = point_quantities(dNdxi, X[lo:hi], q[lo:hi])
        upf = up.reshape(-1, *up.shape[2:])
        eps = strains(upf, nonlinear)
        sig = np.einsum("as,pa->ps", D, eps)
        S = strain_sens(upf, nonlinear)
        scale = np.repeat(theta[lo:hi], P)
        wr = np.einsum("pesd,pe,p->pds", S, sig, scale, optimize=True)
        Gf = G.reshape(-1, *G.shape[2:])
        r = np.einsum("pdn,pds->pns", Gf, wr).reshape(hi - lo, P, n, s)
        out[lo:hi] = np.einsum("epns,ep,p->ens", r, detJ, w)
    return out


MASS, VON_MISES, STRESS, STRAIN = 0, 1, 2, 3  # STRESS / STRAIN: the bar model's functions (Iso1D.py:203-225)
REDUCTIONS = ("sum", "mean", "integrate", "min", "max", "ksmax", "ksmin")
KS_RHO = 100.0  # FEMpy/Utils/KSAgg.py:26


def element_function(dNdxi, w, X, q, func, reduction, model, E, nu, rho, nonlinear=False):
    """Point function + element reduction (reference Element.py:163-273, IsoPlaneStress.py:187-227).

    reduction None returns the (E, P) point values.  The `integrate` reduction is sum_p v detJ w,
    without the thickness factor, as in the reference (Element.py:249-259)."""
    numEl = X.shape[0]
    P = len(w)
    out = np.empty((numEl, P)) if reduction is None else np.empty(numEl)
    D = d_matrix(model, E, nu)
    for lo in range(0, numEl, CHUNK):
        hi = min(lo + CHUNK, numEl)
        detJ, G, up = point_quantities(dNdxi, X[lo:hi], q[lo:hi])
        if func == MASS:
            vals = np.full((hi - lo, P), float(rho))
        elif func == VON_MISES:
            eps = strains(up.reshape(-1, *up.shape[2:]), nonlinear)
            sig = np.einsum("as,pa->ps", D, eps)
            vals = von_mises(sig, model, nu).reshape(hi - lo, P)
        elif func in (STRESS, STRAIN) and model == BAR_1D:
            eps = strains(up.reshape(-1, *up.shape[2:]), nonlinear)
            vals = (eps if func == STRAIN else np.einsum("as,pa->ps", D, eps)).reshape(hi - lo, P)
        else:
            raise ValueError("unknown function id")
        if reduction is None:
            out[lo:hi] = vals
            continue
        red = reduction.lower()
        assert red in REDUCTIONS, "elementReductionType not valid"
        if red == "sum":
            out[lo:hi] = vals.sum(axis=1)
        elif red == "mean":
            out[lo:hi] = np.average(vals, axis=1)
        elif red == "min":
            out[lo:hi] = vals.min(axis=1)
        elif red == "max":
            out[lo:hi] = vals.max(axis=1)
        elif red in ("ksmax", "ksmin"):
            # Element.py:261-265: both names reach ksAgg(values, "max") (the "ksmin" branch below it is unreachable);
            # KSAgg.py:52-53: max + log(mean(exp(rho (g - max)))) / rho
            mx = vals.max(axis=1)
            out[lo:hi] = mx + 1.0 / KS_RHO * np.log(1.0 / P * np.sum(np.exp(KS_RHO * (vals - mx[:, None])), axis=1))
        else:
            out[lo:hi] = np.einsum("ep,ep,p->e", vals, detJ, w)
    return out


# ---------------------------------------------------------------------------------------------
# assembly utilities (reference FEMpy/Utils/AssemblyUtils.py)
# ---------------------------------------------------------------------------------------------
def dof_from_node_inds(nodeInds, numStates):
    """dof = node*numStates + s, last axis widened (reference FEMpy/Model.py:387-414)."""
    nodeInds = np.asarray(nodeInds, dtype=np.int64)
    dof = nodeInds[..., None] * numStates + np.arange(numStates, dtype=np.int64)
    return dof.reshape(*nodeInds.shape[:-1], nodeInds.shape[-1] * numStates)


def local_matrices_to_coo(localMats, localDOF):
    """Element-major, row-major COO triplets with exactly-zero values dropped (AssemblyUtils.py:168-214)."""
    localMats = np.asarray(localMats)
    localDOF = np.asarray(localDOF, dtype=np.int64)
    nd = localDOF.shape[1]
    rows = np.repeat(localDOF, nd, axis=1).reshape(-1)
    cols = np.tile(localDOF, (1, nd)).reshape(-1)
    vals = localMats.reshape(-1)
    keep = np.flatnonzero(vals)
    return rows[keep], cols[keep], vals[keep]


def scatter_local_residuals(localRes, conn, globalRes):
    """globalRes[node*s + st] += localRes[e, n, st], in place (AssemblyUtils.py:217-238)."""
    s = localRes.shape[2]
    dof = dof_from_node_inds(conn, s).reshape(-1)
    np.add.at(globalRes, dof, localRes.reshape(-1))


def bc_dict_to_lists(bcDict):
    bcDOF, bcValues = [], []
    for key in bcDict:
        bcDOF += list(bcDict[key]["DOF"])
        bcValues += list(bcDict[key]["Value"])
    return bcDOF, bcValues


def loads_dict_to_vector(loadsDict, numDOF):
    """Fancy-index `+=`: duplicates inside one entry do not accumulate (AssemblyUtils.py:140-165)."""
    f = np.zeros(numDOF)
    for key in loadsDict:
        f[loadsDict[key]["DOF"]] += loadsDict[key]["Value"]
    return f


def apply_bcs_to_matrix(rows, cols, vals, bcDOF):
    """Zero every entry whose row is a BC DOF, then append (bc, bc, 1.0) per list occurrence
    (AssemblyUtils.py:26-62)."""
    bcDOF = np.asarray(bcDOF, dtype=np.int64)
    vals = vals.copy()
    vals[np.isin(rows, bcDOF)] = 0.0
    return np.append(rows, bcDOF), np.append(cols, bcDOF), np.append(vals, np.ones(len(bcDOF)))


def apply_bcs_to_vector(vec, state, bcDOF, bcValues):
    vec[np.asarray(bcDOF, dtype=np.int64)] = state[np.asarray(bcDOF, dtype=np.int64)] - np.asarray(bcValues, dtype=float)
    return vec


# ---------------------------------------------------------------------------------------------
# problem-level drivers (reference FEMpy/Problem.py:565-661, 282-355)
# ---------------------------------------------------------------------------------------------
class Block:
    """One element type of a mesh: tables + connectivity (element IDs run across blocks in order)."""

    def __init__(self, N, dNdxi, w, conn):
        self.N = np.asarray(N, dtype=float)
        self.dNdxi = np.asarray(dNdxi, dtype=float)
        self.w = np.asarray(w, dtype=float)
        self.conn = np.asarray(conn, dtype=np.int64)


def _gather(arr, conn):
    return arr[conn]


def assemble_matrix(blocks, coords, states, thickness, model, E, nu, bcDOF=None, nonlinear=False, dropZeros=True):
    """Global stiffness matrix as scipy CSR.  dropZeros=True is the reference's behaviour
    (value-dependent pattern); dropZeros=False gives the structural pattern."""
    numDOF = states.size
    s = states.shape[1]
    R, C, V = [], [], []
    e0 = 0
    for b in blocks:
        ne = b.conn.shape[0]
        Ke = element_jacobians(b.dNdxi, b.w, _gather(coords, b.conn), _gather(states, b.conn), thickness[e0 : e0 + ne], model, E, nu, nonlinear)
        dof = dof_from_node_inds(b.conn, s)
        if dropZeros:
            r, c, v = local_matrices_to_coo(Ke, dof)
        else:
            nd = dof.shape[1]
            r, c, v = np.repeat(dof, nd, axis=1).reshape(-1), np.tile(dof, (1, nd)).reshape(-1), Ke.reshape(-1)
        R.append(r), C.append(c), V.append(v)
        e0 += ne
    R, C, V = np.concatenate(R), np.concatenate(C), np.concatenate(V)
    if bcDOF is not None:
        R, C, V = apply_bcs_to_matrix(R, C, V, bcDOF)
    return csr_array(coo_array((V, (R, C)), shape=(numDOF, numDOF)))


def assemble_residual(blocks, coords, states, thickness, model, E, nu, loads=None, bcDOF=None, bcValues=None, nonlinear=False):
    numDOF = states.size
    res = np.zeros(numDOF)
    e0 = 0
    for b in blocks:
        ne = b.conn.shape[0]
        Re = element_residuals(b.dNdxi, b.w, _gather(coords, b.conn), _gather(states, b.conn), thickness[e0 : e0 + ne], model, E, nu, nonlinear)
        scatter_local_residuals(Re, b.conn, res)
        e0 += ne
    if loads is not None:
        res -= loads
    if bcDOF is not None:
        apply_bcs_to_vector(res, states.reshape(-1), bcDOF, bcValues)
    return res


def compute_function(blocks, coords, states, func, reduction, model, E, nu, rho, nonlinear=False):
    return [
        element_function(b.dNdxi, b.w, _gather(coords, b.conn), _gather(states, b.conn), func, reduction, model, E, nu, rho, nonlinear)
        for b in blocks
    ]


def structural_pattern(blocks, numNodes, numStates, extraDiagDOF=None):
    """Union of all element DOF x DOF couplings (+ optional extra diagonal entries), sorted CSR."""
    R, C = [], []
    for b in blocks:
        dof = dof_from_node_inds(b.conn, numStates)
        nd = dof.shape[1]
        R.append(np.repeat(dof, nd, axis=1).reshape(-1))
        C.append(np.tile(dof, (1, nd)).reshape(-1))
    if extraDiagDOF is not None and len(extraDiagDOF):
        R.append(np.asarray(extraDiagDOF, dtype=np.int64))
        C.append(np.asarray(extraDiagDOF, dtype=np.int64))
    R, C = np.concatenate(R), np.concatenate(C)
    n = numNodes * numStates
    m = csr_array(coo_array((np.ones(len(R)), (R, C)), shape=(n, n)))
    m.sort_indices()
    return m.indptr.astype(np.int64), m.indices.astype(np.int64)
