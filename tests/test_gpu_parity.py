"""Parity of the CUDA path (through the C ABI / the Model-Problem mirror) with the oracle, the golden
fixtures produced by the unmodified reference, and the reference's own known-answer vectors.

Bar (BASELINE.json north_star): sparsity pattern and DOF maps bit-exact; stiffness, residual and stress
values within 1e-12 relative (to the largest magnitude of the array -- summation order differs)."""
import numpy as np
import pytest
from scipy.sparse import csr_array

from fempy_b200 import Constitutive, AssemblyPlan, AssemblyUtils, meshes
from hostmirror import FEMpyModel
from oracle import fem_oracle as fo
from test_oracle_golden import COO_COLS, COO_DOF, COO_MATS, COO_ROWS, COO_VALS, KNOWN_QUAD_K, KNOWN_QUAD_X, two_quad_problem
from util import ELEMENT_FOR, FAMILY_ELEMENT, MAT, build_problem, golden, make_cm, mesh_case, mesh_case_names, oracle_blocks, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-12


def problem_from_case(c):
    cm = make_cm(c["model"], c["linear"])
    m = FEMpyModel(cm, nodeCoords=c["coords"], connectivity=c["conn"])
    m.setDesignVariables({"Thickness": c["thick"]})
    p = m.addProblem("case")
    p.BCs["golden"] = {"DOF": [int(d) for d in c["bcDOF"]], "Value": [float(v) for v in c["bcVal"]]}
    nz = np.flatnonzero(c["loads"])
    p.loads["golden"] = {"DOF": [int(d) for d in nz], "Value": [float(v) for v in c["loads"][nz]]}
    p.updateState(c["states"])
    return m, p


# ---- fixtures from the unmodified reference ----------------------------------------------------------------------
@pytest.mark.parametrize("name", mesh_case_names(golden("meshes.npz")))
def test_assembled_meshes_against_reference(name):
    g = golden("meshes.npz")
    c = mesh_case(g, name)
    m, p = problem_from_case(c)
    K = p._assembleMatrix(p.states, applyBCs=c["applyBCs"])
    R = p._assembleResidual(p.states, applyBCs=c["applyBCs"])
    assert isinstance(K, csr_array) and K.shape == (p.numDOF, p.numDOF)
    np.testing.assert_array_equal(K.indptr, c["indptr"])  # bit-exact pattern
    np.testing.assert_array_equal(K.indices, c["indices"])
    assert K.has_sorted_indices
    assert relerr(K.data, c["data"]) < TOL
    assert relerr(R, c["R"]) < TOL
    # fused K+R call gives the same answer as the two separate drivers
    K2, R2 = p.assembleSystem(applyBCs=c["applyBCs"])
    assert relerr(K2.data, c["data"]) < TOL and relerr(R2, c["R"]) < TOL
    for red in ("mean", "max", "min", "sum", "integrate"):
        vals = p.computeFunction("Von-Mises-Stress", red)
        for k in c["conn"]:
            assert relerr(vals[k], g[f"{name}_vm_{red}_{k}"]) < TOL, (red, k)
    mass = p.computeFunction("Mass", "integrate")
    for k in c["conn"]:
        assert relerr(mass[k], g[f"{name}_mass_integrate_{k}"]) < TOL
    np.testing.assert_allclose(p.computeFunction("Mass", "integrate", "sum"), g[name + "_mass_total"], rtol=1e-12)
    np.testing.assert_allclose(p.computeFunction("Von-Mises-Stress", "mean", "max"), g[name + "_vm_mean_max"], rtol=1e-12)


@pytest.mark.parametrize("family", sorted(FAMILY_ELEMENT))
@pytest.mark.parametrize("model", [0, 1])
@pytest.mark.parametrize("linear", [True, False])
def test_element_level_against_reference(family, model, linear):
    """the finer seams: Element.computeResidualJacobians / computeResiduals / computeFunction"""
    g = golden("elements.npz")
    el = ELEMENT_FOR[FAMILY_ELEMENT[family]]()
    cm = make_cm(model, linear)
    X, q, t = g[family + "_X"], g[family + "_q"], g[family + "_t"]
    key = f"{family}_m{model}_{'lin' if linear else 'nl'}"
    dvs = {"Thickness": t}
    Ke = el.computeResidualJacobians(q, X, dvs, cm)
    Re = el.computeResiduals(q, X, dvs, cm)
    assert Ke.shape == g[key + "_Ke"].shape and Re.shape == g[key + "_Re"].shape
    assert relerr(Ke, g[key + "_Ke"]) < TOL
    assert relerr(Re, g[key + "_Re"]) < TOL
    vm = el.computeFunction(X, q, dvs, cm.getFunction("Von-Mises-Stress"), None)
    assert vm.shape == g[key + "_vm"].shape and relerr(vm, g[key + "_vm"]) < TOL
    assert relerr(el.computeFunction(X, q, dvs, cm.getFunction("Von-Mises-Stress"), "mean"), g[key + "_vmmean"]) < TOL
    assert relerr(el.computeFunction(X, q, dvs, cm.getFunction("Mass"), "integrate"), g[key + "_massint"]) < TOL


EXTRA_RULES = {"quad1": (1, 3, 4), "quad2": (2, 4), "quad3": (3,), "tri1": (2, 3), "tri2": (1, 3), "tri3": (1, 3)}


def custom_function(states, stateGradients, coords, dvs):
    """the caller-supplied point function of oracle/gen_golden.py: gen_points"""
    return states[:, 0] * coords[:, 1] + stateGradients[:, 0, 1] * dvs["Thickness"] - 3.0 * stateGradients[:, 1, 0] * coords[:, 0]


@pytest.mark.parametrize("family", sorted(FAMILY_ELEMENT))
def test_point_quantities_against_reference(family):
    """Element._computeFunctionEvaluationQuantities (Element.py:275-368: _interpolationProduct, Jacobians, state gradients
    and their sensitivities) at arbitrary parametric points through femb_point_quantities, and Element.computeFunction
    with a function that is not one of the models' own (evaluated on the host from the device's point quantities)"""
    g, pg = golden("elements.npz"), golden("points.npz")
    el = ELEMENT_FOR[FAMILY_ELEMENT[family]]()
    X, q, t = g[family + "_X"], g[family + "_q"], g[family + "_t"]
    names = ["Coord", "State", "StateGrad", "Jac", "JacDet", "StateGradSens", "DVs", "NPrimeParam"]
    pq = el._computeFunctionEvaluationQuantities(pg[family + "_pts"], q, X, {"Thickness": t}, names)
    for key in names[:6]:
        want = pg[f"{family}_{key}"]
        assert pq[key].shape == want.shape and relerr(pq[key], want) < TOL, key
    np.testing.assert_array_equal(pq["DVs"]["Thickness"], pg[family + "_DVs"])
    assert pq["NPrimeParam"].shape == (5, 2, el.numNodes)
    for red in (None, "mean", "integrate", "max", "ksmax"):
        v = el.computeFunction(X, q, {"Thickness": t}, custom_function, red)
        want = pg[f"{family}_custom_{red}"]
        assert v.shape == want.shape and relerr(v, want) < TOL, red


@pytest.mark.parametrize("family", sorted(FAMILY_ELEMENT))
def test_non_default_quadrature_against_reference(family):
    """`intOrder=` of the element-level methods (Element.py:163,370,426): n x n Gauss rules for the quads, the 1- / 3- /
    4-point rules for the triangles, on the scatter / element-block / function kernels"""
    g, pg = golden("elements.npz"), golden("points.npz")
    el = ELEMENT_FOR[FAMILY_ELEMENT[family]]()
    X, q, t = g[family + "_X"], g[family + "_q"], g[family + "_t"]
    dvs = {"Thickness": t}
    cm, cmnl = make_cm(0, True), make_cm(1, False)
    for order in EXTRA_RULES[family]:
        key = f"{family}_int{order}"
        assert relerr(el.computeResidualJacobians(q, X, dvs, cm, intOrder=order), pg[key + "_Ke"]) < TOL, key
        assert relerr(el.computeResiduals(q, X, dvs, cm, intOrder=order), pg[key + "_Re"]) < TOL, key
        assert relerr(el.computeResidualJacobians(q, X, dvs, cmnl, intOrder=order), pg[key + "_Ke_nl"]) < TOL, key
        vm = el.computeFunction(X, q, dvs, cm.getFunction("Von-Mises-Stress"), None, intOrder=order)
        assert vm.shape == pg[key + "_vm"].shape and relerr(vm, pg[key + "_vm"]) < TOL, key
        assert relerr(el.computeFunction(X, q, dvs, cm.getFunction("Mass"), "integrate", intOrder=order), pg[key + "_massint"]) < TOL, key
        assert relerr(el.computeFunction(X, q, dvs, custom_function, "integrate", intOrder=order), pg[key + "_custom_integrate"]) < TOL, key


def test_model_on_a_non_default_rule_assembles_on_the_scatter_kernels():
    """a plan whose block carries a non-default rule (Q4 with the 3 x 3 rule) leaves the row-owner path and still matches
    the oracle evaluated with the same tables"""
    coords, conn = meshes.quad4_grid(13, 9)
    coords = coords + 0.01 * np.random.default_rng(4).random(coords.shape)
    el = ELEMENT_FOR["quad"]()
    N_, dN, w = el.tables(3)
    N = coords.shape[0]
    rng = np.random.default_rng(8)
    states, thick, loads = 1e-3 * rng.random((N, 2)), rng.random(conn["quad"].shape[0]) + 0.5, rng.random(2 * N)
    bcDOF, bcVal = np.array([0, 1, 7]), np.array([0.0, 1e-4, 0.0])
    plan = AssemblyPlan(N)
    plan.addBlock(N_, dN, w, conn["quad"])
    plan.finalize()
    assert plan.info("rows_path") == 0 and plan.info("fan_path") == 0
    K, R = plan.assemble(coords, states, thick, make_cm(0).material(), bcDOF, bcVal, True, loads)
    blocks = [fo.Block(N_, dN, w, conn["quad"])]
    Ko = fo.assemble_matrix(blocks, coords, states, thick, 0, MAT["E"], MAT["nu"], bcDOF, dropZeros=False)
    Ko.sort_indices()
    Ro = fo.assemble_residual(blocks, coords, states, thick, 0, MAT["E"], MAT["nu"], loads, bcDOF, bcVal)
    np.testing.assert_array_equal(plan.pattern()[1], Ko.indices)
    assert relerr(K, Ko.data) < TOL and relerr(R, Ro) < TOL
    plan.close()


@pytest.mark.parametrize("family", sorted(FAMILY_ELEMENT))
@pytest.mark.parametrize("model", [0, 1])
def test_ks_element_reduction_against_reference(family, model):
    """KS element reduction (Element.py:261-265, KSAgg.py:25-58) through femb_function: both names are the KS max"""
    g, ks = golden("elements.npz"), golden("ks.npz")
    el = ELEMENT_FOR[FAMILY_ELEMENT[family]]()
    X, q, t = g[family + "_X"], g[family + "_q"], g[family + "_t"]
    for tag, Emod in (("stiff", MAT["E"]), ("soft", 2.0)):
        cm = make_cm(model, True, E=Emod)
        for red in ("ksmax", "ksmin"):
            v = el.computeFunction(X, q, {"Thickness": t}, cm.getFunction("Von-Mises-Stress"), red)
            assert v.shape == (X.shape[0],) and relerr(v, ks[f"{family}_m{model}_E{tag}_{red}"]) < TOL, (tag, red)


def test_ks_reduction_on_a_mesh_and_global_ks_error():
    """Problem.computeFunction: element-level KS per element type == oracle; the global KS names raise TypeError as in
    the reference (Problem.py:340,347 call ksMax()/ksMin() without their argument)"""
    coords, conn = meshes.quad4_grid(24, 17)
    m, p = build_problem(coords, conn, 0, E=3.0, states=1e-2 * np.random.default_rng(5).random(coords.shape))
    got = p.computeFunction("Von-Mises-Stress", "ksmax")["quad"]
    want = fo.compute_function(oracle_blocks(conn), coords, p.states, fo.VON_MISES, "ksmax", 0, 3.0, MAT["nu"], MAT["rho"])[0]
    assert relerr(got, want) < TOL
    hard = p.computeFunction("Von-Mises-Stress", "max")["quad"]
    assert np.all(got <= hard + 1e-15) and np.all(got >= hard + np.log(1.0 / 4) / 100.0 - 1e-15)  # KS-mean bounds of the max
    with pytest.raises(TypeError):
        p.computeFunction("Von-Mises-Stress", "mean", "ksmax")


def test_lbracket_config1():
    """BASELINE config 1 (Examples/LBracketExample.py): pattern, K, R, solve, von Mises, mass"""
    g = golden("lbracket.npz")
    E, nu, rho, t = g["material"]
    from fempy_b200 import Constitutive

    cm = Constitutive.IsoPlaneStress(E, nu, rho, t)
    model = FEMpyModel(cm, nodeCoords=g["coords"], connectivity={"quad": g["conn"].astype(np.int64)})
    model.addFixedBCToNodes(name="Fixed", nodeInds=[int(i) for i in g["top"]], dof=[0, 1], value=0.0)
    prob = model.addProblem("Vertical-Load")
    prob.addLoadToNodes(name="RightEdgeLoad", nodeInds=[int(i) for i in g["right"]], dof=1, value=-1e5, totalLoad=True)
    K = prob._assembleMatrix(prob.states)
    R = prob._assembleResidual(prob.states)
    np.testing.assert_array_equal(K.indptr, g["indptr"])
    np.testing.assert_array_equal(K.indices, g["indices"])
    assert K.nnz == 297028
    assert relerr(K.data, g["data"]) < TOL
    assert relerr(R, g["R"]) < TOL
    prob.solve()
    assert relerr(prob.states, g["u"]) < 1e-5  # through a direct solve: conditioning-limited (cond ~ 1e9)
    np.testing.assert_allclose(np.abs(prob.states).max(), 3.285941388626e-02, rtol=1e-5)
    prob.updateState(g["u"])
    assert relerr(prob.computeFunction("Von-Mises-Stress", "mean")["quad"], g["vm_mean"]) < TOL
    np.testing.assert_allclose(prob.computeFunction("Von-Mises-Stress", "mean", "max"), 1.632125596200e09, rtol=1e-9)
    np.testing.assert_allclose(prob.computeFunction("Mass", "integrate", "sum"), 1.779200000000e03, rtol=1e-9)
    assert relerr(prob._assembleResidual(prob.states), g["R_solved"]) < 1e-6 or np.abs(g["R_solved"]).max() < 1e-3


# ---- device-side model: coordinates / thickness resident between calls ---------------------------------
def test_resident_geometry_matches_explicit_inputs():
    """femb_plan_set_geometry: NULL coords / thickness use the resident copies, bit-identical to explicit inputs; an explicit
    pointer ends the residency; NULL without a resident copy is an error"""
    from fempy_b200._lib import FembError

    coords, conn, states, thick, loads, bcDOF, bcVal = _path_case("quad", 33)
    plan = _raw_plan("quad", coords, conn)
    mat = make_cm(0).material()
    K0, R0 = plan.assemble(coords, states, thick, mat, bcDOF, bcVal, True, loads)
    with pytest.raises(FembError, match="resident"):
        plan.assemble(None, states, None, mat, bcDOF, bcVal, True, loads)
    plan.setGeometry(coords, thick)
    K1, R1 = plan.assemble(None, states, None, mat, bcDOF, bcVal, True, loads)
    np.testing.assert_array_equal(K1, K0)
    np.testing.assert_array_equal(R1, R0)
    vm0 = plan.function(coords, states, mat, 1, "mean")  # explicit coords: that array is no longer resident
    with pytest.raises(FembError, match="resident"):
        plan.function(None, states, mat, 1, "mean")
    plan.setGeometry(coords, None)
    np.testing.assert_array_equal(plan.function(None, states, mat, 1, "mean"), vm0)
    with pytest.raises(FembError, match="resident"):
        plan.assemble(None, states, None, mat, bcDOF, bcVal, True, loads)  # thickness was dropped
    plan.close()


def test_model_geometry_changes_reach_the_device():
    """FEMpyModel.setCoordinates / setDesignVariables invalidate the resident copies (version counter)"""
    coords, conn = meshes.quad4_grid(14, 9)
    m, p = build_problem(coords, conn, 0, states=1e-3 * np.random.default_rng(1).random(coords.shape))
    blocks = oracle_blocks(conn)

    def oracle_K():
        Ko = fo.assemble_matrix(blocks, m.nodeCoords, p.states, m.dvs["Thickness"], 0, MAT["E"], MAT["nu"], None, dropZeros=False)
        Ko.sort_indices()
        return Ko.data

    assert relerr(p._assembleMatrix(p.states, applyBCs=False).data, oracle_K()) < TOL
    m.setCoordinates(m.nodeCoords * np.array([1.5, 0.75]))
    assert relerr(p._assembleMatrix(p.states, applyBCs=False).data, oracle_K()) < TOL
    m.setDesignVariables({"Thickness": np.random.default_rng(2).random(m.dvs["Thickness"].shape) + 0.5})
    assert relerr(p._assembleMatrix(p.states, applyBCs=False).data, oracle_K()) < TOL
    assert relerr(p.computeFunction("Von-Mises-Stress", "mean")["quad"],
                  fo.compute_function(blocks, m.nodeCoords, p.states, fo.VON_MISES, "mean", 0, MAT["E"], MAT["nu"], MAT["rho"])[0]) < TOL


# ---- device-resident linear solve: the hand-off after assembly ------------------------------------
def test_lbracket_solve_on_device():
    """config 1 solved without the CSR values leaving the GPU (femb_assemble_solve): the state equals the reference's
    direct solve to the conditioning of the problem, and the converged state has no residual"""
    g = golden("lbracket.npz")
    E, nu, rho, t = g["material"]
    from fempy_b200 import Constitutive

    model = FEMpyModel(Constitutive.IsoPlaneStress(E, nu, rho, t), nodeCoords=g["coords"], connectivity={"quad": g["conn"].astype(np.int64)})
    model.addFixedBCToNodes(name="Fixed", nodeInds=[int(i) for i in g["top"]], dof=[0, 1], value=0.0)
    prob = model.addProblem("Vertical-Load")
    prob.addLoadToNodes(name="RightEdgeLoad", nodeInds=[int(i) for i in g["right"]], dof=1, value=-1e5, totalLoad=True)
    info = prob.solveOnDevice(relTol=1e-11)
    assert 0 < info["iterations"] < 50000 and info["relativeResidual"] <= 1e-11
    assert relerr(prob.Res, g["R"]) < TOL  # the residual the step started from
    assert relerr(prob.states, g["u"]) < 1e-5
    np.testing.assert_allclose(np.abs(prob.states).max(), 3.285941388626e-02, rtol=1e-5)
    R1 = prob._assembleResidual(prob.states)
    assert np.abs(R1).max() < 1e-6 * np.abs(g["R"]).max()


@pytest.mark.parametrize("elType,n", [("quad", 40), ("triangle", 31), ("quad9", 12)])
def test_pcg_matches_direct_solve(elType, n):
    """femb_solve_pcg on the assembled values (BC rows kept, prescribed values != 0, duplicate BCs) against scipy's direct
    solve.  The reference hands the row-only-BC matrix to the factorisation as it is; with prescribed values != 0 that
    system is badly scaled (BC rows ~ 1, stiffness rows ~ 1e10) and SuperLU returns the constrained DOFs to ~1e-3 only
    (tools/dbg_pcg.py), so the yardstick here is the direct solve of the equivalent reduced system
    K_ff x_f = b_f - K_fc x_c with x_c = b_c / count."""
    from scipy.sparse.linalg import spsolve

    coords, conn, states, thick, loads, bcDOF, bcVal = _path_case(elType, n)
    thick = thick * 100.0
    plan = _raw_plan(elType, coords, conn)
    K, R = plan.assemble(coords, states, thick, make_cm(0).material(), bcDOF, bcVal, True, 1e6 * loads)
    indptr, indices = plan.pattern()
    A = csr_array((K, indices, indptr), shape=(R.size, R.size))
    free = np.setdiff1d(np.arange(R.size), bcDOF)
    con = np.unique(bcDOF)
    want = np.zeros(R.size)
    want[con] = (-R / np.maximum(np.bincount(bcDOF, minlength=R.size), 1))[con]
    want[free] = spsolve(A[free][:, free].tocsc(), (-R - A @ want)[free])
    asIs = spsolve(A.tocsc(), -R)  # what the reference's host solve returns
    assert relerr(asIs[free], want[free]) < 1e-3
    x, iters, res = plan.solvePCG(K, -R, relTol=1e-12, maxIter=20000)
    assert 0 < iters < 20000 and res <= 1e-12
    trueRes = np.abs((A @ x + R)[free]).max() / np.abs(R).max()
    assert trueRes <= 1e-9, trueRes
    assert relerr(x, want) < 1e-7, (relerr(x, want), trueRes, iters)
    # constrained DOFs: x_c = b_c / count exactly as the row equation count * x_c = b_c says
    cnt = np.bincount(bcDOF, minlength=R.size)
    np.testing.assert_allclose(x[np.unique(bcDOF)], (-R / np.maximum(cnt, 1))[np.unique(bcDOF)], rtol=1e-15)
    plan.close()


def test_pcg_without_bcs_and_bad_arguments():
    """no BC set: plain PCG on K + shift (SPD); argument checks"""
    coords, conn = meshes.quad4_grid(9, 7)
    plan = _raw_plan("quad", coords, conn)
    N = coords.shape[0]
    K, _ = plan.assemble(coords, np.zeros((N, 2)), np.ones(conn["quad"].shape[0]), make_cm(0).material(), None, None, False, None)
    indptr, indices = plan.pattern()
    A = csr_array((K, indices, indptr), shape=(2 * N, 2 * N))
    diagPos = np.flatnonzero(indices == np.repeat(np.arange(2 * N), np.diff(indptr)))
    Ks = K.copy()
    Ks[diagPos] += 1e9  # rigid-body modes out of the null space
    b = np.random.default_rng(0).random(2 * N)
    x, iters, res = plan.solvePCG(Ks, b, relTol=1e-12)
    As = csr_array((Ks, indices, indptr), shape=A.shape)
    assert relerr(As @ x, b) < 1e-10
    from fempy_b200._lib import FembError

    with pytest.raises(FembError):
        plan.solvePCG(Ks, b, relTol=0.0)
    plan.close()


# ---- the reference's own known-answer vectors, through the CUDA path ------------------------------------------------
def test_local_matrices_to_coo_reference_vector():
    """reference tests/testAssembleutils.py:55-81"""
    r, c, v = AssemblyUtils.localMatricesToCOOArrays(COO_MATS, COO_DOF)
    np.testing.assert_equal(r, COO_ROWS)
    np.testing.assert_equal(c, COO_COLS)
    np.testing.assert_equal(v, COO_VALS)
    assert r.dtype == np.int64 and c.dtype == np.int64


def test_scatter_local_residuals():
    """AssemblyUtils.py:217-238 against the oracle restatement"""
    rng = np.random.default_rng(3)
    conn = rng.integers(0, 50, size=(200, 4))
    res = rng.random((200, 4, 2))
    out = np.arange(100, dtype=float)
    ref = out.copy()
    fo.scatter_local_residuals(res, conn, ref)
    AssemblyUtils.scatterLocalResiduals(res, conn, out)
    np.testing.assert_allclose(out, ref, rtol=1e-13)


def test_known_quad_stiffness_matrix():
    """reference tests/testElements.py:34-45,80-82,169"""
    el = ELEMENT_FOR["quad"]()
    Ke = el.computeResidualJacobians(np.zeros((1, 4, 2)), KNOWN_QUAD_X[None], {"Thickness": np.ones(1)}, make_cm(1))
    np.testing.assert_allclose(Ke[0], KNOWN_QUAD_K, rtol=1e-3)


@pytest.mark.parametrize("family", sorted(FAMILY_ELEMENT))
def test_zero_residual_and_fd_consistency(family):
    """reference tests/testElements.py:135-180: zero residual at zero state; K dq == dR (1e-5)"""
    el = ELEMENT_FOR[FAMILY_ELEMENT[family]]()
    g = golden("elements.npz")
    X = g[family + "_X"]
    rng = np.random.default_rng(5)
    dvs = {"Thickness": rng.random(X.shape[0])}
    cm = make_cm(1)
    res = el.computeResiduals(np.zeros_like(X), X, dvs, cm)
    assert res.shape == (X.shape[0], el.numNodes, 2)
    np.testing.assert_allclose(res, 0, atol=1e-10)
    q = rng.random(X.shape)
    Jac = el.computeResidualJacobians(q, X, dvs, cm)
    res0 = el.computeResiduals(q, X, dvs, cm)
    for _ in range(3):
        pert = rng.random(q.shape) * 1e-6 - 0.5e-6
        dR = el.computeResiduals(q + pert, X, dvs, cm) - res0
        for i in range(X.shape[0]):
            np.testing.assert_allclose(dR[i].flatten(), Jac[i] @ pert[i].flatten(), atol=1e-5, rtol=1e-5)


def test_two_quad_problem_reference_test():
    """reference tests/testProblem.py:73-143 verbatim in spirit: sparsity, exact RHS, solve, functions"""
    coords, conn, _, _, _, sparsity = two_quad_problem()
    m, p = build_problem(coords, conn, 0, E=71.7e9, nu=0.33, rho=2780, t=5e-3)
    p.addFixedBCToNodes(name="YFixed", nodeInds=[1, 5], dof=1, value=1.0)
    p.addFixedBCToNodes(name="XFixed", nodeInds=5, dof=0, value=0.1)
    p.addLoadToNodes(name="YLoad", nodeInds=[2], dof=1, value=-10.0)
    p.updateJacobian(applyBCs=False)
    np.testing.assert_equal(p.Jacobian.nonzero(), np.nonzero(sparsity))
    p.markJacOutOfDate()
    expectedRHS = np.zeros(12)
    expectedRHS[[3, 5, 11, 10]] = [-1.0, 10.0, -1.0, -0.1]
    p.updateResidual()
    np.testing.assert_equal(p.Res, expectedRHS)
    p.solve()
    p.updateResidual()
    np.testing.assert_allclose(p.Res, np.zeros(12), rtol=1e-6, atol=1e-6)
    u = p.states.flatten()
    np.testing.assert_approx_equal(u[3], 1.0, significant=10)
    np.testing.assert_approx_equal(u[11], 1.0, significant=10)
    np.testing.assert_approx_equal(u[10], 0.1, significant=10)
    sparsity[[3, 10, 11], :] = 0
    sparsity[[3, 10, 11], [3, 10, 11]] = 1.0
    np.testing.assert_equal(p.Jacobian.nonzero(), np.nonzero(sparsity))
    np.testing.assert_allclose(2780 * 2, p.computeFunction("Mass", elementReductionType="integrate", globalReductionType="sum"))
    np.testing.assert_equal({"quad": np.array([2780.0, 2780.0])}, p.computeFunction("Mass", elementReductionType="max"))
    np.testing.assert_equal(2780 * 2, p.computeFunction("Mass", elementReductionType="min", globalReductionType="sum"))
    with pytest.raises(ValueError):
        p.computeFunction("Pressure", "max")
    with pytest.raises(AssertionError):
        p.computeFunction("Mass", "median")


# ---- against the oracle at sizes it finishes in seconds -------------------------------------------------------------
MEDIUM = [("quad", 128, 0), ("quad", 96, 1), ("quad9", 40, 0), ("triangle", 100, 0), ("triangle", 64, 1)]


@pytest.mark.parametrize("elType,n,model", MEDIUM)
def test_medium_meshes_against_oracle(elType, n, model):
    coords, conn = meshes.GENERATORS[elType](n, n)
    N = coords.shape[0]
    numEl = conn[elType].shape[0]
    rng = np.random.default_rng(0)
    states = 1e-3 * rng.random((N, 2))
    thick = rng.random(numEl) * 1e-2 + 1e-3
    left, right = meshes.edge_nodes(coords, x=0.0), meshes.edge_nodes(coords, x=2.0)
    m, p = build_problem(coords, conn, model, thickness=thick, states=states,
                         bcs=[("Fixed", left, [0, 1], 0.0)], loads=[("Load", right, [0], 1e4, False)])
    K, R = p.assembleSystem()
    blocks = oracle_blocks(conn)
    ip, ix = fo.structural_pattern(blocks, N, 2)
    np.testing.assert_array_equal(K.indptr, ip)
    np.testing.assert_array_equal(K.indices, ix)
    bcDOF, bcVal = AssemblyUtils.convertBCDictToLists(p.getBCs())
    loads = AssemblyUtils.convertLoadsDictToVector(p.loads, 2 * N)
    Ko = fo.assemble_matrix(blocks, coords, states, thick, model, MAT["E"], MAT["nu"], bcDOF, dropZeros=False)
    Ko.sort_indices()
    assert relerr(K.data, Ko.data) < TOL
    Ro = fo.assemble_residual(blocks, coords, states, thick, model, MAT["E"], MAT["nu"], loads, bcDOF, bcVal)
    assert relerr(R, Ro) < TOL
    vm = p.computeFunction("Von-Mises-Stress", "mean")[elType]
    vmo = fo.compute_function(blocks, coords, states, fo.VON_MISES, "mean", model, MAT["E"], MAT["nu"], MAT["rho"])[0]
    assert relerr(vm, vmo) < TOL


# ---- full-size configurations through size-independent properties --------------------------------------------------
FULL = [("quad", 512, "config 2: 526k DOF"), ("quad9", 250, "config 3 family: 502k DOF"), ("triangle", 700, "config 4 family: 983k DOF"),
        ("quad9", 500, "config 3: 2.0M DOF, nnz 64M"), ("triangle", 1580, "config 4: 5.0M DOF, nnz 70M")]


@pytest.mark.parametrize("elType,n,label", FULL)
def test_full_size_properties(elType, n, label):
    coords, conn = meshes.GENERATORS[elType](n, n)
    N = coords.shape[0]
    numEl = conn[elType].shape[0]
    rng = np.random.default_rng(1)
    thick = rng.random(numEl) * 1e-2 + 1e-3
    m, p = build_problem(coords, conn, 0, thickness=thick)
    K = p._assembleMatrix(p.states, applyBCs=False)
    scale = np.abs(K.data).max()
    # nnz of a structured grid
    expected = {"quad": 4 * (3 * n + 1) ** 2, "quad9": 4 * (8 * n + 1) ** 2, "triangle": 4 * (7 * n * n + 6 * n + 1)}[elType]
    assert K.nnz == expected
    assert np.all(np.diff(K.indptr) > 0) and K.has_sorted_indices
    # symmetry
    D = (K - K.T).tocsr()
    assert D.nnz == 0 or np.abs(D.data).max() < 1e-12 * scale
    # rigid-body modes are in the null space: two translations and the linearised rotation
    for mode in (np.tile([1.0, 0.0], N), np.tile([0.0, 1.0], N), np.stack((-coords[:, 1], coords[:, 0]), axis=1).reshape(-1)):
        assert np.abs(K @ mode).max() < 1e-10 * scale * np.abs(mode).max()
    # linear model: the stress-based residual equals K u (two independent code paths)
    u = 1e-3 * rng.random((N, 2))
    R = p._assembleResidual(u, applyBCs=False)
    Ku = K @ u.reshape(-1)
    assert np.abs(R - Ku).max() < 1e-12 * np.abs(Ku).max() * 50
    # patch test: a linear displacement field gives a constant stress state
    p.updateState(np.stack((1e-3 * coords[:, 0] + 2e-4 * coords[:, 1], -3e-4 * coords[:, 0] + 5e-4 * coords[:, 1]), axis=1))
    vm = p.computeFunction("Von-Mises-Stress", None)[elType]
    Dm = fo.d_matrix(0, MAT["E"], MAT["nu"])
    sig = Dm @ np.array([1e-3, 5e-4, 2e-4 - 3e-4])
    vm_exact = np.sqrt(sig[0] ** 2 - sig[0] * sig[1] + sig[1] ** 2 + 3 * sig[2] ** 2)
    np.testing.assert_allclose(vm, vm_exact, rtol=1e-9)
    # total mass = rho * area of the warped domain (trapezoid: width 2, heights 2 and 1)
    np.testing.assert_allclose(p.computeFunction("Mass", "integrate", "sum"), MAT["rho"] * 3.0, rtol=1e-11)
    # idempotence: a second assembly reproduces the first exactly (write-once path: fixed summation order)
    K2 = p._assembleMatrix(np.zeros((N, 2)), applyBCs=False)
    np.testing.assert_array_equal(K2.data, K.data)
    assert np.abs(K2.data - K.data).max() < 1e-13 * scale


# ---- the assembly paths: row-owner kernel (default), slot-map scatter with FP64 atomics (general fallback: 16-node
#      families, node degree above the row-owner limit) ------------------------------------------------------------------
PATHS = {"scatter": 0, "rows": 2}


def _raw_plan(elType, coords, conn, path="rows", balance=None, fan=None, fanCompact=None):
    el = ELEMENT_FOR[elType]()
    plan = AssemblyPlan(coords.shape[0])
    plan.setOption("assembly_path", PATHS[path])
    if fan is not None:
        plan.setOption("fan_rows", fan)
    if fanCompact is not None:
        plan.setOption("fan_compact", fanCompact)
    if balance is not None:
        plan.setOption("balance_rows", balance)
    N_, dN, w = el.tables()
    plan.addBlock(N_, dN, w, conn[elType])
    plan.finalize()
    return plan


def _path_case(elType, n, seed=7):
    coords, conn = meshes.GENERATORS[elType](n, n - 3)
    N = coords.shape[0]
    numEl = conn[elType].shape[0]
    rng = np.random.default_rng(seed)
    states = 1e-3 * rng.random((N, 2))
    thick = rng.random(numEl) * 1e-2 + 1e-3
    loads = rng.random(2 * N)
    left = meshes.edge_nodes(coords, x=0.0)
    bcDOF = np.concatenate((2 * left, 2 * left[:3] + 1, 2 * left[:2]))  # with duplicates
    bcVal = rng.random(bcDOF.size) * 1e-4
    return coords, conn, states, thick, loads, bcDOF, bcVal


def _check_against_oracle(plan, elType, case, launchesExpected):
    coords, conn, states, thick, loads, bcDOF, bcVal = case
    mat = make_cm(0).material()
    K1, R1 = plan.assemble(coords, states, thick, mat, bcDOF, bcVal, True, loads)
    assert plan.lastLaunches == launchesExpected
    blocks = oracle_blocks(conn)
    Ko = fo.assemble_matrix(blocks, coords, states, thick, 0, MAT["E"], MAT["nu"], bcDOF, dropZeros=False)
    Ko.sort_indices()
    Ro = fo.assemble_residual(blocks, coords, states, thick, 0, MAT["E"], MAT["nu"], loads, bcDOF, bcVal)
    ip, ix = plan.pattern()
    np.testing.assert_array_equal(ip, Ko.indptr)
    np.testing.assert_array_equal(ix, Ko.indices)
    assert relerr(K1, Ko.data) < TOL and relerr(R1, Ro) < TOL
    # K-only and R-only variants, and no BCs
    K2, _ = plan.assemble(coords, states, thick, mat, None, None, False, None, wantR=False)
    _, R2 = plan.assemble(coords, states, thick, mat, None, None, False, loads, wantK=False)
    Kn = fo.assemble_matrix(blocks, coords, states, thick, 0, MAT["E"], MAT["nu"], None, dropZeros=False)
    Kn.sort_indices()
    Rn = fo.assemble_residual(blocks, coords, states, thick, 0, MAT["E"], MAT["nu"], loads)
    assert relerr(K2, Kn.data) < TOL and relerr(R2, Rn) < TOL
    return K1, R1


@pytest.mark.parametrize("elType,n", [("quad", 37), ("triangle", 29), ("quad9", 14)])
def test_default_path_matches_scatter_path_and_oracle(elType, n):
    """default: row-owner kernel for single-block meshes of 3/4/6/9/10-node elements (1 launch); scatter kernels otherwise"""
    case = _path_case(elType, n)
    coords, conn = case[0], case[1]
    plan = _raw_plan(elType, coords, conn)
    assert plan.info("assembly_path") == 2
    onRowsPath = plan.info("rows_path") == 1
    assert onRowsPath
    K1, R1 = _check_against_oracle(plan, elType, case, launchesExpected=1 if onRowsPath else 3)
    plan.setOption("assembly_path", 0)
    mat = make_cm(0).material()
    K0, R0 = plan.assemble(coords, case[2], case[3], mat, case[5], case[6], True, case[4])
    assert plan.lastLaunches == 3
    assert relerr(K1, K0) < 1e-13 and relerr(R1, R0) < 1e-13
    plan.close()


def _renumbered(case, seed=3):
    """the same mesh with randomly permuted node numbers (no locality in the numbering, irregular row order)"""
    coords, conn, states, thick, loads, bcDOF, bcVal = case
    N = coords.shape[0]
    perm = np.random.default_rng(seed).permutation(N)  # new id of old node
    inv = np.argsort(perm)
    conn2 = {k: perm[v] for k, v in conn.items()}
    loads2 = loads.reshape(N, 2)[inv].ravel()
    bc2 = 2 * perm[bcDOF // 2] + bcDOF % 2
    return coords[inv], conn2, states[inv], thick, loads2, bc2, bcVal


@pytest.mark.parametrize("elType,n", [("quad", 37), ("triangle", 29), ("quad9", 11), ("triangle6", 12)])
@pytest.mark.parametrize("renumber", [False, True])
@pytest.mark.parametrize("balance", [0, 2])
def test_row_owner_kernel_matches_oracle(elType, n, renumber, balance):
    """default path: one thread per node row, every K value and R entry written exactly once = 1 launch;
    in node order (balance 0) and with the balanced row schedule forced (2: windows of nodes ordered by corner count)"""
    case = _path_case(elType, n)
    if renumber:
        case = _renumbered(case)
    plan = _raw_plan(elType, case[0], case[1], path="rows", balance=balance)
    assert plan.info("rows_path") == 1
    assert plan.info("rows_balanced") == (1 if balance == 2 else 0)
    assert 0 < plan.info("row_iters_balanced") <= plan.info("row_iters_node_order")
    _check_against_oracle(plan, elType, case, launchesExpected=1)
    # Green-strain model through the same kernel family
    coords, conn, states, thick, loads, bcDOF, bcVal = case
    matNL = make_cm(1, linear=False).material()
    K, R = plan.assemble(coords, 50 * states, thick, matNL, bcDOF, bcVal, True, loads)
    assert plan.lastLaunches == 1
    blocks = oracle_blocks(conn)
    Ko = fo.assemble_matrix(blocks, coords, 50 * states, thick, 1, MAT["E"], MAT["nu"], bcDOF, nonlinear=True, dropZeros=False)
    Ko.sort_indices()
    Ro = fo.assemble_residual(blocks, coords, 50 * states, thick, 1, MAT["E"], MAT["nu"], loads, bcDOF, bcVal, nonlinear=True)
    assert relerr(K, Ko.data) < TOL and relerr(R, Ro) < TOL
    plan.close()


@pytest.mark.parametrize("elType,n", [("quad", 37), ("triangle", 29)])
@pytest.mark.parametrize("renumber", [False, True])
@pytest.mark.parametrize("fan,compact", [(1, 1), (1, 0), (0, 1)])
def test_fan_kernel_matches_oracle(elType, n, renumber, fan, compact):
    """write-once fan kernel (single block of 3- / 4-node elements, small strain): corners of a node chained into fans,
    every row block finished in registers and stored once; compact (24-bit differences, element id inline) and wide
    records; fan_rows = 0 runs the accumulating row-owner kernel on the same plan"""
    case = _path_case(elType, n)
    if renumber:
        case = _renumbered(case)
    plan = _raw_plan(elType, case[0], case[1], fan=fan, fanCompact=compact)
    assert plan.info("rows_path") == 1 and plan.info("fan_path") == fan
    assert plan.info("fan_compact") == (1 if compact else 0)  # quads: 24-bit differences; triangles: 8-byte chained records
    K1, R1 = _check_against_oracle(plan, elType, case, launchesExpected=1)
    plan.setOption("fan_rows", 1 - fan)  # the other kernel on the same plan: same values to rounding
    mat = make_cm(0).material()
    K0, R0 = plan.assemble(case[0], case[2], case[3], mat, case[5], case[6], True, case[4])
    assert relerr(K1, K0) < 1e-13 and relerr(R1, R0) < 1e-13
    # K only / R only variants
    Kk, _ = plan.assemble(case[0], case[2], case[3], mat, case[5], case[6], True, case[4], wantR=False)
    _, Rr = plan.assemble(case[0], case[2], case[3], mat, case[5], case[6], True, case[4], wantK=False)
    np.testing.assert_array_equal(Kk, K0)
    np.testing.assert_array_equal(Rr, R0)
    plan.close()


def test_chained_triangle_records_fall_back_without_locality_and_look_ahead_is_only_a_hint():
    """8-byte chained 3-node records need |node - owner| < 2^15: a 70 k-node mesh with random numbering keeps the 16-byte
    records (same values as the scatter kernels); the L2 look-ahead distance never changes a bit of the result"""
    case = _renumbered(_path_case("triangle", 268))
    coords, conn, states, thick, loads, bcDOF, bcVal = case
    assert coords.shape[0] > 65536
    plan = _raw_plan("triangle", coords, conn)
    assert plan.info("fan_path") == 1 and plan.info("fan_compact") == 0
    mat = make_cm(0).material()
    K, R = plan.assemble(coords, states, thick, mat, bcDOF, bcVal, True, loads)
    ref = _raw_plan("triangle", coords, conn, path="scatter")
    K0, R0 = ref.assemble(coords, states, thick, mat, bcDOF, bcVal, True, loads)
    np.testing.assert_array_equal(plan.pattern()[1], ref.pattern()[1])
    assert relerr(K, K0) < 1e-13 and relerr(R, R0) < 1e-13
    ref.close()
    plan.close()
    ordered = _path_case("triangle", 268)
    plan = _raw_plan("triangle", ordered[0], ordered[1])
    assert plan.info("fan_compact") == 1
    results = []
    for pf in (-1, 0, 7, 100000):
        plan.setOption("fan_prefetch", pf)
        results.append(plan.assemble(ordered[0], ordered[2], ordered[3], mat, ordered[5], ordered[6], True, ordered[4]))
    for Kp, Rp in results[1:]:
        np.testing.assert_array_equal(Kp, results[0][0])
        np.testing.assert_array_equal(Rp, results[0][1])
    plan.close()


def _check_small(coords, conn, elType, expectFan, expectRows=1):
    N = coords.shape[0]
    numEl = conn[elType].shape[0]
    rng = np.random.default_rng(5)
    case = (coords, conn, 1e-3 * rng.random((N, 2)), rng.random(numEl) + 0.5, rng.random(2 * N), np.array([0, 1, 3]), np.array([1e-4, 0.0, 2e-4]))
    plan = _raw_plan(elType, coords, conn)
    assert plan.info("rows_path") == expectRows and plan.info("fan_path") == expectFan
    _check_against_oracle(plan, elType, case, launchesExpected=1 if expectRows else 3)
    plan.close()


def test_fan_kernel_open_fans_bow_ties_and_fallbacks():
    """fans that are open (boundary nodes), several fans at one node (two quads touching in a point), and meshes whose
    corners do not chain (one element with reversed orientation): the plan falls back to the accumulating kernel"""
    # 2 x 1 quads: every node has an open fan
    coords, conn = meshes.quad4_grid(2, 1)
    _check_small(coords + 0.01 * np.random.default_rng(1).random(coords.shape), conn, "quad", expectFan=1)
    # bow tie: two quads sharing only node 2
    coords = np.array([[0.0, 0.0], [1.0, 0.1], [1.1, 1.0], [0.0, 1.2], [2.2, 1.1], [2.3, 2.0], [1.2, 2.1]])
    conn = {"quad": np.array([[0, 1, 2, 3], [2, 4, 5, 6]])}
    _check_small(coords, conn, "quad", expectFan=1)
    # 3 x 3 quads with one element listed clockwise: its corners do not chain with their neighbours
    coords, conn = meshes.quad4_grid(3, 3)
    coords = coords + 0.01 * np.random.default_rng(2).random(coords.shape)
    q = conn["quad"].copy()
    q[4] = q[4][[0, 3, 2, 1]]
    _check_small(coords, {"quad": q}, "quad", expectFan=0)
    # triangles: a closed fan of 5 around node 0 and the same fan with one triangle reversed
    ang = 2 * np.pi * np.arange(5) / 5
    coords = np.vstack(([[0.0, 0.0]], np.stack((np.cos(ang), np.sin(ang)), axis=1) * (1.0 + 0.1 * np.cos(2 * ang))[:, None]))
    tri = np.stack((np.zeros(5, dtype=np.int64), 1 + np.arange(5), 1 + (np.arange(5) + 1) % 5), axis=1)
    _check_small(coords, {"triangle": tri}, "triangle", expectFan=1)
    tri2 = tri.copy()
    tri2[2] = tri2[2][[0, 2, 1]]
    _check_small(coords, {"triangle": tri2}, "triangle", expectFan=0)
    # triangle bow tie: two fans at node 2 -- the chained 8-byte records do not apply, the plan keeps the 16-byte ones
    coords = np.array([[0.0, 0.0], [1.0, 0.1], [1.1, 1.0], [2.2, 1.1], [2.3, 2.0], [0.1, 1.1]])
    tri = np.array([[0, 1, 2], [2, 3, 4], [0, 2, 5]])
    plan = _raw_plan("triangle", coords, {"triangle": tri})
    assert plan.info("fan_path") == 1 and plan.info("fan_compact") == 0
    plan.close()
    _check_small(coords, {"triangle": tri}, "triangle", expectFan=1)


@pytest.mark.parametrize("oriented", [True, False])
def test_unstructured_delaunay_triangulation(oriented):
    """an unstructured triangle mesh (Delaunay triangulation of random points: node valences 3..10+, no numbering
    locality): consistently oriented it takes the fan kernel (wide or chained records, whatever fits); with the
    triangulation's own mixed orientations the corners do not chain and the plan takes the accumulating kernel"""
    from scipy.spatial import Delaunay

    rng = np.random.default_rng(21)
    coords = rng.random((4000, 2)) * np.array([2.0, 1.0])
    tri = Delaunay(coords).simplices.astype(np.int64)
    a, b, c = coords[tri[:, 0]], coords[tri[:, 1]], coords[tri[:, 2]]
    area2 = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1])
    tri = tri[np.abs(area2) > 1e-9]  # drop slivers on the hull
    area2 = area2[np.abs(area2) > 1e-9]
    if oriented:
        tri[area2 < 0] = tri[area2 < 0][:, [0, 2, 1]]
    else:
        flip = rng.random(tri.shape[0]) < 0.5
        tri[flip] = tri[flip][:, [0, 2, 1]]
    used = np.unique(tri)
    assert used.size == coords.shape[0]
    N, numEl = coords.shape[0], tri.shape[0]
    conn = {"triangle": tri}
    case = (coords, conn, 1e-3 * rng.random((N, 2)), rng.random(numEl) + 0.5, rng.random(2 * N), np.array([0, 1, 9, 2 * N - 1]),
            np.array([1e-4, 0.0, 2e-4, -1e-4]))
    plan = _raw_plan("triangle", coords, conn)
    assert plan.info("rows_path") == 1 and plan.info("fan_path") == (1 if oriented else 0)
    assert plan.info("max_degree") >= 9
    _check_against_oracle(plan, "triangle", case, launchesExpected=1)
    plan.close()


@pytest.mark.parametrize("seed", range(8))
def test_random_meshes_numberings_and_boundary_conditions(seed):
    """randomised sweep: element type, mesh size, node renumbering, constitutive model, BC / load sets (duplicates, negative
    indices) drawn from the seed -- K pattern bit-exact, K / R / von Mises to 1e-12 against the oracle"""
    rng = np.random.default_rng(1000 + seed)
    elType = ["quad", "triangle", "quad9", "triangle6"][seed % 4]
    nx, ny = int(rng.integers(2, 40)), int(rng.integers(2, 40))
    if elType in ("quad9", "triangle6"):
        nx, ny = max(2, nx // 3), max(2, ny // 3)
    coords, conn = meshes.GENERATORS[elType](nx, ny)
    coords = coords + 0.2 / max(nx, ny) / 3 * rng.random(coords.shape)
    N, numEl = coords.shape[0], conn[elType].shape[0]
    states, thick, loads = 1e-3 * rng.random((N, 2)), rng.random(numEl) + 0.5, rng.random(2 * N)
    nbc = int(rng.integers(1, 12))
    bcDOF = rng.integers(-2 * N, 2 * N, nbc)
    bcDOF = np.concatenate((bcDOF, bcDOF[: nbc // 2]))  # duplicates: the occurrence count lands on the diagonal
    bcVal = rng.random(bcDOF.size) * 1e-4
    case = (coords, conn, states, thick, loads, bcDOF, bcVal)
    if rng.random() < 0.5:
        case = _renumbered((coords, conn, states, thick, loads, np.where(bcDOF < 0, bcDOF + 2 * N, bcDOF), bcVal), seed=seed)
    coords, conn, states, thick, loads, bcDOF, bcVal = case
    model = int(rng.integers(0, 2))
    plan = _raw_plan(elType, coords, conn)
    mat = make_cm(model).material()
    K, R = plan.assemble(coords, states, thick, mat, bcDOF, bcVal, True, loads)
    blocks = oracle_blocks(conn)
    bcPos = np.where(np.asarray(bcDOF) < 0, np.asarray(bcDOF) + 2 * N, bcDOF)
    Ko = fo.assemble_matrix(blocks, coords, states, thick, model, MAT["E"], MAT["nu"], bcPos, dropZeros=False)
    Ko.sort_indices()
    Ro = fo.assemble_residual(blocks, coords, states, thick, model, MAT["E"], MAT["nu"], loads, bcPos, bcVal)
    np.testing.assert_array_equal(plan.pattern()[0], Ko.indptr)
    np.testing.assert_array_equal(plan.pattern()[1], Ko.indices)
    assert relerr(K, Ko.data) < TOL and relerr(R, Ro) < TOL, (elType, nx, ny)
    vm = plan.function(coords, states, mat, 1, "max")
    assert relerr(vm, fo.compute_function(blocks, coords, states, fo.VON_MISES, "max", model, MAT["E"], MAT["nu"], MAT["rho"])[0]) < TOL
    plan.close()


def test_tables_without_cyclic_symmetry_leave_the_row_owner_path():
    """the rotated corner records need basis + rule invariant under cyclic renumbering; a 4-node block whose tables are
    in tensor order (0,1,2,3 = BL,BR,TL,TR) is not: the plan runs the scatter kernels for it and stays exact"""
    coords, conn = meshes.quad4_grid(9, 7)
    coords = coords + 0.01 * np.random.default_rng(3).random(coords.shape)
    perm = [0, 1, 3, 2]
    N_, dN, w = ELEMENT_FOR["quad"]().tables()
    connT = conn["quad"][:, perm]  # tensor-ordered connectivity goes with tensor-ordered tables
    N = coords.shape[0]
    rng = np.random.default_rng(6)
    states, thick, loads = 1e-3 * rng.random((N, 2)), rng.random(63) + 0.5, rng.random(2 * N)
    bcDOF, bcVal = np.array([0, 1, 5]), np.array([0.0, 1e-4, 0.0])
    plan = AssemblyPlan(N)
    plan.addBlock(N_[:, perm], dN[:, :, perm], w, connT)
    plan.finalize()
    assert plan.info("rows_path") == 0 and plan.info("fan_path") == 0 and plan.info("closed_form_q4") == 0
    K, R = plan.assemble(coords, states, thick, make_cm(0).material(), bcDOF, bcVal, True, loads)
    blocks = oracle_blocks(conn)
    Ko = fo.assemble_matrix(blocks, coords, states, thick, 0, MAT["E"], MAT["nu"], bcDOF, dropZeros=False)
    Ko.sort_indices()
    Ro = fo.assemble_residual(blocks, coords, states, thick, 0, MAT["E"], MAT["nu"], loads, bcDOF, bcVal)
    np.testing.assert_array_equal(plan.pattern()[1], Ko.indices)
    assert relerr(K, Ko.data) < TOL and relerr(R, Ro) < TOL
    # a rule with unequal weights breaks the symmetry as well; the same rule with the standard tables keeps it
    w2 = w * np.array([1.0, 1.1, 0.9, 1.0])
    plan2 = AssemblyPlan(N)
    plan2.addBlock(N_, dN, w2, conn["quad"])
    plan2.finalize()
    assert plan2.info("rows_path") == 0
    plan.close()
    plan2.close()


@pytest.mark.parametrize("order", [("quad", "triangle"), ("triangle", "quad")])
@pytest.mark.parametrize("balance", [0, 2])
def test_mixed_mesh_row_owner_matches_oracle(order, balance):
    """several element blocks: one row-owner launch per block (later blocks add to the rows the earlier ones wrote,
    the last one applies loads and BCs) -- still no atomics, 2 launches for 2 blocks"""
    coords, cq = meshes.quad4_grid(31, 23)
    coords = coords + 4e-3 * np.random.default_rng(2).random(coords.shape)
    col = np.tile(np.arange(31), 23)
    rest = cq["quad"][col >= 14]
    conn = {"quad": cq["quad"][col < 14], "triangle": np.vstack((rest[:, [0, 1, 2]], rest[:, [0, 2, 3]]))}
    conn = {k: conn[k] for k in order}
    N = coords.shape[0]
    numEl = sum(c.shape[0] for c in conn.values())
    rng = np.random.default_rng(11)
    states = 1e-3 * rng.random((N, 2))
    thick = rng.random(numEl) * 1e-2 + 1e-3
    loads = rng.random(2 * N)
    left = np.arange(0, N, 32)  # first node of every grid row
    bcDOF = np.concatenate((2 * left, 2 * left + 1, 2 * left[:2]))
    bcVal = rng.random(bcDOF.size) * 1e-4
    plan = AssemblyPlan(N)
    plan.setOption("balance_rows", balance)
    for elType, c in conn.items():
        plan.addBlock(*ELEMENT_FOR[elType]().tables(), c)
    plan.finalize()
    assert plan.info("rows_path") == 1 and plan.info("rows_balanced") == (1 if balance == 2 else 0)
    case = (coords, conn, states, thick, loads, bcDOF, bcVal)
    K1, R1 = _check_against_oracle(plan, None, case, launchesExpected=2)
    plan.setOption("assembly_path", 0)  # scatter kernels: memset-free init + one launch per block + BC kernel
    K0, R0 = plan.assemble(coords, states, thick, make_cm(0).material(), bcDOF, bcVal, True, loads)
    assert plan.lastLaunches == 4
    assert relerr(K1, K0) < 1e-13 and relerr(R1, R0) < 1e-13
    plan.close()


@pytest.mark.parametrize("scale", [1e-6, 1e6])
def test_coordinate_scale_invariance(scale):
    """the division-free gradients (hardware reciprocal seed + Newton steps) keep 1e-12 far away from unit-sized cells"""
    coords, conn, states, thick, loads, bcDOF, bcVal = _path_case("quad", 21)
    case = (coords * scale, conn, states * scale, thick, loads, bcDOF, bcVal * scale)
    plan = _raw_plan("quad", case[0], conn)
    _check_against_oracle(plan, "quad", case, launchesExpected=1)
    plan.close()


def test_high_degree_node_falls_back_to_scatter_kernels():
    """a node shared by 60 triangles has 61 neighbours: more than the row-owner kernel's shared-memory rows hold"""
    m = 60
    ang = 2 * np.pi * np.arange(m) / m
    coords = np.vstack(([[0.0, 0.0]], np.stack((np.cos(ang), np.sin(ang)), axis=1) * (1.0 + 0.1 * np.cos(3 * ang))[:, None]))
    conn = {"triangle": np.stack((np.zeros(m, dtype=np.int64), 1 + np.arange(m), 1 + (np.arange(m) + 1) % m), axis=1)}
    N = coords.shape[0]
    rng = np.random.default_rng(4)
    case = (coords, conn, 1e-3 * rng.random((N, 2)), rng.random(m) + 0.5, rng.random(2 * N), np.array([2, 3, 5]), np.array([1e-4, 0.0, 2e-4]))
    plan = _raw_plan("triangle", coords, conn)
    assert plan.info("max_degree") == 61 and plan.info("rows_path") == 0
    _check_against_oracle(plan, "triangle", case, launchesExpected=3)
    plan.close()


def test_write_once_path_is_deterministic(path="rows"):
    """write-once assembly has a fixed summation order: two runs are bit-identical"""
    coords, conn = meshes.quad4_grid(200, 150)
    N = coords.shape[0]
    states = 1e-3 * np.random.default_rng(0).random((N, 2))
    plan = _raw_plan("quad", coords, conn, path=path)
    mat = make_cm(0).material()
    K1, R1 = plan.assemble(coords, states, np.ones(30000), mat, applyBCs=False)
    K2, R2 = plan.assemble(coords, states, np.ones(30000), mat, applyBCs=False)
    np.testing.assert_array_equal(K1, K2)
    np.testing.assert_array_equal(R1, R2)
    plan2 = _raw_plan("quad", coords, conn, path=path)  # a second plan build gives the same program
    K3, R3 = plan2.assemble(coords, states, np.ones(30000), mat, applyBCs=False)
    np.testing.assert_array_equal(K1, K3)
    np.testing.assert_array_equal(R1, R3)


def test_isolated_nodes_get_minus_load():
    coords = np.array([[0.0, 0.0], [2.0, 0.1], [2.2, 1.0], [0.1, 1.5], [5.0, 5.0], [6.0, 7.0]])
    conn = {"quad": np.array([[0, 1, 2, 3]])}
    for path in ("rows", "scatter"):
        plan = _raw_plan("quad", coords, conn, path=path)
        loads = np.arange(12.0)
        _, R = plan.assemble(coords, np.zeros((6, 2)), np.ones(1), make_cm(0).material(), applyBCs=False, loads=loads)
        np.testing.assert_allclose(R, -loads, atol=1e-3)  # zero state: R = -loads everywhere
        np.testing.assert_array_equal(R[8:], -loads[8:])


# ---- edge cases ------------------------------------------------------------------------------------------------------------
def test_single_element_and_isolated_nodes():
    coords = np.array([[0.0, 0.0], [2.0, 0.1], [2.2, 1.0], [0.1, 1.5], [5.0, 5.0], [6.0, 7.0]])  # nodes 4, 5 unused
    conn = {"quad": np.array([[0, 1, 2, 3]])}
    m, p = build_problem(coords, conn, 0)
    K = p._assembleMatrix(p.states, applyBCs=False)
    assert K.shape == (12, 12) and K.nnz == 64
    np.testing.assert_array_equal(K.indptr, [0, 8, 16, 24, 32, 40, 48, 56, 64, 64, 64, 64, 64])
    N_, dN, w = ELEMENT_FOR["quad"]().tables()
    Ke = fo.element_jacobians(dN, w, coords[None, :4], np.zeros((1, 4, 2)), np.ones(1), 0, MAT["E"], MAT["nu"])
    assert relerr(K.toarray()[:8, :8], Ke[0]) < TOL
    p.addFixedBCToNodes("orphan", [4], [0], 0.0)
    from fempy_b200._lib import FembError

    with pytest.raises(FembError, match="belongs to no element"):
        p._assembleMatrix(p.states, applyBCs=True)


def test_bc_semantics_duplicates_and_empty():
    """rows zeroed but kept, +1.0 per occurrence on the diagonal, last value wins in R (AssemblyUtils.py:55-60,93)"""
    coords, conn = meshes.quad4_grid(3, 2)
    coords = coords + 1e-3 * np.random.default_rng(0).random(coords.shape)
    states = 1e-3 * np.random.default_rng(1).random(coords.shape)
    m, p = build_problem(coords, conn, 0, states=states)
    K0 = p._assembleMatrix(p.states, applyBCs=True)  # no BCs defined at all
    Kf = p._assembleMatrix(p.states, applyBCs=False)
    np.testing.assert_array_equal(K0.data, Kf.data) if False else None
    assert relerr(K0.data, Kf.data) < 1e-14
    p.BCs["a"] = {"DOF": [5, 2, 5], "Value": [0.5, 0.25, 0.75]}
    K = p._assembleMatrix(p.states)
    R = p._assembleResidual(p.states)
    np.testing.assert_array_equal(K.indptr, Kf.indptr)  # BC rows stay in the pattern
    Kd = K.toarray()
    assert Kd[5, 5] == 2.0 and Kd[2, 2] == 1.0
    assert np.count_nonzero(Kd[5]) == 1 and np.count_nonzero(Kd[2]) == 1
    assert np.count_nonzero(Kd[:, 5]) > 1  # columns are NOT eliminated
    u = states.reshape(-1)
    assert R[5] == u[5] - 0.75 and R[2] == u[2] - 0.25
    blocks = oracle_blocks(conn)
    Ko = fo.assemble_matrix(blocks, coords, states, np.ones(6), 0, MAT["E"], MAT["nu"], [5, 2, 5])
    assert relerr(Kd, Ko.toarray()) < TOL


def test_inverted_element_signed_determinant():
    """clockwise connectivity: detJ < 0 is used as is (LinAlg.py:89 has no abs)"""
    coords, conn = meshes.quad4_grid(2, 2)
    cw = {"quad": conn["quad"][:, ::-1].copy()}
    m, p = build_problem(coords, cw, 0)
    K = p._assembleMatrix(p.states, applyBCs=False)
    blocks = oracle_blocks(cw)
    Ko = fo.assemble_matrix(blocks, coords, np.zeros_like(coords), np.ones(4), 0, MAT["E"], MAT["nu"], dropZeros=False)
    Ko.sort_indices()
    assert relerr(K.data, Ko.data) < TOL
    assert K.diagonal().max() < 0
    assert p.computeFunction("Mass", "integrate", "sum") < 0


def test_result_arrays_come_from_the_pinned_pool_with_value_semantics():
    """every assemble call returns its own arrays (as the reference does); they live in page-locked memory that goes back
    to the pool when the last reference -- including views such as a scipy matrix's data -- is dropped"""
    import gc

    from fempy_b200 import _pinned

    coords, conn, states, thick, loads, bcDOF, bcVal = _path_case("quad", 21)
    plan = _raw_plan("quad", coords, conn)
    mat = make_cm(0).material()
    K1, R1 = plan.assemble(coords, states, thick, mat, bcDOF, bcVal, True, loads)
    K2, R2 = plan.assemble(coords, 2.0 * states, thick, mat, bcDOF, bcVal, True, loads)
    assert not np.shares_memory(K1, K2) and not np.shares_memory(R1, R2)
    np.testing.assert_array_equal(K1, K2)          # linear model: K does not depend on the state ...
    assert np.abs(R1 - R2).max() > 0               # ... R does, and the first result is still intact
    Ko = fo.assemble_residual(oracle_blocks(conn), coords, states, thick, 0, MAT["E"], MAT["nu"], loads, bcDOF, bcVal)
    assert relerr(R1, Ko) < TOL
    pool = _pinned.pool(None)
    M = csr_array((K1, plan.pattern()[1], plan.pattern()[0]), shape=(plan.numDOF, plan.numDOF))
    addr = K1.ctypes.data
    del K1
    gc.collect()
    assert not pool._free.get(M.data.nbytes)       # the matrix still holds the buffer
    del M
    gc.collect()
    assert addr in pool._free.get(8 * plan.nnz, [])
    K3, _ = plan.assemble(coords, states, thick, mat, bcDOF, bcVal, True, loads)
    assert K3.ctypes.data == addr                  # recycled, not re-allocated
    np.testing.assert_array_equal(K3, K2)
    plan.close()


def test_connectivity_from_device_memory():
    """femb_plan_add_block_dev: the mesh connectivity is taken from device memory; same plan, same values"""
    import torch

    coords, conn, states, thick, loads, bcDOF, bcVal = _path_case("quad", 25)
    el = ELEMENT_FOR["quad"]()
    d_conn = torch.from_numpy(conn["quad"]).cuda()
    plan = AssemblyPlan(coords.shape[0])
    plan.addBlockDev(*el.tables(), d_conn, conn["quad"].shape[0])
    plan.finalize()
    ref = _raw_plan("quad", coords, conn)
    np.testing.assert_array_equal(plan.pattern()[1], ref.pattern()[1])
    mat = make_cm(0).material()
    K, R = plan.assemble(coords, states, thick, mat, bcDOF, bcVal, True, loads)
    K0, R0 = ref.assemble(coords, states, thick, mat, bcDOF, bcVal, True, loads)
    np.testing.assert_array_equal(K, K0)
    np.testing.assert_array_equal(R, R0)
    bad = torch.full((3, 4), 10**6, dtype=torch.int64).cuda()
    with pytest.raises(Exception, match="outside"):
        AssemblyPlan(16).addBlockDev(*el.tables(), bad, 3)
    plan.close()
    ref.close()


def test_bad_inputs_raise():
    from fempy_b200._lib import FembError

    plan = AssemblyPlan(4)
    N, dN, w = ELEMENT_FOR["quad"]().tables()
    with pytest.raises(FembError, match="outside"):
        plan.addBlock(N, dN, w, np.array([[0, 1, 2, 7]]))
    with pytest.raises(FembError, match="unsupported element family"):
        plan.addBlock(N[:2], dN[:2], w[:2], np.array([[0, 1, 2, 3]]))  # 4 nodes with a 2-point rule: no such family
    with pytest.raises(FembError, match="not finalized"):
        plan.pattern()
    plan.addBlock(N, dN, w, np.array([[0, 1, 2, 3]]))
    plan.finalize()
    with pytest.raises(ValueError):
        plan.assemble(np.zeros((3, 2)), np.zeros((4, 2)), np.ones(1), make_cm(0).material())
    with pytest.raises(FembError, match="one displacement state per dimension"):
        AssemblyPlan(4, numDim=3, numStates=2)
    with pytest.raises(FembError, match="constitutive model"):  # a 2-D plan with a 3-D material
        plan.assemble(np.zeros((4, 2)), np.zeros((4, 2)), np.ones(1), Constitutive.Iso3D(1.0, 0.3, 1.0).material())
    plan.close()


def test_problem_level_nonlinear_newton_step():
    """Green-strain model: K and R at a non-zero state against the oracle (state-dependent Jacobian)"""
    coords, conn = meshes.quad9_grid(6, 5)
    N = coords.shape[0]
    rng = np.random.default_rng(2)
    states = 5e-2 * rng.random((N, 2))
    m, p = build_problem(coords, conn, 1, linear=False, states=states, bcs=[("Fixed", meshes.edge_nodes(coords, x=0.0), [0, 1], 0.0)])
    assert not p.isLinear
    K, R = p.assembleSystem()
    blocks = oracle_blocks(conn)
    bcDOF, bcVal = AssemblyUtils.convertBCDictToLists(p.getBCs())
    Ko = fo.assemble_matrix(blocks, coords, states, np.ones(30), 1, MAT["E"], MAT["nu"], bcDOF, nonlinear=True, dropZeros=False)
    Ko.sort_indices()
    Ro = fo.assemble_residual(blocks, coords, states, np.ones(30), 1, MAT["E"], MAT["nu"], None, bcDOF, bcVal, nonlinear=True)
    assert relerr(K.data, Ko.data) < TOL and relerr(R, Ro) < TOL


# ---- multi-GPU (needs >= 2 devices; the CPU suite covers the same logic over gloo) ------------------------------------
def test_multi_gpu_partitioned_assembly_matches_single_gpu():
    import subprocess
    import sys

    import torch

    n = min(torch.cuda.device_count(), 4)
    if n < 2:
        pytest.skip("needs at least 2 GPUs (tests/test_dist_cpu.py covers the partition/exchange logic over gloo)")
    from util import ROOT
    import os

    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "check_multi_gpu.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "MULTI_GPU_PARITY_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
