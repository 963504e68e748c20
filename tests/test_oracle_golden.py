"""Pin the CPU oracle (oracle/fem_oracle.py):
 (a) against the known-answer vectors held by the reference's own tests, and
 (b) against outputs of the unmodified reference stored under tests/golden/ (oracle/gen_golden.py).
Runs without a GPU and without /root/reference."""
import numpy as np
import pytest

from oracle import fem_oracle as fo
from util import ELEMENT_FOR, FAMILY_ELEMENT, MAT, golden, mesh_case, mesh_case_names, oracle_blocks, relerr

TOL = 1e-12


# ---- (a) the reference's own golden vectors ----------------------------------------------------------------------
def test_bc_dict_to_lists_reference_vector():
    """reference tests/testAssembleutils.py:31-42"""
    BCDict = {
        "BC1Name": {"DOF": [0, 1, 2], "Value": [0, 0, 0]},
        "BC2Name": {"DOF": [13, 46, 1385], "Value": [1.0, 1.0, -1.0]},
        "BC3Name": {"DOF": [837, 25], "Value": [1.0, 1.0]},
    }
    dof, val = fo.bc_dict_to_lists(BCDict)
    assert dof == [0, 1, 2, 13, 46, 1385, 837, 25]
    assert val == [0, 0, 0, 1.0, 1.0, -1.0, 1.0, 1.0]


def test_loads_dict_to_vector_reference_vector():
    """reference tests/testAssembleutils.py:44-53"""
    loadsDict = {
        "Load1Name": {"DOF": [0, 1, 2], "Value": [0, 0, -1]},
        "Load2Name": {"DOF": [4, 5, 8], "Value": [1.0, 1.0, -1.0]},
        "Load3Name": {"DOF": [2, 10], "Value": [3.0, 4.0]},
    }
    np.testing.assert_equal(
        fo.loads_dict_to_vector(loadsDict, 12), np.array([0.0, 0.0, 2.0, 0.0, 1.0, 1.0, 0.0, 0.0, -1.0, 0.0, 4.0, 0.0])
    )


COO_MATS = np.array(
    [
        [[1, 0, 0, 0], [0, 1, 0, 0], [0, 1, 1, 0], [0, 0, 0, 1]],
        [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 2, 2], [0, 0, 0, 2]],
        [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 2, 2], [0, 0, 0, 2]],
        [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 2, 2], [0, 0, 0, 2]],
    ]
)
COO_DOF = np.array([[0, 1, 2, 3], [2, 3, 4, 5], [4, 5, 6, 7], [6, 7, 0, 1]])
COO_ROWS = np.array([0, 1, 2, 2, 3, 2, 3, 4, 4, 5, 4, 5, 6, 6, 7, 6, 7, 0, 0, 1])
COO_COLS = np.array([0, 1, 1, 2, 3, 2, 3, 4, 5, 5, 4, 5, 6, 7, 7, 6, 7, 0, 1, 1])
COO_VALS = np.array([1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 2, 2, 2, 1, 1, 2, 2, 2])


def test_local_matrices_to_coo_reference_vector():
    """reference tests/testAssembleutils.py:55-81: row-major order, wrap-around DOFs, zero dropping"""
    r, c, v = fo.local_matrices_to_coo(COO_MATS, COO_DOF)
    np.testing.assert_equal(r, COO_ROWS)
    np.testing.assert_equal(c, COO_COLS)
    np.testing.assert_equal(v, COO_VALS)


def test_dof_map_reference_vector():
    """reference tests/testModel.py:99-104"""
    np.testing.assert_equal(fo.dof_from_node_inds(np.array([3, 8, 11, 23]), 2), [6, 7, 16, 17, 22, 23, 46, 47])


KNOWN_QUAD_K = 1.0e10 * np.array(
    [
        [4.4786, 2.3299, -1.5533, 0.7249, -3.1583, -2.0710, 0.2330, -0.9837],
        [2.3299, 4.6080, 0.0518, 1.5533, -2.0710, -2.8994, -0.3107, -3.2618],
        [-1.5533, 0.0518, 3.4430, -0.7766, -1.0873, -0.5695, -0.8025, 1.2944],
        [0.7249, 1.5533, -0.7766, 5.6435, -1.2426, -4.9704, 1.2944, -2.2263],
        [-3.1583, -2.0710, -1.0873, -1.2426, 5.7988, 2.5888, -1.5533, 0.7249],
        [-2.0710, -2.8994, -0.5695, -4.9704, 2.5888, 6.3166, 0.0518, 1.5533],
        [0.2330, -0.3107, -0.8025, 1.2944, -1.5533, 0.0518, 2.1228, -1.0355],
        [-0.9837, -3.2618, 1.2944, -2.2263, 0.7249, 1.5533, -1.0355, 3.9349],
    ]
)
KNOWN_QUAD_X = np.array([[0.0, 0.0], [2.0, 0.0], [2.0, 1.0], [0.0, 2.0]])


def test_known_quad_stiffness_matrix():
    """reference tests/testElements.py:34-45,80-82,169: plane-strain Q4, E=70e9, nu=0.3, t=1, rtol 1e-3"""
    N, dN, w = ELEMENT_FOR["quad"]().tables()
    Ke = fo.element_jacobians(dN, w, KNOWN_QUAD_X[None], np.zeros((1, 4, 2)), np.ones(1), fo.PLANE_STRAIN, 70e9, 0.3)
    np.testing.assert_allclose(Ke[0], KNOWN_QUAD_K, rtol=1e-3)


def two_quad_problem():
    """reference tests/testProblem.py:27-71 (coordinates perturbed as there, seed fixed here)"""
    rng = np.random.default_rng(0)
    coords = np.array([[0.0, 1.0], [0.0, 0.0], [1.2, 1.0], [0.9, 0.0], [2.0, 1.0], [2.0, 0.0]]) + 1e-14 * rng.random((6, 2))
    conn = {"quad": np.array([[0, 1, 3, 2], [2, 3, 5, 4]])}
    bcDOF, bcVal = [3, 11, 10], [1.0, 1.0, 0.1]
    loads = np.zeros(12)
    loads[5] = -10.0
    expectedSparsity = np.zeros((12, 12))
    expectedSparsity[:4, :8] = 1
    expectedSparsity[4:8, :] = 1
    expectedSparsity[8:, 4:] = 1
    return coords, conn, bcDOF, bcVal, loads, expectedSparsity


def test_two_quad_sparsity_rhs_and_solve():
    """reference tests/testProblem.py:73-125"""
    from scipy.sparse.linalg import spsolve

    coords, conn, bcDOF, bcVal, loads, sparsity = two_quad_problem()
    blocks = oracle_blocks(conn)
    states = np.zeros((6, 2))
    thick = np.full(2, 5e-3)
    K = fo.assemble_matrix(blocks, coords, states, thick, fo.PLANE_STRESS, 71.7e9, 0.33)
    np.testing.assert_equal(K.nonzero(), np.nonzero(sparsity))
    R = fo.assemble_residual(blocks, coords, states, thick, fo.PLANE_STRESS, 71.7e9, 0.33, loads, bcDOF, bcVal)
    expectedRHS = np.zeros(12)
    expectedRHS[[3, 5, 11, 10]] = [-1.0, 10.0, -1.0, -0.1]
    np.testing.assert_equal(R, expectedRHS)
    Kbc = fo.assemble_matrix(blocks, coords, states, thick, fo.PLANE_STRESS, 71.7e9, 0.33, bcDOF)
    sparsity[[3, 10, 11], :] = 0
    sparsity[[3, 10, 11], [3, 10, 11]] = 1.0
    np.testing.assert_equal(Kbc.nonzero(), np.nonzero(sparsity))
    u = spsolve(Kbc.tocsc(), -R)
    R2 = fo.assemble_residual(blocks, coords, u.reshape(6, 2), thick, fo.PLANE_STRESS, 71.7e9, 0.33, loads, bcDOF, bcVal)
    np.testing.assert_allclose(R2, 0, rtol=1e-6, atol=1e-6)
    for dof, val in zip(bcDOF, bcVal):
        np.testing.assert_approx_equal(u[dof], val, significant=10)


def test_two_quad_mass_functions():
    """reference tests/testProblem.py:127-143"""
    coords, conn, *_ = two_quad_problem()
    blocks = oracle_blocks(conn)
    rho = 2780
    mass = fo.compute_function(blocks, coords, np.zeros((6, 2)), fo.MASS, "integrate", fo.PLANE_STRESS, 71.7e9, 0.33, rho)[0]
    np.testing.assert_allclose(mass.sum(), rho * 2)
    mx = fo.compute_function(blocks, coords, np.zeros((6, 2)), fo.MASS, "max", fo.PLANE_STRESS, 71.7e9, 0.33, rho)[0]
    np.testing.assert_equal(mx, np.array([rho, rho]))


def test_iso_stress_matrices():
    """reference tests/testISOstresses.py:39-92 (D eps to 1e-12) -- the D matrices written out by hand"""
    E, nu = 70e9, 0.3
    D = fo.d_matrix(fo.PLANE_STRESS, E, nu)
    c = E / (1 - nu**2)
    np.testing.assert_allclose(D, [[c, c * nu, 0], [c * nu, c, 0], [0, 0, c * (1 - nu) / 2]], rtol=1e-15)
    D = fo.d_matrix(fo.PLANE_STRAIN, E, nu)
    c = E / ((1 + nu) * (1 - 2 * nu))
    np.testing.assert_allclose(D, [[c * (1 - nu), c * nu, 0], [c * nu, c * (1 - nu), 0], [0, 0, c * (1 - 2 * nu) / 2]], rtol=1e-15)


def test_det_inv_against_numpy():
    """reference tests/testLinAlg.py:39-85: det/inv of 1000 random 2x2 and 3x3 matrices, atol 1e-12"""
    rng = np.random.default_rng(0)
    for d in (2, 3):
        A = rng.random((1000, d, d))
        np.testing.assert_allclose(fo.det_stack(A), np.linalg.det(A), atol=1e-12)
        np.testing.assert_allclose(fo.inv_stack(A) @ A, np.broadcast_to(np.eye(d), A.shape), atol=1e-9)


# ---- (b) fixtures produced by the unmodified reference --------------------------------------------------------------
@pytest.mark.parametrize("family", sorted(FAMILY_ELEMENT))
@pytest.mark.parametrize("model", [0, 1])
@pytest.mark.parametrize("linear", [True, False])
def test_element_level_against_reference(family, model, linear):
    g = golden("elements.npz")
    N, dN, w = ELEMENT_FOR[FAMILY_ELEMENT[family]]().tables()
    X, q, t = g[family + "_X"], g[family + "_q"], g[family + "_t"]
    key = f"{family}_m{model}_{'lin' if linear else 'nl'}"
    Ke = fo.element_jacobians(dN, w, X, q, t, model, MAT["E"], MAT["nu"], nonlinear=not linear)
    Re = fo.element_residuals(dN, w, X, q, t, model, MAT["E"], MAT["nu"], nonlinear=not linear)
    assert relerr(Ke, g[key + "_Ke"]) < TOL
    assert relerr(Re, g[key + "_Re"]) < TOL
    vm = fo.element_function(dN, w, X, q, fo.VON_MISES, None, model, MAT["E"], MAT["nu"], MAT["rho"], not linear)
    assert relerr(vm, g[key + "_vm"]) < TOL
    vmm = fo.element_function(dN, w, X, q, fo.VON_MISES, "mean", model, MAT["E"], MAT["nu"], MAT["rho"], not linear)
    assert relerr(vmm, g[key + "_vmmean"]) < TOL
    mi = fo.element_function(dN, w, X, q, fo.MASS, "integrate", model, MAT["E"], MAT["nu"], MAT["rho"])
    assert relerr(mi, g[key + "_massint"]) < TOL


EXTRA_RULES = {"quad1": (1, 3, 4), "quad2": (2, 4), "quad3": (3,), "tri1": (2, 3), "tri2": (1, 3), "tri3": (1, 3)}


@pytest.mark.parametrize("family", sorted(FAMILY_ELEMENT))
def test_point_quantities_and_non_default_quadrature_against_reference(family):
    """oracle vs the reference's _computeFunctionEvaluationQuantities at arbitrary points (Element.py:275-368) and its
    element-level methods with `intOrder=` (points.npz, oracle/gen_golden.py: gen_points)"""
    g, pg = golden("elements.npz"), golden("points.npz")
    el = ELEMENT_FOR[FAMILY_ELEMENT[family]]()
    X, q, t = g[family + "_X"], g[family + "_q"], g[family + "_t"]
    pts = pg[family + "_pts"]
    N, dN = el.computeShapeFunctions(pts), el.computeShapeFunctionGradients(pts)
    detJ, G, up = fo.point_quantities(dN, X, q)
    P = pts.shape[0]
    assert relerr(detJ.reshape(-1), pg[family + "_JacDet"]) < TOL
    assert relerr(G.reshape(-1, 2, el.numNodes), pg[family + "_StateGradSens"]) < TOL
    assert relerr(up.reshape(-1, 2, 2), pg[family + "_StateGrad"]) < TOL
    assert relerr(np.einsum("pn,end->epd", N, X).reshape(-1, 2), pg[family + "_Coord"]) < TOL
    assert relerr(np.einsum("pn,ens->eps", N, q).reshape(-1, 2), pg[family + "_State"]) < TOL
    assert relerr(np.einsum("pdn,enk->epdk", dN, X).reshape(-1, 2, 2), pg[family + "_Jac"]) < TOL
    assert pg[family + "_DVs"].shape == (X.shape[0] * P,)
    for order in EXTRA_RULES[family]:
        _, dNo, w = el.tables(order)
        key = f"{family}_int{order}"
        assert relerr(fo.element_jacobians(dNo, w, X, q, t, 0, MAT["E"], MAT["nu"]), pg[key + "_Ke"]) < TOL, key
        assert relerr(fo.element_residuals(dNo, w, X, q, t, 0, MAT["E"], MAT["nu"]), pg[key + "_Re"]) < TOL, key
        assert relerr(fo.element_jacobians(dNo, w, X, q, t, 1, MAT["E"], MAT["nu"], nonlinear=True), pg[key + "_Ke_nl"]) < TOL, key
        assert relerr(fo.element_function(dNo, w, X, q, fo.VON_MISES, None, 0, MAT["E"], MAT["nu"], MAT["rho"]), pg[key + "_vm"]) < TOL, key
        assert relerr(fo.element_function(dNo, w, X, q, fo.MASS, "integrate", 0, MAT["E"], MAT["nu"], MAT["rho"]), pg[key + "_massint"]) < TOL, key


@pytest.mark.parametrize("family", sorted(FAMILY_ELEMENT))
@pytest.mark.parametrize("model", [0, 1])
def test_ks_element_reduction_against_reference(family, model):
    """Element.py:261-265 + Utils/KSAgg.py:25-58: "ksmax" and "ksmin" both give the KS max (rho = 100)"""
    g, ks = golden("elements.npz"), golden("ks.npz")
    N, dN, w = ELEMENT_FOR[FAMILY_ELEMENT[family]]().tables()
    X, q = g[family + "_X"], g[family + "_q"]
    for tag, Emod in (("stiff", MAT["E"]), ("soft", 2.0)):
        for red in ("ksmax", "ksmin"):
            v = fo.element_function(dN, w, X, q, fo.VON_MISES, red, model, Emod, MAT["nu"], MAT["rho"])
            assert relerr(v, ks[f"{family}_m{model}_E{tag}_{red}"]) < TOL, (tag, red)
    np.testing.assert_array_equal(ks[f"{family}_m{model}_Esoft_ksmax"], ks[f"{family}_m{model}_Esoft_ksmin"])


@pytest.mark.parametrize("name", mesh_case_names(golden("meshes.npz")))
def test_assembled_meshes_against_reference(name):
    g = golden("meshes.npz")
    c = mesh_case(g, name)
    blocks = oracle_blocks(c["conn"])
    N = c["coords"].shape[0]
    bc = (c["bcDOF"], c["bcVal"]) if c["applyBCs"] else (None, None)
    K = fo.assemble_matrix(blocks, c["coords"], c["states"], c["thick"], c["model"], MAT["E"], MAT["nu"], bc[0], nonlinear=not c["linear"])
    K.sort_indices()
    # the fixtures use perturbed coordinates, so the reference's value-dependent pattern is the structural one
    ip, ix = fo.structural_pattern(blocks, N, 2)
    np.testing.assert_array_equal(c["indptr"], ip)
    np.testing.assert_array_equal(c["indices"], ix)
    np.testing.assert_array_equal(K.indptr, ip)
    np.testing.assert_array_equal(K.indices, ix)
    assert relerr(K.data, c["data"]) < TOL
    R = fo.assemble_residual(blocks, c["coords"], c["states"], c["thick"], c["model"], MAT["E"], MAT["nu"], c["loads"], bc[0], bc[1],
                             nonlinear=not c["linear"])
    assert relerr(R, c["R"]) < TOL
    for red in ("mean", "max", "min", "sum", "integrate"):
        vals = fo.compute_function(blocks, c["coords"], c["states"], fo.VON_MISES, red, c["model"], MAT["E"], MAT["nu"], MAT["rho"], not c["linear"])
        for v, k in zip(vals, c["conn"]):
            assert relerr(v, g[f"{name}_vm_{red}_{k}"]) < TOL


def test_lbracket_against_reference():
    """BASELINE config 1: Examples/LBracketExample.py vertical-load case"""
    g = golden("lbracket.npz")
    E, nu, rho, t = g["material"]
    conn = {"quad": g["conn"].astype(np.int64)}
    blocks = oracle_blocks(conn)
    coords = g["coords"]
    N = coords.shape[0]
    assert (N, conn["quad"].shape[0]) == (8385, 8184)
    top, right = g["top"], g["right"]
    bcDOF = np.stack((2 * top, 2 * top + 1), axis=1).reshape(-1)
    loads = np.zeros(2 * N)
    loads[2 * right + 1] += -1e5 / len(right)
    thick = np.full(8184, t)
    ip, ix = fo.structural_pattern(blocks, N, 2)
    np.testing.assert_array_equal(g["indptr"], ip)
    np.testing.assert_array_equal(g["indices"], ix)
    assert len(ix) == 297028 and len(bcDOF) == 82 and len(right) == 41
    K = fo.assemble_matrix(blocks, coords, np.zeros((N, 2)), thick, fo.PLANE_STRESS, E, nu, bcDOF)
    K.sort_indices()
    np.testing.assert_array_equal(K.indices, ix)
    assert relerr(K.data, g["data"]) < TOL
    R = fo.assemble_residual(blocks, coords, np.zeros((N, 2)), thick, fo.PLANE_STRESS, E, nu, loads, bcDOF, np.zeros(82))
    assert relerr(R, g["R"]) < TOL
    vm = fo.compute_function(blocks, coords, g["u"], fo.VON_MISES, "mean", fo.PLANE_STRESS, E, nu, rho)[0]
    assert relerr(vm, g["vm_mean"]) < TOL
    np.testing.assert_allclose(vm.max(), 1.632125596200e09, rtol=1e-9)
    mass = fo.compute_function(blocks, coords, g["u"], fo.MASS, "integrate", fo.PLANE_STRESS, E, nu, rho)[0].sum()
    np.testing.assert_allclose(mass, 1.779200000000e03, rtol=1e-9)
    np.testing.assert_allclose(np.abs(g["u"]).max(), 3.285941388626e-02, rtol=1e-9)
