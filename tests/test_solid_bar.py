"""The 3-D and 1-D families: HexElement3D order 1-2 with Iso3D, LineElement1D order 1-3 with Iso1D.

tests/golden/solid_bar.npz holds what the unmodified reference computes (oracle/gen_golden.py: gen_solid_bar): element-level
K_e / R_e / functions for small and Green strain, and assembled systems on a hexahedron grid and on line meshes.  CPU tests:
host tables + oracle reproduce them.  GPU tests: the generic-dimension CUDA path (csrc/generic.cu) through the element-level
mirror and through AssemblyPlan -- pattern bit-exact, values to 1e-12.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from util import MAT, golden, relerr

from fempy_b200 import AssemblyPlan, Constitutive, Elements
from oracle import fem_oracle as fo

TOL = 1e-12
FAMILIES = {"hex1": lambda: Elements.HexElement3D(1), "hex2": lambda: Elements.HexElement3D(2), "line1": lambda: Elements.LineElement1D(1),
            "line2": lambda: Elements.LineElement1D(2), "line3": lambda: Elements.LineElement1D(3)}  # fmt: skip
MESHES = ["mesh_hex", "mesh_hex_nl", "mesh_line", "mesh_line3", "mesh_line4_nl"]


def model_of(name):
    return 2 if "hex" in name else 3


def constitutive(model, linear):
    return Constitutive.Iso3D(MAT["E"], MAT["nu"], MAT["rho"], linear=linear) if model == 2 else Constitutive.Iso1D(MAT["E"], MAT["rho"], MAT["t"], linear=linear)


def functions_of(model):
    return (("Von-Mises-Stress", fo.VON_MISES),) if model == 2 else (("Stress", fo.STRESS), ("Strain", fo.STRAIN))


def element_for_mesh(elType):
    return {"hexahedron": Elements.HexElement3D(1), "line": Elements.LineElement1D(1), "line3": Elements.LineElement1D(2), "line4": Elements.LineElement1D(3)}[elType]


# ---- CPU: tables + oracle against the reference --------------------------------------------------------------------
@pytest.mark.parametrize("family", sorted(FAMILIES))
@pytest.mark.parametrize("linear", [True, False])
def test_oracle_element_level_against_reference(family, linear):
    g = golden("solid_bar.npz")
    el, model = FAMILIES[family](), model_of(family)
    X, q, t = g[family + "_X"], g[family + "_q"], g[family + "_t"]
    theta = np.ones(X.shape[0]) if model == 2 else t
    _, dN, w = el.tables()
    key = f"{family}_{'lin' if linear else 'nl'}"
    Ke = fo.element_jacobians(dN, w, X, q, theta, model, MAT["E"], MAT["nu"], nonlinear=not linear)
    Re = fo.element_residuals(dN, w, X, q, theta, model, MAT["E"], MAT["nu"], nonlinear=not linear)
    assert relerr(Ke, g[key + "_Ke"]) < TOL and relerr(Re, g[key + "_Re"]) < TOL
    assert relerr(fo.element_function(dN, w, X, q, fo.MASS, "integrate", model, MAT["E"], MAT["nu"], MAT["rho"], not linear), g[key + "_massint"]) < TOL
    for fname, fid in functions_of(model):
        for red in (None, "mean", "max"):
            want = g[f"{key}_{fname}" + ("" if red is None else f"_{red}")]
            assert relerr(fo.element_function(dN, w, X, q, fid, red, model, MAT["E"], MAT["nu"], MAT["rho"], not linear), want) < TOL, (fname, red)


def _mesh(g, name):
    elType = str(g[name + "_eltype"][0])
    model, linear = (int(v) for v in g[name + "_meta"])
    el = element_for_mesh(elType)
    coords, conn, states = g[name + "_coords"], g[name + "_conn"], g[name + "_states"]
    theta = np.ones(conn.shape[0]) if model == 2 else g[name + "_scale"]
    return el, model, bool(linear), coords, conn, states, theta


@pytest.mark.parametrize("name", MESHES)
def test_oracle_assembled_against_reference(name):
    g = golden("solid_bar.npz")
    el, model, linear, coords, conn, states, theta = _mesh(g, name)
    N_, dN, w = el.tables()
    blocks = [fo.Block(N_, dN, w, conn)]
    K = fo.assemble_matrix(blocks, coords, states, theta, model, MAT["E"], MAT["nu"], g[name + "_bcDOF"], nonlinear=not linear)
    K.sort_indices()
    R = fo.assemble_residual(blocks, coords, states, theta, model, MAT["E"], MAT["nu"], g[name + "_loads"], g[name + "_bcDOF"], g[name + "_bcVal"], not linear)
    np.testing.assert_array_equal(K.indptr, g[name + "_indptr"])
    np.testing.assert_array_equal(K.indices, g[name + "_indices"])
    assert relerr(K.data, g[name + "_data"]) < TOL and relerr(R, g[name + "_R"]) < TOL
    fid = fo.VON_MISES if model == 2 else fo.STRESS
    assert relerr(fo.compute_function(blocks, coords, states, fid, "mean", model, MAT["E"], MAT["nu"], MAT["rho"], not linear)[0], g[name + "_func_mean"]) < TOL


def test_mirror_classes_and_registry():
    assert Elements.HexElement3D(2).numNodes == 27 and Elements.HexElement3D(2).quadratureOrder == 3 and Elements.HexElement3D(1).quadratureOrder == 2
    assert Elements.LineElement1D(3).numNodes == 4 and Elements.LineElement1D(3).quadratureOrder == 2
    with pytest.raises(ValueError):
        Elements.HexElement3D(3)
    assert isinstance(Elements.elementFromMeshioName("hexahedron", 3), Elements.HexElement3D)
    assert Elements.elementFromMeshioName("line3", 1).order == 2 and Elements.elementFromMeshioName("hexahedron27", 3) is None
    cm = Constitutive.Iso1D(1.0, 2.0, 3.0)
    assert cm.functionNames == ["Mass", "Stress", "Strain"] and cm.designVariables["Area"]["defaultValue"] == 3.0
    with pytest.raises(ValueError):
        Constitutive.Iso3D(1.0, 0.3, 1.0).getFunction("Stress")
    for el in (Elements.HexElement3D(1), Elements.HexElement3D(2), Elements.LineElement1D(3)):
        X = el.getReferenceElementCoordinates()
        np.testing.assert_allclose(el.computeShapeFunctions(X), np.eye(el.numNodes), atol=1e-12)


# ---- GPU: the generic-dimension CUDA path ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("family", sorted(FAMILIES))
@pytest.mark.parametrize("linear", [True, False])
def test_element_level_against_reference(family, linear):
    g = golden("solid_bar.npz")
    el, model = FAMILIES[family](), model_of(family)
    cm = constitutive(model, linear)
    X, q, t = g[family + "_X"], g[family + "_q"], g[family + "_t"]
    dvs = {} if model == 2 else {"Area": t}
    key = f"{family}_{'lin' if linear else 'nl'}"
    Ke, Re = el.computeResidualJacobians(q, X, dvs, cm), el.computeResiduals(q, X, dvs, cm)
    assert Ke.shape == g[key + "_Ke"].shape and Re.shape == g[key + "_Re"].shape
    assert relerr(Ke, g[key + "_Ke"]) < TOL and relerr(Re, g[key + "_Re"]) < TOL
    assert relerr(el.computeFunction(X, q, dvs, cm.getFunction("Mass"), "integrate"), g[key + "_massint"]) < TOL
    for fname, _ in functions_of(model):
        for red in (None, "mean", "max"):
            want = g[f"{key}_{fname}" + ("" if red is None else f"_{red}")]
            got = el.computeFunction(X, q, dvs, cm.getFunction(fname), red)
            assert got.shape == want.shape and relerr(got, want) < TOL, (fname, red)


@pytest.mark.gpu
@pytest.mark.parametrize("name", MESHES)
def test_assembled_against_reference(name):
    g = golden("solid_bar.npz")
    el, model, linear, coords, conn, states, theta = _mesh(g, name)
    plan = AssemblyPlan(coords.shape[0], coords.shape[1], coords.shape[1])
    plan.addBlock(*el.tables(), conn)
    plan.finalize()
    mat = constitutive(model, linear).material()
    K, R = plan.assemble(coords, states, theta, mat, g[name + "_bcDOF"], g[name + "_bcVal"], True, g[name + "_loads"])
    indptr, indices = plan.pattern()
    np.testing.assert_array_equal(indptr, g[name + "_indptr"])
    np.testing.assert_array_equal(indices, g[name + "_indices"])
    assert relerr(K, g[name + "_data"]) < TOL and relerr(R, g[name + "_R"]) < TOL
    fid = 1 if model == 2 else 2
    assert relerr(plan.function(coords, states, mat, fid, "mean"), g[name + "_func_mean"]) < TOL
    assert abs(plan.function(coords, states, mat, 0, "integrate").sum() - float(g[name + "_mass"])) < 1e-12 * float(g[name + "_mass"])
    # K-only / R-only, no BCs: against the oracle
    Kn, _ = plan.assemble(coords, states, theta, mat, None, None, False, None, wantR=False)
    N_, dN, w = el.tables()
    Ko = fo.assemble_matrix([fo.Block(N_, dN, w, conn)], coords, states, theta, model, MAT["E"], MAT["nu"], None, nonlinear=not linear, dropZeros=False)
    Ko.sort_indices()
    assert relerr(Kn, Ko.data) < TOL
    assert sp.csr_array((Kn, indices, indptr)).shape[0] == plan.numDOF
    with pytest.raises(Exception, match="constitutive model"):
        plan.assemble(coords, states, theta, Constitutive.IsoPlaneStress(1.0, 0.3, 1.0, 1.0).material(), None, None, False, None)
    plan.close()
