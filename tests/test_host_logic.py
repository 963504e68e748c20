"""Host-side mirror of the reference's Model / Problem / AssemblyUtils interface (no GPU needed)."""
import numpy as np
import pytest

from fempy_b200 import AssemblyUtils, Constitutive, meshes
from hostmirror import FEMpyModel
from util import build_problem, make_cm


def test_bc_and_load_converters_reference_vectors():
    """reference tests/testAssembleutils.py:31-53"""
    BCDict = {
        "BC1Name": {"DOF": [0, 1, 2], "Value": [0, 0, 0]},
        "BC2Name": {"DOF": [13, 46, 1385], "Value": [1.0, 1.0, -1.0]},
        "BC3Name": {"DOF": [837, 25], "Value": [1.0, 1.0]},
    }
    dof, val = AssemblyUtils.convertBCDictToLists(BCDict)
    assert dof == [0, 1, 2, 13, 46, 1385, 837, 25] and val == [0, 0, 0, 1.0, 1.0, -1.0, 1.0, 1.0]
    loadsDict = {
        "Load1Name": {"DOF": [0, 1, 2], "Value": [0, 0, -1]},
        "Load2Name": {"DOF": [4, 5, 8], "Value": [1.0, 1.0, -1.0]},
        "Load3Name": {"DOF": [2, 10], "Value": [3.0, 4.0]},
    }
    np.testing.assert_equal(
        AssemblyUtils.convertLoadsDictToVector(loadsDict, 12), np.array([0.0, 0.0, 2.0, 0.0, 1.0, 1.0, 0.0, 0.0, -1.0, 0.0, 4.0, 0.0])
    )


def test_apply_bcs_to_coo_and_vector():
    """AssemblyUtils.py:26-94 semantics (the example at the bottom of that file)"""
    rows, cols = np.array([0, 0, 1, 2]), np.array([1, 2, 2, 0])
    vals = np.array([1, 0.5, 2, 3])
    r, c, v = AssemblyUtils.applyBCsToMatrix(rows, cols, vals, np.array([1, 2]))
    np.testing.assert_equal(r, [0, 0, 1, 2, 1, 2])
    np.testing.assert_equal(c, [1, 2, 2, 0, 1, 2])
    np.testing.assert_equal(v, [1, 0.5, 0, 0, 1, 1])
    vec = AssemblyUtils.applyBCsToVector(np.arange(4.0), np.array([5.0, 6.0, 7.0, 8.0]), [1, 2], [10.0, 15.0])
    np.testing.assert_equal(vec, [0.0, -4.0, -8.0, 3.0])


def test_model_dof_map_and_registry():
    """reference tests/testModel.py:90-104 + Model.py:123-151"""
    coords, conn = meshes.quad4_grid(4, 6)
    m = FEMpyModel(make_cm(0), nodeCoords=coords, connectivity=conn)
    np.testing.assert_equal(m.getDOFfromNodeInds(np.array([3, 8, 11, 23])), [6, 7, 16, 17, 22, 23, 46, 47])
    assert m.getDOFfromNodeInds(np.array([3, 8, 11, 23])).dtype == np.int64
    assert m.numNodes == 35 and m.numDOF == 70 and m.numElements == 24 and m.numDim == 2
    el = m.elements["quad"]
    assert el["DOF"].shape == (24, 8) and el["numElements"] == 24
    np.testing.assert_equal(el["elementIDs"], np.arange(24))
    np.testing.assert_equal(el["DOF"][:, 0::2], 2 * conn["quad"])
    np.testing.assert_equal(m.dvs["Thickness"], np.ones(24))
    m.addFixedBCToNodes("Fixed", [0, 5], dof=[0, 1], value=0.0)
    assert m.BCs["Fixed"]["DOF"] == [0, 1, 10, 11]
    np.testing.assert_equal(m.getElementCoordinates("quad")[3], coords[conn["quad"][3]])


def test_model_mixed_types_and_errors():
    coords, cq = meshes.quad4_grid(4, 2)
    conn = {"quad": cq["quad"][:4], "triangle": np.vstack((cq["quad"][4:, [0, 1, 2]], cq["quad"][4:, [0, 2, 3]])), "vertex": np.array([[0]])}
    with pytest.warns(UserWarning, match="not supported"):
        m = FEMpyModel(make_cm(1), nodeCoords=coords, connectivity=conn)
    assert list(m.elements) == ["quad", "triangle"] and m.numElements == 12
    np.testing.assert_equal(m.elements["triangle"]["elementIDs"], np.arange(4, 12))
    np.testing.assert_equal(m.getElementDVs("triangle")["Thickness"], np.ones(8))
    with pytest.raises(ValueError, match="expected shape"):
        m.setDesignVariables({"Thickness": np.ones(3)})
    with pytest.raises(ValueError, match="2D constitutive"):
        FEMpyModel(make_cm(0), nodeCoords=np.array([[0.0, 0], [1, 0], [2, 0]]), connectivity={"quad": np.array([[0, 1, 2, 0]])})
    with pytest.raises(Exception, match="both needed"):
        FEMpyModel(make_cm(0))


def test_problem_bcs_loads_and_flags():
    """reference FEMpy/Problem.py:357-483, 200-280"""
    coords, conn = meshes.quad4_grid(3, 3)
    m, p = build_problem(coords, conn, 0)
    m.addFixedBCToNodes("ModelBC", [0], dof=[0, 1], value=0.0)
    p.addFixedBCToNodes(name="YFixed", nodeInds=[1, 5], dof=1, value=1.0)
    p.addFixedBCToNodes(name="XFixed", nodeInds=5, dof=0, value=0.1)
    p.addLoadToNodes(name="L", nodeInds=[2, 3], dof=[0, 1], value=[4.0, -2.0], totalLoad=True)
    assert p.BCs["YFixed"] == {"DOF": [3, 11], "Value": [1.0, 1.0]}
    assert p.BCs["XFixed"] == {"DOF": [10], "Value": [0.1]}
    assert p.loads["L"] == {"DOF": [4, 5, 6, 7], "Value": [2.0, -1.0, 2.0, -1.0]}
    assert list(p.getBCs()) == ["ModelBC", "YFixed", "XFixed"]
    with pytest.raises(Exception, match="one value per entry"):
        p.addFixedBCToNodes("bad", [0], dof=[0, 1], value=[1.0])
    with pytest.raises(NotImplementedError):
        p.addBodyLoad("g", [0, -9.81])
    p.markResUpToDate(), p.markJacUpToDate()
    p.incrementState(np.ones(p.numDOF))
    assert not p.resUpToDate and p.jacUpToDate  # linear model keeps its Jacobian (Problem.py:215-217)
    np.testing.assert_equal(p.getElementStates("quad"), np.ones((9, 4, 2)))
    p.reset()
    assert p.states.sum() == 0


def test_constitutive_mirror():
    """reference tests/testConstitutive.py:93-107: bad function name raises ValueError"""
    cm = Constitutive.IsoPlaneStress(70e9, 0.3, 2700.0, 1e-2)
    assert cm.numStates == 2 and cm.numStrains == 3 and cm.isLinear and cm.functionNames == ["Mass", "Von-Mises-Stress"]
    assert cm.designVariables["Thickness"]["defaultValue"] == 1e-2
    with pytest.raises(ValueError, match="not a valid function name"):
        cm.getFunction("Pressure")
    assert cm.getFunction("von-mises-stress").funcId == 1
    mat = Constitutive.IsoPlaneStrain(1.0, 0.25, 3.0, 1.0, linear=False).material()
    assert (mat.E, mat.nu, mat.rho, mat.model, mat.nonlinear) == (1.0, 0.25, 3.0, 1, 1)


def test_mesh_generators():
    coords, conn = meshes.quad4_grid(512, 512)
    assert coords.shape == (263169, 2) and conn["quad"].shape == (262144, 4)
    np.testing.assert_equal(conn["quad"][0], [0, 1, 514, 513])
    assert coords[:, 0].max() == 2.0
    c9, q9 = meshes.quad9_grid(3, 2)
    assert c9.shape == (35, 2)
    np.testing.assert_equal(q9["quad9"][0], [0, 2, 16, 14, 1, 9, 15, 7, 8])
    ct, t3 = meshes.tri3_grid(3, 2)
    assert t3["triangle"].shape == (12, 3)
    np.testing.assert_equal(t3["triangle"][0], [0, 1, 5])
    np.testing.assert_equal(t3["triangle"][6], [0, 5, 4])
    # every element counter-clockwise: positive signed area
    for c, k in ((coords, conn["quad"][:50]), (c9, q9["quad9"][:, :4]), (ct, t3["triangle"])):
        x, y = c[k, 0], c[k, 1]
        area = 0.5 * (x * np.roll(y, -1, axis=1) - np.roll(x, -1, axis=1) * y).sum(axis=1)
        assert (area > 0).all()
