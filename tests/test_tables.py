"""Host tables (basis functions, quadrature, node orderings) against the reference's tables
(golden fixture generated from FEMpy/Elements/*.py by oracle/gen_golden.py)."""
import numpy as np
import pytest

from util import ELEMENT_FOR, FAMILY_ELEMENT, golden


@pytest.mark.parametrize("family", sorted(FAMILY_ELEMENT))
def test_tables_match_reference(family):
    g = golden("tables.npz")
    el = ELEMENT_FOR[FAMILY_ELEMENT[family]]()
    pts = el.getIntegrationPointCoords()
    np.testing.assert_array_equal(pts, g[family + "_pts"])
    np.testing.assert_array_equal(el.getIntegrationPointWeights(), g[family + "_w"])
    np.testing.assert_allclose(el.computeShapeFunctions(pts), g[family + "_N"], rtol=0, atol=5e-15)
    np.testing.assert_allclose(el.computeShapeFunctionGradients(pts), g[family + "_dN"], rtol=0, atol=5e-14)
    np.testing.assert_allclose(el.getReferenceElementCoordinates(), g[family + "_refcoords"], rtol=0, atol=1e-15)


@pytest.mark.parametrize("family", sorted(FAMILY_ELEMENT))
def test_partition_of_unity_and_gradient_sum(family):
    """Shape functions sum to one and their gradients to zero (reference tests/testElements.py:96-98)."""
    el = ELEMENT_FOR[FAMILY_ELEMENT[family]]()
    rng = np.random.default_rng(1)
    pts = rng.random((5, 2)) * 0.5
    np.testing.assert_allclose(el.computeShapeFunctions(pts).sum(axis=1), 1.0, atol=1e-12)
    np.testing.assert_allclose(el.computeShapeFunctionGradients(pts).sum(axis=2), 0.0, atol=1e-11)


@pytest.mark.parametrize("family", sorted(FAMILY_ELEMENT))
def test_kronecker_at_reference_nodes(family):
    el = ELEMENT_FOR[FAMILY_ELEMENT[family]]()
    X = el.getReferenceElementCoordinates()
    np.testing.assert_allclose(el.computeShapeFunctions(X), np.eye(el.numNodes), atol=1e-12)


def test_element_sizes_and_errors():
    from fempy_b200 import Elements

    assert Elements.QuadElement2D(2).numNodes == 9 and Elements.QuadElement2D(2).quadratureOrder == 3
    assert Elements.QuadElement2D(1).quadratureOrder == 2 and Elements.QuadElement2D(3).quadratureOrder == 4
    assert Elements.TriElement2D(1).quadratureOrder == 1 and Elements.TriElement2D(3).numNodes == 10
    with pytest.raises(ValueError):
        Elements.QuadElement2D(4)
    with pytest.raises(ValueError):
        Elements.TriElement2D(0)
    assert Elements.elementFromMeshioName("quad8", 2) is None
    assert Elements.elementFromMeshioName("quad9", 2).order == 2
    assert Elements.elementFromMeshioName("hexahedron", 3).numNodes == 8 and Elements.elementFromMeshioName("tetra", 3) is None
