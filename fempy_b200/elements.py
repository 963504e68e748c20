"""Host-side mirror of the reference's element families
(reference FEMpy/Elements/Element.py:32-486, QuadElement2D.py:42-186, TriElement2D.py:69-169, HexElement3D.py, LineElement1D.py).

The table generators (shape functions, gradients, quadrature) are host code; the batched per-element
methods -- `computeResidualJacobians`, `computeResiduals`, `computeFunction` -- keep the reference's
names, argument order and return shapes but run on the GPU through the C ABI: element arrays
(E, n, d) are flattened into a throw-away plan with E*n nodes.
"""
import numpy as np

from . import tables as T
from .plan import AssemblyPlan


class Element:
    """Base class: sizes + the batched element-level seams (Element.py:163-486)."""

    def __init__(self, numNodes, numDim, quadratureOrder, numStates=None):
        if numDim not in [1, 2, 3]:
            raise ValueError(f"Sorry, FEMpy doesn't support {numDim}-dimensional problems")
        self.numNodes = numNodes
        self.numDim = numDim
        self.numStates = numStates if numStates is not None else numDim
        self.numDOF = self.numNodes * self.numStates
        self.quadratureOrder = quadratureOrder
        self.name = f"{self.numNodes}Node-{self.numStates}Disp-{self.numDim}D-Element"

    # -- tables, implemented by the families ----------------------------------------------------
    def computeShapeFunctions(self, paramCoords):
        raise NotImplementedError

    def computeShapeFunctionGradients(self, paramCoords):
        raise NotImplementedError

    def getIntegrationPointWeights(self, order=None):
        raise NotImplementedError

    def getIntegrationPointCoords(self, order=None):
        raise NotImplementedError

    def getReferenceElementCoordinates(self):
        raise NotImplementedError

    def tables(self, intOrder=None):
        """(N, dNdxi, w) at the element's Gauss points."""
        order = self.quadratureOrder if intOrder is None else intOrder
        pts = np.atleast_2d(self.getIntegrationPointCoords(order))
        return self.computeShapeFunctions(pts), self.computeShapeFunctionGradients(pts), self.getIntegrationPointWeights(order)

    # -- batched element-level methods on the GPU ---------------------------------------------------
    def _flatPlan(self, nodeCoords, intOrder):
        nodeCoords = np.ascontiguousarray(nodeCoords, dtype=np.float64)
        numElements = nodeCoords.shape[0]
        if nodeCoords.shape[1:] != (self.numNodes, self.numDim):
            raise ValueError("nodeCoords must be numElements x numNodes x numDim")
        plan = AssemblyPlan(numElements * self.numNodes, self.numDim, self.numStates)
        N, dN, w = self.tables(intOrder)
        conn = np.arange(numElements * self.numNodes, dtype=np.int64).reshape(numElements, self.numNodes)
        plan.addBlock(N, dN, w, conn)
        plan.finalize()
        return plan, nodeCoords.reshape(-1, self.numDim)

    def _thickness(self, designVars, numElements):
        """the per-element volume scaling (ConstitutiveModel.computeVolumeScaling): thickness in 2-D, cross-sectional area
        in 1-D (Iso1D.py), 1 in 3-D (Iso3D.py has no design variable)"""
        name = {1: "Area", 2: "Thickness"}.get(self.numDim)
        vals = 1.0 if name is None else designVars[name]
        return np.ascontiguousarray(np.broadcast_to(np.asarray(vals, dtype=np.float64), (numElements,)))

    def computeResidualJacobians(self, nodeStates, nodeCoords, designVars, constitutiveModel, intOrder=None):
        """(E, n*s, n*s) element stiffness matrices (Element.py:426-486)."""
        plan, X = self._flatPlan(nodeCoords, intOrder)
        q = np.ascontiguousarray(nodeStates, dtype=np.float64).reshape(-1, self.numStates)
        numElements = X.shape[0] // self.numNodes
        Ke, _ = plan.elementBlocks(X, q, self._thickness(designVars, numElements), constitutiveModel.material(), wantR=False)
        plan.close()
        return Ke.reshape(numElements, self.numDOF, self.numDOF)

    def computeResiduals(self, nodeStates, nodeCoords, designVars, constitutiveModel, intOrder=None):
        """(E, n, s) element residuals (Element.py:370-424)."""
        plan, X = self._flatPlan(nodeCoords, intOrder)
        q = np.ascontiguousarray(nodeStates, dtype=np.float64).reshape(-1, self.numStates)
        numElements = X.shape[0] // self.numNodes
        _, Re = plan.elementBlocks(X, q, self._thickness(designVars, numElements), constitutiveModel.material(), wantK=False)
        plan.close()
        return Re.reshape(numElements, self.numNodes, self.numStates)

    def computeFunction(self, nodeCoords, nodeStates, elementDVs, function, elementReductionType, intOrder=None):
        """Element function values (Element.py:163-273); note the argument order differs from the two
        methods above, as in the reference.  `function` is what `constitutiveModel.getFunction` returned: the
        functions of this package's models carry an id and are evaluated AND reduced in one fused kernel; any other
        callable f(states, stateGradients, coords, dvs) -- e.g. the closures the reference's models return -- is evaluated
        on the host from the point quantities the device computes (Coord, State, StateGrad, JacDet)."""
        if elementReductionType is not None:
            assert elementReductionType.lower() in ["sum", "mean", "integrate", "min", "max", "ksmax", "ksmin"], \
                "elementReductionType not valid"
        if not hasattr(function, "funcId"):
            return self._computeCallable(nodeCoords, nodeStates, elementDVs, function, elementReductionType, intOrder)
        plan, X = self._flatPlan(nodeCoords, intOrder)
        q = np.ascontiguousarray(nodeStates, dtype=np.float64).reshape(-1, self.numStates)
        numElements = X.shape[0] // self.numNodes
        vals = plan.function(X, q, function.constitutiveModel.material(), function.funcId, elementReductionType)
        plan.close()
        return vals.reshape(numElements, -1) if elementReductionType is None else vals

    def _computeCallable(self, nodeCoords, nodeStates, elementDVs, function, elementReductionType, intOrder):
        order = self.quadratureOrder if intOrder is None else intOrder
        w = self.getIntegrationPointWeights(order)
        pts = np.atleast_2d(self.getIntegrationPointCoords(order))
        numElements, numPoints = np.shape(nodeCoords)[0], len(w)
        pq = self._computeFunctionEvaluationQuantities(pts, nodeStates, nodeCoords, elementDVs, ["Coord", "State", "StateGrad", "DVs", "JacDet"])
        values = np.asarray(function(pq["State"], pq["StateGrad"], pq["Coord"], pq["DVs"]), dtype=np.float64).reshape(numElements, numPoints)
        if elementReductionType is None:
            return values
        red = elementReductionType.lower()
        if red in ("sum", "mean", "min", "max"):
            return {"sum": np.sum, "mean": np.average, "min": np.min, "max": np.max}[red](values, axis=1)
        if red == "integrate":
            return np.einsum("ep,ep,p->e", values, pq["JacDet"].reshape(numElements, numPoints), w)
        # "ksmax" and "ksmin" both take the reference's "ksmax" branch (Element.py:261-265): max + log(mean(exp(rho (v - max)))) / rho
        rho = 100.0
        m = values.max(axis=1)
        return m + np.log(np.exp(rho * (values - m[:, None])).sum(axis=1) / numPoints) / rho

    def _computeFunctionEvaluationQuantities(self, paramCoords, nodeStates, nodeCoords, designVars, quantities):
        """Element._computeFunctionEvaluationQuantities (Element.py:275-368) with the per-element products on the device
        (femb_point_quantities): N / NPrimeParam are host tables, Coord, State, StateGrad, Jac, JacDet, StateGradSens come
        back as (numElements*numPoints, ...) arrays, DVs as a dict of per-point arrays.  (The reference never returns
        "JacInv": its test for that key is case-mangled, Element.py:338.)"""
        paramCoords = np.atleast_2d(np.asarray(paramCoords, dtype=np.float64))
        nodeCoords = np.ascontiguousarray(nodeCoords, dtype=np.float64)
        numElements, numPoints = nodeCoords.shape[0], paramCoords.shape[0]
        if nodeCoords.shape[1:] != (self.numNodes, self.numDim):
            raise ValueError("nodeCoords must be numElements x numNodes x numDim")
        lower = [q.lower() for q in quantities]
        out = {}
        N = self.computeShapeFunctions(paramCoords)
        dN = self.computeShapeFunctionGradients(paramCoords)
        names = {"coord": "Coord", "state": "State", "stategrad": "StateGrad", "jac": "Jac", "jacdet": "JacDet", "stategradsens": "StateGradSens"}
        want = tuple(names[q] for q in names if q in lower)
        if want:
            plan = AssemblyPlan(numElements * self.numNodes, self.numDim, self.numStates)
            plan.setOption("points_only", 1)
            conn = np.arange(numElements * self.numNodes, dtype=np.int64).reshape(numElements, self.numNodes)
            plan.addBlock(N, dN, np.ones(numPoints), conn)
            plan.finalize()
            q = np.ascontiguousarray(nodeStates, dtype=np.float64).reshape(-1, self.numStates)
            out.update(plan.pointQuantities(nodeCoords.reshape(-1, self.numDim), q, want))
            plan.close()
        if "nprimeparam" in lower:
            out["NPrimeParam"] = dN
        if "dvs" in lower:
            out["DVs"] = {name: np.ascontiguousarray(np.repeat(vals, numPoints)) for name, vals in designVars.items()}
        return out


class QuadElement2D(Element):
    """Lagrange quadrilateral of order 1-3 in meshio node order (QuadElement2D.py:42-186)."""

    def __init__(self, order=1, numStates=None, quadratureOrder=None):
        if order not in [1, 2, 3]:
            raise ValueError("Quad elements only support orders 1, 2 and 3")
        self.order = order
        if quadratureOrder is None:
            quadratureOrder = int(np.ceil((2 * order + 1) / 2))
        super().__init__((order + 1) ** 2, numDim=2, quadratureOrder=quadratureOrder, numStates=numStates)
        self.name = f"Order{self.order}-LagrangeQuad"
        self.shapeFuncToNodeOrder = T.QUAD_NODE_ORDER[order]

    def computeShapeFunctions(self, paramCoords):
        N = T.lagrange2d(paramCoords[:, 0], paramCoords[:, 1], self.order + 1)
        return np.ascontiguousarray(N[:, self.shapeFuncToNodeOrder])

    def computeShapeFunctionGradients(self, paramCoords):
        dN = T.lagrange2d_deriv(paramCoords[:, 0], paramCoords[:, 1], self.order + 1)
        return np.ascontiguousarray(dN[:, :, self.shapeFuncToNodeOrder])

    def getIntegrationPointWeights(self, order=None):
        return T.gauss_weights(self.numDim, self.quadratureOrder if order is None else order)

    def getIntegrationPointCoords(self, order=None):
        return T.gauss_points(self.numDim, self.quadratureOrder if order is None else order)

    def getReferenceElementCoordinates(self):
        n = self.order + 1
        x = np.tile(np.linspace(-1, 1, n), n)
        y = np.repeat(np.linspace(-1, 1, n), n)
        return np.vstack((x[self.shapeFuncToNodeOrder], y[self.shapeFuncToNodeOrder])).T


class TriElement2D(Element):
    """Lagrange triangle of order 1-3 (TriElement2D.py:69-169)."""

    def __init__(self, order=1, numStates=None, quadratureOrder=None):
        if order not in [1, 2, 3]:
            raise ValueError("Triangular elements only support orders 1, 2 and 3")
        self.order = order
        if quadratureOrder is None:
            quadratureOrder = int(np.ceil((order + 1) / 2))
        super().__init__((order + 1) * (order + 2) // 2, numDim=2, quadratureOrder=quadratureOrder, numStates=numStates)
        self.name = f"Order{self.order}-LagrangeTri"

    def computeShapeFunctions(self, paramCoords):
        return T.lagrange_tri(paramCoords[:, 0], paramCoords[:, 1], self.order)

    def computeShapeFunctionGradients(self, paramCoords):
        return T.lagrange_tri_deriv(paramCoords[:, 0], paramCoords[:, 1], self.order)

    def getIntegrationPointWeights(self, order=None):
        return T.tri_weights(self.quadratureOrder if order is None else order)

    def getIntegrationPointCoords(self, order=None):
        return T.tri_points(self.quadratureOrder if order is None else order)

    def getReferenceElementCoordinates(self):
        if self.order == 1:
            return np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
        if self.order == 2:
            return np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [0.5, 0.0], [0.5, 0.5], [0.0, 0.5]])
        t = 1 / 3
        return np.array([[0, 0], [1, 0], [0, 1], [t, 0], [2 * t, 0], [2 * t, t], [t, 2 * t], [0, 2 * t], [0, t], [t, t]], dtype=float)


class HexElement3D(Element):
    """Lagrange hexahedron of order 1-2 in meshio node order (HexElement3D.py:30-188)."""

    def __init__(self, order=1, numStates=None, quadratureOrder=None):
        if order not in [1, 2]:
            raise ValueError("Hex elements only support orders 1 and 2")
        self.order = order
        if quadratureOrder is None:
            quadratureOrder = int(np.ceil((2 * order + 1) / 2))
        super().__init__((order + 1) ** 3, numDim=3, quadratureOrder=quadratureOrder, numStates=numStates)
        self.name = f"Order{self.order}-LagrangeHex"
        self.shapeFuncToNodeOrder = T._hex_node_order(order)

    def computeShapeFunctions(self, paramCoords):
        N = T.lagrange3d(paramCoords[:, 0], paramCoords[:, 1], paramCoords[:, 2], self.order + 1)
        return np.ascontiguousarray(N[:, self.shapeFuncToNodeOrder])

    def computeShapeFunctionGradients(self, paramCoords):
        dN = T.lagrange3d_deriv(paramCoords[:, 0], paramCoords[:, 1], paramCoords[:, 2], self.order + 1)
        return np.ascontiguousarray(dN[:, :, self.shapeFuncToNodeOrder])

    def getIntegrationPointWeights(self, order=None):
        return T.gauss_weights(self.numDim, self.quadratureOrder if order is None else order)

    def getIntegrationPointCoords(self, order=None):
        return T.gauss_points(self.numDim, self.quadratureOrder if order is None else order)

    def getReferenceElementCoordinates(self):
        n = self.order + 1
        p = np.linspace(-1, 1, n)
        x, y, z = np.tile(p, n * n), np.tile(np.repeat(p, n), n), np.repeat(p, n * n)
        o = self.shapeFuncToNodeOrder
        return np.vstack((x[o], y[o], z[o])).T


class LineElement1D(Element):
    """Lagrange line element of order 1-3 (LineElement1D.py:30-180)."""

    def __init__(self, order=1, numStates=None, quadratureOrder=None):
        if order not in [1, 2, 3]:
            raise ValueError("Line elements only support orders 1, 2 and 3")
        self.order = order
        if quadratureOrder is None:
            quadratureOrder = int(np.ceil((order + 1) / 2))
        super().__init__(order + 1, numDim=1, quadratureOrder=quadratureOrder, numStates=numStates)
        self.name = f"Order{self.order}-LagrangeLine"
        # LineElement1D._getNodeReordering (LineElement1D.py:158-180): [0, 1], [0, 2, 1], [0, 2, 3, 1] -- for order 3 this is
        # not meshio's "ends first" order; connectivity must follow the element's own order (meshes.line_grid does)
        self.shapeFuncToNodeOrder = np.array([0] + list(range(2, order + 1)) + [1]) if order > 1 else np.array([0, 1])

    def computeShapeFunctions(self, paramCoords):
        N = T.lagrange1d(np.asarray(paramCoords).reshape(-1), self.order + 1)
        return np.ascontiguousarray(N[:, self.shapeFuncToNodeOrder])

    def computeShapeFunctionGradients(self, paramCoords):
        dN = T.lagrange1d_deriv(np.asarray(paramCoords).reshape(-1), self.order + 1)
        return np.ascontiguousarray(dN[:, None, self.shapeFuncToNodeOrder])

    def getIntegrationPointWeights(self, order=None):
        return T.gauss_weights(self.numDim, self.quadratureOrder if order is None else order)

    def getIntegrationPointCoords(self, order=None):
        return np.asarray(T.gauss_points(self.numDim, self.quadratureOrder if order is None else order)).reshape(-1, 1)

    def getReferenceElementCoordinates(self):
        return np.linspace(-1, 1, self.numNodes).reshape(-1, 1)[self.shapeFuncToNodeOrder]


def elementFromMeshioName(meshioName, numStates):
    """meshio cell-type name -> element object, or None if unsupported (Model.py:429-488)."""
    name = meshioName.lower()
    if name[:4] == "quad":
        if name == "quad":
            return QuadElement2D(order=1, numStates=numStates)
        numNodes = int(name[4:])
        if numNodes != 8:
            order = np.sqrt(numNodes) - 1
            if int(order) == order and int(order) in (1, 2, 3):
                return QuadElement2D(order=int(order), numStates=numStates)
        return None
    if name == "triangle":
        return TriElement2D(order=1, numStates=numStates)
    if name == "triangle6":
        return TriElement2D(order=2, numStates=numStates)
    if name == "triangle10":
        return TriElement2D(order=3, numStates=numStates)
    if name == "hexahedron":  # 27-node hexahedra: element level only, as in the reference (Model.py:470-476 fails for them)
        return HexElement3D(order=1, numStates=numStates)
    if name[:4] == "line":
        order = 1 if name == "line" else int(name[4:]) - 1
        return LineElement1D(order=order, numStates=numStates) if order in (1, 2, 3) else None
    return None
