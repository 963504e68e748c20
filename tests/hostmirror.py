"""TEST SUPPORT (not product code): a small stand-in for the reference's FEMpyModel / FEMpyProblem host classes.

The product's drop-in is `fempy_b200.B200AssemblyMixin` mixed into the REFERENCE's own `FEMpyProblem`
(`fempy_b200.b200ProblemClass`, exercised by tests/test_reference_plugin.py with the reference from oracle/_ref).
The parity tests, however, must also run where the reference and its dependencies (mdolab-baseclasses, meshio) are
not installed, so they drive the mixin through this stand-in instead.  It offers the attribute and method names the
mixin and the tests read -- model arrays, element registry, DOF numbering, design variables,
BC / load bookkeeping, the solve driver with its up-to-date flags -- and nothing of mesh I/O or output.
"""
import numpy as np
from scipy.sparse.linalg import factorized

from fempy_b200 import B200AssemblyMixin, Elements


class _Options:
    def __init__(self, defaults, given, owner):
        self._opts = dict(defaults)
        for key, val in (given or {}).items():
            if key not in self._opts:
                raise KeyError(f"{key!r} is not an option of {owner}")
            self._opts[key] = val

    def getOption(self, name):
        return self._opts[name]

    def setOption(self, name, value):
        self._opts[name] = value


def _expand_node_dofs(nodeInds, dof, value, numStates, divide=1.0):
    """node list x dof list -> flat (DOF, value) lists; a scalar value is repeated per dof."""
    nodes = np.atleast_1d(np.asarray(nodeInds, dtype=np.int64))
    dofs = np.atleast_1d(np.asarray(dof, dtype=np.int64))
    vals = np.full(dofs.size, float(value)) if np.isscalar(value) else np.asarray(value, dtype=float)
    if vals.size != dofs.size:
        raise Exception("give one value, or one value per entry of the DOF list")
    flatDof = (nodes[:, None] * numStates + dofs[None, :]).ravel()
    flatVal = np.tile(vals / divide, nodes.size)
    return [int(d) for d in flatDof], [float(v) for v in flatVal]


class FEMpyModel(_Options):
    """Mesh arrays -> what an assembly needs: 2-D node coordinates, per-type element registry, DVs, model BCs."""

    def __init__(self, constitutiveModel, meshFileName=None, nodeCoords=None, connectivity=None, options=None):
        _Options.__init__(self, {"outputDir": "./", "outputFormat": "vtk", "outputFunctions": []}, options, "FEMpyModel")
        if meshFileName is not None:
            raise NotImplementedError("mesh files are read by meshio in the reference; this stand-in takes arrays")
        if nodeCoords is None or connectivity is None:
            raise Exception("nodeCoords and connectivity are both needed")
        self.constitutiveModel = constitutiveModel
        xyz = np.atleast_2d(np.asarray(nodeCoords, dtype=np.float64))
        extent = np.ptp(xyz, axis=0) > 0.0  # a dimension without extent is dropped
        self.activeDimensions = [int(i) for i in np.flatnonzero(extent)]
        self.inactiveDimensions = [i for i in range(3) if i not in self.activeDimensions]
        if len(self.activeDimensions) != self.numDim:
            raise ValueError(f"{self.numDim}D constitutive model on a mesh with extent in {len(self.activeDimensions)} dimensions")
        self.nodeCoords = np.ascontiguousarray(xyz[:, self.activeDimensions])
        self.numNodes = xyz.shape[0]
        self.numDOF = self.numNodes * self.numStates
        self.connectivity = {k: np.array(v, dtype=np.int64) for k, v in connectivity.items()}
        self.elements, self.numElements = {}, 0
        for elType, conn in self.connectivity.items():
            elObject = Elements.elementFromMeshioName(elType, self.numStates)
            if elObject is None:
                import warnings

                warnings.warn(f"element type {elType} is not supported and is skipped", stacklevel=2)
                continue
            count = conn.shape[0]
            self.elements[elType] = dict(connectivity=conn, DOF=self.getDOFfromNodeInds(conn), elementObject=elObject,
                                         numElements=count, elementIDs=self.numElements + np.arange(count))
            self.numElements += count
        self.dvs = {name: np.full(self.numElements, spec["defaultValue"]) for name, spec in constitutiveModel.designVariables.items()}
        self.problems, self.BCs = [], {}

    numDim = property(lambda self: self.constitutiveModel.numDim)
    numStates = property(lambda self: self.constitutiveModel.numStates)

    def _stale(self):
        for p in self.problems:
            p.markResOutOfDate()
            p.markJacOutOfDate()

    def getCoordinates(self, force3D=False):
        if not force3D or self.numDim == 3:
            return self.nodeCoords.copy()
        out = np.zeros((self.numNodes, 3))
        out[:, self.activeDimensions] = self.nodeCoords
        return out

    def setCoordinates(self, nodeCoords):
        nodeCoords = np.asarray(nodeCoords, dtype=np.float64)
        if nodeCoords.shape[1] == self.numDim:
            self.nodeCoords = nodeCoords.copy()
        elif nodeCoords.shape[1] == 3:
            self.nodeCoords = np.ascontiguousarray(nodeCoords[:, self.activeDimensions])
        else:
            raise ValueError(f"{nodeCoords.shape[1]}D coordinates for a {self.numDim}D problem")
        self._stale()

    def setDesignVariables(self, dvs):
        for name, vals in dvs.items():
            if np.shape(vals) != self.dvs[name].shape:
                raise ValueError(f"design variable {name!r}: expected shape {self.dvs[name].shape}, got {np.shape(vals)}")
            self.dvs[name] = np.array(vals, dtype=np.float64)
        self._stale()

    def getDesignVariables(self):
        return {k: v.copy() for k, v in self.dvs.items()}

    def getElementCoordinates(self, elementType):
        return self.nodeCoords[self.elements[elementType]["connectivity"]]

    def getElementDVs(self, elementType):
        ids = self.elements[elementType]["elementIDs"]
        return {name: vals[ids] for name, vals in self.dvs.items()}

    def getDOFfromNodeInds(self, nodeIndices):
        """(..., n) node ids -> (..., n * numStates) int64 DOF ids, the states of a node adjacent."""
        nodes = np.asarray(nodeIndices, dtype=np.int64)
        dof = nodes[..., None] * self.numStates + np.arange(self.numStates, dtype=np.int64)
        return dof.reshape(nodes.shape[:-1] + (nodes.shape[-1] * self.numStates,))

    def addFixedBCToNodes(self, name, nodeInds, dof, value):
        d, v = _expand_node_dofs(nodeInds, dof, value, self.numStates)
        self.BCs[name] = {"DOF": d, "Value": v}

    def addProblem(self, name, options=None):
        self.problems.append(FEMpyProblem(name, self, options))
        return self.problems[-1]


class _HostProblem(_Options):
    """Load case: states, BC / load dictionaries, residual / Jacobian caching flags, the Newton-step driver."""

    def __init__(self, name, model, options=None):
        _Options.__init__(self, {"printTiming": False, "outputDir": model.getOption("outputDir"),
                                 "outputFormat": model.getOption("outputFormat"), "outputFunctions": model.getOption("outputFunctions")},
                          {k: v for k, v in (options or {}).items() if v is not None}, "FEMpyProblem")
        self.name, self.model = name, model
        self.states = np.zeros((model.numNodes, model.numStates))
        self.Res = np.zeros(model.numDOF)
        self.Jacobian = self.factorizedJac = None
        self.resUpToDate = self.jacUpToDate = self.jacIsFactorized = False
        self.BCs, self.loads, self.solveCounter = {}, {}, 0

    constitutiveModel = property(lambda self: self.model.constitutiveModel)
    isLinear = property(lambda self: self.model.constitutiveModel.isLinear)
    numDim = property(lambda self: self.model.numDim)
    numNodes = property(lambda self: self.model.numNodes)
    numDOF = property(lambda self: self.model.numDOF)
    numStates = property(lambda self: self.model.numStates)
    elements = property(lambda self: self.model.elements)

    # flags
    def markResOutOfDate(self):
        self.resUpToDate = False

    def markJacOutOfDate(self):
        self.jacUpToDate = False

    def markResUpToDate(self):
        self.resUpToDate = True

    def markJacUpToDate(self):
        self.jacUpToDate = True

    def _statesChanged(self):
        self.resUpToDate = False
        self.jacUpToDate = self.jacUpToDate and self.isLinear  # a linear model keeps its Jacobian

    def updateState(self, newStates):
        self.states[...] = np.reshape(newStates, self.states.shape)
        self._statesChanged()

    def incrementState(self, update):
        self.states += np.reshape(update, self.states.shape)
        self._statesChanged()

    def reset(self):
        self.updateState(0.0 * self.states)

    # assembly drivers (the private _assemble* methods come from B200AssemblyMixin)
    def updateResidual(self, applyBCs=True):
        if not self.resUpToDate:
            self.Res = self._assembleResidual(self.states, applyBCs=applyBCs)
            self.resUpToDate = True

    def updateJacobian(self, applyBCs=True):
        if not self.jacUpToDate:
            self.Jacobian = self._assembleMatrix(self.states, applyBCs=applyBCs)
            self.jacUpToDate, self.jacIsFactorized = True, False

    def solveJacLinear(self, RHS):
        self.updateJacobian()
        if not self.jacIsFactorized:
            self.factorizedJac = factorized(self.Jacobian.tocsc())
            self.jacIsFactorized = True
        return self.factorizedJac(RHS)

    def solve(self):
        import time

        stamps = {"Start": time.time()}
        self.updateResidual(applyBCs=True)
        stamps["ResAssembled"] = time.time()
        self.updateJacobian(applyBCs=True)
        stamps["JacAssembled"] = time.time()
        step = self.solveJacLinear(-self.Res)
        stamps["Solved"] = time.time()
        self.incrementState(step)
        self.solveCounter += 1
        return stamps

    # BCs and loads
    def addFixedBCToNodes(self, name, nodeInds, dof, value):
        d, v = _expand_node_dofs(nodeInds, dof, value, self.numStates)
        self.BCs[name] = {"DOF": d, "Value": v}

    def addLoadToNodes(self, name, nodeInds, dof, value, totalLoad=False):
        share = float(len(np.atleast_1d(nodeInds))) if totalLoad else 1.0
        d, v = _expand_node_dofs(nodeInds, dof, value, self.numStates, divide=share)
        self.loads[name] = {"DOF": d, "Value": v}

    def addBodyLoad(self, name, loadingFunction):
        raise NotImplementedError("body loads do not exist in the reference either")

    def getBCs(self):
        merged = {k: {"DOF": list(v["DOF"]), "Value": list(v["Value"])} for k, v in self.model.BCs.items()}
        merged.update(self.BCs)
        return merged

    def getElementStates(self, elementType):
        return self.states[self.elements[elementType]["connectivity"]]


class FEMpyProblem(B200AssemblyMixin, _HostProblem):
    """The stand-in load case with the assembly hot path on the GPU."""
