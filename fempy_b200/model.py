"""Host-side mirror of the reference's FEMpyModel (reference FEMpy/Model.py:51-505): mesh arrays,
active-dimension detection, DOF maps, element registry, design variables and model-level BCs.

Mesh file I/O (meshio) is out of scope: models are built from arrays.
"""
import copy
import warnings

import numpy as np

from .elements import elementFromMeshioName
from .problem import FEMpyProblem


class FEMpyModel:
    def __init__(self, constitutiveModel, meshFileName=None, nodeCoords=None, connectivity=None, options=None):
        self.options = {"outputDir": "./", "outputFormat": "vtk", "outputFunctions": []}
        for key, value in (options or {}).items():
            if key not in self.options:
                raise KeyError(f"Option {key} is not a valid FEMpyModel option")
            self.options[key] = value
        self.constitutiveModel = constitutiveModel
        self._b200_geomVersion = 0  # counts setCoordinates / setDesignVariables calls: the device copy follows it
        if meshFileName is not None:
            raise NotImplementedError("mesh file input needs meshio, which is outside the B200 hot path; pass nodeCoords and connectivity")
        if nodeCoords is None or connectivity is None:
            raise Exception("Must supply either meshFileName or nodeCoords and connectivity")
        nodeCoords = np.asarray(nodeCoords)
        self.nodeCoords = np.zeros((nodeCoords.shape[0], 3))
        self.nodeCoords[:, : nodeCoords.shape[1]] = nodeCoords
        self.connectivity = copy.deepcopy(connectivity)

        # keep only the dimensions in which the mesh has extent (Model.py:100-117)
        self.numNodes = self.nodeCoords.shape[0]
        self.activeDimensions, self.inactiveDimensions = [], []
        for ii in range(3):
            if self.nodeCoords[:, ii].max() != self.nodeCoords[:, ii].min():
                self.activeDimensions.append(ii)
            else:
                self.inactiveDimensions.append(ii)
        if self.numDim != len(self.activeDimensions):
            raise ValueError(
                f"You have chosen a {self.numDim}D constitutive model model but the mesh is {len(self.activeDimensions)}D"
            )
        self.nodeCoords = np.ascontiguousarray(self.nodeCoords[:, self.activeDimensions])
        self.numDOF = self.numNodes * self.numStates

        # element registry (Model.py:123-142)
        self.elements = {}
        self.numElements = 0
        for elType in self.connectivity:
            elObject = elementFromMeshioName(elType, self.numStates)
            if elObject is None:
                warnings.warn(f"Element type {elType} is not supported by FEMpy and will be ignored", stacklevel=2)
                continue
            conn = np.ascontiguousarray(self.connectivity[elType], dtype=np.int64)
            n = conn.shape[0]
            self.elements[elType] = {
                "connectivity": conn,
                "DOF": self.getDOFfromNodeInds(conn),
                "elementObject": elObject,
                "numElements": n,
                "elementIDs": np.arange(self.numElements, self.numElements + n),
            }
            self.numElements += n

        self.dvs = {}
        for dvName, dv in self.constitutiveModel.designVariables.items():
            self.dvs[dvName] = dv["defaultValue"] * np.ones(self.numElements)
        self.problems = []
        self.BCs = {}

    numDim = property(lambda self: self.constitutiveModel.numDim)
    numStates = property(lambda self: self.constitutiveModel.numStates)
    problemNames = property(lambda self: [p.name for p in self.problems])

    def getOption(self, name):
        return self.options[name]

    def setOption(self, name, value):
        self.options[name] = value

    def getCoordinates(self, force3D=False):
        currentCoords = np.copy(self.nodeCoords)
        if not force3D or self.numDim == 3:
            return currentCoords
        coords = np.zeros((self.numNodes, 3))
        coords[:, self.activeDimensions] = currentCoords
        return coords

    def setCoordinates(self, nodeCoords):
        """Store new node coordinates (Model.py:240-256; the reference returns a copy instead of storing
        when the width equals numDim -- here the coordinates are stored in both cases)."""
        size = nodeCoords.shape[1]
        if size == self.numDim:
            self.nodeCoords = np.ascontiguousarray(nodeCoords, dtype=np.float64).copy()
        elif self.numDim != 3 and size == 3:
            self.nodeCoords = np.ascontiguousarray(nodeCoords[:, self.activeDimensions], dtype=np.float64)
        else:
            raise ValueError(
                f"Invalid number of coordinate dimensions, problem is {self.numDim}D but {size}D coordinates were supplied"
            )
        self._b200_geomVersion += 1  # the device copy of the coordinates is stale (problem.py: _b200Inputs)
        for prob in self.problems:
            prob.markResOutOfDate()
            prob.markJacOutOfDate()

    def setDesignVariables(self, dvs):
        for dvName in dvs:
            if dvs[dvName].shape == self.dvs[dvName].shape:
                self.dvs[dvName] = np.copy(dvs[dvName])
            else:
                raise ValueError(
                    f"Invalid shape for design variable '{dvName}', expected {self.dvs[dvName].shape} but got {dvs[dvName].shape}"
                )
        self._b200_geomVersion += 1
        for prob in self.problems:
            prob.markResOutOfDate()
            prob.markJacOutOfDate()

    def getDesignVariables(self):
        return copy.deepcopy(self.dvs)

    def getElementCoordinates(self, elementType):
        """(E, n, d) node coordinates per element (Model.py:288-307), as one fancy-index gather."""
        return self.nodeCoords[self.elements[elementType]["connectivity"]]

    def getElementDVs(self, elementType):
        elementInds = self.elements[elementType]["elementIDs"]
        return {dv: self.dvs[dv][elementInds] for dv in self.dvs}

    def addFixedBCToNodes(self, name, nodeInds, dof, value):
        """Model-level BC applied to every problem (Model.py:331-372)."""
        if isinstance(nodeInds, (int, np.integer)):
            nodeInds = [nodeInds]
        if isinstance(dof, (int, np.integer)):
            dof = [dof]
        if isinstance(value, (float, int)):
            value = [value] * len(dof)
        elif len(value) != len(dof):
            raise Exception("value should be a single entry or a list of values the same length as the DOF list.")
        dofNodes, valDOF = [], []
        for i in range(len(nodeInds)):
            for j in range(len(dof)):
                dofNodes.append(int(nodeInds[i]) * self.numStates + dof[j])
                valDOF.append(value[j])
        self.BCs[name] = {"DOF": dofNodes, "Value": valDOF}

    def addProblem(self, name, options=None):
        problem = FEMpyProblem(name, self, options)
        self.problems.append(problem)
        return problem

    def getDOFfromNodeInds(self, nodeIndices):
        """dof = node * numStates + s, last axis widened by numStates, int64 (Model.py:387-414)."""
        nodeIndices = np.asarray(nodeIndices, dtype=np.int64)
        shape = list(nodeIndices.shape)
        shape[-1] *= self.numStates
        dof = nodeIndices.reshape(-1, 1) * self.numStates + np.arange(self.numStates, dtype=np.int64)
        return dof.reshape(shape)
