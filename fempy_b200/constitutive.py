"""Host-side mirror of the reference's 2-D isotropic constitutive models
(reference FEMpy/Constitutive/IsoPlaneStress.py:28-227, IsoPlaneStrain.py, ConstitutiveModel.py:44-92).

These objects only carry parameters: the strain / stress / weak-form arithmetic the reference runs
in Numba (ConstitutiveModel.py:240-460) is fused into the CUDA kernels (csrc/elem_math.cuh).
`getFunction` returns a small descriptor that the CUDA path maps to a device function id; unknown
names raise ValueError exactly as the reference does (IsoPlaneStress.py:206-209).
"""
import numpy as np

PLANE_STRESS, PLANE_STRAIN, ISO_3D, ISO_1D = 0, 1, 2, 3
FUNC_MASS, FUNC_VON_MISES, FUNC_STRESS, FUNC_STRAIN = 0, 1, 2, 3
FUNCTION_IDS = {"mass": FUNC_MASS, "von-mises-stress": FUNC_VON_MISES, "stress": FUNC_STRESS, "strain": FUNC_STRAIN}


class PointFunction:
    """Descriptor of a point function evaluated on the device (what `getFunction` returns)."""

    def __init__(self, name, funcId, constitutiveModel):
        self.name = name
        self.funcId = funcId
        self.constitutiveModel = constitutiveModel

    def __repr__(self):
        return f"PointFunction({self.name!r})"


class ConstitutiveModel:
    def __init__(self, numDim, stateNames, strainNames, stressNames, designVars, functionNames, linear=True):
        self.numDim = numDim
        self.stateNames = stateNames
        self.numStates = len(stateNames)
        self.strainNames = strainNames
        self.numStrains = len(strainNames)
        if len(stressNames) != self.numStrains:
            raise ValueError("Number of strains must equal number of stresses")
        self.stressNames = stressNames
        self.designVariables = designVars
        self.numDesignVariables = len(designVars)
        self.functionNames = functionNames
        self.lowerCaseFuncNames = [func.lower() for func in self.functionNames]
        self.isLinear = linear

    def getFunction(self, name):
        if name.lower() not in self.lowerCaseFuncNames:
            raise ValueError(
                f"{name} is not a valid function name for this constitutive model, valid choices are {self.functionNames}"
            )
        return PointFunction(name, FUNCTION_IDS[name.lower()], self)


class _IsoPlane(ConstitutiveModel):
    modelId = None

    def __init__(self, E, nu, rho, t, linear=True):
        designVars = {"Thickness": {"defaultValue": t}}
        stateNames = ["X-Displacement", "Y-Displacement"]
        strainNames = ["e_xx", "e_yy", "gamma_xy"]
        stressNames = ["sigma_xx", "sigma_yy", "tau_xy"]
        functionNames = ["Mass", "Von-Mises-Stress"]
        self.E = E
        self.nu = nu
        self.rho = rho
        super().__init__(2, stateNames, strainNames, stressNames, designVars, functionNames, linear)

    def material(self):
        """The `femb_material` struct for this model."""
        from ._lib import Material

        return Material(float(self.E), float(self.nu), float(self.rho), int(self.modelId), int(not self.isLinear))


class IsoPlaneStress(_IsoPlane):
    """2-D isotropic plane stress; design variable "Thickness" (reference IsoPlaneStress.py:28-72)."""

    modelId = PLANE_STRESS

    def stressStrainMatrix(self):
        E, nu = self.E, self.nu
        return E / (1.0 - nu**2) * np.array([[1.0, nu, 0.0], [nu, 1.0, 0.0], [0.0, 0.0, (1.0 - nu) / 2.0]])


class IsoPlaneStrain(_IsoPlane):
    """2-D isotropic plane strain; design variable "Thickness" (reference IsoPlaneStrain.py)."""

    modelId = PLANE_STRAIN

    def stressStrainMatrix(self):
        E, nu = self.E, self.nu
        return (
            E
            / ((1.0 + nu) * (1.0 - 2.0 * nu))
            * np.array([[1.0 - nu, nu, 0.0], [nu, 1.0 - nu, 0.0], [0.0, 0.0, (1.0 - 2.0 * nu) / 2.0]])
        )


class Iso3D(ConstitutiveModel):
    """3-D isotropic elasticity, no design variable (reference Iso3D.py:28-217)."""

    modelId = ISO_3D

    def __init__(self, E, nu, rho, linear=True):
        self.E, self.nu, self.rho = E, nu, rho
        super().__init__(3, ["X-Displacement", "Y-Displacement", "Z-Displacement"],
                         ["e_xx", "e_yy", "e_zz", "gamma_xy", "gamma_xz", "gamma_yz"],
                         ["sigma_xx", "sigma_yy", "sigma_zz", "tau_xy", "tau_xz", "tau_yz"], {}, ["Mass", "Von-Mises-Stress"], linear)

    def material(self):
        from ._lib import Material

        return Material(float(self.E), float(self.nu), float(self.rho), ISO_3D, int(not self.isLinear))

    def stressStrainMatrix(self):
        E, nu = self.E, self.nu
        D = np.zeros((6, 6))
        D[:3, :3] = nu
        D[[0, 1, 2], [0, 1, 2]] = 1.0 - nu
        D[[3, 4, 5], [3, 4, 5]] = (1.0 - 2.0 * nu) / 2.0
        return E / ((1.0 + nu) * (1.0 - 2.0 * nu)) * D


class Iso1D(ConstitutiveModel):
    """Axial bar; design variable "Area" (reference Iso1D.py:28-225)."""

    modelId = ISO_1D

    def __init__(self, E, rho, A, linear=True):
        self.E, self.rho, self.nu = E, rho, 0.0
        super().__init__(1, ["X-Displacement"], ["e_xx"], ["sigma_xx"], {"Area": {"defaultValue": A}}, ["Mass", "Stress", "Strain"], linear)

    def material(self):
        from ._lib import Material

        return Material(float(self.E), 0.0, float(self.rho), ISO_1D, int(not self.isLinear))

    def stressStrainMatrix(self):
        return np.array([[self.E]])
