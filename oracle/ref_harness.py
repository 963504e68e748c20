"""Import and drive the UNMODIFIED reference (FEMpy at /root/reference) inside the build container.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  /root/reference does not exist on the GPU box; there the
pip-installed copy under ``oracle/_ref`` (``oracle/make_ref.sh``, git-ignored) is used.  Users:
``oracle/gen_golden.py`` (fixture generation), ``oracle/ref_timing.py`` (the timed CPU baseline of bench.py)
and the tests that pin the oracle, which skip when no reference is found.

Three stub packages under ``oracle/stubs`` (baseclasses, meshio, parameterized) replace the
reference's un-installed dependencies; the reference sources themselves are
imported from where they lie and are never copied.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_STUBS = os.path.join(_HERE, "stubs")


def _find_reference():
    """The reference checkout in the build container, else the pip-installed copy oracle/make_ref.sh left in
    oracle/_ref (git-ignored, travels to the GPU box)."""
    for cand in (os.environ.get("FEMPY_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "FEMpy", "Problem.py")):
            return cand
    return os.environ.get("FEMPY_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_reference()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "FEMpy"))


def import_reference():
    """Returns the reference ``FEMpy`` package (numba kernels compile on first import, ~15 s)."""
    if not available():
        raise ImportError(f"reference not found at {REFERENCE_ROOT}")
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/fempy_ref_numba_cache")
    for p in (REFERENCE_ROOT, _STUBS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import FEMpy  # noqa: E402
    return FEMpy


def make_constitutive(fp, model, E, nu, rho, t, linear=True):
    """model 0 plane stress, 1 plane strain (t = thickness), 2 Iso3D, 3 Iso1D (t = cross-sectional area)"""
    if model == 2:
        return fp.Constitutive.Iso3D(E, nu, rho, linear=linear)
    if model == 3:
        return fp.Constitutive.Iso1D(E, rho, t, linear=linear)
    cls = {0: fp.Constitutive.IsoPlaneStress, 1: fp.Constitutive.IsoPlaneStrain}[model]
    return cls(E, nu, rho, t, linear=linear)


def build_problem(fp, coords, connectivity, model, E, nu, rho, t, linear=True, thickness=None,
                  bcs=None, loads=None, states=None):
    """Reference FEMpyModel + FEMpyProblem from arrays.

    bcs:   list of (name, nodeInds, dof, value)  -> problem.addFixedBCToNodes
    loads: list of (name, nodeInds, dof, value, totalLoad) -> problem.addLoadToNodes
    """
    import numpy as np

    con = make_constitutive(fp, model, E, nu, rho, t, linear)
    femModel = fp.FEMpyModel(con, nodeCoords=np.array(coords), connectivity={k: np.array(v) for k, v in connectivity.items()})
    if thickness is not None:
        femModel.setDesignVariables({("Area" if model == 3 else "Thickness"): np.array(thickness)})
    prob = femModel.addProblem("oracle")
    for name, nodes, dof, value in bcs or []:
        prob.addFixedBCToNodes(name, [int(n) for n in nodes], dof, value)
    for name, nodes, dof, value, total in loads or []:
        prob.addLoadToNodes(name, [int(n) for n in nodes], dof, value, totalLoad=total)
    if states is not None:
        prob.updateState(np.array(states))
    return femModel, prob


def assemble(prob, applyBCs=True):
    """(K csr, R) through the reference's own private assembly drivers (Problem.py:565-661)."""
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        K = prob._assembleMatrix(prob.states, applyBCs=applyBCs)
        R = prob._assembleResidual(prob.states, applyBCs=applyBCs)
    K.sort_indices()
    return K, R
