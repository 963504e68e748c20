"""Shared helpers for the tests: golden fixtures, oracle blocks, model construction."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from fempy_b200 import Constitutive, Elements, meshes  # noqa: E402
from hostmirror import FEMpyModel  # noqa: E402
from oracle import fem_oracle as fo  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
MAT = dict(E=70e9, nu=0.3, rho=2700.0, t=1.0)

ELEMENT_FOR = {
    "quad": lambda: Elements.QuadElement2D(1), "quad9": lambda: Elements.QuadElement2D(2),
    "quad16": lambda: Elements.QuadElement2D(3), "triangle": lambda: Elements.TriElement2D(1),
    "triangle6": lambda: Elements.TriElement2D(2), "triangle10": lambda: Elements.TriElement2D(3),
}  # fmt: skip
FAMILY_ELEMENT = {"quad1": "quad", "quad2": "quad9", "quad3": "quad16", "tri1": "triangle", "tri2": "triangle6", "tri3": "triangle10"}


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def oracle_blocks(connectivity):
    """fem_oracle.Block list for a connectivity dict, tables from this package's generators
    (themselves checked against the reference's tables in test_tables.py)."""
    blocks = []
    for elType, conn in connectivity.items():
        N, dN, w = ELEMENT_FOR[elType]().tables()
        blocks.append(fo.Block(N, dN, w, conn))
    return blocks


def make_cm(model, linear=True, E=MAT["E"], nu=MAT["nu"], rho=MAT["rho"], t=MAT["t"]):
    cls = Constitutive.IsoPlaneStress if model == 0 else Constitutive.IsoPlaneStrain
    return cls(E, nu, rho, t, linear=linear)


def mesh_case(g, name):
    """Rebuild the inputs of one meshes.npz case: dict with coords, conn, states, thick, bcDOF, ..."""
    conn = {k[len(name) + 6 :]: g[k] for k in g.files if k.startswith(name + "_conn_")}
    # keep the reference's dict order: quads before triangles (as generated)
    conn = {k: conn[k] for k in sorted(conn, key=lambda s: (not s.startswith("quad"), s))}
    model, linear, applyBCs = (int(v) for v in g[name + "_meta"])
    return dict(
        coords=g[name + "_coords"], conn=conn, states=g[name + "_states"], thick=g[name + "_thick"],
        bcDOF=g[name + "_bcDOF"], bcVal=g[name + "_bcVal"], loads=g[name + "_loads"], model=model,
        linear=bool(linear), applyBCs=bool(applyBCs),
        indptr=g[name + "_indptr"], indices=g[name + "_indices"], data=g[name + "_data"], R=g[name + "_R"],
    )  # fmt: skip


def mesh_case_names(g):
    return sorted(k[: -len("_meta")] for k in g.files if k.endswith("_meta"))


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = max(np.abs(b).max(), 1e-300)
    return np.abs(a - b).max() / scale


def build_problem(coords, conn, model, linear=True, thickness=None, states=None, bcs=(), loads=(), **mat):
    cm = make_cm(model, linear, **mat)
    m = FEMpyModel(cm, nodeCoords=coords, connectivity=conn)
    if thickness is not None:
        m.setDesignVariables({"Thickness": np.asarray(thickness)})
    p = m.addProblem("test")
    for name, nodes, dof, value in bcs:
        p.addFixedBCToNodes(name, [int(n) for n in nodes], dof, value)
    for name, nodes, dof, value, total in loads:
        p.addLoadToNodes(name, [int(n) for n in nodes], dof, value, totalLoad=total)
    if states is not None:
        p.updateState(np.asarray(states))
    return m, p
