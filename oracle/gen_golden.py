"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_harness.py) in the build container.

TEST INFRASTRUCTURE.  Run once here:  python oracle/gen_golden.py
The fixtures are committed; the GPU box (which has no /root/reference) only reads them.

Fixtures
  tables.npz      Gauss points/weights, N and dN/dxi of Quad order 1-3 and Tri order 1-3
  elements.npz    element-level K_e, R_e, von Mises point values for random elements of every family
                  x {plane stress, plane strain} x {linear, Green strain}  (the "finer seams" of SURVEY 8b)
  meshes.npz      assembled K (CSR), R, element functions on small quad / quad9 / triangle / mixed meshes
  ks.npz          element-level KS reductions ("ksmax" / "ksmin", Element.py:261-265) of von Mises on the elements.npz inputs
  points.npz      point quantities at arbitrary parametric points, non-default quadrature (intOrder), a caller-supplied function
  solid_bar.npz   HexElement3D / Iso3D and LineElement1D / Iso1D: element level + assembled hexahedron and line meshes
  lbracket.npz    config 1: Examples/Meshes/LBracket.msh mesh arrays + assembled K, R, solution, functions
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from fempy_b200 import meshes  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
MAT = dict(E=70e9, nu=0.3, rho=2700.0, t=1.0)


def families(fp):
    fam = {}
    for o in (1, 2, 3):
        fam[f"quad{o}"] = fp.Elements.QuadElement2D(order=o)
        fam[f"tri{o}"] = fp.Elements.TriElement2D(order=o)
    return fam


def gen_tables(fp):
    out = {}
    for name, el in families(fp).items():
        pts = el.getIntegrationPointCoords()
        out[f"{name}_pts"] = pts
        out[f"{name}_w"] = el.getIntegrationPointWeights()
        out[f"{name}_N"] = el.computeShapeFunctions(pts)
        out[f"{name}_dN"] = el.computeShapeFunctionGradients(pts)
        out[f"{name}_refcoords"] = el.getReferenceElementCoordinates()
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **out)


def gen_elements(fp):
    out = {}
    numEl = 6
    for name, el in families(fp).items():
        rng = np.random.default_rng(42)
        X = np.zeros((numEl, el.numNodes, 2))
        for i in range(numEl):
            X[i] = el.getRandomElementCoordinates(rng=rng)
        if name == "quad1":  # the reference's known-answer element (tests/testElements.py:80-82)
            X[0] = np.array([[0.0, 0.0], [2.0, 0.0], [2.0, 1.0], [0.0, 2.0]])
        q = 0.05 * rng.random(X.shape)
        thick = rng.random(numEl) + 0.5
        if name == "quad1":
            thick[0] = 1.0
        out[f"{name}_X"], out[f"{name}_q"], out[f"{name}_t"] = X, q, thick
        for model in (0, 1):
            for linear in (True, False):
                cm = rh.make_constitutive(fp, model, MAT["E"], MAT["nu"], MAT["rho"], MAT["t"], linear)
                dvs = {"Thickness": thick}
                key = f"{name}_m{model}_{'lin' if linear else 'nl'}"
                out[key + "_Ke"] = el.computeResidualJacobians(q, X, dvs, cm)
                out[key + "_Re"] = el.computeResiduals(q, X, dvs, cm)
                out[key + "_vm"] = el.computeFunction(X, q, dvs, cm.getFunction("Von-Mises-Stress"), None)
                out[key + "_vmmean"] = el.computeFunction(X, q, dvs, cm.getFunction("Von-Mises-Stress"), "mean")
                out[key + "_massint"] = el.computeFunction(X, q, dvs, cm.getFunction("Mass"), "integrate")
    np.savez_compressed(os.path.join(OUT, "elements.npz"), **out)


def gen_ks(fp):
    """KS element reductions (rho = 100, FEMpy/Utils/KSAgg.py:25-58) on the inputs of elements.npz.  Both names take
    the reference's first branch (Element.py:261-265), i.e. the KS *max*; a soft material (E = 2) keeps the
    exponentials away from 0/1 so the aggregate is not just max + log(1/P)/rho."""
    g = np.load(os.path.join(OUT, "elements.npz"))
    out = {}
    for name, el in families(fp).items():
        X, q, thick = g[f"{name}_X"], g[f"{name}_q"], g[f"{name}_t"]
        for model in (0, 1):
            for Emod in (MAT["E"], 2.0):
                cm = rh.make_constitutive(fp, model, Emod, MAT["nu"], MAT["rho"], MAT["t"], True)
                key = f"{name}_m{model}_E{'soft' if Emod == 2.0 else 'stiff'}"
                for red in ("ksmax", "ksmin"):
                    out[f"{key}_{red}"] = el.computeFunction(X, q, {"Thickness": thick}, cm.getFunction("Von-Mises-Stress"), red)
    np.savez_compressed(os.path.join(OUT, "ks.npz"), **out)


EXTRA_RULES = {"quad1": (1, 3, 4), "quad2": (2, 4), "quad3": (3,), "tri1": (2, 3), "tri2": (1, 3), "tri3": (1, 3)}  # intOrder values


def custom_function(states, stateGradients, coords, dvs):
    """an element function that is NOT one of the constitutive models' own: it reads every argument the reference hands
    to a point function (Element.py:203-214)"""
    return states[:, 0] * coords[:, 1] + stateGradients[:, 0, 1] * dvs["Thickness"] - 3.0 * stateGradients[:, 1, 0] * coords[:, 0]


def gen_points(fp):
    """points.npz: Element._computeFunctionEvaluationQuantities (Element.py:275-368) at random parametric points, the
    element-level methods with non-default quadrature (`intOrder=`, Element.py:163,370,426) and computeFunction with a
    caller-supplied point function, all on the inputs of elements.npz."""
    g = np.load(os.path.join(OUT, "elements.npz"))
    out = {}
    for name, el in families(fp).items():
        X, q, thick = g[f"{name}_X"], g[f"{name}_q"], g[f"{name}_t"]
        dvs = {"Thickness": thick}
        rng = np.random.default_rng(7)
        pts = rng.random((5, 2)) * (0.5 if name.startswith("tri") else 2.0) - (0.0 if name.startswith("tri") else 1.0)
        out[f"{name}_pts"] = pts
        pq = el._computeFunctionEvaluationQuantities(pts, q, X, dvs, ["Coord", "State", "StateGrad", "Jac", "JacDet", "StateGradSens", "DVs"])
        for key in ("Coord", "State", "StateGrad", "Jac", "JacDet", "StateGradSens"):
            out[f"{name}_{key}"] = pq[key]
        out[f"{name}_DVs"] = pq["DVs"]["Thickness"]
        for red in (None, "mean", "integrate", "max", "ksmax"):
            out[f"{name}_custom_{red}"] = el.computeFunction(X, q, dvs, custom_function, red)
        cm = rh.make_constitutive(fp, 0, MAT["E"], MAT["nu"], MAT["rho"], MAT["t"], True)
        cmnl = rh.make_constitutive(fp, 1, MAT["E"], MAT["nu"], MAT["rho"], MAT["t"], False)
        for order in EXTRA_RULES[name]:
            key = f"{name}_int{order}"
            out[key + "_Ke"] = el.computeResidualJacobians(q, X, dvs, cm, intOrder=order)
            out[key + "_Re"] = el.computeResiduals(q, X, dvs, cm, intOrder=order)
            out[key + "_Ke_nl"] = el.computeResidualJacobians(q, X, dvs, cmnl, intOrder=order)
            out[key + "_vm"] = el.computeFunction(X, q, dvs, cm.getFunction("Von-Mises-Stress"), None, intOrder=order)
            out[key + "_massint"] = el.computeFunction(X, q, dvs, cm.getFunction("Mass"), "integrate", intOrder=order)
            out[key + "_custom_integrate"] = el.computeFunction(X, q, dvs, custom_function, "integrate", intOrder=order)
    np.savez_compressed(os.path.join(OUT, "points.npz"), **out)


def gen_solid_bar(fp):
    """solid_bar.npz: the 3-D and 1-D families: HexElement3D order 1-2 with Iso3D, LineElement1D order 1-3
    with Iso1D -- element-level K_e / R_e / functions for linear and Green strain, and assembled systems on a hexahedron
    grid and on line meshes (the reference's model registry builds 27-node hexahedra only at element level:
    FEMpy/Model.py:470-476 reads an undefined name for them)."""
    out = {}
    numEl = 5
    fams = {"hex1": fp.Elements.HexElement3D(order=1), "hex2": fp.Elements.HexElement3D(order=2)}
    fams.update({f"line{o}": fp.Elements.LineElement1D(order=o) for o in (1, 2, 3)})
    for name, el in fams.items():
        rng = np.random.default_rng(43)
        X = np.zeros((numEl, el.numNodes, el.numDim))
        for i in range(numEl):
            X[i] = el.getRandomElementCoordinates(rng=rng)
        q = 0.05 * rng.random((numEl, el.numNodes, el.numStates))
        scale = rng.random(numEl) + 0.5
        out[f"{name}_X"], out[f"{name}_q"], out[f"{name}_t"] = X, q, scale
        model = 2 if name.startswith("hex") else 3
        dvs = {} if model == 2 else {"Area": scale}
        for linear in (True, False):
            cm = rh.make_constitutive(fp, model, MAT["E"], MAT["nu"], MAT["rho"], MAT["t"], linear)
            key = f"{name}_{'lin' if linear else 'nl'}"
            out[key + "_Ke"] = el.computeResidualJacobians(q, X, dvs, cm)
            out[key + "_Re"] = el.computeResiduals(q, X, dvs, cm)
            out[key + "_massint"] = el.computeFunction(X, q, dvs, cm.getFunction("Mass"), "integrate")
            for fname in (("Von-Mises-Stress",) if model == 2 else ("Stress", "Strain")):
                out[f"{key}_{fname}"] = el.computeFunction(X, q, dvs, cm.getFunction(fname), None)
                out[f"{key}_{fname}_mean"] = el.computeFunction(X, q, dvs, cm.getFunction(fname), "mean")
                out[f"{key}_{fname}_max"] = el.computeFunction(X, q, dvs, cm.getFunction(fname), "max")
    # assembled systems
    cases = {"mesh_hex": (meshes.hex8_grid(4, 3, 2), 2, True), "mesh_hex_nl": (meshes.hex8_grid(3, 2, 2), 2, False),
             "mesh_line": (meshes.line_grid(12, 1), 3, True), "mesh_line3": (meshes.line_grid(7, 2), 3, True), "mesh_line4_nl": (meshes.line_grid(5, 3), 3, False)}
    for ci, (name, ((coords, conn), model, linear)) in enumerate(cases.items()):
        rng = np.random.default_rng(300 + ci)
        if model == 2:
            coords = coords + 1e-2 * rng.random(coords.shape)
        else:
            coords = coords + np.concatenate(([0.0], 2e-2 * rng.random(coords.shape[0] - 2), [0.0])).reshape(-1, 1)
        elType = next(iter(conn))
        N, numEl = coords.shape[0], conn[elType].shape[0]
        states = 1e-3 * rng.random((N, coords.shape[1]))
        scale = rng.random(numEl) + 0.5
        left = np.flatnonzero(coords[:, 0] == coords[:, 0].min()) if model == 3 else np.flatnonzero(np.arange(N) % (conn[elType][0, 1] - conn[elType][0, 0] + (4 if name == "mesh_hex" else 3)) == 0)
        right = np.array([N - 1]) if model == 3 else np.flatnonzero(coords[:, 0] > 0.99)
        dofs = list(range(coords.shape[1]))
        _, prob = rh.build_problem(fp, coords, conn, model, MAT["E"], MAT["nu"], MAT["rho"], MAT["t"], linear, thickness=None if model == 2 else scale,
                                   bcs=[("Fixed", left, dofs, 0.0)], loads=[("Load", right, [0], 1e4, False)], states=states)
        K, R = rh.assemble(prob)
        bcDOF, bcVal = fp.Utils.AssemblyUtils.convertBCDictToLists(prob.getBCs())
        fname = "Von-Mises-Stress" if model == 2 else "Stress"
        with contextlib.redirect_stdout(io.StringIO()):
            f = prob.computeFunction(fname, "mean")[elType]
            mass = prob.computeFunction("Mass", "integrate", "sum")
        out.update({f"{name}_coords": coords, f"{name}_conn": conn[elType], f"{name}_eltype": np.array([elType]), f"{name}_states": states,
                    f"{name}_scale": scale, f"{name}_bcDOF": np.array(bcDOF), f"{name}_bcVal": np.array(bcVal, dtype=float),
                    f"{name}_loads": fp.Utils.AssemblyUtils.convertLoadsDictToVector(prob.loads, K.shape[0]),
                    f"{name}_meta": np.array([model, int(linear)]), f"{name}_indptr": K.indptr, f"{name}_indices": K.indices,
                    f"{name}_data": K.data, f"{name}_R": R, f"{name}_func_mean": f, f"{name}_mass": np.array(mass)})
    np.savez_compressed(os.path.join(OUT, "solid_bar.npz"), **out)


def mixed_grid(nx, ny):
    """Left half of the grid as quads, right half as triangles (two element types in one model)."""
    coords, cq = meshes.quad4_grid(nx, ny)
    conn = cq["quad"]
    col = np.tile(np.arange(nx), ny)
    quads = conn[col < nx // 2]
    rest = conn[col >= nx // 2]
    tris = np.vstack((rest[:, [0, 1, 2]], rest[:, [0, 2, 3]]))
    return coords, {"quad": quads, "triangle": tris}


def perturb(coords, seed):
    """Break the vertical/horizontal edge alignment so no element entry is an exact zero and the
    reference's value-dependent pattern equals the structural one (cf. tests/testProblem.py:52-54)."""
    rng = np.random.default_rng(seed)
    return coords + 1e-3 * rng.random(coords.shape)


MESH_CASES = {
    # name: (generator, args, model, linear, applyBCs)
    "quad_ps": (meshes.quad4_grid, (6, 4), 0, True, True),
    "quad_pe": (meshes.quad4_grid, (5, 5), 1, True, True),
    "quad_ps_nobc": (meshes.quad4_grid, (4, 3), 0, True, False),
    "quad_ps_nl": (meshes.quad4_grid, (4, 4), 0, False, True),
    "quad9_ps": (meshes.quad9_grid, (3, 3), 0, True, True),
    "quad9_pe": (meshes.quad9_grid, (4, 2), 1, True, True),
    "tri_ps": (meshes.tri3_grid, (5, 4), 0, True, True),
    "tri_pe": (meshes.tri3_grid, (4, 4), 1, True, True),
    "mixed_ps": (mixed_grid, (6, 3), 0, True, True),
    "quad_ps_32": (meshes.quad4_grid, (32, 32), 0, True, True),
    "quad9_ps_10": (meshes.quad9_grid, (10, 10), 0, True, True),
    "tri_ps_20": (meshes.tri3_grid, (20, 20), 0, True, True),
}


def gen_meshes(fp):
    out = {}
    for ci, (name, (gen, args, model, linear, applyBCs)) in enumerate(MESH_CASES.items()):
        rng = np.random.default_rng(100 + ci)
        coords, conn = gen(*args)
        coords = perturb(coords, ci)
        N = coords.shape[0]
        numEl = sum(c.shape[0] for c in conn.values())
        states = (1e-3 if linear else 5e-2) * rng.random((N, 2))
        thick = rng.random(numEl) * 1e-2 + 1e-3
        left = np.argwhere(coords[:, 0] < 1.5e-3).flatten()
        right = np.argwhere(coords[:, 0] > 2.0 - 1e-9).flatten()
        # BCs: left edge fixed to non-zero values; one DOF repeated in a second entry (diagonal 2.0,
        # later value wins in the residual).  Loads: right edge + a second entry hitting the same DOFs.
        bcs = [("Fixed", left, [0, 1], [0.0, 1e-4]), ("Again", [int(left[0])], [1], 2e-4)]
        loads = [("Load", right, [0], 1e4, False), ("Load2", right[:2], [0, 1], [5e3, -2e3], True)]
        m, p = rh.build_problem(fp, coords, conn, model, MAT["E"], MAT["nu"], MAT["rho"], MAT["t"], linear=linear,
                                thickness=thick, bcs=bcs, loads=loads, states=states)
        K, R = rh.assemble(p, applyBCs=applyBCs)
        out[name + "_coords"], out[name + "_states"], out[name + "_thick"] = coords, states, thick
        for k, c in conn.items():
            out[f"{name}_conn_{k}"] = c
        bcDOF, bcVal = [], []
        for key, bc in p.getBCs().items():
            bcDOF += bc["DOF"]
            bcVal += bc["Value"]
        out[name + "_bcDOF"], out[name + "_bcVal"] = np.array(bcDOF, dtype=np.int64), np.array(bcVal, dtype=float)
        loadVec = np.zeros(2 * N)
        for key in p.loads:
            loadVec[p.loads[key]["DOF"]] += p.loads[key]["Value"]
        out[name + "_loads"] = loadVec
        out[name + "_meta"] = np.array([model, int(linear), int(applyBCs)])
        out[name + "_indptr"], out[name + "_indices"], out[name + "_data"] = K.indptr, K.indices, K.data
        out[name + "_R"] = R
        for red in ("mean", "max", "min", "sum", "integrate"):
            vals = p.computeFunction("Von-Mises-Stress", red)
            for k in conn:
                out[f"{name}_vm_{red}_{k}"] = vals[k]
        vals = p.computeFunction("Mass", "integrate")
        for k in conn:
            out[f"{name}_mass_integrate_{k}"] = vals[k]
        out[name + "_mass_total"] = np.array(p.computeFunction("Mass", "integrate", "sum"))
        out[name + "_vm_mean_max"] = np.array(p.computeFunction("Von-Mises-Stress", "mean", "max"))
    np.savez_compressed(os.path.join(OUT, "meshes.npz"), **out)


def gen_lbracket(fp):
    """Config 1: the vertical-load case of Examples/LBracketExample.py:28-62."""
    import contextlib
    import io

    E, nu, rho, t = 71.7e9, 0.33, 2780, 5e-3
    cm = fp.Constitutive.IsoPlaneStress(E, nu, rho, t)
    model = fp.FEMpyModel(cm, meshFileName=os.path.join(rh.REFERENCE_ROOT, "Examples", "Meshes", "LBracket.msh"))
    coords = model.getCoordinates()
    top = np.argwhere(coords[:, 1] == np.max(coords[:, 1])).flatten()
    model.addFixedBCToNodes(name="Fixed", nodeInds=[int(i) for i in top], dof=[0, 1], value=0.0)
    prob = model.addProblem("Vertical-Load")
    right = np.argwhere(coords[:, 0] == np.max(coords[:, 0])).flatten()
    prob.addLoadToNodes(name="RightEdgeLoad", nodeInds=[int(i) for i in right], dof=1, value=-1e5, totalLoad=True)
    K, R = rh.assemble(prob)
    with contextlib.redirect_stdout(io.StringIO()):
        prob.solve()
    out = dict(
        coords=coords, conn=model.elements["quad"]["connectivity"].astype(np.int32), top=top, right=right,
        indptr=K.indptr, indices=K.indices.astype(np.int32), data=K.data, R=R, u=prob.states,
        vm_mean=prob.computeFunction("Von-Mises-Stress", "mean")["quad"],
        vm_mean_max=np.array(prob.computeFunction("Von-Mises-Stress", "mean", "max")),
        mass=np.array(prob.computeFunction("Mass", "integrate", "sum")),
        material=np.array([E, nu, rho, t]),
    )
    K2, R2 = rh.assemble(prob)  # residual at the solved state (should be ~0 except rounding)
    out["R_solved"] = R2
    np.savez_compressed(os.path.join(OUT, "lbracket.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    fp = rh.import_reference()
    if sys.argv[1:] == ["ks"]:  # later addition: leaves the other fixtures untouched
        gen_ks(fp)
        sys.exit(0)
    if sys.argv[1:] == ["points"]:  # round 2 addition
        gen_points(fp)
        sys.exit(0)
    if sys.argv[1:] == ["solid"]:  # round 2 addition
        gen_solid_bar(fp)
        sys.exit(0)
    gen_tables(fp)
    gen_elements(fp)
    gen_meshes(fp)
    gen_lbracket(fp)
    gen_ks(fp)
    gen_points(fp)
    gen_solid_bar(fp)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
