"""Synthetic structured meshes for the benchmark configurations (BASELINE.json configs 2-5).

The Q4 generator reproduces the node numbering, connectivity and warp of the reference's scaling
benchmark (reference Examples/Benchmarks/PlotQuadMeshScaling.py:34-56): nodes on a (nx+1) x (ny+1)
grid numbered x-fastest, coordinates warped by (x, y) -> (2x, (2 - x) y), counter-clockwise
connectivity [LL, LR, UR, UL].  The warp removes every exactly-zero entry from the element matrices,
so the reference's value-dependent sparsity pattern equals the structural one.
The Q9 and T3 variants use the same node grid construction.
"""
import numpy as np


def warp(x, y):
    return 2.0 * x, (2.0 - x) * y


def _grid_nodes(nxNodes, nyNodes, warped):
    x = np.tile(np.linspace(0.0, 1.0, nxNodes), nyNodes)
    y = np.repeat(np.linspace(0.0, 1.0, nyNodes), nxNodes)
    if warped:
        x, y = warp(x, y)
    return np.array([x, y]).T.copy()


def quad4_grid(nx, ny, warped=True):
    """nodeCoords (N,2) f64 and {"quad": (nx*ny, 4) int64}."""
    coords = _grid_nodes(nx + 1, ny + 1, warped)
    ll = np.tile(np.arange(nx), ny) + np.repeat((nx + 1) * np.arange(ny), nx)
    conn = np.stack((ll, ll + 1, ll + nx + 2, ll + nx + 1), axis=1).astype(np.int64)
    return coords, {"quad": conn}


def quad9_grid(nx, ny, warped=True):
    """9-node quads on a (2nx+1) x (2ny+1) node grid, meshio quad9 node order:
    4 corners CCW, 4 edge mid-points (bottom, right, top, left), centre."""
    W = 2 * nx + 1
    coords = _grid_nodes(W, 2 * ny + 1, warped)
    ll = np.tile(2 * np.arange(nx), ny) + np.repeat(2 * W * np.arange(ny), nx)
    conn = np.stack(
        (ll, ll + 2, ll + 2 * W + 2, ll + 2 * W, ll + 1, ll + W + 2, ll + 2 * W + 1, ll + W, ll + W + 1), axis=1
    ).astype(np.int64)
    return coords, {"quad9": conn}


def tri3_grid(nx, ny, warped=True):
    """Each grid cell split on its LL-UR diagonal: all lower triangles [LL, LR, UR] first, then all
    upper triangles [LL, UR, UL]."""
    coords = _grid_nodes(nx + 1, ny + 1, warped)
    ll = np.tile(np.arange(nx), ny) + np.repeat((nx + 1) * np.arange(ny), nx)
    lower = np.stack((ll, ll + 1, ll + nx + 2), axis=1)
    upper = np.stack((ll, ll + nx + 2, ll + nx + 1), axis=1)
    return coords, {"triangle": np.vstack((lower, upper)).astype(np.int64)}


def tri6_grid(nx, ny, warped=True):
    """6-node triangles on the quad9 node grid: every cell split on its LL-UR diagonal, meshio triangle6 node
    order (3 corners, then the mid-points of edges 0-1, 1-2, 2-0); lower triangles first, then upper."""
    W = 2 * nx + 1
    coords = _grid_nodes(W, 2 * ny + 1, warped)
    ll = np.tile(2 * np.arange(nx), ny) + np.repeat(2 * W * np.arange(ny), nx)
    lower = np.stack((ll, ll + 2, ll + 2 * W + 2, ll + 1, ll + W + 2, ll + W + 1), axis=1)
    upper = np.stack((ll, ll + 2 * W + 2, ll + 2 * W, ll + W + 1, ll + 2 * W + 1, ll + W), axis=1)
    return coords, {"triangle6": np.vstack((lower, upper)).astype(np.int64)}


def edge_nodes(coords, x=None, y=None):
    """Indices of nodes lying exactly on x == value or y == value (as the reference scripts select
    boundary nodes, e.g. PlotQuadMeshScaling.py:65-66)."""
    if x is not None:
        return np.argwhere(coords[:, 0] == x).flatten()
    return np.argwhere(coords[:, 1] == y).flatten()


GENERATORS = {"quad": quad4_grid, "quad9": quad9_grid, "triangle": tri3_grid, "triangle6": tri6_grid}


def hex8_grid(nx, ny, nz):
    """nodeCoords (N,3) on the unit cube and {"hexahedron": (nx*ny*nz, 8) int64}: nodes numbered x fastest, then y, then z;
    meshio hexahedron order (bottom face counter-clockwise seen from above, then the top face)."""
    x, y, z = np.linspace(0.0, 1.0, nx + 1), np.linspace(0.0, 1.0, ny + 1), np.linspace(0.0, 1.0, nz + 1)
    coords = np.stack((np.tile(x, (ny + 1) * (nz + 1)), np.tile(np.repeat(y, nx + 1), nz + 1), np.repeat(z, (nx + 1) * (ny + 1))), axis=1)
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    ll = (i + (nx + 1) * (j + (ny + 1) * k)).transpose(2, 1, 0).reshape(-1)
    dy, dz = nx + 1, (nx + 1) * (ny + 1)
    conn = np.stack((ll, ll + 1, ll + 1 + dy, ll + dy, ll + dz, ll + 1 + dz, ll + 1 + dy + dz, ll + dy + dz), axis=1).astype(np.int64)
    return coords, {"hexahedron": conn}


def line_grid(n, order=1):
    """nodeCoords (N,1) on [0, 1] and {"line" | "line3" | "line4": (n, order+1) int64} with the element nodes in the order
    the reference's LineElement1D places them along the element (LineElement1D.py:158-180): [0, 1], [0, 2, 1], [0, 2, 3, 1]
    as positions on the equispaced reference grid."""
    N = n * order + 1
    coords = np.linspace(0.0, 1.0, N).reshape(-1, 1)
    first = order * np.arange(n)
    position = [0] + list(range(2, order + 1)) + [1] if order > 1 else [0, 1]
    conn = np.stack([first + k for k in position], axis=1).astype(np.int64)
    return coords, {"line" if order == 1 else f"line{order + 1}": conn}


GENERATORS.update({"hexahedron": hex8_grid})
