"""Host-side basis-function and quadrature tables for the element families on the hot path.

These are evaluated once per element type at a handful of Gauss points and handed to the CUDA
library, which keeps them in ``__constant__``/shared memory.  They restate (not copy) the
reference generators:

* 1-D Lagrange polynomials on equispaced nodes in [-1, 1] and their tensor products
  (reference FEMpy/Basis/LagrangePoly.py:26-138).  The reference differentiates by complex step
  (``:52-73``); here the derivative is the exact product-rule sum, which agrees to rounding.
* Triangle Lagrange bases of order 1-3 (reference FEMpy/Basis/LagrangePoly.py:216-358).
* Tensor Gauss-Legendre rules with the point order "x slow, y fast"
  (reference FEMpy/Quadrature/GaussQuad.py:26-122) and the triangle rules of order 1-3
  (reference FEMpy/Quadrature/TriQuad.py:26-86).
* The shape-function -> node reordering of the quad family (meshio node order,
  reference FEMpy/Elements/QuadElement2D.py:160-186).
"""
from functools import lru_cache

import numpy as np


# ---------------------------------------------------------------------------------------------
# 1-D Lagrange basis on equispaced nodes
# ---------------------------------------------------------------------------------------------
def lagrange1d(x, n):
    """Values of the n equispaced-node Lagrange polynomials at points x -> (len(x), n)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    xi = np.linspace(-1.0, 1.0, n)
    N = np.ones((x.size, n))
    for m in range(n):
        for i in range(n):
            if i != m:
                N[:, m] *= (x - xi[i]) / (xi[m] - xi[i])
    return N


def lagrange1d_deriv(x, n):
    """d/dx of the n equispaced-node Lagrange polynomials at points x -> (len(x), n)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    xi = np.linspace(-1.0, 1.0, n)
    dN = np.zeros((x.size, n))
    for m in range(n):
        for k in range(n):
            if k == m:
                continue
            term = np.full(x.size, 1.0 / (xi[m] - xi[k]))
            for i in range(n):
                if i != m and i != k:
                    term *= (x - xi[i]) / (xi[m] - xi[i])
            dN[:, m] += term
    return dN


def lagrange2d(x, y, n):
    """Tensor-product basis, shape function j*n+k = Nx_k(x) * Ny_j(y) -> (P, n*n)."""
    Nx, Ny = lagrange1d(x, n), lagrange1d(y, n)
    return np.einsum("pj,pk->pjk", Ny, Nx).reshape(Nx.shape[0], n * n)


def lagrange2d_deriv(x, y, n):
    """Parametric gradient of the tensor-product basis -> (P, 2, n*n)."""
    Nx, Ny = lagrange1d(x, n), lagrange1d(y, n)
    dNx, dNy = lagrange1d_deriv(x, n), lagrange1d_deriv(y, n)
    out = np.zeros((Nx.shape[0], 2, n * n))
    out[:, 0, :] = np.einsum("pj,pk->pjk", Ny, dNx).reshape(-1, n * n)
    out[:, 1, :] = np.einsum("pj,pk->pjk", dNy, Nx).reshape(-1, n * n)
    return out


def lagrange3d(x, y, z, n):
    """Tensor-product basis in 3-D, index z*n*n + y*n + x -> (P, n**3)."""
    Nx, Ny, Nz = lagrange1d(x, n), lagrange1d(y, n), lagrange1d(z, n)
    return np.einsum("pk,pj,pi->pkji", Nz, Ny, Nx).reshape(Nx.shape[0], n**3)


def lagrange3d_deriv(x, y, z, n):
    """Parametric gradient of the 3-D tensor-product basis -> (P, 3, n**3)."""
    Nx, Ny, Nz = lagrange1d(x, n), lagrange1d(y, n), lagrange1d(z, n)
    dNx, dNy, dNz = lagrange1d_deriv(x, n), lagrange1d_deriv(y, n), lagrange1d_deriv(z, n)
    P = Nx.shape[0]
    out = np.zeros((P, 3, n**3))
    out[:, 0, :] = np.einsum("pk,pj,pi->pkji", Nz, Ny, dNx).reshape(P, -1)
    out[:, 1, :] = np.einsum("pk,pj,pi->pkji", Nz, dNy, Nx).reshape(P, -1)
    out[:, 2, :] = np.einsum("pk,pj,pi->pkji", dNz, Ny, Nx).reshape(P, -1)
    return out


# ---------------------------------------------------------------------------------------------
# Triangle Lagrange bases (area-coordinate construction, node order of the reference:
# 3 vertices, then edge nodes counter-clockwise starting on the edge 0-1, then the centroid)
# ---------------------------------------------------------------------------------------------
def _tri_basis_and_grad(x, y, order):
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    L = np.stack((1.0 - x - y, x, y), axis=1)  # barycentric coordinates L0, L1, L2
    dL = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])  # dL_i/d(x,y)
    P = x.size
    nn = (order + 1) * (order + 2) // 2
    N = np.zeros((P, nn))
    dN = np.zeros((P, 2, nn))

    def setfun(k, val, dval_dL):
        # dval_dL: (P,3) partial derivatives w.r.t. L0..L2 treated as independent
        N[:, k] = val
        dN[:, :, k] = dval_dL @ dL

    if order == 1:
        for i in range(3):
            g = np.zeros((P, 3))
            g[:, i] = 1.0
            setfun(i, L[:, i], g)
    elif order == 2:
        for i in range(3):
            g = np.zeros((P, 3))
            g[:, i] = 4.0 * L[:, i] - 1.0
            setfun(i, L[:, i] * (2.0 * L[:, i] - 1.0), g)
        for k, (i, j) in zip((3, 4, 5), ((0, 1), (1, 2), (2, 0))):
            g = np.zeros((P, 3))
            g[:, i] = 4.0 * L[:, j]
            g[:, j] = 4.0 * L[:, i]
            setfun(k, 4.0 * L[:, i] * L[:, j], g)
    elif order == 3:
        for i in range(3):
            Li = L[:, i]
            g = np.zeros((P, 3))
            g[:, i] = 0.5 * ((3 * Li - 1) * (3 * Li - 2) + 3 * Li * (3 * Li - 2) + 3 * Li * (3 * Li - 1))
            setfun(i, 0.5 * Li * (3 * Li - 1) * (3 * Li - 2), g)
        # edge nodes: on edge i->j the node nearer i is 4.5*Li*Lj*(3Li-1), nearer j is 4.5*Li*Lj*(3Lj-1)
        edgeNodes = ((3, 0, 1), (4, 1, 0), (5, 1, 2), (6, 2, 1), (7, 2, 0), (8, 0, 2))
        for k, i, j in edgeNodes:
            Li, Lj = L[:, i], L[:, j]
            g = np.zeros((P, 3))
            g[:, i] = 4.5 * Lj * (6 * Li - 1)
            g[:, j] = 4.5 * Li * (3 * Li - 1)
            setfun(k, 4.5 * Li * Lj * (3 * Li - 1), g)
        g = np.stack((27 * L[:, 1] * L[:, 2], 27 * L[:, 0] * L[:, 2], 27 * L[:, 0] * L[:, 1]), axis=1)
        setfun(9, 27 * L[:, 0] * L[:, 1] * L[:, 2], g)
    else:
        raise ValueError("Triangular elements only support orders 1, 2 and 3")
    return N, dN


def lagrange_tri(x, y, order):
    return _tri_basis_and_grad(x, y, order)[0]


def lagrange_tri_deriv(x, y, order):
    return _tri_basis_and_grad(x, y, order)[1]


# ---------------------------------------------------------------------------------------------
# Quadrature rules
# ---------------------------------------------------------------------------------------------
@lru_cache(maxsize=None)
def _gauss1d(order):
    return np.polynomial.legendre.leggauss(order)


def gauss_points(numDim, order):
    """Tensor Gauss-Legendre points, first coordinate slowest -> (order**numDim, numDim)."""
    p, _ = _gauss1d(order)
    if numDim == 1:
        return p.copy()
    if numDim == 2:
        return np.stack((np.repeat(p, order), np.tile(p, order)), axis=1)
    if numDim == 3:
        X = np.tile(np.repeat(p, order), order)
        Y = np.tile(np.tile(p, order), order)
        Z = np.repeat(p, order**2)
        return np.stack((X, Y, Z), axis=1)
    raise ValueError(f"Gauss quadrature only available in 1, 2 or 3 dimensions, not {numDim}")


def gauss_weights(numDim, order):
    _, w = _gauss1d(order)
    if numDim == 1:
        return w.copy()
    if numDim == 2:
        return np.repeat(w, order) * np.tile(w, order)
    if numDim == 3:
        return np.tile(np.repeat(w, order), order) * np.tile(np.tile(w, order), order) * np.repeat(w, order**2)
    raise ValueError(f"Gauss quadrature only available in 1, 2 or 3 dimensions, not {numDim}")


def tri_points(order):
    if order == 1:
        return np.array([[1 / 3, 1 / 3]])
    if order == 2:
        return np.array([[0.5, 0.5], [0.0, 0.5], [0.5, 0.0]])
    if order == 3:
        return np.array([[1 / 3, 1 / 3], [0.6, 0.2], [0.2, 0.6], [0.2, 0.2]])
    raise ValueError(f"Triangular quadrature of order {order} not supported")


def tri_weights(order):
    if order == 1:
        w = np.array([1.0])
    elif order == 2:
        w = np.ones(3) / 3
    elif order == 3:
        w = np.array([-27 / 28, 25 / 48, 25 / 48, 25 / 48])
    else:
        raise ValueError(f"Triangular quadrature of order {order} not supported")
    return 0.5 * w


# ---------------------------------------------------------------------------------------------
# Node orderings (shape function index -> meshio node index)
# ---------------------------------------------------------------------------------------------
QUAD_NODE_ORDER = {
    1: np.array([0, 1, 3, 2]),
    2: np.array([0, 2, 8, 6, 1, 5, 7, 3, 4]),
    3: np.array([0, 3, 15, 12, 1, 2, 7, 11, 14, 13, 8, 4, 5, 6, 10, 9]),
}


def _hex_node_order(order):
    """meshio/VTK node order for Lagrange hexahedra of order 1 and 2 expressed as indices into the
    z-slow/y/x-fast tensor grid (reference FEMpy/Elements/HexElement3D.py:164-188)."""
    n = order + 1

    def idx(i, j, k):
        return k * n * n + j * n + i

    m = n - 1
    corners = [idx(0, 0, 0), idx(m, 0, 0), idx(m, m, 0), idx(0, m, 0), idx(0, 0, m), idx(m, 0, m), idx(m, m, m), idx(0, m, m)]
    if order == 1:
        return np.array(corners)
    if order == 2:
        edges = [
            idx(1, 0, 0), idx(2, 1, 0), idx(1, 2, 0), idx(0, 1, 0),
            idx(1, 0, 2), idx(2, 1, 2), idx(1, 2, 2), idx(0, 1, 2),
            idx(0, 0, 1), idx(2, 0, 1), idx(2, 2, 1), idx(0, 2, 1),
        ]  # fmt: skip
        faces = [idx(0, 1, 1), idx(2, 1, 1), idx(1, 0, 1), idx(1, 2, 1), idx(1, 1, 0), idx(1, 1, 2)]
        return np.array(corners + edges + faces + [idx(1, 1, 1)])
    raise ValueError("Hex elements only support orders 1 and 2")
