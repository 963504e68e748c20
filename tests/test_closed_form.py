"""The algebra of the write-once fan kernel (fempy_b200/csrc/assemble_fan.cuh), restated line by line in numpy and checked
on the CPU against the oracle's Gauss-quadrature element matrices (reference Element.py:426-486).

The kernel never forms K_e: thread i evaluates, for each element of its fan, the node-row strip K_e[2a:2a+2, :] of its own
corner a in closed form (bilinear quad: four moments of theta / (8 detJ) over the 2x2 Gauss rule; linear triangle: constant
gradients).  The element is rotated so that the owner is local node 0, which is exact because the standard tables are
invariant under a cyclic renumbering.  These tests pin exactly those two claims without a GPU; the CUDA path itself is
compared with the oracle in test_gpu_parity.py.
"""
import numpy as np
import pytest

from util import ELEMENT_FOR, fo

E_MOD, NU = 70e9, 0.3


def d_entries(model):
    D = fo.d_matrix(model, E_MOD, NU)
    return D[0, 0], D[0, 1], D[1, 1], D[2, 2]


def apply_d(d, s00, s01, s10, s11):
    """2x2 block from the geometric sums S_dD = sum_p w theta detJ dN_a/dx_d dN_b/dx_D (assemble_fan.cuh: apply_d)."""
    d00, d01, d11, d22 = d
    return np.array([[d00 * s00 + d22 * s11, d01 * s01 + d22 * s10], [d01 * s10 + d22 * s01, d11 * s11 + d22 * s00]])


def quad_row_blocks(own, X1, X2, X3, th, d):
    """Row strip of the bilinear quad for the corner at `own` (local node 0 of the rotated element own, X1, X2, X3)."""
    G, THIRD = 0.57735026918962576451, 1.0 / 3.0
    ya, yb, yc, yd, ye, yf = X1[1] - X3[1], X2[1] - own[1], X3[1] - X2[1], own[1] - X1[1], X2[1] - X1[1], own[1] - X3[1]
    xa, xb, xc, xd, xe, xf = X3[0] - X1[0], own[0] - X2[0], X2[0] - X3[0], X1[0] - own[0], X1[0] - X2[0], X3[0] - own[0]
    dd0 = xb * ya - xa * yb
    gx, ge = G * (xc * yd - xd * yc), G * (xe * yf - xf * ye)
    tp, tm, th8 = dd0 + gx, dd0 - gx, 0.125 * th
    spp, spm, smp, smm = th8 / (tp + ge), th8 / (tp - ge), th8 / (tm + ge), th8 / (tm - ge)
    up, um, vp, vm = spp + spm, smp + smm, spp - spm, smp - smm
    m0, mx, me, mxe = up + um, G * (up - um), G * (vp + vm), THIRD * (vp - vm)
    m3 = THIRD * m0
    P0, Q0, R0 = m0 * ya + mx * yc + me * ye, mx * ya + m3 * yc + mxe * ye, me * ya + mxe * yc + m3 * ye
    P1, Q1, R1 = m0 * xa + mx * xc + me * xe, mx * xa + m3 * xc + mxe * xe, me * xa + mxe * xc + m3 * xe
    # coefficient rows of 8 detJ grad N_b = C_b + xi X_b + eta E_b  (y parts | x parts), b = 0..3
    Cy, Cx = (ya, yb, -ya, -yb), (xa, xb, -xa, -xb)
    Xy, Xx = (yc, -yc, yd, -yd), (xc, -xc, xd, -xd)
    Ey, Ex = (ye, yf, -yf, -ye), (xe, xf, -xf, -xe)
    blocks = []
    for b in range(4):
        s00 = P0 * Cy[b] + Q0 * Xy[b] + R0 * Ey[b]
        s01 = P0 * Cx[b] + Q0 * Xx[b] + R0 * Ex[b]
        s10 = P1 * Cy[b] + Q1 * Xy[b] + R1 * Ey[b]
        s11 = P1 * Cx[b] + Q1 * Xx[b] + R1 * Ex[b]
        blocks.append(apply_d(d, s00, s01, s10, s11))
    return blocks


def tri_row_blocks(own, X1, X2, th, d):
    g0 = (X1[1] - X2[1], X2[1] - own[1], own[1] - X1[1])
    g1 = (X2[0] - X1[0], own[0] - X2[0], X1[0] - own[0])
    det = g1[2] * g0[1] - g1[1] * g0[2]
    sc = 0.5 * th / det
    A0, A1 = sc * g0[0], sc * g1[0]
    return [apply_d(d, A0 * g0[b], A0 * g1[b], A1 * g0[b], A1 * g1[b]) for b in range(3)]


def oracle_ke(elType, X, th, model):
    _, dN, w = ELEMENT_FOR[elType]().tables()
    n = X.shape[0]
    return fo.element_jacobians(dN, w, X[None], np.zeros((1, n, 2)), np.array([th]), model, E_MOD, NU)[0]


def random_quad(rng):
    base = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]]) * rng.uniform(0.2, 3.0, 2)
    phi = rng.uniform(0, 2 * np.pi)
    Rm = np.array([[np.cos(phi), -np.sin(phi)], [np.sin(phi), np.cos(phi)]])
    return (base + rng.uniform(-0.12, 0.12, (4, 2)) * base.max()) @ Rm.T + rng.uniform(-5, 5, 2)


@pytest.mark.parametrize("model", [fo.PLANE_STRESS, fo.PLANE_STRAIN])
def test_quad_corner_strip_is_the_gauss_quadrature_row(model):
    rng = np.random.default_rng(7)
    d = d_entries(model)
    for _ in range(50):
        X, th = random_quad(rng), rng.uniform(0.1, 2.0)
        Ke = oracle_ke("quad", X, th, model)
        scale = np.abs(Ke).max()
        for a in range(4):  # the corner a owns rows 2a, 2a+1; the element rotated so that a is local node 0
            rot = [(a + k) % 4 for k in range(4)]
            blocks = quad_row_blocks(X[rot[0]], X[rot[1]], X[rot[2]], X[rot[3]], th, d)
            for k, b in enumerate(rot):
                np.testing.assert_allclose(blocks[k], Ke[2 * a : 2 * a + 2, 2 * b : 2 * b + 2], rtol=0, atol=2e-13 * scale)


@pytest.mark.parametrize("model", [fo.PLANE_STRESS, fo.PLANE_STRAIN])
def test_triangle_corner_strip_is_the_gauss_quadrature_row(model):
    rng = np.random.default_rng(11)
    d = d_entries(model)
    for _ in range(50):
        X = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]]) * rng.uniform(0.2, 3.0, 2) + rng.uniform(-0.2, 0.2, (3, 2)) + rng.uniform(-5, 5, 2)
        th = rng.uniform(0.1, 2.0)
        Ke = oracle_ke("triangle", X, th, model)
        scale = np.abs(Ke).max()
        for a in range(3):
            rot = [(a + k) % 3 for k in range(3)]
            blocks = tri_row_blocks(X[rot[0]], X[rot[1]], X[rot[2]], th, d)
            for k, b in enumerate(rot):
                np.testing.assert_allclose(blocks[k], Ke[2 * a : 2 * a + 2, 2 * b : 2 * b + 2], rtol=0, atol=2e-13 * scale)


def test_inverted_quad_keeps_the_signed_determinant():
    """Clockwise node order: detJ < 0 at every point; the reference keeps the sign (LinAlg det2), so K_e changes sign."""
    rng = np.random.default_rng(3)
    X, th, d = random_quad(rng)[::-1].copy(), 0.7, d_entries(fo.PLANE_STRESS)
    Ke = oracle_ke("quad", X, th, fo.PLANE_STRESS)
    assert Ke[0, 0] < 0
    blocks = quad_row_blocks(X[0], X[1], X[2], X[3], th, d)
    for b in range(4):
        np.testing.assert_allclose(blocks[b], Ke[0:2, 2 * b : 2 * b + 2], rtol=0, atol=2e-13 * np.abs(Ke).max())
