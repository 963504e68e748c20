// Per-element finite-element math in FP64 registers (2-D solid mechanics, 2 states per node).
//
// What the reference evaluates stage by stage through eight Numba kernels and ~1.3 GB of
// temporaries (FEMpy/Elements/Element.py:426-486, :918-1168;
// FEMpy/Constitutive/ConstitutiveModel.py:240-460) is fused here into one pass per Gauss point:
//
//   J = dNdxi X_e            (Element.py:942-978)      signed det, adjugate inverse (LinAlg.py:70-158)
//   G = J^-1 dNdxi           (Element.py:1021-1048)    u' = (G q_e)^T (Element.py:981-1018)
//   B_a = S(u') G_a          S = d eps / d u'  (ContinuumStrains.py:131-174); constant 0/1 when linear
//   K_ab += w detJ theta  B_a^T D B_b                  (ConstitutiveModel.py:389-460 + Element.py:1051-1111,1147-1168)
//   R_a  += w detJ theta  B_a^T sigma,  sigma = D eps  (ConstitutiveModel.py:345-386 + Element.py:1114-1168)
//
// Tables (N, dN/dxi, w at the element's own Gauss points) arrive as a __grid_constant__ kernel
// parameter, i.e. they sit in the constant bank and feed DFMA operands directly.
#pragma once
#include <cstdint>

namespace femb {

template <int NN, int NQ>
struct Tables {
    double dN[NQ][2][NN];  // dN[p][d][a] = dN_a / dxi_d at Gauss point p
    double N[NQ][NN];
    double w[NQ];
};

// Stress-strain matrix entries (isotropic plane models: D02 = D12 = 0, symmetric) + function params
struct DMat {
    double d00, d01, d11, d22;
    double nu, rho;
    int model;      // FEMB_PLANE_STRESS / FEMB_PLANE_STRAIN
    int nonlinear;  // Green strain
};

// the Gauss-point loop is fully unrolled except for the largest families (register pressure)
template <int NN, int NQ>
__host__ __device__ constexpr int point_unroll() {
    return (NN * NQ <= 81) ? NQ : 1;
}

// 1/d to within an ulp or two: hardware seed (rcp.approx.ftz.f64, ~20 bits) + two Newton steps; the IEEE-exact
// division costs ~25 instructions per Gauss point and the result feeds a 1e-12 tolerance
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    r = fma(fma(-d, r, 1.0), r, r);
    r = fma(fma(-d, r, 1.0), r, r);
    return r;
}

// G0[a] = dN_a/dx, G1[a] = dN_a/dy at Gauss point p; returns signed detJ
template <int NN, int NQ>
__device__ __forceinline__ double point_gradients(const Tables<NN, NQ>& tb, int p, const double (&x)[NN],
                                                   const double (&y)[NN], double (&G0)[NN], double (&G1)[NN]) {
    double J00 = 0.0, J01 = 0.0, J10 = 0.0, J11 = 0.0;
#pragma unroll
    for (int a = 0; a < NN; ++a) {
        J00 = fma(tb.dN[p][0][a], x[a], J00);
        J01 = fma(tb.dN[p][0][a], y[a], J01);
        J10 = fma(tb.dN[p][1][a], x[a], J10);
        J11 = fma(tb.dN[p][1][a], y[a], J11);
    }
    const double det = J00 * J11 - J01 * J10;
    const double idet = fast_rcp(det);
    const double i00 = idet * J11, i11 = idet * J00, i01 = -idet * J01, i10 = -idet * J10;
#pragma unroll
    for (int a = 0; a < NN; ++a) {
        G0[a] = i00 * tb.dN[p][0][a] + i01 * tb.dN[p][1][a];
        G1[a] = i10 * tb.dN[p][0][a] + i11 * tb.dN[p][1][a];
    }
    return det;
}

// state gradient u[s][d] = d u_s / d x_d
template <int NN>
__device__ __forceinline__ void state_gradient(const double (&G0)[NN], const double (&G1)[NN], const double (&ux)[NN],
                                               const double (&uy)[NN], double (&u)[2][2]) {
    u[0][0] = u[0][1] = u[1][0] = u[1][1] = 0.0;
#pragma unroll
    for (int a = 0; a < NN; ++a) {
        u[0][0] = fma(G0[a], ux[a], u[0][0]);
        u[0][1] = fma(G1[a], ux[a], u[0][1]);
        u[1][0] = fma(G0[a], uy[a], u[1][0]);
        u[1][1] = fma(G1[a], uy[a], u[1][1]);
    }
}

// stresses sigma = D eps(u')  (ContinuumStrains.py:90-128, IsoStress.py:157-251)
template <bool NL>
__device__ __forceinline__ void stresses(const DMat& D, const double (&u)[2][2], double (&sig)[3]) {
    double e0 = u[0][0], e1 = u[1][1], e2 = u[0][1] + u[1][0];
    if (NL) {
        e0 += 0.5 * (u[0][0] * u[0][0] + u[1][0] * u[1][0]);
        e1 += 0.5 * (u[0][1] * u[0][1] + u[1][1] * u[1][1]);
        e2 += u[0][0] * u[0][1] + u[1][0] * u[1][1];
    }
    sig[0] = D.d00 * e0 + D.d01 * e1;
    sig[1] = D.d01 * e0 + D.d11 * e1;
    sig[2] = D.d22 * e2;
}

// strain-displacement rows of node a:  B[o][s] = sum_d S[o][s][d] G_d,a
template <bool NL>
__device__ __forceinline__ void bmat(const double g0, const double g1, const double (&u)[2][2], double (&B)[3][2]) {
    if (NL) {
        const double f00 = 1.0 + u[0][0], f11 = 1.0 + u[1][1];
        B[0][0] = f00 * g0;
        B[0][1] = u[1][0] * g0;
        B[1][0] = u[0][1] * g1;
        B[1][1] = f11 * g1;
        B[2][0] = u[0][1] * g0 + f00 * g1;
        B[2][1] = f11 * g0 + u[1][0] * g1;
    } else {
        B[0][0] = g0;
        B[0][1] = 0.0;
        B[1][0] = 0.0;
        B[1][1] = g1;
        B[2][0] = g1;
        B[2][1] = g0;
    }
}

__device__ __forceinline__ double von_mises(const DMat& D, const double (&s)[3]) {
    if (D.model == 0) {  // plane stress (VonMises.py:26-43)
        return sqrt(s[0] * s[0] - s[0] * s[1] + s[1] * s[1] + 3.0 * s[2] * s[2]);
    }
    const double s33 = D.nu * (s[0] + s[1]);  // plane strain (VonMises.py:46-66)
    const double a = s[0] - s[1], b = s[1] - s33, c = s33 - s[0];
    return sqrt(0.5 * (a * a + b * b + c * c) + 3.0 * s[2] * s[2]);
}

// ---------------------------------------------------------------------------------------------
// Whole element, one thread: symmetric storage.
//   Kd[a][0..2]  = diagonal node block (a,a): [xx, xy(=yx), yy]
//   Ko[pair][4]  = off-diagonal node block (a,b), a<b, row-major [xx, xy, yx, yy]; (b,a) is its transpose
//   Re[a][2]
// ---------------------------------------------------------------------------------------------
template <int NN>
__host__ __device__ constexpr int pair_index(int a, int b) {  // a < b
    return a * NN - a * (a + 1) / 2 + (b - a - 1);
}

template <int NN, int NQ, bool NL, bool WANT_K, bool WANT_R>
__device__ __forceinline__ void element_sym(const Tables<NN, NQ>& tb, const DMat& D, const double (&x)[NN],
                                            const double (&y)[NN], const double (&ux)[NN], const double (&uy)[NN],
                                            const double theta, double (&Kd)[NN][3],
                                            double (&Ko)[NN * (NN - 1) / 2 + 1][4], double (&Re)[NN][2]) {
#pragma unroll
    for (int a = 0; a < NN; ++a) {
        Kd[a][0] = Kd[a][1] = Kd[a][2] = 0.0;
        Re[a][0] = Re[a][1] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < NN * (NN - 1) / 2 + 1; ++k) Ko[k][0] = Ko[k][1] = Ko[k][2] = Ko[k][3] = 0.0;

#pragma unroll(point_unroll<NN, NQ>())
    for (int p = 0; p < NQ; ++p) {
        double G0[NN], G1[NN];
        const double det = point_gradients<NN, NQ>(tb, p, x, y, G0, G1);
        const double sc = tb.w[p] * det * theta;
        double u[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        if (NL || WANT_R) state_gradient<NN>(G0, G1, ux, uy, u);
        if (WANT_R) {
            double sig[3];
            stresses<NL>(D, u, sig);
#pragma unroll
            for (int a = 0; a < NN; ++a) {
                double B[3][2];
                bmat<NL>(G0[a], G1[a], u, B);
                Re[a][0] = fma(sc, B[0][0] * sig[0] + B[1][0] * sig[1] + B[2][0] * sig[2], Re[a][0]);
                Re[a][1] = fma(sc, B[0][1] * sig[0] + B[1][1] * sig[1] + B[2][1] * sig[2], Re[a][1]);
            }
        }
        if (WANT_K) {
            if (!NL) {
                // linear model: D is constant over the element, so only the geometric sums
                //   S00 = sum sc G0a G0b, S01 = sum sc G0a G1b, S10 = sum sc G1a G0b, S11 = sum sc G1a G1b
                // are accumulated over the Gauss points (4 FMAs per node pair and point); D is applied once, below.
                // Storage while accumulating: Kd[a] = {S00, S01(=S10), S11}, Ko[pair] = {S00, S01, S10, S11}.
#pragma unroll
                for (int a = 0; a < NN; ++a) {
                    const double A0 = G0[a] * sc, A1 = G1[a] * sc;
                    Kd[a][0] = fma(A0, G0[a], Kd[a][0]);
                    Kd[a][1] = fma(A0, G1[a], Kd[a][1]);
                    Kd[a][2] = fma(A1, G1[a], Kd[a][2]);
#pragma unroll
                    for (int b = a + 1; b < NN; ++b) {
                        double(&k)[4] = Ko[pair_index<NN>(a, b)];
                        k[0] = fma(A0, G0[b], k[0]);
                        k[1] = fma(A0, G1[b], k[1]);
                        k[2] = fma(A1, G0[b], k[2]);
                        k[3] = fma(A1, G1[b], k[3]);
                    }
                }
            } else {
                double DB[NN][3][2];  // D B_b
                double Bs[NN][3][2];  // sc * B_a
#pragma unroll
                for (int a = 0; a < NN; ++a) {
                    double B[3][2];
                    bmat<NL>(G0[a], G1[a], u, B);
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        DB[a][0][s] = D.d00 * B[0][s] + D.d01 * B[1][s];
                        DB[a][1][s] = D.d01 * B[0][s] + D.d11 * B[1][s];
                        DB[a][2][s] = D.d22 * B[2][s];
                        Bs[a][0][s] = sc * B[0][s];
                        Bs[a][1][s] = sc * B[1][s];
                        Bs[a][2][s] = sc * B[2][s];
                    }
                }
#pragma unroll
                for (int a = 0; a < NN; ++a) {
                    Kd[a][0] += Bs[a][0][0] * DB[a][0][0] + Bs[a][1][0] * DB[a][1][0] + Bs[a][2][0] * DB[a][2][0];
                    Kd[a][1] += Bs[a][0][0] * DB[a][0][1] + Bs[a][1][0] * DB[a][1][1] + Bs[a][2][0] * DB[a][2][1];
                    Kd[a][2] += Bs[a][0][1] * DB[a][0][1] + Bs[a][1][1] * DB[a][1][1] + Bs[a][2][1] * DB[a][2][1];
#pragma unroll
                    for (int b = a + 1; b < NN; ++b) {
                        double(&k)[4] = Ko[pair_index<NN>(a, b)];
#pragma unroll
                        for (int s = 0; s < 2; ++s)
#pragma unroll
                            for (int S = 0; S < 2; ++S)
                                k[2 * s + S] += Bs[a][0][s] * DB[b][0][S] + Bs[a][1][s] * DB[b][1][S] + Bs[a][2][s] * DB[b][2][S];
                    }
                }
            }
        }
    }
    if (WANT_K && !NL) {  // K_ab = [[D00 S00 + D22 S11, D01 S01 + D22 S10], [D01 S10 + D22 S01, D11 S11 + D22 S00]]
#pragma unroll
        for (int a = 0; a < NN; ++a) {
            const double s00 = Kd[a][0], s01 = Kd[a][1], s11 = Kd[a][2];
            Kd[a][0] = fma(D.d00, s00, D.d22 * s11);
            Kd[a][1] = (D.d01 + D.d22) * s01;
            Kd[a][2] = fma(D.d11, s11, D.d22 * s00);
        }
#pragma unroll
        for (int q = 0; q < NN * (NN - 1) / 2; ++q) {
            const double s00 = Ko[q][0], s01 = Ko[q][1], s10 = Ko[q][2], s11 = Ko[q][3];
            Ko[q][0] = fma(D.d00, s00, D.d22 * s11);
            Ko[q][1] = fma(D.d01, s01, D.d22 * s10);
            Ko[q][2] = fma(D.d01, s10, D.d22 * s01);
            Ko[q][3] = fma(D.d11, s11, D.d22 * s00);
        }
    }
}

// 2x2 node block (a,b) of a symmetric-stored element matrix, row-major [xx, xy, yx, yy]
template <int NN>
__device__ __forceinline__ void sym_block(const double (&Kd)[NN][3], const double (&Ko)[NN * (NN - 1) / 2 + 1][4],
                                          int a, int b, double (&v)[4]) {
    if (a == b) {
        v[0] = Kd[a][0];
        v[1] = Kd[a][1];
        v[2] = Kd[a][1];
        v[3] = Kd[a][2];
    } else if (a < b) {
        const double(&k)[4] = Ko[pair_index<NN>(a, b)];
        v[0] = k[0];
        v[1] = k[1];
        v[2] = k[2];
        v[3] = k[3];
    } else {
        const double(&k)[4] = Ko[pair_index<NN>(b, a)];
        v[0] = k[0];
        v[1] = k[2];
        v[2] = k[1];
        v[3] = k[3];
    }
}


// ---------------------------------------------------------------------------------------------
// One node-row strip of an element, one thread: Krow[b][4] = block (a,b) for all b, Ra[2].
// Used for the larger families (9+ nodes) where a whole element does not fit in registers.
// ---------------------------------------------------------------------------------------------
template <int NN, int NQ, bool NL, bool WANT_K, bool WANT_R>
__device__ __forceinline__ void element_row(const Tables<NN, NQ>& tb, const DMat& D, const int a, const double (&x)[NN],
                                            const double (&y)[NN], const double (&ux)[NN], const double (&uy)[NN],
                                            const double theta, double (&Krow)[NN][4], double (&Ra)[2]) {
#pragma unroll
    for (int b = 0; b < NN; ++b) Krow[b][0] = Krow[b][1] = Krow[b][2] = Krow[b][3] = 0.0;
    Ra[0] = Ra[1] = 0.0;
#pragma unroll(point_unroll<NN, NQ>())
    for (int p = 0; p < NQ; ++p) {
        double G0[NN], G1[NN];
        const double det = point_gradients<NN, NQ>(tb, p, x, y, G0, G1);
        const double sc = tb.w[p] * det * theta;
        double u[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        if (NL || WANT_R) state_gradient<NN>(G0, G1, ux, uy, u);
        // G of "my" node: runtime index a -> select without dynamic register indexing
        double g0 = 0.0, g1 = 0.0;
#pragma unroll
        for (int b = 0; b < NN; ++b) {
            g0 = (b == a) ? G0[b] : g0;
            g1 = (b == a) ? G1[b] : g1;
        }
        double Ba[3][2];
        bmat<NL>(g0, g1, u, Ba);
        if (WANT_R) {
            double sig[3];
            stresses<NL>(D, u, sig);
            Ra[0] = fma(sc, Ba[0][0] * sig[0] + Ba[1][0] * sig[1] + Ba[2][0] * sig[2], Ra[0]);
            Ra[1] = fma(sc, Ba[0][1] * sig[0] + Ba[1][1] * sig[1] + Ba[2][1] * sig[2], Ra[1]);
        }
        if (WANT_K) {
            if (!NL) {
                const double A0 = g0 * sc, A1 = g1 * sc;
                const double c00 = A0 * D.d00, c01 = A0 * D.d01, c10 = A1 * D.d01, c11 = A1 * D.d11;
                const double s0 = A0 * D.d22, s1 = A1 * D.d22;
#pragma unroll
                for (int b = 0; b < NN; ++b) {
                    Krow[b][0] = fma(c00, G0[b], fma(s1, G1[b], Krow[b][0]));
                    Krow[b][1] = fma(c01, G1[b], fma(s1, G0[b], Krow[b][1]));
                    Krow[b][2] = fma(c10, G0[b], fma(s0, G1[b], Krow[b][2]));
                    Krow[b][3] = fma(c11, G1[b], fma(s0, G0[b], Krow[b][3]));
                }
            } else {
                double DBa[3][2];  // sc * D B_a  (D symmetric: B_a^T D B_b = (D B_a)^T B_b)
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    DBa[0][s] = sc * (D.d00 * Ba[0][s] + D.d01 * Ba[1][s]);
                    DBa[1][s] = sc * (D.d01 * Ba[0][s] + D.d11 * Ba[1][s]);
                    DBa[2][s] = sc * (D.d22 * Ba[2][s]);
                }
#pragma unroll
                for (int b = 0; b < NN; ++b) {
                    double Bb[3][2];
                    bmat<NL>(G0[b], G1[b], u, Bb);
#pragma unroll
                    for (int s = 0; s < 2; ++s)
#pragma unroll
                        for (int S = 0; S < 2; ++S)
                            Krow[b][2 * s + S] += DBa[0][s] * Bb[0][S] + DBa[1][s] * Bb[1][S] + DBa[2][s] * Bb[2][S];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Element functions at the Gauss points (Element.py:163-273, IsoPlaneStress.py:187-227): mass density or von Mises
// stress, plus signed detJ for the "integrate" reduction.
// The state gradient needs no per-node dN/dx: with H = sum_b dN_b (x) q_b (reference-space gradient of the state) and
// adj = det * J^-1, u' = adj . H / det.  The point loop is kept rolled so that the tables reach the FP64 pipe through
// uniform registers (an FP64 instruction with a constant-bank operand issues at half rate on sm_100).
// ---------------------------------------------------------------------------------------------
template <int NN, int NQ, bool NL>
__device__ __forceinline__ void point_function(const Tables<NN, NQ>& tb, const DMat& D, const int func, const int p,
                                               const double (&x)[NN], const double (&y)[NN], const double (&ux)[NN],
                                               const double (&uy)[NN], double& val, double& detOut) {
    {
        double J00 = 0.0, J01 = 0.0, J10 = 0.0, J11 = 0.0;
#pragma unroll
        for (int b = 0; b < NN; ++b) {
            J00 = fma(tb.dN[p][0][b], x[b], J00);
            J01 = fma(tb.dN[p][0][b], y[b], J01);
            J10 = fma(tb.dN[p][1][b], x[b], J10);
            J11 = fma(tb.dN[p][1][b], y[b], J11);
        }
        const double det = J00 * J11 - J01 * J10;
        double v = D.rho;
        if (func != 0) {
            double H00 = 0.0, H01 = 0.0, H10 = 0.0, H11 = 0.0;  // H[k][s] = sum_b dN_b/dxi_k * q_b[s]
#pragma unroll
            for (int b = 0; b < NN; ++b) {
                H00 = fma(tb.dN[p][0][b], ux[b], H00);
                H01 = fma(tb.dN[p][0][b], uy[b], H01);
                H10 = fma(tb.dN[p][1][b], ux[b], H10);
                H11 = fma(tb.dN[p][1][b], uy[b], H11);
            }
            const double r = fast_rcp(det);
            double u[2][2], sig[3];  // u[s][d] = d u_s / d x_d = (adj . H)[d][s] / det, adj = [[J11, -J01], [-J10, J00]]
            u[0][0] = (J11 * H00 - J01 * H10) * r;
            u[0][1] = (J00 * H10 - J10 * H00) * r;
            u[1][0] = (J11 * H01 - J01 * H11) * r;
            u[1][1] = (J00 * H11 - J10 * H01) * r;
            stresses<NL>(D, u, sig);
            v = von_mises(D, sig);
        }
        val = v;
        detOut = det;
    }
}

}  // namespace femb
