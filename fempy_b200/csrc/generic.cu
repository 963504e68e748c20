// Generic-dimension path: the 3-D and 1-D families -- HexElement3D order 1-2 with Iso3D
// (FEMpy/Elements/HexElement3D.py, FEMpy/Constitutive/Iso3D.py), LineElement1D order 1-3 with Iso1D (LineElement1D.py,
// Iso1D.py) -- behind the same plan / pattern / C ABI as the 2-D kernels.  Node and point counts are run-time values, the
// tables live in global memory; the kernels restate the reference's evaluation point by point:
//   J = dN X, det, J^-1 (LinAlg.py:70-189), G = J^-1 dN, u' = (G q)^T               (Element.py:942-1048)
//   strains and S = d strain / d u' (ContinuumStrains.py:29-85 1-D, :180-328 3-D), sigma = D eps (IsoStress.py:81-151,259-298)
//   weak residual theta S^T sigma, weak Jacobian theta S^T D S                       (ConstitutiveModel.py:240-460)
//   r = G^T (.), K = G^T (.) G, integrated with det J * w                            (Element.py:1051-1168)
// Parallelisation is the plain one -- a thread per (element, node a, node b) block of K_e, scattered with FP64 atomics
// through the slot map -- because these families are parity targets here, not the optimised configurations.
#include "femb_internal.h"

namespace gen {

constexpr int NNMAX = 27;  // 27-node hexahedron

struct Mat {
    double D[6][6];  // stress-strain matrix, ns x ns used
    double E, nu, rho;
    int nonlinear;
};

template <int DIM>
struct Args {
    long long E;
    long long elemOffset;
    int nn, nq;
    const int* conn;      // node-major
    const int* slot;      // may be NULL (element blocks)
    const int* nrowptr;
    const double* N;      // (nq, nn)
    const double* dN;     // (nq, DIM, nn)
    const double* w;
    const double* coords;  // (numNodes, DIM)
    const double* states;  // (numNodes, DIM)
    const double* scale;   // per element (thickness / area / 1), global element id
    double* K;             // CSR values (scatter) or NULL
    double* R;
    double* Ke;            // (E, nn*DIM, nn*DIM) of this block or NULL
    double* Re;            // (E, nn, DIM)
    double* out;           // function values
};

// signed determinant and inverse (LinAlg.py:54-189)
template <int DIM>
__device__ __forceinline__ double inverse(const double (&J)[DIM][DIM], double (&I)[DIM][DIM]) {
    if constexpr (DIM == 1) {
        I[0][0] = 1.0 / J[0][0];
        return J[0][0];
    } else if constexpr (DIM == 2) {
        const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0], r = 1.0 / det;
        I[0][0] = r * J[1][1], I[1][1] = r * J[0][0], I[0][1] = -r * J[0][1], I[1][0] = -r * J[1][0];
        return det;
    } else {
        const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][0] * J[2][2] - J[1][2] * J[2][0], c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        const double det = J[0][0] * c00 - J[0][1] * c01 + J[0][2] * c02, r = 1.0 / det;
        I[0][0] = r * c00;
        I[0][1] = -r * (J[0][1] * J[2][2] - J[0][2] * J[2][1]);
        I[0][2] = r * (J[0][1] * J[1][2] - J[0][2] * J[1][1]);
        I[1][0] = -r * c01;
        I[1][1] = r * (J[0][0] * J[2][2] - J[0][2] * J[2][0]);
        I[1][2] = -r * (J[0][0] * J[1][2] - J[0][2] * J[1][0]);
        I[2][0] = r * c02;
        I[2][1] = -r * (J[0][0] * J[2][1] - J[0][1] * J[2][0]);
        I[2][2] = r * (J[0][0] * J[1][1] - J[0][1] * J[1][0]);
        return det;
    }
}

// the shear pairs of the engineering strains, in the reference's order: 2-D [xx, yy, xy], 3-D [xx, yy, zz, xy, xz, yz]
template <int DIM>
__device__ __forceinline__ void shear_pair(int k, int& i, int& j) {
    if constexpr (DIM == 2) {
        i = 0, j = 1;
    } else {
        i = (k == 2) ? 1 : 0;
        j = (k == 0) ? 1 : 2;
    }
}

// strains eps[o] and sensitivities S[o][s][d] = d eps_o / d u'[s][d] (u'[s][d] = d u_s / d x_d); Green terms when nl
template <int DIM>
__device__ __forceinline__ void strain_model(const double (&u)[DIM][DIM], bool nl, double (&eps)[DIM * (DIM + 1) / 2],
                                             double (&S)[DIM * (DIM + 1) / 2][DIM][DIM]) {
    constexpr int NS = DIM * (DIM + 1) / 2;
    for (int o = 0; o < NS; ++o)
        for (int s = 0; s < DIM; ++s)
            for (int d = 0; d < DIM; ++d) S[o][s][d] = 0.0;
    for (int i = 0; i < DIM; ++i) {
        eps[i] = u[i][i];
        S[i][i][i] = 1.0;
        if (nl) {
            for (int s = 0; s < DIM; ++s) {
                eps[i] += 0.5 * u[s][i] * u[s][i];
                S[i][s][i] += u[s][i];
            }
        }
    }
    for (int k = 0; k < NS - DIM; ++k) {
        int i, j;
        shear_pair<DIM>(k, i, j);
        eps[DIM + k] = u[i][j] + u[j][i];
        S[DIM + k][i][j] = 1.0;
        S[DIM + k][j][i] = 1.0;
        if (nl) {
            for (int s = 0; s < DIM; ++s) {
                eps[DIM + k] += u[s][i] * u[s][j];
                S[DIM + k][s][i] += u[s][j];
                S[DIM + k][s][j] += u[s][i];
            }
        }
    }
}

// everything of one Gauss point that does not depend on the node pair: J^-1, det, and (when wanted) u'
template <int DIM>
__device__ __forceinline__ double point_geometry(const Args<DIM>& A, int p, const double* X, const double* q, bool wantU,
                                                 double (&I)[DIM][DIM], double (&u)[DIM][DIM]) {
    const double* dNp = A.dN + (size_t)p * DIM * A.nn;
    double J[DIM][DIM];
    for (int i = 0; i < DIM; ++i)
        for (int j = 0; j < DIM; ++j) J[i][j] = 0.0;
    for (int n = 0; n < A.nn; ++n)
        for (int i = 0; i < DIM; ++i) {
            const double g = __ldg(dNp + i * A.nn + n);
            for (int j = 0; j < DIM; ++j) J[i][j] = fma(g, X[n * DIM + j], J[i][j]);
        }
    const double det = inverse<DIM>(J, I);
    if (wantU) {
        double H[DIM][DIM];  // H[k][s] = sum_n dN_n/dxi_k q_n[s]
        for (int k = 0; k < DIM; ++k)
            for (int s = 0; s < DIM; ++s) H[k][s] = 0.0;
        for (int n = 0; n < A.nn; ++n)
            for (int k = 0; k < DIM; ++k) {
                const double g = __ldg(dNp + k * A.nn + n);
                for (int s = 0; s < DIM; ++s) H[k][s] = fma(g, q[n * DIM + s], H[k][s]);
            }
        for (int s = 0; s < DIM; ++s)
            for (int d = 0; d < DIM; ++d) {
                double v = 0.0;
                for (int k = 0; k < DIM; ++k) v = fma(I[d][k], H[k][s], v);
                u[s][d] = v;
            }
    }
    return det;
}

template <int DIM>
__device__ __forceinline__ void load_element(const Args<DIM>& A, long long e, double* X, double* q, bool wantQ) {
    for (int n = 0; n < A.nn; ++n) {
        const int node = __ldg(A.conn + (long long)n * A.E + e);
        for (int j = 0; j < DIM; ++j) {
            X[n * DIM + j] = __ldg(A.coords + (long long)node * DIM + j);
            if (wantQ) q[n * DIM + j] = __ldg(A.states + (long long)node * DIM + j);
        }
    }
}

// B_n[o][s] = sum_d S[o][s][d] G_n[d], G_n = J^-1 dN_n
template <int DIM>
__device__ __forceinline__ void b_matrix(const Args<DIM>& A, int p, int n, const double (&I)[DIM][DIM],
                                         const double (&S)[DIM * (DIM + 1) / 2][DIM][DIM], double (&B)[DIM * (DIM + 1) / 2][DIM]) {
    constexpr int NS = DIM * (DIM + 1) / 2;
    const double* dNp = A.dN + (size_t)p * DIM * A.nn;
    double G[DIM];
    for (int d = 0; d < DIM; ++d) {
        double v = 0.0;
        for (int k = 0; k < DIM; ++k) v = fma(I[d][k], __ldg(dNp + k * A.nn + n), v);
        G[d] = v;
    }
    for (int o = 0; o < NS; ++o)
        for (int s = 0; s < DIM; ++s) {
            double v = 0.0;
            for (int d = 0; d < DIM; ++d) v = fma(S[o][s][d], G[d], v);
            B[o][s] = v;
        }
}

// one thread per (element, a, b): the DIM x DIM block K_e[a][b]
template <int DIM>
__global__ void __launch_bounds__(128) k_block(const Args<DIM> A, const Mat M) {
    constexpr int NS = DIM * (DIM + 1) / 2;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pairs = (long long)A.nn * A.nn;
    if (t >= A.E * pairs) return;
    const long long e = t / pairs;
    const int ab = (int)(t - e * pairs), a = ab / A.nn, b = ab - a * A.nn;
    double X[NNMAX * DIM], q[NNMAX * DIM];
    load_element<DIM>(A, e, X, q, M.nonlinear != 0);
    const double theta = A.scale ? __ldg(A.scale + A.elemOffset + e) : 1.0;
    double Kab[DIM][DIM];
    for (int s = 0; s < DIM; ++s)
        for (int c = 0; c < DIM; ++c) Kab[s][c] = 0.0;
    for (int p = 0; p < A.nq; ++p) {
        double I[DIM][DIM], u[DIM][DIM], eps[NS], S[NS][DIM][DIM], Ba[NS][DIM], Bb[NS][DIM];
        for (int s = 0; s < DIM; ++s)
            for (int d = 0; d < DIM; ++d) u[s][d] = 0.0;
        const double det = point_geometry<DIM>(A, p, X, q, M.nonlinear != 0, I, u);
        strain_model<DIM>(u, M.nonlinear != 0, eps, S);
        b_matrix<DIM>(A, p, a, I, S, Ba);
        b_matrix<DIM>(A, p, b, I, S, Bb);
        const double c = theta * det * __ldg(A.w + p);
        for (int s = 0; s < DIM; ++s) {
            double DB[NS];  // (D B_a)[o][s]
            for (int o = 0; o < NS; ++o) {
                double v = 0.0;
                for (int k = 0; k < NS; ++k) v = fma(M.D[o][k], Ba[k][s], v);
                DB[o] = v;
            }
            for (int cc = 0; cc < DIM; ++cc) {
                double v = 0.0;
                for (int o = 0; o < NS; ++o) v = fma(DB[o], Bb[o][cc], v);
                Kab[s][cc] = fma(c, v, Kab[s][cc]);
            }
        }
    }
    if (A.Ke) {
        const int nd = A.nn * DIM;
        double* base = A.Ke + (size_t)e * nd * nd;
        for (int s = 0; s < DIM; ++s)
            for (int c = 0; c < DIM; ++c) base[(size_t)(a * DIM + s) * nd + b * DIM + c] = Kab[s][c];
    }
    if (A.K) {
        const int node = __ldg(A.conn + (long long)a * A.E + e);
        const int off = __ldg(A.nrowptr + node), deg = __ldg(A.nrowptr + node + 1) - off;
        const int pos = __ldg(A.slot + (long long)ab * A.E + e);
        double* row = A.K + (size_t)DIM * DIM * off + (size_t)DIM * (pos - off);
        for (int s = 0; s < DIM; ++s)
            for (int c = 0; c < DIM; ++c) atomicAdd(row + (size_t)s * DIM * deg + c, Kab[s][c]);
    }
}

// one thread per (element, a): R_e[a]
template <int DIM>
__global__ void __launch_bounds__(128) k_residual(const Args<DIM> A, const Mat M) {
    constexpr int NS = DIM * (DIM + 1) / 2;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= A.E * A.nn) return;
    const long long e = t / A.nn;
    const int a = (int)(t - e * A.nn);
    double X[NNMAX * DIM], q[NNMAX * DIM];
    load_element<DIM>(A, e, X, q, true);
    const double theta = A.scale ? __ldg(A.scale + A.elemOffset + e) : 1.0;
    double Ra[DIM];
    for (int s = 0; s < DIM; ++s) Ra[s] = 0.0;
    for (int p = 0; p < A.nq; ++p) {
        double I[DIM][DIM], u[DIM][DIM], eps[NS], S[NS][DIM][DIM], Ba[NS][DIM], sig[NS];
        const double det = point_geometry<DIM>(A, p, X, q, true, I, u);
        strain_model<DIM>(u, M.nonlinear != 0, eps, S);
        for (int o = 0; o < NS; ++o) {
            double v = 0.0;
            for (int k = 0; k < NS; ++k) v = fma(M.D[k][o], eps[k], v);  // einsum "as,pa->ps" (IsoStress.py:273)
            sig[o] = v;
        }
        b_matrix<DIM>(A, p, a, I, S, Ba);
        const double c = theta * det * __ldg(A.w + p);
        for (int s = 0; s < DIM; ++s) {
            double v = 0.0;
            for (int o = 0; o < NS; ++o) v = fma(Ba[o][s], sig[o], v);
            Ra[s] = fma(c, v, Ra[s]);
        }
    }
    const int node = __ldg(A.conn + (long long)a * A.E + e);
    for (int s = 0; s < DIM; ++s) {
        if (A.Re) A.Re[((size_t)e * A.nn + a) * DIM + s] = Ra[s];
        if (A.R) atomicAdd(A.R + (size_t)node * DIM + s, Ra[s]);
    }
}

// value of the element function at one point
template <int DIM>
__device__ __forceinline__ double point_value(const Mat& M, int func, const double (&u)[DIM][DIM]) {
    constexpr int NS = DIM * (DIM + 1) / 2;
    if (func == FEMB_FUNC_MASS) return M.rho;
    double eps[NS], S[NS][DIM][DIM], sig[NS];
    strain_model<DIM>(u, M.nonlinear != 0, eps, S);
    for (int o = 0; o < NS; ++o) {
        double v = 0.0;
        for (int k = 0; k < NS; ++k) v = fma(M.D[k][o], eps[k], v);
        sig[o] = v;
    }
    if constexpr (DIM == 1) {
        return func == FEMB_FUNC_STRAIN ? eps[0] : sig[0];  // Iso1D.py:203-225
    } else if constexpr (DIM == 3) {  // VonMises.py:70-89
        const double a = sig[0] - sig[1], b = sig[1] - sig[2], c = sig[2] - sig[0];
        return sqrt(0.5 * (a * a + b * b + c * c) + 3.0 * (sig[3] * sig[3] + sig[4] * sig[4] + sig[5] * sig[5]));
    } else {
        return sqrt(sig[0] * sig[0] - sig[0] * sig[1] + sig[1] * sig[1] + 3.0 * sig[2] * sig[2]);
    }
}

// one thread per element: point values + reduction (Element.py:221-273)
template <int DIM>
__global__ void __launch_bounds__(128) k_function(const Args<DIM> A, const Mat M, const int func, const int reduction) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.E) return;
    double X[NNMAX * DIM], q[NNMAX * DIM];
    load_element<DIM>(A, e, X, q, true);
    double r = 0.0;
    for (int pass = 0; pass < (reduction == FEMB_REDUCE_KSMAX ? 2 : 1); ++pass) {
        double s = 0.0;
        for (int p = 0; p < A.nq; ++p) {
            double I[DIM][DIM], u[DIM][DIM];
            const double det = point_geometry<DIM>(A, p, X, q, true, I, u);
            const double v = point_value<DIM>(M, func, u);
            if (pass == 1) {
                s += exp(100.0 * (v - r));  // KSAgg.py:26,52
                continue;
            }
            if (reduction == FEMB_REDUCE_NONE) A.out[e * A.nq + p] = v;
            else if (reduction == FEMB_REDUCE_SUM || reduction == FEMB_REDUCE_MEAN) r += v;
            else if (reduction == FEMB_REDUCE_INTEGRATE) r = fma(v * det, __ldg(A.w + p), r);
            else if (reduction == FEMB_REDUCE_MIN) r = p ? fmin(r, v) : v;
            else r = p ? fmax(r, v) : v;
        }
        if (pass == 1) r += 1.0 / 100.0 * log(1.0 / (double)A.nq * s);
    }
    if (reduction == FEMB_REDUCE_NONE) return;
    if (reduction == FEMB_REDUCE_MEAN) r /= (double)A.nq;
    A.out[e] = r;
}

// one thread per unique BC DOF: zero its row, occurrence count on the diagonal, residual = u - c (AssemblyUtils.py:26-94)
__global__ void k_apply_bcs(long long nBC, int s, const int* __restrict__ bcDof, const double* __restrict__ bcVal,
                            const double* __restrict__ bcCount, const int* __restrict__ nrowptr, const int* __restrict__ ncols,
                            const double* __restrict__ states, double* __restrict__ K, double* __restrict__ R) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nBC) return;
    const int dof = bcDof[t], i = dof / s, r = dof - i * s;
    if (K) {
        const int off = nrowptr[i], deg = nrowptr[i + 1] - off;
        double* row = K + (size_t)s * s * off + (size_t)r * s * deg;
        for (int k = 0; k < deg; ++k)
            for (int c = 0; c < s; ++c) row[(size_t)s * k + c] = (ncols[off + k] == i && c == r) ? bcCount[t] : 0.0;
    }
    if (R) R[dof] = states[dof] - bcVal[t];
}

__global__ void k_init_residual(double* __restrict__ R, const double* __restrict__ loads, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) R[i] = loads ? -loads[i] : 0.0;
}

static Mat make_mat(const femb_material& m, int dim) {
    Mat M;
    memset(&M, 0, sizeof(M));
    M.E = m.E, M.nu = m.nu, M.rho = m.rho, M.nonlinear = m.nonlinear;
    if (dim == 1) {
        M.D[0][0] = m.E;  // IsoStress.py:115-151
    } else if (dim == 3) {  // threeDMat, IsoStress.py:81-112
        const double f = m.E / ((1.0 + m.nu) * (1.0 - 2.0 * m.nu));
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) M.D[i][j] = f * (i == j ? 1.0 - m.nu : m.nu);
        for (int i = 3; i < 6; ++i) M.D[i][i] = f * ((1.0 - 2.0 * m.nu) / 2.0);
    } else {
        const femb::DMat d = make_dmat(m);
        M.D[0][0] = d.d00, M.D[0][1] = M.D[1][0] = d.d01, M.D[1][1] = d.d11, M.D[2][2] = d.d22;
    }
    return M;
}

template <int DIM>
static Args<DIM> block_args(const femb_plan* p, const BlockData& b, const double* d_coords, const double* d_states, const double* d_scale) {
    Args<DIM> A;
    memset(&A, 0, sizeof(A));
    A.E = b.numElems;
    A.elemOffset = b.elemOffset;
    A.nn = b.nn, A.nq = b.nq;
    A.conn = b.d_conn;
    A.slot = b.d_slot;
    A.nrowptr = p->d_nrowptr;
    A.N = b.d_tabN, A.dN = b.d_tabdN, A.w = b.d_tabw;
    A.coords = d_coords, A.states = d_states, A.scale = d_scale;
    return A;
}

template <int DIM>
static int assemble_dim(femb_plan* p, const double* d_coords, const double* d_states, const double* d_scale, const Mat& M, bool bcs,
                        const double* d_loads, bool wk, bool wr, double* d_K, double* d_R) {
    cudaStream_t st = p->stream;
    const int64_t numDOF = femb_plan_num_dof(p), nnz = femb_plan_nnz(p);
    int64_t launches = 0;
    if (wk) CUDA_TRY(cudaMemsetAsync(d_K, 0, nnz * sizeof(double), st));
    if (wr) {
        k_init_residual<<<grid_for(numDOF, 256), 256, 0, st>>>(d_R, d_loads, numDOF);
        launches++;
    }
    for (auto& b : p->blocks) {
        Args<DIM> A = block_args<DIM>(p, b, d_coords, d_states, d_scale);
        if (wk) {
            A.K = d_K;
            k_block<DIM><<<grid_for(b.numElems * b.nn * b.nn, 128), 128, 0, st>>>(A, M);
            launches++;
        }
        if (wr) {
            A.K = nullptr;
            A.R = d_R;
            k_residual<DIM><<<grid_for(b.numElems * b.nn, 128), 128, 0, st>>>(A, M);
            launches++;
        }
    }
    if (bcs) {
        k_apply_bcs<<<grid_for(p->nBC, 128), 128, 0, st>>>(p->nBC, p->numStates, p->d_bcDof, p->d_bcVal, p->d_bcCount, p->d_nrowptr, p->d_ncols,
                                                            d_states, wk ? d_K : nullptr, wr ? d_R : nullptr);
        launches++;
    }
    p->launches = launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace gen

static int check_generic_model(const femb_plan* p, const femb_material* mat, const char* who) {
    const int want = p->numDim == 3 ? FEMB_ISO_3D : FEMB_ISO_1D;
    if (mat->model != want) return fail("%s: a %d-D plan needs constitutive model id %d (got %d)", who, p->numDim, want, mat->model);
    for (auto& b : p->blocks)
        if (b.nn > gen::NNMAX) return fail("%s: at most %d nodes per element on the generic path", who, gen::NNMAX);
    return 0;
}

// femb_assemble_dev for 1-D / 3-D plans (assemble.cu routes here)
int generic_assemble_dev(femb_plan* p, const double* d_coords, const double* d_states, const double* d_scale, const femb_material* mat,
                         bool bcs, const double* d_loads, bool wk, bool wr, double* d_K, double* d_R) {
    TRY(check_generic_model(p, mat, "femb_assemble_dev"));
    TRY(ensure_slot_maps(p));
    const gen::Mat M = gen::make_mat(*mat, p->numDim);
    if (p->numDim == 3) return gen::assemble_dim<3>(p, d_coords, d_states, d_scale, M, bcs, d_loads, wk, wr, d_K, d_R);
    return gen::assemble_dim<1>(p, d_coords, d_states, d_scale, M, bcs, d_loads, wk, wr, d_K, d_R);
}

template <int DIM>
static int function_dim(femb_plan* p, const double* d_coords, const double* d_states, const gen::Mat& M, int func, int reduction, double* d_out) {
    int64_t launches = 0;
    for (auto& b : p->blocks) {
        gen::Args<DIM> A = gen::block_args<DIM>(p, b, d_coords, d_states, nullptr);
        A.out = d_out + (reduction == FEMB_REDUCE_NONE ? b.pointOffset : b.elemOffset);
        gen::k_function<DIM><<<grid_for(b.numElems, 128), 128, 0, p->stream>>>(A, M, func, reduction);
        launches++;
    }
    p->launches = launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int generic_function_dev(femb_plan* p, const double* d_coords, const double* d_states, const femb_material* mat, int func, int reduction,
                         double* d_out) {
    TRY(check_generic_model(p, mat, "femb_function_dev"));
    const bool ok = func == FEMB_FUNC_MASS || (p->numDim == 3 && func == FEMB_FUNC_VON_MISES) ||
                    (p->numDim == 1 && (func == FEMB_FUNC_STRESS || func == FEMB_FUNC_STRAIN));
    if (!ok) return fail("femb_function_dev: function id %d is not defined for the %d-D model", func, p->numDim);
    const gen::Mat M = gen::make_mat(*mat, p->numDim);
    if (p->numDim == 3) return function_dim<3>(p, d_coords, d_states, M, func, reduction, d_out);
    return function_dim<1>(p, d_coords, d_states, M, func, reduction, d_out);
}

template <int DIM>
static int blocks_dim(femb_plan* p, const double* d_coords, const double* d_states, const double* d_scale, const gen::Mat& M, double* d_Ke,
                      double* d_Re) {
    int64_t launches = 0;
    for (auto& b : p->blocks) {
        gen::Args<DIM> A = gen::block_args<DIM>(p, b, d_coords, d_states, d_scale);
        A.slot = nullptr;
        if (d_Ke) {
            A.Ke = d_Ke + b.keOffset;
            gen::k_block<DIM><<<grid_for(b.numElems * b.nn * b.nn, 128), 128, 0, p->stream>>>(A, M);
            launches++;
        }
        if (d_Re) {
            A.Ke = nullptr;
            A.Re = d_Re + b.reOffset;
            gen::k_residual<DIM><<<grid_for(b.numElems * b.nn, 128), 128, 0, p->stream>>>(A, M);
            launches++;
        }
    }
    p->launches = launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int generic_element_blocks_dev(femb_plan* p, const double* d_coords, const double* d_states, const double* d_scale, const femb_material* mat,
                               double* d_Ke, double* d_Re) {
    TRY(check_generic_model(p, mat, "femb_element_blocks"));
    const gen::Mat M = gen::make_mat(*mat, p->numDim);
    if (p->numDim == 3) return blocks_dim<3>(p, d_coords, d_states, d_scale, M, d_Ke, d_Re);
    return blocks_dim<1>(p, d_coords, d_states, d_scale, M, d_Ke, d_Re);
}
