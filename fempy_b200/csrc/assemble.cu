// Fused K + R assembly kernels and their C ABI.
#include "femb_internal.h"

using namespace femb;

// =============================================================================================
// assembly kernels (v0: scatter through the slot map with FP64 RED atomics)
// =============================================================================================
template <int NN>
__device__ __forceinline__ void load_element(const AsmArgs& A, long long e, int (&node)[NN], double (&x)[NN],
                                             double (&y)[NN], double (&ux)[NN], double (&uy)[NN]) {
#pragma unroll
    for (int a = 0; a < NN; ++a) node[a] = __ldg(A.conn + (long long)a * A.E + e);
#pragma unroll
    for (int a = 0; a < NN; ++a) {
        const double2 c = __ldg(A.coords + node[a]);
        const double2 q = __ldg(A.states + node[a]);
        x[a] = c.x;
        y[a] = c.y;
        ux[a] = q.x;
        uy[a] = q.y;
    }
}

__device__ __forceinline__ void red_add(double* p, double v) { atomicAdd(p, v); }

template <int NN, int NQ, bool NL, bool WK, bool WR>
__global__ void __launch_bounds__(128) k_assemble_sym(const __grid_constant__ Tables<NN, NQ> tb, const DMat D,
                                                      const AsmArgs A) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= A.E) return;
    int node[NN];
    double x[NN], y[NN], ux[NN], uy[NN];
    load_element<NN>(A, e, node, x, y, ux, uy);
    const double theta = __ldg(A.thick + A.elemOffset + e);
    double Kd[NN][3], Ko[NN * (NN - 1) / 2 + 1][4], Re[NN][2];
    element_sym<NN, NQ, NL, WK, WR>(tb, D, x, y, ux, uy, theta, Kd, Ko, Re);
    if (WK) {
#pragma unroll
        for (int a = 0; a < NN; ++a) {
            const int off = __ldg(A.nrowptr + node[a]);
            const int deg = __ldg(A.nrowptr + node[a] + 1) - off;
#pragma unroll
            for (int b = 0; b < NN; ++b) {
                const int pos = __ldg(A.slot + (long long)(a * NN + b) * A.E + e);
                double v[4];
                sym_block<NN>(Kd, Ko, a, b, v);
                double* r0 = A.K + 4ll * off + 2 * (pos - off);
                double* r1 = r0 + 2 * deg;
                red_add(r0, v[0]);
                red_add(r0 + 1, v[1]);
                red_add(r1, v[2]);
                red_add(r1 + 1, v[3]);
            }
        }
    }
    if (WR) {
#pragma unroll
        for (int a = 0; a < NN; ++a) {
            red_add(A.R + 2ll * node[a], Re[a][0]);
            red_add(A.R + 2ll * node[a] + 1, Re[a][1]);
        }
    }
}

// one thread per (element, node row a); blockIdx.y = a so a warp holds consecutive elements
template <int NN, int NQ, bool NL, bool WK, bool WR>
__global__ void __launch_bounds__(128) k_assemble_row(const __grid_constant__ Tables<NN, NQ> tb, const DMat D,
                                                      const AsmArgs A) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int a = blockIdx.y;
    if (e >= A.E) return;
    int node[NN];
    double x[NN], y[NN], ux[NN], uy[NN];
    load_element<NN>(A, e, node, x, y, ux, uy);
    const double theta = __ldg(A.thick + A.elemOffset + e);
    double Krow[NN][4], Ra[2];
    element_row<NN, NQ, NL, WK, WR>(tb, D, a, x, y, ux, uy, theta, Krow, Ra);
    int me = 0;
#pragma unroll
    for (int b = 0; b < NN; ++b) me = (b == a) ? node[b] : me;
    if (WK) {
        const int off = __ldg(A.nrowptr + me);
        const int deg = __ldg(A.nrowptr + me + 1) - off;
#pragma unroll
        for (int b = 0; b < NN; ++b) {
            const int pos = __ldg(A.slot + (long long)(a * NN + b) * A.E + e);
            double* r0 = A.K + 4ll * off + 2 * (pos - off);
            double* r1 = r0 + 2 * deg;
            red_add(r0, Krow[b][0]);
            red_add(r0 + 1, Krow[b][1]);
            red_add(r1, Krow[b][2]);
            red_add(r1 + 1, Krow[b][3]);
        }
    }
    if (WR) {
        red_add(A.R + 2ll * me, Ra[0]);
        red_add(A.R + 2ll * me + 1, Ra[1]);
    }
}

// R = -loads (or 0) before the element residuals are accumulated (Problem.py:641,653)
__global__ void k_init_residual(double* __restrict__ R, const double* __restrict__ loads, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) R[i] = loads ? -loads[i] : 0.0;
}

// one warp per (unique) BC DOF: zero its row, set the diagonal to the occurrence count, and set the
// residual entry to u - c (AssemblyUtils.py:26-94)
__global__ void k_apply_bcs(long long nBC, const int* __restrict__ bcDof, const double* __restrict__ bcVal,
                            const double* __restrict__ bcCount, const int* __restrict__ nrowptr,
                            const int* __restrict__ ncols, const double* __restrict__ states, double* __restrict__ K,
                            double* __restrict__ R) {
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= nBC) return;
    const int dof = bcDof[wid];
    const int i = dof >> 1, r = dof & 1;
    if (K) {
        const int off = nrowptr[i], deg = nrowptr[i + 1] - off;
        double* row = K + 4ll * off + 2ll * r * deg;
        for (int k = lane; k < deg; k += 32) {
            const int j = ncols[off + k];
            row[2 * k] = (j == i && r == 0) ? bcCount[wid] : 0.0;
            row[2 * k + 1] = (j == i && r == 1) ? bcCount[wid] : 0.0;
        }
    }
    if (R && lane == 0) R[dof] = states[dof] - bcVal[wid];
}



// =============================================================================================
// write-once assembly over node patches (plan_patch.cu builds the programs)
//
// One CTA per patch.  Phase 1: one thread per patch element evaluates K_e (symmetric storage: NN diagonal
// + NN(NN-1)/2 off-diagonal 2x2 blocks) and R_e in FP64 registers and parks them in shared memory, element
// index fastest (conflict-free stores).  Phase 2: one thread per (owned node, neighbour node) block of the
// global matrix sums its sources and writes both CSR rows of the node with 16-byte stores; the thread on
// the diagonal block also sums the node's residual, subtracts the load and applies the BC.  No atomics, no
// zero-fill, every output written exactly once.
// =============================================================================================
struct PatchArgs {
    long long numNodes;
    int PN;
    const int* order;
    const int* nbStart;
    const unsigned short* nbRow;
    const int* nrowptr;
    const int* ncols;
    const int* pePtr;
    const int* peElem;
    const int* peConn;
    const int* srcBase;
    const unsigned short* srcPtr;
    const unsigned short* src;
    long long elemOffset;
    const double2* coords;
    const double2* states;
    const double* thick;
    const double* loads;
    const unsigned char* bcMask;  // NULL: no BCs in this pass
    const int* bcDof;
    const double* bcVal;
    const double* bcCount;
    int nBC;
    int accumulate;  // != 0: add to K / R (second and later element blocks of a mixed mesh)
    double* K;
    double* R;
};

__device__ __forceinline__ int bc_find(const int* __restrict__ bcDof, int n, int dof) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (bcDof[mid] < dof) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int NN, int NQ, bool NL, bool WK, bool WR>
__global__ void __launch_bounds__(256) k_assemble_patch(const __grid_constant__ Tables<NN, NQ> tb, const DMat D,
                                                        const PatchArgs A) {
    constexpr int NBLK = NN + NN * (NN - 1) / 2;
    constexpr int BLKBITS = (NBLK <= 2) ? 1 : (NBLK <= 4) ? 2 : (NBLK <= 8) ? 3 : (NBLK <= 16) ? 4 : (NBLK <= 32) ? 5 : (NBLK <= 64) ? 6 : 7;
    extern __shared__ double sm[];
    const int p = blockIdx.x;
    const int pe0 = A.pePtr[p];
    const int ne = A.pePtr[p + 1] - pe0;
    const int stride = ne | 1;  // odd stride: the phase-2 gathers of one element's values spread over the banks
    double* Ksm = sm;                                         // [(blk*4 + c) * stride + t]
    double* Rsm = sm + (WK ? (size_t)NBLK * 4 * stride : 0);  // [(a*2 + s) * stride + t]

    // ---- phase 1: element matrices into shared memory ----------------------------------------------------
    for (int t = threadIdx.x; t < ne; t += blockDim.x) {
        const int* cn = A.peConn + (size_t)NN * pe0 + t;
        double x[NN], y[NN], ux[NN], uy[NN];
#pragma unroll
        for (int a = 0; a < NN; ++a) {
            const int node = __ldg(cn + a * ne);
            const double2 c = __ldg(A.coords + node);
            x[a] = c.x;
            y[a] = c.y;
            if (NL || WR) {
                const double2 q = __ldg(A.states + node);
                ux[a] = q.x;
                uy[a] = q.y;
            } else {
                ux[a] = uy[a] = 0.0;
            }
        }
        const double theta = __ldg(A.thick + A.elemOffset + __ldg(A.peElem + pe0 + t));
        double Kd[NN][3], Ko[NN * (NN - 1) / 2 + 1][4], Re[NN][2];
        element_sym<NN, NQ, NL, WK, WR>(tb, D, x, y, ux, uy, theta, Kd, Ko, Re);
        if (WK) {
#pragma unroll
            for (int a = 0; a < NN; ++a) {
                Ksm[(a * 4 + 0) * stride + t] = Kd[a][0];
                Ksm[(a * 4 + 1) * stride + t] = Kd[a][1];
                Ksm[(a * 4 + 2) * stride + t] = Kd[a][1];
                Ksm[(a * 4 + 3) * stride + t] = Kd[a][2];
            }
#pragma unroll
            for (int k = 0; k < NN * (NN - 1) / 2; ++k) {
#pragma unroll
                for (int c = 0; c < 4; ++c) Ksm[((NN + k) * 4 + c) * stride + t] = Ko[k][c];
            }
        }
        if (WR) {
#pragma unroll
            for (int a = 0; a < NN; ++a) {
                Rsm[(a * 2 + 0) * stride + t] = Re[a][0];
                Rsm[(a * 2 + 1) * stride + t] = Re[a][1];
            }
        }
    }
    __syncthreads();

    // ---- phase 2: gather into the CSR rows of the owned nodes ---------------------------------------------
    const long long pos0 = (long long)p * A.PN;
    const long long pos1 = (pos0 + A.PN < A.numNodes) ? pos0 + A.PN : A.numNodes;
    const int nb0 = A.nbStart[pos0];
    const int nbn = A.nbStart[pos1] - nb0;
    const unsigned short* sp = A.srcPtr + nb0 + p;
    const unsigned short* sr = A.src + A.srcBase[p];
    for (int q = threadIdx.x; q < nbn; q += blockDim.x) {
        const long long pos = pos0 + A.nbRow[nb0 + q];
        const int i = __ldg(A.order + pos);
        const int off = __ldg(A.nrowptr + i);
        const int deg = __ldg(A.nrowptr + i + 1) - off;
        const int k = nb0 + q - __ldg(A.nbStart + pos);
        const int s0 = sp[q], s1 = sp[q + 1];
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0, rx = 0.0, ry = 0.0;
        bool diag = false;
        for (int s = s0; s < s1; ++s) {
            const unsigned code = sr[s];
            const unsigned tr = code & 1u;
            const unsigned blk = (code >> 1) & ((1u << BLKBITS) - 1u);
            const unsigned el = code >> (BLKBITS + 1);
            if (WK) {
                const double* b = Ksm + (size_t)(blk * 4) * stride + el;
                const double b1 = b[stride], b2 = b[2 * stride];
                v0 += b[0];
                v1 += tr ? b2 : b1;
                v2 += tr ? b1 : b2;
                v3 += b[3 * stride];
            }
            if (blk < NN) {
                diag = true;
                if (WR) {
                    rx += Rsm[(blk * 2) * stride + el];
                    ry += Rsm[(blk * 2 + 1) * stride + el];
                }
            }
        }
        const unsigned mask = A.bcMask ? A.bcMask[i] : 0u;
        if (mask) {  // constrained rows: zeros except the occurrence count on the diagonal (AssemblyUtils.py:55-60)
            const bool onDiag = (__ldg(A.ncols + off + k) == i);
            if (mask & 1u) {
                v0 = onDiag ? A.bcCount[bc_find(A.bcDof, A.nBC, 2 * i)] : 0.0;
                v1 = 0.0;
            }
            if (mask & 2u) {
                v2 = 0.0;
                v3 = onDiag ? A.bcCount[bc_find(A.bcDof, A.nBC, 2 * i + 1)] : 0.0;
            }
        }
        if (WK) {
            double2* r0 = reinterpret_cast<double2*>(A.K + 4ll * off + 2 * k);
            double2* r1 = reinterpret_cast<double2*>(A.K + 4ll * off + 2 * deg + 2 * k);
            if (A.accumulate) {
                if (s1 > s0) {
                    const double2 o0 = *r0, o1 = *r1;
                    *r0 = make_double2(o0.x + v0, o0.y + v1);
                    *r1 = make_double2(o1.x + v2, o1.y + v3);
                }
            } else {
                *r0 = make_double2(v0, v1);
                *r1 = make_double2(v2, v3);
            }
        }
        if (WR && diag) {
            double2* rr = reinterpret_cast<double2*>(A.R + 2ll * i);
            if (A.accumulate) {
                const double2 o = *rr;
                *rr = make_double2(o.x + rx, o.y + ry);
            } else {
                if (A.loads) {
                    const double2 f = __ldg(reinterpret_cast<const double2*>(A.loads) + i);
                    rx -= f.x;
                    ry -= f.y;
                }
                if (mask) {  // R[bc] = u - c (AssemblyUtils.py:93)
                    const double2 u = __ldg(A.states + i);
                    if (mask & 1u) rx = u.x - A.bcVal[bc_find(A.bcDof, A.nBC, 2 * i)];
                    if (mask & 2u) ry = u.y - A.bcVal[bc_find(A.bcDof, A.nBC, 2 * i + 1)];
                }
                *rr = make_double2(rx, ry);
            }
        }
    }
    // nodes that belong to no element have no node-block: their residual is just -load (Problem.py:641,653)
    if (WR && !A.accumulate) {
        for (long long pos = pos0 + threadIdx.x; pos < pos1; pos += blockDim.x) {
            const int i = __ldg(A.order + pos);
            if (__ldg(A.nrowptr + i + 1) == __ldg(A.nrowptr + i)) {
                double2 f = make_double2(0.0, 0.0);
                if (A.loads) f = __ldg(reinterpret_cast<const double2*>(A.loads) + i);
                *reinterpret_cast<double2*>(A.R + 2ll * i) = make_double2(-f.x, -f.y);
            }
        }
    }
}

template <int NN, int NQ>
static size_t patch_smem_bytes(int maxElems, bool wk, bool wr) {
    constexpr int NBLK = NN + NN * (NN - 1) / 2;
    const size_t stride = (size_t)(maxElems | 1);
    return ((wk ? (size_t)NBLK * 4 : 0) + (wr ? (size_t)NN * 2 : 0)) * stride * sizeof(double);
}

template <int NN, int NQ, bool NL, bool WK, bool WR>
static int launch_patch(const BlockData& b, const DMat& D, const PatchArgs& A, int64_t numPatches, cudaStream_t st) {
    const Tables<NN, NQ> tb = make_tables<NN, NQ>(b);
    const size_t smem = patch_smem_bytes<NN, NQ>(b.prog.maxElems, WK, WR);
    static size_t configured = 0;
    if (smem > configured) {
        CUDA_TRY(cudaFuncSetAttribute(k_assemble_patch<NN, NQ, NL, WK, WR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    int threads = ((b.prog.maxElems + 31) / 32) * 32;
    threads = std::min(256, std::max(64, threads));
    k_assemble_patch<NN, NQ, NL, WK, WR><<<(unsigned)numPatches, threads, smem, st>>>(tb, D, A);
    return 0;
}

template <int NN, int NQ>
static int launch_patch_variant(const BlockData& b, const DMat& D, const PatchArgs& A, int64_t numPatches, bool nl, bool wk,
                                bool wr, cudaStream_t st) {
    if constexpr (NN <= 4) {
        if (!nl) {
            if (wk && wr) return launch_patch<NN, NQ, false, true, true>(b, D, A, numPatches, st);
            if (wk) return launch_patch<NN, NQ, false, true, false>(b, D, A, numPatches, st);
            return launch_patch<NN, NQ, false, false, true>(b, D, A, numPatches, st);
        }
        if (wk && wr) return launch_patch<NN, NQ, true, true, true>(b, D, A, numPatches, st);
        if (wk) return launch_patch<NN, NQ, true, true, false>(b, D, A, numPatches, st);
        return launch_patch<NN, NQ, true, false, true>(b, D, A, numPatches, st);
    } else {
        return fail("no patch kernel for %d-node elements", NN);
    }
}

static constexpr size_t PATCH_SMEM_LIMIT = 200 * 1024;

static bool patch_path_available(const femb_plan* p, bool wk, bool wr) {
    if (!p->usePatch || !p->patchesBuilt) return false;
    for (auto& b : p->blocks) {
        if (!b.prog.built) return false;
        size_t smem = 0;
        if (b.nn == 3) smem = patch_smem_bytes<3, 1>(b.prog.maxElems, wk, wr);
        else if (b.nn == 4) smem = patch_smem_bytes<4, 4>(b.prog.maxElems, wk, wr);
        else return false;
        if (smem > PATCH_SMEM_LIMIT) return false;
    }
    return true;
}

template <int NN, int NQ, bool ROW, bool NL, bool WK, bool WR>
static void launch_assemble(const BlockData& b, const DMat& D, const AsmArgs& A, cudaStream_t st) {
    const Tables<NN, NQ> tb = make_tables<NN, NQ>(b);
    if constexpr (ROW) {
        dim3 grid(grid_for(A.E, 128), NN);
        k_assemble_row<NN, NQ, NL, WK, WR><<<grid, 128, 0, st>>>(tb, D, A);
    } else {
        k_assemble_sym<NN, NQ, NL, WK, WR><<<grid_for(A.E, 128), 128, 0, st>>>(tb, D, A);
    }
}

template <int NN, int NQ, bool ROW>
static void launch_assemble_variant(const BlockData& b, const DMat& D, const AsmArgs& A, bool nl, bool wk, bool wr,
                                    cudaStream_t st) {
    if (!nl) {
        if (wk && wr) launch_assemble<NN, NQ, ROW, false, true, true>(b, D, A, st);
        else if (wk) launch_assemble<NN, NQ, ROW, false, true, false>(b, D, A, st);
        else launch_assemble<NN, NQ, ROW, false, false, true>(b, D, A, st);
    } else {
        if (wk && wr) launch_assemble<NN, NQ, ROW, true, true, true>(b, D, A, st);
        else if (wk) launch_assemble<NN, NQ, ROW, true, true, false>(b, D, A, st);
        else launch_assemble<NN, NQ, ROW, true, false, true>(b, D, A, st);
    }
}

// =============================================================================================
// C ABI: assembly
// =============================================================================================
extern "C" int femb_assemble_dev(femb_plan* p, const double* d_coords, const double* d_states, const double* d_thick,
                                 const femb_material* mat, int32_t applyBCs, const double* d_loads, int32_t want,
                                 double* d_K, double* d_R) {
    TRY(need_final(p, "femb_assemble_dev"));
    if (!mat) return fail("femb_assemble_dev: material is NULL");
    if (mat->model != FEMB_PLANE_STRESS && mat->model != FEMB_PLANE_STRAIN)
        return fail("femb_assemble_dev: unsupported constitutive model id %d", mat->model);
    const bool wk = (want & FEMB_WANT_K) != 0, wr = (want & FEMB_WANT_R) != 0;
    if (!wk && !wr) return fail("femb_assemble_dev: nothing requested");
    if ((wk && !d_K) || (wr && !d_R) || !d_coords || !d_states || !d_thick) return fail("femb_assemble_dev: NULL buffer");
    TRY(use_device(p));
    cudaStream_t st = p->stream;
    const DMat D = make_dmat(*mat);
    const int64_t numDOF = femb_plan_num_dof(p), nnz = femb_plan_nnz(p);
    int64_t launches = 0;
    const bool patchPath = patch_path_available(p, wk, wr);
    const bool multi = p->blocks.size() > 1;
    const bool bcs = applyBCs && p->nBC > 0;
    if (!patchPath || multi) {  // scatter path, or several element types summing into the same rows
        if (wk) CUDA_TRY(cudaMemsetAsync(d_K, 0, nnz * sizeof(double), st));
        if (wr) {
            k_init_residual<<<grid_for(numDOF, 256), 256, 0, st>>>(d_R, d_loads, numDOF);
            launches++;
        }
    }
    std::pair<cudaEvent_t, cudaEvent_t>* ev = nullptr;
    if (p->profile) {
        if (p->profUsed == p->profEvents.size()) {
            std::pair<cudaEvent_t, cudaEvent_t> e;
            CUDA_TRY(cudaEventCreate(&e.first));
            CUDA_TRY(cudaEventCreate(&e.second));
            p->profEvents.push_back(e);
        }
        ev = &p->profEvents[p->profUsed++];
        CUDA_TRY(cudaEventRecord(ev->first, st));
    }
    for (auto& b : p->blocks) {
        if (patchPath) {
            PatchArgs A;
            A.numNodes = p->numNodes;
            A.PN = p->patchNodes;
            A.order = p->d_order;
            A.nbStart = p->d_nbStart;
            A.nbRow = p->d_nbRow;
            A.nrowptr = p->d_nrowptr;
            A.ncols = p->d_ncols;
            A.pePtr = b.prog.d_pePtr;
            A.peElem = b.prog.d_peElem;
            A.peConn = b.prog.d_peConn;
            A.srcBase = b.prog.d_srcBase;
            A.srcPtr = b.prog.d_srcPtr;
            A.src = b.prog.d_src;
            A.elemOffset = b.elemOffset;
            A.coords = (const double2*)d_coords;
            A.states = (const double2*)d_states;
            A.thick = d_thick;
            A.loads = d_loads;
            A.bcMask = (bcs && !multi) ? p->d_bcMask : nullptr;
            A.bcDof = p->d_bcDof;
            A.bcVal = p->d_bcVal;
            A.bcCount = p->d_bcCount;
            A.nBC = (int)p->nBC;
            A.accumulate = multi ? 1 : 0;
            A.K = d_K;
            A.R = d_R;
            FEMB_DISPATCH_FAMILY(b.nn, b.nq, TRY((launch_patch_variant<NN, NQ>(b, D, A, p->numPatches, mat->nonlinear != 0, wk, wr, st))));
        } else {
            AsmArgs A;
            A.E = b.numElems;
            A.elemOffset = b.elemOffset;
            A.conn = b.d_conn;
            A.slot = b.d_slot;
            A.nrowptr = p->d_nrowptr;
            A.coords = (const double2*)d_coords;
            A.states = (const double2*)d_states;
            A.thick = d_thick;
            A.K = d_K;
            A.R = d_R;
            FEMB_DISPATCH_FAMILY(b.nn, b.nq, (launch_assemble_variant<NN, NQ, ROW>(b, D, A, mat->nonlinear != 0, wk, wr, st)));
        }
        launches++;
    }
    if (ev) CUDA_TRY(cudaEventRecord(ev->second, st));
    if (bcs && (!patchPath || multi)) {
        k_apply_bcs<<<grid_for(p->nBC * 32, 128), 128, 0, st>>>(p->nBC, p->d_bcDof, p->d_bcVal, p->d_bcCount, p->d_nrowptr,
                                                               p->d_ncols, d_states, wk ? d_K : nullptr,
                                                               wr ? d_R : nullptr);
        launches++;
    }
    p->launches = launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int femb_plan_profile(femb_plan* p, int32_t enable) {
    if (!p) return fail("femb_plan_profile: plan is NULL");
    p->profile = enable != 0;
    p->profUsed = 0;
    return 0;
}

extern "C" int femb_plan_profile_read(femb_plan* p, double* sumMs, int64_t* count) {
    if (!p || !sumMs || !count) return fail("femb_plan_profile_read: NULL argument");
    TRY(use_device(p));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    double sum = 0.0;
    for (size_t i = 0; i < p->profUsed; ++i) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, p->profEvents[i].first, p->profEvents[i].second));
        sum += ms;
    }
    *sumMs = sum;
    *count = (int64_t)p->profUsed;
    p->profUsed = 0;
    return 0;
}

static int stage(double** buf, const double* host, size_t n, cudaStream_t st) {
    TRY(dalloc(buf, n));
    CUDA_TRY(cudaMemcpyAsync(*buf, host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    return 0;
}

extern "C" int femb_assemble(femb_plan* p, const double* coords, const double* states, const double* thickness,
                             const femb_material* mat, const int64_t* bcDof, const double* bcVal, int64_t nBC,
                             int32_t applyBCs, const double* loads, int32_t want, double* Kdata, double* R) {
    TRY(need_final(p, "femb_assemble"));
    if (!coords || !states || !thickness) return fail("femb_assemble: NULL input");
    const bool wk = (want & FEMB_WANT_K) != 0, wr = (want & FEMB_WANT_R) != 0;
    if ((wk && !Kdata) || (wr && !R)) return fail("femb_assemble: NULL output");
    TRY(use_device(p));
    const int64_t numDOF = femb_plan_num_dof(p), nnz = femb_plan_nnz(p);
    if (applyBCs) TRY(femb_plan_set_bcs(p, bcDof, bcVal, nBC));
    TRY(stage(&p->s_coords, coords, (size_t)p->numNodes * 2, p->stream));
    TRY(stage(&p->s_states, states, (size_t)numDOF, p->stream));
    TRY(stage(&p->s_thick, thickness, (size_t)p->numElems, p->stream));
    if (loads && wr) TRY(stage(&p->s_loads, loads, (size_t)numDOF, p->stream));
    if (wk) TRY(dalloc(&p->s_K, (size_t)nnz));
    if (wr) TRY(dalloc(&p->s_R, (size_t)numDOF));
    TRY(femb_assemble_dev(p, p->s_coords, p->s_states, p->s_thick, mat, applyBCs, (loads && wr) ? p->s_loads : nullptr,
                          want, p->s_K, p->s_R));
    if (wk) CUDA_TRY(cudaMemcpyAsync(Kdata, p->s_K, nnz * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    if (wr) CUDA_TRY(cudaMemcpyAsync(R, p->s_R, numDOF * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}
