// Fused K + R assembly kernels and their C ABI.
#include "assemble_fan.cuh"
#include "assemble_rows.cuh"
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
// row-owner path (assemble_rows.cuh): one thread per node row, redundant geometry, no atomics / barriers
// =============================================================================================
bool rows_path_available(const femb_plan* p) {
    // every block needs corner records (<= 12 nodes per element: 16-node quads stay on the scatter kernels, their strip
    // alone is 64 doubles per thread)
    if (p->assemblyPath != 2 || !p->rowsBuilt || p->maxDeg > rows::MAXD) return false;
    for (auto& b : p->blocks) {
        if (b.nn <= 4 && !b.cyclic) return false;  // rotated corner records need tables invariant under cyclic renumbering
        if (!default_rule(b.nn, b.nq)) return false;  // non-default quadrature rules run on the scatter kernels
    }
    const size_t smem = (size_t)2 * p->maxDeg * rows::BLOCK * (sizeof(double2) + 1) + rows::BLOCK * sizeof(double2);
    return smem <= 200 * 1024;
}

// =============================================================================================
// write-once fan kernel (assemble_fan.cuh): single block of 3- / 4-node elements with the standard tables, small strain
// =============================================================================================
bool fan_path_available(const femb_plan* p) {
    return rows_path_available(p) && p->useFan && p->blocks.size() == 1 && p->blocks[0].fanOK && p->blocks[0].d_fanHdr && p->blocks[0].closedForm;
}

template <int NN, bool WK, bool WR, bool CMP>
static int launch_fan(const femb_plan* p, const DMat& D, const fan::Args& A, cudaStream_t st) {
    const size_t smem = (size_t)2 * p->maxDeg * fan::BLOCK * sizeof(double2);  // row staging
    static size_t configured[64] = {};
    size_t& conf = configured[p->device & 63];
    if (smem > conf) {
        CUDA_TRY(cudaFuncSetAttribute(fan::k_assemble_fan<NN, WK, WR, CMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conf = smem;
    }
    {   // Shared-memory carve-out: 85 % of what the register-limited number of warps would need -- L1 capacity for the node
        // data the corners share is worth more than the last resident warps (sweep on one box: config 2 29.6 us at
        // 60-72 %, 30.9 at 80-100 %, 31.0 at 55 %; config 4 219 us at 68-75 %, 237 at 80-100 %, 227 at 60 %).
        static int carved[64] = {};
        int& cv = carved[p->device & 63];
        const int want = p->fanCarveout >= 0 ? p->fanCarveout
                                             : std::min(100, (int)((fan::resident_warps<NN>() * (smem + 1024) * 85 + 228 * 1024 - 1) / (228 * 1024)));
        if (cv != want + 1) {
            CUDA_TRY(cudaFuncSetAttribute(fan::k_assemble_fan<NN, WK, WR, CMP>, cudaFuncAttributePreferredSharedMemoryCarveout, want));
            cv = want + 1;
        }
    }
    fan::k_assemble_fan<NN, WK, WR, CMP><<<grid_for(p->numNodes, fan::BLOCK), fan::BLOCK, smem, st>>>(D, A);
    return 0;
}

template <int NN>
static int launch_fan_variant(const femb_plan* p, const DMat& D, const fan::Args& A, bool wk, bool wr, cudaStream_t st) {
    {
        if (p->blocks[0].fanCompact) {
            if (wk && wr) return launch_fan<NN, true, true, true>(p, D, A, st);
            if (wk) return launch_fan<NN, true, false, true>(p, D, A, st);
            return launch_fan<NN, false, true, true>(p, D, A, st);
        }
    }
    if (wk && wr) return launch_fan<NN, true, true, false>(p, D, A, st);
    if (wk) return launch_fan<NN, true, false, false>(p, D, A, st);
    return launch_fan<NN, false, true, false>(p, D, A, st);
}

template <int NN, int NQ, bool NL, bool WK, bool WR, bool MB, bool PERM, bool CF>
static int launch_rows_cf(const femb_plan* p, const BlockData& b, const DMat& D, const rows::Args& A, cudaStream_t st) {
    const Tables<NN, NQ> tb = make_tables<NN, NQ>(b);
    // accumulator rows + residual pairs + write-out owner table
    const size_t smem = (size_t)2 * p->maxDeg * rows::BLOCK * (sizeof(double2) + 1) + rows::BLOCK * sizeof(double2);
    // per device (a process may hold plans on several GPUs): the opt-in above 48 KB is remembered per device ordinal
    static size_t configured[64] = {};
    size_t& conf = configured[p->device & 63];
    if (smem > conf) {
        CUDA_TRY(cudaFuncSetAttribute(rows::k_assemble_rows<NN, NQ, NL, WK, WR, MB, PERM, CF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conf = smem;
    }
    const long long slots = PERM ? p->rowSlots : p->numNodes;
    rows::k_assemble_rows<NN, NQ, NL, WK, WR, MB, PERM, CF><<<grid_for(slots, rows::BLOCK), rows::BLOCK, smem, st>>>(tb, D, A);
    return 0;
}

// 4-node blocks whose tables are the standard bilinear quad with the 2x2 Gauss rule (plan.cu: standard_q4) take the
// closed-form kernel for the small-strain model
template <int NN, int NQ, bool NL, bool WK, bool WR, bool MB, bool PERM>
static int launch_rows_mb(const femb_plan* p, const BlockData& b, const DMat& D, const rows::Args& A, cudaStream_t st) {
    if constexpr (NN == 4 && NQ == 4 && !NL && FEMB_ROT4) {
        if (b.closedForm) return launch_rows_cf<NN, NQ, NL, WK, WR, MB, PERM, true>(p, b, D, A, st);
    }
    return launch_rows_cf<NN, NQ, NL, WK, WR, MB, PERM, false>(p, b, D, A, st);
}

template <int NN, int NQ, bool NL, bool WK, bool WR>
static int launch_rows(const femb_plan* p, const BlockData& b, const DMat& D, const rows::Args& A, cudaStream_t st) {
    if (p->rowsBalanced) {
        if (p->blocks.size() > 1) return launch_rows_mb<NN, NQ, NL, WK, WR, true, true>(p, b, D, A, st);
        return launch_rows_mb<NN, NQ, NL, WK, WR, false, true>(p, b, D, A, st);
    }
    if (p->blocks.size() > 1) return launch_rows_mb<NN, NQ, NL, WK, WR, true, false>(p, b, D, A, st);
    return launch_rows_mb<NN, NQ, NL, WK, WR, false, false>(p, b, D, A, st);
}

template <int NN, int NQ>
static int launch_rows_variant(const femb_plan* p, const BlockData& b, const DMat& D, const rows::Args& A, bool nl, bool wk,
                               bool wr, cudaStream_t st) {
    if constexpr (NN <= 12) {
        if (!nl) {
            if (wk && wr) return launch_rows<NN, NQ, false, true, true>(p, b, D, A, st);
            if (wk) return launch_rows<NN, NQ, false, true, false>(p, b, D, A, st);
            return launch_rows<NN, NQ, false, false, true>(p, b, D, A, st);
        }
        if (wk && wr) return launch_rows<NN, NQ, true, true, true>(p, b, D, A, st);
        if (wk) return launch_rows<NN, NQ, true, true, false>(p, b, D, A, st);
        return launch_rows<NN, NQ, true, false, true>(p, b, D, A, st);
    } else {
        return fail("no row-owner kernel for %d-node elements", NN);
    }
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
    const bool wk = (want & FEMB_WANT_K) != 0, wr = (want & FEMB_WANT_R) != 0;
    if (!wk && !wr) return fail("femb_assemble_dev: nothing requested");
    if ((wk && !d_K) || (wr && !d_R) || !d_coords || !d_states || !d_thick) return fail("femb_assemble_dev: NULL buffer");
    TRY(use_device(p));
    if (p->numDim != 2)  // 1-D / 3-D families: generic.cu
        return generic_assemble_dev(p, d_coords, d_states, d_thick, mat, applyBCs && p->nBC > 0, d_loads, wk, wr, d_K, d_R);
    if (mat->model != FEMB_PLANE_STRESS && mat->model != FEMB_PLANE_STRAIN)
        return fail("femb_assemble_dev: unsupported constitutive model id %d for a 2-D plan", mat->model);
    cudaStream_t st = p->stream;
    const DMat D = make_dmat(*mat);
    const int64_t numDOF = femb_plan_num_dof(p), nnz = femb_plan_nnz(p);
    int64_t launches = 0;
    const bool rowsPath = rows_path_available(p);
    const bool bcs = applyBCs && p->nBC > 0;
    if (!rowsPath) TRY(ensure_slot_maps(p));
    if (!rowsPath) {  // scatter path: the element blocks are added into zeroed K / -loads
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
    const bool fanPath = rowsPath && !mat->nonlinear && fan_path_available(p);
    for (auto& b : p->blocks) {
        if (fanPath) {
            fan::Args A;
            A.numNodes = p->numNodes;
            A.hdr = b.d_fanHdr;
            A.rest = b.d_fanRest;
            A.coords = (const double2*)d_coords;
            A.states = (const double2*)d_states;
            A.thick = d_thick;
            A.loads = (const double2*)d_loads;
            A.useBC = bcs ? 1 : 0;
            A.bcNodeCnt = (const double2*)p->d_bcNodeCnt;
            A.bcNodeVal = (const double2*)p->d_bcNodeVal;
            A.maxDeg = p->maxDeg;
            A.zero = 0;
            {   // L2 look-ahead distance: by default a sixth of the resident wave (SMs x resident warps per SM / 6) -- far enough
                // for the prefetch to land, near enough that the lines are still in L2 when their warp starts (measured:
                // 200..600 warps at configs 2 / 5, 600..900 at config 4; a whole wave ahead loses 2..5 %)
                if (p->numSMs <= 0 && cudaDeviceGetAttribute(&p->numSMs, cudaDevAttrMultiProcessorCount, p->device) != cudaSuccess) p->numSMs = 148;
                const int sms = p->numSMs;
                const long long warps = p->fanPrefetch >= 0 ? p->fanPrefetch : (long long)sms * (b.nn == 3 ? fan::resident_warps<3>() : fan::resident_warps<4>()) / 6;
                const long long pfNodes = std::min<long long>(warps * fan::BLOCK, p->numNodes);
                const double perNode = (double)pfNodes / (double)p->numNodes;
                A.pfNodes = (int)pfNodes;
                A.pfRest = (int)(perNode * (double)b.fanRestTotal);  // in records (16 / 32 bytes, 8 for compact 3-node records)
                A.pfElems = (int)(perNode * (double)p->numElems);
                A.pfNodeLimit = pfNodes ? (int)(p->numNodes - pfNodes) : 0;
                A.pfRestLimit = pfNodes ? (int)(b.fanRestTotal - A.pfRest) : 0;
                A.pfElemLimit = pfNodes ? (int)(p->numElems - A.pfElems) : 0;
            }
            A.K = d_K;
            A.R = d_R;
            if (b.nn == 3) TRY((launch_fan_variant<3>(p, D, A, wk, wr, st)));
            else TRY((launch_fan_variant<4>(p, D, A, wk, wr, st)));
        } else if (rowsPath) {
            rows::Args A;
            A.numNodes = p->numNodes;
            A.slotRec = p->d_slotRec;
            A.nrowptr = p->d_nrowptr;
            A.ncols = p->d_ncols;
            const size_t k = &b - &p->blocks[0];
            A.cBegin = p->d_blockAdj + k * (size_t)p->numNodes;
            A.cEnd = p->d_blockAdj + (k + 1) * (size_t)p->numNodes;
            A.elemOffset = (int)b.elemOffset;
            A.accumulate = k > 0;
            A.finalise = k + 1 == p->blocks.size();
            A.corners = b.d_corners;
            A.cornerElem = b.d_cornerElem;
            A.zero = 0;
            A.stats = p->d_rowStats;
            A.connE = b.d_connE;
            A.coords = (const double2*)d_coords;
            A.states = (const double2*)d_states;
            A.thick = d_thick;
            A.loads = (const double2*)d_loads;
            A.bcInfo = bcs ? p->d_bcInfo : nullptr;
            A.bcNodeCnt = (const double2*)p->d_bcNodeCnt;
            A.bcNodeVal = (const double2*)p->d_bcNodeVal;
            A.maxDeg = p->maxDeg;
            A.K = d_K;
            A.R = d_R;
            FEMB_DISPATCH_FAMILY(b.nn, b.nq, TRY((launch_rows_variant<NN, NQ>(p, b, D, A, mat->nonlinear != 0, wk, wr, st))));
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
            FEMB_DISPATCH_ANY(b.nn, b.nq, (launch_assemble_variant<NN, NQ, ROW>(b, D, A, mat->nonlinear != 0, wk, wr, st)));
        }
        launches++;
    }
    if (ev) CUDA_TRY(cudaEventRecord(ev->second, st));
    if (bcs && !rowsPath) {
        k_apply_bcs<<<grid_for(p->nBC * 32, 128), 128, 0, st>>>(p->nBC, p->d_bcDof, p->d_bcVal, p->d_bcCount, p->d_nrowptr,
                                                               p->d_ncols, d_states, wk ? d_K : nullptr,
                                                               wr ? d_R : nullptr);
        launches++;
    }
    p->launches = launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// measurement builds only (ROWS_TRACE, tools/rows_trace.py): warp lifetime statistics of the row-owner kernel
extern "C" int femb_plan_row_stats(femb_plan* p, int32_t enable, int64_t* out, int64_t capacity) {
    if (!p) return fail("femb_plan_row_stats: plan is NULL");
    TRY(use_device(p));
    const size_t n = 16 * 256;
    if (out) {
        if (!p->d_rowStats) return fail("femb_plan_row_stats: not enabled");
        CUDA_TRY(cudaStreamSynchronize(p->stream));
        CUDA_TRY(cudaMemcpy(out, p->d_rowStats, std::min<size_t>(n, (size_t)capacity) * sizeof(long long), cudaMemcpyDeviceToHost));
        return 0;
    }
    if (enable) {
        TRY(dalloc(&p->d_rowStats, n));
        CUDA_TRY(cudaMemset(p->d_rowStats, 0, n * sizeof(long long)));
        const long long big = 0x7fffffffffffffffll;  // slots 3 and 5 hold minima
        CUDA_TRY(cudaMemcpy(p->d_rowStats + 3, &big, sizeof(big), cudaMemcpyHostToDevice));
        CUDA_TRY(cudaMemcpy(p->d_rowStats + 5, &big, sizeof(big), cudaMemcpyHostToDevice));
    } else {
        dfree(p->d_rowStats);
    }
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

// Device-side model: node coordinates and per-element thickness are constants of a Newton or
// optimisation loop (FEMpyModel.setCoordinates / setDesignVariables change them, Model.py:240-276); uploaded once here,
// the host-pointer entries then take NULL for them and move only the states (and loads) per call.
extern "C" int femb_plan_set_geometry(femb_plan* p, const double* coords, const double* thickness) {
    TRY(need_final(p, "femb_plan_set_geometry"));
    TRY(use_device(p));
    if (coords) {
        TRY(stage_geometry(p, coords, nullptr, false, "femb_plan_set_geometry"));
        p->coordsResident = true;
    } else {
        p->coordsResident = false;
    }
    if (thickness) {
        TRY(stage(&p->s_thick, thickness, (size_t)p->numElems, p->stream));
        p->thickResident = true;
    } else {
        p->thickResident = false;
    }
    CUDA_TRY(cudaStreamSynchronize(p->stream));  // the caller's arrays may change after the call
    return 0;
}

extern "C" int femb_assemble(femb_plan* p, const double* coords, const double* states, const double* thickness,
                             const femb_material* mat, const int64_t* bcDof, const double* bcVal, int64_t nBC,
                             int32_t applyBCs, const double* loads, int32_t want, double* Kdata, double* R) {
    TRY(need_final(p, "femb_assemble"));
    if (!states) return fail("femb_assemble: NULL input");
    const bool wk = (want & FEMB_WANT_K) != 0, wr = (want & FEMB_WANT_R) != 0;
    if ((wk && !Kdata) || (wr && !R)) return fail("femb_assemble: NULL output");
    TRY(use_device(p));
    const int64_t numDOF = femb_plan_num_dof(p), nnz = femb_plan_nnz(p);
    if (applyBCs) TRY(femb_plan_set_bcs(p, bcDof, bcVal, nBC));
    TRY(stage_geometry(p, coords, thickness, true, "femb_assemble"));
    TRY(stage(&p->s_states, states, (size_t)numDOF, p->stream));
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
