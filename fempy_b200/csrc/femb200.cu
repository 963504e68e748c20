// femb200: B200-native FEM assembly -- plan (pattern + scatter map), fused K/R assembly, BCs,
// element functions, and the C ABI declared in include/femb200.h.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/femb200.h"
#include "elem_math.cuh"

using namespace femb;

// =============================================================================================
// error handling
// =============================================================================================
static thread_local std::string g_err;

static int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define TRY(expr)            \
    do {                     \
        int _r = (expr);     \
        if (_r) return _r;   \
    } while (0)

extern "C" const char* femb_last_error(void) { return g_err.c_str(); }
extern "C" const char* femb_version(void) { return "femb200 0.1 (sm_100a)"; }
extern "C" int femb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// =============================================================================================
// plan
// =============================================================================================
static constexpr int MAX_BLOCKS = 8;

struct BlockData {
    int nn = 0, nq = 0;
    int64_t numElems = 0;
    int64_t elemOffset = 0;   // first global element id
    int64_t pointOffset = 0;  // first global Gauss point (for FEMB_REDUCE_NONE output)
    int64_t keOffset = 0, reOffset = 0;
    std::vector<double> N, dN, w;
    int* d_conn = nullptr;  // node-major int32: conn[a * E + e]
    int* d_slot = nullptr;  // [(a*nn + b) * E + e] -> position of (node_a, node_b) in the node-level CSR
};

struct BlockTable {  // kernel-side view of all blocks (pattern build)
    int numBlocks;
    int nn[MAX_BLOCKS];
    long long numElems[MAX_BLOCKS];
    long long elemOffset[MAX_BLOCKS];
    const int* conn[MAX_BLOCKS];
};

struct femb_plan {
    int device = 0;
    int64_t numNodes = 0;
    int numDim = 2, numStates = 2;
    std::vector<BlockData> blocks;
    int64_t numElems = 0, numPoints = 0, keTotal = 0, reTotal = 0;
    bool finalized = false;
    int* d_nrowptr = nullptr;  // node-level CSR row pointers (numNodes + 1)
    int* d_ncols = nullptr;    // node-level CSR column nodes, sorted within a row
    int64_t nnzNode = 0;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    float buildMs = 0.f;
    int64_t launches = 0;
    // boundary conditions, de-duplicated on the host with the reference's semantics
    int64_t nBC = 0;
    int* d_bcDof = nullptr;
    double* d_bcVal = nullptr;    // last value given for the DOF
    double* d_bcCount = nullptr;  // number of occurrences (diagonal value)
    // scratch for the host-pointer API
    double *s_coords = nullptr, *s_states = nullptr, *s_thick = nullptr, *s_loads = nullptr, *s_K = nullptr,
           *s_R = nullptr, *s_out = nullptr;
    size_t s_outCap = 0;
    int* d_flag = nullptr;
    // optional CUDA-event timing of the dominant assembly kernel(s) (bench.py roofline leg)
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> profEvents;
    size_t profUsed = 0;
};

template <typename T>
static int dalloc(T** p, size_t n) {
    if (*p) return 0;
    CUDA_TRY(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    return 0;
}
template <typename T>
static void dfree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

// =============================================================================================
// exclusive scan (int32), hierarchical: 1024 items per block
// =============================================================================================
static constexpr int SCAN_THREADS = 256;
static constexpr int SCAN_ITEMS = 4;
static constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void k_scan_tiles(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ tileSums, int64_t n) {
    __shared__ int warpSums[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        sum += v[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warpSums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int ws = (lane < SCAN_THREADS / 32) ? warpSums[lane] : 0;
#pragma unroll
        for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, ws, d);
            if (lane >= d) ws += t;
        }
        if (lane < SCAN_THREADS / 32) warpSums[lane] = ws;
    }
    __syncthreads();
    int excl = incl - sum + (warp > 0 ? warpSums[warp - 1] : 0);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = excl;
        excl += v[i];
    }
    if (threadIdx.x == SCAN_THREADS - 1) tileSums[blockIdx.x] = excl;
}

__global__ void k_scan_add(int* __restrict__ out, const int* __restrict__ tileOffsets, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * SCAN_TILE + threadIdx.x;
    const int off = tileOffsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t j = i + (int64_t)k * SCAN_THREADS;
        if (j < n) out[j] += off;
    }
}

__global__ void k_store_total(const int* __restrict__ in, const int* __restrict__ out, int64_t n, int* total) {
    // total = exclusive[n-1] + in[n-1]
    *total = (n > 0) ? out[n - 1] + in[n - 1] : 0;
}

// out[0..n) = exclusive prefix sums of in[0..n); out[n] = total.  in and out must not alias.
static int exclusive_scan(const int* d_in, int* d_out, int64_t n, cudaStream_t st, int64_t& launches) {
    if (n == 0) {
        CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(int), st));
        return 0;
    }
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    int* d_tileSums = nullptr;
    int* d_tileOffsets = nullptr;
    CUDA_TRY(cudaMalloc(&d_tileSums, (tiles + 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_tileOffsets, (tiles + 1) * sizeof(int)));
    k_scan_tiles<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(d_in, d_out, d_tileSums, n);
    launches++;
    if (tiles > 1) {
        TRY(exclusive_scan(d_tileSums, d_tileOffsets, tiles, st, launches));
        k_scan_add<<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(d_out, d_tileOffsets, n);
        launches++;
    }
    k_store_total<<<1, 1, 0, st>>>(d_in, d_out, n, d_out + n);
    launches++;
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaFree(d_tileSums);
    cudaFree(d_tileOffsets);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// =============================================================================================
// pattern build kernels
// =============================================================================================
// int64 row-major (E, nn) connectivity -> node-major int32, with range check
__global__ void k_convert_conn(const long long* __restrict__ conn64, int* __restrict__ conn32, long long E, int nn,
                               long long numNodes, int* __restrict__ flag) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= E * nn) return;
    const long long e = t / nn;
    const int a = (int)(t - e * nn);
    const long long node = conn64[t];
    if (node < 0 || node >= numNodes) {
        atomicExch(flag, 1);
        conn32[(long long)a * E + e] = 0;
        return;
    }
    conn32[(long long)a * E + e] = (int)node;
}

// adjacency counts: number of elements and number of candidate neighbours per node
__global__ void k_count_adj(const int* __restrict__ conn, long long E, int nn, int* __restrict__ adjCount,
                            int* __restrict__ candCount) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= E * nn) return;
    const int node = conn[t];
    atomicAdd(&adjCount[node], 1);
    atomicAdd(&candCount[node], nn);
}

__global__ void k_fill_adj(const int* __restrict__ conn, long long E, int nn, long long elemOffset,
                           const int* __restrict__ adjPtr, int* __restrict__ cursor, int* __restrict__ adj) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= E * nn) return;
    const long long e = t % E;
    const int node = conn[t];
    const int p = atomicAdd(&cursor[node], 1);
    adj[adjPtr[node] + p] = (int)(elemOffset + e);
}

// one thread per node: sorted, de-duplicated list of neighbour nodes (itself included) written into
// its candidate segment; deg[i] = list length.  The adjacency order is non-deterministic (atomics) but
// the sorted result is not.
__global__ void k_build_rows(BlockTable bt, long long numNodes, const int* __restrict__ adjPtr,
                             const int* __restrict__ adj, const int* __restrict__ candPtr, int* __restrict__ cand,
                             int* __restrict__ deg) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numNodes) return;
    int* seg = cand + candPtr[i];
    int count = 0;
    for (int q = adjPtr[i]; q < adjPtr[i + 1]; ++q) {
        const long long ge = adj[q];
        int b = 0;
        while (b + 1 < bt.numBlocks && ge >= bt.elemOffset[b + 1]) ++b;
        const long long e = ge - bt.elemOffset[b];
        const int nn = bt.nn[b];
        const long long E = bt.numElems[b];
        const int* conn = bt.conn[b];
        for (int a = 0; a < nn; ++a) {
            const int j = conn[(long long)a * E + e];
            // binary search for the insertion point
            int lo = 0, hi = count;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (seg[mid] < j) lo = mid + 1; else hi = mid;
            }
            if (lo < count && seg[lo] == j) continue;
            for (int m = count; m > lo; --m) seg[m] = seg[m - 1];
            seg[lo] = j;
            ++count;
        }
    }
    deg[i] = count;
}

__global__ void k_compact_rows(long long numNodes, const int* __restrict__ candPtr, const int* __restrict__ cand,
                               const int* __restrict__ nrowptr, int* __restrict__ ncols) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numNodes) return;
    const int* seg = cand + candPtr[i];
    const int off = nrowptr[i], d = nrowptr[i + 1] - off;
    for (int k = 0; k < d; ++k) ncols[off + k] = seg[k];
}

// slot[(a*nn+b)*E + e] = position of (node_a, node_b) in the node-level CSR
__global__ void k_build_slots(const int* __restrict__ conn, long long E, int nn, const int* __restrict__ nrowptr,
                              const int* __restrict__ ncols, int* __restrict__ slot) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= E * nn * nn) return;
    const long long e = t % E;
    const int ab = (int)(t / E);
    const int a = ab / nn, b = ab - a * nn;
    const int i = conn[(long long)a * E + e], j = conn[(long long)b * E + e];
    int lo = nrowptr[i], hi = nrowptr[i + 1];
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ncols[mid] < j) lo = mid + 1; else hi = mid;
    }
    slot[t] = lo;
}

// DOF-level CSR (int64) from the node-level CSR.  Row s*i + r holds, for every neighbour k of node i,
// the s columns s*j_k .. s*j_k + s-1; its values start at s*s*off_i + r*s*deg_i.
__global__ void k_dof_pattern(long long numNodes, int s, const int* __restrict__ nrowptr, const int* __restrict__ ncols,
                              long long* __restrict__ indptr, long long* __restrict__ indices) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > numNodes) return;
    if (i == numNodes) {
        indptr[(long long)s * numNodes] = (long long)s * s * nrowptr[numNodes];
        return;
    }
    const long long off = nrowptr[i], d = nrowptr[i + 1] - off;
    for (int r = 0; r < s; ++r) {
        const long long start = (long long)s * s * off + (long long)r * s * d;
        indptr[(long long)s * i + r] = start;
        for (long long k = 0; k < d; ++k) {
            const long long j = ncols[off + k];
            for (int c = 0; c < s; ++c) indices[start + s * k + c] = (long long)s * j + c;
        }
    }
}

// =============================================================================================
// assembly kernels (v0: scatter through the slot map with FP64 RED atomics)
// =============================================================================================
struct AsmArgs {
    long long E;
    long long elemOffset;
    const int* conn;
    const int* slot;
    const int* nrowptr;
    const double2* coords;
    const double2* states;
    const double* thick;  // indexed by global element id
    double* K;
    double* R;
};

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

__global__ void k_check_bcs(long long nBC, const int* __restrict__ bcDof, const int* __restrict__ nrowptr, int* flag) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nBC) return;
    const int i = bcDof[t] >> 1;
    if (nrowptr[i + 1] == nrowptr[i]) atomicExch(flag, 2);
}

// =============================================================================================
// element function kernel
// =============================================================================================
struct FuncArgs {
    long long E;
    const int* conn;
    const double2* coords;
    const double2* states;
    double* out;  // E or E*NQ values
};

template <int NN, int NQ, bool NL>
__global__ void __launch_bounds__(128) k_function(const __grid_constant__ Tables<NN, NQ> tb, const DMat D,
                                                  const FuncArgs F, const int func, const int reduction) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= F.E) return;
    double x[NN], y[NN], ux[NN], uy[NN];
#pragma unroll
    for (int a = 0; a < NN; ++a) {
        const int node = __ldg(F.conn + (long long)a * F.E + e);
        const double2 c = __ldg(F.coords + node);
        const double2 q = __ldg(F.states + node);
        x[a] = c.x;
        y[a] = c.y;
        ux[a] = q.x;
        uy[a] = q.y;
    }
    double vals[NQ], dets[NQ];
    element_function<NN, NQ, NL>(tb, D, func, x, y, ux, uy, vals, dets);
    if (reduction == FEMB_REDUCE_NONE) {
#pragma unroll
        for (int p = 0; p < NQ; ++p) F.out[e * NQ + p] = vals[p];
        return;
    }
    double r = 0.0;
    if (reduction == FEMB_REDUCE_SUM || reduction == FEMB_REDUCE_MEAN) {
#pragma unroll
        for (int p = 0; p < NQ; ++p) r += vals[p];
        if (reduction == FEMB_REDUCE_MEAN) r /= (double)NQ;
    } else if (reduction == FEMB_REDUCE_INTEGRATE) {
#pragma unroll
        for (int p = 0; p < NQ; ++p) r = fma(vals[p] * dets[p], tb.w[p], r);
    } else if (reduction == FEMB_REDUCE_MIN) {
        r = vals[0];
#pragma unroll
        for (int p = 1; p < NQ; ++p) r = fmin(r, vals[p]);
    } else {
        r = vals[0];
#pragma unroll
        for (int p = 1; p < NQ; ++p) r = fmax(r, vals[p]);
    }
    F.out[e] = r;
}

// =============================================================================================
// per-element blocks (debug / finer seam): Ke (E, 2n, 2n), Re (E, n, 2)
// =============================================================================================
template <int NN, int NQ, bool NL>
__global__ void __launch_bounds__(128) k_element_blocks(const __grid_constant__ Tables<NN, NQ> tb, const DMat D,
                                                        const AsmArgs A, double* __restrict__ Ke,
                                                        double* __restrict__ Re) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int a = blockIdx.y;
    if (e >= A.E) return;
    int node[NN];
    double x[NN], y[NN], ux[NN], uy[NN];
    load_element<NN>(A, e, node, x, y, ux, uy);
    const double theta = __ldg(A.thick + A.elemOffset + e);
    double Krow[NN][4], Ra[2];
    element_row<NN, NQ, NL, true, true>(tb, D, a, x, y, ux, uy, theta, Krow, Ra);
    constexpr int ND = 2 * NN;
    if (Ke) {
        double* r0 = Ke + (e * ND + 2 * a) * ND;
#pragma unroll
        for (int b = 0; b < NN; ++b) {
            r0[2 * b] = Krow[b][0];
            r0[2 * b + 1] = Krow[b][1];
            r0[ND + 2 * b] = Krow[b][2];
            r0[ND + 2 * b + 1] = Krow[b][3];
        }
    }
    if (Re) {
        Re[(e * NN + a) * 2] = Ra[0];
        Re[(e * NN + a) * 2 + 1] = Ra[1];
    }
}

// =============================================================================================
// dispatch over element families
// =============================================================================================
template <int NN, int NQ>
static Tables<NN, NQ> make_tables(const BlockData& b) {
    Tables<NN, NQ> t;
    for (int p = 0; p < NQ; ++p) {
        for (int d = 0; d < 2; ++d)
            for (int a = 0; a < NN; ++a) t.dN[p][d][a] = b.dN[(p * 2 + d) * NN + a];
        for (int a = 0; a < NN; ++a) t.N[p][a] = b.N[p * NN + a];
        t.w[p] = b.w[p];
    }
    return t;
}

static DMat make_dmat(const femb_material& m) {
    DMat D;
    if (m.model == FEMB_PLANE_STRESS) {  // IsoStress.py:25-50
        const double f = m.E / (1.0 - m.nu * m.nu);
        D.d00 = f * 1.0;
        D.d01 = f * m.nu;
        D.d11 = f * 1.0;
        D.d22 = f * ((1.0 - m.nu) / 2.0);
    } else {  // IsoStress.py:53-78
        const double f = m.E / ((1.0 + m.nu) * (1.0 - 2.0 * m.nu));
        D.d00 = f * (1.0 - m.nu);
        D.d01 = f * m.nu;
        D.d11 = f * (1.0 - m.nu);
        D.d22 = f * ((1.0 - 2.0 * m.nu) / 2.0);
    }
    D.nu = m.nu;
    D.rho = m.rho;
    D.model = m.model;
    D.nonlinear = m.nonlinear;
    return D;
}

static bool family_supported(int nn, int nq) {
    return (nn == 3 && nq == 1) || (nn == 6 && nq == 3) || (nn == 10 && nq == 3) || (nn == 4 && nq == 4) ||
           (nn == 9 && nq == 9) || (nn == 16 && nq == 16);
}

// FAMILY(nn, nq, ROWSPLIT, body) -- body sees constexpr NN, NQ, ROW
#define FEMB_DISPATCH_FAMILY(nn, nq, BODY)                                                     \
    do {                                                                                       \
        if ((nn) == 3 && (nq) == 1) { constexpr int NN = 3, NQ = 1; constexpr bool ROW = false; (void)ROW; BODY; }        \
        else if ((nn) == 4 && (nq) == 4) { constexpr int NN = 4, NQ = 4; constexpr bool ROW = false; (void)ROW; BODY; }   \
        else if ((nn) == 6 && (nq) == 3) { constexpr int NN = 6, NQ = 3; constexpr bool ROW = true; (void)ROW; BODY; }    \
        else if ((nn) == 9 && (nq) == 9) { constexpr int NN = 9, NQ = 9; constexpr bool ROW = true; (void)ROW; BODY; }    \
        else if ((nn) == 10 && (nq) == 3) { constexpr int NN = 10, NQ = 3; constexpr bool ROW = true; (void)ROW; BODY; }  \
        else if ((nn) == 16 && (nq) == 16) { constexpr int NN = 16, NQ = 16; constexpr bool ROW = true; (void)ROW; BODY; } \
        else return fail("unsupported element family: %d nodes, %d quadrature points", (nn), (nq));            \
    } while (0)

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
// C ABI: plan
// =============================================================================================
static int use_device(const femb_plan* p) {
    CUDA_TRY(cudaSetDevice(p->device));
    return 0;
}

extern "C" int femb_plan_create(int64_t numNodes, int32_t numDim, int32_t numStates, int32_t device, femb_plan** out) {
    if (!out) return fail("femb_plan_create: out is NULL");
    *out = nullptr;
    if (numDim != 2 || numStates != 2)
        return fail("femb_plan_create: only 2-D models with 2 states per node are supported (got numDim=%d, numStates=%d)",
                    numDim, numStates);
    if (numNodes <= 0 || numNodes >= (1ll << 30)) return fail("femb_plan_create: numNodes out of range");
    int ndev = femb_device_count();
    if (ndev == 0) return fail("femb_plan_create: no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("femb_plan_create: device %d out of range (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    femb_plan* p = new femb_plan();
    p->device = device;
    p->numNodes = numNodes;
    p->numDim = numDim;
    p->numStates = numStates;
    if (cudaStreamCreateWithFlags(&p->ownStream, cudaStreamNonBlocking) != cudaSuccess) {
        delete p;
        return fail("femb_plan_create: cudaStreamCreate failed");
    }
    p->stream = p->ownStream;
    if (dalloc(&p->d_flag, 1)) {
        delete p;
        return 1;
    }
    *out = p;
    return 0;
}

extern "C" int femb_plan_add_block(femb_plan* p, int32_t nn, int32_t nq, const double* N, const double* dNdxi,
                                   const double* w, int64_t numElems, const int64_t* conn) {
    if (!p) return fail("femb_plan_add_block: plan is NULL");
    if (p->finalized) return fail("femb_plan_add_block: plan already finalized");
    if ((int)p->blocks.size() >= MAX_BLOCKS) return fail("femb_plan_add_block: at most %d element types", MAX_BLOCKS);
    if (!family_supported(nn, nq))
        return fail("femb_plan_add_block: unsupported element family (%d nodes, %d quadrature points)", nn, nq);
    if (numElems <= 0 || !conn || !N || !dNdxi || !w) return fail("femb_plan_add_block: bad arguments");
    if (numElems * (int64_t)nn * nn >= (1ll << 31)) return fail("femb_plan_add_block: block too large for int32 slot map");
    TRY(use_device(p));
    BlockData b;
    b.nn = nn;
    b.nq = nq;
    b.numElems = numElems;
    b.elemOffset = p->numElems;
    b.pointOffset = p->numPoints;
    b.keOffset = p->keTotal;
    b.reOffset = p->reTotal;
    b.N.assign(N, N + (size_t)nq * nn);
    b.dN.assign(dNdxi, dNdxi + (size_t)nq * 2 * nn);
    b.w.assign(w, w + nq);
    long long* d_conn64 = nullptr;
    const size_t cnt = (size_t)numElems * nn;
    CUDA_TRY(cudaMalloc(&d_conn64, cnt * sizeof(long long)));
    CUDA_TRY(cudaMalloc(&b.d_conn, cnt * sizeof(int)));
    CUDA_TRY(cudaMemcpyAsync(d_conn64, conn, cnt * sizeof(long long), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemsetAsync(p->d_flag, 0, sizeof(int), p->stream));
    k_convert_conn<<<grid_for((int64_t)cnt, 256), 256, 0, p->stream>>>(d_conn64, b.d_conn, numElems, nn, p->numNodes,
                                                                        p->d_flag);
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, p->d_flag, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    cudaFree(d_conn64);
    if (flag) {
        cudaFree(b.d_conn);
        return fail("femb_plan_add_block: connectivity refers to a node outside [0, %lld)", (long long)p->numNodes);
    }
    p->numElems += numElems;
    p->numPoints += numElems * nq;
    p->keTotal += numElems * (int64_t)(2 * nn) * (2 * nn);
    p->reTotal += numElems * (int64_t)nn * 2;
    p->blocks.push_back(std::move(b));
    return 0;
}

extern "C" int femb_plan_finalize(femb_plan* p) {
    if (!p) return fail("femb_plan_finalize: plan is NULL");
    if (p->finalized) return 0;
    if (p->blocks.empty()) return fail("femb_plan_finalize: no element blocks");
    TRY(use_device(p));
    cudaStream_t st = p->stream;
    cudaEvent_t ev0, ev1;
    CUDA_TRY(cudaEventCreate(&ev0));
    CUDA_TRY(cudaEventCreate(&ev1));
    CUDA_TRY(cudaEventRecord(ev0, st));
    const int64_t N = p->numNodes;
    int64_t launches = 0;
    int *d_adjCount = nullptr, *d_candCount = nullptr, *d_adjPtr = nullptr, *d_candPtr = nullptr, *d_cursor = nullptr;
    int *d_adj = nullptr, *d_cand = nullptr, *d_deg = nullptr;
    CUDA_TRY(cudaMalloc(&d_adjCount, (N + 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_candCount, (N + 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_adjPtr, (N + 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_candPtr, (N + 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_cursor, (N + 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_deg, (N + 1) * sizeof(int)));
    CUDA_TRY(cudaMemsetAsync(d_adjCount, 0, (N + 1) * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(d_candCount, 0, (N + 1) * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(d_cursor, 0, (N + 1) * sizeof(int), st));
    int64_t totalAdj = 0, totalCand = 0;
    for (auto& b : p->blocks) {
        k_count_adj<<<grid_for(b.numElems * b.nn, 256), 256, 0, st>>>(b.d_conn, b.numElems, b.nn, d_adjCount, d_candCount);
        launches++;
        totalAdj += b.numElems * b.nn;
        totalCand += b.numElems * (int64_t)b.nn * b.nn;
    }
    if (totalCand >= (1ll << 31)) return fail("femb_plan_finalize: mesh too large for int32 candidate offsets");
    TRY(exclusive_scan(d_adjCount, d_adjPtr, N, st, launches));
    TRY(exclusive_scan(d_candCount, d_candPtr, N, st, launches));
    CUDA_TRY(cudaMalloc(&d_adj, std::max<int64_t>(totalAdj, 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_cand, std::max<int64_t>(totalCand, 1) * sizeof(int)));
    BlockTable bt;
    bt.numBlocks = (int)p->blocks.size();
    for (int i = 0; i < bt.numBlocks; ++i) {
        auto& b = p->blocks[i];
        bt.nn[i] = b.nn;
        bt.numElems[i] = b.numElems;
        bt.elemOffset[i] = b.elemOffset;
        bt.conn[i] = b.d_conn;
        k_fill_adj<<<grid_for(b.numElems * b.nn, 256), 256, 0, st>>>(b.d_conn, b.numElems, b.nn, b.elemOffset, d_adjPtr,
                                                                     d_cursor, d_adj);
        launches++;
    }
    k_build_rows<<<grid_for(N, 128), 128, 0, st>>>(bt, N, d_adjPtr, d_adj, d_candPtr, d_cand, d_deg);
    launches++;
    CUDA_TRY(cudaMalloc(&p->d_nrowptr, (N + 1) * sizeof(int)));
    TRY(exclusive_scan(d_deg, p->d_nrowptr, N, st, launches));
    int nnzNode = 0;
    CUDA_TRY(cudaMemcpyAsync(&nnzNode, p->d_nrowptr + N, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    p->nnzNode = nnzNode;
    CUDA_TRY(cudaMalloc(&p->d_ncols, std::max<int64_t>(nnzNode, 1) * sizeof(int)));
    k_compact_rows<<<grid_for(N, 128), 128, 0, st>>>(N, d_candPtr, d_cand, p->d_nrowptr, p->d_ncols);
    launches++;
    for (auto& b : p->blocks) {
        const int64_t cnt = b.numElems * (int64_t)b.nn * b.nn;
        CUDA_TRY(cudaMalloc(&b.d_slot, cnt * sizeof(int)));
        k_build_slots<<<grid_for(cnt, 256), 256, 0, st>>>(b.d_conn, b.numElems, b.nn, p->d_nrowptr, p->d_ncols, b.d_slot);
        launches++;
    }
    CUDA_TRY(cudaEventRecord(ev1, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    cudaEventElapsedTime(&p->buildMs, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    cudaFree(d_adjCount);
    cudaFree(d_candCount);
    cudaFree(d_adjPtr);
    cudaFree(d_candPtr);
    cudaFree(d_cursor);
    cudaFree(d_adj);
    cudaFree(d_cand);
    cudaFree(d_deg);
    p->launches = launches;
    p->finalized = true;
    return 0;
}

extern "C" void femb_plan_destroy(femb_plan* p) {
    if (!p) return;
    cudaSetDevice(p->device);
    if (p->ownStream) cudaStreamSynchronize(p->ownStream);
    for (auto& b : p->blocks) {
        dfree(b.d_conn);
        dfree(b.d_slot);
    }
    dfree(p->d_nrowptr);
    dfree(p->d_ncols);
    dfree(p->d_bcDof);
    dfree(p->d_bcVal);
    dfree(p->d_bcCount);
    dfree(p->s_coords);
    dfree(p->s_states);
    dfree(p->s_thick);
    dfree(p->s_loads);
    dfree(p->s_K);
    dfree(p->s_R);
    dfree(p->s_out);
    dfree(p->d_flag);
    for (auto& e : p->profEvents) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    if (p->ownStream) cudaStreamDestroy(p->ownStream);
    delete p;
}

extern "C" int femb_plan_set_stream(femb_plan* p, void* cudaStream) {
    if (!p) return fail("femb_plan_set_stream: plan is NULL");
    p->stream = cudaStream ? (cudaStream_t)cudaStream : p->ownStream;
    return 0;
}

extern "C" int femb_plan_sync(femb_plan* p) {
    if (!p) return fail("femb_plan_sync: plan is NULL");
    TRY(use_device(p));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

extern "C" int64_t femb_plan_num_dof(const femb_plan* p) { return p ? p->numNodes * p->numStates : -1; }
extern "C" int64_t femb_plan_num_elems(const femb_plan* p) { return p ? p->numElems : -1; }
extern "C" int64_t femb_plan_nnz(const femb_plan* p) {
    return (p && p->finalized) ? p->nnzNode * (int64_t)p->numStates * p->numStates : -1;
}
extern "C" double femb_plan_build_ms(const femb_plan* p) { return p ? (double)p->buildMs : -1.0; }
extern "C" int64_t femb_plan_last_launches(const femb_plan* p) { return p ? p->launches : -1; }

static int need_final(const femb_plan* p, const char* who) {
    if (!p) return fail("%s: plan is NULL", who);
    if (!p->finalized) return fail("%s: plan is not finalized", who);
    return 0;
}

extern "C" int femb_plan_pattern_dev(femb_plan* p, int64_t* d_indptr, int64_t* d_indices) {
    TRY(need_final(p, "femb_plan_pattern_dev"));
    TRY(use_device(p));
    k_dof_pattern<<<grid_for(p->numNodes + 1, 128), 128, 0, p->stream>>>(
        p->numNodes, p->numStates, p->d_nrowptr, p->d_ncols, (long long*)d_indptr, (long long*)d_indices);
    p->launches = 1;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int femb_plan_pattern(femb_plan* p, int64_t* indptr, int64_t* indices) {
    TRY(need_final(p, "femb_plan_pattern"));
    TRY(use_device(p));
    const int64_t nd = femb_plan_num_dof(p), nnz = femb_plan_nnz(p);
    long long *d_ip = nullptr, *d_ix = nullptr;
    CUDA_TRY(cudaMalloc(&d_ip, (nd + 1) * sizeof(long long)));
    CUDA_TRY(cudaMalloc(&d_ix, std::max<int64_t>(nnz, 1) * sizeof(long long)));
    int rc = femb_plan_pattern_dev(p, (int64_t*)d_ip, (int64_t*)d_ix);
    if (!rc) {
        cudaError_t e1 = cudaMemcpyAsync(indptr, d_ip, (nd + 1) * sizeof(long long), cudaMemcpyDeviceToHost, p->stream);
        cudaError_t e2 = cudaMemcpyAsync(indices, d_ix, nnz * sizeof(long long), cudaMemcpyDeviceToHost, p->stream);
        cudaError_t e3 = cudaStreamSynchronize(p->stream);
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) rc = fail("femb_plan_pattern: copy back failed");
    }
    cudaFree(d_ip);
    cudaFree(d_ix);
    return rc;
}

// =============================================================================================
// C ABI: boundary conditions
// =============================================================================================
extern "C" int femb_plan_set_bcs(femb_plan* p, const int64_t* bcDof, const double* bcVal, int64_t nBC) {
    TRY(need_final(p, "femb_plan_set_bcs"));
    TRY(use_device(p));
    dfree(p->d_bcDof);
    dfree(p->d_bcVal);
    dfree(p->d_bcCount);
    p->nBC = 0;
    if (nBC <= 0) return 0;
    if (!bcDof || !bcVal) return fail("femb_plan_set_bcs: NULL BC arrays");
    const int64_t numDOF = femb_plan_num_dof(p);
    // reference semantics (AssemblyUtils.py:55-60,93): every occurrence adds 1.0 to the diagonal, the
    // last value assigned to a DOF wins in the residual
    std::unordered_map<int64_t, size_t> where;
    std::vector<int> dofs;
    std::vector<double> vals, counts;
    for (int64_t k = 0; k < nBC; ++k) {
        int64_t d = bcDof[k];
        if (d < 0) d += numDOF;  // numpy-style negative index
        if (d < 0 || d >= numDOF) return fail("femb_plan_set_bcs: BC DOF %lld out of range", (long long)bcDof[k]);
        auto it = where.find(d);
        if (it == where.end()) {
            where.emplace(d, dofs.size());
            dofs.push_back((int)d);
            vals.push_back(bcVal[k]);
            counts.push_back(1.0);
        } else {
            vals[it->second] = bcVal[k];
            counts[it->second] += 1.0;
        }
    }
    const size_t n = dofs.size();
    CUDA_TRY(cudaMalloc(&p->d_bcDof, n * sizeof(int)));
    CUDA_TRY(cudaMalloc(&p->d_bcVal, n * sizeof(double)));
    CUDA_TRY(cudaMalloc(&p->d_bcCount, n * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(p->d_bcDof, dofs.data(), n * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->d_bcVal, vals.data(), n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->d_bcCount, counts.data(), n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemsetAsync(p->d_flag, 0, sizeof(int), p->stream));
    k_check_bcs<<<grid_for((int64_t)n, 256), 256, 0, p->stream>>>((long long)n, p->d_bcDof, p->d_nrowptr, p->d_flag);
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, p->d_flag, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (flag) return fail("femb_plan_set_bcs: a BC is applied to a node that belongs to no element (no diagonal entry in the pattern)");
    p->nBC = (int64_t)n;
    return 0;
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
    if (wk) {
        CUDA_TRY(cudaMemsetAsync(d_K, 0, nnz * sizeof(double), st));
    }
    if (wr) {
        k_init_residual<<<grid_for(numDOF, 256), 256, 0, st>>>(d_R, d_loads, numDOF);
        launches++;
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
        launches++;
    }
    if (ev) CUDA_TRY(cudaEventRecord(ev->second, st));
    if (applyBCs && p->nBC > 0) {
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

// =============================================================================================
// C ABI: element functions
// =============================================================================================
template <int NN, int NQ>
static void launch_function(const BlockData& b, const DMat& D, const FuncArgs& F, bool nl, int func, int reduction,
                            cudaStream_t st) {
    const Tables<NN, NQ> tb = make_tables<NN, NQ>(b);
    if (nl) k_function<NN, NQ, true><<<grid_for(F.E, 128), 128, 0, st>>>(tb, D, F, func, reduction);
    else k_function<NN, NQ, false><<<grid_for(F.E, 128), 128, 0, st>>>(tb, D, F, func, reduction);
}

extern "C" int femb_function_dev(femb_plan* p, const double* d_coords, const double* d_states, const femb_material* mat,
                                 int32_t func, int32_t reduction, double* d_out) {
    TRY(need_final(p, "femb_function_dev"));
    if (!mat || !d_coords || !d_states || !d_out) return fail("femb_function_dev: NULL argument");
    if (func != FEMB_FUNC_MASS && func != FEMB_FUNC_VON_MISES) return fail("femb_function_dev: unknown function id %d", func);
    if (reduction < FEMB_REDUCE_NONE || reduction > FEMB_REDUCE_MAX)
        return fail("femb_function_dev: unknown reduction id %d", reduction);
    TRY(use_device(p));
    const DMat D = make_dmat(*mat);
    int64_t launches = 0;
    for (auto& b : p->blocks) {
        FuncArgs F;
        F.E = b.numElems;
        F.conn = b.d_conn;
        F.coords = (const double2*)d_coords;
        F.states = (const double2*)d_states;
        F.out = d_out + (reduction == FEMB_REDUCE_NONE ? b.pointOffset : b.elemOffset);
        FEMB_DISPATCH_FAMILY(b.nn, b.nq, (launch_function<NN, NQ>(b, D, F, mat->nonlinear != 0, func, reduction, p->stream)));
        launches++;
    }
    p->launches = launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int femb_function(femb_plan* p, const double* coords, const double* states, const femb_material* mat,
                             int32_t func, int32_t reduction, double* out) {
    TRY(need_final(p, "femb_function"));
    if (!coords || !states || !out) return fail("femb_function: NULL argument");
    TRY(use_device(p));
    const int64_t numDOF = femb_plan_num_dof(p);
    const size_t nout = (size_t)(reduction == FEMB_REDUCE_NONE ? p->numPoints : p->numElems);
    TRY(stage(&p->s_coords, coords, (size_t)p->numNodes * 2, p->stream));
    TRY(stage(&p->s_states, states, (size_t)numDOF, p->stream));
    if (p->s_outCap < nout) {
        dfree(p->s_out);
        TRY(dalloc(&p->s_out, nout));
        p->s_outCap = nout;
    }
    TRY(femb_function_dev(p, p->s_coords, p->s_states, mat, func, reduction, p->s_out));
    CUDA_TRY(cudaMemcpyAsync(out, p->s_out, nout * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

// =============================================================================================
// C ABI: per-element blocks
// =============================================================================================
template <int NN, int NQ>
static void launch_blocks(const BlockData& b, const DMat& D, const AsmArgs& A, bool nl, double* Ke, double* Re,
                          cudaStream_t st) {
    const Tables<NN, NQ> tb = make_tables<NN, NQ>(b);
    dim3 grid(grid_for(A.E, 128), NN);
    if (nl) k_element_blocks<NN, NQ, true><<<grid, 128, 0, st>>>(tb, D, A, Ke, Re);
    else k_element_blocks<NN, NQ, false><<<grid, 128, 0, st>>>(tb, D, A, Ke, Re);
}

extern "C" int femb_element_blocks(femb_plan* p, const double* coords, const double* states, const double* thickness,
                                   const femb_material* mat, double* Ke, double* Re) {
    TRY(need_final(p, "femb_element_blocks"));
    if (!coords || !states || !thickness || !mat) return fail("femb_element_blocks: NULL argument");
    TRY(use_device(p));
    const int64_t numDOF = femb_plan_num_dof(p);
    TRY(stage(&p->s_coords, coords, (size_t)p->numNodes * 2, p->stream));
    TRY(stage(&p->s_states, states, (size_t)numDOF, p->stream));
    TRY(stage(&p->s_thick, thickness, (size_t)p->numElems, p->stream));
    double *d_Ke = nullptr, *d_Re = nullptr;
    if (Ke) CUDA_TRY(cudaMalloc(&d_Ke, p->keTotal * sizeof(double)));
    if (Re) CUDA_TRY(cudaMalloc(&d_Re, p->reTotal * sizeof(double)));
    const DMat D = make_dmat(*mat);
    int64_t launches = 0;
    for (auto& b : p->blocks) {
        AsmArgs A;
        A.E = b.numElems;
        A.elemOffset = b.elemOffset;
        A.conn = b.d_conn;
        A.slot = nullptr;
        A.nrowptr = nullptr;
        A.coords = (const double2*)p->s_coords;
        A.states = (const double2*)p->s_states;
        A.thick = p->s_thick;
        A.K = nullptr;
        A.R = nullptr;
        FEMB_DISPATCH_FAMILY(b.nn, b.nq, (launch_blocks<NN, NQ>(b, D, A, mat->nonlinear != 0, d_Ke ? d_Ke + b.keOffset : nullptr,
                                                               d_Re ? d_Re + b.reOffset : nullptr, p->stream)));
        launches++;
    }
    p->launches = launches;
    int rc = 0;
    if (Ke && cudaMemcpyAsync(Ke, d_Ke, p->keTotal * sizeof(double), cudaMemcpyDeviceToHost, p->stream) != cudaSuccess) rc = 1;
    if (Re && cudaMemcpyAsync(Re, d_Re, p->reTotal * sizeof(double), cudaMemcpyDeviceToHost, p->stream) != cudaSuccess) rc = 1;
    if (cudaStreamSynchronize(p->stream) != cudaSuccess) rc = 1;
    cudaFree(d_Ke);
    cudaFree(d_Re);
    if (rc) return fail("femb_element_blocks: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// =============================================================================================
// C ABI: AssemblyUtils equivalents on the device
// =============================================================================================
__global__ void k_flag_nonzero(const double* __restrict__ vals, int* __restrict__ flags, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (vals[i] != 0.0) ? 1 : 0;
}

__global__ void k_coo_compact(const double* __restrict__ vals, const long long* __restrict__ dof, const int* __restrict__ flags,
                              const int* __restrict__ pos, long long n, int nd, long long* __restrict__ rows,
                              long long* __restrict__ cols, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    const long long e = i / ((long long)nd * nd);
    const int rc = (int)(i - e * nd * nd);
    const int r = rc / nd, c = rc - r * nd;
    const int q = pos[i];
    rows[q] = dof[e * nd + r];
    cols[q] = dof[e * nd + c];
    out[q] = vals[i];
}

extern "C" int femb_local_matrices_to_coo(const double* localMats, const int64_t* localDOF, int64_t numElems, int32_t nd,
                                          int32_t device, int64_t* rows, int64_t* cols, double* vals, int64_t* numOut) {
    if (!localMats || !localDOF || !rows || !cols || !vals || !numOut) return fail("femb_local_matrices_to_coo: NULL argument");
    if (femb_device_count() == 0) return fail("femb_local_matrices_to_coo: no CUDA device available");
    const int64_t n = numElems * (int64_t)nd * nd;
    if (n >= (1ll << 31)) return fail("femb_local_matrices_to_coo: too many entries");
    *numOut = 0;
    if (n == 0) return 0;
    CUDA_TRY(cudaSetDevice(device));
    double *d_vals = nullptr, *d_out = nullptr;
    long long *d_dof = nullptr, *d_rows = nullptr, *d_cols = nullptr;
    int *d_flags = nullptr, *d_pos = nullptr;
    CUDA_TRY(cudaMalloc(&d_vals, n * sizeof(double)));
    CUDA_TRY(cudaMalloc(&d_out, n * sizeof(double)));
    CUDA_TRY(cudaMalloc(&d_dof, numElems * nd * sizeof(long long)));
    CUDA_TRY(cudaMalloc(&d_rows, n * sizeof(long long)));
    CUDA_TRY(cudaMalloc(&d_cols, n * sizeof(long long)));
    CUDA_TRY(cudaMalloc(&d_flags, n * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_pos, (n + 1) * sizeof(int)));
    CUDA_TRY(cudaMemcpy(d_vals, localMats, n * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_dof, localDOF, numElems * nd * sizeof(long long), cudaMemcpyHostToDevice));
    int64_t launches = 0;
    k_flag_nonzero<<<grid_for(n, 256), 256>>>(d_vals, d_flags, n);
    TRY(exclusive_scan(d_flags, d_pos, n, 0, launches));
    k_coo_compact<<<grid_for(n, 256), 256>>>(d_vals, d_dof, d_flags, d_pos, n, nd, d_rows, d_cols, d_out);
    int total = 0;
    CUDA_TRY(cudaMemcpy(&total, d_pos + n, sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(rows, d_rows, total * sizeof(long long), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(cols, d_cols, total * sizeof(long long), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(vals, d_out, total * sizeof(double), cudaMemcpyDeviceToHost));
    *numOut = total;
    cudaFree(d_vals);
    cudaFree(d_out);
    cudaFree(d_dof);
    cudaFree(d_rows);
    cudaFree(d_cols);
    cudaFree(d_flags);
    cudaFree(d_pos);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

__global__ void k_scatter_residuals(const double* __restrict__ localRes, const long long* __restrict__ conn, long long n,
                                    int s, long long numNodes, double* __restrict__ R, int* flag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over (element, node)
    if (i >= n) return;
    const long long node = conn[i];
    if (node < 0 || node >= numNodes) {
        atomicExch(flag, 1);
        return;
    }
    for (int st = 0; st < s; ++st) atomicAdd(R + node * s + st, localRes[i * s + st]);
}

extern "C" int femb_scatter_local_residuals(const double* localRes, const int64_t* conn, int64_t numElems, int32_t nn,
                                            int32_t numStates, int64_t numNodes, int32_t device, double* globalResidual) {
    if (!localRes || !conn || !globalResidual) return fail("femb_scatter_local_residuals: NULL argument");
    if (femb_device_count() == 0) return fail("femb_scatter_local_residuals: no CUDA device available");
    CUDA_TRY(cudaSetDevice(device));
    const int64_t n = numElems * nn, nd = numNodes * numStates;
    double *d_res = nullptr, *d_R = nullptr;
    long long* d_conn = nullptr;
    int* d_flag = nullptr;
    CUDA_TRY(cudaMalloc(&d_res, std::max<int64_t>(n * numStates, 1) * sizeof(double)));
    CUDA_TRY(cudaMalloc(&d_R, std::max<int64_t>(nd, 1) * sizeof(double)));
    CUDA_TRY(cudaMalloc(&d_conn, std::max<int64_t>(n, 1) * sizeof(long long)));
    CUDA_TRY(cudaMalloc(&d_flag, sizeof(int)));
    CUDA_TRY(cudaMemset(d_flag, 0, sizeof(int)));
    CUDA_TRY(cudaMemcpy(d_res, localRes, n * numStates * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_R, globalResidual, nd * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_conn, conn, n * sizeof(long long), cudaMemcpyHostToDevice));
    if (n > 0) k_scatter_residuals<<<grid_for(n, 256), 256>>>(d_res, d_conn, n, numStates, numNodes, d_R, d_flag);
    int flag = 0;
    CUDA_TRY(cudaMemcpy(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(globalResidual, d_R, nd * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_res);
    cudaFree(d_R);
    cudaFree(d_conn);
    cudaFree(d_flag);
    if (flag) return fail("femb_scatter_local_residuals: connectivity out of range");
    return 0;
}
