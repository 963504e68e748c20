// Internal declarations shared by the translation units of libfemb200.so.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/femb200.h"
#include "elem_math.cuh"

// ---------------------------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------------------------
#ifndef FEMB_ROT4
#define FEMB_ROT4 1  // 4-node quads: corner records hold the element rotated to the owner (plan.cu: k_corners); 0: {e, a | slots} + connectivity load
#endif
int femb_fail(const char* fmt, ...);
#define fail femb_fail

#define CUDA_TRY(expr)                                                                                                  \
    do {                                                                                                                \
        cudaError_t _e = (expr);                                                                                        \
        if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define TRY(expr)          \
    do {                   \
        int _r = (expr);   \
        if (_r) return _r; \
    } while (0)

// ---------------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------------
static constexpr int MAX_BLOCKS = 8;

struct BlockData {
    int nn = 0, nq = 0;
    int64_t numElems = 0;
    int64_t elemOffset = 0;   // first global element id
    int64_t pointOffset = 0;  // first global Gauss point (for FEMB_REDUCE_NONE output)
    int64_t keOffset = 0, reOffset = 0;
    std::vector<double> N, dN, w;
    bool closedForm = false;  // 4 nodes: standard bilinear tables + 2x2 Gauss rule -> closed-form row-owner kernel (plan.cu: standard_q4)
    bool cyclic = true;       // 3/4 nodes: tables invariant under cyclic renumbering -> corner records rotated to the owner
    int* d_conn = nullptr;  // node-major int32: conn[a * E + e]
    double *d_tabN = nullptr, *d_tabdN = nullptr, *d_tabw = nullptr;  // 1-D / 3-D plans: the block's tables on the device (generic.cu)
    int* d_slot = nullptr;  // [(a*nn + b) * E + e] -> position of (node_a, node_b) in the node-level CSR
    int4* d_connE = nullptr;   // element-major connectivity, ceil(nn/4) words per element (row-owner kernel)
    void* d_corners = nullptr; // corner records of this block, indexed by the global corner number (plan.cu: k_corners)
    int* d_cornerElem = nullptr;  // rotated 4-node records: element id of each corner
    int4* d_fanRec = nullptr;     // plan-time: fan-ordered corner records of the write-once kernel (plan.cu: k_fan_corners)
    int* d_fanElem = nullptr;     // plan-time, 4 nodes: element id of each fan-ordered corner
    int4* d_fanHdr = nullptr;     // per node {row start, rest index, corners | degree << 8, BC word} + first corner record (plan.cu: k_fan_headers)
    int4* d_fanRest = nullptr;    // records of every node's further corners
    long long fanRestTotal = 0;   // records in d_fanRest
    bool fanCompact = false;      // 4 nodes: node / element ids as 24-bit differences to the owner, no element array
    bool fanOK = false;           // the corners of every node chain into fans: the write-once kernel applies
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
    int numSMs = 0;  // cached on first use
    int64_t numNodes = 0;
    int numDim = 2, numStates = 2;
    std::vector<BlockData> blocks;
    int64_t numElems = 0, numPoints = 0, keTotal = 0, reTotal = 0;
    bool finalized = false;
    int* d_nrowptr = nullptr;  // node-level CSR row pointers (numNodes + 1)
    int* d_ncols = nullptr;    // node-level CSR column nodes, sorted within a row
    int64_t nnzNode = 0;
    int* d_adjPtr = nullptr;   // node -> adjacent elements (global element ids), sorted per node
    int* d_adj = nullptr;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    float buildMs = 0.f;
    int64_t launches = 0;
    long long* d_rowStats = nullptr;  // measurement builds only (ROWS_TRACE): warp lifetime statistics
    int assemblyPath = 2;        // 2 row-owner kernel, 0 slot-map scatter with FP64 atomics
    bool fanCompactAllowed = true;  // option "fan_compact" (tests): 0 keeps the wide fan records even when the differences fit
    int fanPrefetch = -1;        // option "fan_prefetch": L2 look-ahead of the fan kernel in 32-node warps (-1: a sixth of the resident wave, 0: off)
    bool pointsOnly = false;     // option "points_only": blocks of any size, femb_point_quantities only
    int fanCarveout = -1;        // option "fan_carveout": shared-memory carve-out of the fan kernel in percent (-1: what the resident warps need)
    bool useFan = true;          // option "fan_rows": write-once fan kernel where it applies (else the accumulating row-owner kernel)
    int* d_blockAdj = nullptr;   // (numBlocks+1) x numNodes: start of each block's corners in a node's adjacency list (row-owner kernel)
    int* d_rowOrder = nullptr;   // plan-time temporary: balanced row order (plan.cu: k_row_order), node of every thread slot, -1 = padding
    int4* d_slotRec = nullptr;   // balanced row schedule: one record per thread slot (plan.cu: k_slot_records); NULL = node order
    long long rowSlots = 0;      // thread slots of the balanced schedule (multiple of the sort window)
    long long rowIters[2] = {0, 0};  // sum over warps of the longest corner walk: node order / balanced order
    bool rowsBalanced = false;
    int balanceRows = 1;         // option "balance_rows": 0 never, 1 when it saves > 8 % of the corner iterations, 2 always
    bool rowsBuilt = false;      // corner records / element-major connectivity exist for every block
    int maxDeg = 0;              // largest node degree
    // boundary conditions, de-duplicated on the host with the reference's semantics
    std::vector<int64_t> bcRawDof;  // the lists as last given (identical lists are not processed again)
    std::vector<double> bcRawVal;
    bool bcValid = false;
    int64_t nBC = 0;
    int* d_bcDof = nullptr;       // sorted ascending
    double* d_bcVal = nullptr;    // last value given for the DOF
    double* d_bcCount = nullptr;  // number of occurrences (diagonal value)
    unsigned char* d_bcMask = nullptr;  // (numNodes) bit r set when DOF 2*node + r is constrained
    int* d_bcInfo = nullptr;        // (numNodes) (index of the constrained node + 1) << 9 | position in own row << 2 | DOF bits; 0 = unconstrained
    double* d_bcNodeCnt = nullptr;  // per constrained node: diagonal values (occurrence counts) of its two DOFs
    double* d_bcNodeVal = nullptr;  // per constrained node: prescribed values of its two DOFs
    // scratch for the host-pointer API
    double *s_coords = nullptr, *s_states = nullptr, *s_thick = nullptr, *s_loads = nullptr, *s_K = nullptr,
           *s_R = nullptr, *s_out = nullptr;
    size_t s_outCap = 0;
    bool coordsResident = false, thickResident = false;  // femb_plan_set_geometry: s_coords / s_thick hold the model's arrays
    void* s_solve = nullptr;  // work vectors + scalars of femb_solve_pcg_dev (solve.cu)
    int* d_flag = nullptr;
    // optional CUDA-event timing of the dominant assembly kernel(s) (bench.py roofline leg)
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> profEvents;
    size_t profUsed = 0;
};

template <typename T>
static inline int dalloc(T** p, size_t n) {
    if (*p) return 0;
    CUDA_TRY(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    return 0;
}
template <typename T>
static inline void dfree(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

// device scratch released when the scope ends: the C ABI's error paths return early (CUDA_TRY / TRY)
template <typename T>
struct Scratch {
    T* p = nullptr;
    Scratch() = default;
    Scratch(const Scratch&) = delete;
    Scratch& operator=(const Scratch&) = delete;
    ~Scratch() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t n) { return cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T)); }
    operator T*() const { return p; }
};
struct ScopedEvent {
    cudaEvent_t e = nullptr;
    ~ScopedEvent() {
        if (e) cudaEventDestroy(e);
    }
    cudaError_t create() { return cudaEventCreate(&e); }
    operator cudaEvent_t() const { return e; }
};

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

static inline int use_device(const femb_plan* p) {
    CUDA_TRY(cudaSetDevice(p->device));
    return 0;
}

// host-pointer entries: upload coords / thickness, or -- pointer NULL -- use the copy femb_plan_set_geometry left on the
// device.  An explicit pointer overwrites the scratch, so the resident promise ends there.
static inline int stage_geometry(femb_plan* p, const double* coords, const double* thickness, bool needThickness, const char* who) {
    if (coords) {
        if (!p->s_coords) CUDA_TRY(cudaMalloc((void**)&p->s_coords, std::max<size_t>((size_t)p->numNodes * p->numDim, 1) * sizeof(double)));
        CUDA_TRY(cudaMemcpyAsync(p->s_coords, coords, (size_t)p->numNodes * p->numDim * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        p->coordsResident = false;
    } else if (!p->coordsResident) {
        return fail("%s: coords is NULL and no resident copy exists (femb_plan_set_geometry)", who);
    }
    if (!needThickness) return 0;
    if (thickness) {
        if (!p->s_thick) CUDA_TRY(cudaMalloc((void**)&p->s_thick, std::max<size_t>((size_t)p->numElems, 1) * sizeof(double)));
        CUDA_TRY(cudaMemcpyAsync(p->s_thick, thickness, (size_t)p->numElems * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        p->thickResident = false;
    } else if (!p->thickResident) {
        return fail("%s: thickness is NULL and no resident copy exists (femb_plan_set_geometry)", who);
    }
    return 0;
}

static inline int need_final(const femb_plan* p, const char* who) {
    if (!p) return fail("%s: plan is NULL", who);
    if (!p->finalized) return fail("%s: plan is not finalized", who);
    return 0;
}

// out[0..n) = exclusive prefix sums of in[0..n); out[n] = total.  in and out must not alias.  Synchronises.
int exclusive_scan(const int* d_in, int* d_out, int64_t n, cudaStream_t st, int64_t& launches);

bool rows_path_available(const femb_plan* p);
bool fan_path_available(const femb_plan* p);  // assemble.cu: the write-once fan kernel will run (small-strain model)
// generic.cu: the 1-D / 3-D families (run-time node and point counts, FP64-atomic scatter)
int generic_assemble_dev(femb_plan* p, const double* d_coords, const double* d_states, const double* d_scale, const femb_material* mat,
                         bool bcs, const double* d_loads, bool wk, bool wr, double* d_K, double* d_R);
int generic_function_dev(femb_plan* p, const double* d_coords, const double* d_states, const femb_material* mat, int func, int reduction,
                         double* d_out);
int generic_element_blocks_dev(femb_plan* p, const double* d_coords, const double* d_states, const double* d_scale, const femb_material* mat,
                               double* d_Ke, double* d_Re);
int ensure_slot_maps(femb_plan* p);  // plan.cu: builds the scatter kernels' slot maps on first use

// ---------------------------------------------------------------------------------------------
// kernel argument blocks
// ---------------------------------------------------------------------------------------------
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

struct FuncArgs {
    long long E;
    const int* conn;
    const double2* coords;
    const double2* states;
    double* out;  // E or E*NQ values
};

// ---------------------------------------------------------------------------------------------
// element families
// ---------------------------------------------------------------------------------------------
template <int NN, int NQ>
static inline femb::Tables<NN, NQ> make_tables(const BlockData& b) {
    femb::Tables<NN, NQ> t;
    for (int p = 0; p < NQ; ++p) {
        for (int d = 0; d < 2; ++d)
            for (int a = 0; a < NN; ++a) t.dN[p][d][a] = b.dN[(p * 2 + d) * NN + a];
        for (int a = 0; a < NN; ++a) t.N[p][a] = b.N[p * NN + a];
        t.w[p] = b.w[p];
    }
    return t;
}

static inline femb::DMat make_dmat(const femb_material& m) {
    femb::DMat D;
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

// the families' default rules (QuadElement2D.py:65-67, TriElement2D.py:92-94): the row-owner / fan kernels are built for these
static inline bool default_rule(int nn, int nq) {
    return (nn == 3 && nq == 1) || (nn == 6 && nq == 3) || (nn == 10 && nq == 3) || (nn == 4 && nq == 4) ||
           (nn == 9 && nq == 9) || (nn == 16 && nq == 16);
}
// other quadrature rules (`intOrder=` of the element-level methods, Element.py:163,370,426; a model may also be built on
// them): n x n Gauss rules with n = 1..4 for the quads, the 1- / 3- / 4-point rules for the triangles
// (FEMpy/Quadrature/TriQuad.py:28-35).  These run on the scatter kernels, the element-block and the function kernels.
static inline bool extra_rule(int nn, int nq) {
    return (nn == 4 && (nq == 1 || nq == 9 || nq == 16)) || (nn == 9 && (nq == 4 || nq == 16)) || (nn == 16 && nq == 9) ||
           (nn == 3 && (nq == 3 || nq == 4)) || (nn == 6 && (nq == 1 || nq == 4)) || (nn == 10 && (nq == 1 || nq == 4));
}
static inline bool family_supported(int nn, int nq) { return default_rule(nn, nq) || extra_rule(nn, nq); }

// body sees constexpr NN, NQ and ROW (row-split kernels for the families that do not fit one thread)
#define FEMB_DISPATCH_FAMILY(nn, nq, BODY)                                                                                 \
    do {                                                                                                                   \
        if ((nn) == 3 && (nq) == 1) { constexpr int NN = 3, NQ = 1; constexpr bool ROW = false; (void)ROW; BODY; }         \
        else if ((nn) == 4 && (nq) == 4) { constexpr int NN = 4, NQ = 4; constexpr bool ROW = false; (void)ROW; BODY; }    \
        else if ((nn) == 6 && (nq) == 3) { constexpr int NN = 6, NQ = 3; constexpr bool ROW = true; (void)ROW; BODY; }     \
        else if ((nn) == 9 && (nq) == 9) { constexpr int NN = 9, NQ = 9; constexpr bool ROW = true; (void)ROW; BODY; }     \
        else if ((nn) == 10 && (nq) == 3) { constexpr int NN = 10, NQ = 3; constexpr bool ROW = true; (void)ROW; BODY; }   \
        else if ((nn) == 16 && (nq) == 16) { constexpr int NN = 16, NQ = 16; constexpr bool ROW = true; (void)ROW; BODY; } \
        else return fail("unsupported element family: %d nodes, %d quadrature points", (nn), (nq));                        \
    } while (0)

// every supported (nodes, rule) pair: the default rules above + the extra rules (generic kernels only)
#define FEMB_CASE(n_, q_, row_, BODY) if ((nn_) == n_ && (nq_) == q_) { constexpr int NN = n_, NQ = q_; constexpr bool ROW = row_; (void)ROW; BODY; }
#define FEMB_DISPATCH_ANY(nnArg, nqArg, BODY)                                                                              \
    do {                                                                                                                   \
        const int nn_ = (nnArg), nq_ = (nqArg);                                                                            \
        FEMB_CASE(3, 1, false, BODY) else FEMB_CASE(3, 3, false, BODY) else FEMB_CASE(3, 4, false, BODY)                   \
        else FEMB_CASE(4, 4, false, BODY) else FEMB_CASE(4, 1, false, BODY) else FEMB_CASE(4, 9, false, BODY)              \
        else FEMB_CASE(4, 16, false, BODY) else FEMB_CASE(6, 3, true, BODY) else FEMB_CASE(6, 1, true, BODY)               \
        else FEMB_CASE(6, 4, true, BODY) else FEMB_CASE(9, 9, true, BODY) else FEMB_CASE(9, 4, true, BODY)                 \
        else FEMB_CASE(9, 16, true, BODY) else FEMB_CASE(10, 3, true, BODY) else FEMB_CASE(10, 1, true, BODY)              \
        else FEMB_CASE(10, 4, true, BODY) else FEMB_CASE(16, 16, true, BODY) else FEMB_CASE(16, 9, true, BODY)             \
        else return fail("unsupported element family: %d nodes, %d quadrature points", nn_, nq_);                          \
    } while (0)
