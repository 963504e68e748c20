// Device-resident linear solve of the assembled system K x = b (the hand-off after assembly).
//
// Replaces, on request, the host factorisation in FEMpyProblem.solveJacLinear (FEMpy/Problem.py:171-198: scipy
// `factorized` / pypardiso on the csr_array) with a Jacobi-preconditioned conjugate-gradient iteration that runs on
// the plan's node-block CSR values where the assembly kernel left them: no D2H of K, no csr_array rebuild.
//
// The reference applies boundary conditions to ROWS only (AssemblyUtils.applyBCsToMatrix, AssemblyUtils.py:26-62): a
// constrained row is zero except for its diagonal (the occurrence count), the columns keep their couplings, so K is not
// symmetric.  With x_c = b_c / count known, the free block is solved instead:  K_ff x_f = b_f - K_fc x_c,  which is
// symmetric positive definite for the elasticity models.  In the iteration this is a mask: search directions are zero
// on constrained DOFs (their columns drop out) and the product is zeroed on constrained rows.
//
// Per iteration three launches, no host synchronisation (alpha, beta live in device scalars, double-buffered by iteration
// parity); the host reads the residual norm every `CHECK` iterations.
//   k_spmv   q = K p (4 lanes per node row: the two DOF rows of a node are one contiguous run),  pq = p.q
//   k_update alpha = rz/pq;  x += alpha p;  r -= alpha q;  z = r / diag;  rz' = r.z;  rr = r.r
//   k_direct beta = rz'/rz;  p = z + beta p
// HBM traffic per iteration: K values once (8 B/nnz) + node columns (1 B/nnz) + 9 vector passes.
#include "femb_internal.h"

using namespace femb;

namespace {

#ifndef SPMV_LANES
#define SPMV_LANES 4  // A/B at config 2 (deg 9): 1 lane 44.1, 2 lanes 39.7, 4 lanes 40.5, 8 lanes 48.5, 16 lanes 67.1 us per iteration
#endif
constexpr int LANES = SPMV_LANES;  // lanes that share a node row in the product
constexpr int CHECK = 25;  // iterations between convergence checks

struct Scal {  // device scalars, [parity]
    double rz[2], pq[2], rr[2];
};

__device__ __forceinline__ void block_add(double v, double* target) {
    __shared__ double part[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane == 0) part[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < (int)(blockDim.x >> 5) ? part[lane] : 0.0;
#pragma unroll
        for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) atomicAdd(target, v);
    }
    __syncthreads();
}

// q = K v for the node-block CSR; masked != 0: rows of constrained DOFs give 0 (v is zero there by construction).
// dotTarget (optional) accumulates v.q.  zero0 / zero1 (optional): scalars reset for the kernels that follow.
__global__ void __launch_bounds__(256) k_spmv(long long N, const int* __restrict__ nrowptr, const int* __restrict__ ncols,
                                              const double2* __restrict__ K2, const double2* __restrict__ v,
                                              const unsigned char* __restrict__ mask, int masked, double2* __restrict__ q,
                                              double* dotTarget, double* zero0, double* zero1) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (zero0) *zero0 = 0.0;
        if (zero1) *zero1 = 0.0;
    }
    const long long node = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const int j = threadIdx.x & (LANES - 1);
    double s0 = 0.0, s1 = 0.0, dot = 0.0;
    if (node < N) {
        const int off = __ldg(nrowptr + node), deg = __ldg(nrowptr + node + 1) - off;
        const double2* row0 = K2 + 2ll * off;  // DOF row 2*node: deg (xx, xy) pairs, then DOF row 2*node + 1: (yx, yy) pairs
        const double2* row1 = row0 + deg;
        for (int k = j; k < deg; k += LANES) {
            const double2 xv = __ldg(v + __ldg(ncols + off + k));
            const double2 a = __ldg(row0 + k), b = __ldg(row1 + k);
            s0 = fma(a.x, xv.x, fma(a.y, xv.y, s0));
            s1 = fma(b.x, xv.x, fma(b.y, xv.y, s1));
        }
    }
#pragma unroll
    for (int d = LANES / 2; d; d >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, d);
        s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    }
    if (node < N && j == 0) {
        const unsigned m = (masked && mask) ? mask[node] : 0u;
        if (m & 1u) s0 = 0.0;
        if (m & 2u) s1 = 0.0;
        q[node] = make_double2(s0, s1);
        if (dotTarget) {
            const double2 pv = v[node];
            dot = pv.x * s0 + pv.y * s1;
        }
    }
    if (dotTarget) block_add(dot, dotTarget);
}

// 1 / diagonal of every DOF (position of the node in its own sorted row found by a walk over the row)
__global__ void k_diag_inv(long long N, const int* __restrict__ nrowptr, const int* __restrict__ ncols, const double2* __restrict__ K2,
                           double2* __restrict__ dinv) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int off = nrowptr[i], deg = nrowptr[i + 1] - off;
    double2 d = make_double2(1.0, 1.0);  // nodes outside every element: identity
    for (int k = 0; k < deg; ++k)
        if (ncols[off + k] == (int)i) d = make_double2(K2[2ll * off + k].x, K2[2ll * off + deg + k].y);
    dinv[i] = make_double2(d.x != 0.0 ? 1.0 / d.x : 1.0, d.y != 0.0 ? 1.0 / d.y : 1.0);
}

// x = 0 except x_c = b_c / count on the constrained DOFs
__global__ void k_known_values(long long nBC, const int* __restrict__ bcDof, const double* __restrict__ bcCount, const double* __restrict__ b,
                               double* __restrict__ x) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nBC) x[bcDof[k]] = b[bcDof[k]] / bcCount[k];
}

// r = mask(b - q);  z = r / diag;  p = z;  rz = r.z;  rr = r.r
__global__ void __launch_bounds__(256) k_start(long long N, const double2* __restrict__ b, const double2* __restrict__ q,
                                               const double2* __restrict__ dinv, const unsigned char* __restrict__ mask,
                                               double2* __restrict__ r, double2* __restrict__ p, Scal* S) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double rz = 0.0, rr = 0.0;
    if (i < N) {
        const unsigned m = mask ? mask[i] : 0u;
        double2 rv = make_double2(b[i].x - q[i].x, b[i].y - q[i].y);
        if (m & 1u) rv.x = 0.0;
        if (m & 2u) rv.y = 0.0;
        const double2 di = dinv[i], z = make_double2(rv.x * di.x, rv.y * di.y);
        r[i] = rv;
        p[i] = z;
        rz = rv.x * z.x + rv.y * z.y;
        rr = rv.x * rv.x + rv.y * rv.y;
    }
    block_add(rz, &S->rz[0]);
    block_add(rr, &S->rr[0]);
}

// iteration `par` parity: alpha = rz[par] / pq[par];  x += alpha p;  r -= alpha q;  z = r / diag;  rz[par^1] += r.z;  rr[par^1] += r.r
__global__ void __launch_bounds__(256) k_update(long long N, int par, const double2* __restrict__ p, const double2* __restrict__ q,
                                                const double2* __restrict__ dinv, double2* __restrict__ x, double2* __restrict__ r,
                                                double2* __restrict__ z, Scal* S) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double pq = S->pq[par];
    const double alpha = pq != 0.0 ? S->rz[par] / pq : 0.0;
    double rz = 0.0, rr = 0.0;
    if (i < N) {
        const double2 pv = p[i], qv = q[i], di = dinv[i];
        double2 xv = x[i], rv = r[i];
        xv.x = fma(alpha, pv.x, xv.x);
        xv.y = fma(alpha, pv.y, xv.y);
        rv.x = fma(-alpha, qv.x, rv.x);
        rv.y = fma(-alpha, qv.y, rv.y);
        const double2 zv = make_double2(rv.x * di.x, rv.y * di.y);
        x[i] = xv;
        r[i] = rv;
        z[i] = zv;
        rz = rv.x * zv.x + rv.y * zv.y;
        rr = rv.x * rv.x + rv.y * rv.y;
    }
    block_add(rz, &S->rz[par ^ 1]);
    block_add(rr, &S->rr[par ^ 1]);
}

// beta = rz[par^1] / rz[par];  p = z + beta p
__global__ void __launch_bounds__(256) k_direct(long long N, int par, const double2* __restrict__ z, double2* __restrict__ p, Scal* S) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const double old = S->rz[par];
    const double beta = old != 0.0 ? S->rz[par ^ 1] / old : 0.0;
    if (i == 0) S->pq[par] = 0.0;  // consumed by k_update; free for iteration it + 2
    if (i < N) {
        const double2 zv = z[i], pv = p[i];
        p[i] = make_double2(fma(beta, pv.x, zv.x), fma(beta, pv.y, zv.y));
    }
}

}  // namespace

extern "C" int femb_solve_pcg_dev(femb_plan* p, const double* d_K, const double* d_b, double* d_x, double relTol, int32_t maxIter,
                                  int32_t* itersOut, double* relResOut) {
    TRY(need_final(p, "femb_solve_pcg_dev"));
    if (!d_K || !d_b || !d_x) return fail("femb_solve_pcg_dev: NULL argument");
    if (p->numStates != 2) return fail("femb_solve_pcg_dev: written for 2 states per node");
    if (!(relTol > 0.0) || maxIter < 1) return fail("femb_solve_pcg_dev: relTol must be > 0 and maxIter >= 1");
    TRY(use_device(p));
    cudaStream_t st = p->stream;
    const long long N = p->numNodes;
    const unsigned char* mask = p->bcValid ? p->d_bcMask : nullptr;
    const double2* K2 = (const double2*)d_K;
    // work vectors r, p, q, z, 1/diag (double2 per node) + scalars: kept in the plan between solves
    if (!p->s_solve) CUDA_TRY(cudaMalloc(&p->s_solve, (size_t)N * 5 * sizeof(double2) + sizeof(Scal)));
    double2* r = (double2*)p->s_solve;
    double2 *pv = r + N, *q = pv + N, *z = q + N, *dinv = z + N;
    Scal* S = (Scal*)(dinv + N);
    double2* x = (double2*)d_x;
    const unsigned gN = grid_for(N, 256), gS = grid_for(N * LANES, 256);

    CUDA_TRY(cudaMemsetAsync(S, 0, sizeof(Scal), st));
    CUDA_TRY(cudaMemsetAsync(d_x, 0, (size_t)N * sizeof(double2), st));
    k_diag_inv<<<gN, 256, 0, st>>>(N, p->d_nrowptr, p->d_ncols, K2, dinv);
    if (mask && p->nBC > 0) k_known_values<<<grid_for(p->nBC, 256), 256, 0, st>>>(p->nBC, p->d_bcDof, p->d_bcCount, d_b, d_x);
    k_spmv<<<gS, 256, 0, st>>>(N, p->d_nrowptr, p->d_ncols, K2, x, mask, 0, q, nullptr, nullptr, nullptr);  // K x_known (all rows)
    k_start<<<gN, 256, 0, st>>>(N, (const double2*)d_b, q, dinv, mask, r, pv, S);
    int64_t launches = 4 + (mask ? 1 : 0);
    Scal h;
    CUDA_TRY(cudaMemcpyAsync(&h, S, sizeof(Scal), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const double rr0 = h.rr[0];
    int it = 0;
    double rr = rr0;
    if (rr0 > 0.0) {
        const double target = relTol * relTol * rr0;
        while (it < maxIter && rr > target) {
            const int batch = std::min(CHECK, maxIter - it);
            for (int k = 0; k < batch; ++k, ++it) {
                const int par = it & 1;
                // the accumulators of this iteration's k_update (parity par^1) are reset by its k_spmv; k_direct resets pq[par]
                // once k_update has consumed it
                k_spmv<<<gS, 256, 0, st>>>(N, p->d_nrowptr, p->d_ncols, K2, pv, mask, 1, q, &S->pq[par], &S->rz[par ^ 1], &S->rr[par ^ 1]);
                k_update<<<gN, 256, 0, st>>>(N, par, pv, q, dinv, x, r, z, S);
                k_direct<<<gN, 256, 0, st>>>(N, par, z, pv, S);
            }
            launches += 3 * batch;
            CUDA_TRY(cudaMemcpyAsync(&h, S, sizeof(Scal), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            rr = h.rr[it & 1];
            if (!(rr == rr)) return fail("femb_solve_pcg_dev: the iteration broke down (NaN residual after %d iterations); is K positive definite on the free DOFs?", it);
        }
    }
    CUDA_TRY(cudaGetLastError());
    p->launches = launches;
    if (itersOut) *itersOut = it;
    if (relResOut) *relResOut = rr0 > 0.0 ? sqrt(rr / rr0) : 0.0;
    return 0;
}

extern "C" int femb_solve_pcg(femb_plan* p, const double* K, const double* b, double* x, double relTol, int32_t maxIter, int32_t* itersOut,
                              double* relResOut) {
    TRY(need_final(p, "femb_solve_pcg"));
    if (!K || !b || !x) return fail("femb_solve_pcg: NULL argument");
    TRY(use_device(p));
    const size_t nnz = (size_t)femb_plan_nnz(p), numDOF = (size_t)femb_plan_num_dof(p);
    TRY(dalloc(&p->s_K, nnz));
    TRY(dalloc(&p->s_R, numDOF));
    TRY(dalloc(&p->s_loads, numDOF));
    CUDA_TRY(cudaMemcpyAsync(p->s_K, K, nnz * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->s_loads, b, numDOF * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    TRY(femb_solve_pcg_dev(p, p->s_K, p->s_loads, p->s_R, relTol, maxIter, itersOut, relResOut));
    CUDA_TRY(cudaMemcpyAsync(x, p->s_R, numDOF * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

__global__ void k_negate(long long n, const double* __restrict__ a, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = -a[i];
}

static int stage_in(double** buf, const double* host, size_t n, cudaStream_t st) {
    TRY(dalloc(buf, n));
    CUDA_TRY(cudaMemcpyAsync(*buf, host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    return 0;
}

// One Newton / linear step with everything between the inputs and the state update on the device
// (FEMpyProblem.solve, Problem.py:137-169: residual + Jacobian assembly with BCs, update = K^-1 (-R)):
// H2D of the nodal inputs, ONE fused assembly, PCG on the values where the kernel left them, D2H of the update
// (and of R when asked) -- the CSR values never cross PCIe.
extern "C" int femb_assemble_solve(femb_plan* p, const double* coords, const double* states, const double* thickness,
                                   const femb_material* mat, const int64_t* bcDof, const double* bcVal, int64_t nBC,
                                   const double* loads, double relTol, int32_t maxIter, double* update, double* R,
                                   int32_t* itersOut, double* relResOut) {
    TRY(need_final(p, "femb_assemble_solve"));
    if (!states || !mat || !update) return fail("femb_assemble_solve: NULL argument");
    TRY(use_device(p));
    const size_t numDOF = (size_t)femb_plan_num_dof(p), nnz = (size_t)femb_plan_nnz(p);
    TRY(femb_plan_set_bcs(p, bcDof, bcVal, nBC));
    TRY(stage_geometry(p, coords, thickness, true, "femb_assemble_solve"));
    TRY(stage_in(&p->s_states, states, numDOF, p->stream));
    if (loads) TRY(stage_in(&p->s_loads, loads, numDOF, p->stream));
    else TRY(dalloc(&p->s_loads, numDOF));
    TRY(dalloc(&p->s_K, nnz));
    TRY(dalloc(&p->s_R, numDOF));
    TRY(femb_assemble_dev(p, p->s_coords, p->s_states, p->s_thick, mat, 1, loads ? p->s_loads : nullptr, FEMB_WANT_K | FEMB_WANT_R,
                          p->s_K, p->s_R));
    const int64_t asmLaunches = p->launches;
    if (R) CUDA_TRY(cudaMemcpyAsync(R, p->s_R, numDOF * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    k_negate<<<grid_for((int64_t)numDOF, 256), 256, 0, p->stream>>>((long long)numDOF, p->s_R, p->s_loads);  // rhs = -R (loads consumed)
    TRY(femb_solve_pcg_dev(p, p->s_K, p->s_loads, p->s_states, relTol, maxIter, itersOut, relResOut));  // states consumed: holds the update
    p->launches += asmLaunches + 1;
    CUDA_TRY(cudaMemcpyAsync(update, p->s_states, numDOF * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}
