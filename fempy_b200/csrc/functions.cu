// Element functions (von Mises, mass), per-element blocks and the AssemblyUtils equivalents.
#include "femb_internal.h"

using namespace femb;

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

static int stage(double** buf, const double* host, size_t n, cudaStream_t st) {
    TRY(dalloc(buf, n));
    CUDA_TRY(cudaMemcpyAsync(*buf, host, n * sizeof(double), cudaMemcpyHostToDevice, st));
    return 0;
}

// =============================================================================================
// element function kernel
// =============================================================================================
// KS: separate instantiation for the KS aggregate (second pass + exp/log) so the plain reductions keep their registers
template <int NN, int NQ, bool NL, bool KS>
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
    // rolled loop over the Gauss points, reduction on the fly (Element.py:221-273)
    double r = 0.0;
#pragma unroll 1
    for (int p = 0; p < NQ; ++p) {
        double v, det;
        point_function<NN, NQ, NL>(tb, D, func, p, x, y, ux, uy, v, det);
        if (reduction == FEMB_REDUCE_NONE) F.out[e * NQ + p] = v;
        else if (reduction == FEMB_REDUCE_SUM || reduction == FEMB_REDUCE_MEAN) r += v;
        else if (reduction == FEMB_REDUCE_INTEGRATE) r = fma(v * det, tb.w[p], r);
        else if (reduction == FEMB_REDUCE_MIN) r = p ? fmin(r, v) : v;
        else r = p ? fmax(r, v) : v;
    }
    if (reduction == FEMB_REDUCE_NONE) return;
    if (reduction == FEMB_REDUCE_MEAN) r /= (double)NQ;
    if (KS) {  // r holds the maximum; second pass re-evaluates the (bit-identical) point values
        constexpr double rho = 100.0;      // KSAgg.py:26
        double s = 0.0;
#pragma unroll 1
        for (int p = 0; p < NQ; ++p) {
            double v, det;
            point_function<NN, NQ, NL>(tb, D, func, p, x, y, ux, uy, v, det);
            s += exp(rho * (v - r));
        }
        r += 1.0 / rho * log(1.0 / (double)NQ * s);
    }
    F.out[e] = r;
}

// Closed-form variant for the standard bilinear quad with the 2 x 2 Gauss rule, small strain (the plan checks the tables:
// plan.cu standard_q4): with A_f = f1 - f0 + f2 - f3, B_f = f3 - f0 + f2 - f1, C_f = f0 - f1 + f2 - f3 for a nodal field f,
//     4 df/dxi = A_f + eta C_f,   4 df/deta = B_f + xi C_f,
// so J and H = dN q cost one FMA per entry and point instead of a 4-term table contraction (the kernel is FP64-bound at
// config 5: 75 -> 50 FP64 instructions per point).  The factors 1/4 cancel in u' = J^-1 H.  Element-major connectivity:
// one 16-byte load per element.  Same reductions, same point order as the tables (xi_p, eta_p read from them on the host).
struct Q4Points {
    double xi[4], eta[4], w[4];
};
template <bool KS>
__global__ void __launch_bounds__(128) k_function_q4(const __grid_constant__ Q4Points P, const DMat D, const long long E,
                                                     const int4* __restrict__ connE, const double2* __restrict__ coords,
                                                     const double2* __restrict__ states, double* __restrict__ out, const int func,
                                                     const int reduction) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int4 c = __ldg(connE + e);
    const double2 X0 = __ldg(coords + c.x), X1 = __ldg(coords + c.y), X2 = __ldg(coords + c.z), X3 = __ldg(coords + c.w);
    const double2 q0 = __ldg(states + c.x), q1 = __ldg(states + c.y), q2 = __ldg(states + c.z), q3 = __ldg(states + c.w);
    const double ax = (X1.x - X0.x) + (X2.x - X3.x), bx = (X3.x - X0.x) + (X2.x - X1.x), cx = (X0.x - X1.x) + (X2.x - X3.x);
    const double ay = (X1.y - X0.y) + (X2.y - X3.y), by = (X3.y - X0.y) + (X2.y - X1.y), cy = (X0.y - X1.y) + (X2.y - X3.y);
    const double au = (q1.x - q0.x) + (q2.x - q3.x), bu = (q3.x - q0.x) + (q2.x - q1.x), cu = (q0.x - q1.x) + (q2.x - q3.x);
    const double av = (q1.y - q0.y) + (q2.y - q3.y), bv = (q3.y - q0.y) + (q2.y - q1.y), cv = (q0.y - q1.y) + (q2.y - q3.y);
    auto value = [&](int p, double& det) {
        const double xi = P.xi[p], eta = P.eta[p];
        const double J00 = fma(eta, cx, ax), J01 = fma(eta, cy, ay), J10 = fma(xi, cx, bx), J11 = fma(xi, cy, by);  // 4 J
        const double d4 = J00 * J11 - J01 * J10;
        det = 0.0625 * d4;
        if (func == FEMB_FUNC_MASS) return D.rho;
        const double H00 = fma(eta, cu, au), H01 = fma(eta, cv, av), H10 = fma(xi, cu, bu), H11 = fma(xi, cv, bv);  // 4 H[k][s]
        const double r = fast_rcp(d4);
        double u[2][2], sig[3];
        u[0][0] = (J11 * H00 - J01 * H10) * r;
        u[0][1] = (J00 * H10 - J10 * H00) * r;
        u[1][0] = (J11 * H01 - J01 * H11) * r;
        u[1][1] = (J00 * H11 - J10 * H01) * r;
        stresses<false>(D, u, sig);
        return von_mises(D, sig);
    };
    double r = 0.0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        double det;
        const double v = value(p, det);
        if (reduction == FEMB_REDUCE_NONE) out[e * 4 + p] = v;
        else if (reduction == FEMB_REDUCE_SUM || reduction == FEMB_REDUCE_MEAN) r += v;
        else if (reduction == FEMB_REDUCE_INTEGRATE) r = fma(v * det, P.w[p], r);
        else if (reduction == FEMB_REDUCE_MIN) r = p ? fmin(r, v) : v;
        else r = p ? fmax(r, v) : v;
    }
    if (reduction == FEMB_REDUCE_NONE) return;
    if (reduction == FEMB_REDUCE_MEAN) r *= 0.25;
    if (KS) {  // r holds the maximum (KSAgg.py:25-58)
        double s = 0.0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            double det;
            s += exp(100.0 * (value(p, det) - r));
        }
        r += 1.0 / 100.0 * log(0.25 * s);
    }
    out[e] = r;
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
// C ABI: element functions
// =============================================================================================
template <int NN, int NQ>
static void launch_function(const BlockData& b, const DMat& D, const FuncArgs& F, bool nl, int func, int reduction,
                            cudaStream_t st) {
    const Tables<NN, NQ> tb = make_tables<NN, NQ>(b);
    const bool ks = reduction == FEMB_REDUCE_KSMAX;
    if (nl && ks) k_function<NN, NQ, true, true><<<grid_for(F.E, 128), 128, 0, st>>>(tb, D, F, func, reduction);
    else if (nl) k_function<NN, NQ, true, false><<<grid_for(F.E, 128), 128, 0, st>>>(tb, D, F, func, reduction);
    else if (ks) k_function<NN, NQ, false, true><<<grid_for(F.E, 128), 128, 0, st>>>(tb, D, F, func, reduction);
    else k_function<NN, NQ, false, false><<<grid_for(F.E, 128), 128, 0, st>>>(tb, D, F, func, reduction);
}

extern "C" int femb_function_dev(femb_plan* p, const double* d_coords, const double* d_states, const femb_material* mat,
                                 int32_t func, int32_t reduction, double* d_out) {
    TRY(need_final(p, "femb_function_dev"));
    if (!mat || !d_coords || !d_states || !d_out) return fail("femb_function_dev: NULL argument");
    if (reduction < FEMB_REDUCE_NONE || reduction > FEMB_REDUCE_KSMAX)
        return fail("femb_function_dev: unknown reduction id %d", reduction);
    TRY(use_device(p));
    if (p->numDim != 2) return generic_function_dev(p, d_coords, d_states, mat, func, reduction, d_out);
    if (func != FEMB_FUNC_MASS && func != FEMB_FUNC_VON_MISES) return fail("femb_function_dev: unknown function id %d", func);
    const DMat D = make_dmat(*mat);
    int64_t launches = 0;
    for (auto& b : p->blocks) {
        FuncArgs F;
        F.E = b.numElems;
        F.conn = b.d_conn;
        F.coords = (const double2*)d_coords;
        F.states = (const double2*)d_states;
        F.out = d_out + (reduction == FEMB_REDUCE_NONE ? b.pointOffset : b.elemOffset);
        if (b.nn == 4 && b.nq == 4 && b.closedForm && b.d_connE && !mat->nonlinear) {  // standard bilinear quad: closed form
            Q4Points P;
            for (int q = 0; q < 4; ++q) {  // dN_0/dxi = -(1 - eta)/4, dN_0/deta = -(1 - xi)/4
                P.eta[q] = 1.0 + 4.0 * b.dN[(size_t)(q * 2) * 4];
                P.xi[q] = 1.0 + 4.0 * b.dN[(size_t)(q * 2 + 1) * 4];
                P.w[q] = b.w[q];
            }
            if (reduction == FEMB_REDUCE_KSMAX)
                k_function_q4<true><<<grid_for(F.E, 128), 128, 0, p->stream>>>(P, D, F.E, b.d_connE, F.coords, F.states, F.out, func, reduction);
            else
                k_function_q4<false><<<grid_for(F.E, 128), 128, 0, p->stream>>>(P, D, F.E, b.d_connE, F.coords, F.states, F.out, func, reduction);
            launches++;
            continue;
        }
        FEMB_DISPATCH_ANY(b.nn, b.nq, (launch_function<NN, NQ>(b, D, F, mat->nonlinear != 0, func, reduction, p->stream)));
        launches++;
    }
    p->launches = launches;
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int femb_function(femb_plan* p, const double* coords, const double* states, const femb_material* mat,
                             int32_t func, int32_t reduction, double* out) {
    TRY(need_final(p, "femb_function"));
    if (!states || !out) return fail("femb_function: NULL argument");
    TRY(use_device(p));
    const int64_t numDOF = femb_plan_num_dof(p);
    const size_t nout = (size_t)(reduction == FEMB_REDUCE_NONE ? p->numPoints : p->numElems);
    TRY(stage_geometry(p, coords, nullptr, false, "femb_function"));
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
// point quantities: Element._computeFunctionEvaluationQuantities (FEMpy/Elements/Element.py:275-368) at the block's points
//   Coord / State   sum_a N_a(p) X_a, sum_a N_a(p) q_a                (_interpolationProduct, Element.py:918-939)
//   Jac             J[i][j] = sum_a dN_a/dxi_i X_a[j]                 (_computeNPrimeCoordProduct, :942-978)
//   JacDet          signed determinant                                (LinAlg.py:70-90)
//   StateGrad       u'[s][d] = sum_a (J^-1 dN_a)[d] q_a[s]            (_computeUPrimeProduct, :981-1018)
//   StateGradSens   (J^-1 dN)[d][a]                                   (_computeDUPrimeDqProduct, :1021-1048)
// One thread per (element, point); node count and point count are run-time values (tables in global memory), so any
// family and any rule -- or any set of parametric points -- runs through this one kernel.
// =============================================================================================
struct PointArgs {
    long long E;
    int nn, nq;
    const int* conn;         // node-major
    const double* N;         // (nq, nn)
    const double* dN;        // (nq, 2, nn)
    const double2* coords;
    const double2* states;
    double *coord, *state, *grad, *jac, *det, *sens;  // (E*nq, ...) outputs of this block, any may be NULL
};

__global__ void __launch_bounds__(128) k_point_quantities(const PointArgs P) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.E * P.nq) return;
    const long long e = t / P.nq;
    const int p = (int)(t - e * P.nq);
    const double* Np = P.N + (size_t)p * P.nn;
    const double* d0 = P.dN + (size_t)p * 2 * P.nn;
    const double* d1 = d0 + P.nn;
    double cx = 0.0, cy = 0.0, sx = 0.0, sy = 0.0, J00 = 0.0, J01 = 0.0, J10 = 0.0, J11 = 0.0, H00 = 0.0, H01 = 0.0, H10 = 0.0, H11 = 0.0;
    for (int a = 0; a < P.nn; ++a) {
        const int node = __ldg(P.conn + (long long)a * P.E + e);
        const double2 X = __ldg(P.coords + node), q = __ldg(P.states + node);
        const double n = __ldg(Np + a), g0 = __ldg(d0 + a), g1 = __ldg(d1 + a);
        cx = fma(n, X.x, cx), cy = fma(n, X.y, cy), sx = fma(n, q.x, sx), sy = fma(n, q.y, sy);
        J00 = fma(g0, X.x, J00), J01 = fma(g0, X.y, J01), J10 = fma(g1, X.x, J10), J11 = fma(g1, X.y, J11);
        H00 = fma(g0, q.x, H00), H01 = fma(g0, q.y, H01), H10 = fma(g1, q.x, H10), H11 = fma(g1, q.y, H11);
    }
    const double det = J00 * J11 - J01 * J10;
    const double r = 1.0 / det;
    const double i00 = J11 * r, i01 = -J01 * r, i10 = -J10 * r, i11 = J00 * r;  // J^-1
    if (P.coord) P.coord[2 * t] = cx, P.coord[2 * t + 1] = cy;
    if (P.state) P.state[2 * t] = sx, P.state[2 * t + 1] = sy;
    if (P.jac) P.jac[4 * t] = J00, P.jac[4 * t + 1] = J01, P.jac[4 * t + 2] = J10, P.jac[4 * t + 3] = J11;
    if (P.det) P.det[t] = det;
    if (P.grad) {  // u'[s][d] = sum_k Jinv[d][k] H[k][s]
        P.grad[4 * t] = i00 * H00 + i01 * H10;
        P.grad[4 * t + 1] = i10 * H00 + i11 * H10;
        P.grad[4 * t + 2] = i00 * H01 + i01 * H11;
        P.grad[4 * t + 3] = i10 * H01 + i11 * H11;
    }
    if (P.sens) {
        double* s = P.sens + (size_t)t * 2 * P.nn;
        for (int a = 0; a < P.nn; ++a) {
            const double g0 = __ldg(d0 + a), g1 = __ldg(d1 + a);
            s[a] = i00 * g0 + i01 * g1;
            s[P.nn + a] = i10 * g0 + i11 * g1;
        }
    }
}

extern "C" int femb_point_quantities(femb_plan* p, const double* coords, const double* states, double* coord, double* state,
                                     double* stateGrad, double* jac, double* jacDet, double* stateGradSens) {
    TRY(need_final(p, "femb_point_quantities"));
    if (!states) return fail("femb_point_quantities: states is NULL");
    if (!coord && !state && !stateGrad && !jac && !jacDet && !stateGradSens) return fail("femb_point_quantities: nothing requested");
    if (p->numDim != 2) return fail("femb_point_quantities: 2-D plans only");
    TRY(use_device(p));
    cudaStream_t st = p->stream;
    TRY(stage_geometry(p, coords, nullptr, false, "femb_point_quantities"));
    TRY(stage(&p->s_states, states, (size_t)femb_plan_num_dof(p), st));
    const size_t np = (size_t)p->numPoints;
    size_t sensTotal = 0;
    for (auto& b : p->blocks) sensTotal += (size_t)b.numElems * b.nq * 2 * b.nn;
    Scratch<double> d_coord, d_state, d_grad, d_jac, d_det, d_sens, d_tab;
    if (coord) CUDA_TRY(d_coord.alloc(2 * np));
    if (state) CUDA_TRY(d_state.alloc(2 * np));
    if (stateGrad) CUDA_TRY(d_grad.alloc(4 * np));
    if (jac) CUDA_TRY(d_jac.alloc(4 * np));
    if (jacDet) CUDA_TRY(d_det.alloc(np));
    if (stateGradSens) CUDA_TRY(d_sens.alloc(sensTotal));
    size_t tabTotal = 0;
    for (auto& b : p->blocks) tabTotal += b.N.size() + b.dN.size();
    CUDA_TRY(d_tab.alloc(tabTotal));
    size_t tabOff = 0, sensOff = 0;
    int64_t launches = 0;
    for (auto& b : p->blocks) {
        CUDA_TRY(cudaMemcpyAsync(d_tab.p + tabOff, b.N.data(), b.N.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(d_tab.p + tabOff + b.N.size(), b.dN.data(), b.dN.size() * sizeof(double), cudaMemcpyHostToDevice, st));
        PointArgs A;
        A.E = b.numElems;
        A.nn = b.nn;
        A.nq = b.nq;
        A.conn = b.d_conn;
        A.N = d_tab.p + tabOff;
        A.dN = d_tab.p + tabOff + b.N.size();
        A.coords = (const double2*)p->s_coords;
        A.states = (const double2*)p->s_states;
        const size_t po = (size_t)b.pointOffset;
        A.coord = coord ? d_coord.p + 2 * po : nullptr;
        A.state = state ? d_state.p + 2 * po : nullptr;
        A.grad = stateGrad ? d_grad.p + 4 * po : nullptr;
        A.jac = jac ? d_jac.p + 4 * po : nullptr;
        A.det = jacDet ? d_det.p + po : nullptr;
        A.sens = stateGradSens ? d_sens.p + sensOff : nullptr;
        k_point_quantities<<<grid_for(b.numElems * b.nq, 128), 128, 0, st>>>(A);
        tabOff += b.N.size() + b.dN.size();
        sensOff += (size_t)b.numElems * b.nq * 2 * b.nn;
        launches++;
    }
    p->launches = launches;
    if (coord) CUDA_TRY(cudaMemcpyAsync(coord, d_coord, 2 * np * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (state) CUDA_TRY(cudaMemcpyAsync(state, d_state, 2 * np * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (stateGrad) CUDA_TRY(cudaMemcpyAsync(stateGrad, d_grad, 4 * np * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (jac) CUDA_TRY(cudaMemcpyAsync(jac, d_jac, 4 * np * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (jacDet) CUDA_TRY(cudaMemcpyAsync(jacDet, d_det, np * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (stateGradSens) CUDA_TRY(cudaMemcpyAsync(stateGradSens, d_sens, sensTotal * sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
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
    TRY(stage(&p->s_coords, coords, (size_t)p->numNodes * p->numDim, p->stream));
    TRY(stage(&p->s_states, states, (size_t)numDOF, p->stream));
    TRY(stage(&p->s_thick, thickness, (size_t)p->numElems, p->stream));
    Scratch<double> d_Ke, d_Re;
    if (Ke) CUDA_TRY(d_Ke.alloc(p->keTotal));
    if (Re) CUDA_TRY(d_Re.alloc(p->reTotal));
    if (p->numDim != 2) {
        TRY(generic_element_blocks_dev(p, p->s_coords, p->s_states, p->s_thick, mat, d_Ke.p, d_Re.p));
        if (Ke) CUDA_TRY(cudaMemcpyAsync(Ke, d_Ke, p->keTotal * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
        if (Re) CUDA_TRY(cudaMemcpyAsync(Re, d_Re, p->reTotal * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
        CUDA_TRY(cudaStreamSynchronize(p->stream));
        return 0;
    }
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
        FEMB_DISPATCH_ANY(b.nn, b.nq, (launch_blocks<NN, NQ>(b, D, A, mat->nonlinear != 0, d_Ke.p ? d_Ke.p + b.keOffset : nullptr,
                                                               d_Re.p ? d_Re.p + b.reOffset : nullptr, p->stream)));
        launches++;
    }
    p->launches = launches;
    int rc = 0;
    if (Ke && cudaMemcpyAsync(Ke, d_Ke, p->keTotal * sizeof(double), cudaMemcpyDeviceToHost, p->stream) != cudaSuccess) rc = 1;
    if (Re && cudaMemcpyAsync(Re, d_Re, p->reTotal * sizeof(double), cudaMemcpyDeviceToHost, p->stream) != cudaSuccess) rc = 1;
    if (cudaStreamSynchronize(p->stream) != cudaSuccess) rc = 1;
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
    Scratch<double> d_vals, d_out;
    Scratch<long long> d_dof, d_rows, d_cols;
    Scratch<int> d_flags, d_pos;
    CUDA_TRY(d_vals.alloc(n));
    CUDA_TRY(d_out.alloc(n));
    CUDA_TRY(d_dof.alloc(numElems * nd));
    CUDA_TRY(d_rows.alloc(n));
    CUDA_TRY(d_cols.alloc(n));
    CUDA_TRY(d_flags.alloc(n));
    CUDA_TRY(d_pos.alloc(n + 1));
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
    Scratch<double> d_res, d_R;
    Scratch<long long> d_conn;
    Scratch<int> d_flag;
    CUDA_TRY(d_res.alloc(n * numStates));
    CUDA_TRY(d_R.alloc(nd));
    CUDA_TRY(d_conn.alloc(n));
    CUDA_TRY(d_flag.alloc(1));
    CUDA_TRY(cudaMemset(d_flag, 0, sizeof(int)));
    CUDA_TRY(cudaMemcpy(d_res, localRes, n * numStates * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_R, globalResidual, nd * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_conn, conn, n * sizeof(long long), cudaMemcpyHostToDevice));
    if (n > 0) k_scatter_residuals<<<grid_for(n, 256), 256>>>(d_res, d_conn, n, numStates, numNodes, d_R, d_flag);
    int flag = 0;
    CUDA_TRY(cudaMemcpy(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(globalResidual, d_R, nd * sizeof(double), cudaMemcpyDeviceToHost));
    if (flag) return fail("femb_scatter_local_residuals: connectivity out of range");
    return 0;
}
