// Row-owner assembly (sm_100a): one thread per node row, no atomics, no barriers, no inter-thread accumulation.
//
// Thread i owns the two DOF rows of node i.  It walks the node's (sorted) element adjacency; for every adjacent
// element it re-evaluates the element geometry at the Gauss points and forms ONLY the node-row strip of K_e that
// belongs to node i (4 node-pair blocks for a quad) plus its residual pair, and adds them into a private row of
// accumulators in shared memory.  The element geometry is therefore evaluated once per (node, element) corner --
// about 2.7x the minimal FP64 work -- which buys: independent threads at high occupancy (the thread never holds a
// whole K_e), a fixed summation order (bit-reproducible), and every CSR value / residual entry written exactly once.
//
// Shared-memory accumulators: acc[m][t] (double2), m = k for DOF row 0 and deg + k for DOF row 1 of the thread's
// node, stored at m*BLOCK + ((t + m) & (BLOCK-1)).  The swizzle keeps the read-modify-writes conflict-free when the
// lanes of a warp use the same slot pattern (structured meshes) and makes the final row-by-row copy conflict-free:
// the lanes of a warp then stream one node's 32*deg-byte run (both DOF rows are contiguous in the CSR values) per step.
#pragma once
#include "femb_internal.h"

namespace rows {

#ifndef ROWS_KNOCK
#define ROWS_KNOCK 0  // measurement builds only (tools/knock_variants.sh): bit mask of kernel stages replaced by stubs
#endif
#ifndef ROWS_TRACE
#define ROWS_TRACE 0  // measurement builds only: warp lifetime statistics into Args::stats
#endif
static constexpr int KNOCK = ROWS_KNOCK;  // 1 zero-fill, 2 node-data gathers, 4 corner/connectivity loads, 8 write-out, 16 one Gauss point,
                                          // 32 accumulator read-modify-writes, 64 D application / residual product
static constexpr int POINT_UNROLL = 1;  // Gauss-point loop kept rolled: fewer live registers, 20+ warps per SM
static constexpr int BLOCK = 32;
static constexpr int MAXD = 48;  // largest node degree (neighbours incl. itself): shared-memory rows; the corner records hold 7-bit positions

struct Args {
    long long numNodes;
    const int4* slotRec;     // balanced schedule (PERM kernels): {node or -1, first value of the row, degree | corners << 8, first corner}
    const int* nrowptr;
    const int* ncols;
    const int* cBegin;       // per node: first / one-past-last corner of this element block (global corner numbers)
    const int* cEnd;
    int elemOffset;          // first global element id of the block (thickness is indexed by global element id)
    int accumulate;          // 1: K and R already hold the contributions of earlier element blocks
    int finalise;            // 1: last block -- apply loads and boundary conditions
    const void* corners;     // corner records (plan.cu: k_corners)
    const int* cornerElem;   // rotated 4-node records: element id of each corner
    const int4* connE;       // element-major connectivity, ceil(NN/4) words per element
    const double2* coords;
    const double2* states;
    const double* thick;
    const double2* loads;           // may be NULL
    const int* bcInfo;              // NULL: no BCs; per node (index of the constrained node + 1) << 9 | position in its own row << 2 | DOF bits
    const double2* bcNodeCnt;       // per constrained node: diagonal values (occurrence counts) of DOF 0 / DOF 1
    const double2* bcNodeVal;       // per constrained node: prescribed values
    int maxDeg;
    long long* stats;        // measurement builds only (ROWS_TRACE, tools/rows_trace.py): warp lifetimes
    int zero;                // always 0: a runtime-uniform index that keeps constants out of FP64 operand slots
    double* K;
    double* R;
};

__device__ __forceinline__ void st_stream(double2* p, const double2 v) {  // K is written once and not re-read here
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y));
}

// resident threads per SM the register allocation aims at: 20 warps for the 3/4-node families (<= 96 registers),
// 12 for 6 nodes, 8 for 9/10 nodes (the node-row strip alone is 4*NN doubles)
template <int NN, bool NL>
__host__ __device__ constexpr int resident_threads() {
#ifndef ROWS_T3_THREADS
#define ROWS_T3_THREADS 640
#endif
    return NN == 3 && !NL ? ROWS_T3_THREADS : NN <= 4 ? (NL ? 512 : 640) : (NN <= 6 ? 384 : 256);
}

// MB: the plan has several element blocks (one launch per block: later blocks add to what earlier ones left in K and R,
// the last one applies loads and boundary conditions); single-block plans compile those switches away
// PERM: the thread's node comes from the plan's balanced row schedule (nodes of a window ordered by corner count) instead
// of the thread index; the warp's rows are then 32 separate runs of the CSR values instead of one
// CF: closed-form bilinear quad (4 nodes, 2x2 Gauss rule, small strain; plan.cu checks the block's tables): the geometry of
// a corner is 12 coordinate differences and four reciprocals, the Gauss-point sum is carried by four moments of w*theta/det
template <int NN, int NQ, bool NL, bool WK, bool WR, bool MB, bool PERM, bool CF>
__global__ void __launch_bounds__(BLOCK, resident_threads<NN, NL>() / BLOCK) k_assemble_rows(const __grid_constant__ femb::Tables<NN, NQ> tb, const __grid_constant__ femb::DMat D,
                                                        const Args A) {
    using namespace femb;
    static_assert(NN <= 12, "corner records hold at most 12 slot positions");
    static_assert(!CF || (NN == 4 && NQ == 4 && !NL && FEMB_ROT4), "closed form: rotated bilinear quads, small strain");
    constexpr int CW = (NN + 3) / 4;  // connectivity words per element
    extern __shared__ double2 acc[];  // [2*maxDeg][BLOCK], swizzled
    const bool accumulate = MB && A.accumulate, finalise = !MB || A.finalise;
    const int elemOffset = MB ? A.elemOffset : 0;
#if ROWS_TRACE
    long long tStart = clock64(), gStart;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gStart));
#endif
    const int tid = threadIdx.x, lane = tid & 31;
    const long long slot0 = (long long)blockIdx.x * BLOCK + tid;
    int4 srec = make_int4(0, 0, 0, 0);
    if (PERM) srec = __ldg(A.slotRec + slot0);
    const long long i = PERM ? (long long)srec.x : slot0;
    const bool valid = PERM ? i >= 0 : i < A.numNodes;
    const int off = PERM ? srec.y : __ldg(A.nrowptr + (valid ? i : A.numNodes));  // tail lanes: end of the values, degree 0
    const int deg = PERM ? (srec.z & 0xff) : valid ? __ldg(A.nrowptr + i + 1) - off : 0;
    auto slot = [&](int m) -> double2& { return acc[m * BLOCK + ((tid + m) & (BLOCK - 1))]; };
    // loads the kernel waits on later are issued here: the corner range (start of the corner loop) and the BC word (epilogue)
    // (balanced single-block plans carry the range in the slot record)
    const int c0 = (PERM && !MB) ? srec.w : valid ? __ldg(A.cBegin + i) : 0;
    const int c1 = (PERM && !MB) ? srec.w + (int)((unsigned)srec.z >> 8) : valid ? __ldg(A.cEnd + i) : 0;
    const unsigned info = (valid && A.bcInfo && finalise) ? (unsigned)__ldg(A.bcInfo + i) : 0u;

    // The rows of the warp's 32 consecutive nodes are one contiguous run of the CSR values (both DOF rows of a node
    // are adjacent and rows are stored in node order): [2*off_first, 2*off_last + 2*deg_last) in double2 units.  A byte
    // table maps a position of the run to the lane that owns it; the warp streams the run in 512-byte steps when it
    // writes the rows out (and when it reads back what earlier element blocks left there).
    // (base / start / total are recomputed at the write-out rather than carried across the corner loop)
    double2* K2 = reinterpret_cast<double2*>(A.K);
    const int wbase = tid - lane;
    unsigned char* owner = reinterpret_cast<unsigned char*>(acc + 2 * A.maxDeg * BLOCK + BLOCK) + 2 * A.maxDeg * wbase;
    double2& racc = acc[2 * A.maxDeg * BLOCK + tid];  // residual pair of the node: kept in shared memory, not in registers
    // my first position in the warp's run of values (PERM: in the concatenation of the 32 separate runs)
    int ownStart = 0;
    if (WK && !(KNOCK & 8)) {
        if (PERM) {
            int incl = 2 * deg;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += up;
            }
            ownStart = incl - 2 * deg;
        } else {
            ownStart = 2 * (off - __shfl_sync(0xffffffffu, off, 0));
        }
    }
    const int ownGlobal = 2 * off - ownStart;  // PERM: value position (double2 units) = ownGlobal of the owning lane + run position
    // owner table + zeroed accumulators: thread-private stores, so they run under the first corner's loads
    auto fill = [&]() {
        if (WK && !(KNOCK & 8))
            for (int m = 0; m < 2 * deg; ++m) owner[ownStart + m] = (unsigned char)lane;
        if (WK && !(KNOCK & 1))
            for (int m = 0; m < 2 * deg; ++m) slot(m) = make_double2(0.0, 0.0);
        if (WR) racc = make_double2(0.0, 0.0);
    };
    bool fillPending = !accumulate;
    if (accumulate) {
        if (WK && !(KNOCK & 8))
            for (int m = 0; m < 2 * deg; ++m) owner[ownStart + m] = (unsigned char)lane;
        if (WK) {
            const int base = __shfl_sync(0xffffffffu, off, 0);
            const int start = ownStart;
            const int total = __shfl_sync(0xffffffffu, start + 2 * deg, 31);  // invalid tail lanes hold off = end, deg = 0
            __syncwarp();
            for (int f0 = 0; f0 < total; f0 += 32) {
                const int f = f0 + lane;
                const bool ok = f < total;
                const int r = ok ? owner[f] : 0;
                const int mm = f - __shfl_sync(0xffffffffu, start, r);
                const long long g = PERM ? (long long)__shfl_sync(0xffffffffu, ownGlobal, r) + f : 2ll * base + f;
                if (ok) acc[mm * BLOCK + ((wbase + r + mm) & (BLOCK - 1))] = K2[g];
            }
            __syncwarp();
        }
        if (WR) racc = valid ? reinterpret_cast<const double2*>(A.R)[i] : make_double2(0.0, 0.0);
    }

#ifndef ROWS_VAR
#define ROWS_VAR 1  // bit 1: wide families load the next corner's record and connectivity one iteration ahead
#endif
    // Wide families (6-10 nodes) run at 8-12 warps per SM with registers to spare: the next corner's record and
    // connectivity words are fetched one iteration ahead, which takes two of the three dependent loads
    // (record -> connectivity -> coordinates) off the critical path.
    constexpr bool AHEAD = NN > 4 && (ROWS_VAR & 1);
    // 3-node triangles: the corner record holds the element rotated so that the owner is local node 0 and names the other
    // two nodes (plan.cu: k_corners) -- no connectivity load, the owner's coordinates / state are read once per thread
    constexpr bool ROT = NN == 3;
    // 4-node quads, same idea ({slots, the three other nodes} + element id aside); the owner's coordinates / state are
    // re-read per corner (one coalesced request per warp) because the 96-register budget has no room to keep them
    constexpr bool ROT4 = NN == 4 && FEMB_ROT4;
    double2 ownX = make_double2(0.0, 0.0), ownU = make_double2(0.0, 0.0);
    if ((ROT || CF) && valid) {
        ownX = __ldg(A.coords + i);
        ownU = __ldg(A.states + i);
    }
    auto load_rec = [&](int c) -> int4 {
        if constexpr (NN <= 4) {
            const int2 r = __ldg(reinterpret_cast<const int2*>(A.corners) + c);
            return make_int4(r.x, r.y, 0, 0);
        } else {
            return __ldg(reinterpret_cast<const int4*>(A.corners) + c);
        }
    };
    constexpr int EMASK = NN <= 4 ? 0x7fffffff : 0x0fffffff;
    int4 recAhead = make_int4(0, 0, 0, 0), connAhead[CW];
    if (AHEAD && c0 < c1) {
        recAhead = load_rec(c0);
#pragma unroll
        for (int w = 0; w < CW; ++w) connAhead[w] = __ldg(A.connE + (long long)(recAhead.x & EMASK) * CW + w);
    }
    for (int c = c0; c < c1; ++c) {
        int e, a, node[4 * CW];
        unsigned kw[CW];  // slot positions: 7-bit fields from bit 4 (NN <= 4) or bytes
        if constexpr (ROT) {
            const int4 rec = __ldg(reinterpret_cast<const int4*>(A.corners) + c);
            e = rec.x;
            kw[0] = (unsigned)rec.y;
            a = 0;
            node[0] = (int)i;
            node[1] = rec.z;
            node[2] = rec.w;
            node[3] = -1;
        } else if constexpr (ROT4) {
            const int4 rec = __ldg(reinterpret_cast<const int4*>(A.corners) + c);
            e = __ldg(A.cornerElem + c);
            kw[0] = (unsigned)rec.x << 4;  // same field positions as the unrotated record (a = 0 in the low bits)
            a = 0;
            node[0] = (int)i;
            node[1] = rec.y;
            node[2] = rec.z;
            node[3] = rec.w;
        } else if constexpr (NN <= 4) {
            const int2 rec = AHEAD ? make_int2(recAhead.x, recAhead.y)
                             : (KNOCK & 4) ? make_int2(c & 0xffff, ((c - c0) & 3) | (1 << 11) | (2 << 18) | (3 << 25))
                                           : __ldg(reinterpret_cast<const int2*>(A.corners) + c);
            e = rec.x;
            kw[0] = (unsigned)rec.y;
            a = rec.y & 0xf;
        } else {
            const int4 rec = AHEAD ? recAhead : __ldg(reinterpret_cast<const int4*>(A.corners) + c);
            e = rec.x & 0x0fffffff;
            a = (int)((unsigned)rec.x >> 28);
            kw[0] = (unsigned)rec.y;
            if (CW > 1) kw[1] = (unsigned)rec.z;
            if (CW > 2) kw[CW - 1] = (unsigned)rec.w;
        }
#pragma unroll
        for (int w = 0; w < ((ROT || ROT4) ? 0 : CW); ++w) {
            const int4 cn = AHEAD ? connAhead[w] : (KNOCK & 4) ? make_int4((int)i, (int)i + 1, (int)i + 2, (int)i + 3) : __ldg(A.connE + (long long)e * CW + w);
            node[4 * w] = cn.x;
            node[4 * w + 1] = cn.y;
            node[4 * w + 2] = cn.z;
            node[4 * w + 3] = cn.w;
        }
        if (AHEAD && c + 1 < c1) {
            recAhead = load_rec(c + 1);
#pragma unroll
            for (int w = 0; w < CW; ++w) connAhead[w] = __ldg(A.connE + (long long)(recAhead.x & EMASK) * CW + w);
        }
        double Kr[NN][4];  // node-row strip of K_e: blocks (a, b), b = 0..NN-1 -- geometric sums S[d][D] until D is applied
        double ra0 = 0.0, ra1 = 0.0;
        if constexpr (CF) {
            // Closed form of the bilinear quad with the owner as local node 0 (nodes counter-clockwise).  With
            //   8 det J dN_b/dx = C0_b + xi X0_b + eta E0_b,   8 det J dN_b/dy = C1_b + xi X1_b + eta E1_b
            // (every coefficient is one difference of two node coordinates), 8 det J = d0 + xi dx + eta de and the 2x2
            // Gauss rule (xi, eta = +-g, w = 1), the geometric sums over the four points are
            //   S[d][D]_0b = P_d C_Db + Q_d X_Db + R_d E_Db,
            // P, Q, R built from the moments m0, m_xi, m_eta, m_xi_eta of s_p = theta / (8 det'_p)  (sum_p s_p xi^2 = m0 / 3).
            // (FEMpy/Elements/Element.py:942-1111,1147-1168 evaluate the same integrals point by point.)
            const double2 X1 = __ldg(A.coords + node[1]), X2 = __ldg(A.coords + node[2]), X3 = __ldg(A.coords + node[3]);
            const double th8 = 0.125 * __ldg(A.thick + elemOffset + e);
            if (fillPending) {
                fill();
                fillPending = false;
            }
            const double ya = X1.y - X3.y, yb = X2.y - ownX.y, yc = X3.y - X2.y, yd = ownX.y - X1.y, ye = X2.y - X1.y, yf = ownX.y - X3.y;
            const double xa = X3.x - X1.x, xb = ownX.x - X2.x, xc = X2.x - X3.x, xd = X1.x - ownX.x, xe = X1.x - X2.x, xf = X3.x - ownX.x;
            constexpr double G = 0.57735026918962576451, THIRD = 1.0 / 3.0;
            const double d0 = xb * ya - xa * yb;
            const double gx = G * (xc * yd - xd * yc), ge = G * (xe * yf - xf * ye);
            const double tp = d0 + gx, tm = d0 - gx;
            const double spp = th8 * fast_rcp(tp + ge), spm = th8 * fast_rcp(tp - ge);
            const double smp = th8 * fast_rcp(tm + ge), smm = th8 * fast_rcp(tm - ge);
            const double up = spp + spm, um = smp + smm, vp = spp - spm, vm = smp - smm;
            const double m0 = up + um, mx = G * (up - um), me = G * (vp + vm), mxe = THIRD * (vp - vm), m3 = THIRD * m0;
            const double P0 = fma(m0, ya, fma(mx, yc, me * ye)), Q0 = fma(mx, ya, fma(m3, yc, mxe * ye)), R0 = fma(me, ya, fma(mxe, yc, m3 * ye));
            const double P1 = fma(m0, xa, fma(mx, xc, me * xe)), Q1 = fma(mx, xa, fma(m3, xc, mxe * xe)), R1 = fma(me, xa, fma(mxe, xc, m3 * xe));
            if (WK || WR) {
                // b:        0            1            2            3
                // C_Db:  ya | xa      yb | xb     -ya | -xa    -yb | -xb
                // X_Db:  yc | xc     -yc | -xc     yd | xd     -yd | -xd
                // E_Db:  ye | xe      yf | xf     -yf | -xf    -ye | -xe
                Kr[0][0] = fma(P0, ya, fma(Q0, yc, R0 * ye));
                Kr[0][1] = fma(P0, xa, fma(Q0, xc, R0 * xe));
                Kr[0][2] = fma(P1, ya, fma(Q1, yc, R1 * ye));
                Kr[0][3] = fma(P1, xa, fma(Q1, xc, R1 * xe));
                Kr[1][0] = fma(P0, yb, fma(R0, yf, -(Q0 * yc)));
                Kr[1][1] = fma(P0, xb, fma(R0, xf, -(Q0 * xc)));
                Kr[1][2] = fma(P1, yb, fma(R1, yf, -(Q1 * yc)));
                Kr[1][3] = fma(P1, xb, fma(R1, xf, -(Q1 * xc)));
                Kr[2][0] = fma(Q0, yd, -fma(P0, ya, R0 * yf));
                Kr[2][1] = fma(Q0, xd, -fma(P0, xa, R0 * xf));
                Kr[2][2] = fma(Q1, yd, -fma(P1, ya, R1 * yf));
                Kr[2][3] = fma(Q1, xd, -fma(P1, xa, R1 * xf));
                Kr[3][0] = -fma(P0, yb, fma(Q0, yd, R0 * ye));
                Kr[3][1] = -fma(P0, xb, fma(Q0, xd, R0 * xe));
                Kr[3][2] = -fma(P1, yb, fma(Q1, yd, R1 * ye));
                Kr[3][3] = -fma(P1, xb, fma(Q1, xd, R1 * xe));
            }
        } else {
            double x[NN], y[NN], ux[NN], uy[NN];
#pragma unroll
            for (int b = 0; b < NN; ++b) {
                const double2 X = (ROT && b == 0) ? ownX
                                  : (KNOCK & 2) ? make_double2(0.01 * (node[b] & 511) + ((b == 1 || b == 2) ? 0.01 : 0.0), 1e-5 * node[b] + (b >= 2 ? 0.01 : 0.0))
                                                : __ldg(A.coords + node[b]);
                x[b] = X.x;
                y[b] = X.y;
                if (NL) {
                    const double2 U = (ROT && b == 0) ? ownU : __ldg(A.states + node[b]);
                    ux[b] = U.x;
                    uy[b] = U.y;
                } else {  // linear model: the states are fetched after the strip (short live range)
                    ux[b] = uy[b] = 0.0;
                }
            }
            const double theta = (KNOCK & 2) ? 1.0 + 1e-9 * e : __ldg(A.thick + elemOffset + e);
            if (fillPending) {
                fill();
                fillPending = false;
            }

            // node-row strip of K_e: blocks (a, b), b = 0..NN-1, and the residual pair of node a
#pragma unroll
            for (int b = 0; b < NN; ++b) Kr[b][0] = Kr[b][1] = Kr[b][2] = Kr[b][3] = 0.0;
#pragma unroll POINT_UNROLL
            for (int p = 0; p < ((KNOCK & 16) ? 1 : NQ); ++p) {
                if (!NL) {
                    // Linear model, division-free gradients: with adj = det * J^-1 and Gh = adj * dN (so G = Gh / det),
                    //   sc * G_a G_b = (w * theta / det) * Gh_a Gh_b,
                    // so only the scalar w*theta/det carries the reciprocal.  The gradient of "my" shape function comes
                    // from the table with a per-thread index (constant-bank load) instead of register selects, and the
                    // tables are read through uniform registers (rolled loop): FP64 instructions with a constant-bank
                    // operand issue at half rate on sm_100 (tools/micro/fp64_micro2.cu).
                    double J00 = 0.0, J01 = 0.0, J10 = 0.0, J11 = 0.0;
#pragma unroll
                    for (int b = 0; b < NN; ++b) {
                        J00 = fma(tb.dN[p][0][b], x[b], J00);
                        J01 = fma(tb.dN[p][0][b], y[b], J01);
                        J10 = fma(tb.dN[p][1][b], x[b], J10);
                        J11 = fma(tb.dN[p][1][b], y[b], J11);
                    }
                    const double det = J00 * J11 - J01 * J10;
                    const double sc = tb.w[p] * theta * fast_rcp(det);
                    const double da0 = tb.dN[p][0][a], da1 = tb.dN[p][1][a];
                    const double A0 = (J11 * da0 - J01 * da1) * sc, A1 = (J00 * da1 - J10 * da0) * sc;
                    if (WK || WR) {
#pragma unroll
                        for (int b = 0; b < NN; ++b) {
                            const double g0 = J11 * tb.dN[p][0][b] - J01 * tb.dN[p][1][b];
                            const double g1 = J00 * tb.dN[p][1][b] - J10 * tb.dN[p][0][b];
                            Kr[b][0] = fma(A0, g0, Kr[b][0]);
                            Kr[b][1] = fma(A0, g1, Kr[b][1]);
                            Kr[b][2] = fma(A1, g0, Kr[b][2]);
                            Kr[b][3] = fma(A1, g1, Kr[b][3]);
                        }
                    }
                    continue;
                }
                double G0[NN], G1[NN];
                const double det = point_gradients<NN, NQ>(tb, p, x, y, G0, G1);
                const double sc = tb.w[p] * det * theta;
                double g0 = 0.0, g1 = 0.0;  // gradient of "my" shape function: runtime index -> selects, no dynamic registers
#pragma unroll
                for (int b = 0; b < NN; ++b) {
                    g0 = (b == a) ? G0[b] : g0;
                    g1 = (b == a) ? G1[b] : g1;
                }
                if (!NL) {
                    // linear model: accumulate the geometric sums sc * G_d,a * G_D,b; D is applied once below
                    if (WK || WR) {
                        const double A0 = g0 * sc, A1 = g1 * sc;
#pragma unroll
                        for (int b = 0; b < NN; ++b) {
                            Kr[b][0] = fma(A0, G0[b], Kr[b][0]);
                            Kr[b][1] = fma(A0, G1[b], Kr[b][1]);
                            Kr[b][2] = fma(A1, G0[b], Kr[b][2]);
                            Kr[b][3] = fma(A1, G1[b], Kr[b][3]);
                        }
                    }
                } else {
                    double u[2][2];
                    state_gradient<NN>(G0, G1, ux, uy, u);
                    double Ba[3][2];
                    bmat<true>(g0, g1, u, Ba);
                    if (WR) {
                        double sig[3];
                        stresses<true>(D, u, sig);
                        ra0 = fma(sc, Ba[0][0] * sig[0] + Ba[1][0] * sig[1] + Ba[2][0] * sig[2], ra0);
                        ra1 = fma(sc, Ba[0][1] * sig[0] + Ba[1][1] * sig[1] + Ba[2][1] * sig[2], ra1);
                    }
                    if (WK) {
                        double DBa[3][2];
#pragma unroll
                        for (int s = 0; s < 2; ++s) {
                            DBa[0][s] = sc * (D.d00 * Ba[0][s] + D.d01 * Ba[1][s]);
                            DBa[1][s] = sc * (D.d01 * Ba[0][s] + D.d11 * Ba[1][s]);
                            DBa[2][s] = sc * (D.d22 * Ba[2][s]);
                        }
#pragma unroll
                        for (int b = 0; b < NN; ++b) {
                            double Bb[3][2];
                            bmat<true>(G0[b], G1[b], u, Bb);
#pragma unroll
                            for (int s = 0; s < 2; ++s)
#pragma unroll
                                for (int S = 0; S < 2; ++S)
                                    Kr[b][2 * s + S] += DBa[0][s] * Bb[0][S] + DBa[1][s] * Bb[1][S] + DBa[2][s] * Bb[2][S];
                        }
                    }
                }
            }
        }
        if (!NL && !(KNOCK & 64)) {  // [[D00 S00 + D22 S11, D01 S01 + D22 S10], [D01 S10 + D22 S01, D11 S11 + D22 S00]]
            const double* Dv = &D.d00 + A.zero;  // runtime-uniform offset (0): the entries arrive in uniform registers
            const double d00 = Dv[0], d01 = Dv[1], d11 = Dv[2], d22 = Dv[3];
#pragma unroll
            for (int b = 0; b < NN; ++b) {
                const double s00 = Kr[b][0], s01 = Kr[b][1], s10 = Kr[b][2], s11 = Kr[b][3];
                Kr[b][0] = fma(d00, s00, d22 * s11);
                Kr[b][1] = fma(d01, s01, d22 * s10);
                Kr[b][2] = fma(d01, s10, d22 * s01);
                Kr[b][3] = fma(d11, s11, d22 * s00);
                if (WR) {  // R_a = sum_b K_ab q_b
                    const double2 q = ((ROT || CF) && b == 0) ? ownU : (KNOCK & 2) ? make_double2(1e-3, 1e-4 * node[b]) : __ldg(A.states + node[b]);
                    ra0 = fma(Kr[b][0], q.x, fma(Kr[b][1], q.y, ra0));
                    ra1 = fma(Kr[b][2], q.x, fma(Kr[b][3], q.y, ra1));
                }
            }
        }
        if (WR) {
            const double2 o = racc;
            racc = make_double2(o.x + ra0, o.y + ra1);
        }
        if (KNOCK & 32) {  // keep the strip alive without touching shared memory
#pragma unroll
            for (int b = 0; b < NN; ++b) racc.x += Kr[b][0] + Kr[b][1] + Kr[b][2] + Kr[b][3];
        } else if (WK) {
#pragma unroll
            for (int b = 0; b < NN; ++b) {
                const int kk = (NN <= 4) ? (kw[0] >> (4 + 7 * b)) & 0x7f : (kw[b >> 2] >> (8 * (b & 3))) & 0xff;  // position of node b in my sorted row (plan time)
                double2& s0 = slot(kk);
                double2& s1 = slot(deg + kk);
                const double2 o0 = s0, o1 = s1;
                s0 = make_double2(o0.x + Kr[b][0], o0.y + Kr[b][1]);
                s1 = make_double2(o1.x + Kr[b][2], o1.y + Kr[b][3]);
            }
        }
    }

    if (fillPending) fill();  // nodes without a corner in this element block

    // ---- BC rows, loads, residual ---------------------------------------------------------------------------
    // A warp waits for its slowest lane, so the (rare) constrained nodes take everything from per-constrained-node
    // arrays built by femb_plan_set_bcs: one level of loads behind bcInfo.
    const unsigned m = info & 3u;
    if (m && WK) {  // constrained rows: zeros, occurrence count on the diagonal (AssemblyUtils.py:55-60)
        const double2 cnt = __ldg(A.bcNodeCnt + (info >> 9) - 1);
        const int diagK = (info >> 2) & 0x7f;  // position of the node in its own sorted row
        for (int r = 0; r < 2; ++r) {
            if (m & (1u << r)) {
                for (int k = 0; k < deg; ++k) slot(r * deg + k) = make_double2(0.0, 0.0);
                slot(r * deg + diagK) = r ? make_double2(0.0, cnt.y) : make_double2(cnt.x, 0.0);
            }
        }
    }
    // residual epilogue split around the write-out: its loads are issued here and consumed after the K stores
    double2 fLoad = make_double2(0.0, 0.0), uBC = make_double2(0.0, 0.0), valBC = make_double2(0.0, 0.0);
    if (WR && valid) {
        if (A.loads && finalise) fLoad = __ldg(A.loads + i);
        if (m) {
            uBC = __ldg(A.states + i);
            valBC = __ldg(A.bcNodeVal + (info >> 9) - 1);
        }
    }

    // ---- write-out: stream the warp's run of CSR values (owner table built at the top) ------------------------------
    if (WK && !(KNOCK & 8)) {
        const int base = __shfl_sync(0xffffffffu, off, 0);
        const int start = PERM ? ownStart : 2 * (off - base);  // (node order: recomputed rather than carried across the corner loop)
        const int total = __shfl_sync(0xffffffffu, start + 2 * deg, 31);  // invalid tail lanes hold off = end, deg = 0
        __syncwarp();
        for (int f0 = 0; f0 < total; f0 += 32) {
            const int f = f0 + lane;
            const bool ok = f < total;
            const int r = ok ? owner[f] : 0;
            const int mm = f - __shfl_sync(0xffffffffu, start, r);
            const long long g = PERM ? (long long)__shfl_sync(0xffffffffu, ownGlobal, r) + f : 2ll * base + f;
            if (ok) st_stream(K2 + g, acc[mm * BLOCK + ((wbase + r + mm) & (BLOCK - 1))]);
        }
    }
    if (WR && valid) {
        double rx = racc.x - fLoad.x, ry = racc.y - fLoad.y;  // R = scatter(R_e) - loads (Problem.py:653)
        if (m & 1u) rx = uBC.x - valBC.x;                      // R[bc] = u - c (AssemblyUtils.py:93)
        if (m & 2u) ry = uBC.y - valBC.y;
        reinterpret_cast<double2*>(A.R)[i] = make_double2(rx, ry);
    }
#if ROWS_TRACE
    if (A.stats && lane == 0) {  // [0] sum of warp lifetimes (cycles) [1] warps [2] longest [3] first start (ns) [4] last end (ns) [5] shortest
        long long gEnd;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gEnd));
        const long long life = clock64() - tStart;
        atomicAdd((unsigned long long*)A.stats + 0, (unsigned long long)life);
        atomicAdd((unsigned long long*)A.stats + 1, 1ull);
        atomicMax(A.stats + 2, life);
        atomicMin(A.stats + 3, gStart);
        atomicMax(A.stats + 4, gEnd);
        atomicMin(A.stats + 5, life);
        atomicAdd((unsigned long long*)A.stats + 16 + min(63, (int)((gStart - A.stats[3]) / 1000)), 1ull);  // starts per us
        atomicAdd((unsigned long long*)A.stats + 96 + min(63, (int)(life / 2000)), 1ull);                  // lifetimes per 2000 cycles
    }
#endif
}

}  // namespace rows
