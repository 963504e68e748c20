// Write-once row assembly over fan-ordered corners (sm_100a): 3- and 4-node elements, small strain, single element block.
//
// Like the row-owner kernel (assemble_rows.cuh) one thread owns the two DOF rows of one node, evaluates for every adjacent
// element only the node-row strip of K_e, and the warp streams its 32 finished rows out as one contiguous run of the CSR
// values.  What differs is how a row comes together.  The plan orders the corners of a node as a fan (plan.cu:
// k_fan_corners): the last node of one corner is the first node of the next.  Walking the fan the thread finishes every
// 2x2 block of its row in registers --
//     own node        sum over all corners                  (kept in registers, stored after the walk)
//     edge neighbour  last block of corner k + first block of corner k+1   (carried in registers from k to k+1)
//     diagonal node   one corner                            (4-node elements)
// -- and stores it ONCE into the row staging area in shared memory: no zero-fill, no read-modify-write.  The accumulating
// kernel spends 44 % of the L1TEX data pipe on shared-memory wavefronts (ncu, round 2), two thirds of them the per-corner
// read-modify-writes this kernel does not have.
//
// Element math (closed forms of the reference's point-by-point evaluation, FEMpy/Elements/Element.py:942-1168 with the
// tables of QuadElement2D / TriElement2D order 1; the plan checks the block's tables, plan.cu: standard_q4 / standard_t3):
//   4 nodes  with 8 det J dN_b/dx = C0_b + xi X0_b + eta E0_b, 8 det J dN_b/dy = C1_b + xi X1_b + eta E1_b (every
//            coefficient one difference of two node coordinates), 8 det J = d0 + xi dx + eta de and the 2x2 Gauss rule,
//            the geometric sums are S[d][D]_0b = P_d C_Db + Q_d X_Db + R_d E_Db with P, Q, R built from four moments of
//            theta / (8 det J) over the Gauss points;
//   3 nodes  constant gradients: S[d][D]_0b = theta / (2 det J) * Gh_d0 Gh_Db, Gh = adj(J) dN.
// D is applied per block: K = [[D00 S00 + D22 S11, D01 S01 + D22 S10], [D01 S10 + D22 S01, D11 S11 + D22 S00]].
//
// Memory side: per-node headers with the first corner record inline (one load level before the first
// element), records two corners ahead / node data one corner ahead in registers, an L2 look-ahead for the warps a
// sixth of a resident wave ahead, and -- 4-node elements -- the finished run written out by ONE cp.async.bulk per warp.
#pragma once
#include "assemble_rows.cuh"

namespace fan {

static constexpr int BLOCK = 32;

// Per-node input of the kernel (plan.cu: k_fan_headers).  Everything a thread needs before its first element arrives
// with ONE level of loads: header {first row position, index of the node's further records, corners | degree << 8,
// BC word (femb_plan_set_bcs)} and, behind it, the record of the node's first corner; the records of the further corners
// sit in a second array.  (The first version read row pointer, corner pointer and BC word from three arrays and then
// the first record: two dependent levels, 13 % of the stall samples at config 2, 15 % at config 4.)
//   record, 3 nodes           {slots | flags, node 1, node 2, element}
//   record, 3 nodes compact   further corners only, 8 bytes: the first node is the previous corner's last (plan.cu: k_fan_rest3)
//   record, 4 nodes compact   {slots | flags, three node ids and the element id as signed 24-bit differences to the owner}
//   record, 4 nodes wide      {slots | flags, node 1, node 2, node 3} {element, -, -, -}   (numberings without locality)
template <int NN, bool CMP>
__host__ __device__ constexpr int rec_words() {
    return (NN == 4 && !CMP) ? 2 : 1;
}

struct Args {
    long long numNodes;
    const int4* hdr;          // (1 + rec_words) int4 per node, node count padded to a multiple of BLOCK
    const int4* rest;         // records of the corners after the first, rec_words int4 each
    const double2* coords;
    const double2* states;
    const double* thick;
    const double2* loads;     // may be NULL
    int useBC;                // 0: ignore the headers' BC words
    const double2* bcNodeCnt;
    const double2* bcNodeVal;
    int maxDeg;
    int zero;                 // always 0 (runtime-uniform index: keeps D out of the constant-bank operand slot)
    // L2 look-ahead (0 = off): every thread asks L2 for the inputs of the node pfNodes ahead of its own -- a fraction of
    // a resident wave of warps -- so that later warps find headers, records, node data and thickness in L2 instead of DRAM
    // (limits = array length - distance, 0 when off: an index below its limit has a valid look-ahead target)
    int pfNodes, pfNodeLimit, pfRest, pfRestLimit, pfElems, pfElemLimit;
    double* K;
    double* R;
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int NN>
__host__ __device__ constexpr int resident_warps() {
#ifndef FAN_WARPS4
#define FAN_WARPS4 16
#endif
#ifndef FAN_WARPS3
#define FAN_WARPS3 24
#endif
    return NN == 4 ? FAN_WARPS4 : FAN_WARPS3;
}
// write-out by one cp.async.bulk per warp from the unswizzled staging area (bit 0: 3-node elements, bit 1: 4-node
// elements).  Measured: 4 nodes 32.3 -> 30.3 us at config 2, 2.34 -> 2.19 ms at config 5; 3 nodes 256 -> 265 us at config 4
#ifndef FAN_QSTASH
#define FAN_QSTASH 1
#endif
#ifndef FAN_QLAST
#define FAN_QLAST 1
#endif
#ifndef FAN_BULK
#define FAN_BULK 2
#endif

// streaming inputs (headers, corner records): read once, kept out of L1 so that the node data the corners share stays there
__device__ __forceinline__ int4 ld_stream(const int4* p) {
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int2 ld_stream(const int2* p) {
    int2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
// volatile loads: the compiler keeps the request where it is written -- as early as the address is known -- instead of
// sinking it next to its use (lower register pressure, but the full memory latency on the critical path)
__device__ __forceinline__ double2 ld_early(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_early(const double* p) {
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

template <int NN, bool WK, bool WR, bool CMP>
__global__ void __launch_bounds__(BLOCK, resident_warps<NN>()) k_assemble_fan(const __grid_constant__ femb::DMat D, const Args A) {
    using femb::fast_rcp;
    static_assert(NN == 3 || NN == 4, "fan kernel: 3- and 4-node elements");
    constexpr int RW = rec_words<NN, CMP>();
    // Row staging in the layout of the output: the warp's 32 rows are one contiguous run of the CSR values (both DOF rows
    // of a node are adjacent, rows in node order), position p of the run sits at acc[swz(p)].  The swizzle (bit 3 of p
    // flips bit 0) spreads the lanes' block stores -- 2*deg positions apart -- over all 16-byte bank groups and leaves the
    // consecutive positions the write-out reads conflict-free.
    extern __shared__ double2 acc[];  // [<= 2*maxDeg*BLOCK]
    const int lane = threadIdx.x;
    const long long i = (long long)blockIdx.x * BLOCK + lane;
    const bool valid = i < A.numNodes;
    // ---- level 1: header + first record (the header array is padded: tail lanes read {end of the values, 0, 0, 0}) --
    const int4* hp = A.hdr + (1 + RW) * i;
    const int4 h = ld_stream(hp);
    int4 rec = ld_stream(hp + 1);
    int el = 0;
    if (RW == 2) el = ld_stream(reinterpret_cast<const int*>(hp + 2));
    double2 ownX = make_double2(0.0, 0.0), ownU = make_double2(0.0, 0.0);
    if (valid) {
        ownX = __ldg(A.coords + i);
        ownU = __ldg(A.states + i);
    }
    const int off = h.x, nc = h.z & 0xff, deg = (h.z >> 8) & 0xff;
    const unsigned info = A.useBC ? (unsigned)h.w : 0u;
    constexpr bool T3C = NN == 3 && CMP;              // 8-byte records for the further corners of 3-node fans
    const int4* rest = A.rest + (long long)RW * h.y;  // record of corner c >= 1 at rest[RW * (c - 1)]
    const int2* rest2 = reinterpret_cast<const int2*>(A.rest) + h.y;  // T3C
    // record (and element id) of corner c >= 1
    auto load_rec = [&](int c, int4& r, int& relem) {
        if constexpr (T3C) {
            const int2 v = ld_stream(rest2 + (c - 1));
            r = make_int4(v.x, v.y, 0, 0);
        } else {
            r = ld_stream(rest + RW * (c - 1));
            if (RW == 2) relem = ld_stream(reinterpret_cast<const int*>(rest + RW * (c - 1) + 1));
        }
    };
    if (i < A.pfNodeLimit) {  // L2 look-ahead for the warp that will own node i + pfNodes
        const long long j = i + A.pfNodes;
        prefetch_l2(A.hdr + (1 + RW) * j);
        if (RW == 2) prefetch_l2(A.hdr + (1 + RW) * j + 2);
        prefetch_l2(A.coords + j);
        prefetch_l2(A.states + j);
        if (WR && A.loads) prefetch_l2(A.loads + j);
    }
    if (h.y < A.pfRestLimit) {
        if (T3C) prefetch_l2(rest2 + A.pfRest);
        else prefetch_l2(rest + RW * A.pfRest);
    }
    // node ids / element id of a record (CMP: four signed 24-bit differences to my node id)
    auto decode = [&](const int4 r, const int relem, int& n1, int& n2, int& n3, int& e) {
        n1 = r.y, n2 = r.z, n3 = r.w;  // 3 nodes: the fourth word is the element id
        e = NN == 4 ? relem : r.w;
        if constexpr (CMP && NN == 4) {
            const int me = (int)i;
            e = me + (r.w >> 8);
            n3 = me + ((int)(__funnelshift_r((unsigned)r.z, (unsigned)r.w, 16) << 8) >> 8);
            n2 = me + ((int)(__funnelshift_r((unsigned)r.y, (unsigned)r.z, 24) << 8) >> 8);
            n1 = me + ((r.y << 8) >> 8);
        }
    };
    // ---- level 2: the first corner's nodes and thickness, the second corner's record -------------------------------
    // Software pipeline over the corners: records run two corners ahead, coordinates and thickness one corner ahead
    // (requested as soon as the current corner has turned its own coordinates into differences), so that the dependent
    // chain record -> coordinates -> geometry waits on memory for the first corner only.
    int n1 = 0, n2 = 0, n3 = 0, e = 0;
    double2 X1 = make_double2(0.0, 0.0), X2 = X1, X3 = X1;
    double theta = 0.0;
    int4 rec1 = make_int4(0, 0, 0, 0);
    int el1 = 0;
    if (nc > 0) {
        decode(rec, el, n1, n2, n3, e);
        X1 = ld_early(A.coords + n1);
        X2 = ld_early(A.coords + n2);
        if (NN == 4) X3 = ld_early(A.coords + n3);
        theta = ld_early(A.thick + e);
        if (NN == 4 && e < A.pfElemLimit) prefetch_l2(A.thick + e + A.pfElems);
        if (nc > 1) load_rec(1, rec1, el1);
    }
    const int firstSlot = ((unsigned)rec.x >> 7) & 0x7f, firstNode = n1;  // closed fans wrap into the first corner's first neighbour
    unsigned kw = (unsigned)rec.x, lastFlags = 0u;

    const int base = __shfl_sync(0xffffffffu, off, 0);
    const int ownStart = 2 * (off - base);  // my first position in the warp's run of values
    constexpr bool BULK = ((FAN_BULK >> (NN - 3)) & 1) != 0;
    auto swz = [](int p) { return BULK ? p : p ^ ((p >> 3) & 1); };
    auto put = [&](int m, const double2 v) { acc[swz(ownStart + m)] = v; };

    // D through registers (an FP64 instruction with a constant-bank operand issues at half rate on sm_100)
    const double* Dv = &D.d00 + A.zero;
    const double d00 = Dv[0], d01 = Dv[1], d11 = Dv[2], d22 = Dv[3];

    double o0 = 0.0, o1 = 0.0, o2 = 0.0, o3 = 0.0;  // own block, summed over the fan
    double k0 = 0.0, k1 = 0.0, k2 = 0.0, k3 = 0.0;  // block carried from the previous corner into this one's first neighbour
    double ra0 = 0.0, ra1 = 0.0;                     // residual pair R_i = sum_j K_ij q_j, taken from the FINISHED row blocks

    // geometric sums -> 2x2 block: K = [[D00 S00 + D22 S11, D01 S01 + D22 S10], [D01 S10 + D22 S01, D11 S11 + D22 S00]]
    auto apply_d = [&](double s00, double s01, double s10, double s11, double& b0, double& b1, double& b2, double& b3) {
        b0 = fma(d00, s00, d22 * s11);
        b1 = fma(d01, s01, d22 * s10);
        b2 = fma(d01, s10, d22 * s01);
        b3 = fma(d11, s11, d22 * s00);
    };
    auto residual = [&](double b0, double b1, double b2, double b3, const double2 q) {
        if (WR) {
            ra0 = fma(b0, q.x, fma(b1, q.y, ra0));
            ra1 = fma(b2, q.x, fma(b3, q.y, ra1));
        }
    };
    // a finished row block: its share of the residual, then its place in the row
    auto store = [&](int k, double b0, double b1, double b2, double b3, const double2 q) {
        residual(b0, b1, b2, b3, q);
        if (WK) {
            put(k, make_double2(b0, b1));
            put(deg + k, make_double2(b2, b3));
        }
    };

    for (int c = 0; c < nc; ++c) {
        const int nLast = NN == 4 ? n3 : n2;
        const double th = theta;
        double2 q1 = make_double2(0.0, 0.0), q2 = q1;
        if (WR) {
            q1 = ld_early(A.states + n1);
            if (NN == 4) q2 = ld_early(A.states + n2);
        }
        const unsigned kwc = kw;
        lastFlags = kwc;
        double2 qL = make_double2(0.0, 0.0);  // an ending fan's last neighbour state: requested here, not after the corner
        if (FAN_QLAST && WR && !(kwc & (3u << 28))) qL = ld_early(A.states + nLast);
        int4 rec2 = make_int4(0, 0, 0, 0);
        int el2 = 0;
        if (c + 2 < nc) load_rec(c + 2, rec2, el2);
        if (NN == 3 && e < A.pfElemLimit) prefetch_l2(A.thick + e + A.pfElems);
        // the next corner's coordinates and thickness replace this corner's once the geometry has consumed them
        auto next_corner = [&]() {
            if (c + 1 < nc) {
                if constexpr (T3C) {  // one chain: this corner's last node is the next one's first (coordinates are at hand)
                    n1 = n2;
                    X1 = X2;
                    n2 = (int)i + (int)(short)(((unsigned)rec1.y & 0x1ffu) | ((((unsigned)rec1.x >> 21) & 0x7fu) << 9));
                    e = (int)i + (rec1.y >> 9);
                } else {
                    decode(rec1, el1, n1, n2, n3, e);
                    X1 = ld_early(A.coords + n1);
                }
                X2 = ld_early(A.coords + n2);
                if (NN == 4) X3 = ld_early(A.coords + n3);
                theta = ld_early(A.thick + e);
                kw = (unsigned)rec1.x;
            }
        };
        double b0, b1, b2, b3;
        if constexpr (NN == 4) {
            const double ya = X1.y - X3.y, yb = X2.y - ownX.y, yc = X3.y - X2.y, yd = ownX.y - X1.y, ye = X2.y - X1.y, yf = ownX.y - X3.y;
            const double xa = X3.x - X1.x, xb = ownX.x - X2.x, xc = X2.x - X3.x, xd = X1.x - ownX.x, xe = X1.x - X2.x, xf = X3.x - ownX.x;
            constexpr double G = 0.57735026918962576451, THIRD = 1.0 / 3.0;
            const double dd0 = xb * ya - xa * yb;
            const double gx = G * (xc * yd - xd * yc), ge = G * (xe * yf - xf * ye);
            const double tp = dd0 + gx, tm = dd0 - gx, th8 = 0.125 * th;
            const double spp = th8 * fast_rcp(tp + ge), spm = th8 * fast_rcp(tp - ge);
            const double smp = th8 * fast_rcp(tm + ge), smm = th8 * fast_rcp(tm - ge);
            const double up = spp + spm, um = smp + smm, vp = spp - spm, vm = smp - smm;
            const double m0 = up + um, mx = G * (up - um), me = G * (vp + vm), mxe = THIRD * (vp - vm), m3 = THIRD * m0;
            const double P0 = fma(m0, ya, fma(mx, yc, me * ye)), Q0 = fma(mx, ya, fma(m3, yc, mxe * ye)), R0 = fma(me, ya, fma(mxe, yc, m3 * ye));
            const double P1 = fma(m0, xa, fma(mx, xc, me * xe)), Q1 = fma(mx, xa, fma(m3, xc, mxe * xe)), R1 = fma(me, xa, fma(mxe, xc, m3 * xe));
            next_corner();
            // b:        0            1            2            3
            // C_Db:  ya | xa      yb | xb     -ya | -xa    -yb | -xb
            // X_Db:  yc | xc     -yc | -xc     yd | xd     -yd | -xd
            // E_Db:  ye | xe      yf | xf     -yf | -xf    -ye | -xe
            apply_d(fma(P0, ya, fma(Q0, yc, R0 * ye)), fma(P0, xa, fma(Q0, xc, R0 * xe)), fma(P1, ya, fma(Q1, yc, R1 * ye)),
                    fma(P1, xa, fma(Q1, xc, R1 * xe)), b0, b1, b2, b3);
            o0 += b0, o1 += b1, o2 += b2, o3 += b3;
            apply_d(fma(P0, yb, fma(R0, yf, -(Q0 * yc))), fma(P0, xb, fma(R0, xf, -(Q0 * xc))), fma(P1, yb, fma(R1, yf, -(Q1 * yc))),
                    fma(P1, xb, fma(R1, xf, -(Q1 * xc))), b0, b1, b2, b3);
            b0 += k0, b1 += k1, b2 += k2, b3 += k3;
            store((kwc >> 7) & 0x7f, b0, b1, b2, b3, q1);
            if (FAN_QSTASH && WR && c == 0) acc[swz(ownStart + (int)(kwc & 0x7f))] = q1;
            apply_d(fma(Q0, yd, -fma(P0, ya, R0 * yf)), fma(Q0, xd, -fma(P0, xa, R0 * xf)), fma(Q1, yd, -fma(P1, ya, R1 * yf)),
                    fma(Q1, xd, -fma(P1, xa, R1 * xf)), b0, b1, b2, b3);
            store((kwc >> 14) & 0x7f, b0, b1, b2, b3, q2);
            apply_d(-fma(P0, yb, fma(Q0, yd, R0 * ye)), -fma(P0, xb, fma(Q0, xd, R0 * xe)), -fma(P1, yb, fma(Q1, yd, R1 * ye)),
                    -fma(P1, xb, fma(Q1, xd, R1 * xe)), k0, k1, k2, k3);

        } else {
            // Gh_0 = (y1 - y2, y2 - y0, y0 - y1), Gh_1 = (x2 - x1, x0 - x2, x1 - x0), 2 * area = det
            const double g00 = X1.y - X2.y, g01 = X2.y - ownX.y, g02 = ownX.y - X1.y;
            const double g10 = X2.x - X1.x, g11 = ownX.x - X2.x, g12 = X1.x - ownX.x;
            next_corner();
            const double det = g12 * g01 - g11 * g02;  // (x1 - x0)(y2 - y0) - (x2 - x0)(y1 - y0)
            const double sc = 0.5 * th * fast_rcp(det);
            const double A0 = sc * g00, A1 = sc * g10;
            apply_d(A0 * g00, A0 * g10, A1 * g00, A1 * g10, b0, b1, b2, b3);
            o0 += b0, o1 += b1, o2 += b2, o3 += b3;
            apply_d(A0 * g01, A0 * g11, A1 * g01, A1 * g11, b0, b1, b2, b3);
            b0 += k0, b1 += k1, b2 += k2, b3 += k3;
            store((kwc >> 7) & 0x7f, b0, b1, b2, b3, q1);
            if (FAN_QSTASH && WR && c == 0) acc[swz(ownStart + (int)(kwc & 0x7f))] = q1;
            apply_d(A0 * g02, A0 * g12, A1 * g02, A1 * g12, k0, k1, k2, k3);
        }
        if (!(kwc & (3u << 28))) {  // the fan ends here (open end or chain break): the carried block is final
            store((kwc >> (7 * (NN - 1))) & 0x7f, k0, k1, k2, k3, !WR ? make_double2(0.0, 0.0) : FAN_QLAST ? qL : __ldg(A.states + nLast));
            k0 = k1 = k2 = k3 = 0.0;
        }
        rec1 = rec2;
        el1 = el2;
    }
    if (nc > 0) {
        if (lastFlags & (1u << 29)) {  // closed fan: the last corner continues into the first one's first neighbour
            // FAN_QSTASH: the state of the first corner's first neighbour waited in my own block's staging slot (free until
            // the own block is stored below): no dependent global load in the epilogue
            residual(k0, k1, k2, k3, !WR ? make_double2(0.0, 0.0) : FAN_QSTASH ? acc[swz(ownStart + (int)(lastFlags & 0x7f))] : __ldg(A.states + firstNode));
            if (WK) {
                double2& s0 = acc[swz(ownStart + firstSlot)];
                double2& s1 = acc[swz(ownStart + deg + firstSlot)];
                const double2 p0 = s0, p1 = s1;
                s0 = make_double2(p0.x + k0, p0.y + k1);
                s1 = make_double2(p1.x + k2, p1.y + k3);
            }
        }
        store((int)(lastFlags & 0x7f), o0, o1, o2, o3, ownU);  // own block
    }

    // ---- BC rows, loads, residual: as in the row-owner kernel ------------------------------------------------------
    const unsigned m = info & 3u;
    if (m && WK) {  // constrained rows: zeros, occurrence count on the diagonal (AssemblyUtils.py:55-60)
        const double2 cnt = __ldg(A.bcNodeCnt + (info >> 9) - 1);
        const int diagK = (info >> 2) & 0x7f;
        for (int r = 0; r < 2; ++r) {
            if (m & (1u << r)) {
                for (int k = 0; k < deg; ++k) put(r * deg + k, make_double2(0.0, 0.0));
                put(r * deg + diagK, r ? make_double2(0.0, cnt.y) : make_double2(cnt.x, 0.0));
            }
        }
    }
    double2 fLoad = make_double2(0.0, 0.0), valBC = make_double2(0.0, 0.0);
    if (WR && valid) {
        if (A.loads) fLoad = __ldg(A.loads + i);
        if (m) valBC = __ldg(A.bcNodeVal + (info >> 9) - 1);
    }
    // ---- write-out: stream the warp's run of CSR values (both DOF rows of a node are adjacent, rows in node order) ---
    if (WK) {
        double2* K2 = reinterpret_cast<double2*>(A.K) + 2ll * base;
        const int total = __shfl_sync(0xffffffffu, ownStart + 2 * deg, 31);  // padded tail lanes hold off = end, deg = 0
        if constexpr (BULK) {
            // the staged run IS the output (no swizzle): one bulk copy shared -> global by the async proxy replaces the
            // warp's LDS / STG loop; the warp only waits until the copy engine has read the staging area
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && total > 0) {
                const unsigned src = (unsigned)__cvta_generic_to_shared(acc);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(K2), "r"(src), "r"(total * 16) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        } else {
            __syncwarp();
#pragma unroll 4
            for (int f = lane; f < total; f += 32) rows::st_stream(K2 + f, acc[swz(f)]);
        }
    }
    if (WR && valid) {
        double rx = ra0 - fLoad.x, ry = ra1 - fLoad.y;  // R = scatter(R_e) - loads (Problem.py:653)
        if (m & 1u) rx = ownU.x - valBC.x;             // R[bc] = u - c (AssemblyUtils.py:93)
        if (m & 2u) ry = ownU.y - valBC.y;
        reinterpret_cast<double2*>(A.R)[i] = make_double2(rx, ry);
    }
}

}  // namespace fan
