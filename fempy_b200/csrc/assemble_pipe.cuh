// Write-once assembly over node patches: persistent kernel of independent, self-prefetching warp teams (sm_100a).
//
// One CTA per SM; its warps form NUM_TEAMS teams of TEAM warps.  Team g of CTA c walks the patches
// p = c*NUM_TEAMS + g, + gridDim.x*NUM_TEAMS, ...  A team is a complete software pipeline of its own -- no
// producer/consumer warps, no block-wide barriers (for TEAM == 1 the only synchronisation is __syncwarp):
//
//   prefetch (2 ahead)   lane 0 issues a cp.async.bulk of the patch's program blob (patch_blob.h) into one of
//                        three private input stages (mbarrier complete_tx);
//   prefetch (1 ahead)   once that blob has landed the team gathers the coordinates / states / loads /
//                        thickness / BC mask the patch touches with 16-byte cp.async -- the HBM latency of both
//                        steps hides behind the arithmetic of the current patch;
//   compute              one thread per patch element: K_e (10 symmetric 2x2 blocks for a quad) and R_e in FP64
//                        registers, inputs from shared memory only;
//   scatter              conflict-free adds into the team's shared-memory accumulators, which mirror the global
//                        CSR layout of the owned nodes' rows: one phase per local node a, split into plan-time
//                        colours so that no two threads touch the same row at once (fixed order => bit-
//                        reproducible sums, no atomics); the 8 read-modify-writes of a phase are batched;
//   fix-up + write-out   BC rows / loads / residual in shared memory, then the accumulators stream to K with
//                        16-byte st.global.cs (they ARE the CSR rows of the owned nodes), residual pairs directly.
//
// No global atomics, no zero-fill of K, no separate BC or load pass: every CSR value and residual entry is
// written exactly once.
#pragma once
#include <type_traits>

#include "femb_internal.h"
#include "patch_blob.h"

namespace pipe {

template <int TEAM>
struct Cfg {  // 10 warps (<= 204 registers) for teams of 1 or 2 warps; 12 warps (<= 170 registers) = 3 teams of 4
    static constexpr int WARPS = (TEAM == 4) ? 12 : 10;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int NUM_TEAMS = WARPS / TEAM;
};
static constexpr int NI = 3;              // input stages per team

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void st_stream(double2* p, double a, double b) {  // K is never re-read by this kernel
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b));
}
template <int TEAM>
__device__ __forceinline__ void team_sync(int team) {
    if (TEAM == 1) {
        __syncwarp();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(TEAM * 32) : "memory");
    }
}

struct Args {
    long long numPatches;
    int PN;
    const int* blobPtr;             // 16-byte units
    const unsigned char* blob;
    const unsigned char* maskPerm;  // NULL: no BCs in this pass
    const int* bcDof;
    const double* bcVal;
    const double* bcCount;
    int nBC;
    const double2* coords;
    const double2* states;
    const double2* loads;  // may be NULL
    const double* thick;
    double* K;
    double* R;
    int accumulate;    // != 0: add to K / R (second and later element blocks of a mixed mesh)
    int debug;         // measurement only
    long long* stats;  // measurement only: per-CTA cycle counters (16 per CTA) or NULL
    // shared-memory carve-up (bytes), computed on the host from the plan's maxima
    int inStride, offX, offU, offL, offTH, offMask, accStride, offRacc, teamStride, offTeams;
};

struct SmemPlan {
    int inStride, offX, offU, offL, offTH, offMask, accStride, offRacc, teamStride, offTeams, total;
};

static inline SmemPlan smem_plan(int numTeams, int maxBlobBytes, int maxLocal, int maxElems, int maxNodeBlocks, int PN, bool wk) {
    SmemPlan S;
    S.offX = pad16(maxBlobBytes);
    S.offU = S.offX + 16 * maxLocal;
    S.offL = S.offU + 16 * maxLocal;
    S.offTH = S.offL + 16 * PN;
    S.offMask = S.offTH + pad16(8 * maxElems);
    S.inStride = S.offMask + pad16(PN);
    S.offRacc = wk ? 16 * (2 * maxNodeBlocks + PN) : 0;  // one pad double2 per owned node
    S.accStride = S.offRacc + 32 * PN;  // two residual partial-sum arrays (half-element kernel)
    S.teamStride = NI * S.inStride + S.accStride;
    S.offTeams = 256;  // mbarriers (NI per team) live in the first 256 bytes
    S.total = S.offTeams + numTeams * S.teamStride;
    return S;
}

__device__ __forceinline__ int bc_find(const int* __restrict__ bcDof, int n, int dof) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (bcDof[mid] < dof) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int NN, int NQ, bool NL, bool WK, bool WR, int TEAM, bool TIMED>
__global__ void __launch_bounds__(Cfg<TEAM>::THREADS, 1) k_assemble_pipe(const __grid_constant__ femb::Tables<NN, NQ> tb,
                                                              const femb::DMat D, const Args A) {
    using namespace femb;
    static_assert(NN <= 4, "scatter records hold at most 4 nodes per element");
    constexpr int NUM_TEAMS = Cfg<TEAM>::NUM_TEAMS;
    constexpr int TT = TEAM * 32;  // threads of a team = max elements of a patch
    constexpr bool NEED_U = NL || WR;
    constexpr bool R_FROM_K = WK && WR && !NL;  // linear model: R_e = K_e q_e after the Gauss loop

    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);

    const int warp = threadIdx.x >> 5;
    const int team = warp / TEAM;
    if (team >= NUM_TEAMS) return;
    const int tid = threadIdx.x - team * TT;
    uint64_t* blobBar = bars + team * NI;
    if (tid == 0) {
        for (int i = 0; i < NI; ++i) mbar_init(&blobBar[i], 1);
        fence_barrier_init();
    }
    team_sync<TEAM>(team);

    unsigned char* teamSmem = smem + A.offTeams + team * A.teamStride;
    unsigned char* accBytes = teamSmem + NI * A.inStride;
    double2* acc2 = reinterpret_cast<double2*>(accBytes);
    double2* racc = reinterpret_cast<double2*>(accBytes + A.offRacc);

    const long long slot = (long long)blockIdx.x * NUM_TEAMS + team;
    const long long step = (long long)gridDim.x * NUM_TEAMS;
    const int n = (slot < A.numPatches) ? (int)((A.numPatches - slot + step - 1) / step) : 0;
    const bool haveMask = A.maskPerm != nullptr;

    auto issue_blob = [&](int k, int begin, int end) {  // one thread: bulk copy of patch k's blob into stage k % NI
        const uint32_t bytes = 16u * (uint32_t)(end - begin);
        mbar_arrive_expect_tx(&blobBar[k % NI], bytes);
        bulk_g2s(teamSmem + (k % NI) * A.inStride, A.blob + 16ll * begin, bytes, &blobBar[k % NI]);
    };
    auto blob_range = [&](int k, int& begin, int& end) {  // plain loads, consumed an iteration later
        const long long p = slot + (long long)k * step;
        begin = (k < n) ? __ldg(A.blobPtr + p) : 0;
        end = (k < n) ? __ldg(A.blobPtr + p + 1) : 0;
    };
    auto issue_gathers = [&](int k) {  // whole team: per-node / per-element inputs of patch k (blob must have landed)
        unsigned char* in = teamSmem + (k % NI) * A.inStride;
        mbar_wait(&blobBar[k % NI], (k / NI) & 1);
        const int* hdr = reinterpret_cast<const int*>(in);
        const int ne = hdr[0], nOwned = hdr[1], nLocal = hdr[2];
        const BlobLayout L = blob_layout(NN, ne, nOwned, nLocal, hdr[3]);
        const int* lnode = reinterpret_cast<const int*>(in + L.offLNODE);
        const int* elem = reinterpret_cast<const int*>(in + L.offELEM);
        double2* Xs = reinterpret_cast<double2*>(in + A.offX);
        double2* Us = reinterpret_cast<double2*>(in + A.offU);
        double2* Ls = reinterpret_cast<double2*>(in + A.offL);
        double* THs = reinterpret_cast<double*>(in + A.offTH);
        for (int l = tid; l < nLocal; l += TT) {
            const int g = lnode[l];
            cp_async16(&Xs[l], A.coords + g);
            if (NEED_U) cp_async16(&Us[l], A.states + g);
            if (WR && A.loads && l < nOwned) cp_async16(&Ls[l], A.loads + g);
        }
        for (int t = tid; t < ne; t += TT) cp_async8(&THs[t], A.thick + elem[t]);
        if (haveMask) {
            const long long p = slot + (long long)k * step;
            for (int m = tid; m < A.PN / 4; m += TT) cp_async4(in + A.offMask + 4 * m, A.maskPerm + p * A.PN + 4 * m);
        }
    };

    // ---- prologue: blobs of patches 0 and 1 in flight, gathers of patch 0 issued -------------------------------
    int nextBegin = 0, nextEnd = 0;  // blob range of patch k + 2, loaded while patch k - 1 is processed
    if (tid == 0) {
        int b0, e0, b1, e1;
        blob_range(0, b0, e0);
        blob_range(1, b1, e1);
        if (n > 0) issue_blob(0, b0, e0);
        if (n > 1) issue_blob(1, b1, e1);
        blob_range(2, nextBegin, nextEnd);
    }
    if (n > 0) issue_gathers(0);

    long long cWait = 0, cCompute = 0, cScatter = 0, cOut = 0;
    const bool timed = TIMED && A.stats && tid == 0 && team < 2;
    const long long tStart = clock64();

    for (int k = 0; k < n; ++k) {
        long long c0 = timed ? clock64() : 0;
        unsigned char* in = teamSmem + (k % NI) * A.inStride;
        cp_async_wait_all();     // this thread's share of the gathers of patch k
        team_sync<TEAM>(team);   // ... everyone's; also: the previous write-out no longer reads acc / stage (k-1) % NI
        if (tid == 0) {
            if (k + 2 < n) {
                fence_proxy_async();  // generic-proxy reads of stage (k-1) % NI are done before the bulk engine overwrites it
                issue_blob(k + 2, nextBegin, nextEnd);
            }
            blob_range(k + 3, nextBegin, nextEnd);
        }
        if (k + 1 < n) issue_gathers(k + 1);
        if (timed) { const long long c1 = clock64(); cWait += c1 - c0; c0 = c1; }

        const int* hdr = reinterpret_cast<const int*>(in);
        const int ne = hdr[0], nOwned = hdr[1], nnb = hdr[3];
        const BlobLayout L = blob_layout(NN, ne, nOwned, hdr[2], nnb);
        const int* lnode = reinterpret_cast<const int*>(in + L.offLNODE);
        const int2* node = reinterpret_cast<const int2*>(in + L.offNODE);
        const unsigned short* nbRow = reinterpret_cast<const unsigned short*>(in + L.offNBROW);
        const ElemInfo* einfo = reinterpret_cast<const ElemInfo*>(in + L.offEINFO);
        const unsigned short* conn = reinterpret_cast<const unsigned short*>(in + L.offCONN);
        const double2* Xs = reinterpret_cast<const double2*>(in + A.offX);
        const double2* Us = reinterpret_cast<const double2*>(in + A.offU);
        const double2* Ls = reinterpret_cast<const double2*>(in + A.offL);
        const double* THs = reinterpret_cast<const double*>(in + A.offTH);
        const unsigned char* mask = in + A.offMask;

        // ---- zero the accumulators ----------------------------------------------------------------------------
        if (WK)
            for (int idx = tid; idx < 2 * nnb + nOwned; idx += TT) acc2[idx] = make_double2(0.0, 0.0);
        if (WR)
            for (int r = tid; r < nOwned; r += TT) racc[r] = make_double2(0.0, 0.0);

        // ---- element matrix / residual in registers -------------------------------------------------------------
        const int t = tid;
        const bool active = t < ne;
        double Kd[NN][3], Ko[NN * (NN - 1) / 2 + 1][4], Re[NN][2];
        int ln[NN];
        if (active) {
            double x[NN], y[NN], ux[NN], uy[NN];
#pragma unroll
            for (int a = 0; a < NN; ++a) {
                ln[a] = conn[a * ne + t];
                const double2 c = Xs[ln[a]];
                x[a] = c.x;
                y[a] = c.y;
                if (NEED_U && !R_FROM_K) {
                    const double2 q = Us[ln[a]];
                    ux[a] = q.x;
                    uy[a] = q.y;
                } else {
                    ux[a] = uy[a] = 0.0;
                }
            }
            const double theta = THs[t];
            if (A.debug & 8) {  // measurement only: skip the Gauss loop
#pragma unroll
                for (int a = 0; a < NN; ++a) {
                    Kd[a][0] = Kd[a][1] = Kd[a][2] = x[a];
                    Re[a][0] = Re[a][1] = y[a];
                }
#pragma unroll
                for (int q = 0; q < NN * (NN - 1) / 2 + 1; ++q) Ko[q][0] = Ko[q][1] = Ko[q][2] = Ko[q][3] = theta;
            } else {
                element_sym<NN, NQ, NL, WK, (WR && !R_FROM_K)>(tb, D, x, y, ux, uy, theta, Kd, Ko, Re);
            }
            if (R_FROM_K) {
#pragma unroll
                for (int a = 0; a < NN; ++a) {
                    const double2 q = Us[ln[a]];
                    ux[a] = q.x;
                    uy[a] = q.y;
                }
#pragma unroll
                for (int a = 0; a < NN; ++a) {
                    double r0 = 0.0, r1 = 0.0;
#pragma unroll
                    for (int b = 0; b < NN; ++b) {
                        double v[4];
                        sym_block<NN>(Kd, Ko, a, b, v);
                        r0 = fma(v[0], ux[b], fma(v[1], uy[b], r0));
                        r1 = fma(v[2], ux[b], fma(v[3], uy[b], r1));
                    }
                    Re[a][0] = r0;
                    Re[a][1] = r1;
                }
            }
        } else {
#pragma unroll
            for (int a = 0; a < NN; ++a) ln[a] = 0;
        }
        team_sync<TEAM>(team);  // accumulators are zero
        if (timed) { const long long c1 = clock64(); cCompute += c1 - c0; c0 = c1; }

        // ---- scatter-add: phase a = local node a, colours make same-phase updates disjoint --------------------------
        // the element's 32-byte scatter record, as two 128-bit loads kept in registers (fields extracted by shifts)
        const uint4 i0 = reinterpret_cast<const uint4*>(einfo + (active ? t : 0))[0];
        const uint4 i1 = reinterpret_cast<const uint4*>(einfo + (active ? t : 0))[1];
        const unsigned rowBasePk[2] = {i0.x, i0.y};
        const unsigned degPk = i0.z;
        const unsigned colourPk = active ? i0.w : 0xffffffffu;
        const unsigned kPk[4] = {i1.x, i1.y, i1.z, i1.w};
#pragma unroll
        for (int a = 0; a < NN; ++a) {
            const int numColours = hdr[4 + a];
            const int colour = (colourPk >> (8 * a)) & 0xff;
            for (int c = 0; c < numColours; ++c) {
                if (colour == c && !(A.debug & 4)) {
                    if (WK) {
                        double2* row0 = acc2 + ((rowBasePk[a >> 1] >> (16 * (a & 1))) & 0xffff);
                        const int deg = (degPk >> (8 * a)) & 0xff;
                        const unsigned kpack = kPk[a];
                        double2 s0[NN], s1[NN];
#pragma unroll
                        for (int b = 0; b < NN; ++b) {  // the NN blocks of a row are distinct: batch the loads
                            const int kk = (kpack >> (8 * b)) & 0xff;
                            s0[b] = row0[kk];
                            s1[b] = row0[kk + deg];
                        }
#pragma unroll
                        for (int b = 0; b < NN; ++b) {
                            double v[4];
                            sym_block<NN>(Kd, Ko, a, b, v);
                            s0[b].x += v[0];
                            s0[b].y += v[1];
                            s1[b].x += v[2];
                            s1[b].y += v[3];
                        }
#pragma unroll
                        for (int b = 0; b < NN; ++b) {
                            const int kk = (kpack >> (8 * b)) & 0xff;
                            row0[kk] = s0[b];
                            row0[kk + deg] = s1[b];
                        }
                    }
                    if (WR) {
                        double2 s = racc[ln[a]];
                        s.x += Re[a][0];
                        s.y += Re[a][1];
                        racc[ln[a]] = s;
                    }
                }
                team_sync<TEAM>(team);
            }
        }

        // ---- constrained rows: zeros, occurrence count on the diagonal (AssemblyUtils.py:55-60) --------------------
        if (haveMask) {
            if (WK) {
                for (int r = tid; r < nOwned; r += TT) {
                    const unsigned m = mask[r];
                    if (m) {
                        const int2 nd = node[r];
                        const int start = nd.y & 0xffff, deg = (nd.y >> 16) & 0xff, diagK = (nd.y >> 24) & 0xff;
                        double* rows = reinterpret_cast<double*>(acc2 + 2 * start + r);
                        const int gi = lnode[r];
                        for (int c = 0; c < 2; ++c) {
                            if (m & (1u << c)) {
                                for (int q = 0; q < 2 * deg; ++q) rows[c * 2 * deg + q] = 0.0;
                                rows[c * 2 * deg + 2 * diagK + c] = A.bcCount[bc_find(A.bcDof, A.nBC, 2 * gi + c)];
                            }
                        }
                    }
                }
            }
            team_sync<TEAM>(team);
        }
        if (timed) { const long long c1 = clock64(); cScatter += c1 - c0; c0 = c1; }

        // ---- write-out: the accumulators mirror the CSR rows, so this is a 16-byte-per-thread streaming copy --------
        if (WK && !(A.debug & 16)) {
            double2* K2 = reinterpret_cast<double2*>(A.K);
            const int total = 2 * nnb;
            for (int base = tid; base < total; base += 4 * TT) {  // four 16-byte pieces in flight per thread
                int rr[4];
                int2 nd[4];
                double2 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) rr[u] = (base + u * TT < total) ? nbRow[(base + u * TT) >> 1] : 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    nd[u] = node[rr[u]];
                    v[u] = acc2[min(base + u * TT, total - 1) + rr[u]];  // one pad double2 per preceding owned node
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int sidx = base + u * TT;
                    if (sidx < total) {
                        const long long g = 2ll * nd[u].x + (sidx - 2 * (nd[u].y & 0xffff));
                        if (!A.accumulate) {
                            st_stream(K2 + g, v[u].x, v[u].y);
                        } else {
                            const double2 o = K2[g];
                            K2[g] = make_double2(o.x + v[u].x, o.y + v[u].y);
                        }
                    }
                }
            }
        }
        if (WR && !(A.debug & 32)) {
            for (int r = tid; r < nOwned; r += TT) {
                const int gi = lnode[r];
                double2 s = racc[r];
                double2* rr = reinterpret_cast<double2*>(A.R) + gi;
                if (A.accumulate) {
                    const double2 o = *rr;
                    *rr = make_double2(o.x + s.x, o.y + s.y);
                } else {
                    if (A.loads) {  // R = scatter(R_e) - loads (Problem.py:653)
                        const double2 f = Ls[r];
                        s.x -= f.x;
                        s.y -= f.y;
                    }
                    const unsigned m = haveMask ? mask[r] : 0u;
                    if (m) {  // R[bc] = u - c (AssemblyUtils.py:93)
                        const double2 u = Us[r];
                        if (m & 1u) s.x = u.x - A.bcVal[bc_find(A.bcDof, A.nBC, 2 * gi)];
                        if (m & 2u) s.y = u.y - A.bcVal[bc_find(A.bcDof, A.nBC, 2 * gi + 1)];
                    }
                    *rr = s;
                }
            }
        }
        if (timed) { const long long c1 = clock64(); cOut += c1 - c0; c0 = c1; }
    }
    if (timed) {
        long long* st = A.stats + 16ll * blockIdx.x + 6 * team;
        st[0] = clock64() - tStart;
        st[1] = n;
        st[2] = cWait;
        st[3] = cCompute;
        st[4] = cScatter;
        st[5] = cOut;
    }
}

}  // namespace pipe
