// Plan: mesh topology + element tables -> node adjacency, sorted CSR pattern, scatter map, BC staging.
#include <cmath>

#include "femb_internal.h"

using namespace femb;

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

// the atomic fill leaves each node's element list in arbitrary order: sort it (short lists, one thread per node)
__global__ void k_sort_adj(long long numNodes, const int* __restrict__ adjPtr, int* __restrict__ adj) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= numNodes) return;
    const int lo = adjPtr[i], hi = adjPtr[i + 1];
    for (int m = lo + 1; m < hi; ++m) {
        const int v = adj[m];
        int q = m - 1;
        while (q >= lo && adj[q] > v) {
            adj[q + 1] = adj[q];
            --q;
        }
        adj[q + 1] = v;
    }
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


// element-major connectivity for the row-owner kernel: ceil(nn/4) int4 words per element, padded with -1
__global__ void k_conn_elem_major(const int* __restrict__ conn, long long E, int nn, int4* __restrict__ connE) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int cw = (nn + 3) / 4;
    for (int w = 0; w < cw; ++w) {
        int v[4] = {-1, -1, -1, -1};
        for (int a = 4 * w; a < nn && a < 4 * w + 4; ++a) v[a - 4 * w] = conn[(long long)a * E + e];
        connE[e * cw + w] = make_int4(v[0], v[1], v[2], v[3]);
    }
}

// Corner records of the row-owner kernel: for every (node, adjacent element), in adjacency order, the element id, the
// node's local index a in the element and the position k_b of each element node in the node's sorted neighbour list.
//   nn <= 4: int2 {e, a | k_0 << 4 | k_1 << 11 | k_2 << 18 | k_3 << 25}            (7-bit positions)
//   nn <= 12: int4 {e | a << 28, k_0..k_3, k_4..k_7, k_8..k_11}                      (one byte per position)
//   nn == 3:  int4 {e, 0 | k_a << 4 | k_a+1 << 11 | k_a+2 << 18, node a+1, node a+2}: the triangle ROTATED so that the
//             owning node is local node 0 (the linear triangle and its centroid rule are invariant under cyclic
//             renumbering).  The kernel then needs neither the connectivity load nor the owner's own coordinates and
//             state per corner: 6 instead of 11 gathers per corner of a kernel that ncu shows L1TEX-bound (81 %).
// (the records of a block are indexed by the global corner number; e is the element id within the block)
__global__ void k_corners(long long N, const int* __restrict__ nrowptr, const int* __restrict__ ncols,
                          const int* __restrict__ cBegin, const int* __restrict__ cEnd, const int* __restrict__ adj,
                          const int4* __restrict__ connE, int nn, int elemOffset, void* __restrict__ corners, int* __restrict__ cornerElem) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int off = nrowptr[i], deg = nrowptr[i + 1] - off;
    const int cw = (nn + 3) / 4;
    for (int c = cBegin[i]; c < cEnd[i]; ++c) {
        const int e = adj[c] - elemOffset;
        unsigned a = 0, packed = 0, kb[3] = {0u, 0u, 0u};
        for (int b = 0; b < nn; ++b) {
            const int4 cn = connE[(long long)e * cw + (b >> 2)];
            const int node = (b & 3) == 0 ? cn.x : (b & 3) == 1 ? cn.y : (b & 3) == 2 ? cn.z : cn.w;
            if (node == (int)i) a = (unsigned)b;
            int lo = 0, hi = deg;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (ncols[off + mid] < node) lo = mid + 1; else hi = mid;
            }
            if (nn <= 4) packed |= (unsigned)lo << (4 + 7 * b);
            else kb[b >> 2] |= (unsigned)lo << (8 * (b & 3));
        }
        if (nn == 3) {
            const int4 cn = connE[e];
            const int nd[3] = {cn.x, cn.y, cn.z};
            unsigned rot = 0;
            for (int b = 0; b < 3; ++b) rot |= ((packed >> (4 + 7 * ((a + b) % 3))) & 0x7fu) << (4 + 7 * b);
            reinterpret_cast<int4*>(corners)[c] = make_int4(e, (int)rot, nd[(a + 1) % 3], nd[(a + 2) % 3]);
        } else if (nn == 4 && FEMB_ROT4) {  // the same rotation for the bilinear quad (its 2x2 rule is invariant under it):
            const int4 cn = connE[e];         // {k_a | k_a+1 << 7 | k_a+2 << 14 | k_a+3 << 21, node a+1, node a+2, node a+3} + element id aside
            const int nd[4] = {cn.x, cn.y, cn.z, cn.w};
            unsigned rot = 0;
            for (int b = 0; b < 4; ++b) rot |= ((packed >> (4 + 7 * ((a + b) & 3))) & 0x7fu) << (7 * b);
            reinterpret_cast<int4*>(corners)[c] = make_int4((int)rot, nd[(a + 1) & 3], nd[(a + 2) & 3], nd[(a + 3) & 3]);
            cornerElem[c] = e;
        } else if (nn <= 4) reinterpret_cast<int2*>(corners)[c] = make_int2(e, (int)(packed | a));
        else reinterpret_cast<int4*>(corners)[c] = make_int4((int)((unsigned)e | (a << 28)), (int)kb[0], (int)kb[1], (int)kb[2]);
    }
}

// first corner of node i whose (global) element id is >= firstElem: start of a block's corners in the node's sorted list
__global__ void k_block_adj(long long N, const int* __restrict__ adjPtr, const int* __restrict__ adj, long long firstElem,
                            int* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int lo = adjPtr[i], hi = adjPtr[i + 1];
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (adj[mid] < firstElem) lo = mid + 1; else hi = mid;
    }
    out[i] = lo;
}

__global__ void k_max_int(const int* __restrict__ v, long long n, int* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicMax(out, v[i]);
}

// Row schedule of the row-owner kernel: inside every window of ROW_WINDOW consecutive nodes the nodes are ordered by
// falling corner count (stable), so that the 32 nodes a warp owns walk the same number of corners (a 9-node quad mesh
// alternates nodes with 4 / 2 / 1 adjacent elements: in node order a warp idles on a quarter of its corner iterations).
// iters[0] / iters[1]: sum over warps of the longest corner walk in node order / in the balanced order.
constexpr int ROW_WINDOW = 256;
__global__ void __launch_bounds__(ROW_WINDOW) k_row_order(long long N, const int* __restrict__ cFirst, const int* __restrict__ cLast,
                                                         int* __restrict__ order, unsigned long long* __restrict__ iters) {
    __shared__ int cnt[ROW_WINDOW];
    __shared__ int sorted[ROW_WINDOW];
    const int t = threadIdx.x;
    const long long i = (long long)blockIdx.x * ROW_WINDOW + t;
    const int mine = i < N ? cLast[i] - cFirst[i] : -1;  // padding slots sort to the end
    cnt[t] = mine;
    __syncthreads();
    int rank = 0;
    for (int u = 0; u < ROW_WINDOW; ++u) rank += (cnt[u] > mine) || (cnt[u] == mine && u < t);
    order[(long long)blockIdx.x * ROW_WINDOW + rank] = i < N ? (int)i : -1;
    sorted[rank] = mine;
    __syncthreads();
    const int a = __reduce_max_sync(0xffffffffu, max(mine, 0)), b = __reduce_max_sync(0xffffffffu, max(sorted[t], 0));
    if ((t & 31) == 0) {
        atomicAdd(iters, (unsigned long long)a);
        atomicAdd(iters + 1, (unsigned long long)b);
    }
}

// slot records of the balanced schedule: everything a thread needs before its corner loop in one coalesced 16-byte load
// {node (-1: padding), first value of the node row, degree | corners of the first block << 8, first corner of the first block}
__global__ void k_slot_records(long long slots, const int* __restrict__ order, const int* __restrict__ nrowptr,
                               const int* __restrict__ cFirst, const int* __restrict__ cLast, int4* __restrict__ rec) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= slots) return;
    const int i = order[s];
    if (i < 0) {
        rec[s] = make_int4(-1, 0, 0, 0);
        return;
    }
    const int off = nrowptr[i], c0 = cFirst[i];
    rec[s] = make_int4(i, off, (nrowptr[i + 1] - off) | ((cLast[i] - c0) << 8), c0);
}

// adds the position of each constrained node in its own sorted row to its bcInfo word: (index + 1) << 9 | diagK << 2 | bits
__global__ void k_bc_diag(long long N, const int* __restrict__ nrowptr, const int* __restrict__ ncols, int* __restrict__ info) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int v = info[i];
    if (!v) return;
    int diagK = 0;
    for (int c = nrowptr[i]; c < nrowptr[i + 1]; ++c) diagK += (ncols[c] < (int)i) ? 1 : 0;
    info[i] = ((v >> 2) << 9) | (diagK << 2) | (v & 3);
}

__global__ void k_check_bcs(long long nBC, int s, const int* __restrict__ bcDof, const int* __restrict__ nrowptr, int* flag) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nBC) return;
    const int i = bcDof[t] / s;
    if (nrowptr[i + 1] == nrowptr[i]) atomicExch(flag, 2);
}


// =============================================================================================
// table checks of 3- and 4-node blocks (host, at add_block)
// =============================================================================================
// Geometric sums S[a][b][d][D] = sum_p w_p / det_p * Gh_d,a Gh_D,b (Gh = adj(J) dN) of one fixed, irregular element:
// what the assembly kernels integrate, up to the constitutive matrix.
static void geometric_sums(int nn, int nq, const double* dN, const double* w, const double* X, double* S) {
    std::fill(S, S + nn * nn * 4, 0.0);
    for (int p = 0; p < nq; ++p) {
        const double* d0 = dN + (size_t)(p * 2) * nn;
        const double* d1 = d0 + nn;
        double J00 = 0, J01 = 0, J10 = 0, J11 = 0;
        for (int a = 0; a < nn; ++a) {
            J00 += d0[a] * X[2 * a];
            J01 += d0[a] * X[2 * a + 1];
            J10 += d1[a] * X[2 * a];
            J11 += d1[a] * X[2 * a + 1];
        }
        const double det = J00 * J11 - J01 * J10;
        for (int a = 0; a < nn; ++a)
            for (int b = 0; b < nn; ++b) {
                const double ga[2] = {J11 * d0[a] - J01 * d1[a], J00 * d1[a] - J10 * d0[a]};
                const double gb[2] = {J11 * d0[b] - J01 * d1[b], J00 * d1[b] - J10 * d0[b]};
                for (int d = 0; d < 2; ++d)
                    for (int D = 0; D < 2; ++D) S[((a * nn + b) * 2 + d) * 2 + D] += w[p] / det * ga[d] * gb[D];
            }
    }
}

// The rotated corner records of the row-owner kernel (k_corners) hand every thread its element renumbered cyclically so
// that the owning node is local node 0, and the kernel evaluates it with the caller's tables as they are.  That is exact
// only when basis and rule are invariant under a cyclic renumbering of the nodes (counter-clockwise node order, rotation-
// symmetric rule: the reference's Tri/Quad order 1).  Checked numerically on one irregular element: the matrix of the
// renumbered element must be the renumbered matrix.
static bool tables_cyclic_invariant(int nn, int nq, const double* dN, const double* w) {
    static const double X0[8] = {0.03, -0.02, 1.11, 0.07, 1.26, 0.93, -0.09, 1.17};  // (3 nodes: the first three)
    std::vector<double> S((size_t)nn * nn * 4), Sr(S.size()), Xr(2 * nn);
    geometric_sums(nn, nq, dN, w, X0, S.data());
    double scale = 0.0;
    for (double v : S) scale = std::max(scale, std::fabs(v));
    for (int r = 1; r < nn; ++r) {
        for (int a = 0; a < nn; ++a) {
            Xr[2 * a] = X0[2 * ((a + r) % nn)];
            Xr[2 * a + 1] = X0[2 * ((a + r) % nn) + 1];
        }
        geometric_sums(nn, nq, dN, w, Xr.data(), Sr.data());
        for (int a = 0; a < nn; ++a)
            for (int b = 0; b < nn; ++b)
                for (int k = 0; k < 4; ++k)
                    if (!(std::fabs(Sr[(a * nn + b) * 4 + k] - S[(((a + r) % nn) * nn + (b + r) % nn) * 4 + k]) <= 1e-12 * scale)) return false;
    }
    return true;
}

// true when a 4-node block carries the standard bilinear quad -- nodes counter-clockwise from (-1, -1), 2x2 Gauss rule
// (points (+-1/sqrt 3, +-1/sqrt 3) in any order, unit weights): QuadElement2D order 1 (FEMpy/Elements/QuadElement2D.py:
// 65-111,181-186) -- for which the row-owner kernel has a closed form
static bool standard_q4(int nq, const double* dN, const double* w) {
    if (nq != 4) return false;
    static const double xs[4] = {-1, 1, 1, -1}, es[4] = {-1, -1, 1, 1};
    const double g = 0.57735026918962576451, tol = 1e-13;
    int seen = 0;
    for (int p = 0; p < 4; ++p) {
        if (std::fabs(w[p] - 1.0) > tol) return false;
        const double* d0 = dN + (size_t)(p * 2) * 4;
        const double* d1 = d0 + 4;
        const double eta = 1.0 + 4.0 * d0[0], xi = 1.0 + 4.0 * d1[0];  // dN_0/dxi = -(1 - eta)/4, dN_0/deta = -(1 - xi)/4
        if (std::fabs(std::fabs(xi) - g) > tol || std::fabs(std::fabs(eta) - g) > tol) return false;
        for (int a = 0; a < 4; ++a)
            if (std::fabs(d0[a] - 0.25 * xs[a] * (1.0 + es[a] * eta)) > tol || std::fabs(d1[a] - 0.25 * es[a] * (1.0 + xs[a] * xi)) > tol) return false;
        seen |= 1 << ((xi > 0 ? 1 : 0) | (eta > 0 ? 2 : 0));
    }
    return seen == 15;
}


// Fan-ordered corner records of the write-once kernel (assemble_fan.cuh), 3- and 4-node blocks of single-block plans.
// The corners of a node are chained so that the last node of one corner is the first node of the next (consistently
// oriented elements around a node form such a fan: an edge is shared by the two elements on either side of it).  Walking
// the fan, a thread then completes every block of its row in registers -- own block: sum over all corners; edge
// neighbour: the last block of one corner + the first block of the next; diagonal neighbour: one corner -- and stores it
// ONCE: no zero-fill, no read-modify-write of shared-memory accumulators.
//   record: {k_0 | k_1 << 7 | .. | linkNext << 28 | closedLast << 29, node 1, node 2, node 3 (4 nodes) or element id (3 nodes)}
//   linkNext: the next corner of the node continues this one; closedLast (last corner only): it continues into the first.
// bad[0] is set when some node's corners cannot be ordered that way (a row position would receive two unlinked
// contributions: inconsistent element orientation, elements sharing two nodes but no edge); the plan then keeps the
// accumulating row-owner kernel.
constexpr int FAN_MAXC = 48;
__global__ void k_fan_corners(long long N, int nn, const int* __restrict__ cBegin, const int* __restrict__ cEnd,
                              const int4* __restrict__ rotRec, const int* __restrict__ rotElem, int4* __restrict__ fanRec,
                              int* __restrict__ fanElem, int* __restrict__ bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int c0 = cBegin[i], nc = cEnd[i] - c0;
    if (nc <= 0) return;
    if (nc > FAN_MAXC) {
        atomicExch(bad, 1);
        return;
    }
    int first[FAN_MAXC], last[FAN_MAXC], order[FAN_MAXC];
    unsigned slots[FAN_MAXC];
    bool used[FAN_MAXC], link[FAN_MAXC];
    for (int k = 0; k < nc; ++k) {
        const int4 r = rotRec[c0 + k];
        if (nn == 3) {  // {e, slots << 4, n1, n2}
            first[k] = r.z;
            last[k] = r.w;
            slots[k] = ((unsigned)r.y >> 4) & 0x1fffffu;
        } else {  // {slots, n1, n2, n3}
            first[k] = r.y;
            last[k] = r.w;
            slots[k] = (unsigned)r.x & 0x0fffffffu;
        }
        used[k] = false;
    }
    // chains: start at a corner nobody continues into (open fan); when none is left the rest are closed loops
    int placed = 0, chains = 0;
    bool closed = false;
    while (placed < nc) {
        int start = -1;
        for (int k = 0; k < nc && start < 0; ++k) {
            if (used[k]) continue;
            bool hasPred = false;
            for (int j = 0; j < nc; ++j) hasPred = hasPred || (!used[j] && j != k && last[j] == first[k]);
            if (!hasPred) start = k;
        }
        bool loop = false;
        if (start < 0) {
            loop = true;
            for (int k = 0; k < nc && start < 0; ++k)
                if (!used[k]) start = k;
        }
        ++chains;
        int cur = start;
        while (true) {
            used[cur] = true;
            order[placed] = cur;
            int next = -1;
            for (int j = 0; j < nc && next < 0; ++j)
                if (!used[j] && first[j] == last[cur]) next = j;
            link[placed] = next >= 0;
            ++placed;
            if (next < 0) break;
            cur = next;
        }
        if (loop && chains == 1 && placed == nc && last[cur] == first[start]) closed = true;
    }
    // every row position other than the node's own may be hit twice only through a link (or the closing wrap)
    const int diag = (int)(slots[0] & 0x7fu);
    bool ok = true;
    for (int a = 0; a < nc && ok; ++a) {
        const int ka = order[a];
        ok = ok && (int)(slots[ka] & 0x7fu) == diag;
        for (int ba = 1; ba < nn && ok; ++ba) {
            const int sa = (int)((slots[ka] >> (7 * ba)) & 0x7fu);
            if (sa == diag) ok = false;  // an element naming the node twice
            for (int b2 = a; b2 < nc && ok; ++b2) {
                const int kb = order[b2];
                for (int bb = 1; bb < nn && ok; ++bb) {
                    if (b2 == a && bb <= ba) continue;
                    const int sb = (int)((slots[kb] >> (7 * bb)) & 0x7fu);
                    if (sa != sb) continue;
                    const bool linked = (b2 == a + 1 && link[a] && ba == nn - 1 && bb == 1) ||
                                        (closed && a == 0 && b2 == nc - 1 && ba == 1 && bb == nn - 1 && nc > 1);
                    if (!linked) ok = false;
                }
            }
        }
    }
    if (closed && nc == 1) ok = false;
    if (!ok) {
        atomicExch(bad, 1);
        return;
    }
    for (int k = 0; k < nc; ++k) {
        const int4 r = rotRec[c0 + order[k]];
        const unsigned flags = (link[k] ? 1u << 28 : 0u) | ((closed && k == nc - 1) ? 1u << 29 : 0u);
        if (nn == 3) {
            fanRec[c0 + k] = make_int4((int)(slots[order[k]] | flags), r.z, r.w, r.x);
        } else {
            fanRec[c0 + k] = make_int4((int)(slots[order[k]] | flags), r.y, r.z, r.w);
            fanElem[c0 + k] = rotElem[c0 + order[k]];
        }
    }
}

// Compact 4-node fan records: node and element ids as signed 24-bit differences to the owning node's id, so that the
// element id (thickness index) rides in the same 16 bytes:
//   {slots | flags, d1 | d2 << 24, d2 >> 8 | d3 << 16, d3 >> 16 | de << 8}      d_k = node_k - i, de = element - i
// Numberings with some locality fit (|difference| < 2^23); otherwise the plan keeps the wide records + element array.
__global__ void k_fan_fits(long long N, const int* __restrict__ cBegin, const int* __restrict__ cEnd, const int4* __restrict__ fanRec,
                           const int* __restrict__ fanElem, int* __restrict__ bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int lim = 1 << 23;
    for (int c = cBegin[i]; c < cEnd[i]; ++c) {
        const int4 r = fanRec[c];
        const long long d[4] = {r.y - i, r.z - i, r.w - i, fanElem[c] - i};
        for (int k = 0; k < 4; ++k)
            if (d[k] < -lim || d[k] >= lim) atomicExch(bad, 1);
    }
}
__global__ void k_fan_compact(long long N, const int* __restrict__ cBegin, const int* __restrict__ cEnd, int4* __restrict__ fanRec,
                              const int* __restrict__ fanElem) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    for (int c = cBegin[i]; c < cEnd[i]; ++c) {
        const int4 r = fanRec[c];
        const unsigned d1 = (unsigned)(r.y - (int)i) & 0xffffffu, d2 = (unsigned)(r.z - (int)i) & 0xffffffu;
        const unsigned d3 = (unsigned)(r.w - (int)i) & 0xffffffu, de = (unsigned)(fanElem[c] - (int)i) & 0xffffffu;
        fanRec[c] = make_int4(r.x, (int)(d1 | (d2 << 24)), (int)((d2 >> 8) | (d3 << 16)), (int)((d3 >> 16) | (de << 8)));
    }
}

// Kernel-side layout of the fan records (assemble_fan.cuh: Args): per node a header {first row position, index of its
// further records, corners | degree << 8, BC word} followed by the record of its first corner (rw int4: 2 for wide 4-node
// records, whose second word carries the element id), and the records of all further corners in `rest`.  The node count is
// padded to a multiple of 32: the kernel's tail lanes read {end of the values, 0, 0, 0}.
__global__ void k_fan_rest_count(long long N, const int* __restrict__ cBegin, const int* __restrict__ cEnd, int* __restrict__ cnt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) cnt[i] = max(cEnd[i] - cBegin[i] - 1, 0);
}
__global__ void k_fan_headers(long long N, long long Npad, int rw, const int* __restrict__ nrowptr, const int* __restrict__ cBegin,
                              const int* __restrict__ cEnd, const int* __restrict__ restPtr, const int4* __restrict__ fanRec,
                              const int* __restrict__ fanElem, int4* __restrict__ hdr, int4* __restrict__ rest) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Npad) return;
    int4* h = hdr + (1 + rw) * i;
    const int4 zero = make_int4(0, 0, 0, 0);
    if (i >= N) {
        h[0] = make_int4(nrowptr[N], 0, 0, 0);
        h[1] = zero;
        if (rw == 2) h[2] = zero;
        return;
    }
    const int c0 = cBegin[i], nc = cEnd[i] - c0, off = nrowptr[i], deg = nrowptr[i + 1] - off, r0 = restPtr[i];
    h[0] = make_int4(off, r0, nc | (deg << 8), 0);
    h[1] = nc > 0 ? fanRec[c0] : zero;
    if (rw == 2) h[2] = make_int4(nc > 0 ? fanElem[c0] : 0, 0, 0, 0);
    for (int k = 1; k < nc; ++k) {
        rest[(long long)rw * (r0 + k - 1)] = fanRec[c0 + k];
        if (rw == 2) rest[(long long)rw * (r0 + k - 1) + 1] = make_int4(fanElem[c0 + k], 0, 0, 0);
    }
}
// Compact 3-node fan records (8 bytes): in a single chain the first node of a corner is the last node of the one before it,
// so a further corner only names its second node and its element -- as signed differences to the owner's id:
//   x = slots (21 bits) | (node 2 - i) bits 9..15 << 21 | flags (bits 28, 29)       y = (node 2 - i) bits 0..8 | (element - i) << 9
// Fits when every node's corners form ONE chain and |node 2 - i| < 2^15, |element - i| < 2^22 (checked here); else the
// plan keeps the 16-byte records.
__global__ void k_fan_fits3(long long N, const int* __restrict__ cBegin, const int* __restrict__ cEnd, const int4* __restrict__ fanRec,
                            int* __restrict__ bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int c0 = cBegin[i], c1 = cEnd[i];
    for (int c = c0; c < c1; ++c) {
        const int4 r = fanRec[c];  // {slots | flags, node 1, node 2, element}
        if (c + 1 < c1 && !((unsigned)r.x & (1u << 28))) atomicExch(bad, 1);  // chain break: the next corner starts a new fan
        if (c > c0) {
            const long long d2 = r.z - i, de = r.w - i;
            if (d2 < -(1 << 15) || d2 >= (1 << 15) || de < -(1 << 22) || de >= (1 << 22)) atomicExch(bad, 1);
        }
    }
}
__global__ void k_fan_rest3(long long N, const int* __restrict__ cBegin, const int* __restrict__ cEnd, const int* __restrict__ restPtr,
                            const int4* __restrict__ fanRec, int2* __restrict__ rest) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int c0 = cBegin[i], nc = cEnd[i] - c0, r0 = restPtr[i];
    for (int k = 1; k < nc; ++k) {
        const int4 r = fanRec[c0 + k];
        const unsigned d2 = (unsigned)(r.z - (int)i) & 0xffffu, de = (unsigned)(r.w - (int)i) & 0x7fffffu;
        rest[r0 + k - 1] = make_int2((int)(((unsigned)r.x & 0x301fffffu) | ((d2 >> 9) << 21)), (int)((d2 & 0x1ffu) | (de << 9)));
    }
}

__global__ void k_fan_header_bcs(long long N, int stride, const int* __restrict__ bcInfo, int4* __restrict__ hdr) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) hdr[stride * i].w = bcInfo[i];
}

// standard linear triangle: N = (1 - xi - eta, xi, eta), centroid rule with weight 1/2 (TriElement2D order 1,
// FEMpy/Elements/TriElement2D.py:69-169, Quadrature/TriQuad.py:78-86)
static bool standard_t3(int nq, const double* dN, const double* w) {
    if (nq != 1 || std::fabs(w[0] - 0.5) > 1e-14) return false;
    static const double ref[6] = {-1, 1, 0, -1, 0, 1};
    for (int k = 0; k < 6; ++k)
        if (std::fabs(dN[k] - ref[k]) > 1e-14) return false;
    return true;
}

// =============================================================================================
// C ABI: plan
// =============================================================================================
extern "C" int femb_plan_create(int64_t numNodes, int32_t numDim, int32_t numStates, int32_t device, femb_plan** out) {
    if (!out) return fail("femb_plan_create: out is NULL");
    *out = nullptr;
    if (numDim < 1 || numDim > 3 || numStates != numDim)
        return fail("femb_plan_create: supported are 1-D, 2-D and 3-D models with one displacement state per dimension (got numDim=%d, "
                    "numStates=%d)", numDim, numStates);
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

static int add_block(femb_plan* p, int32_t nn, int32_t nq, const double* N, const double* dNdxi, const double* w, int64_t numElems,
                     const int64_t* conn, bool connOnDevice);

extern "C" int femb_plan_add_block(femb_plan* p, int32_t nn, int32_t nq, const double* N, const double* dNdxi,
                                   const double* w, int64_t numElems, const int64_t* conn) {
    return add_block(p, nn, nq, N, dNdxi, w, numElems, conn, false);
}

// the same with the connectivity already on the plan's device (mesh generated or ingested there): no host copy of the mesh
extern "C" int femb_plan_add_block_dev(femb_plan* p, int32_t nn, int32_t nq, const double* N, const double* dNdxi,
                                       const double* w, int64_t numElems, const int64_t* d_conn) {
    return add_block(p, nn, nq, N, dNdxi, w, numElems, d_conn, true);
}

static int add_block(femb_plan* p, int32_t nn, int32_t nq, const double* N, const double* dNdxi, const double* w, int64_t numElems,
                     const int64_t* conn, bool connOnDevice) {
    if (!p) return fail("femb_plan_add_block: plan is NULL");
    if (p->finalized) return fail("femb_plan_add_block: plan already finalized");
    if ((int)p->blocks.size() >= MAX_BLOCKS) return fail("femb_plan_add_block: at most %d element types", MAX_BLOCKS);
    const bool generic = p->numDim != 2;  // 1-D / 3-D: any Lagrange family of up to 27 nodes (generic.cu)
    if (generic ? !(nn >= 2 && nn <= 27 && nq >= 1 && nq <= 125) : (!family_supported(nn, nq) && !(p->pointsOnly && nn >= 1 && nn <= 64 && nq >= 1)))
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
    b.dN.assign(dNdxi, dNdxi + (size_t)nq * p->numDim * nn);
    b.w.assign(w, w + nq);
    if (!generic) {
        if (nn <= 4) b.cyclic = tables_cyclic_invariant(nn, nq, dNdxi, w);
        b.closedForm = b.cyclic && ((nn == 4 && standard_q4(nq, dNdxi, w)) || (nn == 3 && standard_t3(nq, dNdxi, w)));
    } else {
        CUDA_TRY(cudaMalloc(&b.d_tabN, b.N.size() * sizeof(double)));
        CUDA_TRY(cudaMalloc(&b.d_tabdN, b.dN.size() * sizeof(double)));
        CUDA_TRY(cudaMalloc(&b.d_tabw, b.w.size() * sizeof(double)));
        CUDA_TRY(cudaMemcpyAsync(b.d_tabN, b.N.data(), b.N.size() * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        CUDA_TRY(cudaMemcpyAsync(b.d_tabdN, b.dN.data(), b.dN.size() * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        CUDA_TRY(cudaMemcpyAsync(b.d_tabw, b.w.data(), b.w.size() * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    }
    Scratch<long long> d_conn64;
    const size_t cnt = (size_t)numElems * nn;
    CUDA_TRY(d_conn64.alloc(cnt));
    CUDA_TRY(cudaMalloc(&b.d_conn, cnt * sizeof(int)));
    CUDA_TRY(cudaMemcpyAsync(d_conn64, conn, cnt * sizeof(long long), connOnDevice ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemsetAsync(p->d_flag, 0, sizeof(int), p->stream));
    k_convert_conn<<<grid_for((int64_t)cnt, 256), 256, 0, p->stream>>>(d_conn64, b.d_conn, numElems, nn, p->numNodes,
                                                                        p->d_flag);
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, p->d_flag, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (flag) {
        cudaFree(b.d_conn);
        return fail("femb_plan_add_block: connectivity refers to a node outside [0, %lld)", (long long)p->numNodes);
    }
    p->numElems += numElems;
    p->numPoints += numElems * nq;
    p->keTotal += numElems * (int64_t)(p->numStates * nn) * (p->numStates * nn);
    p->reTotal += numElems * (int64_t)nn * p->numStates;
    p->blocks.push_back(std::move(b));
    return 0;
}

extern "C" int femb_plan_finalize(femb_plan* p) {
    if (!p) return fail("femb_plan_finalize: plan is NULL");
    if (p->finalized) return 0;
    if (p->blocks.empty()) return fail("femb_plan_finalize: no element blocks");
    TRY(use_device(p));
    cudaStream_t st = p->stream;
    ScopedEvent ev0, ev1;
    CUDA_TRY(ev0.create());
    CUDA_TRY(ev1.create());
    CUDA_TRY(cudaEventRecord(ev0, st));
    const int64_t N = p->numNodes;
    int64_t launches = 0;
    Scratch<int> d_adjCount, d_candCount, d_candPtr, d_cursor, d_cand, d_deg;  // plan-time temporaries (freed on every path)
    int*& d_adjPtr = p->d_adjPtr;
    int*& d_adj = p->d_adj;
    CUDA_TRY(d_adjCount.alloc(N + 1));
    CUDA_TRY(d_candCount.alloc(N + 1));
    CUDA_TRY(cudaMalloc(&d_adjPtr, (N + 1) * sizeof(int)));
    CUDA_TRY(d_candPtr.alloc(N + 1));
    CUDA_TRY(d_cursor.alloc(N + 1));
    CUDA_TRY(d_deg.alloc(N + 1));
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
    CUDA_TRY(d_cand.alloc((size_t)totalCand));
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
    k_sort_adj<<<grid_for(N, 128), 128, 0, st>>>(N, d_adjPtr, d_adj);
    launches++;
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
    {  // largest node degree + element-major connectivity for the row-owner kernel
        CUDA_TRY(cudaMemsetAsync(p->d_flag, 0, sizeof(int), st));
        k_max_int<<<grid_for(N, 256), 256, 0, st>>>(d_deg, N, p->d_flag);
        CUDA_TRY(cudaMemcpyAsync(&p->maxDeg, p->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
        launches++;
        bool rowsOk = p->numDim == 2;  // the row-owner / fan kernels are 2-D
        for (auto& b : p->blocks) rowsOk = rowsOk && b.nn <= 12 && b.numElems < (1ll << 28) && default_rule(b.nn, b.nq);
        if (rowsOk) {
            // row-owner kernel: per block element-major connectivity, corner records (indexed by the global corner
            // number) and, per node, where the block's corners start in the node's adjacency list (the list is sorted
            // by global element id and a block is a contiguous id range)
            const size_t nb = p->blocks.size();
            CUDA_TRY(cudaMalloc(&p->d_blockAdj, (nb + 1) * (size_t)N * sizeof(int)));
            for (size_t k = 0; k <= nb; ++k) {
                const long long firstElem = k < nb ? p->blocks[k].elemOffset : (long long)p->numElems;
                k_block_adj<<<grid_for(N, 256), 256, 0, st>>>(N, d_adjPtr, d_adj, firstElem, p->d_blockAdj + k * (size_t)N);
                launches++;
            }
            for (auto& b : p->blocks) {
                const int cw = (b.nn + 3) / 4;
                CUDA_TRY(cudaMalloc(&b.d_connE, b.numElems * cw * sizeof(int4)));
                k_conn_elem_major<<<grid_for(b.numElems, 256), 256, 0, st>>>(b.d_conn, b.numElems, b.nn, b.d_connE);
                CUDA_TRY(cudaMalloc(&b.d_corners, std::max<int64_t>(totalAdj, 1) * ((b.nn == 4 && !FEMB_ROT4) ? sizeof(int2) : sizeof(int4))));
                if (b.nn == 4 && FEMB_ROT4) CUDA_TRY(cudaMalloc(&b.d_cornerElem, std::max<int64_t>(totalAdj, 1) * sizeof(int)));
                const size_t k = &b - &p->blocks[0];
                k_corners<<<grid_for(N, 128), 128, 0, st>>>(N, p->d_nrowptr, p->d_ncols, p->d_blockAdj + k * (size_t)N,
                                                           p->d_blockAdj + (k + 1) * (size_t)N, d_adj, b.d_connE, b.nn,
                                                           (int)b.elemOffset, b.d_corners, b.d_cornerElem);
                launches += 2;
            }
            p->rowsBuilt = true;
            if (nb == 1 && p->blocks[0].nn <= 4 && p->blocks[0].closedForm && (p->blocks[0].nn == 3 || FEMB_ROT4)) {
                auto& b = p->blocks[0];
                CUDA_TRY(cudaMalloc(&b.d_fanRec, std::max<int64_t>(totalAdj, 1) * sizeof(int4)));
                if (b.nn == 4) CUDA_TRY(cudaMalloc(&b.d_fanElem, std::max<int64_t>(totalAdj, 1) * sizeof(int)));
                CUDA_TRY(cudaMemsetAsync(p->d_flag, 0, sizeof(int), st));
                k_fan_corners<<<grid_for(N, 128), 128, 0, st>>>(N, b.nn, p->d_blockAdj, p->d_blockAdj + (size_t)N, (const int4*)b.d_corners,
                                                               b.d_cornerElem, b.d_fanRec, b.d_fanElem, p->d_flag);
                launches++;
                int bad = 0;
                CUDA_TRY(cudaMemcpyAsync(&bad, p->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                b.fanOK = bad == 0;
                if (!b.fanOK) {
                    dfree(b.d_fanRec);
                    dfree(b.d_fanElem);
                } else if (b.nn == 3 && p->fanCompactAllowed) {  // 8-byte records where every node has one chain and the numbering is local
                    CUDA_TRY(cudaMemsetAsync(p->d_flag, 0, sizeof(int), st));
                    k_fan_fits3<<<grid_for(N, 128), 128, 0, st>>>(N, p->d_blockAdj, p->d_blockAdj + (size_t)N, b.d_fanRec, p->d_flag);
                    CUDA_TRY(cudaMemcpyAsync(&bad, p->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
                    CUDA_TRY(cudaStreamSynchronize(st));
                    launches++;
                    b.fanCompact = !bad;
                } else if (b.nn == 4 && p->fanCompactAllowed) {  // element id into the record when the numbering has the locality for it
                    CUDA_TRY(cudaMemsetAsync(p->d_flag, 0, sizeof(int), st));
                    k_fan_fits<<<grid_for(N, 128), 128, 0, st>>>(N, p->d_blockAdj, p->d_blockAdj + (size_t)N, b.d_fanRec, b.d_fanElem, p->d_flag);
                    CUDA_TRY(cudaMemcpyAsync(&bad, p->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
                    CUDA_TRY(cudaStreamSynchronize(st));
                    launches++;
                    if (!bad) {
                        k_fan_compact<<<grid_for(N, 128), 128, 0, st>>>(N, p->d_blockAdj, p->d_blockAdj + (size_t)N, b.d_fanRec, b.d_fanElem);
                        CUDA_TRY(cudaStreamSynchronize(st));
                        launches++;
                        dfree(b.d_fanElem);
                        b.fanCompact = true;
                    }
                }
                if (b.fanOK) {  // kernel-side layout: per-node headers + first record, further records behind
                    const int rw = (b.nn == 4 && !b.fanCompact) ? 2 : 1;
                    const long long Npad = (N + 31) / 32 * 32;
                    Scratch<int> d_cnt, d_restPtr;
                    CUDA_TRY(d_cnt.alloc(N + 1));
                    CUDA_TRY(d_restPtr.alloc(N + 1));
                    k_fan_rest_count<<<grid_for(N, 256), 256, 0, st>>>(N, p->d_blockAdj, p->d_blockAdj + (size_t)N, d_cnt);
                    launches++;
                    int rc = exclusive_scan(d_cnt, d_restPtr, N, st, launches);
                    int restTotal = 0;
                    if (!rc && cudaMemcpyAsync(&restTotal, d_restPtr + N, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = 1;
                    if (!rc && cudaStreamSynchronize(st) != cudaSuccess) rc = 1;
                    b.fanRestTotal = restTotal;
                    if (!rc && cudaMalloc(&b.d_fanHdr, (size_t)Npad * (1 + rw) * sizeof(int4)) != cudaSuccess) rc = 1;
                    if (!rc && cudaMalloc(&b.d_fanRest, std::max<size_t>((size_t)restTotal * rw, 1) * sizeof(int4)) != cudaSuccess) rc = 1;
                    if (!rc) {
                        k_fan_headers<<<grid_for(Npad, 128), 128, 0, st>>>(N, Npad, rw, p->d_nrowptr, p->d_blockAdj, p->d_blockAdj + (size_t)N,
                                                                         d_restPtr, b.d_fanRec, b.d_fanElem, b.d_fanHdr, b.d_fanRest);
                        launches++;
                        if (b.nn == 3 && b.fanCompact) {  // the further corners again, as 8-byte records over the same allocation
                            k_fan_rest3<<<grid_for(N, 128), 128, 0, st>>>(N, p->d_blockAdj, p->d_blockAdj + (size_t)N, d_restPtr, b.d_fanRec,
                                                                          (int2*)b.d_fanRest);
                            launches++;
                        }
                        if (cudaStreamSynchronize(st) != cudaSuccess) rc = 1;
                    }
                    dfree(b.d_fanRec);  // the kernel reads headers + rest only
                    dfree(b.d_fanElem);
                    if (rc) return fail("femb_plan_finalize: building the fan headers failed (%s)", cudaGetErrorString(cudaGetLastError()));
                }
            }
            {  // balanced row schedule: kept only when it saves more than 8 % of the warps' corner iterations
                const long long windows = (N + ROW_WINDOW - 1) / ROW_WINDOW;
                Scratch<unsigned long long> d_iters;
                unsigned long long iters[2] = {0, 0};
                CUDA_TRY(cudaMalloc(&p->d_rowOrder, windows * ROW_WINDOW * sizeof(int)));
                CUDA_TRY(d_iters.alloc(2));
                CUDA_TRY(cudaMemsetAsync(d_iters, 0, sizeof(iters), st));
                k_row_order<<<(unsigned)windows, ROW_WINDOW, 0, st>>>(N, p->d_blockAdj, p->d_blockAdj + nb * (size_t)N, p->d_rowOrder, d_iters);
                launches++;
                CUDA_TRY(cudaMemcpyAsync(iters, d_iters, sizeof(iters), cudaMemcpyDeviceToHost, st));
                CUDA_TRY(cudaStreamSynchronize(st));
                p->rowSlots = windows * ROW_WINDOW;
                p->rowIters[0] = (long long)iters[0];
                p->rowIters[1] = (long long)iters[1];
                p->rowsBalanced = p->balanceRows != 0 && (p->balanceRows == 2 || (double)iters[1] < 0.92 * (double)iters[0]);
                if (p->rowsBalanced) {
                    CUDA_TRY(cudaMalloc(&p->d_slotRec, p->rowSlots * sizeof(int4)));
                    k_slot_records<<<grid_for(p->rowSlots, 256), 256, 0, st>>>(p->rowSlots, p->d_rowOrder, p->d_nrowptr, p->d_blockAdj,
                                                                             p->d_blockAdj + (size_t)N, p->d_slotRec);
                    launches++;
                    CUDA_TRY(cudaStreamSynchronize(st));
                }
                cudaFree(p->d_rowOrder);
                p->d_rowOrder = nullptr;
            }
        }
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(ev1, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    cudaEventElapsedTime(&p->buildMs, ev0, ev1);
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
    dfree(p->d_adjPtr);
    dfree(p->d_adj);
    dfree(p->d_blockAdj);
    dfree(p->d_rowOrder);
    if (p->s_solve) cudaFree(p->s_solve);
    p->s_solve = nullptr;
    if (p->d_slotRec) cudaFree(p->d_slotRec);
    p->d_slotRec = nullptr;
    for (auto& b : p->blocks) {
        dfree(b.d_connE);
        if (b.d_corners) cudaFree(b.d_corners);
        b.d_corners = nullptr;
        if (b.d_cornerElem) cudaFree(b.d_cornerElem);
        b.d_cornerElem = nullptr;
        dfree(b.d_fanRec);
        dfree(b.d_fanElem);
        dfree(b.d_fanHdr);
        dfree(b.d_fanRest);
        dfree(b.d_tabN);
        dfree(b.d_tabdN);
        dfree(b.d_tabw);
    }
    dfree(p->d_bcMask);
    dfree(p->d_bcInfo);
    dfree(p->d_bcNodeCnt);
    dfree(p->d_bcNodeVal);
    dfree(p->d_rowStats);
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

extern "C" int femb_plan_set_stream(femb_plan* p, void* cudaStream, int32_t external) {
    if (!p) return fail("femb_plan_set_stream: plan is NULL");
    p->stream = external ? (cudaStream_t)cudaStream : p->ownStream;  // external NULL = the legacy default stream
    return 0;
}

extern "C" int femb_plan_set_option(femb_plan* p, const char* name, int64_t value) {
    if (!p || !name) return fail("femb_plan_set_option: NULL argument");
    if (!strcmp(name, "assembly_path")) {
        if (value != 0 && value != 2) return fail("femb_plan_set_option: assembly_path must be 0 (scatter) or 2 (row owner)");
        p->assemblyPath = (int)value;
    } else if (!strcmp(name, "fan_compact")) {
        if (p->finalized) return fail("femb_plan_set_option: fan_compact must be set before finalize");
        p->fanCompactAllowed = value != 0;
    } else if (!strcmp(name, "points_only")) {
        // blocks with any number of nodes / points: such a plan serves femb_point_quantities only (every other compute
        // entry rejects a family it has no kernel for)
        if (!p->blocks.empty()) return fail("femb_plan_set_option: points_only must be set before the first block");
        p->pointsOnly = value != 0;
    } else if (!strcmp(name, "fan_prefetch")) {
        if (value < -1) return fail("femb_plan_set_option: fan_prefetch must be -1 (default), 0 (off) or a distance in 32-node warps");
        p->fanPrefetch = (int)value;
    } else if (!strcmp(name, "fan_carveout")) {
        if (value < -1 || value > 100) return fail("femb_plan_set_option: fan_carveout is a percentage (or -1 for the default)");
        p->fanCarveout = (int)value;
    } else if (!strcmp(name, "fan_rows")) {
        if (value != 0 && value != 1) return fail("femb_plan_set_option: fan_rows must be 0 or 1");
        p->useFan = value != 0;
    } else if (!strcmp(name, "balance_rows")) {
        if (p->finalized) return fail("femb_plan_set_option: balance_rows must be set before finalize");
        if (value < 0 || value > 2) return fail("femb_plan_set_option: balance_rows must be 0 (never), 1 (when it pays) or 2 (always)");
        p->balanceRows = (int)value;
    } else {
        return fail("femb_plan_set_option: unknown option %s", name);
    }
    return 0;
}

extern "C" int64_t femb_plan_get_info(const femb_plan* p, const char* name) {
    if (!p || !name) return -1;
    if (!strcmp(name, "assembly_path")) return p->assemblyPath;
    if (!strcmp(name, "rows_path")) return rows_path_available(p) ? 1 : 0;
    if (!strcmp(name, "max_degree")) return p->maxDeg;
    if (!strcmp(name, "rows_balanced")) return p->rowsBalanced ? 1 : 0;
    if (!strcmp(name, "fan_path")) return fan_path_available(p) ? 1 : 0;
    if (!strcmp(name, "fan_compact")) return (!p->blocks.empty() && p->blocks[0].fanCompact) ? 1 : 0;
    if (!strcmp(name, "closed_form_q4")) {
        bool any = false, all = true;
        for (auto& b : p->blocks)
            if (b.nn == 4) {
                any = true;
                all = all && b.closedForm;
            }
        return (any && all) ? 1 : 0;
    }
    if (!strcmp(name, "row_iters_node_order")) return p->rowIters[0];
    if (!strcmp(name, "row_iters_balanced")) return p->rowIters[1];
    return -1;
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
    Scratch<long long> d_ip, d_ix;
    CUDA_TRY(d_ip.alloc(nd + 1));
    CUDA_TRY(d_ix.alloc((size_t)nnz));
    int rc = femb_plan_pattern_dev(p, (int64_t*)d_ip.p, (int64_t*)d_ix.p);
    if (!rc) {
        cudaError_t e1 = cudaMemcpyAsync(indptr, d_ip, (nd + 1) * sizeof(long long), cudaMemcpyDeviceToHost, p->stream);
        cudaError_t e2 = cudaMemcpyAsync(indices, d_ix, nnz * sizeof(long long), cudaMemcpyDeviceToHost, p->stream);
        cudaError_t e3 = cudaStreamSynchronize(p->stream);
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) rc = fail("femb_plan_pattern: copy back failed");
    }
    return rc;
}

// The element -> CSR slot map of the scatter kernels (4*nn^2 bytes per element) is only built when those kernels are
// going to run: the row-owner path never reads it.
int ensure_slot_maps(femb_plan* p) {
    for (auto& b : p->blocks) {
        if (b.d_slot) continue;
        const int64_t cnt = b.numElems * (int64_t)b.nn * b.nn;
        CUDA_TRY(cudaMalloc(&b.d_slot, cnt * sizeof(int)));
        k_build_slots<<<grid_for(cnt, 256), 256, 0, p->stream>>>(b.d_conn, b.numElems, b.nn, p->d_nrowptr, p->d_ncols, b.d_slot);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// =============================================================================================
// C ABI: boundary conditions
// =============================================================================================
extern "C" int femb_plan_set_bcs(femb_plan* p, const int64_t* bcDof, const double* bcVal, int64_t nBC) {
    TRY(need_final(p, "femb_plan_set_bcs"));
    TRY(use_device(p));
    // The host-pointer API passes the BC lists on every call (the reference rebuilds them per assembly,
    // Problem.py:471-483); identical lists are recognised and cost nothing after the first call.
    if (nBC > 0 && bcDof && bcVal && p->bcValid && (int64_t)p->bcRawDof.size() == nBC &&
        !memcmp(p->bcRawDof.data(), bcDof, nBC * sizeof(int64_t)) && !memcmp(p->bcRawVal.data(), bcVal, nBC * sizeof(double)))
        return 0;
    p->bcValid = false;
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
    {  // ascending DOF order: the fused kernel looks constrained DOFs up by binary search
        std::vector<size_t> perm(n);
        for (size_t i = 0; i < n; ++i) perm[i] = i;
        std::sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return dofs[a] < dofs[b]; });
        std::vector<int> d2(n);
        std::vector<double> v2(n), c2(n);
        for (size_t i = 0; i < n; ++i) {
            d2[i] = dofs[perm[i]];
            v2[i] = vals[perm[i]];
            c2[i] = counts[perm[i]];
        }
        dofs.swap(d2);
        vals.swap(v2);
        counts.swap(c2);
    }
    std::vector<unsigned char> mask((size_t)p->numNodes, 0);
    if (p->numStates == 2) {
        for (size_t i = 0; i < n; ++i) mask[dofs[i] >> 1] |= (unsigned char)(1u << (dofs[i] & 1));
        TRY(dalloc(&p->d_bcMask, (size_t)p->numNodes));
        CUDA_TRY(cudaMemcpyAsync(p->d_bcMask, mask.data(), mask.size(), cudaMemcpyHostToDevice, p->stream));
    }
    if (p->numStates == 2) {  // row-owner kernel: per node (index of the constrained node + 1) << 9 | diagK << 2 | DOF bits (0 = unconstrained) and, per
       // constrained node, the diagonal values (occurrence counts) and the prescribed values of its two DOFs
        std::vector<int> info((size_t)p->numNodes, 0);
        std::vector<double> cnt2, val2;
        for (size_t i = 0; i < n; ++i) {
            const size_t node = (size_t)(dofs[i] >> 1);
            if (!info[node]) {
                info[node] = (int)((cnt2.size() / 2 + 1) << 2) | mask[node];
                cnt2.push_back(0.0); cnt2.push_back(0.0);
                val2.push_back(0.0); val2.push_back(0.0);
            }
            const size_t j = (size_t)(info[node] >> 2) - 1;
            cnt2[2 * j + (dofs[i] & 1)] = counts[i];
            val2[2 * j + (dofs[i] & 1)] = vals[i];
        }
        if (cnt2.size() / 2 >= (1u << 22)) return fail("femb_plan_set_bcs: more than 2^22 constrained nodes do not fit the per-node BC word");
        dfree(p->d_bcInfo); dfree(p->d_bcNodeCnt); dfree(p->d_bcNodeVal);
        TRY(dalloc(&p->d_bcInfo, (size_t)p->numNodes));
        TRY(dalloc(&p->d_bcNodeCnt, cnt2.size()));
        TRY(dalloc(&p->d_bcNodeVal, val2.size()));
        CUDA_TRY(cudaMemcpyAsync(p->d_bcInfo, info.data(), info.size() * sizeof(int), cudaMemcpyHostToDevice, p->stream));
        CUDA_TRY(cudaMemcpyAsync(p->d_bcNodeCnt, cnt2.data(), cnt2.size() * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        CUDA_TRY(cudaMemcpyAsync(p->d_bcNodeVal, val2.data(), val2.size() * sizeof(double), cudaMemcpyHostToDevice, p->stream));
        k_bc_diag<<<grid_for(p->numNodes, 256), 256, 0, p->stream>>>(p->numNodes, p->d_nrowptr, p->d_ncols, p->d_bcInfo);
        for (auto& b : p->blocks)  // the fan kernel reads the BC word from its per-node header
            if (b.d_fanHdr)
                k_fan_header_bcs<<<grid_for(p->numNodes, 256), 256, 0, p->stream>>>(p->numNodes, 1 + ((b.nn == 4 && !b.fanCompact) ? 2 : 1),
                                                                                  p->d_bcInfo, b.d_fanHdr);
        CUDA_TRY(cudaStreamSynchronize(p->stream));  // the host vectors go out of scope
    }
    CUDA_TRY(cudaMalloc(&p->d_bcDof, n * sizeof(int)));
    CUDA_TRY(cudaMalloc(&p->d_bcVal, n * sizeof(double)));
    CUDA_TRY(cudaMalloc(&p->d_bcCount, n * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(p->d_bcDof, dofs.data(), n * sizeof(int), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->d_bcVal, vals.data(), n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemcpyAsync(p->d_bcCount, counts.data(), n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    CUDA_TRY(cudaMemsetAsync(p->d_flag, 0, sizeof(int), p->stream));
    k_check_bcs<<<grid_for((int64_t)n, 256), 256, 0, p->stream>>>((long long)n, p->numStates, p->d_bcDof, p->d_nrowptr, p->d_flag);
    int flag = 0;
    CUDA_TRY(cudaMemcpyAsync(&flag, p->d_flag, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    if (flag) return fail("femb_plan_set_bcs: a BC is applied to a node that belongs to no element (no diagonal entry in the pattern)");
    p->nBC = (int64_t)n;
    p->bcRawDof.assign(bcDof, bcDof + nBC);
    p->bcRawVal.assign(bcVal, bcVal + nBC);
    p->bcValid = true;
    return 0;
}
