// Patch programs for the write-once assembly (built once per plan, on the GPU).
//
// The node set is cut into patches of `patchNodes` consecutive nodes along a locality-preserving order
// (Morton order of the hint coordinates when the caller supplied them, node-id order otherwise).  For
// every element block the builder then records, per patch,
//   * the elements touching an owned node (sorted; evaluated by that patch's CTA, halo elements included),
//   * a patch-contiguous copy of their connectivity, and
//   * for every (owned node, neighbour node) block of the global matrix the list of element sub-blocks
//     that sum into it ("sources": element slot, sub-block id, transpose bit -- 16 bits each).
// The assembly kernel (assemble.cu: k_assemble_patch) needs no atomics and no zero-fill: every CSR value
// and residual entry is written exactly once.
#include <cub/device/device_radix_sort.cuh>

#include "femb_internal.h"

static constexpr int DEFAULT_PATCH_NODES = 128;
static constexpr int MAX_PATCH_ELEMS = 1024;

bool patch_family_supported(int nn, int nq) { return (nn == 3 && nq == 1) || (nn == 4 && nq == 4); }

static int blk_bits(int nn) {
    const int nblk = nn + nn * (nn - 1) / 2;
    int b = 1;
    while ((1 << b) < nblk) ++b;
    return b;
}

// ---------------------------------------------------------------------------------------------
// node order
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned spread_bits16(unsigned v) {
    v &= 0xffffu;
    v = (v | (v << 8)) & 0x00ff00ffu;
    v = (v | (v << 4)) & 0x0f0f0f0fu;
    v = (v | (v << 2)) & 0x33333333u;
    v = (v | (v << 1)) & 0x55555555u;
    return v;
}

__global__ void k_morton_keys(const double* __restrict__ coords, long long N, double minx, double miny, double sx,
                              double sy, unsigned* __restrict__ keys, int* __restrict__ ids) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double qx = (coords[2 * i] - minx) * sx, qy = (coords[2 * i + 1] - miny) * sy;
    const unsigned ix = (unsigned)fmin(fmax(qx, 0.0), 65535.0), iy = (unsigned)fmin(fmax(qy, 0.0), 65535.0);
    keys[i] = spread_bits16(ix) | (spread_bits16(iy) << 1);
    ids[i] = (int)i;
}

__global__ void k_iota(int* __restrict__ ids, long long N) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) ids[i] = (int)i;
}

__global__ void k_minmax(const double* __restrict__ coords, long long N, double* __restrict__ out) {
    // out[0..3] = minx, miny, maxx, maxy; one block, strided (plan-time only)
    __shared__ double s[4][256];
    double mnx = 1e300, mny = 1e300, mxx = -1e300, mxy = -1e300;
    for (long long i = threadIdx.x; i < N; i += blockDim.x) {
        mnx = fmin(mnx, coords[2 * i]);
        mxx = fmax(mxx, coords[2 * i]);
        mny = fmin(mny, coords[2 * i + 1]);
        mxy = fmax(mxy, coords[2 * i + 1]);
    }
    s[0][threadIdx.x] = mnx;
    s[1][threadIdx.x] = mny;
    s[2][threadIdx.x] = mxx;
    s[3][threadIdx.x] = mxy;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < (int)blockDim.x; ++t) {
            s[0][0] = fmin(s[0][0], s[0][t]);
            s[1][0] = fmin(s[1][0], s[1][t]);
            s[2][0] = fmax(s[2][0], s[2][t]);
            s[3][0] = fmax(s[3][0], s[3][t]);
        }
        for (int c = 0; c < 4; ++c) out[c] = s[c][0];
    }
}

__global__ void k_inverse_perm(const int* __restrict__ order, long long N, int* __restrict__ nodePos) {
    const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos < N) nodePos[order[pos]] = (int)pos;
}

// ---------------------------------------------------------------------------------------------
// patch element lists
// ---------------------------------------------------------------------------------------------
// distinct patches among the nodes of element e; returns their number
__device__ __forceinline__ int element_patches(const int* __restrict__ conn, long long E, int nn, long long e,
                                               const int* __restrict__ nodePos, int PN, int (&q)[16]) {
    int n = 0;
    for (int a = 0; a < nn; ++a) {
        const int patch = nodePos[conn[(long long)a * E + e]] / PN;
        bool seen = false;
        for (int m = 0; m < n; ++m) seen = seen || (q[m] == patch);
        if (!seen) q[n++] = patch;
    }
    return n;
}

__global__ void k_pe_count(const int* __restrict__ conn, long long E, int nn, const int* __restrict__ nodePos, int PN,
                           int* __restrict__ count) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int q[16];
    const int n = element_patches(conn, E, nn, e, nodePos, PN, q);
    for (int m = 0; m < n; ++m) atomicAdd(&count[q[m]], 1);
}

__global__ void k_pe_fill(const int* __restrict__ conn, long long E, int nn, const int* __restrict__ nodePos, int PN,
                          const int* __restrict__ pePtr, int* __restrict__ cursor, int* __restrict__ peElem) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int q[16];
    const int n = element_patches(conn, E, nn, e, nodePos, PN, q);
    for (int m = 0; m < n; ++m) peElem[pePtr[q[m]] + atomicAdd(&cursor[q[m]], 1)] = (int)e;
}

// one CTA per patch: sort the patch's element list (rank sort in shared memory), record the largest patch
__global__ void k_pe_sort(const int* __restrict__ pePtr, int* __restrict__ peElem, int* __restrict__ maxElems,
                          int* __restrict__ overflow) {
    __shared__ int buf[MAX_PATCH_ELEMS];
    const int p = blockIdx.x;
    const int base = pePtr[p], n = pePtr[p + 1] - base;
    if (threadIdx.x == 0) atomicMax(maxElems, n);
    if (n > MAX_PATCH_ELEMS) {
        if (threadIdx.x == 0) atomicExch(overflow, 1);
        return;
    }
    for (int t = threadIdx.x; t < n; t += blockDim.x) buf[t] = peElem[base + t];
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int v = buf[t];
        int rank = 0;
        for (int u = 0; u < n; ++u) rank += (buf[u] < v);
        peElem[base + rank] = v;
    }
}

// patch-contiguous connectivity copy: peConn[nn*base + a*n + t] = conn[a*E + peElem[base+t]]
__global__ void k_pe_conn(const int* __restrict__ pePtr, const int* __restrict__ peElem, const int* __restrict__ conn,
                          long long E, int nn, int* __restrict__ peConn) {
    const int p = blockIdx.x;
    const int base = pePtr[p], n = pePtr[p + 1] - base;
    for (int t = threadIdx.x; t < n * nn; t += blockDim.x) {
        const int a = t / n, el = t - a * n;
        peConn[(long long)nn * base + t] = conn[(long long)a * E + peElem[base + el]];
    }
}

// ---------------------------------------------------------------------------------------------
// node blocks and sources
// ---------------------------------------------------------------------------------------------
__global__ void k_deg_perm(const int* __restrict__ order, const int* __restrict__ nrowptr, long long N,
                           int* __restrict__ degPerm) {
    const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= N) return;
    const int i = order[pos];
    degPerm[pos] = nrowptr[i + 1] - nrowptr[i];
}

__global__ void k_nb_row(long long N, int PN, const int* __restrict__ nbStart, unsigned short* __restrict__ nbRow) {
    const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= N) return;
    const unsigned short r = (unsigned short)(pos % PN);
    for (int g = nbStart[pos]; g < nbStart[pos + 1]; ++g) nbRow[g] = r;
}

__device__ __forceinline__ int lower_bound_int(const int* __restrict__ a, int n, int v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct SrcArgs {
    long long N;
    int PN;
    const int* order;
    const int* nbStart;
    const int* nrowptr;
    const int* ncols;
    const int* adjPtr;
    const int* adj;
    long long elemOffset, E;
    int nn;
    const int* conn;
    const int* pePtr;
    const int* peElem;
    int blkBits;
};

// One thread per node (in patch order) walks the node's sorted element adjacency; FILL == false counts the
// sources of each of its node-blocks, FILL == true writes them (cnt is then the per-node-block cursor).
// A node's node-blocks are touched by this thread only, so no atomics are needed and the order is fixed.
template <bool FILL>
__global__ void k_sources(const SrcArgs S, int* __restrict__ cnt, const int* __restrict__ srcPtrG,
                          unsigned short* __restrict__ src) {
    const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= S.N) return;
    const int i = S.order[pos];
    const int off = S.nrowptr[i], deg = S.nrowptr[i + 1] - off;
    const int nb0 = S.nbStart[pos];
    const int patch = (int)(pos / S.PN);
    const int peBase = S.pePtr[patch], peN = S.pePtr[patch + 1] - peBase;
    for (int q = S.adjPtr[i]; q < S.adjPtr[i + 1]; ++q) {
        const long long ge = S.adj[q];
        if (ge < S.elemOffset || ge >= S.elemOffset + S.E) continue;
        const long long e = ge - S.elemOffset;
        int a = 0;
        for (int m = 0; m < S.nn; ++m)
            if (S.conn[(long long)m * S.E + e] == i) a = m;
        int elemLocal = 0;
        if (FILL) elemLocal = lower_bound_int(S.peElem + peBase, peN, (int)e);
        for (int b = 0; b < S.nn; ++b) {
            const int j = S.conn[(long long)b * S.E + e];
            const int k = lower_bound_int(S.ncols + off, deg, j);
            if (!FILL) {
                cnt[nb0 + k] += 1;
            } else {
                int blk, tr = 0;
                if (a == b) {
                    blk = a;
                } else if (a < b) {
                    blk = S.nn + a * S.nn - a * (a + 1) / 2 + (b - a - 1);
                } else {
                    blk = S.nn + b * S.nn - b * (b + 1) / 2 + (a - b - 1);
                    tr = 1;
                }
                const int c = cnt[nb0 + k]++;
                src[srcPtrG[nb0 + k] + c] = (unsigned short)((elemLocal << (S.blkBits + 1)) | (blk << 1) | tr);
            }
        }
    }
}

// patch-relative 16-bit source offsets (one extra entry per patch) + per-patch source base
__global__ void k_src_ptr16(long long N, int PN, long long numPatches, const int* __restrict__ nbStart,
                            const int* __restrict__ srcPtrG, unsigned short* __restrict__ srcPtr16,
                            int* __restrict__ srcBase, int* __restrict__ overflow) {
    const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= N) return;
    const long long patch = pos / PN;
    const long long first = patch * PN, last = min(first + PN, N);
    const int base = srcPtrG[nbStart[first]];
    for (int g = nbStart[pos]; g < nbStart[pos + 1]; ++g) srcPtr16[g + patch] = (unsigned short)(srcPtrG[g] - base);
    if (pos == last - 1) {
        const int total = srcPtrG[nbStart[last]] - base;
        if (total > 65535) atomicExch(overflow, 1);
        srcPtr16[nbStart[last] + patch] = (unsigned short)total;
    }
    if (pos == first) srcBase[patch] = base;
    if (pos == N - 1) srcBase[numPatches] = srcPtrG[nbStart[N]];
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
static void free_program(PatchProgram& g) {
    dfree(g.d_pePtr);
    dfree(g.d_peElem);
    dfree(g.d_peConn);
    dfree(g.d_srcBase);
    dfree(g.d_srcPtr);
    dfree(g.d_src);
    g.built = false;
}

void free_patches(femb_plan* p) {
    for (auto& b : p->blocks) free_program(b.prog);
    dfree(p->d_order);
    dfree(p->d_nbStart);
    dfree(p->d_nbRow);
    p->patchesBuilt = false;
}

static int build_order(femb_plan* p, int64_t& launches) {
    const int64_t N = p->numNodes;
    cudaStream_t st = p->stream;
    CUDA_TRY(cudaMalloc(&p->d_order, N * sizeof(int)));
    if (!p->d_hintCoords) {
        k_iota<<<grid_for(N, 256), 256, 0, st>>>(p->d_order, N);
        launches++;
        return 0;
    }
    double* d_mm = nullptr;
    CUDA_TRY(cudaMalloc(&d_mm, 4 * sizeof(double)));
    k_minmax<<<1, 256, 0, st>>>(p->d_hintCoords, N, d_mm);
    launches++;
    double mm[4];
    CUDA_TRY(cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaFree(d_mm);
    const double sx = (mm[2] > mm[0]) ? 65535.0 / (mm[2] - mm[0]) : 0.0;
    const double sy = (mm[3] > mm[1]) ? 65535.0 / (mm[3] - mm[1]) : 0.0;
    unsigned *d_keys = nullptr, *d_keysOut = nullptr;
    int* d_ids = nullptr;
    CUDA_TRY(cudaMalloc(&d_keys, N * sizeof(unsigned)));
    CUDA_TRY(cudaMalloc(&d_keysOut, N * sizeof(unsigned)));
    CUDA_TRY(cudaMalloc(&d_ids, N * sizeof(int)));
    k_morton_keys<<<grid_for(N, 256), 256, 0, st>>>(p->d_hintCoords, N, mm[0], mm[1], sx, sy, d_keys, d_ids);
    launches++;
    size_t tempBytes = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tempBytes, d_keys, d_keysOut, d_ids, p->d_order, (int)N, 0, 32, st));
    void* d_temp = nullptr;
    CUDA_TRY(cudaMalloc(&d_temp, std::max<size_t>(tempBytes, 1)));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(d_temp, tempBytes, d_keys, d_keysOut, d_ids, p->d_order, (int)N, 0, 32, st));
    launches += 8;
    CUDA_TRY(cudaStreamSynchronize(st));
    cudaFree(d_temp);
    cudaFree(d_keys);
    cudaFree(d_keysOut);
    cudaFree(d_ids);
    return 0;
}

int build_patches(femb_plan* p, int64_t& launches) {
    bool any = false;
    for (auto& b : p->blocks) any = any || patch_family_supported(b.nn, b.nq);
    if (!any) return 0;
    const int64_t N = p->numNodes;
    cudaStream_t st = p->stream;
    if (p->patchNodes <= 0) p->patchNodes = DEFAULT_PATCH_NODES;
    const int PN = p->patchNodes;
    p->numPatches = (N + PN - 1) / PN;
    const int64_t NP = p->numPatches;

    TRY(build_order(p, launches));
    int* d_nodePos = nullptr;
    CUDA_TRY(cudaMalloc(&d_nodePos, N * sizeof(int)));
    k_inverse_perm<<<grid_for(N, 256), 256, 0, st>>>(p->d_order, N, d_nodePos);
    launches++;

    // node blocks in patch order
    int* d_degPerm = nullptr;
    CUDA_TRY(cudaMalloc(&d_degPerm, (N + 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&p->d_nbStart, (N + 1) * sizeof(int)));
    k_deg_perm<<<grid_for(N, 256), 256, 0, st>>>(p->d_order, p->d_nrowptr, N, d_degPerm);
    launches++;
    TRY(exclusive_scan(d_degPerm, p->d_nbStart, N, st, launches));
    CUDA_TRY(cudaMalloc(&p->d_nbRow, std::max<int64_t>(p->nnzNode, 1) * sizeof(uint16_t)));
    k_nb_row<<<grid_for(N, 256), 256, 0, st>>>(N, PN, p->d_nbStart, p->d_nbRow);
    launches++;
    cudaFree(d_degPerm);

    int *d_count = nullptr, *d_cursor = nullptr, *d_cnt = nullptr, *d_srcPtrG = nullptr;
    CUDA_TRY(cudaMalloc(&d_count, (NP + 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_cursor, (NP + 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_cnt, (p->nnzNode + 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_srcPtrG, (p->nnzNode + 2) * sizeof(int)));
    int* d_flags = nullptr;  // [0] = max elems, [1] = overflow
    CUDA_TRY(cudaMalloc(&d_flags, 2 * sizeof(int)));

    for (auto& b : p->blocks) {
        if (!patch_family_supported(b.nn, b.nq)) continue;
        PatchProgram& g = b.prog;
        const int64_t E = b.numElems;
        CUDA_TRY(cudaMemsetAsync(d_count, 0, (NP + 1) * sizeof(int), st));
        CUDA_TRY(cudaMemsetAsync(d_cursor, 0, (NP + 1) * sizeof(int), st));
        CUDA_TRY(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), st));
        k_pe_count<<<grid_for(E, 256), 256, 0, st>>>(b.d_conn, E, b.nn, d_nodePos, PN, d_count);
        launches++;
        CUDA_TRY(cudaMalloc(&g.d_pePtr, (NP + 1) * sizeof(int)));
        TRY(exclusive_scan(d_count, g.d_pePtr, NP, st, launches));
        int totalPE = 0;
        CUDA_TRY(cudaMemcpyAsync(&totalPE, g.d_pePtr + NP, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        g.totalElems = totalPE;
        CUDA_TRY(cudaMalloc(&g.d_peElem, std::max(totalPE, 1) * sizeof(int)));
        CUDA_TRY(cudaMalloc(&g.d_peConn, std::max<int64_t>((int64_t)totalPE * b.nn, 1) * sizeof(int)));
        k_pe_fill<<<grid_for(E, 256), 256, 0, st>>>(b.d_conn, E, b.nn, d_nodePos, PN, g.d_pePtr, d_cursor, g.d_peElem);
        k_pe_sort<<<(unsigned)NP, 256, 0, st>>>(g.d_pePtr, g.d_peElem, d_flags, d_flags + 1);
        launches += 2;
        int flags[2];
        CUDA_TRY(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        g.maxElems = flags[0];
        const int bits = blk_bits(b.nn);
        if (flags[1] || g.maxElems > (1 << (15 - bits))) {  // patch too populated for the 16-bit source encoding
            free_program(g);
            continue;
        }
        k_pe_conn<<<(unsigned)NP, 256, 0, st>>>(g.d_pePtr, g.d_peElem, b.d_conn, E, b.nn, g.d_peConn);
        launches++;

        SrcArgs S;
        S.N = N;
        S.PN = PN;
        S.order = p->d_order;
        S.nbStart = p->d_nbStart;
        S.nrowptr = p->d_nrowptr;
        S.ncols = p->d_ncols;
        S.adjPtr = p->d_adjPtr;
        S.adj = p->d_adj;
        S.elemOffset = b.elemOffset;
        S.E = E;
        S.nn = b.nn;
        S.conn = b.d_conn;
        S.pePtr = g.d_pePtr;
        S.peElem = g.d_peElem;
        S.blkBits = bits;
        CUDA_TRY(cudaMemsetAsync(d_cnt, 0, (p->nnzNode + 1) * sizeof(int), st));
        k_sources<false><<<grid_for(N, 128), 128, 0, st>>>(S, d_cnt, nullptr, nullptr);
        launches++;
        TRY(exclusive_scan(d_cnt, d_srcPtrG, p->nnzNode, st, launches));
        int totalSrc = 0;
        CUDA_TRY(cudaMemcpyAsync(&totalSrc, d_srcPtrG + p->nnzNode, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        g.totalSources = totalSrc;
        CUDA_TRY(cudaMalloc(&g.d_src, std::max(totalSrc, 1) * sizeof(uint16_t)));
        CUDA_TRY(cudaMalloc(&g.d_srcPtr, (p->nnzNode + NP + 1) * sizeof(uint16_t)));
        CUDA_TRY(cudaMalloc(&g.d_srcBase, (NP + 1) * sizeof(int)));
        CUDA_TRY(cudaMemsetAsync(d_cnt, 0, (p->nnzNode + 1) * sizeof(int), st));
        CUDA_TRY(cudaMemsetAsync(d_flags, 0, 2 * sizeof(int), st));
        k_sources<true><<<grid_for(N, 128), 128, 0, st>>>(S, d_cnt, d_srcPtrG, g.d_src);
        k_src_ptr16<<<grid_for(N, 256), 256, 0, st>>>(N, PN, NP, p->d_nbStart, d_srcPtrG, g.d_srcPtr, g.d_srcBase, d_flags + 1);
        launches += 2;
        CUDA_TRY(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaGetLastError());
        if (flags[1]) {
            free_program(g);
            continue;
        }
        g.built = true;
    }
    cudaFree(d_count);
    cudaFree(d_cursor);
    cudaFree(d_cnt);
    cudaFree(d_srcPtrG);
    cudaFree(d_flags);
    cudaFree(d_nodePos);
    p->patchesBuilt = true;
    return 0;
}
