// Patch programs for the write-once assembly (built once per plan, on the GPU).
//
// The node set is cut into patches of `patchNodes` consecutive nodes along a locality-preserving order
// (Morton order of the hint coordinates when the caller supplied them, node-id order otherwise).  For
// every element block the builder then records, per patch,
//   * the elements touching an owned node (sorted; evaluated by that patch's CTA, halo elements included),
//   * a patch-local copy of their connectivity (local node ids), and
//   * for every element a scatter record: where the CSR rows of its owned nodes sit in the patch's
//     shared-memory accumulators, and a colour per local node so that same-colour updates never collide.
// All of it is packed into one contiguous "program blob" per patch (patch_blob.h).  The assembly kernel
// (assemble_pipe.cuh) needs no atomics and no zero-fill of K: every CSR value and residual entry is written
// exactly once, in a fixed summation order.
#include <cub/device/device_radix_sort.cuh>

#include "femb_internal.h"
#include "patch_blob.h"

static constexpr int DEFAULT_PATCH_NODES_PER_WARP = 20;  // a team of w warps takes patches of <= 32*w elements (assemble_pipe.cuh)
static constexpr int MAX_PATCH_ELEMS = 1024;
static constexpr int MAX_PATCH_CONN = 4096;  // ne * nn entries sorted in shared memory when the halo is built

bool patch_family_supported(int nn, int nq) { return (nn == 3 && nq == 1) || (nn == 4 && nq == 4); }

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

__device__ __forceinline__ int lower_bound_int(const int* __restrict__ a, int n, int v) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------
// program blobs
// ---------------------------------------------------------------------------------------------
struct BlobArgs {
    long long N;
    int PN;
    int nn;
    long long E, elemOffset;
    const int* order;
    const int* nodePos;
    const int* nbStart;
    const int* nrowptr;
    const int* ncols;
    const int* adjPtr;
    const int* adj;
    const int* conn;
    const int* pePtr;
    const int* peElem;
    int* halo;      // scratch: nn * pePtr[p] + h
    int* nHalo;     // per patch
    int* units;     // per patch blob size in 16-byte units
    int* stats;     // [0] max blob bytes, [1] max local nodes, [2] overflow, [3] max node-blocks per patch, [4] max colours
};

// one CTA per patch: sorted unique list of the halo nodes (nodes of patch elements that the patch does not own)
__global__ void k_blob_halo(const BlobArgs B) {
    __shared__ int cand[MAX_PATCH_CONN];
    __shared__ int sorted[MAX_PATCH_CONN];
    const int p = blockIdx.x;
    const int pe0 = B.pePtr[p], ne = B.pePtr[p + 1] - pe0;
    const int M = ne * B.nn;
    const long long pos0 = (long long)p * B.PN;
    const long long pos1 = (pos0 + B.PN < B.N) ? pos0 + B.PN : B.N;
    if (M > MAX_PATCH_CONN) {
        if (threadIdx.x == 0) atomicExch(&B.stats[2], 1);
        return;
    }
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const int a = m / ne, t = m - a * ne;
        const int g = B.conn[(long long)a * B.E + B.peElem[pe0 + t]];
        const int pos = B.nodePos[g];
        cand[m] = (pos >= pos0 && pos < pos1) ? 0x7fffffff : g;
    }
    __syncthreads();
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const int v = cand[m];
        int rank = 0;
        for (int u = 0; u < M; ++u) rank += (cand[u] < v) || (cand[u] == v && u < m);
        sorted[rank] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int* out = B.halo + (long long)B.nn * pe0;
        int h = 0;
        for (int m = 0; m < M; ++m) {
            const int v = sorted[m];
            if (v == 0x7fffffff) break;
            if (h == 0 || out[h - 1] != v) out[h++] = v;
        }
        B.nHalo[p] = h;
        const int nOwned = (int)(pos1 - pos0);
        const int nnb = B.nbStart[pos1] - B.nbStart[pos0];
        const BlobLayout L = blob_layout(B.nn, ne, nOwned, nOwned + h, nnb);
        B.units[p] = L.bytes / 16;
        atomicMax(&B.stats[0], L.bytes);
        atomicMax(&B.stats[1], nOwned + h);
        atomicMax(&B.stats[3], nnb);
        if (2 * nnb + nOwned > 65535) atomicExch(&B.stats[2], 1);
    }
}

__global__ void k_blob_fill(const BlobArgs B, const int* __restrict__ blobPtr, unsigned char* __restrict__ blob) {
    __shared__ int numColours[4];
    const int p = blockIdx.x;
    const int pe0 = B.pePtr[p], ne = B.pePtr[p + 1] - pe0;
    const long long pos0 = (long long)p * B.PN;
    const long long pos1 = (pos0 + B.PN < B.N) ? pos0 + B.PN : B.N;
    const int nOwned = (int)(pos1 - pos0);
    const int h = B.nHalo[p];
    const int* halo = B.halo + (long long)B.nn * pe0;
    const int nb0 = B.nbStart[pos0], nnb = B.nbStart[pos1] - nb0;
    const BlobLayout L = blob_layout(B.nn, ne, nOwned, nOwned + h, nnb);
    unsigned char* base = blob + 16ll * blobPtr[p];
    int* hdr = reinterpret_cast<int*>(base);
    int* lnode = reinterpret_cast<int*>(base + L.offLNODE);
    int2* node = reinterpret_cast<int2*>(base + L.offNODE);
    unsigned short* nbRow = reinterpret_cast<unsigned short*>(base + L.offNBROW);
    ElemInfo* einfo = reinterpret_cast<ElemInfo*>(base + L.offEINFO);
    unsigned short* connOut = reinterpret_cast<unsigned short*>(base + L.offCONN);
    int* elemOut = reinterpret_cast<int*>(base + L.offELEM);
    if (threadIdx.x < 4) numColours[threadIdx.x] = 0;
    __syncthreads();
    for (int r = threadIdx.x; r < nOwned; r += blockDim.x) {
        const int i = B.order[pos0 + r];
        lnode[r] = i;
        const int off = B.nrowptr[i], deg = B.nrowptr[i + 1] - off;
        const int start = B.nbStart[pos0 + r] - nb0;
        const int diagK = deg ? lower_bound_int(B.ncols + off, deg, i) : 0;
        if (deg > 255 || start > 32767) atomicExch(&B.stats[2], 1);
        node[r] = make_int2(off, (start & 0xffff) | ((deg & 0xff) << 16) | ((diagK & 0xff) << 24));
        for (int k = 0; k < deg; ++k) nbRow[start + k] = (unsigned short)r;
    }
    for (int m = threadIdx.x; m < h; m += blockDim.x) lnode[nOwned + m] = halo[m];
    for (int m = threadIdx.x; m < ne * B.nn; m += blockDim.x) {
        const int a = m / ne, t = m - a * ne;
        const int g = B.conn[(long long)a * B.E + B.peElem[pe0 + t]];
        const int pos = B.nodePos[g];
        const int local = (pos >= pos0 && pos < pos1) ? (int)(pos - pos0) : nOwned + lower_bound_int(halo, h, g);
        connOut[m] = (unsigned short)local;
    }
    // scatter program: for every (element, local node a) whose node the patch owns, where the node's rows start in
    // the accumulators, where each of the element's nodes sits in that row, and the colour = rank of the element
    // among the patch elements that have the same node as THEIR a-th node (same-colour updates never collide)
    for (int t = threadIdx.x; t < ne; t += blockDim.x) {
        const int e = B.peElem[pe0 + t];
        elemOut[t] = (int)(B.elemOffset + e);
        ElemInfo info;
        int nodes[4];
        for (int a = 0; a < 4; ++a) nodes[a] = (a < B.nn) ? B.conn[(long long)a * B.E + e] : -1;
        for (int a = 0; a < 4; ++a) {
            info.rowBase[a] = 0;
            info.deg[a] = 0;
            info.colour[a] = 0xff;
            for (int b = 0; b < 4; ++b) info.k[a * 4 + b] = 0;
            if (a >= B.nn) continue;
            const int i = nodes[a];
            const int pos = B.nodePos[i];
            if (pos < pos0 || pos >= pos1) continue;
            const int off = B.nrowptr[i], deg = B.nrowptr[i + 1] - off;
            info.rowBase[a] = (unsigned short)(2 * (B.nbStart[pos] - nb0) + (pos - pos0));  // + one pad double2 per preceding owned node
            info.deg[a] = (unsigned char)deg;
            for (int b = 0; b < B.nn; ++b) info.k[a * 4 + b] = (unsigned char)lower_bound_int(B.ncols + off, deg, nodes[b]);
            int colour = 0;
            for (int q = B.adjPtr[i]; q < B.adjPtr[i + 1]; ++q) {
                const long long ge = B.adj[q];
                if (ge < B.elemOffset || ge >= B.elemOffset + B.E) continue;
                const long long e2 = ge - B.elemOffset;
                if (e2 < e && B.conn[(long long)a * B.E + e2] == i) ++colour;
            }
            if (colour >= 255) atomicExch(&B.stats[2], 1);
            info.colour[a] = (unsigned char)colour;
            atomicMax(&numColours[a], colour + 1);
        }
        einfo[t] = info;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        const int v[8] = {ne, nOwned, nOwned + h, nnb, numColours[0], numColours[1], numColours[2], numColours[3]};
        hdr[threadIdx.x] = v[threadIdx.x];
    }
    if (threadIdx.x == 0) atomicMax(&B.stats[4], max(max(numColours[0], numColours[1]), max(numColours[2], numColours[3])));
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
static void free_program(PatchProgram& g) {
    dfree(g.d_pePtr);
    dfree(g.d_peElem);
    dfree(g.d_blobPtr);
    dfree(g.d_blob);
    g.built = false;
}

void free_patches(femb_plan* p) {
    for (auto& b : p->blocks) free_program(b.prog);
    dfree(p->d_order);
    dfree(p->d_nbStart);
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

static int build_patches_once(femb_plan* p, int64_t& launches) {
    bool any = false;
    for (auto& b : p->blocks) any = any || patch_family_supported(b.nn, b.nq);
    if (!any) return 0;
    const int64_t N = p->numNodes;
    cudaStream_t st = p->stream;
    if (p->patchNodes <= 0) p->patchNodes = DEFAULT_PATCH_NODES_PER_WARP * p->teamWarps;
    p->patchNodes = std::min(1024, ((p->patchNodes + 3) / 4) * 4);  // multiple of 4: a patch's slice of the permuted BC mask is fetched 4 bytes at a time
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
    cudaFree(d_degPerm);

    int *d_count = nullptr, *d_cursor = nullptr, *d_nHalo = nullptr, *d_units = nullptr;
    CUDA_TRY(cudaMalloc(&d_count, (NP + 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_cursor, (NP + 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_nHalo, (NP + 1) * sizeof(int)));
    CUDA_TRY(cudaMalloc(&d_units, (NP + 1) * sizeof(int)));
    int* d_flags = nullptr;  // [0] = max elems, [1] = overflow, [2..4] = blob stats (max bytes, max local nodes, overflow)
    CUDA_TRY(cudaMalloc(&d_flags, 8 * sizeof(int)));

    for (auto& b : p->blocks) {
        if (!patch_family_supported(b.nn, b.nq)) continue;
        PatchProgram& g = b.prog;
        const int64_t E = b.numElems;
        CUDA_TRY(cudaMemsetAsync(d_count, 0, (NP + 1) * sizeof(int), st));
        CUDA_TRY(cudaMemsetAsync(d_cursor, 0, (NP + 1) * sizeof(int), st));
        CUDA_TRY(cudaMemsetAsync(d_flags, 0, 8 * sizeof(int), st));
        k_pe_count<<<grid_for(E, 256), 256, 0, st>>>(b.d_conn, E, b.nn, d_nodePos, PN, d_count);
        launches++;
        CUDA_TRY(cudaMalloc(&g.d_pePtr, (NP + 1) * sizeof(int)));
        TRY(exclusive_scan(d_count, g.d_pePtr, NP, st, launches));
        int totalPE = 0;
        CUDA_TRY(cudaMemcpyAsync(&totalPE, g.d_pePtr + NP, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        g.totalElems = totalPE;
        CUDA_TRY(cudaMalloc(&g.d_peElem, std::max(totalPE, 1) * sizeof(int)));
        k_pe_fill<<<grid_for(E, 256), 256, 0, st>>>(b.d_conn, E, b.nn, d_nodePos, PN, g.d_pePtr, d_cursor, g.d_peElem);
        k_pe_sort<<<(unsigned)NP, 256, 0, st>>>(g.d_pePtr, g.d_peElem, d_flags, d_flags + 1);
        launches += 2;
        int flags[8];
        CUDA_TRY(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        g.maxElems = flags[0];
        if (flags[1] || b.nn > 4) {  // patch too populated, or a family without a patch kernel
            free_program(g);
            continue;
        }

        // pack everything a patch needs into one contiguous blob
        int* d_halo = nullptr;
        CUDA_TRY(cudaMalloc(&d_halo, std::max<int64_t>((int64_t)totalPE * b.nn, 1) * sizeof(int)));
        BlobArgs B;
        B.N = N;
        B.PN = PN;
        B.nn = b.nn;
        B.E = E;
        B.elemOffset = b.elemOffset;
        B.order = p->d_order;
        B.nodePos = d_nodePos;
        B.nbStart = p->d_nbStart;
        B.nrowptr = p->d_nrowptr;
        B.ncols = p->d_ncols;
        B.conn = b.d_conn;
        B.pePtr = g.d_pePtr;
        B.peElem = g.d_peElem;
        B.adjPtr = p->d_adjPtr;
        B.adj = p->d_adj;
        B.halo = d_halo;
        B.nHalo = d_nHalo;
        B.units = d_units;
        B.stats = d_flags + 2;
        k_blob_halo<<<(unsigned)NP, 256, 0, st>>>(B);
        launches++;
        CUDA_TRY(cudaMalloc(&g.d_blobPtr, (NP + 1) * sizeof(int)));
        TRY(exclusive_scan(d_units, g.d_blobPtr, NP, st, launches));
        int totalUnits = 0;
        CUDA_TRY(cudaMemcpyAsync(&totalUnits, g.d_blobPtr + NP, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (flags[4]) {
            cudaFree(d_halo);
            free_program(g);
            continue;
        }
        g.blobBytes = 16ll * totalUnits;
        g.maxBlobBytes = flags[2];
        g.maxLocal = flags[3];
        g.maxNodeBlocks = flags[5];
        CUDA_TRY(cudaMalloc(&g.d_blob, std::max<int64_t>(g.blobBytes, 16)));
        k_blob_fill<<<(unsigned)NP, 256, 0, st>>>(B, g.d_blobPtr, g.d_blob);
        launches++;
        CUDA_TRY(cudaMemcpyAsync(flags, d_flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaGetLastError());
        cudaFree(d_halo);
        dfree(g.d_peElem);
        if (flags[4]) {
            free_program(g);
            continue;
        }
        g.maxColours = flags[6];
        g.built = true;
    }
    cudaFree(d_count);
    cudaFree(d_cursor);
    cudaFree(d_nHalo);
    cudaFree(d_units);
    cudaFree(d_flags);
    cudaFree(d_nodePos);
    p->patchesBuilt = true;
    return 0;
}

// Patches must not hold more elements than a compute warpgroup has threads: shrink the patch size until the
// most populated patch fits (unless the caller fixed the size).
int build_patches(femb_plan* p, int64_t& launches) {
    const bool userFixed = p->patchNodes > 0;
    for (int attempt = 0; attempt < 6; ++attempt) {
        TRY(build_patches_once(p, launches));
        int worst = 0;
        for (auto& b : p->blocks) worst = std::max(worst, b.prog.maxElems);
        const int target = 32 * p->teamWarps;  // one thread per element
        if (userFixed || worst <= target || p->patchNodes <= 4) return 0;
        const int next = std::max(4, (int)((int64_t)p->patchNodes * target / worst) / 4 * 4);
        const int shrunk = (next >= p->patchNodes) ? p->patchNodes - 4 : next;
        free_patches(p);
        p->patchNodes = shrunk;
    }
    return 0;
}
