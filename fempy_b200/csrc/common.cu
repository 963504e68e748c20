// Error state and the exclusive scan shared by the plan builders.
#include "femb_internal.h"

static thread_local std::string g_err;

int femb_fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

extern "C" const char* femb_last_error(void) { return g_err.c_str(); }
extern "C" const char* femb_version(void) { return "femb200 0.2 (sm_100a)"; }
extern "C" int femb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

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
int exclusive_scan(const int* d_in, int* d_out, int64_t n, cudaStream_t st, int64_t& launches) {
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

// Page-locked host buffers for the Python layer's output pool (fempy_b200/_pinned.py): a D2H copy into pinned memory runs
// at PCIe speed, into fresh pageable memory at a quarter of it (page faults + the driver's staged copy).
extern "C" int femb_host_alloc(size_t bytes, void** out) {
    if (!out) return fail("femb_host_alloc: out is NULL");
    *out = nullptr;
    if (femb_device_count() == 0) return fail("femb_host_alloc: no CUDA device available");
    CUDA_TRY(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return 0;
}

extern "C" int femb_host_free(void* ptr) {
    if (ptr) CUDA_TRY(cudaFreeHost(ptr));
    return 0;
}
