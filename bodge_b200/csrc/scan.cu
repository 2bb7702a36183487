// Exclusive prefix sum over int32 built from warp shuffles: reduce tiles -> scan tile sums ->
// re-scan tiles with their offsets.  Used for indptr construction (skeleton, generic lattices)
// and for zero-block compaction.  Three short launches; n up to 2^31-1.
#include "bdg_internal.h"

namespace {

constexpr int kThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kThreads * kItems;

__device__ __forceinline__ int warp_inclusive_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int up = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += up;
    }
    return v;
}

// Inclusive scan of one value per thread across the CTA; returns the inclusive prefix and the
// CTA total through `total`.
__device__ __forceinline__ int block_inclusive_scan(int v, int &total) {
    __shared__ int warp_sums[kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int s = lane < kThreads / 32 ? warp_sums[lane] : 0;
        s = warp_inclusive_scan(s);
        if (lane < kThreads / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    if (warp > 0) incl += warp_sums[warp - 1];
    total = warp_sums[kThreads / 32 - 1];
    __syncthreads();
    return incl;
}

__global__ void __launch_bounds__(kThreads) scan_tile_sums(const int32_t *__restrict__ in, int64_t n,
                                                           int32_t *__restrict__ tile_sums) {
    const int64_t base = (int64_t)blockIdx.x * kTile;
    int acc = 0;
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
        int64_t idx = base + it * kThreads + threadIdx.x;  // coalesced
        if (idx < n) acc += in[idx];
    }
    int total;
    block_inclusive_scan(acc, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// Single CTA: in-place exclusive scan of the tile sums, carrying the running offset across chunks.
__global__ void __launch_bounds__(kThreads) scan_of_tile_sums(int32_t *tile_sums, int64_t n_tiles,
                                                              int32_t *total_out) {
    int carry = 0;
    for (int64_t base = 0; base < n_tiles; base += kThreads) {
        int64_t idx = base + threadIdx.x;
        int v = idx < n_tiles ? tile_sums[idx] : 0;
        int total;
        int incl = block_inclusive_scan(v, total);
        if (idx < n_tiles) tile_sums[idx] = carry + incl - v;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(kThreads) scan_apply(const int32_t *__restrict__ in, int64_t n,
                                                       const int32_t *__restrict__ tile_offsets,
                                                       int32_t *__restrict__ out) {
    // Blocked arrangement: thread t owns items [t*kItems, (t+1)*kItems) of the tile so a single
    // value per thread enters the CTA scan.
    const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
    int vals[kItems];
    int sum = 0;
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
        int64_t idx = base + it;
        vals[it] = idx < n ? in[idx] : 0;
        sum += vals[it];
    }
    int total;
    int incl = block_inclusive_scan(sum, total);
    int run = tile_offsets[blockIdx.x] + incl - sum;
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
        int64_t idx = base + it;
        if (idx < n) out[idx] = run;
        run += vals[it];
    }
}

}  // namespace

// out[i] = sum(in[0..i)), i < n.  `in` and `out` may alias.  total_dev (device int32, optional)
// receives sum(in[0..n)).
int exclusive_scan_i32(bdg_system *sys, const int32_t *in, int32_t *out, int64_t n, int32_t *total_dev) {
    if (n <= 0) {
        if (total_dev) BDG_CUDA(cudaMemsetAsync(total_dev, 0, sizeof(int32_t), sys->stream));
        return BDG_OK;
    }
    const int64_t n_tiles = ceil_div(n, kTile);
    BDG_TRY(ensure_scratch(sys, 3, (size_t)n_tiles * sizeof(int32_t)));
    int32_t *tile_sums = sys->scratch_i32[3].as<int32_t>();
    scan_tile_sums<<<(unsigned)n_tiles, kThreads, 0, sys->stream>>>(in, n, tile_sums);
    scan_of_tile_sums<<<1, kThreads, 0, sys->stream>>>(tile_sums, n_tiles, total_dev);
    scan_apply<<<(unsigned)n_tiles, kThreads, 0, sys->stream>>>(in, n, tile_sums, out);
    BDG_CUDA(cudaGetLastError());
    return BDG_OK;
}
