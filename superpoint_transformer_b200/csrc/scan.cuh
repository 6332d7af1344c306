// scan.cuh — int32 exclusive prefix sum (three launches: tile sums, tile offsets, apply), shared
// by the stable grouping (index.cu) and the selection kernels (select.cu).  Lengths stay below
// 2^31 (checked by the callers).  The kernels are `static`: each translation unit gets its own
// copy (the library is built without relocatable device code).
#pragma once
#include "common.cuh"

namespace spt {

// ---------------------------------------------------------------- scan
constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

// inclusive block scan of one value per thread; returns inclusive prefix and
// the block total through smem.
__device__ __forceinline__ int block_inclusive_scan(int v, int* total) {
  __shared__ int warp_tot[kScanThreads / kWarp];
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) warp_tot[w] = v;
  __syncthreads();
  if (w == 0) {
    int t = (lane < (int)(blockDim.x >> 5)) ? warp_tot[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int u = __shfl_up_sync(kFull, t, o);
      if (lane >= o) t += u;
    }
    warp_tot[lane] = t;  // inclusive totals of warps
  }
  __syncthreads();
  int add = (w > 0) ? warp_tot[w - 1] : 0;
  *total = warp_tot[(blockDim.x >> 5) - 1];
  __syncthreads();
  return v + add;
}

static __global__ void __launch_bounds__(kScanThreads)
k_scan_tile_sums(const int32_t* __restrict__ in, int64_t n,
                 int32_t* __restrict__ tile_sums) {
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) s += in[base + i];
  int total;
  block_inclusive_scan(s, &total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block: exclusive scan of tile_sums (any length) in place
static __global__ void __launch_bounds__(kScanThreads)
k_scan_tile_offsets(int32_t* __restrict__ tile_sums, int64_t num_tiles) {
  int carry = 0;
  for (int64_t base = 0; base < num_tiles; base += kScanThreads) {
    int64_t i = base + threadIdx.x;
    int v = (i < num_tiles) ? tile_sums[i] : 0;
    int total;
    int inc = block_inclusive_scan(v, &total);
    if (i < num_tiles) tile_sums[i] = carry + inc - v;
    carry += total;
  }
}

static __global__ void __launch_bounds__(kScanThreads)
k_scan_apply(const int32_t* __restrict__ in, int64_t n,
             const int32_t* __restrict__ tile_offsets, int32_t* __restrict__ out) {
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  int total;
  int inc = block_inclusive_scan(s, &total);
  int run = tile_offsets[blockIdx.x] + inc - s;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

inline size_t scan_workspace_bytes(int64_t n) {
  return align_up((size_t)ceil_div(n > 0 ? n : 1, kScanTile) * 4, 256);
}

// out[i] = sum(in[0..i)) for i in [0, n); pass n = count + 1 (with in[count] readable) to get
// the total in out[count].  `tile_sums`: scan_workspace_bytes(n) of scratch; in != out.
inline void exclusive_scan_i32(const int32_t* in, int64_t n, int32_t* out, int32_t* tile_sums,
                               cudaStream_t st) {
  if (n <= 0) return;
  const int64_t tiles = ceil_div(n, kScanTile);
  k_scan_tile_sums<<<(int)tiles, kScanThreads, 0, st>>>(in, n, tile_sums);
  k_scan_tile_offsets<<<1, kScanThreads, 0, st>>>(tile_sums, tiles);
  k_scan_apply<<<(int)tiles, kScanThreads, 0, st>>>(in, n, tile_sums, out);
}

}  // namespace spt
