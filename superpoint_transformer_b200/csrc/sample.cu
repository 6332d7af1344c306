// sample.cu — per-segment sampling without replacement: the device form of
// `sparse_sample` (reference src/utils/sparse.py:142-243, "huge BOTTLENECK" at :214-216), the
// core of NAG.get_sampling / SampleSubNodes (src/data/nag.py:662-711,
// src/transforms/sampling.py:656-715).
//
// The reference shuffles ALL elements (randperm), sorts them by segment and keeps the first
// n_samples[g] of every segment: two global passes through a sort for what is a local decision.
// Here the elements are already grouped (the stable CSR of spt_group_index) and every segment
// draws its own uniformly random k-subset from a counter-based generator (Philox-4x32-10, keyed
// by the caller's seed and the segment id — reproducible, independent of the launch geometry):
//   * segments of <= kShortSegment candidates: one thread each, selection sampling (Knuth's
//     Algorithm S: element j of the remaining r is taken with probability needed / r) — exact
//     k-subsets, one pass, no storage;
//   * longer segments: one warp each, the k smallest of per-element 32-bit random keys found by
//     a 32-step radix select (keys are recomputed, never stored), ties at the threshold taken
//     in position order.
// Both give every k-subset of the segment the same probability (up to the 2^-32 granularity of
// the integer comparison).  Inside a segment the output keeps the position order of the
// candidates (the reference's order inside a segment is random; the SET is what is sampled).
#include "common.cuh"

namespace spt {

constexpr int kShortSegment = 256;
constexpr int kSampleThreads = 128;

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__device__ __forceinline__ uint32_t pick(const uint4& v, int i) {
  return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w;
}

__device__ __forceinline__ int64_t element_of(const int32_t* __restrict__ seg_perm,
                                              const int64_t* __restrict__ elem_ids,
                                              int64_t slot) {
  const int64_t c = seg_perm[slot];
  return elem_ids ? elem_ids[c] : c;
}

// one thread per segment: Algorithm S over the segment's candidates
static __global__ void __launch_bounds__(kSampleThreads)
k_sample_short(const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_perm,
               int64_t num_segments, const int64_t* __restrict__ n_samples,
               const int64_t* __restrict__ out_ptr, const int64_t* __restrict__ elem_ids,
               uint2 key, int64_t* __restrict__ out, int32_t* __restrict__ long_list,
               int32_t* __restrict__ long_count) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= num_segments) return;
  const int64_t base = seg_ptr[g];
  const int64_t size = seg_ptr[g + 1] - base;
  int64_t needed = n_samples[g];
  if (needed > size) needed = size;
  if (needed <= 0) return;
  if (size > kShortSegment) {
    long_list[atomicAdd(long_count, 1)] = (int32_t)g;
    return;
  }
  int64_t o = out_ptr[g];
  uint4 r4 = make_uint4(0, 0, 0, 0);
  for (int64_t j = 0; j < size && needed > 0; ++j) {
    if ((j & 3) == 0)
      r4 = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)(j >> 2), 0u),
                         key);
    const uint32_t remaining = (uint32_t)(size - j);
    // floor(r * remaining / 2^32) is uniform on [0, remaining)
    if ((int64_t)__umulhi(pick(r4, (int)(j & 3)), remaining) < needed) {
      out[o++] = element_of(seg_perm, elem_ids, base + j);
      --needed;
    }
  }
}

__device__ __forceinline__ uint32_t element_key(int64_t g, int64_t j, uint2 key) {
  return philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)j,
                                  0x80000000u | (uint32_t)(j >> 32)), key).x;
}

// one warp per long segment: radix select of the k-th smallest key, then an ordered compaction
static __global__ void __launch_bounds__(kSampleThreads)
k_sample_long(const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_perm,
              const int64_t* __restrict__ n_samples, const int64_t* __restrict__ out_ptr,
              const int64_t* __restrict__ elem_ids, uint2 key, int64_t* __restrict__ out,
              const int32_t* __restrict__ long_list, const int32_t* __restrict__ long_count) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int num_long = *long_count;
  for (int w = blockIdx.x * warps_per_block + (threadIdx.x >> 5); w < num_long;
       w += gridDim.x * warps_per_block) {
    const int64_t g = long_list[w];
    const int64_t base = seg_ptr[g];
    const int64_t size = seg_ptr[g + 1] - base;
    int64_t k = n_samples[g];
    if (k > size) k = size;
    uint32_t threshold = 0xffffffffu;
    int64_t take_equal = 0;          // how many of the elements with key == threshold to take
    const bool all = (k == size);
    if (!all) {
      uint32_t prefix = 0;
      int64_t want = k;              // rank (1-based) of the threshold among the candidates left
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t hi_mask = bit == 31 ? 0u : ~((2u << bit) - 1u);
        int64_t cnt0 = 0;
        for (int64_t j = lane; j < size; j += 32) {
          const uint32_t kk = element_key(g, j, key);
          cnt0 += ((kk & hi_mask) == prefix) && !((kk >> bit) & 1u);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt0 += __shfl_xor_sync(kFull, cnt0, o);
        if (want > cnt0) {
          prefix |= 1u << bit;
          want -= cnt0;
        }
      }
      threshold = prefix;
      take_equal = want;
    }
    int64_t o = out_ptr[g];
    int64_t seen_equal = 0;
    for (int64_t j0 = 0; j0 < size; j0 += 32) {
      const int64_t j = j0 + lane;
      bool sel = false, eq = false;
      if (j < size) {
        if (all) {
          sel = true;
        } else {
          const uint32_t kk = element_key(g, j, key);
          eq = kk == threshold;
          sel = kk < threshold;
        }
      }
      const unsigned eq_ballot = __ballot_sync(kFull, eq);
      const unsigned lt = (1u << lane) - 1u;
      if (eq && seen_equal + __popc(eq_ballot & lt) < take_equal) sel = true;
      seen_equal += __popc(eq_ballot);
      const unsigned sel_ballot = __ballot_sync(kFull, sel);
      if (sel) out[o + __popc(sel_ballot & lt)] = element_of(seg_perm, elem_ids, base + j);
      o += __popc(sel_ballot);
    }
  }
}

}  // namespace spt

using namespace spt;

extern "C" {

// ws: [0, 256) counter of long segments | long_list[num_segments]
size_t spt_sparse_sample_workspace_bytes(int64_t num_segments) {
  if (num_segments < 0) return 0;
  return 256 + align_up((size_t)(num_segments > 0 ? num_segments : 1) * 4, 256);
}

int spt_sparse_sample(const int32_t* seg_ptr, const int32_t* seg_perm, int64_t num_segments,
                      const int64_t* n_samples, const int64_t* out_ptr, const int64_t* elem_ids,
                      uint64_t seed, int64_t* out, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SPT_REQUIRE(num_segments >= 0, SPT_E_INVALID, "sparse_sample: negative size");
  SPT_REQUIRE(num_segments < 2147483647LL, SPT_E_TOO_LARGE,
              "sparse_sample: num_segments=%lld exceeds int32 internals", (long long)num_segments);
  if (num_segments == 0) return SPT_OK;
  SPT_REQUIRE(seg_ptr && seg_perm && n_samples && out_ptr && out && ws, SPT_E_INVALID,
              "sparse_sample: null pointer");
  const size_t need = spt_sparse_sample_workspace_bytes(num_segments);
  SPT_REQUIRE(ws_bytes >= need, SPT_E_WORKSPACE, "sparse_sample: workspace %zu < %zu", ws_bytes,
              need);
  int32_t* long_count = (int32_t*)ws;
  int32_t* long_list = (int32_t*)((char*)ws + 256);
  cudaError_t ce = cudaMemsetAsync(ws, 0, 256, st);
  if (ce != cudaSuccess) {
    set_error("sparse_sample memset: %s", cudaGetErrorString(ce));
    return (int)ce;
  }
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  const int64_t blocks = ceil_div(num_segments, kSampleThreads);
  k_sample_short<<<(unsigned)blocks, kSampleThreads, 0, st>>>(
      seg_ptr, seg_perm, num_segments, n_samples, out_ptr, elem_ids, key, out, long_list,
      long_count);
  k_sample_long<<<device_sm_count() * 4, kSampleThreads, 0, st>>>(
      seg_ptr, seg_perm, n_samples, out_ptr, elem_ids, key, out, long_list, long_count);
  return check_launch("sparse_sample");
}

}  // extern "C"
