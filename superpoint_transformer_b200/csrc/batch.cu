// batch.cu — on-device batch construction: the integer work of
// NAGBatch.from_nag_list -> Batch.from_data_list -> CSRBatch.from_list (reference
// src/data/nag.py:878-898, src/data/data.py:1154-1242, src/data/csr.py:676-757): concatenate the
// per-item index tensors while adding the per-item offsets (edge_index by the node count,
// super_index by the parent count, Cluster.points by the child count, Cluster.pointers by the
// running point count), and emit the `batch` vector.  One launch per key for any number of
// items (a segment table lives in device memory); bit-exact int64 arithmetic.
#include "common.cuh"

namespace spt {

constexpr int kMaxSegThreads = 256;

// out[prefix[s] + j] = src_s[j] + offset[s]   (src_s == nullptr: out = s, the `batch` vector;
// skip_first: drop element 0 of every segment but the first — CSR pointer concatenation)
__global__ void __launch_bounds__(kMaxSegThreads)
k_concat_offset_i64(const int64_t* const* __restrict__ srcs, const int64_t* __restrict__ prefix,
                    const int64_t* __restrict__ offsets, int num_seg, int64_t total,
                    int skip_first, int64_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    // binary search: largest s with prefix[s] <= i   (prefix[0] = 0, prefix[num_seg] = total)
    int lo = 0, hi = num_seg;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (prefix[mid] <= i) lo = mid; else hi = mid;
    }
    const int64_t j = i - prefix[lo] + ((skip_first && lo > 0) ? 1 : 0);
    const int64_t* src = srcs ? srcs[lo] : nullptr;
    out[i] = (src ? src[j] : (int64_t)lo) + (offsets ? offsets[lo] : 0);
  }
}

}  // namespace spt

using namespace spt;

extern "C" {

int spt_concat_offset_i64(const int64_t* const* srcs, const int64_t* prefix,
                          const int64_t* offsets, int num_segments, int64_t total,
                          int skip_first, int64_t* out, void* stream_) {
  SPT_REQUIRE(num_segments >= 0 && total >= 0, SPT_E_INVALID, "concat_offset: negative size");
  if (total == 0 || num_segments == 0) return SPT_OK;
  SPT_REQUIRE(prefix && out, SPT_E_INVALID, "concat_offset: null pointer");
  int64_t blocks = ceil_div(total, kMaxSegThreads);
  if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
  k_concat_offset_i64<<<(unsigned)blocks, kMaxSegThreads, 0, (cudaStream_t)stream_>>>(
      srcs, prefix, offsets, num_segments, total, skip_first, out);
  return check_launch("concat_offset_i64");
}

}  // extern "C"
