// index.cu — integer index structures (bit-exact): stable group-by-key (COO->CSR),
// permutation helpers, int64 segment sums.
//
// Design (B200): the reference never builds a CSR — it scatters with atomics on
// unsorted COO (src/nn/attention.py:315).  Here one stable CSR is built per graph
// and reused by every block of a stage, forward and backward.  The build is a
// counting sort specialised for "many small groups": histogram (L2 atomics on
// int32 counters), a 3-kernel exclusive scan, an atomic fill, then a per-group
// sort of the (short) id lists to restore the stable (original) order, so the
// result is deterministic and equals torch.sort(stable=True).
#include "common.cuh"
#include "scan.cuh"

namespace spt {

// ---------------------------------------------------------------- error string
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------- histogram
__global__ void k_count_keys(const int64_t* __restrict__ key, int64_t n,
                             int64_t num_groups, int32_t* __restrict__ cnt,
                             int32_t* __restrict__ err) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int64_t g = key[i];
    if (g < 0 || g >= num_groups) {
      atomicAdd(err, 1);
      continue;
    }
    atomicAdd(&cnt[g], 1);
  }
}

// ---------------------------------------------------------------- fill
// slot = ptr[g] + (remaining count - 1); order inside a group is arbitrary here
// and is canonicalised by the per-group sort below.
__global__ void k_fill_slots(const int64_t* __restrict__ key, int64_t n,
                             int64_t num_groups, const int32_t* __restrict__ ptr,
                             int32_t* __restrict__ cnt, int32_t* __restrict__ perm) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    int64_t g = key[i];
    if (g < 0 || g >= num_groups) continue;
    int r = atomicSub(&cnt[g], 1) - 1;
    perm[ptr[g] + r] = (int32_t)i;
  }
}

// ---------------------------------------------------------------- per-group sort
// All-ascending bitonic network (flip + half-cleaners): virtual +inf padding
// needs no storage because every compare-exchange moves the min to the lower
// index.
constexpr int kSortWarps = 4;
constexpr int kSortMaxShort = 1024;

__global__ void __launch_bounds__(kSortWarps * kWarp)
k_sort_groups_short(const int32_t* __restrict__ ptr, int64_t num_groups,
                    int32_t* __restrict__ perm, int32_t* __restrict__ long_list,
                    int32_t* __restrict__ long_count) {
  __shared__ int32_t buf[kSortWarps][kSortMaxShort];
  int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t g = (int64_t)blockIdx.x * kSortWarps + w;
  if (g >= num_groups) return;
  int b = ptr[g], e = ptr[g + 1];
  int deg = e - b;
  if (deg <= 1) return;
  if (deg > kSortMaxShort) {
    if (lane == 0) {
      int slot = atomicAdd(long_count, 1);
      long_list[slot] = (int32_t)g;
    }
    return;
  }
  int32_t* s = buf[w];
  for (int i = lane; i < deg; i += 32) s[i] = perm[b + i];
  __syncwarp();
  int n2 = 1;
  while (n2 < deg) n2 <<= 1;
  for (int k = 2; k <= n2; k <<= 1) {
    // flip
    for (int i = lane; i < deg; i += 32) {
      int l = i ^ (k - 1);
      if (l > i && l < deg) {
        int a = s[i], c = s[l];
        if (a > c) { s[i] = c; s[l] = a; }
      }
    }
    __syncwarp();
    for (int j = k >> 2; j > 0; j >>= 1) {
      for (int i = lane; i < deg; i += 32) {
        int l = i ^ j;
        if (l > i && l < deg) {
          int a = s[i], c = s[l];
          if (a > c) { s[i] = c; s[l] = a; }
        }
      }
      __syncwarp();
    }
  }
  for (int i = lane; i < deg; i += 32) perm[b + i] = s[i];
}

// groups longer than kSortMaxShort: one CTA per listed group, network run in
// global memory (L2-resident).
__global__ void __launch_bounds__(1024)
k_sort_groups_long(const int32_t* __restrict__ ptr, int32_t* __restrict__ perm,
                   const int32_t* __restrict__ long_list,
                   const int32_t* __restrict__ long_count) {
  int nlong = *long_count;
  for (int li = blockIdx.x; li < nlong; li += gridDim.x) {
    int g = long_list[li];
    int b = ptr[g];
    int deg = ptr[g + 1] - b;
    int32_t* s = perm + b;
    int64_t n2 = 1;
    while (n2 < deg) n2 <<= 1;
    for (int64_t k = 2; k <= n2; k <<= 1) {
      for (int i = threadIdx.x; i < deg; i += blockDim.x) {
        int64_t l = (int64_t)i ^ (k - 1);
        if (l > i && l < deg) {
          int a = s[i], c = s[l];
          if (a > c) { s[i] = c; s[l] = a; }
        }
      }
      __syncthreads();
      for (int64_t j = k >> 2; j > 0; j >>= 1) {
        for (int i = threadIdx.x; i < deg; i += blockDim.x) {
          int64_t l = (int64_t)i ^ j;
          if (l > i && l < deg) {
            int a = s[i], c = s[l];
            if (a > c) { s[i] = c; s[l] = a; }
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();
  }
}

__global__ void k_gather_other(const int64_t* __restrict__ other,
                               const int32_t* __restrict__ perm, int64_t n,
                               int32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = (int32_t)other[perm[i]];
}

__global__ void k_invert_perm(const int32_t* __restrict__ perm, int64_t n,
                              int32_t* __restrict__ inv) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) inv[perm[i]] = (int32_t)i;
}

__global__ void k_gather_i32(const int32_t* __restrict__ src,
                             const int32_t* __restrict__ idx, int64_t n,
                             int32_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = src[idx[i]];
}

// out[j] = g for ptr[g] <= j < ptr[g + 1]: one warp per group (groups are short)
__global__ void k_expand_pointers(const int32_t* __restrict__ ptr, int64_t num_groups,
                                  int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= num_groups) return;
  const int b = ptr[g], e = ptr[g + 1];
  for (int j = b + lane; j < e; j += 32) out[j] = (int32_t)g;
}

// one warp per group; int64 exact sums
__global__ void k_segment_sum_i64(const int64_t* __restrict__ values,
                                  const int32_t* __restrict__ ptr,
                                  const int32_t* __restrict__ points,
                                  int64_t num_groups, int64_t* __restrict__ out) {
  int lane = threadIdx.x & 31;
  int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (g >= num_groups) return;
  int b = ptr[g], e = ptr[g + 1];
  long long s = 0;
  for (int i = b + lane; i < e; i += 32) {
    if (values) {
      int c = points ? points[i] : i;
      s += values[c];
    } else {
      s += 1;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
  if (lane == 0) out[g] = s;
}

static inline int grid_for(int64_t n, int threads, int max_blocks = device_sm_count() * 16) {
  int64_t b = ceil_div(n, threads);
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

}  // namespace spt

using namespace spt;

extern "C" {

int spt_abi_version(void) { return SPT_ABI_VERSION; }
const char* spt_last_error(void) { return g_err; }
const char* spt_build_info(void) {
  return "libspt_b200 abi=1 arch=sm_100a cuda="
#define SPT_STR2(x) #x
#define SPT_STR(x) SPT_STR2(x)
      SPT_STR(__CUDACC_VER_MAJOR__) "." SPT_STR(__CUDACC_VER_MINOR__);
}

// ws layout: [0,256): err counter (int32) + long_count (int32 at +4)
//            cnt[num_groups+1] | tile_sums[num_tiles] | long_list[num_groups]
size_t spt_group_index_workspace_bytes(int64_t n, int64_t num_groups) {
  (void)n;
  if (num_groups < 0) return 0;
  int64_t ns = num_groups + 1;
  int64_t tiles = ceil_div(ns, kScanTile);
  size_t b = 256;
  b += align_up((size_t)ns * 4, 256);
  b += align_up((size_t)tiles * 4, 256);
  b += align_up((size_t)(num_groups > 0 ? num_groups : 1) * 4, 256);
  return b;
}

int spt_group_index(const int64_t* key, const int64_t* other, int64_t n,
                    int64_t num_groups, int32_t* ptr, int32_t* perm,
                    int32_t* other_sorted, void* ws, size_t ws_bytes,
                    void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SPT_REQUIRE(n >= 0 && num_groups >= 0, SPT_E_INVALID, "group_index: negative size");
  SPT_REQUIRE(n < 2147483647LL && num_groups < 2147483646LL, SPT_E_TOO_LARGE,
              "group_index: n=%lld / num_groups=%lld exceed int32 internals",
              (long long)n, (long long)num_groups);
  SPT_REQUIRE(ptr && ws && (n == 0 || (key && perm)), SPT_E_INVALID,
              "group_index: null pointer");
  SPT_REQUIRE(!(other && !other_sorted), SPT_E_INVALID,
              "group_index: other given without other_sorted");
  size_t need = spt_group_index_workspace_bytes(n, num_groups);
  SPT_REQUIRE(ws_bytes >= need, SPT_E_WORKSPACE,
              "group_index: workspace %zu < %zu", ws_bytes, need);

  int64_t ns = num_groups + 1;
  int64_t tiles = ceil_div(ns, kScanTile);
  char* base = (char*)ws;
  int32_t* err = (int32_t*)base;
  int32_t* long_count = err + 1;
  int32_t* cnt = (int32_t*)(base + 256);
  int32_t* tile_sums = (int32_t*)(base + 256 + align_up((size_t)ns * 4, 256));
  int32_t* long_list =
      (int32_t*)((char*)tile_sums + align_up((size_t)tiles * 4, 256));

  cudaError_t ce = cudaMemsetAsync(ws, 0, 256 + (size_t)ns * 4, st);
  if (ce != cudaSuccess) {
    set_error("group_index memset: %s", cudaGetErrorString(ce));
    return (int)ce;
  }
  if (n > 0) {
    k_count_keys<<<grid_for(n, 256), 256, 0, st>>>(key, n, num_groups, cnt, err);
  }
  exclusive_scan_i32(cnt, ns, ptr, tile_sums, st);
  if (n > 0) {
    k_fill_slots<<<grid_for(n, 256), 256, 0, st>>>(key, n, num_groups, ptr, cnt, perm);
    if (num_groups > 0) {
      int64_t blocks = ceil_div(num_groups, kSortWarps);
      k_sort_groups_short<<<(unsigned)blocks, kSortWarps * kWarp, 0, st>>>(
          ptr, num_groups, perm, long_list, long_count);
      k_sort_groups_long<<<64, 1024, 0, st>>>(ptr, perm, long_list, long_count);
    }
    if (other) {
      k_gather_other<<<grid_for(n, 256), 256, 0, st>>>(other, perm, n, other_sorted);
    }
  }
  return check_launch("group_index");
}

int spt_invert_permutation(const int32_t* perm, int64_t n, int32_t* inv,
                           void* stream_) {
  SPT_REQUIRE(n >= 0, SPT_E_INVALID, "invert_permutation: negative size");
  if (n == 0) return SPT_OK;
  SPT_REQUIRE(perm && inv, SPT_E_INVALID, "invert_permutation: null pointer");
  k_invert_perm<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream_>>>(perm, n, inv);
  return check_launch("invert_permutation");
}

int spt_gather_i32(const int32_t* src, const int32_t* idx, int64_t n,
                   int32_t* out, void* stream_) {
  SPT_REQUIRE(n >= 0, SPT_E_INVALID, "gather_i32: negative size");
  if (n == 0) return SPT_OK;
  SPT_REQUIRE(src && idx && out, SPT_E_INVALID, "gather_i32: null pointer");
  k_gather_i32<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream_>>>(src, idx, n, out);
  return check_launch("gather_i32");
}

int spt_expand_pointers_i32(const int32_t* ptr, int64_t num_groups, int32_t* out,
                            void* stream_) {
  SPT_REQUIRE(num_groups >= 0, SPT_E_INVALID, "expand_pointers: negative size");
  if (num_groups == 0) return SPT_OK;
  SPT_REQUIRE(ptr && out, SPT_E_INVALID, "expand_pointers: null pointer");
  k_expand_pointers<<<(unsigned)ceil_div(num_groups * 32, 256), 256, 0,
                      (cudaStream_t)stream_>>>(ptr, num_groups, out);
  return check_launch("expand_pointers");
}

int spt_segment_sum_i64(const int64_t* values, const int32_t* ptr,
                        const int32_t* points, int64_t num_groups, int64_t* out,
                        void* stream_) {
  SPT_REQUIRE(num_groups >= 0, SPT_E_INVALID, "segment_sum_i64: negative size");
  if (num_groups == 0) return SPT_OK;
  SPT_REQUIRE(ptr && out, SPT_E_INVALID, "segment_sum_i64: null pointer");
  int64_t threads = num_groups * 32;
  k_segment_sum_i64<<<(unsigned)ceil_div(threads, 256), 256, 0,
                      (cudaStream_t)stream_>>>(values, ptr, points, num_groups, out);
  return check_launch("segment_sum_i64");
}

}  // extern "C"
