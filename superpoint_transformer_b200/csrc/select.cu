// select.cu — on-device node selection of a nested partition: the integer
// work of NAG.select -> Data.select -> Cluster.select -> CSRData.__getitem__ (reference
// src/data/nag.py:306-399, src/data/data.py:286-470, src/data/cluster.py:79-140,
// src/data/csr.py:328-393), which runs on every training batch right before the hot path.
//
// Design (B200): the reference relabels surviving ids with `consecutive_cluster`
// (torch.unique(sorted=True, return_inverse=True): a device sort, flagged "bottleneck" at
// cluster.py:128-130) and compacts edges with torch.where (a host round trip per call).  Ids
// here are dense ([0, num_ids)), so a relabel is a presence bitmap + an exclusive scan — O(n)
// streaming passes, no sort, bit-exact by construction (the rank of an id among the present ones
// IS its position in the sorted unique list).  Edge and CSR selection are order-preserving
// compactions on the same scan.  All kernels are HBM-bound integer streams: grid-stride,
// coalesced int64 traffic; the random accesses (rank[id], reindex[node]) hit tables of the size
// of one level, which the 126 MB L2 holds.
//
// Data-dependent output sizes are produced in two phases (count, then write) so that the caller
// — who has to allocate — reads exactly one small `counts` vector per phase-1 call.
#include "common.cuh"
#include "scan.cuh"
#include <vector>

namespace spt {

constexpr int kSelThreads = 256;

static inline int sel_grid(int64_t n) {
  int64_t b = ceil_div(n > 0 ? n : 1, kSelThreads);
  const int64_t cap = (int64_t)device_sm_count() * 16;
  return (int)(b < cap ? b : cap);
}

#define SPT_GRID_STRIDE(i, n)                                            \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x,       \
               stride_ = (int64_t)gridDim.x * blockDim.x;                \
       i < (n); i += stride_)

// ---------------------------------------------------------------- consecutive relabel
static __global__ void k_mark_ids(const int64_t* __restrict__ ids, int64_t n, int64_t num_ids,
                                  int32_t* __restrict__ flags, int64_t* __restrict__ counts) {
  SPT_GRID_STRIDE(i, n) {
    const int64_t v = ids[i];
    if (v < 0 || v >= num_ids) {
      atomicAdd((unsigned long long*)&counts[1], 1ull);
      continue;
    }
    flags[v] = 1;  // same value from every writer
  }
}

static __global__ void k_apply_rank(const int64_t* __restrict__ ids, int64_t n, int64_t num_ids,
                                    const int32_t* __restrict__ rank,
                                    int64_t* __restrict__ new_ids,
                                    const int64_t* __restrict__ payload,
                                    int64_t* __restrict__ payload_by_new) {
  SPT_GRID_STRIDE(i, n) {
    const int64_t v = ids[i];
    if (v < 0 || v >= num_ids) {
      new_ids[i] = -1;
      continue;
    }
    const int64_t r = rank[v];
    new_ids[i] = r;
    if (payload_by_new) payload_by_new[r] = payload[i];
  }
}

static __global__ void k_emit_present(const int32_t* __restrict__ flags,
                                      const int32_t* __restrict__ rank, int64_t num_ids,
                                      int64_t* __restrict__ unique_ids,
                                      int64_t* __restrict__ counts) {
  SPT_GRID_STRIDE(v, num_ids) {
    if (flags[v]) unique_ids[rank[v]] = v;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) counts[0] = rank[num_ids];
}

// ---------------------------------------------------------------- node -> new position table
static __global__ void k_fill_i64(int64_t* __restrict__ p, int64_t n, int64_t value) {
  SPT_GRID_STRIDE(i, n) p[i] = value;
}

static __global__ void k_build_reindex(const int64_t* __restrict__ idx, int64_t K,
                                       int64_t num_nodes, int64_t* __restrict__ reindex,
                                       int64_t* __restrict__ counts) {
  SPT_GRID_STRIDE(j, K) {
    const int64_t v = idx[j];
    bool bad = v < 0 || v >= num_nodes;
    if (!bad) {
      const unsigned long long old = atomicCAS((unsigned long long*)&reindex[v],
                                               (unsigned long long)(int64_t)-1,
                                               (unsigned long long)j);
      bad = old != (unsigned long long)(int64_t)-1;  // duplicate entry
    }
    if (bad) atomicAdd((unsigned long long*)&counts[1], 1ull);
  }
}

// ---------------------------------------------------------------- edge compaction
static __global__ void k_edge_flags(const int64_t* __restrict__ edge_index, int64_t E,
                                    int64_t num_nodes, const int64_t* __restrict__ reindex,
                                    int32_t* __restrict__ flags) {
  SPT_GRID_STRIDE(e, E + 1) {
    int f = 0;
    if (e < E) {
      const int64_t s = edge_index[e], t = edge_index[E + e];
      f = s >= 0 && s < num_nodes && t >= 0 && t < num_nodes && reindex[s] >= 0 &&
          reindex[t] >= 0;
    }
    flags[e] = f;
  }
}

static __global__ void k_edge_count(const int32_t* __restrict__ slot, int64_t E,
                                    int64_t* __restrict__ counts) {
  if (blockIdx.x == 0 && threadIdx.x == 0) counts[0] = slot[E];
}

static __global__ void k_edge_write(const int64_t* __restrict__ edge_index, int64_t E,
                                    const int64_t* __restrict__ reindex,
                                    const int32_t* __restrict__ slot, int64_t num_kept,
                                    int64_t* __restrict__ out_edge_index,
                                    int64_t* __restrict__ idx_edge) {
  SPT_GRID_STRIDE(e, E) {
    const int32_t p = slot[e];
    if (slot[e + 1] == p) continue;
    out_edge_index[p] = reindex[edge_index[e]];
    out_edge_index[num_kept + p] = reindex[edge_index[E + e]];
    idx_edge[p] = e;
  }
}

// ---------------------------------------------------------------- CSR group selection
static __global__ void k_selected_sizes(const int64_t* __restrict__ pointers, int64_t num_groups,
                                        const int64_t* __restrict__ idx, int64_t K,
                                        int32_t* __restrict__ sizes,
                                        int64_t* __restrict__ counts) {
  SPT_GRID_STRIDE(j, K + 1) {
    int32_t s = 0;
    if (j < K) {
      const int64_t g = idx[j];
      if (g < 0 || g >= num_groups) atomicAdd((unsigned long long*)&counts[1], 1ull);
      else s = (int32_t)(pointers[g + 1] - pointers[g]);
    }
    sizes[j] = s;
  }
}

static __global__ void k_widen_pointers(const int32_t* __restrict__ in, int64_t n,
                                        int64_t* __restrict__ out,
                                        int64_t* __restrict__ counts) {
  SPT_GRID_STRIDE(i, n) out[i] = in[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) counts[0] = in[n - 1];
}

// one thread per selected item: group by binary search in the new pointers
static __global__ void k_select_values(const int64_t* __restrict__ pointers,
                                       const int64_t* __restrict__ idx, int64_t K,
                                       const int64_t* __restrict__ new_pointers,
                                       const int64_t* __restrict__ values, int64_t M,
                                       int64_t* __restrict__ out_values,
                                       int64_t* __restrict__ out_group) {
  SPT_GRID_STRIDE(j, M) {
    int64_t lo = 0, hi = K;  // largest g with new_pointers[g] <= j
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (new_pointers[mid] <= j) lo = mid; else hi = mid;
    }
    out_values[j] = values[pointers[idx[lo]] + (j - new_pointers[lo])];
    if (out_group) out_group[j] = lo;
  }
}

// ---------------------------------------------------------------- row gather (any dtype)
template <typename Unit>
static __global__ void k_gather_rows(const Unit* __restrict__ src, int64_t row_units,
                                     const int64_t* __restrict__ idx, int64_t K,
                                     Unit* __restrict__ out) {
  const int64_t total = K * row_units;
  SPT_GRID_STRIDE(t, total) {
    const int64_t r = t / row_units, u = t - r * row_units;
    out[t] = src[idx[r] * row_units + u];
  }
}

// ---------------------------------------------------------------- subgraph sampling helpers
// flags[i] = 1 when node i lies within `r` of a seed of its own batch item (sphere, or
// cylinder around z): the neighbour search of SampleRadiusSubgraphs (reference
// src/transforms/sampling.py:1196-1231 -> knn_brute_force src/utils/neighbors.py:245-295, which
// sorts ALL distances per seed).  The arithmetic follows the reference's tensor expression in
// fp32 without contraction: z' = z * mask_z + batch * z_offset, d = sqrt(dx^2 + dy^2 + dz^2),
// kept when d <= r.  within[s] counts the nodes of seed s (the caller needs it for k_max).
constexpr int kMaxSeeds = 64;

static __global__ void __launch_bounds__(kSelThreads)
k_radius_flags(const float* __restrict__ pos, int64_t N, const int64_t* __restrict__ batch,
               const int64_t* __restrict__ seeds, int num_seeds, float r, int cylindrical,
               const float* __restrict__ z_offset, int32_t* __restrict__ flags,
               int32_t* __restrict__ within) {
  __shared__ float sx[kMaxSeeds], sy[kMaxSeeds], sz[kMaxSeeds];
  __shared__ int32_t scount[kMaxSeeds];
  const float zoff = (batch && z_offset) ? *z_offset : 0.f;
  const float mz = cylindrical ? 0.f : 1.f;
  if (threadIdx.x < num_seeds) {
    const int64_t s = seeds[threadIdx.x];
    sx[threadIdx.x] = pos[3 * s];
    sy[threadIdx.x] = pos[3 * s + 1];
    const float zb = batch ? __fmul_rn((float)batch[s], zoff) : 0.f;
    sz[threadIdx.x] = __fadd_rn(__fmul_rn(pos[3 * s + 2], mz), zb);
    scount[threadIdx.x] = 0;
  }
  __syncthreads();
  SPT_GRID_STRIDE(i, N + 1) {
    int f = 0;
    if (i < N) {
      const float x = pos[3 * i], y = pos[3 * i + 1];
      const float zb = batch ? __fmul_rn((float)batch[i], zoff) : 0.f;
      const float z = __fadd_rn(__fmul_rn(pos[3 * i + 2], mz), zb);
      for (int s = 0; s < num_seeds; ++s) {
        const float dx = __fsub_rn(x, sx[s]), dy = __fsub_rn(y, sy[s]), dz = __fsub_rn(z, sz[s]);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)),
                                   __fmul_rn(dz, dz));
        if (__fsqrt_rn(d2) <= r) {
          f = 1;
          atomicAdd(&scount[s], 1);
        }
      }
    }
    flags[i] = f;
  }
  __syncthreads();
  if (threadIdx.x < num_seeds && scount[threadIdx.x])
    atomicAdd(&within[threadIdx.x], scount[threadIdx.x]);
}

// one hop over undirected edges: out (pre-filled with a copy of in) gains the neighbours of
// every flagged node (torch_geometric.utils.k_hop_subgraph after to_undirected,
// sampling.py:1080-1091)
static __global__ void k_khop_expand(const int64_t* __restrict__ edge_index, int64_t E,
                                     int64_t N, const int32_t* __restrict__ in,
                                     int32_t* __restrict__ out) {
  SPT_GRID_STRIDE(e, E) {
    const int64_t u = edge_index[e], v = edge_index[E + e];
    if (u < 0 || u >= N || v < 0 || v >= N) continue;
    if (in[u]) out[v] = 1;
    if (in[v]) out[u] = 1;
  }
}

static __global__ void k_where_count(const int32_t* __restrict__ slot, int64_t n,
                                     int64_t* __restrict__ counts) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    counts[0] = slot[n];
    counts[1] = 0;
  }
}

static __global__ void k_where_write(const int32_t* __restrict__ slot, int64_t n,
                                     int64_t* __restrict__ out) {
  SPT_GRID_STRIDE(i, n) {
    const int32_t p = slot[i];
    if (slot[i + 1] != p) out[p] = i;
  }
}

static __global__ void k_widen_i32(const int32_t* __restrict__ in, int64_t n,
                                   int64_t* __restrict__ out) {
  SPT_GRID_STRIDE(i, n) out[i] = in[i];
}

// several tensors, one launch: the table travels in the kernel parameters (no device copy)
constexpr int kMultiMax = 16;
struct GatherMulti {
  const void* src[kMultiMax];
  void* dst[kMultiMax];
  int64_t row_units[kMultiMax];       // units per row
  int64_t unit_prefix[kMultiMax + 1]; // exclusive scan of K * row_units
  int32_t unit_log2[kMultiMax];       // log2 of the unit size in bytes: 0, 2, 3 or 4
  int32_t n;
  int64_t src_rows;                   // rows of idx outside [0, src_rows) are skipped
};

static __global__ void k_gather_rows_multi(const GatherMulti tab,
                                           const int64_t* __restrict__ idx) {
  const int64_t total = tab.unit_prefix[tab.n];
  SPT_GRID_STRIDE(t, total) {
    int s = 0;
    while (s + 1 < tab.n && tab.unit_prefix[s + 1] <= t) ++s;
    const int64_t local = t - tab.unit_prefix[s];
    const int64_t ru = tab.row_units[s];
    const int64_t r = local / ru, u = local - r * ru;
    const int64_t row = idx[r];
    if (row < 0 || row >= tab.src_rows) continue;
    const int64_t from = row * ru + u;
    switch (tab.unit_log2[s]) {
      case 4: ((uint4*)tab.dst[s])[local] = ((const uint4*)tab.src[s])[from]; break;
      case 3: ((uint2*)tab.dst[s])[local] = ((const uint2*)tab.src[s])[from]; break;
      case 2: ((uint32_t*)tab.dst[s])[local] = ((const uint32_t*)tab.src[s])[from]; break;
      default: ((uint8_t*)tab.dst[s])[local] = ((const uint8_t*)tab.src[s])[from]; break;
    }
  }
}

}  // namespace spt

using namespace spt;

extern "C" {

// ws: flags[num_ids+1] | rank[num_ids+1] | scan scratch
size_t spt_relabel_consecutive_workspace_bytes(int64_t num_ids) {
  if (num_ids < 0) return 0;
  const size_t a = align_up((size_t)(num_ids + 1) * 4, 256);
  return 2 * a + scan_workspace_bytes(num_ids + 1);
}

int spt_relabel_consecutive_i64(const int64_t* ids, int64_t n, int64_t num_ids,
                                int64_t* new_ids, int64_t* unique_ids, int64_t* counts,
                                const int64_t* payload, int64_t* payload_by_new, void* ws,
                                size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SPT_REQUIRE(n >= 0 && num_ids >= 0, SPT_E_INVALID, "relabel_consecutive: negative size");
  SPT_REQUIRE(n < 2147483647LL && num_ids < 2147483646LL, SPT_E_TOO_LARGE,
              "relabel_consecutive: n=%lld / num_ids=%lld exceed int32 internals", (long long)n,
              (long long)num_ids);
  SPT_REQUIRE(counts && ws && (n == 0 || (ids && new_ids && unique_ids)), SPT_E_INVALID,
              "relabel_consecutive: null pointer");
  SPT_REQUIRE(!payload == !payload_by_new, SPT_E_INVALID,
              "relabel_consecutive: payload and payload_by_new go together");
  const size_t need = spt_relabel_consecutive_workspace_bytes(num_ids);
  SPT_REQUIRE(ws_bytes >= need, SPT_E_WORKSPACE, "relabel_consecutive: workspace %zu < %zu",
              ws_bytes, need);
  const size_t a = align_up((size_t)(num_ids + 1) * 4, 256);
  int32_t* flags = (int32_t*)ws;
  int32_t* rank = (int32_t*)((char*)ws + a);
  int32_t* tiles = (int32_t*)((char*)ws + 2 * a);
  cudaError_t ce = cudaMemsetAsync(flags, 0, (size_t)(num_ids + 1) * 4, st);
  if (ce == cudaSuccess) ce = cudaMemsetAsync(counts, 0, 16, st);
  if (ce != cudaSuccess) {
    set_error("relabel_consecutive memset: %s", cudaGetErrorString(ce));
    return (int)ce;
  }
  if (n > 0) k_mark_ids<<<sel_grid(n), kSelThreads, 0, st>>>(ids, n, num_ids, flags, counts);
  exclusive_scan_i32(flags, num_ids + 1, rank, tiles, st);
  if (n > 0)
    k_apply_rank<<<sel_grid(n), kSelThreads, 0, st>>>(ids, n, num_ids, rank, new_ids, payload,
                                                      payload_by_new);
  k_emit_present<<<sel_grid(num_ids), kSelThreads, 0, st>>>(flags, rank, num_ids, unique_ids,
                                                            counts);
  return check_launch("relabel_consecutive");
}

// ws: flags[E+1] | scan scratch
size_t spt_select_edges_workspace_bytes(int64_t E) {
  if (E < 0) return 0;
  return align_up((size_t)(E + 1) * 4, 256) + scan_workspace_bytes(E + 1);
}

int spt_select_edges_mark(const int64_t* edge_index, int64_t E, const int64_t* idx, int64_t K,
                          int64_t num_nodes, int64_t* reindex, int32_t* slot, int64_t* counts,
                          void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SPT_REQUIRE(E >= 0 && K >= 0 && num_nodes >= 0, SPT_E_INVALID,
              "select_edges_mark: negative size");
  SPT_REQUIRE(E < 2147483646LL, SPT_E_TOO_LARGE, "select_edges_mark: E=%lld exceeds int32 slots",
              (long long)E);
  SPT_REQUIRE(counts && (num_nodes == 0 || reindex) && (K == 0 || idx) &&
                  (E == 0 || (edge_index && slot && ws)),
              SPT_E_INVALID, "select_edges_mark: null pointer");
  const size_t need = E > 0 ? spt_select_edges_workspace_bytes(E) : 0;
  SPT_REQUIRE(ws_bytes >= need, SPT_E_WORKSPACE, "select_edges_mark: workspace %zu < %zu",
              ws_bytes, need);
  cudaError_t ce = cudaMemsetAsync(counts, 0, 16, st);
  if (ce != cudaSuccess) {
    set_error("select_edges_mark memset: %s", cudaGetErrorString(ce));
    return (int)ce;
  }
  if (num_nodes > 0)
    k_fill_i64<<<sel_grid(num_nodes), kSelThreads, 0, st>>>(reindex, num_nodes, -1);
  if (K > 0)
    k_build_reindex<<<sel_grid(K), kSelThreads, 0, st>>>(idx, K, num_nodes, reindex, counts);
  if (E > 0) {
    int32_t* flags = (int32_t*)ws;
    int32_t* tiles = (int32_t*)((char*)ws + align_up((size_t)(E + 1) * 4, 256));
    k_edge_flags<<<sel_grid(E + 1), kSelThreads, 0, st>>>(edge_index, E, num_nodes, reindex,
                                                          flags);
    exclusive_scan_i32(flags, E + 1, slot, tiles, st);
    k_edge_count<<<1, 32, 0, st>>>(slot, E, counts);
  }
  return check_launch("select_edges_mark");
}

int spt_select_edges_write(const int64_t* edge_index, int64_t E, const int64_t* reindex,
                           const int32_t* slot, int64_t num_kept, int64_t* out_edge_index,
                           int64_t* idx_edge, void* stream_) {
  SPT_REQUIRE(E >= 0 && num_kept >= 0 && num_kept <= E, SPT_E_INVALID,
              "select_edges_write: bad sizes");
  if (E == 0 || num_kept == 0) return SPT_OK;
  SPT_REQUIRE(edge_index && reindex && slot && out_edge_index && idx_edge, SPT_E_INVALID,
              "select_edges_write: null pointer");
  k_edge_write<<<sel_grid(E), kSelThreads, 0, (cudaStream_t)stream_>>>(
      edge_index, E, reindex, slot, num_kept, out_edge_index, idx_edge);
  return check_launch("select_edges_write");
}

// ws: sizes[K+1] | scanned[K+1] | scan scratch
size_t spt_csr_select_workspace_bytes(int64_t K) {
  if (K < 0) return 0;
  return 2 * align_up((size_t)(K + 1) * 4, 256) + scan_workspace_bytes(K + 1);
}

int spt_csr_select_pointers(const int64_t* pointers, int64_t num_groups, int64_t num_items,
                            const int64_t* idx, int64_t K, int64_t* new_pointers,
                            int64_t* counts, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SPT_REQUIRE(num_groups >= 0 && K >= 0 && num_items >= 0, SPT_E_INVALID,
              "csr_select_pointers: negative size");
  SPT_REQUIRE(K < 2147483646LL && num_items < 2147483647LL, SPT_E_TOO_LARGE,
              "csr_select_pointers: K=%lld / num_items=%lld exceed int32 internals",
              (long long)K, (long long)num_items);
  SPT_REQUIRE(pointers && new_pointers && counts && ws && (K == 0 || idx), SPT_E_INVALID,
              "csr_select_pointers: null pointer");
  const size_t need = spt_csr_select_workspace_bytes(K);
  SPT_REQUIRE(ws_bytes >= need, SPT_E_WORKSPACE, "csr_select_pointers: workspace %zu < %zu",
              ws_bytes, need);
  const size_t a = align_up((size_t)(K + 1) * 4, 256);
  int32_t* sizes = (int32_t*)ws;
  int32_t* scanned = (int32_t*)((char*)ws + a);
  int32_t* tiles = (int32_t*)((char*)ws + 2 * a);
  cudaError_t ce = cudaMemsetAsync(counts, 0, 16, st);
  if (ce != cudaSuccess) {
    set_error("csr_select_pointers memset: %s", cudaGetErrorString(ce));
    return (int)ce;
  }
  k_selected_sizes<<<sel_grid(K + 1), kSelThreads, 0, st>>>(pointers, num_groups, idx, K, sizes,
                                                            counts);
  exclusive_scan_i32(sizes, K + 1, scanned, tiles, st);
  k_widen_pointers<<<sel_grid(K + 1), kSelThreads, 0, st>>>(scanned, K + 1, new_pointers, counts);
  return check_launch("csr_select_pointers");
}

int spt_csr_select_values_i64(const int64_t* pointers, const int64_t* idx, int64_t K,
                              const int64_t* new_pointers, const int64_t* values, int64_t M,
                              int64_t* out_values, int64_t* out_group, void* stream_) {
  SPT_REQUIRE(K >= 0 && M >= 0, SPT_E_INVALID, "csr_select_values: negative size");
  if (M == 0 || K == 0) return SPT_OK;
  SPT_REQUIRE(pointers && idx && new_pointers && values && out_values, SPT_E_INVALID,
              "csr_select_values: null pointer");
  k_select_values<<<sel_grid(M), kSelThreads, 0, (cudaStream_t)stream_>>>(
      pointers, idx, K, new_pointers, values, M, out_values, out_group);
  return check_launch("csr_select_values");
}

int spt_gather_rows_bytes(const void* src, int64_t row_bytes, const int64_t* idx, int64_t K, void* out,
                    void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SPT_REQUIRE(row_bytes >= 0 && K >= 0, SPT_E_INVALID, "gather_rows_bytes: negative size");
  if (row_bytes == 0 || K == 0) return SPT_OK;
  SPT_REQUIRE(src && idx && out, SPT_E_INVALID, "gather_rows_bytes: null pointer");
  const uintptr_t align = (uintptr_t)src | (uintptr_t)out | (uintptr_t)row_bytes;
  const int64_t units = K * row_bytes;
  if ((align & 15) == 0) {
    k_gather_rows<uint4><<<sel_grid(units / 16), kSelThreads, 0, st>>>(
        (const uint4*)src, row_bytes / 16, idx, K, (uint4*)out);
  } else if ((align & 7) == 0) {
    k_gather_rows<uint2><<<sel_grid(units / 8), kSelThreads, 0, st>>>(
        (const uint2*)src, row_bytes / 8, idx, K, (uint2*)out);
  } else if ((align & 3) == 0) {
    k_gather_rows<uint32_t><<<sel_grid(units / 4), kSelThreads, 0, st>>>(
        (const uint32_t*)src, row_bytes / 4, idx, K, (uint32_t*)out);
  } else {
    k_gather_rows<uint8_t><<<sel_grid(units), kSelThreads, 0, st>>>(
        (const uint8_t*)src, row_bytes, idx, K, (uint8_t*)out);
  }
  return check_launch("gather_rows_bytes");
}

static int gather_rows_multi(const void* const* srcs, void* const* outs,
                             const int64_t* row_bytes, int num_tensors, const int64_t* idx,
                             int64_t K, int64_t src_rows, cudaStream_t st) {
  SPT_REQUIRE(num_tensors >= 0 && K >= 0, SPT_E_INVALID, "gather_rows_multi: negative size");
  if (num_tensors == 0 || K == 0) return SPT_OK;
  SPT_REQUIRE(srcs && outs && row_bytes && idx, SPT_E_INVALID, "gather_rows_multi: null pointer");
  for (int i = 0; i < num_tensors;) {
    GatherMulti tab;
    tab.n = 0;
    tab.src_rows = src_rows;
    tab.unit_prefix[0] = 0;
    for (; i < num_tensors && tab.n < kMultiMax; ++i) {
      const int64_t rb = row_bytes[i];
      SPT_REQUIRE(rb >= 0, SPT_E_INVALID, "gather_rows_multi: negative row size");
      if (rb == 0) continue;
      SPT_REQUIRE(srcs[i] && outs[i], SPT_E_INVALID, "gather_rows_multi: null tensor");
      const uintptr_t align = (uintptr_t)srcs[i] | (uintptr_t)outs[i] | (uintptr_t)rb;
      const int lg = (align & 15) == 0 ? 4 : (align & 7) == 0 ? 3 : (align & 3) == 0 ? 2 : 0;
      const int j = tab.n++;
      tab.src[j] = srcs[i];
      tab.dst[j] = outs[i];
      tab.unit_log2[j] = lg;
      tab.row_units[j] = rb >> lg;
      tab.unit_prefix[j + 1] = tab.unit_prefix[j] + K * (rb >> lg);
    }
    if (tab.n == 0) continue;
    k_gather_rows_multi<<<sel_grid(tab.unit_prefix[tab.n]), kSelThreads, 0, st>>>(tab, idx);
  }
  return check_launch("gather_rows_multi");
}

int spt_gather_rows_multi(const void* const* srcs, void* const* outs, const int64_t* row_bytes,
                          int num_tensors, const int64_t* idx, int64_t K, void* stream_) {
  return gather_rows_multi(srcs, outs, row_bytes, num_tensors, idx, K, INT64_MAX,
                           (cudaStream_t)stream_);
}

int spt_radius_flags(const float* pos, int64_t N, const int64_t* batch, const int64_t* seeds,
                     int num_seeds, float r, int cylindrical, const float* z_offset,
                     int32_t* flags, int32_t* within, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SPT_REQUIRE(N >= 0 && num_seeds >= 0, SPT_E_INVALID, "radius_flags: negative size");
  SPT_REQUIRE(num_seeds <= kMaxSeeds, SPT_E_TOO_LARGE, "radius_flags: more than %d seeds",
              kMaxSeeds);
  SPT_REQUIRE(flags && within && (N == 0 || pos) && (num_seeds == 0 || seeds) &&
                  (!batch || z_offset),
              SPT_E_INVALID, "radius_flags: null pointer");
  cudaError_t ce = cudaMemsetAsync(within, 0, sizeof(int32_t) * (num_seeds > 0 ? num_seeds : 1), st);
  if (ce != cudaSuccess) {
    set_error("radius_flags memset: %s", cudaGetErrorString(ce));
    return (int)ce;
  }
  k_radius_flags<<<sel_grid(N + 1), kSelThreads, 0, st>>>(pos, N, batch, seeds, num_seeds, r,
                                                           cylindrical, z_offset, flags, within);
  return check_launch("radius_flags");
}

int spt_khop_expand(const int64_t* edge_index, int64_t E, int64_t N, const int32_t* flags_in,
                    int32_t* flags_out, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SPT_REQUIRE(E >= 0 && N >= 0, SPT_E_INVALID, "khop_expand: negative size");
  SPT_REQUIRE(flags_in && flags_out && (E == 0 || edge_index), SPT_E_INVALID,
              "khop_expand: null pointer");
  cudaError_t ce = cudaMemcpyAsync(flags_out, flags_in, sizeof(int32_t) * (size_t)(N + 1),
                                   cudaMemcpyDeviceToDevice, st);
  if (ce != cudaSuccess) {
    set_error("khop_expand copy: %s", cudaGetErrorString(ce));
    return (int)ce;
  }
  if (E > 0)
    k_khop_expand<<<sel_grid(E), kSelThreads, 0, st>>>(edge_index, E, N, flags_in, flags_out);
  return check_launch("khop_expand");
}

size_t spt_where_workspace_bytes(int64_t n) { return n < 0 ? 0 : scan_workspace_bytes(n + 1); }

int spt_where_count(const int32_t* flags, int64_t n, int32_t* slot, int64_t* counts, void* ws,
                    size_t ws_bytes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  SPT_REQUIRE(n >= 0 && n < 2147483646LL, SPT_E_INVALID, "where_count: bad size");
  SPT_REQUIRE(flags && slot && counts && ws, SPT_E_INVALID, "where_count: null pointer");
  SPT_REQUIRE(ws_bytes >= spt_where_workspace_bytes(n), SPT_E_WORKSPACE,
              "where_count: workspace too small");
  exclusive_scan_i32(flags, n + 1, slot, (int32_t*)ws, st);
  k_where_count<<<1, 32, 0, st>>>(slot, n, counts);
  return check_launch("where_count");
}

int spt_where_write(const int32_t* slot, int64_t n, int64_t* out, void* stream_) {
  SPT_REQUIRE(n >= 0, SPT_E_INVALID, "where_write: negative size");
  if (n == 0) return SPT_OK;
  SPT_REQUIRE(slot && out, SPT_E_INVALID, "where_write: null pointer");
  k_where_write<<<sel_grid(n), kSelThreads, 0, (cudaStream_t)stream_>>>(slot, n, out);
  return check_launch("where_write");
}

}  // extern "C"

// ---------------------------------------------------------------- one level in one call
namespace {

struct Arena {
  char* base;      // nullptr: sizing pass
  size_t off;
  int64_t take(size_t bytes) {
    off = align_up(off, 256);
    const int64_t o = (int64_t)off;
    off += bytes;
    return o;
  }
  template <typename T>
  T* at(int64_t o) const { return base ? (T*)(base + o) : nullptr; }
};

int64_t* pinned_counts() {
  static thread_local int64_t* p = nullptr;
  static thread_local int64_t fallback[8];
  if (!p && cudaHostAlloc((void**)&p, 64, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    p = fallback;
  }
  return p;
}

#define SPT_TRY(call)            \
  do {                           \
    const int rc_ = (call);      \
    if (rc_ != SPT_OK) return rc_; \
  } while (0)

// `dry`: no launches, worst-case counts -> A.off ends at the arena size the real pass can need
int select_level(const spt_select_level* a, Arena& A, int64_t* layout, cudaStream_t st,
                 bool dry) {
  const int64_t N = a->num_nodes, K = a->num_selected, E = a->edge_index ? a->num_edges : 0;
  const bool has_sub = a->sub_pointers != nullptr, has_super = a->super_index != nullptr;
  const bool upd_sub = has_sub && a->update_sub, upd_super = has_super && a->update_super;
  const int64_t cap_u = K < a->num_super ? K : a->num_super;
  const int n_rows = a->num_node_rows + a->num_edge_rows;
  if (layout)
    for (int i = 0; i < SPT_SEL_ROWS + n_rows; ++i) layout[i] = -1;

  // ---- phase 1: everything whose output size is known up front
  const int64_t o_counts = A.take(64);
  const int64_t o_reindex = A.take((size_t)N * 8);
  int64_t o_slot = -1, o_ws_e = -1;
  const size_t ws_e = E > 0 ? spt_select_edges_workspace_bytes(E) : 0;
  if (E > 0) {
    o_slot = A.take((size_t)(E + 1) * 4);
    o_ws_e = A.take(ws_e);
  }
  int64_t o_newptr = -1, o_ws_c = -1;
  const size_t ws_c = has_sub ? spt_csr_select_workspace_bytes(K) : 0;
  if (has_sub) {
    o_newptr = A.take((size_t)(K + 1) * 8);
    o_ws_c = A.take(ws_c);
  }
  int64_t o_si = -1, o_newsi = -1, o_uniq = -1, o_ws_r = -1;
  const size_t ws_r = upd_super ? spt_relabel_consecutive_workspace_bytes(a->num_super) : 0;
  if (has_super) {
    o_si = A.take((size_t)K * 8);
    if (upd_super) {
      o_newsi = A.take((size_t)K * 8);
      o_uniq = A.take((size_t)cap_u * 8);
      o_ws_r = A.take(ws_r);
    }
  }
  std::vector<void*> outs((size_t)(n_rows > 0 ? n_rows : 1), nullptr);
  for (int i = 0; i < a->num_node_rows; ++i) {
    const int64_t o = A.take((size_t)K * (size_t)a->node_row_bytes[i]);
    outs[i] = A.at<char>(o);
    if (layout) layout[SPT_SEL_ROWS + i] = o;
  }
  int64_t kept = E, items = has_sub ? a->sub_items : 0, parents = cap_u;
  if (!dry) {
    int64_t* counts = A.at<int64_t>(o_counts);
    SPT_TRY(spt_select_edges_mark(a->edge_index, E, a->idx, K, N, A.at<int64_t>(o_reindex),
                                  E > 0 ? A.at<int32_t>(o_slot) : nullptr, counts,
                                  E > 0 ? A.at<char>(o_ws_e) : nullptr, ws_e, st));
    if (has_sub)
      SPT_TRY(spt_csr_select_pointers(a->sub_pointers, N, a->sub_items, a->idx, K,
                                      A.at<int64_t>(o_newptr), counts + 2, A.at<char>(o_ws_c),
                                      ws_c, st));
    else
      cudaMemsetAsync(counts + 2, 0, 16, st);
    // gathers skip rows an invalid idx entry points at (reported right after the read below)
    if (has_super) {
      const void* src1[1] = {a->super_index};
      void* out1[1] = {A.at<int64_t>(o_si)};
      const int64_t rb1[1] = {8};
      SPT_TRY(gather_rows_multi(src1, out1, rb1, 1, a->idx, K, N, st));
    }
    if (upd_super)
      SPT_TRY(spt_relabel_consecutive_i64(A.at<int64_t>(o_si), K, a->num_super,
                                          A.at<int64_t>(o_newsi), A.at<int64_t>(o_uniq),
                                          counts + 4, nullptr, nullptr, A.at<char>(o_ws_r), ws_r,
                                          st));
    else
      cudaMemsetAsync(counts + 4, 0, 16, st);
    if (a->num_node_rows > 0)
      SPT_TRY(gather_rows_multi(a->node_src, outs.data(), a->node_row_bytes, a->num_node_rows,
                                a->idx, K, N, st));
    // the one host read of the level
    int64_t* h = pinned_counts();
    cudaError_t ce = cudaMemcpyAsync(h, counts, 48, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) {
      set_error("data_select: %s", cudaGetErrorString(ce));
      return (int)ce;
    }
    SPT_REQUIRE(h[1] == 0, SPT_E_INDEX,
                "select: %lld index entries are out of range or repeated", (long long)h[1]);
    SPT_REQUIRE(h[3] == 0, SPT_E_INDEX, "select: %lld cluster ids out of range",
                (long long)h[3]);
    SPT_REQUIRE(h[5] == 0, SPT_E_INDEX, "select: %lld super_index entries out of range",
                (long long)h[5]);
    kept = E > 0 ? h[0] : 0;
    items = has_sub ? h[2] : 0;
    parents = upd_super ? h[4] : 0;
    SPT_REQUIRE(items <= a->sub_items, SPT_E_INDEX, "select: selected clusters overlap");
  }
  if (layout) {
    layout[SPT_SEL_NUM_EDGES] = E > 0 ? kept : 0;
    layout[SPT_SEL_NUM_ITEMS] = items;
    layout[SPT_SEL_NUM_PARENTS] = upd_super ? parents : 0;
  }

  // ---- phase 2: sized by the counts
  if (E > 0) {
    const int64_t o_ei = A.take((size_t)kept * 16), o_ie = A.take((size_t)kept * 8);
    for (int i = 0; i < a->num_edge_rows; ++i) {
      const int64_t o = A.take((size_t)kept * (size_t)a->edge_row_bytes[i]);
      outs[a->num_node_rows + i] = A.at<char>(o);
      if (layout) layout[SPT_SEL_ROWS + a->num_node_rows + i] = o;
    }
    if (layout) {
      layout[SPT_SEL_EDGE_INDEX] = o_ei;
      layout[SPT_SEL_IDX_EDGE] = o_ie;
    }
    if (!dry) {
      SPT_TRY(spt_select_edges_write(a->edge_index, E, A.at<int64_t>(o_reindex),
                                     A.at<int32_t>(o_slot), kept, A.at<int64_t>(o_ei),
                                     A.at<int64_t>(o_ie), st));
      if (a->num_edge_rows > 0 && kept > 0)
        SPT_TRY(spt_gather_rows_multi(a->edge_src, outs.data() + a->num_node_rows,
                                      a->edge_row_bytes, a->num_edge_rows, A.at<int64_t>(o_ie),
                                      kept, st));
    }
  }
  if (has_sub) {
    const int64_t o_pts = A.take((size_t)items * 8);
    if (layout) layout[SPT_SEL_SUB_POINTERS] = o_newptr;
    if (!upd_sub) {
      if (layout) layout[SPT_SEL_SUB_POINTS] = o_pts;
      if (!dry)
        SPT_TRY(spt_csr_select_values_i64(a->sub_pointers, a->idx, K, A.at<int64_t>(o_newptr),
                                          a->sub_points, items, A.at<int64_t>(o_pts), nullptr,
                                          st));
    } else {
      const size_t ws2 = spt_relabel_consecutive_workspace_bytes(a->num_sub);
      const int64_t o_grp = A.take((size_t)items * 8), o_new = A.take((size_t)items * 8);
      const int64_t o_isub = A.take((size_t)items * 8), o_ssup = A.take((size_t)items * 8);
      const int64_t o_c2 = A.take(16), o_ws2 = A.take(ws2);
      if (layout) {
        layout[SPT_SEL_SUB_POINTS] = o_new;
        layout[SPT_SEL_IDX_SUB] = o_isub;
        layout[SPT_SEL_SUB_SUPER] = o_ssup;
        layout[SPT_SEL_SUB_COUNTS] = o_c2;
      }
      if (!dry) {
        SPT_TRY(spt_csr_select_values_i64(a->sub_pointers, a->idx, K, A.at<int64_t>(o_newptr),
                                          a->sub_points, items, A.at<int64_t>(o_pts),
                                          A.at<int64_t>(o_grp), st));
        SPT_TRY(spt_relabel_consecutive_i64(A.at<int64_t>(o_pts), items, a->num_sub,
                                            A.at<int64_t>(o_new), A.at<int64_t>(o_isub),
                                            A.at<int64_t>(o_c2), A.at<int64_t>(o_grp),
                                            A.at<int64_t>(o_ssup), A.at<char>(o_ws2), ws2, st));
      }
    }
  }
  if (has_super) {
    if (layout) layout[SPT_SEL_SUPER_INDEX] = upd_super ? o_newsi : o_si;
    if (upd_super) {
      const size_t ws_g = spt_group_index_workspace_bytes(K, parents);
      const int64_t o_p32 = A.take((size_t)(parents + 1) * 4), o_q32 = A.take((size_t)K * 4);
      const int64_t o_ws_g = A.take(ws_g);
      const int64_t o_sp = A.take((size_t)(parents + 1) * 8), o_sq = A.take((size_t)K * 8);
      if (layout) {
        layout[SPT_SEL_IDX_SUPER] = o_uniq;
        layout[SPT_SEL_SUPER_SUB_POINTERS] = o_sp;
        layout[SPT_SEL_SUPER_SUB_POINTS] = o_sq;
      }
      if (!dry) {
        SPT_TRY(spt_group_index(A.at<int64_t>(o_newsi), nullptr, K, parents,
                                A.at<int32_t>(o_p32), A.at<int32_t>(o_q32), nullptr,
                                A.at<char>(o_ws_g), ws_g, st));
        k_widen_i32<<<sel_grid(parents + 1), kSelThreads, 0, st>>>(
            A.at<int32_t>(o_p32), parents + 1, A.at<int64_t>(o_sp));
        if (K > 0)
          k_widen_i32<<<sel_grid(K), kSelThreads, 0, st>>>(A.at<int32_t>(o_q32), K,
                                                           A.at<int64_t>(o_sq));
      }
    }
  }
  return dry ? SPT_OK : check_launch("data_select");
}

int check_level(const spt_select_level* a) {
  SPT_REQUIRE(a, SPT_E_INVALID, "data_select: null level");
  SPT_REQUIRE(a->num_nodes >= 0 && a->num_selected >= 0 && a->num_edges >= 0 &&
                  a->num_node_rows >= 0 && a->num_edge_rows >= 0,
              SPT_E_INVALID, "data_select: negative size");
  SPT_REQUIRE(a->num_selected == 0 || a->idx, SPT_E_INVALID, "data_select: null idx");
  SPT_REQUIRE(!a->sub_pointers || (a->sub_points || a->sub_items == 0), SPT_E_INVALID,
              "data_select: sub_points missing");
  SPT_REQUIRE((a->num_node_rows == 0 || (a->node_src && a->node_row_bytes)) &&
                  (a->num_edge_rows == 0 || (a->edge_src && a->edge_row_bytes)),
              SPT_E_INVALID, "data_select: row tables missing");
  return SPT_OK;
}

}  // namespace

extern "C" {

size_t spt_data_select_arena_bytes(const spt_select_level* level) {
  if (check_level(level) != SPT_OK) return 0;
  Arena A{nullptr, 0};
  select_level(level, A, nullptr, nullptr, true);
  return align_up(A.off, 256);
}

int spt_data_select(const spt_select_level* level, void* arena, size_t arena_bytes,
                    int64_t* layout, void* stream_) {
  SPT_TRY(check_level(level));
  SPT_REQUIRE(arena && layout, SPT_E_INVALID, "data_select: null arena / layout");
  SPT_REQUIRE(((uintptr_t)arena & 255) == 0, SPT_E_INVALID, "data_select: arena not 256-aligned");
  const size_t need = spt_data_select_arena_bytes(level);
  SPT_REQUIRE(arena_bytes >= need, SPT_E_WORKSPACE, "data_select: arena %zu < %zu", arena_bytes,
              need);
  Arena A{(char*)arena, 0};
  return select_level(level, A, layout, (cudaStream_t)stream_, false);
}

}  // extern "C"
