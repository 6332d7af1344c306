/*
 * spt_b200.h — C ABI of libspt_b200.so: the B200 (sm_100a) hot path of the
 * Superpoint Transformer: superpoint-graph self-attention + segment-wise
 * scatter/pool aggregation.
 *
 * The reference (drprojects/superpoint_transformer @ eb959b6) has NO native code
 * and therefore no FFI of its own: every entry point below
 * replaces a *composition of third-party leaf calls* made from the reference's
 * Python glue.  Each declaration cites the reference call-site it replaces
 * (paths relative to /root/reference).  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add to bind them.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer
 *     unless stated otherwise; tensors are contiguous row-major.
 *   - the caller owns every buffer; the library never allocates device memory.
 *     Where scratch is needed the caller passes `ws` of at least
 *     `*_workspace_bytes(...)` bytes (256-byte aligned).
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *     no host synchronisation, no internal streams.  One exception, stated at its
 *     declaration: spt_data_select reads a 48-byte counter vector back in the middle
 *     of the call (output sizes that depend on the data).
 *   - return value: 0 = ok; negative = invalid argument / unsupported shape
 *     (SPT_E_*); positive = cudaError_t of the failed launch.  Details via
 *     spt_last_error() (thread-local, host string).
 *   - API index tensors are int64 (the reference casts everything to int64 in
 *     `NAGCast`, configs/datamodule/semantic/default.yaml:208-210); the internal
 *     CSR arrays are int32 (E, N < 2^31 is checked).
 */
#ifndef SPT_B200_H
#define SPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPT_ABI_VERSION 1

/* status codes (negative = caller error) */
#define SPT_OK 0
#define SPT_E_INVALID -1      /* null pointer / negative size / bad enum      */
#define SPT_E_UNSUPPORTED -2  /* shape outside what the kernels implement     */
#define SPT_E_WORKSPACE -3    /* workspace too small                          */
#define SPT_E_TOO_LARGE -4    /* element count does not fit int32 internals   */
#define SPT_E_INDEX -5        /* an index tensor holds out-of-range or repeated entries */

/* reduce ops for segment pooling (src/nn/pool.py:24-82) */
#define SPT_REDUCE_SUM 0
#define SPT_REDUCE_MEAN 1
#define SPT_REDUCE_MAX 2
#define SPT_REDUCE_MIN 3

/* qk-scale modes (src/utils/nn.py:75-127): D=(dim/num_heads)^-1/2, G=deg(s)^-1/2 */
#define SPT_SCALE_D_TIMES_G 0 /* default (qk_scale=None) and 'd.g'            */
#define SPT_SCALE_D_PLUS_G 1  /* 'd+g'                                        */
#define SPT_SCALE_D 2         /* 'd'                                          */
#define SPT_SCALE_G 3         /* 'g'                                          */
#define SPT_SCALE_CONST 4     /* numeric qk_scale used as is                  */

int spt_abi_version(void);
const char* spt_last_error(void);
/* compile-time facts, for the loader to verify it got the sm_100a build */
const char* spt_build_info(void);

/* ------------------------------------------------------------------------- *
 *  Integer index structure (bit-exact)                                      *
 * ------------------------------------------------------------------------- */

/* Stable grouping of n items by key (key[i] in [0, num_groups)):
 *   ptr[num_groups+1] = exclusive scan of bincount(key)
 *   perm[n]           = stable argsort(key)   (original item id at sorted slot)
 *   other_sorted[n]   = other[perm]  (optional; e.g. edge targets -> CSR col)
 * Replaces the atomics-based scatter the reference relies on
 * (torch_scatter.scatter_sum, src/nn/attention.py:315) by a deterministic CSR,
 * and is the device version of `indices_to_pointers` (src/utils/sparse.py:23-41)
 * with a *stable* order. With key=edge_index[0], other=edge_index[1] it is the
 * COO->CSR build of SURVEY.md §2.2 (1); with key=super_index it reproduces
 * `Cluster.pointers/points` (src/data/cluster.py:19-77).
 * Out-of-range keys are skipped and counted in ws[0] (int32 error counter). */
size_t spt_group_index_workspace_bytes(int64_t n, int64_t num_groups);
int spt_group_index(const int64_t* key, const int64_t* other /*nullable*/,
                    int64_t n, int64_t num_groups, int32_t* ptr, int32_t* perm,
                    int32_t* other_sorted /*nullable*/, void* ws,
                    size_t ws_bytes, void* stream);

/* inv[perm[j]] = j */
int spt_invert_permutation(const int32_t* perm, int64_t n, int32_t* inv,
                           void* stream);
/* out[j] = src[idx[j]]  (int32 payload, int32 index) */
int spt_gather_i32(const int32_t* src, const int32_t* idx, int64_t n,
                   int32_t* out, void* stream);

/* out[j] = g for ptr[g] <= j < ptr[g+1]  (the row of every CSR slot; `out` has ptr[num_groups]
 * entries).  Input of the edge-parallel attention kernels (spt_attn_extras.edge_row).       */
int spt_expand_pointers_i32(const int32_t* ptr, int64_t num_groups, int32_t* out,
                            void* stream);

/* Segment sum of int64 values (values==NULL -> ones), segments given by
 * ptr/points (points==NULL -> identity).  `NAG.get_sub_size`
 * (src/data/nag.py:59-110), used by `NodeSize` (src/transforms/graph.py:1475-1498). */
int spt_segment_sum_i64(const int64_t* values /*nullable*/, const int32_t* ptr,
                        const int32_t* points /*nullable*/, int64_t num_groups,
                        int64_t* out, void* stream);

/* ------------------------------------------------------------------------- *
 *  Row gathers                                                              *
 * ------------------------------------------------------------------------- */

/* out[i, :] = x[idx[i], :]   (fp32 rows of width C).
 * `IndexUnpool.forward` (src/nn/unpool.py:12-13) and the q[s]/k[t]/v[t]
 * expansions of src/nn/attention.py:207-211 when used stand-alone. */
int spt_gather_rows_i64(const float* x, const int64_t* idx, int64_t n_out,
                        int64_t C, float* out, void* stream);
int spt_gather_rows_i32(const float* x, const int32_t* idx, int64_t n_out,
                        int64_t C, float* out, void* stream);

/* x = hi + lo exactly, hi = x truncated to tf32 (13 low mantissa bits cleared).
 * Operand split of the "3xTF32" dense projections: the reference runs its Linear
 * layers as cuBLAS GEMMs (src/nn/attention.py:191,318; src/nn/mlp.py:45); here they
 * are three tensor-core TF32 GEMMs on (hi,lo) pairs with fp32 accumulation, which
 * keeps fp32-level accuracy (~2^-21 relative). */
int spt_split_tf32(const float* x, int64_t n, float* hi, float* lo, void* stream);

/* Dense projections on the 5th-generation tensor cores, fp32-accurate (3xTF32:
 * hi/lo operand split, three tcgen05.mma.kind::tf32 per k-step, fp32 accumulation in
 * TMEM), every operand streamed from HBM once by TMA (csrc/gemm_umma.cu).  Shapes the
 * TMA descriptors cannot take (fewer than 512 / 2048 rows, leading dimensions or C not
 * 16-byte aligned, N > 4096, dW accumulators beyond 512 TMEM columns) run the
 * mma.sync kernels of csrc/gemm.cu with the same numerics.
 *   gemm_nt     : C[M,N]  = A[M,K] . B[N,K]^T + bias[N]   (nn.Linear forward; dX with
 *                 B = W^T).  K, lda, ldb multiples of 4; A, B 16-byte aligned.
 *   gemm_tn_acc : C[N,K] += A[M,N]^T . B[M,K] ; colsumA[N] += column sums of A
 *                 (dW and dbias of nn.Linear; fp32 atomics, caller zero-fills).
 * Replaces the cuBLAS calls behind src/nn/attention.py:191,226,239,295,318 and
 * src/nn/mlp.py:45 (which the reference runs in TF32, src/train.py:93-94). */
int spt_gemm_nt(const float* A, int64_t M, int64_t K, int64_t lda, const float* B,
                int64_t N, int64_t ldb, const float* bias /*nullable*/, float* C,
                int64_t ldc, void* stream);
int spt_gemm_tn_acc(const float* A, int64_t M, int64_t N, int64_t lda, const float* B,
                    int64_t K, int64_t ldb, float* C, int64_t ldc,
                    float* colsumA /*nullable*/, void* stream);

/* ------------------------------------------------------------------------- *
 *  On-device batch construction (src/data/nag.py:878-898, data.py:1154-1242,  *
 *  csr.py:676-757): offset-concatenation of int64 index tensors                *
 * ------------------------------------------------------------------------- */

/* out[prefix[s] + j] = srcs[s][j (+1 if skip_first and s > 0)] + offsets[s] for the
 * `num_segments` device arrays srcs[s] (device table of device pointers); prefix [S+1] =
 * exclusive scan of the emitted lengths (device), offsets [S] (device, nullable = 0).
 * srcs == NULL writes the segment id (the `batch` vector of Batch.from_data_list).
 * skip_first = 1 concatenates CSR pointer arrays (element 0 of later items dropped). */
int spt_concat_offset_i64(const int64_t* const* srcs, const int64_t* prefix,
                          const int64_t* offsets, int num_segments, int64_t total,
                          int skip_first, int64_t* out, void* stream);

/* ------------------------------------------------------------------------- *
 *  On-device node selection (NAG.select src/data/nag.py:306-399, Data.select    *
 *  data.py:286-470, Cluster.select cluster.py:79-140, CSRData.__getitem__ /     *
 *  index_select_pointers csr.py:328-393): csrc/select.cu.  Sizes that depend on *
 *  the data come back in a 2-element device vector `counts` = {count, number of *
 *  invalid input entries}; the caller reads it, allocates, and runs phase 2.     *
 * ------------------------------------------------------------------------- */

/* `consecutive_cluster` (torch_geometric.nn.pool.consecutive, called at cluster.py:131 and
 * data.py:405) for ids in [0, num_ids): new_ids[i] = rank of ids[i] among the distinct values
 * present (ascending), unique_ids[r] = the r-th distinct value (capacity min(n, num_ids)),
 * counts = {number of distinct values, number of out-of-range ids (their new id is -1)}.
 * A presence bitmap + exclusive scan instead of the reference's sort.  Optional
 * payload_by_new[new_ids[i]] = payload[i] (only meaningful when the ids are distinct; this is
 * `Cluster.to_super_index` of the selected clusters, cluster.py:67-77). */
size_t spt_relabel_consecutive_workspace_bytes(int64_t num_ids);
int spt_relabel_consecutive_i64(const int64_t* ids, int64_t n, int64_t num_ids,
                                int64_t* new_ids, int64_t* unique_ids, int64_t* counts,
                                const int64_t* payload /*nullable*/,
                                int64_t* payload_by_new /*nullable*/, void* ws, size_t ws_bytes,
                                void* stream);

/* Data.select's edge update (data.py:356-371), phase 1: reindex[v] = j for v = idx[j], -1
 * elsewhere (num_nodes entries); slot [E+1] = exclusive scan of "both end points survive";
 * counts = {edges kept, entries of idx that are out of range or repeated}.  E = 0 only builds
 * (and validates) the table.  edge_index is [2, E] contiguous. */
size_t spt_select_edges_workspace_bytes(int64_t E);
int spt_select_edges_mark(const int64_t* edge_index, int64_t E, const int64_t* idx, int64_t K,
                          int64_t num_nodes, int64_t* reindex, int32_t* slot, int64_t* counts,
                          void* ws, size_t ws_bytes, void* stream);
/* phase 2: out_edge_index [2, num_kept] = reindex[edge_index[:, kept]] in the original edge
 * order, idx_edge [num_kept] = the kept edge positions (to slice edge attributes). */
int spt_select_edges_write(const int64_t* edge_index, int64_t E, const int64_t* reindex,
                           const int32_t* slot, int64_t num_kept, int64_t* out_edge_index,
                           int64_t* idx_edge, void* stream);

/* CSRData.index_select_pointers (csr.py:328-356), phase 1: new_pointers [K+1] = exclusive scan
 * of the sizes of groups idx[0..K); counts = {selected items, out-of-range group ids}. */
size_t spt_csr_select_workspace_bytes(int64_t K);
int spt_csr_select_pointers(const int64_t* pointers, int64_t num_groups, int64_t num_items,
                            const int64_t* idx, int64_t K, int64_t* new_pointers,
                            int64_t* counts, void* ws, size_t ws_bytes, void* stream);
/* phase 2: out_values[j] = values[val_idx[j]] (the `v[val_idx]` of csr.py:384) for the M
 * selected items; out_group[j] (nullable) = position in idx of the group item j belongs to. */
int spt_csr_select_values_i64(const int64_t* pointers, const int64_t* idx, int64_t K,
                              const int64_t* new_pointers, const int64_t* values, int64_t M,
                              int64_t* out_values, int64_t* out_group /*nullable*/,
                              void* stream);

/* out[j, :] = src[idx[j], :] for rows of row_bytes bytes of any dtype (`item[idx]`,
 * data.py:447-459); idx must be in range (validated by spt_select_edges_mark). */
int spt_gather_rows_bytes(const void* src, int64_t row_bytes, const int64_t* idx, int64_t K,
                          void* out, void* stream);
/* The same for `num_tensors` tensors in one launch per 16 tensors (all node-level or all
 * edge-level attributes of a Data object).  srcs / outs / row_bytes are HOST arrays of device
 * pointers / byte counts: the table is passed in the kernel parameters. */
int spt_gather_rows_multi(const void* const* srcs, void* const* outs, const int64_t* row_bytes,
                          int num_tensors, const int64_t* idx, int64_t K, void* stream);

/* One level of Data.select (src/data/data.py:286-470) in ONE call: every primitive above, with a
 * single host read of the data-dependent sizes in the middle, outputs carved from a
 * caller-provided device arena (256-byte aligned pieces).  The reference issues ~25 tensor ops
 * with three host round trips per level; called piecewise from Python the primitives above cost
 * ~40 allocations and ~4 reads per level, which is what this entry removes.
 *   sub_pointers == NULL: the level has no `sub` (or the caller replaces it: NAG.select on the
 *   levels above the selected one); super_index == NULL likewise.  update_sub / update_super as
 *   in the reference.  node_src / edge_src (+ *_row_bytes): HOST arrays describing the node-level
 *   and edge-level attribute tensors to gather (`item[idx]`, `item[idx_edge]`).
 * layout (HOST, SPT_SEL_ROWS + num_node_rows + num_edge_rows int64): element counts and byte
 * offsets into the arena of the outputs (-1 = absent); int64 outputs unless noted:
 *   [NUM_EDGES] E'   [EDGE_INDEX] [2, E']   [IDX_EDGE] [E']
 *   [NUM_ITEMS] M    [SUB_POINTERS] [K+1]   [SUB_POINTS] [M]   [IDX_SUB] [M]   [SUB_SUPER] [M]
 *                    [SUB_COUNTS] 2 x int64 {distinct point ids, out-of-range ids} (device; for
 *                    an optional check: a valid Cluster gives {M, 0})
 *   [NUM_PARENTS] U  [SUPER_INDEX] [K]      [IDX_SUPER] [U]    [SUPER_SUB_POINTERS] [U+1]
 *                    [SUPER_SUB_POINTS] [K]
 *   [ROWS + i] the i-th node-level output (K rows), then the edge-level ones (E' rows).
 * Returns SPT_E_INDEX when idx (or super_index) holds invalid or repeated entries. */
typedef struct spt_select_level {
  int64_t num_nodes;
  const int64_t* idx;          int64_t num_selected;
  const int64_t* edge_index;   int64_t num_edges;
  const int64_t* sub_pointers; const int64_t* sub_points;
  int64_t sub_items;           int64_t num_sub;
  int update_sub;
  const int64_t* super_index;  int64_t num_super;
  int update_super;
  int num_node_rows;  const void* const* node_src;  const int64_t* node_row_bytes;
  int num_edge_rows;  const void* const* edge_src;  const int64_t* edge_row_bytes;
} spt_select_level;
enum { SPT_SEL_NUM_EDGES = 0, SPT_SEL_NUM_ITEMS, SPT_SEL_NUM_PARENTS, SPT_SEL_EDGE_INDEX,
       SPT_SEL_IDX_EDGE, SPT_SEL_SUB_POINTERS, SPT_SEL_SUB_POINTS, SPT_SEL_IDX_SUB,
       SPT_SEL_SUB_SUPER, SPT_SEL_SUB_COUNTS, SPT_SEL_SUPER_INDEX, SPT_SEL_IDX_SUPER,
       SPT_SEL_SUPER_SUB_POINTERS, SPT_SEL_SUPER_SUB_POINTS, SPT_SEL_ROWS };
size_t spt_data_select_arena_bytes(const spt_select_level* level);
int spt_data_select(const spt_select_level* level, void* arena, size_t arena_bytes,
                    int64_t* layout, void* stream);

/* Subgraph sampling (src/transforms/sampling.py:1003-1231).  Node sets are int32 flag vectors
 * of N+1 entries (flags[N] = 0, the scan's sentinel).
 * spt_radius_flags: flags[i] = 1 when node i is within r of one of the (<= 64) seeds that shares
 * its `batch` id — a sphere, or with cylindrical != 0 a cylinder around z — and within[s] = the
 * number of such nodes of seed s (for the caller's k_max check): the neighbour search of
 * SampleRadiusSubgraphs (:1196-1231 -> knn_brute_force, src/utils/neighbors.py:245-295) in one
 * pass instead of a full sort of the distances.  pos is [N, 3] fp32; z_offset (device scalar,
 * required with batch) is the reference's per-batch-item z shift (neighbors.py:273-279).
 * spt_khop_expand: flags_out = flags_in plus the neighbours of the flagged nodes over the edges
 * taken in both directions — one hop of k_hop_subgraph(to_undirected(edge_index)) (:1080-1091).
 * spt_where_count / spt_where_write: ascending positions of the non-zero flags (two phases:
 * slot [n+1] = exclusive scan, counts[0] = how many). */
int spt_radius_flags(const float* pos, int64_t N, const int64_t* batch /*nullable*/,
                     const int64_t* seeds, int num_seeds, float r, int cylindrical,
                     const float* z_offset /*nullable*/, int32_t* flags, int32_t* within,
                     void* stream);
int spt_khop_expand(const int64_t* edge_index, int64_t E, int64_t N, const int32_t* flags_in,
                    int32_t* flags_out, void* stream);
size_t spt_where_workspace_bytes(int64_t n);
int spt_where_count(const int32_t* flags, int64_t n, int32_t* slot, int64_t* counts, void* ws,
                    size_t ws_bytes, void* stream);
int spt_where_write(const int32_t* slot, int64_t n, int64_t* out, void* stream);

/* ------------------------------------------------------------------------- *
 *  Per-segment sampling without replacement: `sparse_sample`                     *
 *  (src/utils/sparse.py:142-243), the core of NAG.get_sampling / SampleSubNodes   *
 *  (src/data/nag.py:662-711, src/transforms/sampling.py:656-715): csrc/sample.cu  *
 * ------------------------------------------------------------------------- */

/* For every segment g of the stable CSR (seg_ptr [G+1], seg_perm [n]: spt_group_index of the
 * candidates' segment ids) draw a uniformly random subset of min(n_samples[g], size_g)
 * candidates and write their ids — elem_ids[c] for candidate c, or c itself when elem_ids is
 * NULL — to out[out_ptr[g] ...) in candidate order (out_ptr [G+1] = exclusive scan of the
 * clamped n_samples; the caller computes n_samples with the reference's fp32 heuristic,
 * sparse.py:176-189).  Randomness: Philox-4x32-10 keyed by `seed` and the segment id, so the
 * result depends on (seed, CSR) only.  The reference shuffles everything and sorts by segment;
 * the SET sampled per segment has the same distribution, the order inside a segment differs. */
size_t spt_sparse_sample_workspace_bytes(int64_t num_segments);
int spt_sparse_sample(const int32_t* seg_ptr, const int32_t* seg_perm, int64_t num_segments,
                      const int64_t* n_samples, const int64_t* out_ptr,
                      const int64_t* elem_ids /*nullable*/, uint64_t seed, int64_t* out,
                      void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 *  Value RPE of SelfAttentionBlock (src/nn/attention.py:294-301), applied        *
 *  algebraically: y = agg + Wbd . abar + bv (x) sump  (csrc/vrpe.cu)              *
 * ------------------------------------------------------------------------- */

/* Wbd [C, H*F] = blockdiag(Wv[h*Dv:(h+1)*Dv, :]) (transpose = 0) or its transpose
 * [H*F, C] (transpose = 1); heads_share: Wv is the single-head encoder [Dv, F]. */
int spt_vrpe_blockdiag(const float* Wv, int H, int Dv, int F, int heads_share, int transpose,
                       float* out, void* stream);

/* y[n, c] = agg[n, c] + rv[n, c] + sump[n, c / Dv] * bv[c]   (rv, bv nullable) */
int spt_vrpe_epilogue(const float* agg, const float* rv, const float* sump, const float* bv,
                      int64_t N, int H, int Dv, int heads_share, float* y, void* stream);

/* parameter gradients: dbv[c] += sum_n dy[n, c] sump[n, c / Dv];
 * dWv += diagonal blocks of dWbd [C, H*F] (summed over heads when shared).  Accumulating. */
int spt_vrpe_bwd_params(const float* dy, const float* sump, const float* dWbd, int64_t N, int H,
                        int Dv, int F, int heads_share, float* dWv /*nullable*/,
                        float* dbv /*nullable*/, void* stream);

/* ------------------------------------------------------------------------- *
 *  Segment pooling  (src/nn/pool.py:44-82 -> PyG *Aggregation -> scatter)    *
 * ------------------------------------------------------------------------- */

/* out[p, c] = reduce over children i in points[ptr[p]:ptr[p+1]] of x[i, c].
 * Empty segment -> 0 (torch_scatter / PyG semantics).  For MAX/MIN `arg`
 * (int32 [Np, C], nullable) receives the child row of the selected element
 * (first in child order on ties; -1 for empty segments). */
int spt_segment_pool_fwd(const float* x, const int32_t* ptr,
                         const int32_t* points /*nullable = identity*/,
                         int64_t num_parents, int64_t C, int reduce, float* out,
                         int32_t* arg /*nullable*/, void* stream);

/* torch_scatter scatter_mean / scatter_std over CSR segments (reference call sites
 * src/transforms/graph.py:266-285 — per-segment mean / std of point attributes — and
 * :1025-1044; semantics SURVEY.md Appendix A): mean = sum / max(count, 1),
 * std = sqrt(sum (x - mean)^2 / (max(count - 1, 1) + 1e-6)).  Either output may be NULL. */
int spt_segment_mean_std_fwd(const float* x, const int32_t* ptr,
                             const int32_t* points /*nullable = identity*/,
                             int64_t num_parents, int64_t C, float* mean_out /*nullable*/,
                             float* std_out /*nullable*/, void* stream);

/* Superedge descriptors from their level-0 sub-edges: replaces
 * _minimalistic_horizontal_edge_features (src/transforms/graph.py:950-1060): for superedge s
 * with sub-edges j in perm[ptr[s]:ptr[s+1]] (CSR of `se_id`), offsets o_j =
 * points[sp_dst[j]] - points[sp_src[j]]:
 *   out[s, 0:3] = mean o_j ; out[s, 3:6] = clip(unbiased std of o_j in the orthonormal base
 *   built around the mean offset (src/utils/geometry.py:42-77), -2, 2) ;
 *   out[s, 6] = sqrt(mean |o_j|). */
int spt_superedge_features_fwd(const float* points /*[N0,3]*/, const int64_t* sp_src,
                               const int64_t* sp_dst, const int32_t* ptr, const int32_t* perm,
                               int64_t num_superedges, float* out /*[num_superedges,7]*/,
                               void* stream);

/* dx[i, c] from dout[parent(i), c]; gather-form backward (no atomics):
 *   SUM : dx = dout[parent]          MEAN: dx = dout[parent] / max(count, 1)
 *   MAX/MIN: dx = dout[parent] if arg[parent, c] == i else 0               */
int spt_segment_pool_bwd(const float* dout, const int64_t* parent,
                         const int32_t* ptr, const int32_t* arg /*nullable*/,
                         int64_t num_children, int64_t C, int reduce, float* dx,
                         void* stream);

/* ------------------------------------------------------------------------- *
 *  UnitSphereNorm (src/nn/norm.py:53-138, scatter_mean_weighted              *
 *  src/utils/scatter.py:17-38)                                              *
 * ------------------------------------------------------------------------- */

/* Per parent: bbox min/max -> diameter = max span; w-weighted centroid.
 * pos_out[i] = (pos[i] - center[parent(i)]) / (diameter[parent(i)] + 1e-2).
 * parent==NULL (and num_parents==1, ptr={0,N}, points==NULL) is the
 * `_forward` (no idx) case.  w==NULL -> unweighted mean.  */
size_t spt_unitsphere_workspace_bytes(int64_t num_parents);
int spt_unitsphere_fwd(const float* pos, const int64_t* parent /*nullable*/,
                       const int32_t* ptr, const int32_t* points /*nullable*/,
                       const float* w /*nullable*/, int64_t N,
                       int64_t num_parents, float* pos_out,
                       float* diameter /*[num_parents]*/, void* ws,
                       size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------- *
 *  GraphNorm (torch_geometric.nn.norm.GraphNorm, selected by                *
 *  configs/model/semantic/_attention.yaml:9-11, called through              *
 *  src/nn/transformer.py:258-265 and src/nn/mlp.py:89-94)                   *
 * ------------------------------------------------------------------------- */

/* y = weight * (x - mean_scale*mean[b]) * rstd[b] + bias, per graph b=batch[i].
 * batch==NULL -> single graph; with B == 1 `batch` is not read (every row is graph 0).
 * batch need not be sorted.  Outputs mean/rstd [B, C] are saved for backward.
 * Launches: memset + statistics + (finalise fused into) apply; backward: memset + statistics
 * + (coefficients and parameter gradients fused into) apply, when C % 4 == 0 and B*C <= 2048. */
size_t spt_graphnorm_workspace_bytes(int64_t B, int64_t C);
int spt_graphnorm_fwd(const float* x, const int64_t* batch /*nullable*/,
                      int64_t N, int64_t C, int64_t B, const float* weight,
                      const float* bias, const float* mean_scale, float eps,
                      float act_slope /* 1 = none; else fused LeakyReLU(slope),
                        the activation that follows the norm in src/nn/mlp.py:41-55 */,
                      float* y, float* mean /*[B,C]*/, float* rstd /*[B,C]*/,
                      void* ws, size_t ws_bytes, void* stream);
int spt_graphnorm_bwd(const float* x, const float* dy,
                      const int64_t* batch /*nullable*/, int64_t N, int64_t C,
                      int64_t B, const float* weight, const float* mean_scale,
                      const float* mean, const float* rstd,
                      const float* y_act /*forward output, needed iff act_slope != 1*/,
                      float act_slope, float* dx,
                      float* dweight /*[C]*/, float* dbias /*[C]*/,
                      float* dmean_scale /*[C]*/, void* ws, size_t ws_bytes,
                      void* stream);

/* Graph-wise GroupNorm (src/nn/norm.py:141-237, mode='graph') and PyG
 * LayerNorm(mode='graph') (= num_groups 1; the code default of
 * src/nn/transformer.py:137): statistics per (graph, channel group) over
 * nodes x group channels, y = weight * (x - mean) * rstd + bias.
 * eps_outside != 0 reproduces PyG LayerNorm called without `batch`
 * (x / (std + eps)); otherwise rstd = 1/sqrt(var + eps).
 * Workspace: spt_graphnorm_workspace_bytes(B, C). */
int spt_groupnorm_fwd(const float* x, const int64_t* batch /*nullable*/, int64_t N,
                      int64_t C, int64_t B, int64_t num_groups,
                      const float* weight /*nullable*/, const float* bias /*nullable*/,
                      float eps, int eps_outside, float* y, float* mean /*[B,C]*/,
                      float* rstd /*[B,C]*/, void* ws, size_t ws_bytes, void* stream);
int spt_groupnorm_bwd(const float* x, const float* dy,
                      const int64_t* batch /*nullable*/, int64_t N, int64_t C,
                      int64_t B, int64_t num_groups, const float* weight /*nullable*/,
                      const float* mean, const float* rstd, float eps, int eps_outside,
                      float* dx, float* dweight /*[C], nullable*/,
                      float* dbias /*[C], nullable*/, void* ws, size_t ws_bytes,
                      void* stream);

/* ------------------------------------------------------------------------- *
 *  Fused sparse graph attention core                                        *
 *  (src/nn/attention.py:202-315; src/nn/pool.py:196-233 for attentive pool)  *
 * ------------------------------------------------------------------------- */

/* Rows (queries) are CSR rows; edge j (CSR slot) has target col[j] and edge
 * features a[j, :F] (already permuted to CSR order).  For each row s, head h:
 *   q_e = q[s]*scale(s) + Wq a_e + bq      (Wq, Wk: [H*D, F] row-major)
 *   k_e = k[t] + Wk a_e + bk
 *   c   = <q_e, k_e>_h ; p = softmax_row(c)  (denominator + 1e-16, PyG softmax)
 *   agg_v[s,h,:] = sum_e p * v[t,h,:]         ([N, C])
 *   abar [s,h,:] = sum_e p * a_e              ([N, H, F])  -> v_rpe applied by
 *                                              the caller: Wv_h abar + bv*sump
 *   sump [s,h]   = sum_e p                    (1 unless the row is empty)
 * Saved for backward: m[s,h] (row max), z[s,h] (sum exp + 1e-16).
 * q/k/v are given as base pointers + leading dimensions so that the fused
 * qkv Linear output [N, 2HD+C] (attention.py:191-204) or separate q [Np,HD] /
 * kv [Nc, HD+C] buffers (pool.py:187-198) can be used without copies.
 * a, Wq, bq, Wk, bk may each be NULL (no RPE / no bias).                     */
int spt_attn_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk,
                 const float* v, int64_t ldv, const float* a /*[E,F]*/,
                 const int32_t* rowptr, const int32_t* col, int64_t num_rows,
                 int64_t E, int H, int D, int Dv, int F, const float* Wq,
                 const float* bq, const float* Wk, const float* bk,
                 int scale_mode, float scale_value, float* agg_v /*[R,H*Dv]*/,
                 float* abar /*[R,H,F] nullable*/, float* sump /*[R,H]*/,
                 float* m /*[R,H]*/, float* z /*[R,H]*/, void* stream);

/* Optional terms of the attention core (all pointers nullable; a NULL struct = none):
 *   q_e += q_row_add[s] + q_tgt_add[t] ; k_e += k_row_add[s]
 *     -> the node-difference encodings k_delta_rpe / q_delta_rpe of src/nn/attention.py:259-291:
 *        enc(x_t - x_s) = W x_t - W x_s + b, i.e. one dense product per node + these addends;
 *   softmax weights multiplied by drop_mask[e, h] (CSR order) AFTER the normalisation
 *     -> attention dropout (src/nn/attention.py:310-311), mask drawn by the caller.
 * Backward outputs: d_q_row_add[s] = sum_e dq_e, d_k_row_add[s] = sum_e dk_e (rows kernel),
 * d_q_tgt_add[t] = sum_{e -> t} dq_e (targets kernel).  With any of these the generic kernels run. */
typedef struct spt_attn_extras {
  const float* q_row_add; /* [R, H*D] */
  const float* q_tgt_add; /* [T, H*D] */
  const float* k_row_add; /* [R, H*D] */
  const float* drop_mask; /* [E, H]   */
  float* d_q_row_add;     /* [R, H*D] */
  float* d_k_row_add;     /* [R, H*D] */
  /* backward only: gradient flowing into sump [R, H] (non-zero only with a dropout mask, when
   * the caller's v-RPE bias multiplies sump) and the forward's sump */
  const float* d_sump;    /* [R, H] */
  const float* sump;      /* [R, H] */
  /* Workspace of the split kernels (csrc/attention_split.cuh: one edge-parallel pass on the
   * tensor cores + one row-parallel pass, 16 bytes per edge between them; shape families
   * H=4, D=4, Dv=32, F=32 and — csrc/attention_split16.cuh — H=16, D=4, Dv in {4, 8}, F=32).  With ws_logits and
   * edge_row set (and none of the optional terms above) the forward takes that path and leaves
   * the base-2 logits in ws_logits; the backward reads them back and needs ws_ds as well.
   * NULL = the fused row-tile kernels. */
  float* ws_logits;        /* [E, H] */
  const int32_t* edge_row; /* [E] CSR row of every slot (spt_expand_pointers_i32) */
  float* ws_ds;            /* [E, H] backward only */
  /* split kernels only: the gathered value rows stored as bf16 ([T, ldv_bf16] elements; fp32
   * accumulation).  Halves the bytes of the dominant gather and keeps it L2-resident — the
   * "bf16 storage" configuration (BASELINE cfg 3); q, k and the edge features stay fp32. */
  const uint16_t* v_bf16;
  int64_t ldv_bf16;
} spt_attn_extras;

int spt_attn_fwd_ex(const float* q, int64_t ldq, const float* k, int64_t ldk,
                    const float* v, int64_t ldv, const float* a, const int32_t* rowptr,
                    const int32_t* col, int64_t num_rows, int64_t E, int H, int D, int Dv, int F,
                    const float* Wq, const float* bq, const float* Wk, const float* bk,
                    int scale_mode, float scale_value, float* agg_v, float* abar, float* sump,
                    float* m, float* z, const spt_attn_extras* extras, void* stream);
int spt_attn_bwd_rows_ex(const float* q, int64_t ldq, const float* k, int64_t ldk,
                         const float* v, int64_t ldv, const float* a, const int32_t* rowptr,
                         const int32_t* col, int64_t num_rows, int64_t E, int H, int D, int Dv,
                         int F, const float* Wq, const float* bq, const float* Wk, const float* bk,
                         int scale_mode, float scale_value, const float* m, const float* z,
                         const float* agg_v, const float* abar, const float* d_agg_v,
                         const float* d_abar, float* dq, int64_t lddq, float* da, float* dWq,
                         float* dbq, float* dWk, float* dbk, float* Pbuf, float* G,
                         const spt_attn_extras* extras, void* stream);
int spt_attn_bwd_targets_ex(const int32_t* csc_ptr, const int32_t* csc_src,
                            const int32_t* csc2csr, int64_t num_targets, int64_t E, int H, int D,
                            int Dv, const float* Pbuf, const float* G, const float* d_agg_v,
                            float* dk, int64_t lddk, float* dv, int64_t lddv,
                            float* d_q_tgt_add /*[T, H*D] nullable*/, void* stream);

/* bf16 STORAGE of the attention operands (BASELINE cfg 3: "bf16, fp32 accumulate"): q / k / v
 * (leading dimensions in elements) and the CSR-ordered edge features a [E, 32] are bf16 in HBM,
 * weights / statistics / outputs / gradients stay fp32, every sum is fp32 (the RPE products run as
 * bf16 tensor-core MMAs with fp32 accumulation).  Only the shape family of the row-tile kernels
 * (H=4, D=4, Dv=32, F=32); anything else returns SPT_E_UNSUPPORTED (the caller uses the fp32
 * entry points).  The backward writes fp32 dq / da / P / G; d[Wq;Wk] (spt_gemm_tn_acc on G and
 * the fp32 features) and spt_attn_bwd_targets are the fp32 calls. */
int spt_cast_bf16(const float* x, int64_t n, uint16_t* out, void* stream);
int spt_attn_fwd_bf16(const uint16_t* q, int64_t ldq, const uint16_t* k, int64_t ldk,
                      const uint16_t* v, int64_t ldv, const uint16_t* a /*[E,32]*/,
                      const int32_t* rowptr, const int32_t* col, int64_t num_rows, int64_t E,
                      int H, int D, int Dv, int F, const float* Wq, const float* bq,
                      const float* Wk, const float* bk, int scale_mode, float scale_value,
                      float* agg_v, float* abar, float* sump, float* m, float* z, void* stream);
int spt_attn_bwd_rows_bf16(const uint16_t* q, int64_t ldq, const uint16_t* k, int64_t ldk,
                           const uint16_t* v, int64_t ldv, const uint16_t* a,
                           const int32_t* rowptr, const int32_t* col, int64_t num_rows, int64_t E,
                           int H, int D, int Dv, int F, const float* Wq, const float* bq,
                           const float* Wk, const float* bk, int scale_mode, float scale_value,
                           const float* m, const float* z, const float* agg_v, const float* abar,
                           const float* d_agg_v, const float* d_abar, float* dq, int64_t lddq,
                           float* da, float* Pbuf, float* G, void* stream);

/* Backward of spt_attn_fwd, three launches so each can be timed on its own:
 *  (1) rows    : per CSR row, recompute p from (m, z); writes dq [R rows, lddq],
 *                da [E,F] (CSR order, nullable) and the per-edge scratch
 *                P [E,H] (= p) and G [E,2HD] (= [dq_e | dk_e]; the specialised
 *                kernel only fills the dk_e half).  dWq,dWk [HD,F] / dbq,dbk [HD]
 *                (nullable) are ACCUMULATED into (caller zero-fills): fused in the
 *                specialised kernel, a follow-up slab reduction (= entry (3)) in
 *                the generic one.
 *  (2) targets : per target t (edges grouped by target: csc_ptr/csc_src/csc2csr,
 *                csc2csr = CSR slot of each CSC slot), gathers
 *                dv[t] = sum p * d_agg_v[src], dk[t] = sum dk_e — no atomics;
 *                every target row is written (zeros when it has no incoming edge).
 *  (3) weights : dWq,dWk [HD,F] += G^T a ; dbq,dbk [HD] += colsum(G)
 *                (ACCUMULATED with fp32 atomics; the caller zero-fills).        */
int spt_attn_bwd_rows(const float* q, int64_t ldq, const float* k, int64_t ldk,
                      const float* v, int64_t ldv, const float* a,
                      const int32_t* rowptr, const int32_t* col, int64_t num_rows,
                      int64_t E, int H, int D, int Dv, int F, const float* Wq,
                      const float* bq, const float* Wk, const float* bk,
                      int scale_mode, float scale_value, const float* m,
                      const float* z, const float* agg_v, const float* abar,
                      const float* d_agg_v, const float* d_abar /*nullable*/,
                      float* dq, int64_t lddq, float* da /*nullable*/,
                      float* dWq, float* dbq, float* dWk, float* dbk /*nullable*/,
                      float* P, float* G, void* stream);
int spt_attn_bwd_targets(const int32_t* csc_ptr, const int32_t* csc_src,
                         const int32_t* csc2csr, int64_t num_targets, int64_t E,
                         int H, int D, int Dv, const float* P, const float* G,
                         const float* d_agg_v, float* dk, int64_t lddk, float* dv,
                         int64_t lddv, void* stream);
int spt_attn_bwd_weights(const float* G, const float* a, int64_t E, int H, int D,
                         int F, float* dWq, float* dbq, float* dWk, float* dbk,
                         void* stream);

/* ------------------------------------------------------------------------- *
 *  On-the-fly horizontal edge features                                      *
 *  (src/transforms/graph.py:1137-1277 + NAGAddSelfLoops :1419-1452)          *
 * ------------------------------------------------------------------------- */

/* From the trimmed graph se [2, Eh] (int64) and its 7-column attributes
 * ea (fp32 [Eh,7]: mean_off 3, std_off 3, mean_dist 1) builds
 *   edge_index_out [2, 2*Eh + (self_loops? N:0)]  = [se | flip(se) | loops]
 *   edge_attr_out  [same, 18] fp32 in the reference column order
 *     mean_off(3) std_off(3) mean_dist angle_source angle_target normal_angle
 *     log_length log_surface log_volume log_size centroid_dir(3) centroid_dist
 *   self-loop rows are zero (add_self_loops fill_value=0).                   */
int spt_edge_features_fwd(const int64_t* se, const float* ea, const float* pos,
                          const float* normal, const float* log_length,
                          const float* log_surface, const float* log_volume,
                          const float* log_size, int64_t Eh, int64_t N,
                          int add_self_loops, int64_t* edge_index_out,
                          float* edge_attr_out, void* stream);

/* On-the-fly VERTICAL (child -> parent) edge features, default key set of
 * _on_the_fly_vertical_edge_features (src/transforms/graph.py:1336-1416):
 *   v_edge_attr[i] = [centroid_dir(3), sqrt(centroid_dist), |n_child . n_parent|,
 *                     d log_length, d log_surface, d log_volume, d log_size]  ([Nc, 9])
 * child_logs [4, Nc] / parent_logs [4, Np] = stacked (log_length, log_surface,
 * log_volume, log_size); parent = super_index of the child level (int64). */
int spt_vertical_edge_features_fwd(const float* child_pos, const float* parent_pos,
                                   const float* child_normal, const float* parent_normal,
                                   const float* child_logs, const float* parent_logs,
                                   const int64_t* parent, int64_t Nc, int64_t Np,
                                   float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPT_B200_H */
