// attention_umma.cuh — the EDGE passes of the split attention kernels on the tcgen05 tensor cores.
//
// The relative-position encodings R_e = [Wq;Wk] a_e (src/nn/attention.py:225-256 of the
// reference: three per-edge Linear layers) are a dense [E, 32] x [32, 32] product.  One
// persistent CTA per SM works on tiles of 128 consecutive CSR slots:
//   warp 0       TMA producer: 128 x 32 fp32 feature rows (SWIZZLE_128B, L2 evict-first) into a
//                ring of 8 landing slots (128 KB in flight per SM)
//   warps 8-11   splitter: row -> registers -> TF32 hi / lo -> tcgen05.st (A operand in TMEM)
//   warp 1       MMA issuer: 12 x tcgen05.mma.kind::tf32 (M=128, N=32, K=8; lo.hi + hi.lo + hi.hi)
//                per tile, weights resident in shared memory (pre-split, K-major SWIZZLE_128B)
//   warps 4-7, 12-15  two epilogue groups, one per TMEM accumulator, alternating tiles:
//                THREAD = EDGE: tcgen05.ld of the edge's 32 outputs, q row / gathered k row,
//                4 logits, one 16-byte store
//   warp 2       TMEM allocation
// 3xTF32 keeps the product at fp32 accuracy (~2^-21 relative), like csrc/gemm_umma.cu.
#pragma once
#include "umma_common.cuh"
#include "attention_split.cuh"

namespace spt {
namespace umma {  // csrc/gemm_umma.cu
bool make_map_rows32(CUtensorMap* tm, const float* ptr, int64_t rows, int64_t ld, int box_rows);
bool make_map_box(CUtensorMap* tm, const float* ptr, int64_t rows, int64_t cols, int64_t ld,
                  int box_cols, int box_rows);
}
namespace aumma {

using namespace umma;   // mbarrier / TMA / tcgen05 helpers of umma_common.cuh

constexpr int kThreads = 512;
constexpr int kRing = 6;                       // landing slots of 16 KB
constexpr int kAStages = 4;                    // hi / lo operand stages in TMEM (64 columns each)
constexpr int kTileRows = 128;
constexpr uint32_t kTileBytes = kTileRows * 32 * 4;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kAccCols = 32;              // accumulator a: columns [32 a, 32 a + 32)
constexpr uint32_t kAStage0 = 64;              // stage s: hi at 64 + 64 s, lo at + 32
constexpr int kWBytes = 32 * 32 * 4;           // one pre-split weight matrix

struct Smem {
  static constexpr int ring_off = 0;
  static constexpr int qk_off = kRing * (int)kTileBytes;     // [2 groups][2 stages][128][q 16 | k 16]
  static constexpr int ext_off = qk_off + 4 * (int)kTileBytes;  // [2][2][128][rowptr[s], rowptr[s+1], s, -]
  static constexpr int bhi_off = ext_off + 4 * kTileRows * 16;
  static constexpr int blo_off = bhi_off + kWBytes;
  static constexpr int bias_off = blo_off + kWBytes;
  static constexpr int bar_off = bias_off + 128;
  // a_full[R] a_free[R] a_ready[4] a_empty[4] tmem_full[2] tmem_empty[2] + tmem slot
  static constexpr int total = bar_off + (2 * kRing + 2 * kAStages + 4) * 8 + 16;
};

// ---- the epilogue's operands, staged two tiles ahead ------------------------------------
// Thread = edge needs the q row of its source and the gathered k row of its target: two
// dependent global round trips (slot -> row / column id -> 64-byte rows).  Each epilogue thread
// copies them with cp.async into its own 128-byte row of a 2-stage buffer for the tile it will
// process two turns later; the ids of the tile after that wait in registers.
__device__ __forceinline__ void cp_async16_ca(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
struct EdgeIds {
  int rr, c;
  bool valid;
};
__device__ __forceinline__ EdgeIds load_ids(const int32_t* edge_row, const int32_t* col,
                                            int64_t t, int64_t tiles, int row, int64_t E) {
  EdgeIds id;
  const int64_t e = t * kTileRows + row;
  id.valid = t < tiles && e < E;
  id.rr = 0; id.c = 0;
  if (id.valid) { id.rr = edge_row[e]; id.c = col[e]; }
  return id;
}
// q chunks 0..3, k chunks 4..7 of a 128-byte stage row, chunk j at ((j ^ (row & 7)) << 4):
// conflict-free 16-byte reads by the thread that owns the row, and the SWIZZLE_128B pattern of
// the TMA stores that later leave from the same rows.
// The copy is COOPERATIVE: a row-per-thread gather costs one L1 wavefront per lane and
// instruction (32 per LDG/LDGSTS, the limiter of the first version); here 4 lanes fetch the 4
// chunks of one 64-byte row, so an instruction touches 8 rows = 8 wavefronts.  `wbase` is the
// stage address of the warp's first row; ids travel by shuffle from the lanes that own them.
__device__ __forceinline__ void stage_qk(uint32_t wbase, int lane, const EdgeIds& id,
                                         const float* q, int ldq, const float* k, int ldk) {
  const int sub = lane & 3, r8 = lane >> 2;
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int rl = 8 * m + r8;
    const int rr = __shfl_sync(0xffffffffu, id.rr, rl);
    const int c = __shfl_sync(0xffffffffu, id.c, rl);
    const int v = __shfl_sync(0xffffffffu, (int)id.valid, rl);
    if (v) {
      const uint32_t dst = wbase + (uint32_t)rl * 128u;
      const int sw = rl & 7;
      cp_async16_ca(dst + (uint32_t)((sub ^ sw) << 4),
                    reinterpret_cast<const char*>(q + (int64_t)rr * ldq) + 16 * sub);
      cp_async16_cg(dst + (uint32_t)(((4 + sub) ^ sw) << 4),
                    reinterpret_cast<const char*>(k + (int64_t)c * ldk) + 16 * sub);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// extent of my edge's source row (-> q scale) and its id, parked in shared memory: nothing in
// registers waits on these loads until the tile is processed two turns later
__device__ __forceinline__ void stage_extent(uint32_t dst, const EdgeIds& id, const int32_t* rowptr) {
  if (id.valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(rowptr + id.rr) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4), "l"(rowptr + id.rr + 1)
                 : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + 8), "r"(id.rr) : "memory");
  }
}
__device__ __forceinline__ int4 lds_extent(uint32_t src) {
  int4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(src) : "memory");
  return v;
}

__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* tm, int c0, int c1,
                                                 uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}

// [Wq;Wk] ([32 outputs, 32 features], K-major) as TF32 hi / lo in the SWIZZLE_128B layout the
// UMMA smem descriptor reads: output row o = 128 bytes, 16-byte chunk c at ((c ^ (o & 7)) << 4)
__device__ __forceinline__ void stage_weights(unsigned char* bhi, unsigned char* blo, float* bias_s,
                                              const float* Wq, const float* bq, const float* Wk,
                                              const float* bk) {
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int o = i >> 5, f = i & 31;
    const float* W = o < 16 ? Wq : Wk;
    const float w = W ? W[(o & 15) * 32 + f] : 0.f;
    const float hi = tf32_rna(w), lo = tf32_rna(w - hi);
    const int off = o * 128 + (((f >> 2) ^ (o & 7)) << 4) + (f & 3) * 4;
    *reinterpret_cast<float*>(bhi + off) = hi;
    *reinterpret_cast<float*>(blo + off) = lo;
  }
  for (int o = threadIdx.x; o < 32; o += blockDim.x) {
    const float* W = o < 16 ? Wq : Wk;
    const float* b = o < 16 ? bq : bk;
    bias_s[o] = (W && b) ? b[o & 15] : 0.f;
  }
  fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async proxy
}

// splitter role: landing slot row -> hi / lo -> TMEM operand stage (as in k_gemm_nt_umma)
__device__ __forceinline__ void split_row_to_tmem(uint32_t arow, int row, uint32_t ta) {
  float4 x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = lds128(arow + (uint32_t)((j ^ (row & 7)) << 4));
  uint32_t hi[32], lo[32];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float xs[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float h = tf32_rna(xs[c]);
      hi[4 * j + c] = __float_as_uint(h);
      lo[4 * j + c] = __float_as_uint(xs[c] - h);   // fed raw: the tensor core ignores the low 13 bits
    }
  }
  tmem_st32(ta, hi);
  tmem_st32(ta + 32, lo);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// 12 MMAs of one 128 x 32 x 32 tile: D = Alo.Bhi + Ahi.Blo + Ahi.Bhi
__device__ __forceinline__ void issue_tile_mma(uint32_t d, uint32_t a_hi, uint64_t b_hi,
                                               uint64_t b_lo, uint32_t idesc) {
  const uint32_t a_lo = a_hi + 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint64_t o = (uint64_t)(k * 8 * 4 >> 4);   // advance inside the 128-byte swizzle row
    umma_tf32_ts(d, a_lo + k * 8, b_hi + o, idesc, k != 0);
    umma_tf32_ts(d, a_hi + k * 8, b_lo + o, idesc, 1);
    umma_tf32_ts(d, a_hi + k * 8, b_hi + o, idesc, 1);
  }
}

constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) |
                            ((uint32_t)(128 >> 4) << 24);

__global__ void __launch_bounds__(kThreads, 1)
k_edge_logits_umma(const __grid_constant__ CUtensorMap tmA, const split::EdgeFwdArgs P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem =
      (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // swizzle atoms
  unsigned char* ringbuf = smem + Smem::ring_off;
  unsigned char* bhi = smem + Smem::bhi_off;
  unsigned char* blo = smem + Smem::blo_off;
  float* bias_s = reinterpret_cast<float*>(smem + Smem::bias_off);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + Smem::bar_off);
  uint64_t* a_free = a_full + kRing;
  uint64_t* a_ready = a_free + kRing;
  uint64_t* a_empty = a_ready + kAStages;
  uint64_t* tmem_full = a_empty + kAStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tiles = (P.E + kTileRows - 1) / kTileRows;

  if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
  if (warp == 1 && lane == 0) {
    for (int r = 0; r < kRing; ++r) { mbar_init(&a_full[r], 1); mbar_init(&a_free[r], 128); }
    for (int s = 0; s < kAStages; ++s) { mbar_init(&a_ready[s], 128); mbar_init(&a_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  stage_weights(bhi, blo, bias_s, P.Wq, P.bq, P.Wk, P.bk);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    const uint64_t policy = tile::policy_evict_first();
    uint32_t ph = 0;
    int r = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      mbar_wait(&a_free[r], ph ^ 1, 0);
      if (elect_one()) {
        mbar_expect_tx(&a_full[r], kTileBytes);
        tma_load_2d_hint(ringbuf + (size_t)r * kTileBytes, &tmA, 0, (int)(t * kTileRows),
                         &a_full[r], policy);
      }
      __syncwarp();
      if (++r == kRing) { r = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    const uint64_t b_hi = smem_desc_sw128(smem_u32(bhi)), b_lo = smem_desc_sw128(smem_u32(blo));
    uint32_t tl = 0, pha = 0;
    int sa = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++tl) {
      const uint32_t acc = tl & 1, accph = (tl >> 1) & 1;
      mbar_wait(&tmem_empty[acc], accph ^ 1, 2);
      mbar_wait(&a_ready[sa], pha, 3);
      tc_fence_after();
      if (elect_one()) {
        issue_tile_mma(tmem_base + acc * kAccCols, tmem_base + kAStage0 + 64 * sa, b_hi, b_lo,
                       kIdesc);
        umma_commit(&a_empty[sa]);
        umma_commit(&tmem_full[acc]);
      }
      __syncwarp();
      if (++sa == kAStages) { sa = 0; pha ^= 1; }
    }
  } else if (warp >= 8 && warp < 12) {
    // ---------------- splitter ----------------
    const int q = warp - 8, row = q * 32 + lane;
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int r = it % kRing, s = it % kAStages;
      mbar_wait(&a_full[r], (it / kRing) & 1, 4);
      mbar_wait(&a_empty[s], ((it / kAStages) & 1) ^ 1, 5);    // the MMAs that read it retired
      tc_fence_after();
      split_row_to_tmem(smem_u32(ringbuf) + (uint32_t)r * kTileBytes + (uint32_t)row * 128u, row,
                        tmem_base + kAStage0 + 64 * s + ((uint32_t)(q * 32) << 16));
      mbar_arrive(&a_free[r]);      // after the tcgen05.st consumed the loaded values
      tc_fence_before();
      mbar_arrive(&a_ready[s]);
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: thread = edge ----------------
    const int g = warp >= 12 ? 1 : 0;
    const int q = warp & 3, row = q * 32 + lane, sw = row & 7;
    const uint32_t wbase = smem_u32(smem + Smem::qk_off) + (uint32_t)g * 2 * kTileBytes +
                           (uint32_t)(q * 32) * 128u;      // my warp's rows in stage 0
    const uint32_t my = wbase + (uint32_t)lane * 128u;
    const int64_t stride = 2 * (int64_t)gridDim.x;
    const int64_t t0 = blockIdx.x + (int64_t)g * gridDim.x;
    const uint32_t myext = smem_u32(smem + Smem::ext_off) + (uint32_t)g * 2 * (kTileRows * 16) +
                           (uint32_t)row * 16u;
    EdgeIds id = load_ids(P.edge_row, P.col, t0, tiles, row, P.E);
    stage_extent(myext, id, P.rowptr);
    stage_qk(wbase, lane, id, P.q, P.ldq, P.k, P.ldk);
    id = load_ids(P.edge_row, P.col, t0 + stride, tiles, row, P.E);
    stage_extent(myext + kTileRows * 16, id, P.rowptr);
    stage_qk(wbase + kTileBytes, lane, id, P.q, P.ldq, P.k, P.ldk);
    id = load_ids(P.edge_row, P.col, t0 + 2 * stride, tiles, row, P.E);
    uint32_t n = 0;
    for (int64_t t = t0; t < tiles; t += stride, ++n) {
      const uint32_t st = n & 1;
      const int64_t e = t * kTileRows + row;
      const bool valid = e < P.E;
      asm volatile("cp.async.wait_group 1;" ::: "memory");   // this tile's rows have landed
      __syncwarp();                                          // (copied by my warp's lanes)
      const int4 ext = lds_extent(myext + st * (kTileRows * 16));
      const float scale =
          fast::qk_scale_fast(P.scale_mode, P.scale_value, valid ? ext.y - ext.x : 1);
      mbar_wait(&tmem_full[g], n & 1, 7);
      tc_fence_after();
      uint32_t R[32];
      tmem_ld32(tmem_base + g * kAccCols + ((uint32_t)(q * 32) << 16), R);
      tc_fence_before();
      mbar_arrive(&tmem_empty[g]);
      const uint32_t src = my + st * kTileBytes;
      float lg[4];
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float4 q4 = lds128(src + (uint32_t)((h ^ sw) << 4));
        const float4 k4 = lds128(src + (uint32_t)(((4 + h) ^ sw) << 4));
        const float qs[4] = {q4.x, q4.y, q4.z, q4.w};
        const float ks[4] = {k4.x, k4.y, k4.z, k4.w};
        float s = 0.f;
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const float qe = fmaf(qs[d], scale, __uint_as_float(R[4 * h + d]) + bias_s[4 * h + d]);
          const float ke = ks[d] + __uint_as_float(R[16 + 4 * h + d]) + bias_s[16 + 4 * h + d];
          s = fmaf(qe, ke, s);
        }
        lg[h] = s * fast::kLog2e;
      }
      if (valid)
        *reinterpret_cast<float4*>(P.logits + e * P.ldl) = make_float4(lg[0], lg[1], lg[2], lg[3]);
      // the warp's rows of this stage are free: stage the tile two turns ahead, fetch the ids
      // after it
      __syncwarp();
      stage_extent(myext + st * (kTileRows * 16), id, P.rowptr);
      stage_qk(wbase + st * kTileBytes, lane, id, P.q, P.ldq, P.k, P.ldk);
      id = load_ids(P.edge_row, P.col, t + 3 * stride, tiles, row, P.E);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(kTmemCols)
                 : "memory");
  }
}

// host: false = tensor map could not be encoded (the caller runs the CUDA-core pass)
inline bool edge_logits_launch(const split::EdgeFwdArgs& A, cudaStream_t st, int* rc) {
  CUtensorMap tm;
  if (!umma::make_map_rows32(&tm, A.a, A.E, 32, kTileRows)) return false;
  const int smem = Smem::total + 1024;
  static unsigned long long done = 0;
  ensure_dynamic_smem(k_edge_logits_umma, smem, &done);
  const int64_t tiles = (A.E + kTileRows - 1) / kTileRows;
  const int64_t sms = device_sm_count();
  k_edge_logits_umma<<<(unsigned)(tiles < sms ? tiles : sms), kThreads, smem, st>>>(tm, A);
  *rc = check_launch("attn_fwd(edge, umma)");
  return true;
}


// ------------------------------------------------------------------------------------------
// backward edge pass: per tile of 128 edges
//   MMA 1   R = a [Wq;Wk]^T                       (as in the forward)
//   epilogue 1 (thread = edge)  q_e, k_e, dS  ->  G_e = [dS k_e | dS q_e]  -> global (targets
//           kernel, d[Wq;Wk] product) and, split into TF32 hi / lo, into a TMEM operand stage
//   MMA 2   da = G [Wq;Wk]                        (A operand = G from TMEM, into the same
//                                                  accumulator columns, R is dead by then)
//   epilogue 2  da += sum_h p_e,h dAbar_s,h ; store
// TMEM: 2 accumulators (one per epilogue group) + 4 feature stages + 2 G stages = 448 columns.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kGStage0 = kAStage0 + 64 * kAStages;    // G stage g: hi at +64 g, lo at +32

struct SmemBwd {
  static constexpr int ring_off = 0;
  static constexpr int qk_off = kRing * (int)kTileBytes;     // [2 groups][2 stages][128][q | k]
  static constexpr int sp_off = qk_off + 4 * (int)kTileBytes;   // [2][2][128][dS 4 | P 4]
  static constexpr int ext_off = sp_off + 4 * kTileRows * 32;   // [2][2][128][rowptr[s], rowptr[s+1], s, -]
  static constexpr int bhi_off = ext_off + 4 * kTileRows * 16;  // [Wq;Wk]   rows = outputs  (MMA 1)
  static constexpr int blo_off = bhi_off + kWBytes;
  static constexpr int thi_off = blo_off + kWBytes;          // [Wq;Wk]^T rows = features (MMA 2)
  static constexpr int tlo_off = thi_off + kWBytes;
  static constexpr int bias_off = tlo_off + kWBytes;
  static constexpr int bar_off = bias_off + 128;
  // a_full[8] a_free[8] a_ready[4] a_empty[4] r_full[2] acc_free[2] g_ready[2] d_full[2] + slot
  static constexpr int total = bar_off + (2 * kRing + 2 * kAStages + 8) * 8 + 16;
};

// [Wq;Wk]^T: row f (feature) = the 32 outputs, K-major for da[e, f] = sum_o G[e, o] W[o, f]
__device__ __forceinline__ void stage_weights_t(unsigned char* thi, unsigned char* tlo,
                                                const float* Wq, const float* Wk) {
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int f = i >> 5, o = i & 31;
    const float* W = o < 16 ? Wq : Wk;
    const float w = W ? W[(o & 15) * 32 + f] : 0.f;
    const float hi = tf32_rna(w), lo = tf32_rna(w - hi);
    const int off = f * 128 + (((o >> 2) ^ (f & 7)) << 4) + (o & 3) * 4;
    *reinterpret_cast<float*>(thi + off) = hi;
    *reinterpret_cast<float*>(tlo + off) = lo;
  }
  fence_proxy_async();
}

__global__ void __launch_bounds__(kThreads, 1)
k_edge_bwd_umma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmG,
                const __grid_constant__ CUtensorMap tmD, const split::EdgeBwdArgs P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem =
      (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  using L = SmemBwd;
  unsigned char* ringbuf = smem + L::ring_off;
  float* bias_s = reinterpret_cast<float*>(smem + L::bias_off);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + L::bar_off);
  uint64_t* a_free = a_full + kRing;
  uint64_t* a_ready = a_free + kRing;
  uint64_t* a_empty = a_ready + kAStages;
  uint64_t* r_full = a_empty + kAStages;     // [2] MMA 1 of the group's tile has landed
  uint64_t* acc_free = r_full + 2;           // [2] the group has read da (accumulator reusable)
  uint64_t* g_ready = acc_free + 2;          // [2] G hi / lo of the group's tile is in TMEM
  uint64_t* d_full = g_ready + 2;            // [2] MMA 2 has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tiles = (P.E + kTileRows - 1) / kTileRows;

  if (warp == 0 && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
  if (warp == 1 && lane == 0) {
    for (int r = 0; r < kRing; ++r) { mbar_init(&a_full[r], 1); mbar_init(&a_free[r], 128); }
    for (int s = 0; s < kAStages; ++s) { mbar_init(&a_ready[s], 128); mbar_init(&a_empty[s], 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&r_full[a], 1);
      mbar_init(&acc_free[a], 128);
      mbar_init(&g_ready[a], 128);
      mbar_init(&d_full[a], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  stage_weights(smem + L::bhi_off, smem + L::blo_off, bias_s, P.Wq, P.bq, P.Wk, P.bk);
  stage_weights_t(smem + L::thi_off, smem + L::tlo_off, P.Wq, P.Wk);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    const uint64_t policy = tile::policy_evict_first();
    uint32_t ph = 0;
    int r = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
      mbar_wait(&a_free[r], ph ^ 1, 0);
      if (elect_one()) {
        mbar_expect_tx(&a_full[r], kTileBytes);
        tma_load_2d_hint(ringbuf + (size_t)r * kTileBytes, &tmA, 0, (int)(t * kTileRows),
                         &a_full[r], policy);
      }
      __syncwarp();
      if (++r == kRing) { r = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: MMA1(0), MMA1(1), MMA2(0), MMA1(2), MMA2(1), ... ----------
    const uint64_t b_hi = smem_desc_sw128(smem_u32(smem + L::bhi_off));
    const uint64_t b_lo = smem_desc_sw128(smem_u32(smem + L::blo_off));
    const uint64_t t_hi = smem_desc_sw128(smem_u32(smem + L::thi_off));
    const uint64_t t_lo = smem_desc_sw128(smem_u32(smem + L::tlo_off));
    uint32_t tl = 0, pha = 0;
    int sa = 0;
    int64_t n_mine = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) ++n_mine;
    for (int64_t i = 0; i <= n_mine; ++i, ++tl) {
      if (i < n_mine) {                       // MMA 1 of tile tl
        const uint32_t acc = tl & 1, accph = (tl >> 1) & 1;
        mbar_wait(&acc_free[acc], accph ^ 1, 2);
        mbar_wait(&a_ready[sa], pha, 3);
        tc_fence_after();
        if (elect_one()) {
          issue_tile_mma(tmem_base + acc * kAccCols, tmem_base + kAStage0 + 64 * sa, b_hi, b_lo,
                         kIdesc);
          umma_commit(&a_empty[sa]);
          umma_commit(&r_full[acc]);
        }
        __syncwarp();
        if (++sa == kAStages) { sa = 0; pha ^= 1; }
      }
      if (i >= 1) {                           // MMA 2 of tile tl - 1
        const uint32_t pt = tl - 1, acc = pt & 1, accph = (pt >> 1) & 1;
        mbar_wait(&g_ready[acc], accph, 6);
        tc_fence_after();
        if (elect_one()) {
          issue_tile_mma(tmem_base + acc * kAccCols, tmem_base + kGStage0 + 64 * acc, t_hi, t_lo,
                         kIdesc);
          umma_commit(&d_full[acc]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // ---------------- splitter ----------------
    const int q = warp - 8, row = q * 32 + lane;
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int r = it % kRing, s = it % kAStages;
      mbar_wait(&a_full[r], (it / kRing) & 1, 4);
      mbar_wait(&a_empty[s], ((it / kAStages) & 1) ^ 1, 5);
      tc_fence_after();
      split_row_to_tmem(smem_u32(ringbuf) + (uint32_t)r * kTileBytes + (uint32_t)row * 128u, row,
                        tmem_base + kAStage0 + 64 * s + ((uint32_t)(q * 32) << 16));
      mbar_arrive(&a_free[r]);
      tc_fence_before();
      mbar_arrive(&a_ready[s]);
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: thread = edge ----------------
    const int g = warp >= 12 ? 1 : 0;
    const int q = warp & 3, row = q * 32 + lane, sw = row & 7;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t wbase = smem_u32(smem + L::qk_off) + (uint32_t)g * 2 * kTileBytes +
                           (uint32_t)(q * 32) * 128u;      // my warp's rows in stage 0
    const uint32_t my = wbase + (uint32_t)lane * 128u;
    const uint32_t mysp = smem_u32(smem + L::sp_off) + (uint32_t)g * 2 * (kTileRows * 32) +
                          (uint32_t)row * 32u;
    const int64_t stride = 2 * (int64_t)gridDim.x;
    const int64_t t0 = blockIdx.x + (int64_t)g * gridDim.x;
    const uint32_t myext = smem_u32(smem + L::ext_off) + (uint32_t)g * 2 * (kTileRows * 16) +
                           (uint32_t)row * 16u;
    // stage = q / k rows (cooperative copy) + dS / P and the row extent / id of my edge; the
    // dAbar row of the edge is pulled into the L2 on the way
    auto stage_all = [&](uint32_t st, const EdgeIds& id, int64_t t) {
      if (id.valid) {
        const int64_t e = t * kTileRows + row;
        cp_async16_cg(mysp + st * (kTileRows * 32), P.dS + e * P.ld_ds);
        if (P.d_abar) {
          cp_async16_cg(mysp + st * (kTileRows * 32) + 16, P.Pbuf + e * P.ld_ds);
          const char* dab = reinterpret_cast<const char*>(P.d_abar + (int64_t)id.rr * P.ld_dab);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(dab + 128 * j) : "memory");
        }
      }
      stage_extent(myext + st * (kTileRows * 16), id, P.rowptr);
      stage_qk(wbase + st * kTileBytes, lane, id, P.q, P.ldq, P.k, P.ldk);   // commits the group
    };
    // 32 x 128-byte rows of my warp leave through one TMA store (clipped at E by the map)
    auto store_rows = [&](const CUtensorMap* tm, uint32_t src_wbase, int64_t t) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::
                         "l"(tm), "r"(src_wbase), "r"(0), "r"((int)(t * kTileRows + q * 32))
                     : "memory");
        bulk_commit();
      }
    };
    // da of a later head group: added to what the earlier launches stored (TMA reduce-add)
    auto reduce_rows = [&](const CUtensorMap* tm, uint32_t src_wbase, int64_t t) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        asm volatile(
            "cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::
                "l"(tm), "r"(src_wbase), "r"(0), "r"((int)(t * kTileRows + q * 32))
            : "memory");
        bulk_commit();
      }
    };
    // G of one head group of a 16-head problem: two 16-column boxes of G [E, 128] (unswizzled
    // 64-byte staging rows: dq_e half at wsrc, dk_e half at wsrc + 2 KB)
    auto store_halves = [&](const CUtensorMap* tm, uint32_t src_wbase, int64_t t) {
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        const int r0 = (int)(t * kTileRows + q * 32);
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::
                         "l"(tm), "r"(src_wbase), "r"(P.g_col_q), "r"(r0)
                     : "memory");
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::
                         "l"(tm), "r"(src_wbase + 2048u), "r"(P.g_col_k), "r"(r0)
                     : "memory");
        bulk_commit();
      }
    };
    auto rows_read = [&]() {          // the last store has read its rows
      if (lane == 0) bulk_wait_read<0>();
      __syncwarp();
    };
    EdgeIds id = load_ids(P.edge_row, P.col, t0, tiles, row, P.E);
    stage_all(0, id, t0);
    id = load_ids(P.edge_row, P.col, t0 + stride, tiles, row, P.E);
    stage_all(1, id, t0 + stride);
    id = load_ids(P.edge_row, P.col, t0 + 2 * stride, tiles, row, P.E);
    uint32_t n = 0;
    for (int64_t t = t0; t < tiles; t += stride, ++n) {
      const uint32_t st = n & 1, ph = n & 1;
      const int64_t e = t * kTileRows + row;
      const bool valid = e < P.E;
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncwarp();
      const int4 ext = lds_extent(myext + st * (kTileRows * 16));
      const float scale =
          fast::qk_scale_fast(P.scale_mode, P.scale_value, valid ? ext.y - ext.x : 1);
      const int rr = valid ? ext.z : 0;
      const uint32_t src = my + st * kTileBytes, ssp = mysp + st * (kTileRows * 32);
      const uint32_t wsrc = wbase + st * kTileBytes;
      mbar_wait(&r_full[g], ph, 7);
      tc_fence_after();
      uint32_t R[32];
      tmem_ld32(tmem_base + g * kAccCols + lane_off, R);
      // G_e = [dS k_e | dS q_e] (rows past E: zeros)
      float G[32];
      float4 ds4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (valid) ds4 = lds128(ssp);
      const float dsv[4] = {ds4.x, ds4.y, ds4.z, ds4.w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float4 q4 = lds128(src + (uint32_t)((h ^ sw) << 4));
        const float4 k4 = lds128(src + (uint32_t)(((4 + h) ^ sw) << 4));
        const float qs[4] = {q4.x, q4.y, q4.z, q4.w};
        const float ks[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
        for (int d = 0; d < 4; ++d) {
          const float qe = fmaf(qs[d], scale, __uint_as_float(R[4 * h + d]) + bias_s[4 * h + d]);
          const float ke = ks[d] + __uint_as_float(R[16 + 4 * h + d]) + bias_s[16 + 4 * h + d];
          G[4 * h + d] = dsv[h] * ke;
          G[16 + 4 * h + d] = dsv[h] * qe;
        }
      }
      // G leaves through my (now dead) q / k stage row: swizzled staging row + TMA store
      __syncwarp();
      if (P.g_mode) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sts128(wsrc + (uint32_t)lane * 64u + 16u * j,
                 make_float4(G[4 * j], G[4 * j + 1], G[4 * j + 2], G[4 * j + 3]));
          sts128(wsrc + 2048u + (uint32_t)lane * 64u + 16u * j,
                 make_float4(G[16 + 4 * j], G[16 + 4 * j + 1], G[16 + 4 * j + 2], G[16 + 4 * j + 3]));
        }
        store_halves(&tmG, wsrc, t);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(src + (uint32_t)((j ^ sw) << 4),
                 make_float4(G[4 * j], G[4 * j + 1], G[4 * j + 2], G[4 * j + 3]));
        store_rows(&tmG, wsrc, t);
      }
      if (P.da) {
        {
          uint32_t hi[32], lo[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float h = tf32_rna(G[j]);
            hi[j] = __float_as_uint(h);
            lo[j] = __float_as_uint(G[j] - h);   // raw (truncated by the tensor core)
          }
          const uint32_t tg = tmem_base + kGStage0 + 64 * g + lane_off;
          tmem_st32(tg, hi);
          tmem_st32(tg + 32, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(&g_ready[g]);
        // the abar term while MMA 2 runs: sum_h p_e,h dAbar_s,h
        float da[32];
#pragma unroll
        for (int f = 0; f < 32; ++f) da[f] = 0.f;
        if (valid && P.d_abar) {
          const float4 p4 = lds128(ssp + 16);
          const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
          const float4* dab = reinterpret_cast<const float4*>(P.d_abar + (int64_t)rr * P.ld_dab);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
#pragma unroll
            for (int f4 = 0; f4 < 8; ++f4) {
              const float4 x = __ldg(dab + h * 8 + f4);
              da[4 * f4 + 0] = fmaf(pv[h], x.x, da[4 * f4 + 0]);
              da[4 * f4 + 1] = fmaf(pv[h], x.y, da[4 * f4 + 1]);
              da[4 * f4 + 2] = fmaf(pv[h], x.z, da[4 * f4 + 2]);
              da[4 * f4 + 3] = fmaf(pv[h], x.w, da[4 * f4 + 3]);
            }
          }
        }
        mbar_wait(&d_full[g], ph, 8);
        tc_fence_after();
        tmem_ld32(tmem_base + g * kAccCols + lane_off, R);
        tc_fence_before();
        mbar_arrive(&acc_free[g]);
        rows_read();                      // the G store has read the stage rows
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(src + (uint32_t)((j ^ sw) << 4),
                 make_float4(da[4 * j] + __uint_as_float(R[4 * j]),
                             da[4 * j + 1] + __uint_as_float(R[4 * j + 1]),
                             da[4 * j + 2] + __uint_as_float(R[4 * j + 2]),
                             da[4 * j + 3] + __uint_as_float(R[4 * j + 3])));
        if (P.da_reduce) reduce_rows(&tmD, wsrc, t);
        else store_rows(&tmD, wsrc, t);
      } else {
        // no feature gradient wanted: MMA 2 still runs (on zeros) to keep the issue order of
        // the MMA warp
        uint32_t zero[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) zero[j] = 0u;
        const uint32_t tg = tmem_base + kGStage0 + 64 * g + lane_off;
        tmem_st32(tg, zero);
        tmem_st32(tg + 32, zero);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(&g_ready[g]);
        mbar_wait(&d_full[g], ph, 8);
        tc_fence_after();
        tc_fence_before();
        mbar_arrive(&acc_free[g]);
      }
      // the stage rows are free once the last store has read them: stage the tile two turns
      // ahead, fetch the ids after it
      rows_read();
      stage_all(st, id, t + 2 * stride);
      id = load_ids(P.edge_row, P.col, t + 3 * stride, tiles, row, P.E);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (lane == 0) bulk_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(kTmemCols)
                 : "memory");
  }
}

inline bool edge_bwd_launch(const split::EdgeBwdArgs& A, cudaStream_t st, int* rc) {
  CUtensorMap tm, tmG, tmD;
  if (!umma::make_map_rows32(&tm, A.a, A.E, 32, kTileRows)) return false;
  // output maps: one 32-row x 128-byte box per epilogue warp (g_mode: 32 rows x 16 columns of
  // the [E, 128] gradient buffer, unswizzled)
  if (A.g_mode) {
    if (!umma::make_map_box(&tmG, A.G, A.E, 128, 128, 16, 32)) return false;
  } else if (!umma::make_map_rows32(&tmG, A.G, A.E, 32, 32)) {
    return false;
  }
  if (!umma::make_map_rows32(&tmD, A.da ? A.da : A.G, A.E, 32, 32)) return false;
  const int smem = SmemBwd::total + 1024;
  static unsigned long long done = 0;
  ensure_dynamic_smem(k_edge_bwd_umma, smem, &done);
  const int64_t tiles = (A.E + kTileRows - 1) / kTileRows;
  const int64_t sms = device_sm_count();
  k_edge_bwd_umma<<<(unsigned)(tiles < sms ? tiles : sms), kThreads, smem, st>>>(tm, tmG, tmD, A);
  *rc = check_launch("attn_bwd_rows(edge, umma)");
  return true;
}

}  // namespace aumma
}  // namespace spt
