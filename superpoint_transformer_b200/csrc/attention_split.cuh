// attention_split.cuh — the attention core as an EDGE-parallel pass + a ROW-parallel pass.
//
// Replaces, like attention_tile.cuh, the arithmetic of src/nn/attention.py:225-315 of the
// reference for the shipped shape (H=4, D=4, Dv=32, F=32, fp32).  The fused row-tile kernels
// spend ~63 warp instructions per edge because every phase runs in the layout of the worst one
// (profiles/r2_attention_tile_ncu.md); a bare gather of the same v rows runs at 15 TB/s.  Here
// each phase runs in the layout that suits it and the only intermediate is 16 bytes per edge:
//
//   forward
//     edge pass  (thread = edge):  R_e = [Wq;Wk] a_e + b  (32 outputs, dense), q_s, k_t  ->
//                logit2[e, h] = log2(e) * <q_s*scale + Rq_e, k_t + Rk_e>_h        [E, 4]
//                - csrc/attention_umma.cuh: 128 edges per tcgen05 tile (A operand split into
//                  TF32 hi/lo in TMEM), epilogue thread = edge;  k_edge_logits_simple below is
//                  the CUDA-core version (exact fp32; validator and fallback)
//     row pass   (warp = row, lane = 4 value channels): softmax over the row's logits, gather
//                of the v rows, p-weighted sums agg_v / abar, statistics m / z / sump
//   backward
//     row pass   dp_e = <dY_s, v_t> + <dAbar_s, a_e>;  dS = p (dp - sum p dp);  P, dS -> [E, 4];
//                dq_s = sum_e dS (k_t + Rk_e)  through  sum_e dS k_t  and  T_s = sum_e dS a_e
//     edge pass  G_e = [dS (k_t + Rk_e) | dS (q_s*scale + Rq_e)],  da_e = G_e [Wq;Wk] + P dAbar_s
// The statistics (m, z, sump) have the definitions of the fused kernels, so forward and backward
// of the two families can be mixed (tests do).
#pragma once
#include "common.cuh"
#include "attention_fast.cuh"
#include "attention_tile.cuh"

namespace spt {
namespace split {

using fast::f32x2;
using fast::pack2;
using fast::fma2;
using fast::mul2;
using fast::ex2;
using fast::kLog2e;
using fast::kLn2;
using tile::ldg_row16;
using tile::policy_evict_first;
using tile::policy_evict_last;
using tile::fast_rcp;

constexpr int kH = 4, kD = 4, kHD = 16, kF = 32, kDv = 32, kC = 128;
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ inline bool shape_ok(int H, int D, int Dv, int F) {
  return H == kH && D == kD && Dv == kDv && F == kF;
}

// ------------------------------------------------------------------------------------------
// warp reductions of the 4 per-head values every lane holds
// ------------------------------------------------------------------------------------------
// float max through the integer REDUX unit (one instruction): the map below is monotone
__device__ __forceinline__ int f2ord(float x) {
  const int i = __float_as_int(x);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int i) {
  return __int_as_float(i ^ ((i >> 31) & 0x7fffffff));
}
__device__ __forceinline__ float warp_max(float x) {
  return ord2f(__reduce_max_sync(kFull, f2ord(x)));
}
// sum over the warp of 4 values per lane in 6 shuffles: afterwards every lane holds the total
// of head hsel(lane) = 2 * (lane & 1) + ((lane >> 1) & 1); head h sits in lane src_of_head(h)
__device__ __forceinline__ float warp_sum4(float x0, float x1, float x2, float x3, int lane) {
  const bool b0 = lane & 1, b1 = lane & 2;
  float k0 = b0 ? x2 : x0, k1 = b0 ? x3 : x1;
  const float s0 = b0 ? x0 : x2, s1 = b0 ? x1 : x3;
  k0 += __shfl_xor_sync(kFull, s0, 1);
  k1 += __shfl_xor_sync(kFull, s1, 1);
  float kk = b1 ? k1 : k0;
  const float ss = b1 ? k0 : k1;
  kk += __shfl_xor_sync(kFull, ss, 2);
  kk += __shfl_xor_sync(kFull, kk, 4);
  kk += __shfl_xor_sync(kFull, kk, 8);
  kk += __shfl_xor_sync(kFull, kk, 16);
  return kk;
}
__device__ __forceinline__ int hsel_of_lane(int lane) { return 2 * (lane & 1) + ((lane >> 1) & 1); }
__device__ __forceinline__ int src_of_head(int h) { return (h >> 1) | ((h & 1) << 1); }
__device__ __forceinline__ float pick4(const float4& v, int h) {
  return h == 0 ? v.x : h == 1 ? v.y : h == 2 ? v.z : v.w;
}

// streamed 16-byte read of the edge features: L2 evict-first (read once per kernel), but
// allocated in L1 — the 4 head groups of a warp (quarter-warps) read the same 128-byte row, and
// the backward reads it a second time
__device__ __forceinline__ ulonglong2 ldg_stream16(const void* p, uint64_t policy) {
  ulonglong2 v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;"
               : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(policy));
  return v;
}

// ------------------------------------------------------------------------------------------
// forward, edge pass (CUDA cores): thread = edge
// ------------------------------------------------------------------------------------------
struct EdgeFwdArgs {
  const float* q; int ldq;
  const float* k; int ldk;
  const float* a;                       // [E, 32] CSR-ordered edge features
  const int32_t* rowptr; const int32_t* col; const int32_t* edge_row;
  int64_t E;
  const float* Wq; const float* bq; const float* Wk; const float* bk;   // each nullable
  int scale_mode; float scale_value;
  float* logits;                        // [E, ldl]: 4 base-2 logits per edge at logits + e * ldl
  int ldl;                              // 4; 16 when a 16-head problem runs as 4 head groups
};

constexpr int kEdgeThreads = 256;

// W_s[f][o]: o < 16 the q encoder, o >= 16 the k encoder (a missing encoder = zeros)
__device__ __forceinline__ void load_rpe_weights(float (*W_s)[kHD * 2], float* b_s,
                                                 const float* Wq, const float* bq,
                                                 const float* Wk, const float* bk) {
  for (int i = threadIdx.x; i < kF * 2 * kHD; i += blockDim.x) {
    const int f = i >> 5, o = i & 31;
    const float* W = o < kHD ? Wq : Wk;
    W_s[f][o] = W ? W[(o & (kHD - 1)) * kF + f] : 0.f;
  }
  for (int o = threadIdx.x; o < 2 * kHD; o += blockDim.x) {
    const float* W = o < kHD ? Wq : Wk;
    const float* b = o < kHD ? bq : bk;
    b_s[o] = (W && b) ? b[o & (kHD - 1)] : 0.f;
  }
}

__global__ void __launch_bounds__(kEdgeThreads)
k_edge_logits_simple(const EdgeFwdArgs P) {
  __shared__ __align__(16) float W_s[kF][2 * kHD];
  __shared__ __align__(16) float b_s[2 * kHD];
  load_rpe_weights(W_s, b_s, P.Wq, P.bq, P.Wk, P.bk);
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * kEdgeThreads + threadIdx.x;
  if (e >= P.E) return;
  const int row = P.edge_row[e], c = P.col[e];
  float acc[2 * kHD];
#pragma unroll
  for (int o = 0; o < 2 * kHD; ++o) acc[o] = b_s[o];
  const float4* ap = reinterpret_cast<const float4*>(P.a + e * kF);
#pragma unroll
  for (int j = 0; j < kF / 4; ++j) {
    const float4 a4 = __ldg(ap + j);
    const float af[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4* wr = reinterpret_cast<const float4*>(W_s[4 * j + u]);
#pragma unroll
      for (int o4 = 0; o4 < 2 * kHD / 4; ++o4) {
        const float4 w = wr[o4];
        acc[4 * o4 + 0] = fmaf(af[u], w.x, acc[4 * o4 + 0]);
        acc[4 * o4 + 1] = fmaf(af[u], w.y, acc[4 * o4 + 1]);
        acc[4 * o4 + 2] = fmaf(af[u], w.z, acc[4 * o4 + 2]);
        acc[4 * o4 + 3] = fmaf(af[u], w.w, acc[4 * o4 + 3]);
      }
    }
  }
  const float scale =
      fast::qk_scale_fast(P.scale_mode, P.scale_value, P.rowptr[row + 1] - P.rowptr[row]);
  const float4* qp = reinterpret_cast<const float4*>(P.q + (int64_t)row * P.ldq);
  const float4* kp = reinterpret_cast<const float4*>(P.k + (int64_t)c * P.ldk);
  float lg[kH];
#pragma unroll
  for (int h = 0; h < kH; ++h) {
    const float4 q4 = __ldg(qp + h), k4 = __ldg(kp + h);
    float s = (fmaf(q4.x, scale, acc[4 * h + 0])) * (k4.x + acc[kHD + 4 * h + 0]);
    s = fmaf(fmaf(q4.y, scale, acc[4 * h + 1]), k4.y + acc[kHD + 4 * h + 1], s);
    s = fmaf(fmaf(q4.z, scale, acc[4 * h + 2]), k4.z + acc[kHD + 4 * h + 2], s);
    s = fmaf(fmaf(q4.w, scale, acc[4 * h + 3]), k4.w + acc[kHD + 4 * h + 3], s);
    lg[h] = s * kLog2e;
  }
  *reinterpret_cast<float4*>(P.logits + e * P.ldl) = make_float4(lg[0], lg[1], lg[2], lg[3]);
}

// ------------------------------------------------------------------------------------------
// forward, row pass: warp = row, lane = value channels 4*lane.. (head lane>>3) and
// abar[head][4*(lane&7)..]
// ------------------------------------------------------------------------------------------
struct RowFwdArgs {
  const float* logits;                  // [E, 4] base-2 logits of the edge pass
  const void* v; int ldv;               // gathered value rows: fp32, or bf16 (template VBF)
  const float* a;                       // [E, 32]; read only when abar != nullptr
  const int32_t* rowptr; const int32_t* col;
  int64_t num_rows;
  float* agg_v; float* abar; float* sump; float* m; float* z;
  int rows_per_warp;                    // consecutive rows per warp, 1..16
};

constexpr int kRowWarps = 8;
// The row kernels are latency-bound gathers (one L2 round trip of ~1 us per batch of gathered
// rows, a dozen batches per row in the first versions).  What hides it: the edge-feature rows of
// a chunk are staged in shared memory by cp.async as soon as the row's extent is known (no
// registers, overlaps the softmax and the v gathers), so all the registers of the 8-deep batches
// go to the gathered v rows; 24-32 resident warps per SM.
#ifndef SPT_ROW_FWD_CTAS
#define SPT_ROW_FWD_CTAS 3
#endif
#ifndef SPT_ROW_BWD_CTAS
#define SPT_ROW_BWD_CTAS 3
#endif
constexpr int kRowFwdCtas = SPT_ROW_FWD_CTAS;   // 24 warps per SM at <= 85 registers (4 CTAs: spills, 0.148 vs 0.137 ms)

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
// 16-byte asynchronous copy global -> shared, L2 only (streamed data), with an L2 policy
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src),
               "l"(policy)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
// feature rows [tb, tb + n) -> a_s[slot][32] (128-byte rows): lane copies chunk lane & 7 of the
// slots (lane >> 3) + 4 i
__device__ __forceinline__ void stage_features(float* a_s, const float* a, int tb, int n, int lane,
                                               uint64_t policy) {
  const char* src = reinterpret_cast<const char*>(a) + (int64_t)tb * (kF * 4) + 16 * (lane & 7);
  const uint32_t dst = smem_addr(a_s) + 16 * (lane & 7);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int slot = (lane >> 3) + 4 * i;
    if (slot < n) cp_async16(dst + slot * (kF * 4), src + slot * (kF * 4), policy);
  }
  cp_async_commit();
}

// CNT gathered v rows e0.. of the current chunk (p_lane points at p_s[0][head of this lane])
template <int CNT>
__device__ __forceinline__ void row_accumulate_v(int e0, int mycol, const char* vbase,
                                                 unsigned ldvb, uint64_t keep,
                                                 const float* p_lane, f32x2& accv01,
                                                 f32x2& accv23) {
  ulonglong2 vv[CNT];
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    const unsigned tc = (unsigned)__shfl_sync(kFull, mycol, e0 + u);
    vv[u] = ldg_row16(vbase + (uint64_t)tc * (uint64_t)ldvb, keep);
  }
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    const float p = p_lane[(e0 + u) * kH];
    const f32x2 pp = pack2(p, p);
    fma2(accv01, pp, vv[u].x);
    fma2(accv23, pp, vv[u].y);
  }
}

// the same with the staged feature rows consumed in the same pass: one read of p per edge
// (0.127 vs 0.135 ms against a separate abar loop)
// my 4 channels of gathered row `tc`: 16 bytes of fp32, or 8 bytes of bf16 (VBF: the value rows
// stored as bf16, spt_attn_extras.v_bf16 — half the gather bytes, L2-resident)
template <bool VBF>
__device__ __forceinline__ ulonglong2 gather_v4(const char* vbase, unsigned tc, unsigned ldvb,
                                                uint64_t keep) {
  if (VBF) {
    const uint2 w = tile::ldg_row8u(vbase + (uint64_t)tc * (uint64_t)ldvb, keep);
    ulonglong2 v;
    v.x = tile::bf2_to_f32x2(w.x);
    v.y = tile::bf2_to_f32x2(w.y);
    return v;
  }
  return ldg_row16(vbase + (uint64_t)tc * (uint64_t)ldvb, keep);
}

template <int CNT, bool ABAR, bool VBF>
__device__ __forceinline__ void row_accumulate_va(int e0, int mycol, const char* vbase,
                                                  unsigned ldvb, uint64_t keep,
                                                  const float* p_lane, const ulonglong2* a_lane,
                                                  bool wait_a, f32x2& accv01, f32x2& accv23,
                                                  f32x2& acca01, f32x2& acca23) {
  ulonglong2 vv[CNT];
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    const unsigned tc = (unsigned)__shfl_sync(kFull, mycol, e0 + u);
    vv[u] = gather_v4<VBF>(vbase, tc, ldvb, keep);
  }
  if (ABAR && wait_a) {
    cp_async_wait_all();
    __syncwarp();
  }
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    const float p = p_lane[(e0 + u) * kH];
    const f32x2 pp = pack2(p, p);
    if (ABAR) {
      const ulonglong2 a4 = a_lane[(e0 + u) * (kF / 4)];
      fma2(acca01, pp, a4.x);
      fma2(acca23, pp, a4.y);
    }
    fma2(accv01, pp, vv[u].x);
    fma2(accv23, pp, vv[u].y);
  }
}

// one chunk of <= 32 edges of a row: p tile, sum of p, weighted accumulation
template <bool ABAR, bool VBF>
__device__ __forceinline__ void row_fwd_chunk(int n, float4 lg, float4 mx, int mycol, int lane,
                                              float* p_s, const float* a_s, const char* vbase,
                                              unsigned ldvb, float& l, f32x2& accv01,
                                              f32x2& accv23, f32x2& acca01, f32x2& acca23) {
  {
    float4 p;                                     // lanes past n: 2^(-inf) = 0
    p.x = ex2(lg.x - mx.x); p.y = ex2(lg.y - mx.y);
    p.z = ex2(lg.z - mx.z); p.w = ex2(lg.w - mx.w);
    l += warp_sum4(p.x, p.y, p.z, p.w, lane);
    *reinterpret_cast<float4*>(p_s + lane * kH) = p;
    __syncwarp();
  }
  const uint64_t keep = policy_evict_last();
  const float* p_lane = p_s + (lane >> 3);
  int e0 = 0;
  // v rows and staged feature rows consumed together (one read of p per edge); the wait for the
  // staged rows sits behind the first batch of gathers
  const ulonglong2* a_lane = reinterpret_cast<const ulonglong2*>(a_s) + (lane & 7);
  bool first = ABAR;
#pragma unroll 1
  for (; e0 + 8 <= n; e0 += 8) {
    row_accumulate_va<8, ABAR, VBF>(e0, mycol, vbase, ldvb, keep, p_lane, a_lane, first, accv01, accv23,
                               acca01, acca23);
    first = false;
  }
  if (n & 4) {
    row_accumulate_va<4, ABAR, VBF>(e0, mycol, vbase, ldvb, keep, p_lane, a_lane, first, accv01, accv23,
                               acca01, acca23);
    first = false;
    e0 += 4;
  }
  if (n & 2) {
    row_accumulate_va<2, ABAR, VBF>(e0, mycol, vbase, ldvb, keep, p_lane, a_lane, first, accv01, accv23,
                               acca01, acca23);
    first = false;
    e0 += 2;
  }
  if (n & 1)
    row_accumulate_va<1, ABAR, VBF>(e0, mycol, vbase, ldvb, keep, p_lane, a_lane, first, accv01, accv23,
                               acca01, acca23);
}

// per-warp shared memory of the forward row pass
template <bool ABAR>
struct RowFwdTiles {
  float p[32 * kH];              // softmax numerators of the chunk
  float pre_lg[2][32 * kH];      // logits of the first chunk of this / the next row (cp.async)
  int pre_col[2][32];            // its column ids
  float m[kH];                   // row maxima (long rows)
  float a[ABAR ? 32 * kF : 4];   // feature rows of the chunk (cp.async)
};

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_plain(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// A warp walks `rows_per_warp` consecutive rows.  The row extents come in one load, and the
// logits / column ids of the NEXT row's first chunk are copied to shared memory while the
// current row gathers: per row only the batches of gathered v rows remain as dependent round
// trips.
template <bool ABAR, bool VBF>
__global__ void __launch_bounds__(kRowWarps * 32, kRowFwdCtas)
k_row_fwd(const RowFwdArgs P) {
  __shared__ __align__(16) RowFwdTiles<ABAR> tiles[kRowWarps];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = P.rows_per_warp;                  // <= 16
  const int64_t r0 = ((int64_t)blockIdx.x * kRowWarps + w) * R;
  if (r0 >= P.num_rows) return;
  const int nrows = (int)min((int64_t)R, P.num_rows - r0);
  RowFwdTiles<ABAR>& S = tiles[w];
  const int hb = lane >> 3;
  const uint64_t stream = policy_evict_first();
  constexpr unsigned kElt = VBF ? 2u : 4u;
  const char* vbase = reinterpret_cast<const char*>(P.v) + 4 * kElt * lane;
  const unsigned ldvb = (unsigned)P.ldv * kElt;
  const float4* lg4 = reinterpret_cast<const float4*>(P.logits);
  const int rp = P.rowptr[r0 + min(lane, nrows)];             // extents of my rows
  const uint32_t pre_lg = smem_addr(S.pre_lg[0]) + 16 * lane;
  const uint32_t pre_col = smem_addr(S.pre_col[0]) + 4 * lane;
  {
    const int b0 = __shfl_sync(kFull, rp, 0), e0 = __shfl_sync(kFull, rp, 1);
    if (lane < min(32, e0 - b0)) {
      cp_async16_plain(pre_lg, lg4 + b0 + lane);
      cp_async4(pre_col, P.col + b0 + lane);
    }
    cp_async_commit();
  }

#pragma unroll 1
  for (int i = 0; i < nrows; ++i) {
    const int64_t row = r0 + i;
    const int b = __shfl_sync(kFull, rp, i), e = __shfl_sync(kFull, rp, i + 1);
    const int n0 = min(32, e - b);
    cp_async_wait_all();                          // my slot of this row's first chunk has landed
    __syncwarp();                                 // and the previous row's tiles have been read
    float4 lg = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    int mycol = 0;
    if (lane < n0) {
      lg = *reinterpret_cast<const float4*>(&S.pre_lg[i & 1][lane * kH]);
      mycol = S.pre_col[i & 1][lane];
    }
    if (i + 1 < nrows) {                          // the next row's first chunk
      const int e2 = __shfl_sync(kFull, rp, min(i + 2, 31));
      if (lane < min(32, e2 - e)) {
        const uint32_t o = ((i + 1) & 1);
        cp_async16_plain(pre_lg + o * (32 * kH * 4), lg4 + e + lane);
        cp_async4(pre_col + o * (32 * 4), P.col + e + lane);
      }
    }
    if (ABAR && n0 > 0) stage_features(S.a, P.a, b, n0, lane, stream);   // commits
    else cp_async_commit();

    // pass 1: row maxima (later chunks of a long row are read directly)
    const bool single = e - b <= 32;
    float4 mx = lg;
    for (int tb = b + 32; tb < e; tb += 32) {
      const int j = tb + lane;
      if (j < e) {
        const float4 x = __ldg(lg4 + j);
        mx.x = fmaxf(mx.x, x.x); mx.y = fmaxf(mx.y, x.y);
        mx.z = fmaxf(mx.z, x.z); mx.w = fmaxf(mx.w, x.w);
      }
    }
    mx.x = warp_max(mx.x); mx.y = warp_max(mx.y); mx.z = warp_max(mx.z); mx.w = warp_max(mx.w);
    if (lane < kH) P.m[row * kH + lane] = (e > b) ? pick4(mx, lane) * kLn2 : 0.f;   // natural log

    // pass 2: p = 2^(logit - max), sums, weighted accumulation
    f32x2 accv01 = 0ull, accv23 = 0ull, acca01 = 0ull, acca23 = 0ull;
    float l = 0.f;                                // sum of p of head hsel_of_lane(lane)
    if (single) {
      if (e > b)
        row_fwd_chunk<ABAR, VBF>(n0, lg, mx, mycol, lane, S.p, S.a, vbase, ldvb, l, accv01, accv23,
                            acca01, acca23);
    } else {
      // long row: the maxima wait in shared memory while the chunks stream through
      if (lane == 0) *reinterpret_cast<float4*>(S.m) = mx;
#pragma unroll 1
      for (int tb = b; tb < e; tb += 32) {
        if (tb != b) {
          const int j = tb + lane;
          lg = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
          mycol = 0;
          if (j < e) {
            lg = __ldg(lg4 + j);
            mycol = P.col[j];
          }
          __syncwarp();                           // the previous chunk's tiles have been read
          if (ABAR) stage_features(S.a, P.a, tb, min(32, e - tb), lane, stream);
        }
        __syncwarp();
        mx = *reinterpret_cast<const float4*>(S.m);
        row_fwd_chunk<ABAR, VBF>(min(32, e - tb), lg, mx, mycol, lane, S.p, S.a, vbase, ldvb, l,
                            accv01, accv23, acca01, acca23);
      }
    }

    // epilogue: PyG softmax adds 1e-16 to the denominator
    const float zz = l + 1e-16f;
    const float iz = fast_rcp(zz);
    const float inv = __shfl_sync(kFull, iz, src_of_head(hb));
    const f32x2 ii = pack2(inv, inv);
    ulonglong2 o;
    o.x = mul2(accv01, ii); o.y = mul2(accv23, ii);
    *reinterpret_cast<ulonglong2*>(P.agg_v + row * kC + 4 * lane) = o;
    if (ABAR) {
      o.x = mul2(acca01, ii); o.y = mul2(acca23, ii);
      *reinterpret_cast<ulonglong2*>(P.abar + row * (kH * kF) + 4 * lane) = o;
    }
    if (lane < kH) {
      const int h = hsel_of_lane(lane);
      P.z[row * kH + h] = zz;
      P.sump[row * kH + h] = l * iz;
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward, row pass: warp = row, lane = (head hb = lane >> 3, j8 = lane & 7)
//   dp_e,h = <dY_s,h, v_t,h> + <dAbar_s,h, a_e>      (8 edges per transpose-reduce)
//   p = 2^(logit2 - m2) / z ;  dS = p (dp - delta),  delta = <dY, agg> + <dAbar, abar> = sum p dp
//   dq_s = scale * sum_e dS (k_t + Wk a_e + bk) = scale * (U + Wk T + bk sum dS),
//          U = sum_e dS k_t (gathered),  T_h = sum_e dS_e,h a_e
// ------------------------------------------------------------------------------------------
struct RowBwdArgs {
  const float* logits;                  // [E, 4] base-2 logits of the forward edge pass
  const float* k; int ldk;
  const void* v; int ldv;               // gathered value rows: fp32, or bf16 (template VBF)
  const float* a;                       // [E, 32]
  const int32_t* rowptr; const int32_t* col;
  int64_t num_rows;
  const float* Wk; const float* bk;     // k encoder (nullable)
  int scale_mode; float scale_value;
  const float* m; const float* z;
  const float* agg_v; const float* abar; const float* d_agg_v; const float* d_abar;
  float* dq; int lddq;
  float* Pbuf;                          // [E, 4] softmax weights (targets kernel, edge pass)
  float* dS;                            // [E, 4] gradient of the logits (natural units)
  int rows_per_warp;                    // consecutive rows per warp, 1..16
};

__device__ __forceinline__ float hsum2f(f32x2 v) {
  float lo, hi;
  fast::unpack2(v, lo, hi);
  return lo + hi;
}
__device__ __forceinline__ float butterfly8(const float* s, int j8) { return tile::butterfly8(s, j8); }

constexpr int kRowBwdCtas = SPT_ROW_BWD_CTAS;   // 24 warps per SM at <= 85 registers

// per-warp shared memory of the backward row pass
struct RowBwdTiles {
  float a[32 * kF];        // feature rows of the chunk            (cp.async)
  float k[32 * kHD];       // gathered k rows of the chunk         (cp.async)
  float ds[32 * kH];       // dS of the chunk's (edge, head) pairs
};

// partial dot products <dY, v_t> (+ <dAbar, a_e>) of CNT edges e0.. : CNT gathered rows in flight
template <int CNT, bool HAS_DAB, bool VBF>
__device__ __forceinline__ void row_bwd_partials(float* s, int e0, int mycol, const char* vbase,
                                                 unsigned ldvb, uint64_t keep,
                                                 const ulonglong2* a_lane, f32x2 dy01, f32x2 dy23,
                                                 f32x2 dab01, f32x2 dab23) {
  ulonglong2 vv[CNT];
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    const unsigned tc = (unsigned)__shfl_sync(kFull, mycol, e0 + u);
    vv[u] = gather_v4<VBF>(vbase, tc, ldvb, keep);
  }
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    f32x2 acc = mul2(dy01, vv[u].x);
    fma2(acc, dy23, vv[u].y);
    if (HAS_DAB) {
      const ulonglong2 a4 = a_lane[(e0 + u) * (kF / 4)];
      fma2(acc, dab01, a4.x);
      fma2(acc, dab23, a4.y);
    }
    s[u] = hsum2f(acc);
  }
}

template <bool HAS_DAB, bool HAS_WK, bool VBF>
__global__ void __launch_bounds__(kRowWarps * 32, kRowBwdCtas)
k_row_bwd(const RowBwdArgs P) {
  extern __shared__ __align__(16) unsigned char row_bwd_smem[];
  RowBwdTiles* tiles = reinterpret_cast<RowBwdTiles*>(row_bwd_smem);
  float (*Wk_s)[kF] = reinterpret_cast<float (*)[kF]>(row_bwd_smem + kRowWarps * sizeof(RowBwdTiles));
  float* bk_s = &Wk_s[kHD][0];
  if (HAS_WK) {
    for (int i = threadIdx.x; i < kHD * kF; i += blockDim.x) Wk_s[i >> 5][i & 31] = P.Wk[i];
    if (threadIdx.x < kHD) bk_s[threadIdx.x] = P.bk ? P.bk[threadIdx.x] : 0.f;
    __syncthreads();
  }
  constexpr bool NEED_A = HAS_DAB || HAS_WK;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = P.rows_per_warp;
  const int64_t r0 = ((int64_t)blockIdx.x * kRowWarps + w) * R;
  if (r0 >= P.num_rows) return;
  const int nrows = (int)min((int64_t)R, P.num_rows - r0);
  RowBwdTiles& S = tiles[w];
  const int hb = lane >> 3, j8 = lane & 7;
  const uint64_t keep = policy_evict_last(), stream = policy_evict_first();
  constexpr unsigned kElt = VBF ? 2u : 4u;
  const char* vbase = reinterpret_cast<const char*>(P.v) + 4 * kElt * lane;
  const unsigned ldvb = (unsigned)P.ldv * kElt, ldkb = (unsigned)P.ldk * 4u;
  const ulonglong2* a_lane = reinterpret_cast<const ulonglong2*>(S.a) + j8;
  const int rp = P.rowptr[r0 + min(lane, nrows)];             // extents of my rows

#pragma unroll 1
  for (int ri = 0; ri < nrows; ++ri) {
  const int64_t row = r0 + ri;
  const int b = __shfl_sync(kFull, rp, ri), e = __shfl_sync(kFull, rp, ri + 1);
  __syncwarp();                                    // the previous row's tiles have been read
  if (NEED_A && e > b) stage_features(S.a, P.a, b, min(32, e - b), lane, stream);

  const float scale = fast::qk_scale_fast(P.scale_mode, P.scale_value, e - b);
  const float m2 = P.m[row * kH + hb] * kLog2e;
  const float zi = fast_rcp(P.z[row * kH + hb]);
  const float4 dy = *reinterpret_cast<const float4*>(P.d_agg_v + row * kC + 4 * lane);
  float4 dab = make_float4(0.f, 0.f, 0.f, 0.f);
  float delta;
  {
    const float4 ag = *reinterpret_cast<const float4*>(P.agg_v + row * kC + 4 * lane);
    float part = dy.x * ag.x + dy.y * ag.y + dy.z * ag.z + dy.w * ag.w;
    if (HAS_DAB) {
      dab = *reinterpret_cast<const float4*>(P.d_abar + row * (kH * kF) + 4 * lane);
      const float4 ab = *reinterpret_cast<const float4*>(P.abar + row * (kH * kF) + 4 * lane);
      part += dab.x * ab.x + dab.y * ab.y + dab.z * ab.z + dab.w * ab.w;
    }
    part += __shfl_xor_sync(kFull, part, 1);
    part += __shfl_xor_sync(kFull, part, 2);
    part += __shfl_xor_sync(kFull, part, 4);
    delta = part;
  }
  const f32x2 dy01 = pack2(dy.x, dy.y), dy23 = pack2(dy.z, dy.w);
  const f32x2 dab01 = pack2(dab.x, dab.y), dab23 = pack2(dab.z, dab.w);

  f32x2 T01 = 0ull, T23 = 0ull;
  float U = 0.f, SdS = 0.f;
#pragma unroll 1
  for (int tb = b; tb < e; tb += 32) {
    const int n = min(32, e - tb);
    const int mycol = (lane < n) ? P.col[tb + lane] : 0;
    if (tb != b) {
      __syncwarp();                                // the previous chunk's tiles have been read
      if (NEED_A) stage_features(S.a, P.a, tb, n, lane, stream);
    }
    {
      // gathered k rows -> S.k[slot][16]: lane copies chunk lane & 3 of the slots (lane >> 2) + 8 i
      const uint32_t dst = smem_addr(S.k) + 16 * (lane & 3);
      const char* src = reinterpret_cast<const char*>(P.k) + 16 * (lane & 3);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int slot = (lane >> 2) + 8 * i;
        const unsigned tc = (unsigned)__shfl_sync(kFull, mycol, slot);
        if (slot < n) cp_async16(dst + slot * (kHD * 4), src + (uint64_t)tc * (uint64_t)ldkb, keep);
      }
      cp_async_commit();
    }
    if (HAS_DAB) {      // the dot products below read the feature rows (the k rows may still fly)
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncwarp();
    }
#pragma unroll 1
    for (int e0 = 0; e0 < n; e0 += 8) {
      const int cnt = min(8, n - e0);
      const bool valid = j8 < cnt;
      const int64_t slot = (int64_t)(tb + e0 + j8) * kH + hb;
      const float lg = valid ? __ldg(P.logits + slot) : -INFINITY;
      float s[8];
      if (cnt == 8) {
        row_bwd_partials<8, HAS_DAB, VBF>(s, e0, mycol, vbase, ldvb, keep, a_lane, dy01, dy23, dab01,
                                     dab23);
      } else {
        // short last group: 4 + 2 + 1 rows, no wasted gathers
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] = 0.f;
        int o = 0;
        if (cnt & 4) {
          float s4[4];
          row_bwd_partials<4, HAS_DAB, VBF>(s4, e0, mycol, vbase, ldvb, keep, a_lane, dy01, dy23,
                                       dab01, dab23);
          s[0] = s4[0]; s[1] = s4[1]; s[2] = s4[2]; s[3] = s4[3];
          o = 4;
        }
        if (cnt & 2) {
          float s2[2];
          row_bwd_partials<2, HAS_DAB, VBF>(s2, e0 + o, mycol, vbase, ldvb, keep, a_lane, dy01, dy23,
                                       dab01, dab23);
          if (o) { s[4] = s2[0]; s[5] = s2[1]; } else { s[0] = s2[0]; s[1] = s2[1]; }
          o += 2;
        }
        if (cnt & 1) {
          float s1[1];
          row_bwd_partials<1, HAS_DAB, VBF>(s1, e0 + o, mycol, vbase, ldvb, keep, a_lane, dy01, dy23,
                                       dab01, dab23);
          if (o == 0) s[0] = s1[0]; else if (o == 2) s[2] = s1[0];
          else if (o == 4) s[4] = s1[0]; else s[6] = s1[0];
        }
      }
      const float dp = butterfly8(s, j8);                 // edge e0 + j8, head hb
      const float p = ex2(lg - m2) * zi;                   // 0 past the end of the row
      const float ds = p * (dp - delta);
      if (valid) {
        P.Pbuf[slot] = p;
        P.dS[slot] = ds;
        S.ds[(e0 + j8) * kH + hb] = ds;
      }
      SdS += ds;
    }
    // T += dS a, U += dS k from the staged tiles
    cp_async_wait_all();
    __syncwarp();
    if (HAS_WK) {
      const float* ds_lane = S.ds + hb;
#pragma unroll 4
      for (int j = 0; j < n; ++j) {
        const float wj = ds_lane[j * kH];
        const f32x2 ww = pack2(wj, wj);
        const ulonglong2 a4 = a_lane[j * (kF / 4)];
        fma2(T01, ww, a4.x);
        fma2(T23, ww, a4.y);
      }
    }
    {
      // my k element: head hb, dim j8 & 3, edges of parity j8 >> 2
      const float* kp = S.k + hb * kD + (j8 & 3);
#pragma unroll 4
      for (int j = j8 >> 2; j < n; j += 2) U = fmaf(S.ds[j * kH + hb], kp[j * kHD], U);
    }
  }

  // dq of the row
  SdS += __shfl_xor_sync(kFull, SdS, 1);
  SdS += __shfl_xor_sync(kFull, SdS, 2);
  SdS += __shfl_xor_sync(kFull, SdS, 4);
  U += __shfl_xor_sync(kFull, U, 4);                       // lanes j8 and j8 ^ 4: dim j8 & 3
  if (HAS_WK) {
    float t4[4];
    fast::unpack2(T01, t4[0], t4[1]);
    fast::unpack2(T23, t4[2], t4[3]);
    float pd[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const float4 wv = *reinterpret_cast<const float4*>(&Wk_s[hb * kD + d][4 * j8]);
      pd[d] = wv.x * t4[0] + wv.y * t4[1] + wv.z * t4[2] + wv.w * t4[3];
    }
    // 4 partial sums over the 8 lanes of the head: lane ends with dim 2 * bit2 + bit1
    const bool b4 = j8 & 4, b2 = j8 & 2;
    float r0s = (b4 ? pd[2] : pd[0]) + __shfl_xor_sync(kFull, b4 ? pd[0] : pd[2], 4);
    float r1s = (b4 ? pd[3] : pd[1]) + __shfl_xor_sync(kFull, b4 ? pd[1] : pd[3], 4);
    float r = (b2 ? r1s : r0s) + __shfl_xor_sync(kFull, b2 ? r0s : r1s, 2);
    r += __shfl_xor_sync(kFull, r, 1);
    const int d = (b4 ? 2 : 0) + (b2 ? 1 : 0);
    const float Ud = __shfl_sync(kFull, U, (lane & 24) | d);
    if ((j8 & 1) == 0)
      P.dq[row * P.lddq + hb * kD + d] = scale * (Ud + r + bk_s[hb * kD + d] * SdS);
  } else {
    if (j8 < 4) P.dq[row * P.lddq + hb * kD + j8] = scale * U;
  }
  }   // rows of the warp
}

constexpr int kRowBwdSmem = kRowWarps * (int)sizeof(RowBwdTiles) + (kHD * kF + kHD) * 4;

// ------------------------------------------------------------------------------------------
// backward, edge pass (CUDA cores): thread = edge.  G_e = [dS (k_t + Rk) | dS (q_s scale + Rq)],
// da_e = G_e [Wq;Wk] + sum_h p_e,h dAbar_s,h
// ------------------------------------------------------------------------------------------
struct EdgeBwdArgs {
  const float* q; int ldq;
  const float* k; int ldk;
  const float* a;
  const int32_t* rowptr; const int32_t* col; const int32_t* edge_row;
  int64_t E;
  const float* Wq; const float* bq; const float* Wk; const float* bk;
  int scale_mode; float scale_value;
  const float* dS; const float* Pbuf; const float* d_abar;   // d_abar nullable
  float* G;                             // [E, 32]
  float* da;                            // [E, 32] nullable
  // tcgen05 pass only — a 16-head problem run as 4 launches over head groups (the simple pass
  // and the 4-head callers use the defaults 4 / 128 / 0):
  int ld_ds;                            // floats between the dS / P rows of consecutive edges
  int ld_dab;                           // floats between the dAbar rows of consecutive nodes
  int g_mode;                           // 1: G is [E, 128]: dq_e half at column g_col_q, dk_e at g_col_k
  int g_col_q, g_col_k;
  int da_reduce;                        // 1: da += (TMA reduce-add store) instead of da =
};

__global__ void __launch_bounds__(kEdgeThreads)
k_edge_bwd_simple(const EdgeBwdArgs P) {
  __shared__ __align__(16) float W_s[kF][2 * kHD];      // [f][o]
  __shared__ __align__(16) float Wn_s[2 * kHD][kF];     // [o][f]
  __shared__ __align__(16) float b_s[2 * kHD];
  load_rpe_weights(W_s, b_s, P.Wq, P.bq, P.Wk, P.bk);
  for (int i = threadIdx.x; i < 2 * kHD * kF; i += blockDim.x) {
    const int o = i >> 5, f = i & 31;
    const float* W = o < kHD ? P.Wq : P.Wk;
    Wn_s[o][f] = W ? W[(o & (kHD - 1)) * kF + f] : 0.f;
  }
  __syncthreads();
  const int64_t e = (int64_t)blockIdx.x * kEdgeThreads + threadIdx.x;
  if (e >= P.E) return;
  const int row = P.edge_row[e], c = P.col[e];
  float acc[2 * kHD];
#pragma unroll
  for (int o = 0; o < 2 * kHD; ++o) acc[o] = b_s[o];
  const float4* ap = reinterpret_cast<const float4*>(P.a + e * kF);
#pragma unroll
  for (int j = 0; j < kF / 4; ++j) {
    const float4 a4 = __ldg(ap + j);
    const float af[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4* wr = reinterpret_cast<const float4*>(W_s[4 * j + u]);
#pragma unroll
      for (int o4 = 0; o4 < 2 * kHD / 4; ++o4) {
        const float4 w = wr[o4];
        acc[4 * o4 + 0] = fmaf(af[u], w.x, acc[4 * o4 + 0]);
        acc[4 * o4 + 1] = fmaf(af[u], w.y, acc[4 * o4 + 1]);
        acc[4 * o4 + 2] = fmaf(af[u], w.z, acc[4 * o4 + 2]);
        acc[4 * o4 + 3] = fmaf(af[u], w.w, acc[4 * o4 + 3]);
      }
    }
  }
  const float scale =
      fast::qk_scale_fast(P.scale_mode, P.scale_value, P.rowptr[row + 1] - P.rowptr[row]);
  const float4* qp = reinterpret_cast<const float4*>(P.q + (int64_t)row * P.ldq);
  const float4* kp = reinterpret_cast<const float4*>(P.k + (int64_t)c * P.ldk);
  const float4 ds4 = __ldg(reinterpret_cast<const float4*>(P.dS) + e);
  const float dsv[4] = {ds4.x, ds4.y, ds4.z, ds4.w};
  float G[2 * kHD];
#pragma unroll
  for (int h = 0; h < kH; ++h) {
    const float4 q4 = __ldg(qp + h), k4 = __ldg(kp + h);
    const float qv[4] = {q4.x, q4.y, q4.z, q4.w}, kv[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
    for (int d = 0; d < kD; ++d) {
      const float qe = fmaf(qv[d], scale, acc[4 * h + d]);
      const float ke = kv[d] + acc[kHD + 4 * h + d];
      G[4 * h + d] = dsv[h] * ke;
      G[kHD + 4 * h + d] = dsv[h] * qe;
    }
  }
  float4* gp = reinterpret_cast<float4*>(P.G + e * (2 * kHD));
#pragma unroll
  for (int j = 0; j < 2 * kHD / 4; ++j)
    gp[j] = make_float4(G[4 * j], G[4 * j + 1], G[4 * j + 2], G[4 * j + 3]);
  if (!P.da) return;
  float da[kF];
#pragma unroll
  for (int f = 0; f < kF; ++f) da[f] = 0.f;
#pragma unroll
  for (int o = 0; o < 2 * kHD; ++o) {
    const float4* wr = reinterpret_cast<const float4*>(Wn_s[o]);
#pragma unroll
    for (int f4 = 0; f4 < kF / 4; ++f4) {
      const float4 w = wr[f4];
      da[4 * f4 + 0] = fmaf(G[o], w.x, da[4 * f4 + 0]);
      da[4 * f4 + 1] = fmaf(G[o], w.y, da[4 * f4 + 1]);
      da[4 * f4 + 2] = fmaf(G[o], w.z, da[4 * f4 + 2]);
      da[4 * f4 + 3] = fmaf(G[o], w.w, da[4 * f4 + 3]);
    }
  }
  if (P.d_abar) {
    const float4 p4 = __ldg(reinterpret_cast<const float4*>(P.Pbuf) + e);
    const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
    const float4* dab = reinterpret_cast<const float4*>(P.d_abar + (int64_t)row * (kH * kF));
#pragma unroll
    for (int h = 0; h < kH; ++h) {
#pragma unroll
      for (int f4 = 0; f4 < kF / 4; ++f4) {
        const float4 x = __ldg(dab + h * (kF / 4) + f4);
        da[4 * f4 + 0] = fmaf(pv[h], x.x, da[4 * f4 + 0]);
        da[4 * f4 + 1] = fmaf(pv[h], x.y, da[4 * f4 + 1]);
        da[4 * f4 + 2] = fmaf(pv[h], x.z, da[4 * f4 + 2]);
        da[4 * f4 + 3] = fmaf(pv[h], x.w, da[4 * f4 + 3]);
      }
    }
  }
  float4* dp = reinterpret_cast<float4*>(P.da + e * kF);
#pragma unroll
  for (int j = 0; j < kF / 4; ++j)
    dp[j] = make_float4(da[4 * j], da[4 * j + 1], da[4 * j + 2], da[4 * j + 3]);
}

}  // namespace split
}  // namespace spt
