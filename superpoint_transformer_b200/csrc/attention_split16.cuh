// attention_split16.cuh — the split attention kernels for the SHIPPED head layout: H = 16 heads,
// qk_dim D = 4, F = 32, C = 64 (S3DIS / DALES: Dv = 4) or C = 128 (KITTI-360: Dv = 8)
// (configs/model/semantic/_down.yaml:12, _attention.yaml:15,
// configs/experiment/semantic/kitti360.yaml:22-27).  Same decomposition as attention_split.cuh
// (edge-parallel pass, row-parallel pass, 4 bytes per edge and head in between); the generic
// kernels spend ~400 / ~1400 warp instructions per edge on these shapes (5.0 / 17.7 ms at
// E = 1.7 M).  Edge passes: thread = edge on the CUDA cores, 4 heads at a time (exact fp32).
// Row passes: lane = (head h = lane / 2, half = lane & 1): both halves of a head sit in
// adjacent lanes for either C, so every per-head reduction is ONE shfl.xor 1.
#pragma once
#include "attention_split.cuh"

namespace spt {
namespace split16 {

using split::kFull;
using split::warp_max;
using split::cp_async16;
using split::cp_async_commit;
using split::cp_async_wait_all;
using split::stage_features;
using split::smem_addr;
using fast::ex2;
using fast::kLog2e;
using fast::kLn2;
using tile::fast_rcp;
using tile::policy_evict_first;
using tile::policy_evict_last;

constexpr int kH = 16, kD = 4, kHD = 64, kF = 32;
constexpr int kGroups = 4;             // head groups of 4 in the edge passes
constexpr int kEdgeThreads = 256;
constexpr int kRowWarps = 8;

__host__ __device__ inline bool shape_ok(int H, int D, int Dv, int F) {
  return H == kH && D == kD && F == kF && (Dv == 4 || Dv == 8);
}

// ------------------------------------------------------------------------------------------
// edge passes
// ------------------------------------------------------------------------------------------
struct EdgeFwdArgs {
  const float* q; int ldq;
  const float* k; int ldk;
  const float* a;
  const int32_t* rowptr; const int32_t* col; const int32_t* edge_row;
  int64_t E;
  const float* Wq; const float* bq; const float* Wk; const float* bk;
  int scale_mode; float scale_value;
  float* logits;                        // [E, 16] base-2
};

// W_s[g][f][o]: o < 16 the q encoder outputs of heads 4g..4g+3, o >= 16 the k encoder ones
__device__ __forceinline__ void load_weights16(float (*W_s)[kF][32], float (*b_s)[32],
                                               const float* Wq, const float* bq,
                                               const float* Wk, const float* bk) {
  for (int i = threadIdx.x; i < kGroups * kF * 32; i += blockDim.x) {
    const int g = i / (kF * 32), f = (i / 32) % kF, o = i & 31;
    const float* W = o < 16 ? Wq : Wk;
    W_s[g][f][o] = W ? W[(16 * g + (o & 15)) * kF + f] : 0.f;
  }
  for (int i = threadIdx.x; i < kGroups * 32; i += blockDim.x) {
    const int g = i >> 5, o = i & 31;
    const float* W = o < 16 ? Wq : Wk;
    const float* b = o < 16 ? bq : bk;
    b_s[g][o] = (W && b) ? b[16 * g + (o & 15)] : 0.f;
  }
}

// R[o] = b + sum_f a[f] W[f][o] for the 32 outputs of one head group.  Packed FMAs
// (fma.rn.f32x2: two outputs per instruction) — the pass is bound by FMA issue.
__device__ __forceinline__ void rpe_group(float (&acc)[32], const float (&af)[kF],
                                          const float (*Wg)[32], const float* bg) {
  fast::f32x2 acc2[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) acc2[o] = fast::pack2(bg[2 * o], bg[2 * o + 1]);
#pragma unroll
  for (int f = 0; f < kF; ++f) {
    const fast::f32x2 a2 = fast::pack2(af[f], af[f]);
    const ulonglong2* wr = reinterpret_cast<const ulonglong2*>(Wg[f]);
#pragma unroll
    for (int o4 = 0; o4 < 8; ++o4) {
      const ulonglong2 w = wr[o4];
      fast::fma2(acc2[2 * o4], a2, w.x);
      fast::fma2(acc2[2 * o4 + 1], a2, w.y);
    }
  }
#pragma unroll
  for (int o = 0; o < 16; ++o) fast::unpack2(acc2[o], acc[2 * o], acc[2 * o + 1]);
}

// R of TWO edges per thread: every 16-byte weight read from shared memory feeds 8 packed FMAs
// instead of 4 (the one-edge version is bound by the LDS pipe: 1 LDS.128 per 2 FFMA2)
__device__ __forceinline__ void rpe_group2(float (&acc0)[32], float (&acc1)[32],
                                           const float (&af0)[kF], const float (&af1)[kF],
                                           const float (*Wg)[32], const float* bg) {
  fast::f32x2 a0[16], a1[16];
#pragma unroll
  for (int o = 0; o < 16; ++o) a0[o] = a1[o] = fast::pack2(bg[2 * o], bg[2 * o + 1]);
#pragma unroll
  for (int f = 0; f < kF; ++f) {
    const fast::f32x2 x0 = fast::pack2(af0[f], af0[f]), x1 = fast::pack2(af1[f], af1[f]);
    const ulonglong2* wr = reinterpret_cast<const ulonglong2*>(Wg[f]);
#pragma unroll
    for (int o4 = 0; o4 < 8; ++o4) {
      const ulonglong2 w = wr[o4];
      fast::fma2(a0[2 * o4], x0, w.x);
      fast::fma2(a0[2 * o4 + 1], x0, w.y);
      fast::fma2(a1[2 * o4], x1, w.x);
      fast::fma2(a1[2 * o4 + 1], x1, w.y);
    }
  }
#pragma unroll
  for (int o = 0; o < 16; ++o) {
    fast::unpack2(a0[o], acc0[2 * o], acc0[2 * o + 1]);
    fast::unpack2(a1[o], acc1[2 * o], acc1[2 * o + 1]);
  }
}

__device__ __forceinline__ void load_row32(float (&af)[kF], const float* p) {
  const float4* ap = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int j = 0; j < kF / 4; ++j) {
    const float4 t = __ldg(ap + j);
    af[4 * j] = t.x; af[4 * j + 1] = t.y; af[4 * j + 2] = t.z; af[4 * j + 3] = t.w;
  }
}

// logits of 4 heads of one edge from its R (q part acc[0..15], k part acc[16..31])
__device__ __forceinline__ float4 group_logits(const float (&acc)[32], const float4* qp,
                                               const float4* kp, float scale) {
  float lg[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float4 q4 = __ldg(qp + h), k4 = __ldg(kp + h);
    float s = fmaf(q4.x, scale, acc[4 * h + 0]) * (k4.x + acc[16 + 4 * h + 0]);
    s = fmaf(fmaf(q4.y, scale, acc[4 * h + 1]), k4.y + acc[16 + 4 * h + 1], s);
    s = fmaf(fmaf(q4.z, scale, acc[4 * h + 2]), k4.z + acc[16 + 4 * h + 2], s);
    s = fmaf(fmaf(q4.w, scale, acc[4 * h + 3]), k4.w + acc[16 + 4 * h + 3], s);
    lg[h] = s * kLog2e;
  }
  return make_float4(lg[0], lg[1], lg[2], lg[3]);
}

// thread = edges 2t and 2t + 1 (the second one clamped to the last edge; its store is skipped)
__global__ void __launch_bounds__(kEdgeThreads)
k_edge_logits16(const EdgeFwdArgs P) {
  __shared__ __align__(16) float W_s[kGroups][kF][32];
  __shared__ __align__(16) float b_s[kGroups][32];
  load_weights16(W_s, b_s, P.Wq, P.bq, P.Wk, P.bk);
  __syncthreads();
  const int64_t e0 = 2 * ((int64_t)blockIdx.x * kEdgeThreads + threadIdx.x);
  if (e0 >= P.E) return;
  const bool two = e0 + 1 < P.E;
  const int64_t e1 = two ? e0 + 1 : e0;
  const int row0 = P.edge_row[e0], c0 = P.col[e0], row1 = P.edge_row[e1], c1 = P.col[e1];
  float af0[kF], af1[kF];
  load_row32(af0, P.a + e0 * kF);
  load_row32(af1, P.a + e1 * kF);
  const float sc0 =
      fast::qk_scale_fast(P.scale_mode, P.scale_value, P.rowptr[row0 + 1] - P.rowptr[row0]);
  const float sc1 =
      fast::qk_scale_fast(P.scale_mode, P.scale_value, P.rowptr[row1 + 1] - P.rowptr[row1]);
  const float4* qp0 = reinterpret_cast<const float4*>(P.q + (int64_t)row0 * P.ldq);
  const float4* kp0 = reinterpret_cast<const float4*>(P.k + (int64_t)c0 * P.ldk);
  const float4* qp1 = reinterpret_cast<const float4*>(P.q + (int64_t)row1 * P.ldq);
  const float4* kp1 = reinterpret_cast<const float4*>(P.k + (int64_t)c1 * P.ldk);
#pragma unroll 1
  for (int g = 0; g < kGroups; ++g) {
    float acc0[32], acc1[32];
    rpe_group2(acc0, acc1, af0, af1, W_s[g], b_s[g]);
    *reinterpret_cast<float4*>(P.logits + e0 * kH + 4 * g) =
        group_logits(acc0, qp0 + 4 * g, kp0 + 4 * g, sc0);
    if (two)
      *reinterpret_cast<float4*>(P.logits + e1 * kH + 4 * g) =
          group_logits(acc1, qp1 + 4 * g, kp1 + 4 * g, sc1);
  }
}

struct EdgeBwdArgs {
  const float* q; int ldq;
  const float* k; int ldk;
  const float* a;
  const int32_t* rowptr; const int32_t* col; const int32_t* edge_row;
  int64_t E;
  const float* Wq; const float* bq; const float* Wk; const float* bk;
  int scale_mode; float scale_value;
  const float* dS; const float* Pbuf; const float* d_abar;   // [E,16], [E,16], [R,16,32] nullable
  float* G;                             // [E, 128] = [dq_e (64) | dk_e (64)]
  float* da;                            // [E, 32] nullable
};

// G of 4 heads of one edge from its R, dS, q, k; stored to G [E, 128] = [dq_e (64) | dk_e (64)]
__device__ __forceinline__ void group_G(float (&G)[32], const float (&acc)[32], const float4* qp,
                                        const float4* kp, float scale, const float4 ds4,
                                        float* Gq, float* Gk, bool store) {
  const float dsv[4] = {ds4.x, ds4.y, ds4.z, ds4.w};
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float4 q4 = __ldg(qp + h), k4 = __ldg(kp + h);
    const float qv[4] = {q4.x, q4.y, q4.z, q4.w}, kv[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
    for (int d = 0; d < kD; ++d) {
      const float qe = fmaf(qv[d], scale, acc[4 * h + d]);
      const float ke = kv[d] + acc[16 + 4 * h + d];
      G[4 * h + d] = dsv[h] * ke;
      G[16 + 4 * h + d] = dsv[h] * qe;
    }
  }
  if (store) {
    float4* gq = reinterpret_cast<float4*>(Gq);
    float4* gk = reinterpret_cast<float4*>(Gk);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      gq[j] = make_float4(G[4 * j], G[4 * j + 1], G[4 * j + 2], G[4 * j + 3]);
      gk[j] = make_float4(G[16 + 4 * j], G[16 + 4 * j + 1], G[16 + 4 * j + 2], G[16 + 4 * j + 3]);
    }
  }
}

// da += sum_h p_h dAbar[row][h][:] (16 heads), then store
__device__ __forceinline__ void finish_da(float (&da)[kF], const float* Pbuf_e,
                                          const float* dab_row, float* out) {
  if (dab_row) {
    const float4* pp = reinterpret_cast<const float4*>(Pbuf_e);
    const float4* dab = reinterpret_cast<const float4*>(dab_row);
#pragma unroll 1
    for (int g = 0; g < kGroups; ++g) {
      const float4 p4 = __ldg(pp + g);
      const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
#pragma unroll
        for (int f4 = 0; f4 < kF / 4; ++f4) {
          const float4 x = __ldg(dab + (4 * g + h) * (kF / 4) + f4);
          da[4 * f4 + 0] = fmaf(pv[h], x.x, da[4 * f4 + 0]);
          da[4 * f4 + 1] = fmaf(pv[h], x.y, da[4 * f4 + 1]);
          da[4 * f4 + 2] = fmaf(pv[h], x.z, da[4 * f4 + 2]);
          da[4 * f4 + 3] = fmaf(pv[h], x.w, da[4 * f4 + 3]);
        }
      }
    }
  }
  float4* dp = reinterpret_cast<float4*>(out);
#pragma unroll
  for (int j = 0; j < kF / 4; ++j)
    dp[j] = make_float4(da[4 * j], da[4 * j + 1], da[4 * j + 2], da[4 * j + 3]);
}

// thread = edges 2t and 2t + 1: every weight read feeds both (see rpe_group2)
__global__ void __launch_bounds__(kEdgeThreads)
k_edge_bwd16(const EdgeBwdArgs P) {
  extern __shared__ __align__(16) float edge16_smem[];
  float (*W_s)[kF][32] = reinterpret_cast<float (*)[kF][32]>(edge16_smem);            // [g][f][o]
  float (*Wn_s)[32][kF] = reinterpret_cast<float (*)[32][kF]>(edge16_smem + kGroups * kF * 32);  // [g][o][f]
  float (*b_s)[32] = reinterpret_cast<float (*)[32]>(edge16_smem + 2 * kGroups * kF * 32);
  load_weights16(W_s, b_s, P.Wq, P.bq, P.Wk, P.bk);
  for (int i = threadIdx.x; i < kGroups * 32 * kF; i += blockDim.x) {
    const int g = i / (32 * kF), o = (i / kF) % 32, f = i % kF;
    const float* W = o < 16 ? P.Wq : P.Wk;
    Wn_s[g][o][f] = W ? W[(16 * g + (o & 15)) * kF + f] : 0.f;
  }
  __syncthreads();
  const int64_t e0 = 2 * ((int64_t)blockIdx.x * kEdgeThreads + threadIdx.x);
  if (e0 >= P.E) return;
  const bool two = e0 + 1 < P.E;
  const int64_t e1 = two ? e0 + 1 : e0;
  const int row0 = P.edge_row[e0], c0 = P.col[e0], row1 = P.edge_row[e1], c1 = P.col[e1];
  float af0[kF], af1[kF];
  load_row32(af0, P.a + e0 * kF);
  load_row32(af1, P.a + e1 * kF);
  const float sc0 =
      fast::qk_scale_fast(P.scale_mode, P.scale_value, P.rowptr[row0 + 1] - P.rowptr[row0]);
  const float sc1 =
      fast::qk_scale_fast(P.scale_mode, P.scale_value, P.rowptr[row1 + 1] - P.rowptr[row1]);
  const float4* qp0 = reinterpret_cast<const float4*>(P.q + (int64_t)row0 * P.ldq);
  const float4* kp0 = reinterpret_cast<const float4*>(P.k + (int64_t)c0 * P.ldk);
  const float4* qp1 = reinterpret_cast<const float4*>(P.q + (int64_t)row1 * P.ldq);
  const float4* kp1 = reinterpret_cast<const float4*>(P.k + (int64_t)c1 * P.ldk);
  const float4* ds0 = reinterpret_cast<const float4*>(P.dS + e0 * kH);
  const float4* ds1 = reinterpret_cast<const float4*>(P.dS + e1 * kH);
  fast::f32x2 da0[kF / 2], da1[kF / 2];
#pragma unroll
  for (int f = 0; f < kF / 2; ++f) da0[f] = da1[f] = 0ull;
#pragma unroll 1
  for (int g = 0; g < kGroups; ++g) {
    float acc0[32], acc1[32];
    rpe_group2(acc0, acc1, af0, af1, W_s[g], b_s[g]);
    float G0[32], G1[32];
    group_G(G0, acc0, qp0 + 4 * g, kp0 + 4 * g, sc0, __ldg(ds0 + g),
            P.G + e0 * (2 * kHD) + 16 * g, P.G + e0 * (2 * kHD) + kHD + 16 * g, true);
    group_G(G1, acc1, qp1 + 4 * g, kp1 + 4 * g, sc1, __ldg(ds1 + g),
            P.G + e1 * (2 * kHD) + 16 * g, P.G + e1 * (2 * kHD) + kHD + 16 * g, two);
    if (P.da) {
#pragma unroll
      for (int o = 0; o < 32; ++o) {
        const fast::f32x2 g0 = fast::pack2(G0[o], G0[o]), g1 = fast::pack2(G1[o], G1[o]);
        const ulonglong2* wr = reinterpret_cast<const ulonglong2*>(Wn_s[g][o]);
#pragma unroll
        for (int f4 = 0; f4 < kF / 4; ++f4) {
          const ulonglong2 w = wr[f4];
          fast::fma2(da0[2 * f4], g0, w.x);
          fast::fma2(da0[2 * f4 + 1], g0, w.y);
          fast::fma2(da1[2 * f4], g1, w.x);
          fast::fma2(da1[2 * f4 + 1], g1, w.y);
        }
      }
    }
  }
  if (!P.da) return;
  {
    float da[kF];
#pragma unroll
    for (int f = 0; f < kF / 2; ++f) fast::unpack2(da0[f], da[2 * f], da[2 * f + 1]);
    finish_da(da, P.Pbuf + e0 * kH, P.d_abar ? P.d_abar + (int64_t)row0 * (kH * kF) : nullptr,
              P.da + e0 * kF);
  }
  if (two) {
    float da[kF];
#pragma unroll
    for (int f = 0; f < kF / 2; ++f) fast::unpack2(da1[f], da[2 * f], da[2 * f + 1]);
    finish_da(da, P.Pbuf + e1 * kH, P.d_abar ? P.d_abar + (int64_t)row1 * (kH * kF) : nullptr,
              P.da + e1 * kF);
  }
}
constexpr int kEdgeBwdSmem = (2 * kGroups * kF * 32 + kGroups * 32) * 4;

// ------------------------------------------------------------------------------------------
// row passes.  lane = (h = lane >> 1, half = lane & 1); CPL = C / 32 value channels per lane
// (channels lane * CPL .. of the row = channels half * CPL .. of head h).
// ------------------------------------------------------------------------------------------
template <int CPL>
__device__ __forceinline__ void ldg_vals(float (&v)[CPL], const char* p, uint64_t policy) {
  if (CPL == 4) {
    const ulonglong2 t = tile::ldg_row16(p, policy);
    float2 a = *reinterpret_cast<const float2*>(&t.x), b = *reinterpret_cast<const float2*>(&t.y);
    v[0] = a.x; v[1] = a.y; v[2 % CPL] = b.x; v[3 % CPL] = b.y;
  } else {
    const float2 t = tile::ldg_row8(p, policy);
    v[0] = t.x; v[1] = t.y;
  }
}

struct RowFwdArgs {
  const float* logits;                  // [E, 16]
  const float* v; int ldv;
  const float* a;
  const int32_t* rowptr; const int32_t* col;
  int64_t num_rows;
  float* agg_v; float* abar; float* sump; float* m; float* z;
};

struct RowFwdTiles16 {
  float p[32 * kH];                     // softmax numerators of the chunk [slot][head]
  float a[32 * kF];                     // feature rows of the chunk (cp.async)
  float mx[kH];                         // row maxima (base 2)
};

template <int CPL, bool ABAR>
__global__ void __launch_bounds__(kRowWarps * 32)
k_row_fwd16(const RowFwdArgs P) {
  extern __shared__ __align__(16) unsigned char row16_smem[];
  constexpr int C = 32 * CPL;
  RowFwdTiles16& S = reinterpret_cast<RowFwdTiles16*>(row16_smem)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (row >= P.num_rows) return;
  const int h = lane >> 1, half = lane & 1;
  const int b = P.rowptr[row], e = P.rowptr[row + 1];
  const uint64_t keep = policy_evict_last(), stream = policy_evict_first();
  const char* vbase = reinterpret_cast<const char*>(P.v) + 4 * CPL * lane;
  const unsigned ldvb = (unsigned)P.ldv * 4u;
  const float4* lg4 = reinterpret_cast<const float4*>(P.logits);

  // pass 1: maxima of the 16 heads over the row (lane = slot)
  float mx[kH];
#pragma unroll
  for (int i = 0; i < kH; ++i) mx[i] = -INFINITY;
  for (int tb = b; tb < e; tb += 32) {
    const int i = tb + lane;
    if (i < e) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 t = __ldg(lg4 + (int64_t)i * 4 + j);
        mx[4 * j] = fmaxf(mx[4 * j], t.x); mx[4 * j + 1] = fmaxf(mx[4 * j + 1], t.y);
        mx[4 * j + 2] = fmaxf(mx[4 * j + 2], t.z); mx[4 * j + 3] = fmaxf(mx[4 * j + 3], t.w);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kH; ++i) mx[i] = warp_max(mx[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < kH; ++i) S.mx[i] = mx[i];
  }
  __syncwarp();

  float accv[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) accv[i] = 0.f;
  float acca[kH];                       // abar[h'][f = lane], all 16 heads
#pragma unroll
  for (int i = 0; i < kH; ++i) acca[i] = 0.f;
  float l = 0.f;                        // sum of p of head h (both lanes of the pair)
  for (int tb = b; tb < e; tb += 32) {
    const int n = min(32, e - tb);
    __syncwarp();                       // the previous chunk's tiles have been read
    if (ABAR) stage_features(S.a, P.a, tb, n, lane, stream);
    const int i = tb + lane;
    int mycol = 0;
    {
      float4 pr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) pr[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < e) {
        mycol = P.col[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t = __ldg(lg4 + (int64_t)i * 4 + j);
          pr[j] = make_float4(ex2(t.x - mx[4 * j]), ex2(t.y - mx[4 * j + 1]),
                              ex2(t.z - mx[4 * j + 2]), ex2(t.w - mx[4 * j + 3]));
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(S.p + lane * kH + 4 * j) = pr[j];
    }
    __syncwarp();
    {
      // column sums: my head over the slots of my parity, then the other lane of the pair
      float s = 0.f;
      for (int j = half; j < 32; j += 2) s += S.p[j * kH + h];
      l += s + __shfl_xor_sync(kFull, s, 1);
    }
    // gathered value rows
    const float* p_lane = S.p + h;
    int j0 = 0;
#pragma unroll 1
    for (; j0 + 8 <= n; j0 += 8) {
      float vv[8][CPL];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const unsigned tc = (unsigned)__shfl_sync(kFull, mycol, j0 + u);
        ldg_vals<CPL>(vv[u], vbase + (uint64_t)tc * (uint64_t)ldvb, keep);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float p = p_lane[(j0 + u) * kH];
#pragma unroll
        for (int c2 = 0; c2 < CPL; ++c2) accv[c2] = fmaf(p, vv[u][c2], accv[c2]);
      }
    }
    for (; j0 < n; ++j0) {
      const unsigned tc = (unsigned)__shfl_sync(kFull, mycol, j0);
      float vv[CPL];
      ldg_vals<CPL>(vv, vbase + (uint64_t)tc * (uint64_t)ldvb, keep);
      const float p = p_lane[j0 * kH];
#pragma unroll
      for (int c2 = 0; c2 < CPL; ++c2) accv[c2] = fmaf(p, vv[c2], accv[c2]);
    }
    if (ABAR) {
      // abar[h'][f = lane] += p[j][h'] * a[j][lane]
      cp_async_wait_all();
      __syncwarp();
      for (int j = 0; j < n; ++j) {
        const float af = S.a[j * kF + lane];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 pp = *reinterpret_cast<const float4*>(S.p + j * kH + 4 * q4);
          acca[4 * q4] = fmaf(pp.x, af, acca[4 * q4]);
          acca[4 * q4 + 1] = fmaf(pp.y, af, acca[4 * q4 + 1]);
          acca[4 * q4 + 2] = fmaf(pp.z, af, acca[4 * q4 + 2]);
          acca[4 * q4 + 3] = fmaf(pp.w, af, acca[4 * q4 + 3]);
        }
      }
    }
  }

  // epilogue (PyG softmax: + 1e-16 in the denominator)
  const float zz = l + 1e-16f;
  const float iz = fast_rcp(zz);
  {
    float o[CPL];
#pragma unroll
    for (int c2 = 0; c2 < CPL; ++c2) o[c2] = accv[c2] * iz;
    if (CPL == 4)
      *reinterpret_cast<float4*>(P.agg_v + row * C + 4 * lane) =
          make_float4(o[0], o[1], o[2 % CPL], o[3 % CPL]);
    else
      *reinterpret_cast<float2*>(P.agg_v + row * C + 2 * lane) = make_float2(o[0], o[1]);
  }
  if (ABAR) {
#pragma unroll
    for (int hh = 0; hh < kH; ++hh) {
      const float inv = __shfl_sync(kFull, iz, 2 * hh);
      P.abar[row * (kH * kF) + hh * kF + lane] = acca[hh] * inv;
    }
  }
  if (half == 0) {
    P.m[row * kH + h] = (e > b) ? S.mx[h] * kLn2 : 0.f;
    P.z[row * kH + h] = zz;
    P.sump[row * kH + h] = l * iz;
  }
}
constexpr int kRowFwdSmem16 = kRowWarps * (int)sizeof(RowFwdTiles16);

struct RowBwdArgs {
  const float* logits;                  // [E, 16]
  const float* k; int ldk;
  const float* v; int ldv;
  const float* a;
  const int32_t* rowptr; const int32_t* col;
  int64_t num_rows;
  const float* Wk; const float* bk;
  int scale_mode; float scale_value;
  const float* m; const float* z;
  const float* agg_v; const float* abar; const float* d_agg_v; const float* d_abar;
  float* dq; int lddq;
  float* Pbuf; float* dS;               // [E, 16]
};

template <int CPL, bool HAS_DAB, bool HAS_WK>
__global__ void __launch_bounds__(kRowWarps * 32)
k_row_bwd16(const RowBwdArgs P) {
  extern __shared__ __align__(16) unsigned char row16_smem[];
  constexpr int C = 32 * CPL;
  constexpr bool NEED_A = HAS_DAB || HAS_WK;
  float* a_tiles = reinterpret_cast<float*>(row16_smem);                  // [warps][32][32]
  float (*Wk_s)[kF] = reinterpret_cast<float (*)[kF]>(row16_smem + kRowWarps * 32 * kF * 4);
  float* bk_s = &Wk_s[kHD][0];
  if (HAS_WK) {
    for (int i = threadIdx.x; i < kHD * kF; i += blockDim.x) Wk_s[i >> 5][i & 31] = P.Wk[i];
    if (threadIdx.x < kHD) bk_s[threadIdx.x] = P.bk ? P.bk[threadIdx.x] : 0.f;
    __syncthreads();
  }
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kRowWarps + w;
  if (row >= P.num_rows) return;
  float* a_s = a_tiles + w * 32 * kF;
  const int h = lane >> 1, half = lane & 1;
  const int b = P.rowptr[row], e = P.rowptr[row + 1];
  const uint64_t keep = policy_evict_last(), stream = policy_evict_first();
  const char* vbase = reinterpret_cast<const char*>(P.v) + 4 * CPL * lane;
  const char* kbase = reinterpret_cast<const char*>(P.k) + 4 * (h * kD + 2 * half);
  const unsigned ldvb = (unsigned)P.ldv * 4u, ldkb = (unsigned)P.ldk * 4u;

  const float scale = fast::qk_scale_fast(P.scale_mode, P.scale_value, e - b);
  const float m2 = P.m[row * kH + h] * kLog2e;
  const float zi = fast_rcp(P.z[row * kH + h]);
  float dy[CPL], dab[16];
  float delta;
  {
    float ag[CPL];
    ldg_vals<CPL>(dy, reinterpret_cast<const char*>(P.d_agg_v + row * C) + 4 * CPL * lane, keep);
    ldg_vals<CPL>(ag, reinterpret_cast<const char*>(P.agg_v + row * C) + 4 * CPL * lane, keep);
    float part = 0.f;
#pragma unroll
    for (int c2 = 0; c2 < CPL; ++c2) part = fmaf(dy[c2], ag[c2], part);
#pragma unroll
    for (int i = 0; i < 16; ++i) dab[i] = 0.f;
    if (HAS_DAB) {
      const float4* dp = reinterpret_cast<const float4*>(P.d_abar + row * (kH * kF) + h * kF + 16 * half);
      const float4* ap = reinterpret_cast<const float4*>(P.abar + row * (kH * kF) + h * kF + 16 * half);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 d4 = __ldg(dp + j), a4 = __ldg(ap + j);
        dab[4 * j] = d4.x; dab[4 * j + 1] = d4.y; dab[4 * j + 2] = d4.z; dab[4 * j + 3] = d4.w;
        part += d4.x * a4.x + d4.y * a4.y + d4.z * a4.z + d4.w * a4.w;
      }
    }
    delta = part + __shfl_xor_sync(kFull, part, 1);
  }

  float T[16];                          // T[h][16 * half + i] = sum_e dS a
#pragma unroll
  for (int i = 0; i < 16; ++i) T[i] = 0.f;
  float U0 = 0.f, U1 = 0.f, SdS = 0.f;
  for (int tb = b; tb < e; tb += 32) {
    const int n = min(32, e - tb);
    const int mycol = (lane < n) ? P.col[tb + lane] : 0;
    __syncwarp();
    if (NEED_A) {
      stage_features(a_s, P.a, tb, n, lane, stream);
      cp_async_wait_all();
      __syncwarp();
    }
    const float* a_lane = a_s + 16 * half;
#pragma unroll 2
    for (int j = 0; j < n; ++j) {
      const unsigned tc = (unsigned)__shfl_sync(kFull, mycol, j);
      float vv[CPL];
      ldg_vals<CPL>(vv, vbase + (uint64_t)tc * (uint64_t)ldvb, keep);
      const float2 kk = tile::ldg_row8(kbase + (uint64_t)tc * (uint64_t)ldkb, keep);
      const int64_t slot = (int64_t)(tb + j) * kH + h;
      const float lg = __ldg(P.logits + slot);
      float a16[16];
      if (NEED_A) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 t = *reinterpret_cast<const float4*>(a_lane + j * kF + 4 * q4);
          a16[4 * q4] = t.x; a16[4 * q4 + 1] = t.y; a16[4 * q4 + 2] = t.z; a16[4 * q4 + 3] = t.w;
        }
      }
      float part = 0.f;
#pragma unroll
      for (int c2 = 0; c2 < CPL; ++c2) part = fmaf(dy[c2], vv[c2], part);
      if (HAS_DAB) {
#pragma unroll
        for (int i = 0; i < 16; ++i) part = fmaf(dab[i], a16[i], part);
      }
      const float dp = part + __shfl_xor_sync(kFull, part, 1);
      const float p = ex2(lg - m2) * zi;
      const float ds = p * (dp - delta);
      if (half == 0) {
        P.Pbuf[slot] = p;
        P.dS[slot] = ds;
      }
      SdS += ds;
      if (HAS_WK) {
#pragma unroll
        for (int i = 0; i < 16; ++i) T[i] = fmaf(ds, a16[i], T[i]);
      }
      U0 = fmaf(ds, kk.x, U0);
      U1 = fmaf(ds, kk.y, U1);
    }
  }

  // dq[h * 4 + 2 * half + {0, 1}] = scale * (U + Wk T + bk * sum dS)
  float wt[4] = {0.f, 0.f, 0.f, 0.f};
  if (HAS_WK) {
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      const float* wr = &Wk_s[h * kD + d][16 * half];
#pragma unroll
      for (int i = 0; i < 16; ++i) wt[d] = fmaf(wr[i], T[i], wt[d]);
    }
#pragma unroll
    for (int d = 0; d < 4; ++d) wt[d] += __shfl_xor_sync(kFull, wt[d], 1);
  }
  const int d0 = 2 * half;
  const float w0 = half ? wt[2] : wt[0], w1 = half ? wt[3] : wt[1];
  const float b0 = HAS_WK ? bk_s[h * kD + d0] : 0.f, b1 = HAS_WK ? bk_s[h * kD + d0 + 1] : 0.f;
  *reinterpret_cast<float2*>(P.dq + row * P.lddq + h * kD + d0) =
      make_float2(scale * (U0 + w0 + b0 * SdS), scale * (U1 + w1 + b1 * SdS));
}
constexpr int kRowBwdSmem16 = kRowWarps * 32 * kF * 4 + (kHD * kF + kHD) * 4;


// ------------------------------------------------------------------------------------------
// backward, targets pass (CSC): dv[t] = sum_{e -> t} p_e,h dY[s_e],  dk[t] = sum_{e -> t} dk_e.
// Warp per target, the same (head, half) lane layout; no atomics.
// ------------------------------------------------------------------------------------------
struct TgtArgs {
  const int32_t* csc_ptr; const int32_t* csc_src; const int32_t* csc2csr;
  int64_t num_targets;
  const float* Pbuf;                    // [E, 16]
  const float* G;                       // [E, 128]
  const float* d_agg_v;                 // [R, C]
  float* dk; int lddk;
  float* dv; int lddv;
};

template <int CPL>
__global__ void __launch_bounds__(256)
k_attn_bwd_targets16(const TgtArgs P) {
  constexpr int C = 32 * CPL;
  const int lane = threadIdx.x & 31;
  const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= P.num_targets) return;
  const int h = lane >> 1, half = lane & 1;
  const int b = P.csc_ptr[t], e = P.csc_ptr[t + 1];
  const uint64_t keep = policy_evict_last();
  const char* dybase = reinterpret_cast<const char*>(P.d_agg_v) + 4 * CPL * lane;
  const char* gbase = reinterpret_cast<const char*>(P.G) + 4 * (kHD + h * kD + 2 * half);
  float accv[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) accv[i] = 0.f;
  float k0 = 0.f, k1 = 0.f;
  for (int tb = b; tb < e; tb += 32) {
    const int n = min(32, e - tb);
    int mysrc = 0, myslot = 0;
    if (lane < n) {
      mysrc = P.csc_src[tb + lane];
      myslot = P.csc2csr[tb + lane];
    }
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      const unsigned s = (unsigned)__shfl_sync(kFull, mysrc, j);
      const unsigned slot = (unsigned)__shfl_sync(kFull, myslot, j);
      float dy[CPL];
      ldg_vals<CPL>(dy, dybase + (uint64_t)s * (uint64_t)(C * 4), keep);
      const float p = __ldg(P.Pbuf + (uint64_t)slot * kH + h);
      const float2 gk = tile::ldg_row8(gbase + (uint64_t)slot * (uint64_t)(2 * kHD * 4), keep);
#pragma unroll
      for (int i = 0; i < CPL; ++i) accv[i] = fmaf(p, dy[i], accv[i]);
      k0 += gk.x;
      k1 += gk.y;
    }
  }
  if (CPL == 4)
    *reinterpret_cast<float4*>(P.dv + t * P.lddv + 4 * lane) =
        make_float4(accv[0], accv[1], accv[2 % CPL], accv[3 % CPL]);
  else
    *reinterpret_cast<float2*>(P.dv + t * P.lddv + 2 * lane) = make_float2(accv[0], accv[1]);
  *reinterpret_cast<float2*>(P.dk + t * P.lddk + h * kD + 2 * half) = make_float2(k0, k1);
}

}  // namespace split16
}  // namespace spt
