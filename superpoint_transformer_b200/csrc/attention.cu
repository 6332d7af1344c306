// attention.cu — fused sparse graph attention core, shape-generic path.
//
// Replaces the ~20-launch composition of src/nn/attention.py:202-315
// (3x index_select, 3 per-edge RPE Linears, einsum, 4-kernel PyG softmax,
// atomicAdd scatter_sum) by ONE CSR pass per direction:
//   forward : warp per query row, online softmax over its edges, K/V rows
//             gathered straight from the qkv buffer (L2-resident), edge features
//             streamed once; no [E, .] intermediate is materialised.
//   backward: B1 (CSR rows)  recomputes p from the saved (m, z), produces dq
//             row-locally, da per edge, and per-edge g=[dq_e|dk_e] + p scratch;
//             B2 (CSC rows)  gathers dk/dv per target without atomics;
//             B3             reduces dW = g^T a over edge slabs.
// v-RPE is applied algebraically by the caller: since sum_e p_e (Wv a_e + bv) =
// Wv (sum_e p_e a_e) + bv sum_e p_e, the kernel only emits abar = sum_e p_e a_e
// ([N,H,F]) and the per-edge [E,C] v_rpe tensor of attention.py:294-301 is never
// formed (E*C*F MACs -> N*C*F).
//
// This file is the runtime-generic path (any C<=256, H<=32, H*D<=256, F<=128,
// H*F<=512).  Shapes outside raise SPT_E_UNSUPPORTED.
#include "common.cuh"
#include "attention_fast.cuh"
#include "attention_tile.cuh"
#include "attention_split.cuh"
#include "attention_umma.cuh"
#include "attention_split16.cuh"
#include <stdlib.h>

namespace spt {

constexpr int kAttnWarps = 4;
constexpr int kVPL = 8;   // channels per lane  (C   <= 256)
constexpr int kAPL = 16;  // abar entries/lane  (H*F <= 512)
constexpr int kOPL = 8;   // q/k dims per lane  (H*D <= 256)

struct AttnShape {
  int H, D, Dv, F, C, HD, HD2, HF;
};

__device__ __forceinline__ float qk_scale(int mode, float value, int deg) {
  // src/utils/nn.py:75-127; `value` = (dim // num_heads)^-0.5 or the constant
  float g = rsqrtf((float)max(deg, 1));
  switch (mode) {
    case SPT_SCALE_D_TIMES_G: return value * g;
    case SPT_SCALE_D_PLUS_G: return value + g;
    case SPT_SCALE_D: return value;
    case SPT_SCALE_G: return g;
    default: return value;
  }
}

// shared-memory carve-up (floats):
//   Wt[F][HD2+1] | bqk[HD2] | per warp: a_s[F4] r_s[HD2] qs_s[HD] p_s[32] sc_s[32]
//   (+ backward extras)
__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }

struct FwdParams {
  const float* q; int64_t ldq;
  const float* k; int64_t ldk;
  const float* v; int64_t ldv;
  const float* a;
  const int32_t* rowptr; const int32_t* col;
  int64_t num_rows;
  AttnShape s;
  const float* Wq; const float* bq; const float* Wk; const float* bk;
  int scale_mode; float scale_value;
  float* agg_v; float* abar; float* sump; float* m; float* z;
  // optional terms (spt_attn_extras): q_e += q_row_add[s] + q_tgt_add[t], k_e += k_row_add[s],
  // softmax weights multiplied by drop_mask[e, h] after the normalisation
  const float* q_row_add; const float* q_tgt_add; const float* k_row_add; const float* drop_mask;
};

__device__ __forceinline__ void load_weights_smem(float* Wt, float* bqk, const AttnShape& s,
                                                  const float* Wq, const float* bq,
                                                  const float* Wk, const float* bk) {
  int ldw = s.HD2 + 1;
  for (int i = threadIdx.x; i < s.F * s.HD2; i += blockDim.x) {
    int f = i / s.HD2, o = i - f * s.HD2;
    float w = 0.f;
    if (o < s.HD) {
      if (Wq) w = Wq[o * s.F + f];
    } else {
      if (Wk) w = Wk[(o - s.HD) * s.F + f];
    }
    Wt[f * ldw + o] = w;
  }
  for (int o = threadIdx.x; o < s.HD2; o += blockDim.x) {
    float b = 0.f;
    if (o < s.HD) {
      if (bq && Wq) b = bq[o];
    } else {
      if (bk && Wk) b = bk[o - s.HD];
    }
    bqk[o] = b;
  }
}

__global__ void __launch_bounds__(kAttnWarps * kWarp)
k_attn_fwd_generic(FwdParams P) {
  extern __shared__ float smem[];
  const AttnShape s = P.s;
  const int ldw = s.HD2 + 1;
  const int F4 = round4(max(s.F, 1));
  float* Wt = smem;
  float* bqk = Wt + s.F * ldw;
  float* warp_base = bqk + s.HD2;
  const int per_warp = F4 + s.HD2 + 2 * s.HD + 32 + 32 + 32;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* a_s = warp_base + w * per_warp;
  float* r_s = a_s + F4;
  float* qs_s = r_s + s.HD2;
  float* kr_s = qs_s + s.HD;
  float* p_s = kr_s + s.HD;
  float* sc_s = p_s + 32;
  float* inv_s = sc_s + 32;

  load_weights_smem(Wt, bqk, s, P.Wq, P.bq, P.Wk, P.bk);
  __syncthreads();

  const bool has_a = (P.a != nullptr) && s.F > 0;
  int64_t row = (int64_t)blockIdx.x * kAttnWarps + w;
  if (row >= P.num_rows) return;
  const int b = P.rowptr[row], e = P.rowptr[row + 1];
  const float scale = qk_scale(P.scale_mode, P.scale_value, e - b);

  for (int o = lane; o < s.HD; o += 32) {
    qs_s[o] = P.q[row * P.ldq + o] * scale + (P.q_row_add ? P.q_row_add[row * s.HD + o] : 0.f);
    kr_s[o] = P.k_row_add ? P.k_row_add[row * s.HD + o] : 0.f;
  }
  float m_run = -INFINITY, l_run = 0.f, le_run = 0.f;  // lanes < H (le: dropout-masked sum)
  float acc_v[kVPL], acc_a[kAPL];
#pragma unroll
  for (int i = 0; i < kVPL; ++i) acc_v[i] = 0.f;
#pragma unroll
  for (int i = 0; i < kAPL; ++i) acc_a[i] = 0.f;
  __syncwarp();

  for (int j = b; j < e; ++j) {
    const int64_t t = P.col[j];
    if (has_a) {
      for (int f = lane; f < s.F; f += 32) a_s[f] = ldg_stream(P.a + (int64_t)j * s.F + f);
    }
    __syncwarp();
    for (int o = lane; o < s.HD2; o += 32) {
      float acc = bqk[o];
      if (has_a) {
        for (int f = 0; f < s.F; ++f) acc = fmaf(Wt[f * ldw + o], a_s[f], acc);
      }
      float base = (o < s.HD)
                       ? qs_s[o] + (P.q_tgt_add ? P.q_tgt_add[t * s.HD + o] : 0.f)
                       : P.k[t * P.ldk + (o - s.HD)] + kr_s[o - s.HD];
      r_s[o] = base + acc;
    }
    __syncwarp();
    if (lane < s.H) {
      float c = 0.f;
      for (int d = 0; d < s.D; ++d) c = fmaf(r_s[lane * s.D + d], r_s[s.HD + lane * s.D + d], c);
      float m_new = fmaxf(m_run, c);
      float p = expf(c - m_new);
      float alpha = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
      l_run = fmaf(l_run, alpha, p);
      m_run = m_new;
      if (P.drop_mask) p *= P.drop_mask[(int64_t)j * s.H + lane];   // attention dropout
      le_run = fmaf(le_run, alpha, p);
      p_s[lane] = p;
      sc_s[lane] = alpha;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kVPL; ++i) {
      int ch = lane + 32 * i;
      if (ch < s.C) {
        int h = ch / s.Dv;
        acc_v[i] = fmaf(acc_v[i], sc_s[h], p_s[h] * P.v[t * P.ldv + ch]);
      }
    }
    if (has_a && P.abar) {
#pragma unroll
      for (int i = 0; i < kAPL; ++i) {
        int idx = lane + 32 * i;
        if (idx < s.HF) {
          int h = idx / s.F, f = idx - h * s.F;
          acc_a[i] = fmaf(acc_a[i], sc_s[h], p_s[h] * a_s[f]);
        }
      }
    }
    __syncwarp();
  }

  if (lane < s.H) {
    float zden = l_run + 1e-16f;  // PyG softmax: + 1e-16 after the sum
    float inv = 1.f / zden;
    inv_s[lane] = inv;
    P.m[row * s.H + lane] = (e > b) ? m_run : 0.f;
    P.z[row * s.H + lane] = zden;
    P.sump[row * s.H + lane] = le_run * inv;
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < kVPL; ++i) {
    int ch = lane + 32 * i;
    if (ch < s.C) P.agg_v[row * s.C + ch] = acc_v[i] * inv_s[ch / s.Dv];
  }
  if (P.abar) {
#pragma unroll
    for (int i = 0; i < kAPL; ++i) {
      int idx = lane + 32 * i;
      if (idx < s.HF) P.abar[row * s.HF + idx] = acc_a[i] * inv_s[idx / s.F];
    }
  }
}

// ---------------------------------------------------------------- backward B1
struct BwdParams {
  const float* q; int64_t ldq;
  const float* k; int64_t ldk;
  const float* v; int64_t ldv;
  const float* a;
  const int32_t* rowptr; const int32_t* col;
  int64_t num_rows;
  AttnShape s;
  const float* Wq; const float* bq; const float* Wk; const float* bk;
  int scale_mode; float scale_value;
  const float* m; const float* z;
  const float* agg_v; const float* abar;
  const float* d_agg_v; const float* d_abar;
  float* dq; int64_t lddq;
  float* da;
  float* Pbuf; float* G;
  const float* q_row_add; const float* q_tgt_add; const float* k_row_add; const float* drop_mask;
  float* d_q_row_add; float* d_k_row_add;     // [R, HD] row sums of dq_e / dk_e (nullable)
  const float* d_sump; const float* sump;     // [R, H] gradient into / value of sum_e p_e mask_e
};

__global__ void __launch_bounds__(kAttnWarps * kWarp)
k_attn_bwd_rows_generic(BwdParams P) {
  extern __shared__ float smem[];
  const AttnShape s = P.s;
  const int ldw = s.HD2 + 1;
  const int F4 = round4(max(s.F, 1));
  float* Wt = smem;
  float* bqk = Wt + s.F * ldw;
  float* warp_base = bqk + s.HD2;
  // a_s[F4] r_s[HD2] g_s[HD2] qs_s[HD] kr_s[HD] p_s dc_s delta_s m_s zinv_s mk_s (6*32) dy_s[C] dab_s[HF]
  const int per_warp = F4 + 2 * s.HD2 + 2 * s.HD + 6 * 32 + round4(s.C) + round4(max(s.HF, 1));
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* a_s = warp_base + w * per_warp;
  float* r_s = a_s + F4;
  float* g_s = r_s + s.HD2;
  float* qs_s = g_s + s.HD2;
  float* kr_s = qs_s + s.HD;
  float* p_s = kr_s + s.HD;
  float* dc_s = p_s + 32;
  float* delta_s = dc_s + 32;
  float* m_s = delta_s + 32;
  float* zinv_s = m_s + 32;
  float* mk_s = zinv_s + 32;      // dropout multipliers of the current edge
  float* dy_s = mk_s + 32;
  float* dab_s = dy_s + round4(s.C);

  load_weights_smem(Wt, bqk, s, P.Wq, P.bq, P.Wk, P.bk);
  __syncthreads();

  const bool has_a = (P.a != nullptr) && s.F > 0;
  const bool has_dab = has_a && (P.d_abar != nullptr) && (P.abar != nullptr);
  int64_t row = (int64_t)blockIdx.x * kAttnWarps + w;
  if (row >= P.num_rows) return;
  const int b = P.rowptr[row], e = P.rowptr[row + 1];
  const float scale = qk_scale(P.scale_mode, P.scale_value, e - b);

  for (int o = lane; o < s.HD; o += 32) {
    qs_s[o] = P.q[row * P.ldq + o] * scale + (P.q_row_add ? P.q_row_add[row * s.HD + o] : 0.f);
    kr_s[o] = P.k_row_add ? P.k_row_add[row * s.HD + o] : 0.f;
  }
  for (int c = lane; c < s.C; c += 32) dy_s[c] = P.d_agg_v[row * s.C + c];
  if (has_dab)
    for (int i = lane; i < s.HF; i += 32) dab_s[i] = P.d_abar[row * s.HF + i];
  if (lane < s.H) {
    m_s[lane] = P.m[row * s.H + lane];
    zinv_s[lane] = 1.f / P.z[row * s.H + lane];
  }
  __syncwarp();
  // delta[h] = <dY_h, agg_v_h> + <dAbar_h, abar_h>   (= sum_e p_e dp_e)
  for (int h = 0; h < s.H; ++h) {
    float part = 0.f;
    for (int d = lane; d < s.Dv; d += 32)
      part = fmaf(dy_s[h * s.Dv + d], P.agg_v[row * s.C + h * s.Dv + d], part);
    if (has_dab)
      for (int f = lane; f < s.F; f += 32)
        part = fmaf(dab_s[h * s.F + f], P.abar[row * s.HF + h * s.F + f], part);
    part = warp_sum(part);
    if (lane == 0)
      delta_s[h] = part + ((P.d_sump && P.sump) ? P.d_sump[row * s.H + h] * P.sump[row * s.H + h]
                                                 : 0.f);
  }
  float dq_acc[kOPL], dk_acc[kOPL];
#pragma unroll
  for (int i = 0; i < kOPL; ++i) dq_acc[i] = dk_acc[i] = 0.f;
  __syncwarp();

  for (int j = b; j < e; ++j) {
    const int64_t t = P.col[j];
    if (has_a)
      for (int f = lane; f < s.F; f += 32) a_s[f] = ldg_stream(P.a + (int64_t)j * s.F + f);
    __syncwarp();
    for (int o = lane; o < s.HD2; o += 32) {
      float acc = bqk[o];
      if (has_a)
        for (int f = 0; f < s.F; ++f) acc = fmaf(Wt[f * ldw + o], a_s[f], acc);
      float base = (o < s.HD)
                       ? qs_s[o] + (P.q_tgt_add ? P.q_tgt_add[t * s.HD + o] : 0.f)
                       : P.k[t * P.ldk + (o - s.HD)] + kr_s[o - s.HD];
      r_s[o] = base + acc;
    }
    __syncwarp();
    if (lane < s.H) {
      float c = 0.f;
      for (int d = 0; d < s.D; ++d) c = fmaf(r_s[lane * s.D + d], r_s[s.HD + lane * s.D + d], c);
      float p = expf(c - m_s[lane]) * zinv_s[lane];
      // attention dropout: the value path (dv, abar) sees p * mask, the softmax Jacobian p
      const float mk = P.drop_mask ? P.drop_mask[(int64_t)j * s.H + lane] : 1.f;
      p_s[lane] = p;
      mk_s[lane] = mk;
      P.Pbuf[(int64_t)j * s.H + lane] = p * mk;
    }
    __syncwarp();
    for (int h = 0; h < s.H; ++h) {
      float part = 0.f;
      for (int d = lane; d < s.Dv; d += 32)
        part = fmaf(dy_s[h * s.Dv + d], P.v[t * P.ldv + h * s.Dv + d], part);
      if (has_dab)
        for (int f = lane; f < s.F; f += 32) part = fmaf(dab_s[h * s.F + f], a_s[f], part);
      part = warp_sum(part);
      if (lane == 0) {
        if (P.d_sump) part += P.d_sump[row * s.H + h];     // d(sum_e p_e mask_e) / d p_e = mask_e
        dc_s[h] = p_s[h] * (mk_s[h] * part - delta_s[h]);
      }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 2 * kOPL; ++i) {
      int o = lane + 32 * i;
      if (o < s.HD2) {
        int oo = (o < s.HD) ? o : o - s.HD;
        int h = oo / s.D;
        // dq_e = dc * k_e ; dk_e = dc * q_e
        float g = dc_s[h] * ((o < s.HD) ? r_s[s.HD + oo] : r_s[oo]);
        g_s[o] = g;
        P.G[(int64_t)j * s.HD2 + o] = g;
        if (o < s.HD) dq_acc[i % kOPL] += g;
        else dk_acc[i % kOPL] += g;
      }
    }
    __syncwarp();
    if (has_a && P.da) {
      for (int f = lane; f < s.F; f += 32) {
        float acc = 0.f;
        if (has_dab)
          for (int h = 0; h < s.H; ++h)
            acc = fmaf(p_s[h] * mk_s[h], dab_s[h * s.F + f], acc);
        for (int o = 0; o < s.HD2; ++o) acc = fmaf(Wt[f * ldw + o], g_s[o], acc);
        P.da[(int64_t)j * s.F + f] = acc;
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < kOPL; ++i) {
    int o = lane + 32 * i;
    if (o < s.HD) {
      P.dq[row * P.lddq + o] = dq_acc[i] * scale;
      if (P.d_q_row_add) P.d_q_row_add[row * s.HD + o] = dq_acc[i];
    }
  }
  if (P.d_k_row_add) {
    // o = lane + 32 i runs over [0, 2HD): the k half (o >= HD) was accumulated in slot i % kOPL
#pragma unroll
    for (int i = 0; i < 2 * kOPL; ++i) {
      int o = lane + 32 * i;
      if (o >= s.HD && o < s.HD2) P.d_k_row_add[row * s.HD + (o - s.HD)] = dk_acc[i % kOPL];
    }
  }
}

// ---------------------------------------------------------------- backward B2
// warp per target: dv[t] = sum_in p * dY[src] ; dk[t] = sum_in dk_e
__global__ void __launch_bounds__(kAttnWarps * kWarp)
k_attn_bwd_targets_generic(const int32_t* __restrict__ csc_ptr,
                           const int32_t* __restrict__ csc_src,
                           const int32_t* __restrict__ csc2csr, int64_t num_targets,
                           AttnShape s, const float* __restrict__ Pbuf,
                           const float* __restrict__ G, const float* __restrict__ d_agg_v,
                           float* __restrict__ dk, int64_t lddk, float* __restrict__ dv,
                           int64_t lddv, float* __restrict__ d_q_tgt_add) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int64_t t = (int64_t)blockIdx.x * kAttnWarps + w;
  if (t >= num_targets) return;
  float dv_acc[kVPL], dk_acc[kOPL], dqt_acc[kOPL];
#pragma unroll
  for (int i = 0; i < kVPL; ++i) dv_acc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < kOPL; ++i) dk_acc[i] = dqt_acc[i] = 0.f;
  const int b = csc_ptr[t], e = csc_ptr[t + 1];
  for (int jt = b; jt < e; ++jt) {
    const int64_t j = csc2csr[jt];
    const int64_t src = csc_src[jt];
#pragma unroll
    for (int i = 0; i < kVPL; ++i) {
      int ch = lane + 32 * i;
      if (ch < s.C)
        dv_acc[i] = fmaf(Pbuf[j * s.H + ch / s.Dv], d_agg_v[src * s.C + ch], dv_acc[i]);
    }
#pragma unroll
    for (int i = 0; i < kOPL; ++i) {
      int o = lane + 32 * i;
      if (o < s.HD) {
        dk_acc[i] += G[j * s.HD2 + s.HD + o];
        if (d_q_tgt_add) dqt_acc[i] += G[j * s.HD2 + o];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kVPL; ++i) {
    int ch = lane + 32 * i;
    if (ch < s.C) dv[t * lddv + ch] = dv_acc[i];
  }
#pragma unroll
  for (int i = 0; i < kOPL; ++i) {
    int o = lane + 32 * i;
    if (o < s.HD) {
      dk[t * lddk + o] = dk_acc[i];
      if (d_q_tgt_add) d_q_tgt_add[t * s.HD + o] = dqt_acc[i];
    }
  }
}

// ---------------------------------------------------------------- backward B3
// dWqk[o, f] += sum_j G[j, o] * a[j, f] ; dbqk[o] += sum_j G[j, o]
constexpr int kDwThreads = 256;
constexpr int kDwTile = 32;   // edges per smem tile
constexpr int kDwPerThread = 16;

__global__ void __launch_bounds__(kDwThreads)
k_attn_bwd_dw_generic(const float* __restrict__ G, const float* __restrict__ a, int64_t E,
                      AttnShape s, int64_t edges_per_cta, float* __restrict__ dWq,
                      float* __restrict__ dbq, float* __restrict__ dWk,
                      float* __restrict__ dbk) {
  extern __shared__ float smem[];
  float* Gs = smem;                       // [kDwTile][HD2]
  float* As = Gs + kDwTile * s.HD2;       // [kDwTile][F]
  const int64_t e0 = (int64_t)blockIdx.x * edges_per_cta;
  const int64_t e1 = min(e0 + edges_per_cta, E);
  if (e0 >= e1) return;
  const int nout = s.HD2 * s.F;
  for (int obase = 0; obase < nout; obase += kDwThreads * kDwPerThread) {
    float acc[kDwPerThread], accb[kDwPerThread];
#pragma unroll
    for (int i = 0; i < kDwPerThread; ++i) acc[i] = accb[i] = 0.f;
    for (int64_t t0 = e0; t0 < e1; t0 += kDwTile) {
      int nt = (int)min((int64_t)kDwTile, e1 - t0);
      __syncthreads();
      for (int i = threadIdx.x; i < nt * s.HD2; i += kDwThreads) Gs[i] = G[t0 * s.HD2 + i];
      for (int i = threadIdx.x; i < nt * s.F; i += kDwThreads) As[i] = a[t0 * s.F + i];
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kDwPerThread; ++i) {
        int idx = obase + threadIdx.x + kDwThreads * i;
        if (idx < nout) {
          int o = idx / s.F, f = idx - o * s.F;
          float sacc = 0.f, sb = 0.f;
          for (int ee = 0; ee < nt; ++ee) {
            float g = Gs[ee * s.HD2 + o];
            sacc = fmaf(g, As[ee * s.F + f], sacc);
            sb += g;
          }
          acc[i] += sacc;
          accb[i] += sb;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kDwPerThread; ++i) {
      int idx = obase + threadIdx.x + kDwThreads * i;
      if (idx < nout) {
        int o = idx / s.F, f = idx - o * s.F;
        if (o < s.HD) {
          if (dWq) atomicAdd(&dWq[o * s.F + f], acc[i]);
          if (dbq && f == 0) atomicAdd(&dbq[o], accb[i]);
        } else {
          if (dWk) atomicAdd(&dWk[(o - s.HD) * s.F + f], acc[i]);
          if (dbk && f == 0) atomicAdd(&dbk[o - s.HD], accb[i]);
        }
      }
    }
  }
}

// preconditions of the specialised kernels: 16-byte aligned rows, 32-bit offsets
static bool fast_layout_ok(const float* v, const float* a, int64_t ldq, int64_t ldk,
                           int64_t ldv, int64_t rows) {
  return ldv % 4 == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(a) & 15) == 0 && ldq < (1 << 20) && ldk < (1 << 20) &&
         ldv < (1 << 20) && rows * (ldk > ldv ? ldk : ldv) < (int64_t)4000000000LL;
}

static int check_shape(const char* who, int H, int D, int Dv, int F, AttnShape* out) {
  SPT_REQUIRE(H >= 1 && D >= 1 && Dv >= 1 && F >= 0, SPT_E_INVALID, "%s: bad dims", who);
  AttnShape s;
  s.H = H; s.D = D; s.Dv = Dv; s.F = F;
  s.C = H * Dv; s.HD = H * D; s.HD2 = 2 * H * D; s.HF = H * F;
  SPT_REQUIRE(H <= 32 && s.C <= 32 * kVPL && s.HD <= 32 * kOPL && s.HF <= 32 * kAPL && F <= 128,
              SPT_E_UNSUPPORTED,
              "%s: shape H=%d D=%d Dv=%d F=%d outside generic kernel limits "
              "(H<=32, C<=256, H*D<=256, H*F<=512, F<=128)", who, H, D, Dv, F);
  *out = s;
  return SPT_OK;
}

}  // namespace spt

namespace spt {
namespace umma {  // csrc/gemm_umma.cu
bool tn_shape_ok(const float* A, int64_t M, int64_t N, int64_t lda, const float* B, int64_t K,
                 int64_t ldb);
int tn_launch(const float* A, int64_t M, int64_t N, int64_t lda, const float* B, int64_t K,
              int64_t ldb, float* C, int64_t ldc, float* colsum, cudaStream_t stream);
bool make_map_rows32(CUtensorMap* tm, const float* ptr, int64_t rows, int64_t ld, int box_rows);
bool make_map_rows32_bf16(CUtensorMap* tm, const void* ptr, int64_t rows, int box_rows);
}  // namespace umma
}  // namespace spt

using namespace spt;

// CSR rows per warp of the specialised attention kernels: 8 on large levels (long slabs
// amortise the tile-stream prologue), fewer on the coarse levels so that the grid still fills
// the machine `waves` times over (a 20 k-row level with 8 rows/warp is < 1 wave).  Measured
// (cfg 2): backward rows 20 k: 0.30 -> 0.26 ms, 4 k: 0.155 -> 0.087 ms with 3 waves; the forward
// only gains below one wave (4 k: 0.089 -> 0.052 ms; 20 k was slower with 2 rows/warp).
static int rows_per_warp_for(int64_t num_rows, int ctas_per_sm, int waves) {
  // test override: force a launch geometry that the heuristics only pick at scale
  if (const char* ov = getenv("SPT_ATTN_ROWS_PER_WARP")) {
    int r = atoi(ov);
    if (r >= 1 && r <= 64) return r;
  }
  const int64_t target_warps = (int64_t)device_sm_count() * ctas_per_sm * fast::kWarps * waves;
  int64_t r = num_rows / target_warps;
  if (r < 1) r = 1;
  if (r > 8 || (waves == 1 && r >= 4)) r = 8;
  return (int)r;
}

// Row-tile kernels (csrc/attention_tile.cuh): rows per warp.  A warp's edge slab streams
// through its private TMA ring, so longer blocks amortise the ring prologue; the grid should
// still cover every SM a few times over.
static int tile_rows_per_warp(int64_t num_rows, int warps_per_cta) {
  if (const char* ov = getenv("SPT_ATTN_ROWS_PER_WARP")) {
    int r = atoi(ov);
    if (r >= 1 && r <= 64) return r;
  }
  const int sms = device_sm_count();
  const int64_t target_warps = (int64_t)sms * 2 * warps_per_cta * 3;   // >= 3 waves, 2 CTAs/SM
  int64_t r = num_rows / target_warps;
  if (r < 1) r = 1;
  if (r > 8) r = 8;
  return (int)r;
}

// boxes of 8 / 16 / 24 / 32 rows x 32 columns over the CSR-ordered edge features [E, 32]
static bool make_tile_maps(tile::TileMaps* tm, const float* a, int64_t E, int F) {
  for (int i = 0; i < 4; ++i)
    if (!umma::make_map_rows32(&tm->m[i], a, E, F, 8 * (i + 1))) return false;
  return true;
}

static bool make_tile_maps_bf16(tile::TileMaps* tm, const void* a, int64_t E) {
  for (int i = 0; i < 4; ++i)
    if (!umma::make_map_rows32_bf16(&tm->m[i], a, E, 8 * (i + 1))) return false;
  return true;
}

// consecutive rows per warp of the split row kernels: the next row's operands are prefetched
// while the current one gathers, so longer runs help as long as the grid still covers the
// machine ~4 times over
static int split_rows_per_warp(int64_t num_rows) {
  if (const char* ov = getenv("SPT_ATTN_ROWS_PER_WARP")) {
    int r = atoi(ov);
    if (r >= 1 && r <= 16) return r;
  }
  const int64_t warps = (int64_t)device_sm_count() * 32 * 4;
  int64_t r = num_rows / warps;
  return (int)(r < 1 ? 1 : r > 4 ? 4 : r);
}

// split kernels (csrc/attention_split.cuh): 16-byte aligned q / k / v rows, workspace given
static bool split_layout_ok(const float* q, const float* k, const float* v, const float* a,
                            int64_t ldq, int64_t ldk, int64_t ldv, int64_t rows, int64_t E,
                            const spt_attn_extras* ex) {
  const uintptr_t al = (uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)a |
                       (uintptr_t)(ex ? ex->ws_logits : nullptr);
  return ex && ex->ws_logits && ex->edge_row && E > 0 && E < (1ll << 31) - 256 &&
         (al & 15) == 0 && ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldq < (1 << 20) &&
         ldk < (1 << 20) && ldv < (1 << 20) &&
         rows * (ldk > ldv ? ldk : ldv) < (int64_t)4000000000LL && !getenv("SPT_ATTN_NO_SPLIT");
}

static bool split_v_bf16_ok(const spt_attn_extras* ex) {
  return ex && ex->v_bf16 && ((uintptr_t)ex->v_bf16 & 7) == 0 && ex->ldv_bf16 % 4 == 0 &&
         ex->ldv_bf16 > 0 && ex->ldv_bf16 < (1 << 20);
}

static bool tile_layout_ok(const float* q, const float* k, const float* v, const float* a,
                           int64_t ldq, int64_t ldk, int64_t ldv, int64_t rows, int64_t E) {
  const uintptr_t al8 = (uintptr_t)q | (uintptr_t)k;
  return E > 0 && E < (1ll << 31) - 64 && ldv % 4 == 0 && ldq % 2 == 0 && ldk % 2 == 0 &&
         (al8 & 7) == 0 && ((uintptr_t)v & 15) == 0 && ((uintptr_t)a & 15) == 0 &&
         ldq < (1 << 20) && ldk < (1 << 20) && ldv < (1 << 20) &&
         rows * (ldk > ldv ? ldk : ldv) < (int64_t)4000000000LL && !getenv("SPT_ATTN_NO_TILE");
}

extern "C" {

int spt_attn_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v,
                 int64_t ldv, const float* a, const int32_t* rowptr, const int32_t* col,
                 int64_t num_rows, int64_t E, int H, int D, int Dv, int F, const float* Wq,
                 const float* bq, const float* Wk, const float* bk, int scale_mode,
                 float scale_value, float* agg_v, float* abar, float* sump, float* m,
                 float* z, void* stream_) {
  return spt_attn_fwd_ex(q, ldq, k, ldk, v, ldv, a, rowptr, col, num_rows, E, H, D, Dv, F, Wq, bq,
                         Wk, bk, scale_mode, scale_value, agg_v, abar, sump, m, z, nullptr,
                         stream_);
}

static bool has_extras(const spt_attn_extras* ex) {
  return ex && (ex->q_row_add || ex->q_tgt_add || ex->k_row_add || ex->drop_mask);
}

int spt_attn_fwd_ex(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v,
                    int64_t ldv, const float* a, const int32_t* rowptr, const int32_t* col,
                    int64_t num_rows, int64_t E, int H, int D, int Dv, int F, const float* Wq,
                    const float* bq, const float* Wk, const float* bk, int scale_mode,
                    float scale_value, float* agg_v, float* abar, float* sump, float* m,
                    float* z, const spt_attn_extras* ex, void* stream_) {
  const bool extras = has_extras(ex);
  SPT_REQUIRE(num_rows >= 0 && E >= 0, SPT_E_INVALID, "attn_fwd: negative size");
  if (num_rows == 0) return SPT_OK;
  AttnShape s;
  int rc = check_shape("attn_fwd", H, D, Dv, a ? F : 0, &s);
  if (rc != SPT_OK) return rc;
  SPT_REQUIRE(q && k && v && rowptr && (E == 0 || col) && agg_v && sump && m && z,
              SPT_E_INVALID, "attn_fwd: null pointer");
  SPT_REQUIRE(scale_mode >= SPT_SCALE_D_TIMES_G && scale_mode <= SPT_SCALE_CONST,
              SPT_E_INVALID, "attn_fwd: bad scale mode %d", scale_mode);
  cudaStream_t st = (cudaStream_t)stream_;
  if (!extras && a && split16::shape_ok(H, D, Dv, F) &&
      split_layout_ok(q, k, v, a, ldq, ldk, ldv, num_rows, E, ex) &&
      ((uintptr_t)agg_v & 15) == 0) {
    // the shipped head layout (H = 16): CUDA-core edge pass + row pass (attention_split16.cuh)
    split16::EdgeFwdArgs A;
    A.q = q; A.ldq = (int)ldq; A.k = k; A.ldk = (int)ldk; A.a = a;
    A.rowptr = rowptr; A.col = col; A.edge_row = ex->edge_row; A.E = E;
    A.Wq = Wq; A.bq = bq; A.Wk = Wk; A.bk = bk;
    A.scale_mode = scale_mode; A.scale_value = scale_value; A.logits = ex->ws_logits;
    // the logits of 16 heads = 4 independent 4-head problems: the tcgen05 edge pass of the
    // 4-head family, once per head group (q / k / encoder rows offset by 16 g, logits strided);
    // the CUDA-core pass (2 edges per thread) when the tensor map cannot be encoded
    bool done16 = false;
    if (!getenv("SPT_ATTN_EDGE_SIMPLE")) {
      done16 = true;
      for (int g = 0; g < 4 && done16; ++g) {
        split::EdgeFwdArgs Ag;
        Ag.q = q + 16 * g; Ag.ldq = (int)ldq; Ag.k = k + 16 * g; Ag.ldk = (int)ldk; Ag.a = a;
        Ag.rowptr = rowptr; Ag.col = col; Ag.edge_row = ex->edge_row; Ag.E = E;
        Ag.Wq = Wq ? Wq + 16 * g * F : nullptr; Ag.bq = bq ? bq + 16 * g : nullptr;
        Ag.Wk = Wk ? Wk + 16 * g * F : nullptr; Ag.bk = bk ? bk + 16 * g : nullptr;
        Ag.scale_mode = scale_mode; Ag.scale_value = scale_value;
        Ag.logits = ex->ws_logits + 4 * g; Ag.ldl = 16;
        if (!aumma::edge_logits_launch(Ag, st, &rc)) done16 = false;
        else if (rc != SPT_OK) return rc;
      }
    }
    if (!done16) {
      split16::k_edge_logits16<<<(unsigned)ceil_div(ceil_div(E, 2), split16::kEdgeThreads),
                                 split16::kEdgeThreads, 0, st>>>(A);     // 2 edges per thread
      rc = check_launch("attn_fwd(edge16)");
      if (rc != SPT_OK) return rc;
    }
    split16::RowFwdArgs B;
    B.logits = ex->ws_logits; B.v = v; B.ldv = (int)ldv; B.a = a;
    B.rowptr = rowptr; B.col = col; B.num_rows = num_rows;
    B.agg_v = agg_v; B.abar = abar; B.sump = sump; B.m = m; B.z = z;
    const unsigned grid = (unsigned)ceil_div(num_rows, split16::kRowWarps);
    const unsigned thr = split16::kRowWarps * kWarp;
    const int sm = split16::kRowFwdSmem16;
    static unsigned long long fdone[4] = {0, 0, 0, 0};
#define SPT_ROW_FWD16(CPL, AB, IDX)                                                    \
  do {                                                                                 \
    ensure_dynamic_smem(split16::k_row_fwd16<CPL, AB>, sm, &fdone[IDX]);               \
    split16::k_row_fwd16<CPL, AB><<<grid, thr, sm, st>>>(B);                           \
  } while (0)
    if (Dv == 8) { if (abar) SPT_ROW_FWD16(4, true, 0); else SPT_ROW_FWD16(4, false, 1); }
    else { if (abar) SPT_ROW_FWD16(2, true, 2); else SPT_ROW_FWD16(2, false, 3); }
#undef SPT_ROW_FWD16
    return check_launch("attn_fwd(row16)");
  }
  if (!extras && a && split::shape_ok(H, D, Dv, F) &&
      split_layout_ok(q, k, v, a, ldq, ldk, ldv, num_rows, E, ex) &&
      ((uintptr_t)agg_v & 15) == 0 && (!abar || ((uintptr_t)abar & 15) == 0)) {
    // edge pass: base-2 logits [E, 4]; row pass: softmax + gathered sums
    split::EdgeFwdArgs A;
    A.q = q; A.ldq = (int)ldq; A.k = k; A.ldk = (int)ldk; A.a = a;
    A.rowptr = rowptr; A.col = col; A.edge_row = ex->edge_row; A.E = E;
    A.Wq = Wq; A.bq = bq; A.Wk = Wk; A.bk = bk;
    A.scale_mode = scale_mode; A.scale_value = scale_value;
    A.logits = ex->ws_logits; A.ldl = 4;
    // tcgen05 tiles of 128 edges; the CUDA-core pass when the tensor map cannot be encoded
    // (or SPT_ATTN_EDGE_SIMPLE is set: the A/B of the tests)
    if (getenv("SPT_ATTN_EDGE_SIMPLE") || !aumma::edge_logits_launch(A, st, &rc)) {
      split::k_edge_logits_simple<<<(unsigned)ceil_div(E, split::kEdgeThreads),
                                    split::kEdgeThreads, 0, st>>>(A);
      rc = check_launch("attn_fwd(edge)");
    }
    if (rc != SPT_OK) return rc;
    split::RowFwdArgs B;
    const bool vbf = split_v_bf16_ok(ex);
    B.logits = ex->ws_logits; B.a = a;
    B.v = vbf ? (const void*)ex->v_bf16 : (const void*)v;
    B.ldv = vbf ? (int)ex->ldv_bf16 : (int)ldv;
    B.rowptr = rowptr; B.col = col; B.num_rows = num_rows;
    B.agg_v = agg_v; B.abar = abar; B.sump = sump; B.m = m; B.z = z;
    B.rows_per_warp = split_rows_per_warp(num_rows);
    const unsigned grid =
        (unsigned)ceil_div(num_rows, (int64_t)split::kRowWarps * B.rows_per_warp);
    const unsigned thr = split::kRowWarps * kWarp;
    if (abar && vbf) split::k_row_fwd<true, true><<<grid, thr, 0, st>>>(B);
    else if (abar) split::k_row_fwd<true, false><<<grid, thr, 0, st>>>(B);
    else if (vbf) split::k_row_fwd<false, true><<<grid, thr, 0, st>>>(B);
    else split::k_row_fwd<false, false><<<grid, thr, 0, st>>>(B);
    return check_launch("attn_fwd(row)");
  }
  if (!extras && a && tile::shape_ok(H, D, Dv, F) &&
      tile_layout_ok(q, k, v, a, ldq, ldk, ldv, num_rows, E)) {
    tile::TileMaps tmA;
    if (make_tile_maps(&tmA, a, E, F)) {
      tile::FwdArgs A;
      A.q = q; A.ldq = (int)ldq; A.k = k; A.ldk = (int)ldk; A.v = v; A.ldv = (int)ldv;
      A.rowptr = rowptr; A.col = col; A.num_rows = num_rows;
      A.Wq = Wq; A.bq = bq; A.Wk = Wk; A.bk = bk;
      A.scale_mode = scale_mode; A.scale_value = scale_value;
      A.agg_v = agg_v; A.abar = abar; A.sump = sump; A.m = m; A.z = z;
      A.rows_per_warp = tile_rows_per_warp(num_rows, tile::kFwdWarps);
      const int smem = tile::FwdSmem::total + 1024;
      static unsigned long long done = 0;
      ensure_dynamic_smem(tile::k_attn_fwd_tile<false>, smem, &done);
      const int64_t warps = ceil_div(num_rows, A.rows_per_warp);
      tile::k_attn_fwd_tile<false><<<(unsigned)ceil_div(warps, tile::kFwdWarps), tile::kFwdWarps * kWarp,
                              smem, st>>>(tmA, A);
      return check_launch("attn_fwd(tile)");
    }
  }
  if (!extras && a && fast::shape_ok(H, D, Dv, F) &&
      fast_layout_ok(v, a, ldq, ldk, ldv, num_rows)) {
    fast::FwdArgs A;
    A.q = q; A.ldq = (int)ldq; A.k = k; A.ldk = (int)ldk; A.v = v; A.ldv = (int)ldv; A.a = a;
    A.rowptr = rowptr; A.col = col; A.num_rows = num_rows;
    A.Wq = Wq; A.bq = bq; A.Wk = Wk; A.bk = bk;
    A.scale_mode = scale_mode; A.scale_value = scale_value;
    A.agg_v = agg_v; A.abar = abar; A.sump = sump; A.m = m; A.z = z;
    A.rows_per_warp = rows_per_warp_for(num_rows, 5, 1);
    static unsigned long long attr_done = 0;
    ensure_dynamic_smem(fast::k_attn_fwd_fast, (int)fast::fwd_smem_bytes(), &attr_done);
    int64_t warps = ceil_div(num_rows, A.rows_per_warp);
    fast::k_attn_fwd_fast<<<(unsigned)ceil_div(warps, fast::kWarps), fast::kWarps * kWarp,
                            fast::fwd_smem_bytes(), st>>>(A);
    return check_launch("attn_fwd(fast)");
  }
  FwdParams P;
  P.q = q; P.ldq = ldq; P.k = k; P.ldk = ldk; P.v = v; P.ldv = ldv; P.a = a;
  P.rowptr = rowptr; P.col = col; P.num_rows = num_rows; P.s = s;
  P.Wq = a ? Wq : nullptr; P.bq = bq; P.Wk = a ? Wk : nullptr; P.bk = bk;
  P.scale_mode = scale_mode; P.scale_value = scale_value;
  P.agg_v = agg_v; P.abar = a ? abar : nullptr; P.sump = sump; P.m = m; P.z = z;
  P.q_row_add = ex ? ex->q_row_add : nullptr; P.q_tgt_add = ex ? ex->q_tgt_add : nullptr;
  P.k_row_add = ex ? ex->k_row_add : nullptr;
  P.drop_mask = (ex && E > 0) ? ex->drop_mask : nullptr;
  int F4 = round4(s.F > 1 ? s.F : 1);
  int per_warp = F4 + s.HD2 + 2 * s.HD + 96;
  size_t smem = (size_t)(s.F * (s.HD2 + 1) + s.HD2 + kAttnWarps * per_warp) * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(k_attn_fwd_generic, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
  k_attn_fwd_generic<<<(unsigned)ceil_div(num_rows, kAttnWarps), kAttnWarps * kWarp, smem, st>>>(P);
  return check_launch("attn_fwd");
}

int spt_attn_bwd_rows(const float* q, int64_t ldq, const float* k, int64_t ldk,
                      const float* v, int64_t ldv, const float* a, const int32_t* rowptr,
                      const int32_t* col, int64_t num_rows, int64_t E, int H, int D, int Dv,
                      int F, const float* Wq, const float* bq, const float* Wk,
                      const float* bk, int scale_mode, float scale_value, const float* m,
                      const float* z, const float* agg_v, const float* abar,
                      const float* d_agg_v, const float* d_abar, float* dq, int64_t lddq,
                      float* da, float* dWq, float* dbq, float* dWk, float* dbk, float* Pbuf,
                      float* G, void* stream_) {
  return spt_attn_bwd_rows_ex(q, ldq, k, ldk, v, ldv, a, rowptr, col, num_rows, E, H, D, Dv, F, Wq,
                              bq, Wk, bk, scale_mode, scale_value, m, z, agg_v, abar, d_agg_v,
                              d_abar, dq, lddq, da, dWq, dbq, dWk, dbk, Pbuf, G, nullptr, stream_);
}

int spt_attn_bwd_rows_ex(const float* q, int64_t ldq, const float* k, int64_t ldk,
                         const float* v, int64_t ldv, const float* a, const int32_t* rowptr,
                         const int32_t* col, int64_t num_rows, int64_t E, int H, int D, int Dv,
                         int F, const float* Wq, const float* bq, const float* Wk,
                         const float* bk, int scale_mode, float scale_value, const float* m,
                         const float* z, const float* agg_v, const float* abar,
                         const float* d_agg_v, const float* d_abar, float* dq, int64_t lddq,
                         float* da, float* dWq, float* dbq, float* dWk, float* dbk, float* Pbuf,
                         float* G, const spt_attn_extras* ex, void* stream_) {
  const bool extras = has_extras(ex);
  SPT_REQUIRE(num_rows >= 0 && E >= 0, SPT_E_INVALID, "attn_bwd_rows: negative size");
  if (num_rows == 0) return SPT_OK;
  AttnShape s;
  int rc = check_shape("attn_bwd_rows", H, D, Dv, a ? F : 0, &s);
  if (rc != SPT_OK) return rc;
  SPT_REQUIRE(q && k && v && rowptr && m && z && agg_v && d_agg_v && dq &&
                  (E == 0 || (col && Pbuf && G)),
              SPT_E_INVALID, "attn_bwd_rows: null pointer");
  cudaStream_t st = (cudaStream_t)stream_;
  bool tile_done = false;
  if (!extras && a && split16::shape_ok(H, D, Dv, F) &&
      split_layout_ok(q, k, v, a, ldq, ldk, ldv, num_rows, E, ex) && ex->ws_ds &&
      lddq < (1 << 20) && lddq % 2 == 0 &&
      (((uintptr_t)G | (uintptr_t)da | (uintptr_t)d_agg_v | (uintptr_t)agg_v |
        (uintptr_t)abar | (uintptr_t)d_abar | (uintptr_t)Pbuf | (uintptr_t)ex->ws_ds) & 15) == 0 &&
      ((uintptr_t)dq & 7) == 0) {
    const bool has_dab = d_abar && abar;
    split16::RowBwdArgs B;
    B.logits = ex->ws_logits; B.k = k; B.ldk = (int)ldk; B.v = v; B.ldv = (int)ldv; B.a = a;
    B.rowptr = rowptr; B.col = col; B.num_rows = num_rows;
    B.Wk = Wk; B.bk = bk; B.scale_mode = scale_mode; B.scale_value = scale_value;
    B.m = m; B.z = z; B.agg_v = agg_v; B.abar = abar; B.d_agg_v = d_agg_v; B.d_abar = d_abar;
    B.dq = dq; B.lddq = (int)lddq; B.Pbuf = Pbuf; B.dS = ex->ws_ds;
    const unsigned grid = (unsigned)ceil_div(num_rows, split16::kRowWarps);
    const unsigned thr = split16::kRowWarps * kWarp;
    const int rsm = split16::kRowBwdSmem16;
#define SPT_ROW_BWD16(CPL)                                                                     \
  do {                                                                                         \
    if (has_dab && Wk) split16::k_row_bwd16<CPL, true, true><<<grid, thr, rsm, st>>>(B);       \
    else if (has_dab) split16::k_row_bwd16<CPL, true, false><<<grid, thr, rsm, st>>>(B);       \
    else if (Wk) split16::k_row_bwd16<CPL, false, true><<<grid, thr, rsm, st>>>(B);            \
    else split16::k_row_bwd16<CPL, false, false><<<grid, thr, rsm, st>>>(B);                   \
  } while (0)
    if (Dv == 8) SPT_ROW_BWD16(4); else SPT_ROW_BWD16(2);
#undef SPT_ROW_BWD16
    rc = check_launch("attn_bwd_rows(row16)");
    if (rc != SPT_OK) return rc;
    split16::EdgeBwdArgs A;
    A.q = q; A.ldq = (int)ldq; A.k = k; A.ldk = (int)ldk; A.a = a;
    A.rowptr = rowptr; A.col = col; A.edge_row = ex->edge_row; A.E = E;
    A.Wq = Wq; A.bq = bq; A.Wk = Wk; A.bk = bk;
    A.scale_mode = scale_mode; A.scale_value = scale_value;
    A.dS = ex->ws_ds; A.Pbuf = Pbuf; A.d_abar = has_dab ? d_abar : nullptr;
    A.G = G; A.da = da;
    // G and da of 16 heads = the tcgen05 edge pass of the 4-head family once per head group:
    // G halves stored at their columns of [E, 128], da summed by TMA reduce-add stores.  The
    // CUDA-core pass (2 edges per thread) is the fallback.
    bool done16 = false;
    if (!getenv("SPT_ATTN_EDGE_SIMPLE")) {
      done16 = true;
      for (int g = 0; g < 4 && done16; ++g) {
        split::EdgeBwdArgs Ag;
        Ag.q = q + 16 * g; Ag.ldq = (int)ldq; Ag.k = k + 16 * g; Ag.ldk = (int)ldk; Ag.a = a;
        Ag.rowptr = rowptr; Ag.col = col; Ag.edge_row = ex->edge_row; Ag.E = E;
        Ag.Wq = Wq ? Wq + 16 * g * F : nullptr; Ag.bq = bq ? bq + 16 * g : nullptr;
        Ag.Wk = Wk ? Wk + 16 * g * F : nullptr; Ag.bk = bk ? bk + 16 * g : nullptr;
        Ag.scale_mode = scale_mode; Ag.scale_value = scale_value;
        Ag.dS = ex->ws_ds + 4 * g; Ag.Pbuf = Pbuf + 4 * g;
        Ag.d_abar = has_dab ? d_abar + 128 * g : nullptr;
        Ag.G = G; Ag.da = da;
        Ag.ld_ds = 16; Ag.ld_dab = 512; Ag.g_mode = 1;
        Ag.g_col_q = 16 * g; Ag.g_col_k = 64 + 16 * g; Ag.da_reduce = g > 0;
        if (!aumma::edge_bwd_launch(Ag, st, &rc)) done16 = false;
        else if (rc != SPT_OK) return rc;
        if (!done16 && g > 0) {
          set_error("attn_bwd_rows: tensor map failed after the first head group");
          return SPT_E_UNSUPPORTED;
        }
      }
    }
    if (!done16) {
      split16::k_edge_bwd16<<<(unsigned)ceil_div(ceil_div(E, 2), split16::kEdgeThreads),
                              split16::kEdgeThreads, split16::kEdgeBwdSmem, st>>>(A);   // 2 edges / thread
      rc = check_launch("attn_bwd_rows(edge16)");
      if (rc != SPT_OK) return rc;
    }
    // d[Wq;Wk] = G^T a: tcgen05 gemm_tn on the packed gradient pair, else the slab reduction
    if (E > 0 && ((Wq && (dWq || dbq)) || (Wk && (dWk || dbk)))) {
      const bool packed = Wq && Wk && dWq && dWk && dWk == dWq + s.HD * s.F &&
                          ((!dbq && !dbk) || (bq && bk && dbq && dbk && dbk == dbq + s.HD));
      if (packed && E >= 2048 && umma::tn_shape_ok(G, E, s.HD2, s.HD2, a, s.F, s.F))
        return umma::tn_launch(G, E, s.HD2, s.HD2, a, s.F, s.F, dWq, s.F, dbq, st);
      return spt_attn_bwd_weights(G, a, E, H, D, F, Wq ? dWq : nullptr,
                                  (Wq && bq) ? dbq : nullptr, Wk ? dWk : nullptr,
                                  (Wk && bk) ? dbk : nullptr, stream_);
    }
    return SPT_OK;
  }
  if (!extras && a && split::shape_ok(H, D, Dv, F) &&
      split_layout_ok(q, k, v, a, ldq, ldk, ldv, num_rows, E, ex) && ex->ws_ds &&
      lddq < (1 << 20) &&
      (((uintptr_t)G | (uintptr_t)da | (uintptr_t)d_agg_v | (uintptr_t)agg_v |
        (uintptr_t)abar | (uintptr_t)d_abar | (uintptr_t)Pbuf | (uintptr_t)ex->ws_ds) & 15) == 0) {
    // row pass: P, dS [E, 4] and dq; edge pass: G [E, 32] and da
    const bool has_dab = d_abar && abar;
    split::RowBwdArgs B;
    B.logits = ex->ws_logits; B.k = k; B.ldk = (int)ldk; B.v = v; B.ldv = (int)ldv; B.a = a;
    B.rowptr = rowptr; B.col = col; B.num_rows = num_rows;
    B.Wk = Wk; B.bk = bk; B.scale_mode = scale_mode; B.scale_value = scale_value;
    B.m = m; B.z = z; B.agg_v = agg_v; B.abar = abar; B.d_agg_v = d_agg_v; B.d_abar = d_abar;
    B.dq = dq; B.lddq = (int)lddq; B.Pbuf = Pbuf; B.dS = ex->ws_ds;
    B.rows_per_warp = split_rows_per_warp(num_rows);
    const unsigned grid =
        (unsigned)ceil_div(num_rows, (int64_t)split::kRowWarps * B.rows_per_warp);
    const unsigned thr = split::kRowWarps * kWarp;
    const int rsm = split::kRowBwdSmem;
    const bool vbf = split_v_bf16_ok(ex);
    if (vbf) { B.v = ex->v_bf16; B.ldv = (int)ex->ldv_bf16; }
    static unsigned long long rdone[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define SPT_ROW_BWD(DAB, WK, VB, IDX)                                              \
  do {                                                                             \
    ensure_dynamic_smem(split::k_row_bwd<DAB, WK, VB>, rsm, &rdone[IDX]);          \
    split::k_row_bwd<DAB, WK, VB><<<grid, thr, rsm, st>>>(B);                      \
  } while (0)
    if (has_dab && Wk) { if (vbf) SPT_ROW_BWD(true, true, true, 0); else SPT_ROW_BWD(true, true, false, 1); }
    else if (has_dab) { if (vbf) SPT_ROW_BWD(true, false, true, 2); else SPT_ROW_BWD(true, false, false, 3); }
    else if (Wk) { if (vbf) SPT_ROW_BWD(false, true, true, 4); else SPT_ROW_BWD(false, true, false, 5); }
    else { if (vbf) SPT_ROW_BWD(false, false, true, 6); else SPT_ROW_BWD(false, false, false, 7); }
#undef SPT_ROW_BWD
    rc = check_launch("attn_bwd_rows(row)");
    if (rc != SPT_OK) return rc;
    split::EdgeBwdArgs A;
    A.q = q; A.ldq = (int)ldq; A.k = k; A.ldk = (int)ldk; A.a = a;
    A.rowptr = rowptr; A.col = col; A.edge_row = ex->edge_row; A.E = E;
    A.Wq = Wq; A.bq = bq; A.Wk = Wk; A.bk = bk;
    A.scale_mode = scale_mode; A.scale_value = scale_value;
    A.dS = ex->ws_ds; A.Pbuf = Pbuf; A.d_abar = has_dab ? d_abar : nullptr;
    A.G = G; A.da = da;
    A.ld_ds = 4; A.ld_dab = 128; A.g_mode = 0; A.g_col_q = 0; A.g_col_k = 0; A.da_reduce = 0;
    if (getenv("SPT_ATTN_EDGE_SIMPLE") || !aumma::edge_bwd_launch(A, st, &rc)) {
      split::k_edge_bwd_simple<<<(unsigned)ceil_div(E, split::kEdgeThreads), split::kEdgeThreads,
                                 0, st>>>(A);
      rc = check_launch("attn_bwd_rows(edge)");
    }
    if (rc != SPT_OK) return rc;
    tile_done = true;
  }
  if (!tile_done && !extras && a && tile::shape_ok(H, D, Dv, F) &&
      tile_layout_ok(q, k, v, a, ldq, ldk, ldv, num_rows, E) &&
      lddq < (1 << 20) && lddq % 2 == 0 && ((uintptr_t)dq & 7) == 0 &&
      ((uintptr_t)G & 7) == 0 && (!da || ((uintptr_t)da & 7) == 0) &&
      ((uintptr_t)d_agg_v & 15) == 0 && ((uintptr_t)agg_v & 15) == 0 &&
      (!abar || ((uintptr_t)abar & 15) == 0) && (!d_abar || ((uintptr_t)d_abar & 15) == 0)) {
    tile::TileMaps tmA;
    if (make_tile_maps(&tmA, a, E, F)) {
      tile::BwdArgs A;
      A.q = q; A.ldq = (int)ldq; A.k = k; A.ldk = (int)ldk; A.v = v; A.ldv = (int)ldv;
      A.rowptr = rowptr; A.col = col; A.num_rows = num_rows;
      A.Wq = Wq; A.bq = bq; A.Wk = Wk; A.bk = bk;
      A.scale_mode = scale_mode; A.scale_value = scale_value;
      A.m = m; A.z = z; A.agg_v = agg_v; A.abar = abar; A.d_agg_v = d_agg_v; A.d_abar = d_abar;
      A.dq = dq; A.lddq = (int)lddq; A.da = da; A.Pbuf = Pbuf; A.G = G;
      A.rows_per_warp = tile_rows_per_warp(num_rows, tile::kBwdWarps);
      const int smem = tile::BwdSmem::total + 1024;
      static unsigned long long done = 0;
      ensure_dynamic_smem(tile::k_attn_bwd_tile<false>, smem, &done);
      const int64_t warps = ceil_div(num_rows, A.rows_per_warp);
      tile::k_attn_bwd_tile<false><<<(unsigned)ceil_div(warps, tile::kBwdWarps), tile::kBwdWarps * kWarp,
                              smem, st>>>(tmA, A);
      int rc2 = check_launch("attn_bwd_rows(tile)");
      if (rc2 != SPT_OK) return rc2;
      tile_done = true;
    }
  }
  if (tile_done || (!extras && a && fast::shape_ok(H, D, Dv, F) &&
                    fast_layout_ok(v, a, ldq, ldk, ldv, num_rows) && lddq < (1 << 20))) {
    if (!tile_done) {
    fast::BwdArgs A;
    A.q = q; A.ldq = (int)ldq; A.k = k; A.ldk = (int)ldk; A.v = v; A.ldv = (int)ldv; A.a = a;
    A.rowptr = rowptr; A.col = col; A.num_rows = num_rows;
    A.Wq = Wq; A.bq = bq; A.Wk = Wk; A.bk = bk;
    A.scale_mode = scale_mode; A.scale_value = scale_value;
    A.m = m; A.z = z; A.agg_v = agg_v; A.abar = abar; A.d_agg_v = d_agg_v; A.d_abar = d_abar;
    A.dq = dq; A.lddq = (int)lddq; A.da = da; A.Pbuf = Pbuf; A.G = G;
    A.rows_per_warp = rows_per_warp_for(num_rows, 4, 3);
    static unsigned long long attr_done = 0;
    ensure_dynamic_smem(fast::k_attn_bwd_rows_fast, (int)fast::bwd_smem_bytes(), &attr_done);
    int64_t warps = ceil_div(num_rows, A.rows_per_warp);
    fast::k_attn_bwd_rows_fast<<<(unsigned)ceil_div(warps, fast::kWarps), fast::kWarps * kWarp,
                                 fast::bwd_smem_bytes(), st>>>(A);
    int rc2 = check_launch("attn_bwd_rows(fast)");
    if (rc2 != SPT_OK) return rc2;
    }
    if (E > 0 && ((Wq && (dWq || dbq)) || (Wk && (dWk || dbk)))) {
      // d[Wq;Wk] = G^T a, d[bq;bk] = colsum(G): a [E,32]^T [E,32] product.  When the caller
      // hands the four gradients as one contiguous [2HD, F] / [2HD] pair it runs on the
      // tcgen05 gemm_tn kernel (block-diagonal row-group batching, csrc/gemm_umma.cu).
      const bool packed = Wq && Wk && dWq && dWk && dWk == dWq + fast::kHD * fast::kF &&
                          ((!dbq && !dbk) || (bq && bk && dbq && dbk && dbk == dbq + fast::kHD));
      if (packed && E >= 2048 &&
          umma::tn_shape_ok(G, E, 2 * fast::kHD, 2 * fast::kHD, a, fast::kF, fast::kF))
        return umma::tn_launch(G, E, 2 * fast::kHD, 2 * fast::kHD, a, fast::kF, fast::kF, dWq,
                               fast::kF, dbq, st);
      fast::DwArgs W;
      W.G = G; W.a = a; W.E = E;
      W.dWq = Wq ? dWq : nullptr; W.dbq = (Wq && bq) ? dbq : nullptr;
      W.dWk = Wk ? dWk : nullptr; W.dbk = (Wk && bk) ? dbk : nullptr;
      int64_t ctas = ceil_div(E, 1024);
      if (ctas > device_sm_count() * 4) ctas = device_sm_count() * 4;
      W.edges_per_cta = ceil_div(E, ctas);
      ctas = ceil_div(E, W.edges_per_cta);
      fast::k_attn_bwd_dw_fast<<<(unsigned)ctas, 256, 0, st>>>(W);
      return check_launch("attn_bwd_weights(fast)");
    }
    return SPT_OK;
  }
  BwdParams P;
  P.q = q; P.ldq = ldq; P.k = k; P.ldk = ldk; P.v = v; P.ldv = ldv; P.a = a;
  P.rowptr = rowptr; P.col = col; P.num_rows = num_rows; P.s = s;
  P.Wq = a ? Wq : nullptr; P.bq = bq; P.Wk = a ? Wk : nullptr; P.bk = bk;
  P.scale_mode = scale_mode; P.scale_value = scale_value;
  P.m = m; P.z = z; P.agg_v = agg_v; P.abar = abar; P.d_agg_v = d_agg_v; P.d_abar = d_abar;
  P.dq = dq; P.lddq = lddq; P.da = da; P.Pbuf = Pbuf; P.G = G;
  P.q_row_add = ex ? ex->q_row_add : nullptr; P.q_tgt_add = ex ? ex->q_tgt_add : nullptr;
  P.k_row_add = ex ? ex->k_row_add : nullptr;
  P.drop_mask = (ex && E > 0) ? ex->drop_mask : nullptr;
  P.d_q_row_add = ex ? ex->d_q_row_add : nullptr;
  P.d_k_row_add = ex ? ex->d_k_row_add : nullptr;
  P.d_sump = ex ? ex->d_sump : nullptr;
  P.sump = ex ? ex->sump : nullptr;
  int F4 = round4(s.F > 1 ? s.F : 1);
  int per_warp = F4 + 2 * s.HD2 + 2 * s.HD + 192 + round4(s.C) + round4(s.HF > 1 ? s.HF : 1);
  size_t smem = (size_t)(s.F * (s.HD2 + 1) + s.HD2 + kAttnWarps * per_warp) * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(k_attn_bwd_rows_generic, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
  k_attn_bwd_rows_generic<<<(unsigned)ceil_div(num_rows, kAttnWarps), kAttnWarps * kWarp, smem,
                            st>>>(P);
  rc = check_launch("attn_bwd_rows");
  if (rc != SPT_OK) return rc;
  // generic path: dW = G^T a as a separate slab reduction
  if (a && s.F > 0 && E > 0 && (dWq || dWk || dbq || dbk))
    return spt_attn_bwd_weights(G, a, E, H, D, F, Wq ? dWq : nullptr, (Wq && bq) ? dbq : nullptr,
                                Wk ? dWk : nullptr, (Wk && bk) ? dbk : nullptr, stream_);
  return SPT_OK;
}

int spt_attn_bwd_targets(const int32_t* csc_ptr, const int32_t* csc_src,
                         const int32_t* csc2csr, int64_t num_targets, int64_t E, int H, int D,
                         int Dv, const float* Pbuf, const float* G, const float* d_agg_v,
                         float* dk, int64_t lddk, float* dv, int64_t lddv, void* stream_) {
  return spt_attn_bwd_targets_ex(csc_ptr, csc_src, csc2csr, num_targets, E, H, D, Dv, Pbuf, G,
                                 d_agg_v, dk, lddk, dv, lddv, nullptr, stream_);
}

int spt_attn_bwd_targets_ex(const int32_t* csc_ptr, const int32_t* csc_src,
                            const int32_t* csc2csr, int64_t num_targets, int64_t E, int H, int D,
                            int Dv, const float* Pbuf, const float* G, const float* d_agg_v,
                            float* dk, int64_t lddk, float* dv, int64_t lddv,
                            float* d_q_tgt_add, void* stream_) {
  SPT_REQUIRE(num_targets >= 0 && E >= 0, SPT_E_INVALID, "attn_bwd_targets: negative size");
  if (num_targets == 0) return SPT_OK;
  AttnShape s;
  int rc = check_shape("attn_bwd_targets", H, D, Dv, 0, &s);
  if (rc != SPT_OK) return rc;
  SPT_REQUIRE(csc_ptr && dk && dv && (E == 0 || (csc_src && csc2csr && Pbuf && G && d_agg_v)),
              SPT_E_INVALID, "attn_bwd_targets: null pointer");
  if (!d_q_tgt_add && split16::shape_ok(H, D, Dv, split16::kF) && lddv % 4 == 0 &&
      lddk % 2 == 0 && lddk < (1 << 20) && lddv < (1 << 20) && E > 0 &&
      (((uintptr_t)dv | (uintptr_t)d_agg_v) & 15) == 0 &&
      (((uintptr_t)dk | (uintptr_t)G) & 7) == 0) {
    // the shipped head layout (H = 16): csrc/attention_split16.cuh
    split16::TgtArgs T;
    T.csc_ptr = csc_ptr; T.csc_src = csc_src; T.csc2csr = csc2csr; T.num_targets = num_targets;
    T.Pbuf = Pbuf; T.G = G; T.d_agg_v = d_agg_v;
    T.dk = dk; T.lddk = (int)lddk; T.dv = dv; T.lddv = (int)lddv;
    const unsigned grid = (unsigned)ceil_div(num_targets * 32, 256);
    if (Dv == 8) split16::k_attn_bwd_targets16<4><<<grid, 256, 0, (cudaStream_t)stream_>>>(T);
    else split16::k_attn_bwd_targets16<2><<<grid, 256, 0, (cudaStream_t)stream_>>>(T);
    return check_launch("attn_bwd_targets(16)");
  }
  if (!d_q_tgt_add && fast::shape_ok(H, D, Dv, fast::kF) && lddv % 4 == 0 && lddk < (1 << 20) &&
      lddv < (1 << 20) && (reinterpret_cast<uintptr_t>(dv) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(d_agg_v) & 15) == 0) {
    fast::TgtArgs T;
    T.csc_ptr = csc_ptr; T.csc_src = csc_src; T.csc2csr = csc2csr; T.num_targets = num_targets;
    T.Pbuf = Pbuf; T.G = G; T.d_agg_v = d_agg_v;
    T.dk = dk; T.lddk = (int)lddk; T.dv = dv; T.lddv = (int)lddv;
    fast::k_attn_bwd_targets_fast<<<(unsigned)ceil_div(num_targets * 32, 256), 256, 0,
                                    (cudaStream_t)stream_>>>(T);
    return check_launch("attn_bwd_targets(fast)");
  }
  k_attn_bwd_targets_generic<<<(unsigned)ceil_div(num_targets, kAttnWarps), kAttnWarps * kWarp,
                               0, (cudaStream_t)stream_>>>(
      csc_ptr, csc_src, csc2csr, num_targets, s, Pbuf, G, d_agg_v, dk, lddk, dv, lddv,
      d_q_tgt_add);
  return check_launch("attn_bwd_targets");
}

int spt_attn_bwd_weights(const float* G, const float* a, int64_t E, int H, int D, int F,
                         float* dWq, float* dbq, float* dWk, float* dbk, void* stream_) {
  SPT_REQUIRE(E >= 0, SPT_E_INVALID, "attn_bwd_weights: negative size");
  if (E == 0 || F == 0 || !(dWq || dWk || dbq || dbk)) return SPT_OK;
  AttnShape s;
  int rc = check_shape("attn_bwd_weights", H, D, 1, F, &s);
  if (rc != SPT_OK) return rc;
  SPT_REQUIRE(G && a, SPT_E_INVALID, "attn_bwd_weights: null pointer");
  int64_t ctas = ceil_div(E, 2048);
  if (ctas > device_sm_count() * 4) ctas = device_sm_count() * 4;
  int64_t per = ceil_div(E, ctas);
  per = ceil_div(per, kDwTile) * kDwTile;
  ctas = ceil_div(E, per);
  size_t smem = (size_t)kDwTile * (s.HD2 + s.F) * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(k_attn_bwd_dw_generic, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
  k_attn_bwd_dw_generic<<<(unsigned)ctas, kDwThreads, smem, (cudaStream_t)stream_>>>(
      G, a, E, s, per, dWq, dbq, dWk, dbk);
  return check_launch("attn_bwd_weights");
}


// --------------------------------------------------------------------------------------------
// bf16 storage of the attention operands (BASELINE cfg 3): q / k / v and the edge features are
// bf16 in HBM (halves the gather and stream traffic), every sum is fp32; shape family of the
// row-tile kernels only.
// --------------------------------------------------------------------------------------------
static bool bf16_layout_ok(const void* q, const void* k, const void* v, const void* a,
                           int64_t ldq, int64_t ldk, int64_t ldv, int64_t rows, int64_t E) {
  return E > 0 && E < (1ll << 31) - 64 && ldq % 2 == 0 && ldk % 2 == 0 && ldv % 4 == 0 &&
         (((uintptr_t)q | (uintptr_t)k) & 3) == 0 && ((uintptr_t)v & 7) == 0 &&
         ((uintptr_t)a & 15) == 0 && ldq < (1 << 20) && ldk < (1 << 20) && ldv < (1 << 20) &&
         rows * (ldk > ldv ? ldk : ldv) * 2 < (int64_t)4000000000LL;
}

int spt_attn_fwd_bf16(const uint16_t* q, int64_t ldq, const uint16_t* k, int64_t ldk,
                      const uint16_t* v, int64_t ldv, const uint16_t* a, const int32_t* rowptr,
                      const int32_t* col, int64_t num_rows, int64_t E, int H, int D, int Dv, int F,
                      const float* Wq, const float* bq, const float* Wk, const float* bk,
                      int scale_mode, float scale_value, float* agg_v, float* abar, float* sump,
                      float* m, float* z, void* stream_) {
  SPT_REQUIRE(num_rows >= 0 && E >= 0, SPT_E_INVALID, "attn_fwd_bf16: negative size");
  if (num_rows == 0) return SPT_OK;
  SPT_REQUIRE(q && k && v && a && rowptr && col && agg_v && sump && m && z, SPT_E_INVALID,
              "attn_fwd_bf16: null pointer");
  SPT_REQUIRE(tile::shape_ok(H, D, Dv, F) && bf16_layout_ok(q, k, v, a, ldq, ldk, ldv, num_rows, E),
              SPT_E_UNSUPPORTED,
              "attn_fwd_bf16: only H=4, D=4, Dv=32, F=32 with aligned operands and E > 0");
  tile::TileMaps tmA;
  SPT_REQUIRE(make_tile_maps_bf16(&tmA, a, E), SPT_E_UNSUPPORTED,
              "attn_fwd_bf16: cuTensorMapEncodeTiled failed");
  tile::FwdArgs A;
  A.q = q; A.ldq = (int)ldq; A.k = k; A.ldk = (int)ldk; A.v = v; A.ldv = (int)ldv;
  A.rowptr = rowptr; A.col = col; A.num_rows = num_rows;
  A.Wq = Wq; A.bq = bq; A.Wk = Wk; A.bk = bk;
  A.scale_mode = scale_mode; A.scale_value = scale_value;
  A.agg_v = agg_v; A.abar = abar; A.sump = sump; A.m = m; A.z = z;
  A.rows_per_warp = tile_rows_per_warp(num_rows, tile::kFwdWarps);
  const int smem = tile::FwdSmem::total + 1024;
  static unsigned long long done = 0;
  ensure_dynamic_smem(tile::k_attn_fwd_tile<true>, smem, &done);
  const int64_t warps = ceil_div(num_rows, A.rows_per_warp);
  tile::k_attn_fwd_tile<true><<<(unsigned)ceil_div(warps, tile::kFwdWarps), tile::kFwdWarps * kWarp,
                                smem, (cudaStream_t)stream_>>>(tmA, A);
  return check_launch("attn_fwd_bf16");
}

int spt_attn_bwd_rows_bf16(const uint16_t* q, int64_t ldq, const uint16_t* k, int64_t ldk,
                           const uint16_t* v, int64_t ldv, const uint16_t* a,
                           const int32_t* rowptr, const int32_t* col, int64_t num_rows, int64_t E,
                           int H, int D, int Dv, int F, const float* Wq, const float* bq,
                           const float* Wk, const float* bk, int scale_mode, float scale_value,
                           const float* m, const float* z, const float* agg_v, const float* abar,
                           const float* d_agg_v, const float* d_abar, float* dq, int64_t lddq,
                           float* da, float* Pbuf, float* G, void* stream_) {
  SPT_REQUIRE(num_rows >= 0 && E >= 0, SPT_E_INVALID, "attn_bwd_rows_bf16: negative size");
  if (num_rows == 0) return SPT_OK;
  SPT_REQUIRE(q && k && v && a && rowptr && col && m && z && agg_v && d_agg_v && dq && Pbuf && G,
              SPT_E_INVALID, "attn_bwd_rows_bf16: null pointer");
  SPT_REQUIRE(tile::shape_ok(H, D, Dv, F) && bf16_layout_ok(q, k, v, a, ldq, ldk, ldv, num_rows, E) &&
                  lddq < (1 << 20) && lddq % 2 == 0 && ((uintptr_t)dq & 7) == 0 &&
                  ((uintptr_t)G & 7) == 0 && (!da || ((uintptr_t)da & 7) == 0) &&
                  ((uintptr_t)d_agg_v & 15) == 0 && ((uintptr_t)agg_v & 15) == 0 &&
                  (!abar || ((uintptr_t)abar & 15) == 0) && (!d_abar || ((uintptr_t)d_abar & 15) == 0),
              SPT_E_UNSUPPORTED, "attn_bwd_rows_bf16: unsupported shape / alignment");
  tile::TileMaps tmA;
  SPT_REQUIRE(make_tile_maps_bf16(&tmA, a, E), SPT_E_UNSUPPORTED,
              "attn_bwd_rows_bf16: cuTensorMapEncodeTiled failed");
  tile::BwdArgs A;
  A.q = q; A.ldq = (int)ldq; A.k = k; A.ldk = (int)ldk; A.v = v; A.ldv = (int)ldv;
  A.rowptr = rowptr; A.col = col; A.num_rows = num_rows;
  A.Wq = Wq; A.bq = bq; A.Wk = Wk; A.bk = bk;
  A.scale_mode = scale_mode; A.scale_value = scale_value;
  A.m = m; A.z = z; A.agg_v = agg_v; A.abar = abar; A.d_agg_v = d_agg_v; A.d_abar = d_abar;
  A.dq = dq; A.lddq = (int)lddq; A.da = da; A.Pbuf = Pbuf; A.G = G;
  A.rows_per_warp = tile_rows_per_warp(num_rows, tile::kBwdWarps);
  const int smem = tile::BwdSmem::total + 1024;
  static unsigned long long done = 0;
  ensure_dynamic_smem(tile::k_attn_bwd_tile<true>, smem, &done);
  const int64_t warps = ceil_div(num_rows, A.rows_per_warp);
  tile::k_attn_bwd_tile<true><<<(unsigned)ceil_div(warps, tile::kBwdWarps), tile::kBwdWarps * kWarp,
                                smem, (cudaStream_t)stream_>>>(tmA, A);
  return check_launch("attn_bwd_rows_bf16");
}

}  // extern "C"
