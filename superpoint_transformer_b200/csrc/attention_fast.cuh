// attention_fast.cuh — shape-specialised attention kernels for the SPT "4 heads"
// family: H*D = 16 (=> 2*H*D = 32 RPE outputs = one per lane), F = 32 edge
// features, C = H*Dv = 128 (4 contiguous channels per lane, Dv = F = 32).
// This is BASELINE.json cfg 2/3/5 (C=128, H=4, qk_dim=4, in_rpe_dim=32).
//
// Mapping (one warp walks a contiguous block of CSR rows, so its edges are a
// contiguous slab of the CSR-ordered edge-feature matrix):
//   * the slab is streamed HBM -> shared memory with 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), 32 edges (4 KB) per tile, double
//     buffered per warp — no other warp ever touches the tile, so there is not a
//     single __syncthreads() in the steady state;
//   * RPE projections  r = [Wq;Wk] a_e + b : lane o owns output o and keeps row o
//     of the 32x32 weight in registers; a_e is read as 8 broadcast LDS.128;
//   * q_e.k_e per head: lanes 0-15 hold q_e, lanes 16-31 hold k_e -> one
//     shfl_xor(16) + a 4-lane butterfly;
//   * online softmax state (m, l) is replicated in the 8 lanes of each head;
//   * lane l accumulates v channels 4l..4l+3 (one LDG.128 of the gathered row) and
//     abar entries (h=l/8, f=4(l%8)..) (one LDS.128 of the staged tile).
#pragma once
#include "common.cuh"

namespace spt {
namespace fast {

constexpr int kH = 4, kD = 4, kDv = 32, kF = 32;
constexpr int kHD = kH * kD;        // 16
constexpr int kC = kH * kDv;        // 128
constexpr int kTile = 32;           // edges per TMA tile
constexpr int kTileBytes = kTile * kF * 4;
constexpr int kWarps = 8;           // warps per CTA
constexpr int kStages = 2;

__host__ __device__ inline bool shape_ok(int H, int D, int Dv, int F) {
  return H == kH && D == kD && Dv == kDv && F == kF;
}

// ---- mbarrier / TMA bulk copy (PTX ISA 8.x, sm_90+) -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int dbg0 = 0,
                                          int dbg1 = 0) {
  uint32_t done;
#ifdef SPT_WATCHDOG
  long long t0 = clock64();
#endif
  do {
#ifdef SPT_WATCHDOG
    if (clock64() - t0 > 2000000000LL) {
      if ((threadIdx.x & 31) == 0)
        printf("mbar_wait stuck: block %d warp %d parity %u gtile %d gbase %d\n", blockIdx.x,
               threadIdx.x >> 5, parity, dbg0, dbg1);
      __trap();
    }
#endif
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> this CTA's shared memory, completion signalled on `bar`
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem,
                                            uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// Per-warp streaming reader of CSR-ordered edge-feature slabs [e0, e1).  The two
// mbarriers are initialised ONCE per warp (`init`); tiles are numbered globally
// across successive slabs (`open`) so the stage / phase-parity sequence simply
// continues — barriers are never re-initialised.
struct EdgeStream {
  const float* a;       // [E, kF] global
  float* buf;           // warp-private smem: kStages * kTile * kF floats
  uint64_t* bar;        // kStages mbarriers
  int64_t e0, e1;       // current slab
  int64_t tile_base;    // first edge of the tile currently readable
  int gtile;            // global index of that tile
  int gbase;            // global index of the slab's first tile
  int gnext;            // global index the next slab will start at
  int lane;

  __device__ __forceinline__ void init(const float* a_, float* buf_, uint64_t* bar_, int lane_) {
    a = a_; buf = buf_; bar = bar_; lane = lane_;
    gnext = 0;
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < kStages; ++s) mbar_init(&bar[s], 1);
      mbar_fence_init();
    }
    __syncwarp();
  }
  // issue slab-local tile k
  __device__ __forceinline__ void issue(int k) {
    int64_t b = e0 + (int64_t)k * kTile;
    if (b >= e1) return;
    int n = (int)min((int64_t)kTile, e1 - b);
    int s = (gbase + k) % kStages;
    if (lane == 0) {
      mbar_expect_tx(&bar[s], (uint32_t)n * kF * 4);
      tma_load_1d(buf + s * kTile * kF, a + b * kF, (uint32_t)n * kF * 4, &bar[s]);
    }
  }
  __device__ __forceinline__ void wait_current() {
    mbar_wait(&bar[gtile % kStages], (uint32_t)((gtile / kStages) & 1), gtile, gbase);
  }
  // start streaming a new slab; the previous one must have been read to its end
  __device__ __forceinline__ void open(int64_t e0_, int64_t e1_) {
    e0 = e0_; e1 = e1_;
    if (e1 <= e0) return;
    __syncwarp();                   // all lanes finished the previous slab's tiles
    gbase = gnext;
    gnext = gbase + (int)((e1 - e0 + kTile - 1) / kTile);
    gtile = gbase;
    tile_base = e0;
    issue(0);
    issue(1);
    wait_current();
  }
  // pointer to the features of CSR slot j (advances the ring when j leaves the tile);
  // j must be visited in increasing order and every slot of the slab must be visited
  __device__ __forceinline__ const float* row(int64_t j) {
    if (j >= tile_base + kTile) {
      __syncwarp();                           // every lane is done reading the old tile
      issue(gtile - gbase + kStages);         // refill the stage we just released
      ++gtile;
      tile_base += kTile;
      wait_current();
    }
    return buf + (gtile % kStages) * kTile * kF + (int)(j - tile_base) * kF;
  }
};

struct FwdArgs {
  const float* q; int64_t ldq;
  const float* k; int64_t ldk;
  const float* v; int64_t ldv;
  const float* a;
  const int32_t* rowptr; const int32_t* col;
  int64_t num_rows;
  const float* Wq; const float* bq; const float* Wk; const float* bk;
  int scale_mode; float scale_value;
  float* agg_v; float* abar; float* sump; float* m; float* z;
  int rows_per_warp;
};

__device__ __forceinline__ float qk_scale_fast(int mode, float value, int deg) {
  float g = rsqrtf((float)max(deg, 1));
  switch (mode) {
    case SPT_SCALE_D_TIMES_G: return value * g;
    case SPT_SCALE_D_PLUS_G: return value + g;
    case SPT_SCALE_D: return value;
    case SPT_SCALE_G: return g;
    default: return value;
  }
}

// r_o = b_o + sum_f w[f] * a[f], a read as 8 broadcast LDS.128; 4 partial sums
__device__ __forceinline__ float gemv32(const float (&w)[kF], float bias, const float* arow,
                                        float (&av)[kF]) {
  float s0 = bias, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int c = 0; c < kF / 4; ++c) {
    float4 t = *reinterpret_cast<const float4*>(arow + 4 * c);
    av[4 * c + 0] = t.x; av[4 * c + 1] = t.y; av[4 * c + 2] = t.z; av[4 * c + 3] = t.w;
    s0 = fmaf(w[4 * c + 0], t.x, s0);
    s1 = fmaf(w[4 * c + 1], t.y, s1);
    s2 = fmaf(w[4 * c + 2], t.z, s2);
    s3 = fmaf(w[4 * c + 3], t.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

__global__ void __launch_bounds__(kWarps * kWarp)
k_attn_fwd_fast(FwdArgs P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* buf = reinterpret_cast<float*>(smem_raw) + (size_t)w * kStages * kTile * kF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      smem_raw + (size_t)kWarps * kStages * kTileBytes) + w * kStages;

  const int64_t gw = (int64_t)blockIdx.x * kWarps + w;
  const int64_t row0 = gw * P.rows_per_warp;
  if (row0 >= P.num_rows) return;
  const int64_t row1 = min(row0 + (int64_t)P.rows_per_warp, P.num_rows);

  // lane o: row o of [Wq; Wk] and its bias
  float wreg[kF];
  float bias = 0.f;
  {
    const float* W = (lane < kHD) ? P.Wq : P.Wk;
    const float* B = (lane < kHD) ? P.bq : P.bk;
    int o = lane & (kHD - 1);
#pragma unroll
    for (int f = 0; f < kF; ++f) wreg[f] = W ? W[o * kF + f] : 0.f;
    if (W && B) bias = B[o];
  }

  EdgeStream es;
  es.init(P.a, buf, bars, lane);
  es.open(P.rowptr[row0], P.rowptr[row1]);

  const int hsel = (lane >> 3) << 2;   // lane holding compat of my head (0,4,8,12)
  for (int64_t row = row0; row < row1; ++row) {
    const int b = P.rowptr[row], e = P.rowptr[row + 1];
    const float scale = qk_scale_fast(P.scale_mode, P.scale_value, e - b);
    const float qs = (lane < kHD) ? P.q[row * P.ldq + lane] * scale : 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    float4 accv = make_float4(0.f, 0.f, 0.f, 0.f), acca = make_float4(0.f, 0.f, 0.f, 0.f);

    // software prefetch of the gathered key / value rows, one edge ahead
    int64_t t_cur = (b < e) ? P.col[b] : 0;
    float k_cur = 0.f;
    float4 v_cur = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b < e) {
      if (lane >= kHD) k_cur = P.k[t_cur * P.ldk + (lane - kHD)];
      v_cur = *reinterpret_cast<const float4*>(P.v + t_cur * P.ldv + 4 * lane);
    }
    for (int j = b; j < e; ++j) {
      float k_nxt = 0.f;
      float4 v_nxt = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j + 1 < e) {
        int64_t t_n = P.col[j + 1];
        if (lane >= kHD) k_nxt = P.k[t_n * P.ldk + (lane - kHD)];
        v_nxt = *reinterpret_cast<const float4*>(P.v + t_n * P.ldv + 4 * lane);
      }
      const float* arow = es.row(j);
      float av[kF];
      float r = gemv32(wreg, bias, arow, av);
      float val = ((lane < kHD) ? qs : k_cur) + r;          // q_e (lanes<16) | k_e
      float prod = val * __shfl_xor_sync(kFull, val, 16);
      prod += __shfl_xor_sync(kFull, prod, 1);
      prod += __shfl_xor_sync(kFull, prod, 2);              // <q_e,k_e>_h in lanes 4h..4h+3
      float c = __shfl_sync(kFull, prod, hsel);             // compat of my head
      float m_new = fmaxf(m_run, c);
      float alpha = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
      float p = expf(c - m_new);
      l_run = fmaf(l_run, alpha, p);
      m_run = m_new;
      accv.x = fmaf(accv.x, alpha, p * v_cur.x);
      accv.y = fmaf(accv.y, alpha, p * v_cur.y);
      accv.z = fmaf(accv.z, alpha, p * v_cur.z);
      accv.w = fmaf(accv.w, alpha, p * v_cur.w);
      if (P.abar) {
        const int f0 = 4 * (lane & 7);
        // av[] is lane-uniform: select my 4 features without dynamic register indexing
        float4 a4 = *reinterpret_cast<const float4*>(arow + f0);
        acca.x = fmaf(acca.x, alpha, p * a4.x);
        acca.y = fmaf(acca.y, alpha, p * a4.y);
        acca.z = fmaf(acca.z, alpha, p * a4.z);
        acca.w = fmaf(acca.w, alpha, p * a4.w);
      }
      k_cur = k_nxt;
      v_cur = v_nxt;
    }
    const float zden = l_run + 1e-16f;
    const float inv = 1.f / zden;
    *reinterpret_cast<float4*>(P.agg_v + row * kC + 4 * lane) =
        make_float4(accv.x * inv, accv.y * inv, accv.z * inv, accv.w * inv);
    if (P.abar)
      *reinterpret_cast<float4*>(P.abar + row * (kH * kF) + 4 * lane) =
          make_float4(acca.x * inv, acca.y * inv, acca.z * inv, acca.w * inv);
    if ((lane & 7) == 0) {
      int h = lane >> 3;
      P.m[row * kH + h] = (e > b) ? m_run : 0.f;
      P.z[row * kH + h] = zden;
      P.sump[row * kH + h] = l_run * inv;
    }
  }
}

// ------------------------------------------------------------------ backward rows
struct BwdArgs {
  const float* q; int64_t ldq;
  const float* k; int64_t ldk;
  const float* v; int64_t ldv;
  const float* a;
  const int32_t* rowptr; const int32_t* col;
  int64_t num_rows;
  const float* Wq; const float* bq; const float* Wk; const float* bk;
  int scale_mode; float scale_value;
  const float* m; const float* z;
  const float* agg_v; const float* abar;
  const float* d_agg_v; const float* d_abar;
  float* dq; int64_t lddq;
  float* da;
  float* Pbuf;   // [E, H]
  float* G;      // [E, 2HD]  (only the dk_e half [.., HD:2HD] is written)
  float* dWq; float* dbq; float* dWk; float* dbk;   // accumulated (atomics), nullable
  int rows_per_warp;
  int num_row_blocks;   // persistent: warps loop over row blocks
};

// smem per warp: stream (kStages tiles) + g[32]; per CTA: dW reduction [32][33]
__global__ void __launch_bounds__(kWarps * kWarp)
k_attn_bwd_rows_fast(BwdArgs P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* buf = reinterpret_cast<float*>(smem_raw) + (size_t)w * kStages * kTile * kF;
  unsigned char* after = smem_raw + (size_t)kWarps * kStages * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(after) + w * kStages;
  float* g_s = reinterpret_cast<float*>(after + kWarps * kStages * 8) + w * 32;
  float* red = reinterpret_cast<float*>(after + kWarps * kStages * 8) + kWarps * 32;  // [32][33]

  // lane o: row o of [Wq;Wk] (forward GEMV) ; lane f: column f (da GEMV)
  float wrow[kF], wcol[2 * kHD];
  float bias = 0.f;
  {
    const float* W = (lane < kHD) ? P.Wq : P.Wk;
    const float* B = (lane < kHD) ? P.bq : P.bk;
    int o = lane & (kHD - 1);
#pragma unroll
    for (int f = 0; f < kF; ++f) wrow[f] = W ? W[o * kF + f] : 0.f;
    if (W && B) bias = B[o];
#pragma unroll
    for (int oo = 0; oo < kHD; ++oo) {
      wcol[oo] = P.Wq ? P.Wq[oo * kF + lane] : 0.f;
      wcol[kHD + oo] = P.Wk ? P.Wk[oo * kF + lane] : 0.f;
    }
  }
  float dwacc[kF];   // lane o: d[Wq;Wk][o][f]
  float dbacc = 0.f;
#pragma unroll
  for (int f = 0; f < kF; ++f) dwacc[f] = 0.f;
  const bool want_dw = (P.dWq != nullptr) || (P.dWk != nullptr);

  const int hsel = (lane >> 3) << 2;
  const int myhead = lane >> 3;
  EdgeStream es;
  es.init(P.a, buf, bars, lane);

  for (int64_t blk = (int64_t)blockIdx.x * kWarps + w; blk < P.num_row_blocks;
       blk += (int64_t)gridDim.x * kWarps) {
    const int64_t row0 = blk * P.rows_per_warp;
    const int64_t row1 = min(row0 + (int64_t)P.rows_per_warp, P.num_rows);
    es.open(P.rowptr[row0], P.rowptr[row1]);
    for (int64_t row = row0; row < row1; ++row) {
      const int b = P.rowptr[row], e = P.rowptr[row + 1];
      const float scale = qk_scale_fast(P.scale_mode, P.scale_value, e - b);
      const float qs = (lane < kHD) ? P.q[row * P.ldq + lane] * scale : 0.f;
      const float m_h = P.m[row * kH + myhead];
      const float zinv = 1.f / P.z[row * kH + myhead];
      const float4 dy = *reinterpret_cast<const float4*>(P.d_agg_v + row * kC + 4 * lane);
      float4 dab = make_float4(0.f, 0.f, 0.f, 0.f);   // d_abar[row][myhead][4(l%8)..]
      float dabf[kH] = {0.f, 0.f, 0.f, 0.f};          // d_abar[row][h][lane]   (lane = f)
      const bool has_dab = P.d_abar != nullptr && P.abar != nullptr;
      // delta_h = <dY_h, agg_v_h> + <dAbar_h, abar_h>
      float delta;
      {
        float4 ag = *reinterpret_cast<const float4*>(P.agg_v + row * kC + 4 * lane);
        float part = dy.x * ag.x + dy.y * ag.y + dy.z * ag.z + dy.w * ag.w;
        if (has_dab) {
          dab = *reinterpret_cast<const float4*>(P.d_abar + row * (kH * kF) + 4 * lane);
          float4 ab = *reinterpret_cast<const float4*>(P.abar + row * (kH * kF) + 4 * lane);
          part += dab.x * ab.x + dab.y * ab.y + dab.z * ab.z + dab.w * ab.w;
#pragma unroll
          for (int h = 0; h < kH; ++h) dabf[h] = P.d_abar[row * (kH * kF) + h * kF + lane];
        }
        part += __shfl_xor_sync(kFull, part, 1);
        part += __shfl_xor_sync(kFull, part, 2);
        part += __shfl_xor_sync(kFull, part, 4);
        delta = part;   // same value in the 8 lanes of my head
      }
      float dq_acc = 0.f;

      int64_t t_cur = (b < e) ? P.col[b] : 0;
      float k_cur = 0.f;
      float4 v_cur = make_float4(0.f, 0.f, 0.f, 0.f);
      if (b < e) {
        if (lane >= kHD) k_cur = P.k[t_cur * P.ldk + (lane - kHD)];
        v_cur = *reinterpret_cast<const float4*>(P.v + t_cur * P.ldv + 4 * lane);
      }
      for (int j = b; j < e; ++j) {
        float k_nxt = 0.f;
        float4 v_nxt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j + 1 < e) {
          int64_t t_n = P.col[j + 1];
          if (lane >= kHD) k_nxt = P.k[t_n * P.ldk + (lane - kHD)];
          v_nxt = *reinterpret_cast<const float4*>(P.v + t_n * P.ldv + 4 * lane);
        }
        const float* arow = es.row(j);
        float av[kF];
        float r = gemv32(wrow, bias, arow, av);
        float val = ((lane < kHD) ? qs : k_cur) + r;
        float other = __shfl_xor_sync(kFull, val, 16);        // k_e for q lanes, q_e for k lanes
        float prod = val * other;
        prod += __shfl_xor_sync(kFull, prod, 1);
        prod += __shfl_xor_sync(kFull, prod, 2);
        float c = __shfl_sync(kFull, prod, hsel);
        float p = expf(c - m_h) * zinv;                        // softmax weight of my head
        // dp_h = <dY_h, v_h> + <dAbar_h, a>
        float part = dy.x * v_cur.x + dy.y * v_cur.y + dy.z * v_cur.z + dy.w * v_cur.w;
        if (has_dab) {
          float4 a4 = *reinterpret_cast<const float4*>(arow + 4 * (lane & 7));
          part += dab.x * a4.x + dab.y * a4.y + dab.z * a4.z + dab.w * a4.w;
        }
        part += __shfl_xor_sync(kFull, part, 1);
        part += __shfl_xor_sync(kFull, part, 2);
        part += __shfl_xor_sync(kFull, part, 4);
        float dc = p * (part - delta);                         // d compat of my head
        // g_o = dc_{h(o)} * other ; h(o) = (o % 16) / 4 lives in lanes 8*h(o)..
        float dco = __shfl_sync(kFull, dc, ((lane & 15) >> 2) << 3);
        float g = dco * other;                                 // dq_e (lanes<16) | dk_e
        if (lane < kHD) dq_acc += g;
        else P.G[(int64_t)j * (2 * kHD) + lane] = g;
        if ((lane & 7) == 0) P.Pbuf[(int64_t)j * kH + myhead] = p;
        if (want_dw) {
#pragma unroll
          for (int f = 0; f < kF; ++f) dwacc[f] = fmaf(g, av[f], dwacc[f]);
          dbacc += g;
        }
        if (P.da) {
          g_s[lane] = g;
          __syncwarp();
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          if (has_dab) {
            // sum_h p_h * dAbar[row][h][f=lane]
            s0 = __shfl_sync(kFull, p, 0) * dabf[0];
            s1 = __shfl_sync(kFull, p, 8) * dabf[1];
            s2 = __shfl_sync(kFull, p, 16) * dabf[2];
            s3 = __shfl_sync(kFull, p, 24) * dabf[3];
          }
#pragma unroll
          for (int c4 = 0; c4 < (2 * kHD) / 4; ++c4) {
            float4 gg = *reinterpret_cast<const float4*>(g_s + 4 * c4);
            s0 = fmaf(wcol[4 * c4 + 0], gg.x, s0);
            s1 = fmaf(wcol[4 * c4 + 1], gg.y, s1);
            s2 = fmaf(wcol[4 * c4 + 2], gg.z, s2);
            s3 = fmaf(wcol[4 * c4 + 3], gg.w, s3);
          }
          P.da[(int64_t)j * kF + lane] = (s0 + s1) + (s2 + s3);
          __syncwarp();
        }
        k_cur = k_nxt;
        v_cur = v_nxt;
      }
      if (lane < kHD) P.dq[row * P.lddq + lane] = dq_acc * scale;
    }
  }

  // ---- reduce dW over the CTA's warps in shared memory, then one atomic per entry
  if (want_dw) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int f = 0; f < kF; ++f) atomicAdd(&red[lane * 33 + f], dwacc[f]);
    atomicAdd(&red[lane * 33 + 32], dbacc);
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) {
      int o = i / 33, f = i - o * 33;
      float val = red[i];
      if (val == 0.f) continue;
      if (o < kHD) {
        if (f < kF) { if (P.dWq) atomicAdd(&P.dWq[o * kF + f], val); }
        else if (P.dbq) atomicAdd(&P.dbq[o], val);
      } else {
        if (f < kF) { if (P.dWk) atomicAdd(&P.dWk[(o - kHD) * kF + f], val); }
        else if (P.dbk) atomicAdd(&P.dbk[o - kHD], val);
      }
    }
  }
}

inline size_t fwd_smem_bytes() { return (size_t)kWarps * kStages * kTileBytes + kWarps * kStages * 8; }
inline size_t bwd_smem_bytes() {
  return (size_t)kWarps * kStages * kTileBytes + kWarps * kStages * 8 + kWarps * 32 * 4 + 32 * 33 * 4;
}

}  // namespace fast
}  // namespace spt
