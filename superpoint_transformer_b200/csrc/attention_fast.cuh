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
constexpr int kTile = 8;            // CSR slots per tile
constexpr int kWarps = 4;           // warps per CTA (45 KB smem/CTA -> 5 CTAs/SM)
constexpr int kStages = 2;          // tiles in flight per warp
constexpr int kSlotBytes = kC * 4 + kHD * 4;   // one gathered V row (512 B) + K row (64 B)
constexpr int kStageBytes = kTile * kF * 4 + kTile * kSlotBytes;   // a tile + gathered rows
constexpr int kWarpSmem = kStages * kStageBytes;                   // per-warp bytes

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

// Per-warp tile stream over a contiguous slab [e0, e1) of CSR slots.  One tile =
// kTile slots; for every tile the warp stages, kStages tiles ahead of use:
//   * the edge features a[slot, :F]  — ONE 1-D TMA bulk copy (slots are contiguous),
//     completion on a warp-private mbarrier (initialised once; tiles are numbered
//     globally so the stage / phase-parity sequence continues across slabs);
//   * the gathered rows k[col[slot]] (64 B) and v[col[slot]] (512 B) — cp.async (LDGSTS),
//     issued for the whole tile at once: the tile's col[] is fetched with one coalesced
//     load and broadcast by shuffle, every lane copies (and later reads back) its own
//     16-byte / 4-byte piece of each row, one commit group per tile.
struct TileStream {
  const float* a;           // [E, kF]
  const int32_t* col;
  const float* kbase;       // P.k + (lane - kHD)   (dereferenced by k lanes only)
  const float* vbase;       // P.v + 4 * lane
  unsigned ldk, ldv;
  unsigned char* smem;      // warp-private: kStages * kStageBytes
  uint64_t* bar;            // kStages mbarriers
  int lane;
  bool is_k;
  int64_t e0, e1, tile_base;
  int gtile, gbase, gnext;
  int left;                 // slots not yet consumed in the readable tile
  const unsigned char* cur; // stage-relative cursor: a row at cur, gathered rows at cur_g
  const unsigned char* cur_g;

  __device__ __forceinline__ void init() {
    gnext = 0;
    if (lane == 0) {
#pragma unroll
      for (int s = 0; s < kStages; ++s) mbar_init(&bar[s], 1);
      mbar_fence_init();
    }
    __syncwarp();
  }
  __device__ __forceinline__ void issue(int k) {   // slab-local tile k
    const int64_t b = e0 + (int64_t)k * kTile;
    if (b < e1) {
      const int n = (int)min((int64_t)kTile, e1 - b);
      unsigned char* st = smem + ((gbase + k) % kStages) * kStageBytes;
      if (lane == 0) {
        uint64_t* br = &bar[(gbase + k) % kStages];
        mbar_expect_tx(br, (uint32_t)n * kF * 4);
        tma_load_1d(st, a + b * kF, (uint32_t)n * kF * 4, br);
      }
      const int c = (lane < n) ? col[b + lane] : 0;
      unsigned char* g = st + kTile * kF * 4;
#pragma unroll
      for (int e = 0; e < kTile; ++e) {
        if (e < n) {
          const unsigned t = (unsigned)__shfl_sync(kFull, c, e);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(
                           smem_u32(g + e * kSlotBytes + 16 * lane)),
                       "l"(vbase + (size_t)(t * ldv))
                       : "memory");
          if (is_k)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(
                             smem_u32(g + e * kSlotBytes + kC * 4 + 4 * (lane - kHD))),
                         "l"(kbase + (size_t)(t * ldk))
                         : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // (possibly empty) group per tile
  }
  __device__ __forceinline__ void wait_current() {
    mbar_wait(&bar[gtile % kStages], (uint32_t)((gtile / kStages) & 1), gtile, gbase);
    asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");
    left = (int)min((int64_t)kTile, e1 - tile_base);
    cur = smem + (gtile % kStages) * kStageBytes;
    cur_g = cur + kTile * kF * 4;
  }
  __device__ __forceinline__ void open(int64_t e0_, int64_t e1_) {
    e0 = e0_; e1 = e1_;
    left = 0;
    if (e1 <= e0) return;
    __syncwarp();
    gbase = gnext;
    gnext = gbase + (int)((e1 - e0 + kTile - 1) / kTile);
    gtile = gbase;
    tile_base = e0;
#pragma unroll
    for (int k = 0; k < kStages; ++k) issue(k);
    wait_current();
    cur -= kF * 4;            // next() pre-increments
    cur_g -= kSlotBytes;
  }
  // advance to the next CSR slot (strictly in order, every slot exactly once)
  __device__ __forceinline__ void next() {
    if (left == 0) {
      __syncwarp();                           // every lane is done with the old tile
      issue(gtile - gbase + kStages);         // refill the stage just released
      ++gtile;
      tile_base += kTile;
      wait_current();
    } else {
      cur += kF * 4;
      cur_g += kSlotBytes;
    }
    --left;
  }
  __device__ __forceinline__ const float* a_row() const {
    return reinterpret_cast<const float*>(cur);
  }
  __device__ __forceinline__ float4 v() const {
    return *reinterpret_cast<const float4*>(cur_g + 16 * lane);
  }
  __device__ __forceinline__ ulonglong2 v2() const {   // same 16 bytes as two fp32x2 pairs
    return *reinterpret_cast<const ulonglong2*>(cur_g + 16 * lane);
  }
  __device__ __forceinline__ float k() const {
    return *reinterpret_cast<const float*>(cur_g + kC * 4 + 4 * (lane - kHD));
  }
};

struct FwdArgs {
  const float* q; int ldq;
  const float* k; int ldk;
  const float* v; int ldv;
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

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
// 2^x on the SFU (MUFU.EX2, 2 ulp): the softmax is evaluated in base 2 with the
// logits pre-multiplied by log2(e); ex2(-inf) = 0 covers the first edge of a row.
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed fp32x2 math (Blackwell FFMA2 / FMUL2: two fp32 lanes per instruction) ----
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ float hsum2(f32x2 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  return lo + hi;
}
// acc += a * b
__device__ __forceinline__ void fma2(f32x2& acc, f32x2 a, f32x2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
// acc = acc * s + t
__device__ __forceinline__ void scale_add2(f32x2& acc, f32x2 s, f32x2 t) {
  asm("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc) : "l"(s), "l"(t));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// r_o = b_o + sum_f w[f] * a[f]: the weight row is held as 16 packed pairs, a_e is read
// as 8 broadcast LDS.128 (= 2 packed pairs each) -> 16 FFMA2 on two independent chains
__device__ __forceinline__ float gemv32(const f32x2 (&w2)[kF / 2], float bias, const float* arow) {
  f32x2 s01 = pack2(bias, 0.f), s23 = 0ull;
#pragma unroll
  for (int c = 0; c < kF / 4; ++c) {
    const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(arow + 4 * c);
    fma2(s01, w2[2 * c], t.x);
    fma2(s23, w2[2 * c + 1], t.y);
  }
  return hsum2(s01) + hsum2(s23);
}

__global__ void __launch_bounds__(kWarps * kWarp, 5)
k_attn_fwd_fast(FwdArgs P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* mine = smem_raw + (size_t)w * kWarpSmem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kWarps * kWarpSmem) + w * kStages;

  const int64_t gw = (int64_t)blockIdx.x * kWarps + w;
  const int64_t row0 = gw * P.rows_per_warp;
  if (row0 >= P.num_rows) return;
  const int64_t row1 = min(row0 + (int64_t)P.rows_per_warp, P.num_rows);

  // lane o: row o of [Wq; Wk] (16 packed pairs) and its bias
  f32x2 wreg[kF / 2];
  float bias = 0.f;
  {
    const float* W = (lane < kHD) ? P.Wq : P.Wk;
    const float* B = (lane < kHD) ? P.bq : P.bk;
    int o = lane & (kHD - 1);
#pragma unroll
    for (int f = 0; f < kF / 2; ++f)
      wreg[f] = W ? pack2(W[o * kF + 2 * f], W[o * kF + 2 * f + 1]) : 0ull;
    if (W && B) bias = B[o];
  }

  const bool is_k = lane >= kHD;
  const int hsel = (lane >> 3) << 2;           // lane holding the logit of my head
  const int aoff = 4 * (lane & 7);             // my 4 features of a_e (abar)
  const bool want_abar = P.abar != nullptr;
  const int64_t e_begin = P.rowptr[row0], e_end = P.rowptr[row1];

  TileStream ts;
  ts.a = P.a; ts.col = P.col;
  ts.kbase = P.k + (lane & (kHD - 1));
  ts.vbase = P.v + 4 * lane;
  ts.ldk = (unsigned)P.ldk; ts.ldv = (unsigned)P.ldv;
  ts.smem = mine; ts.bar = bars; ts.lane = lane; ts.is_k = is_k;
  ts.init();
  ts.open(e_begin, e_end);

  int b = P.rowptr[row0];
  for (int64_t row = row0; row < row1; ++row) {
    const int e = P.rowptr[row + 1];
    const float scale = qk_scale_fast(P.scale_mode, P.scale_value, e - b);
    const float qs = is_k ? 0.f : P.q[row * P.ldq + lane] * scale;
    float m_run = -INFINITY, l_run = 0.f;      // base-2 running max / sum of my head
    f32x2 accv01 = 0ull, accv23 = 0ull, acca01 = 0ull, acca23 = 0ull;

    for (int j = b; j < e; ++j) {
      ts.next();
      const float* arow = ts.a_row();
      const float r = gemv32(wreg, bias, arow);
      const float k_cur = is_k ? ts.k() : 0.f;
      const ulonglong2 v_cur = ts.v2();
      const float val = (is_k ? k_cur : qs) + r;            // q_e (lanes<16) | k_e
      float prod = val * __shfl_xor_sync(kFull, val, 16);
      prod += __shfl_xor_sync(kFull, prod, 1);
      prod += __shfl_xor_sync(kFull, prod, 2);              // <q_e,k_e>_h in lanes 4h..4h+3
      const float c2 = __shfl_sync(kFull, prod, hsel) * kLog2e;   // logit of my head, base 2
      const float m_new = fmaxf(m_run, c2);
      const float alpha = ex2(m_run - m_new);
      const float p = ex2(c2 - m_new);
      l_run = fmaf(l_run, alpha, p);
      m_run = m_new;
      const f32x2 pp = pack2(p, p), aa = pack2(alpha, alpha);
      scale_add2(accv01, aa, mul2(pp, v_cur.x));            // acc = acc*alpha + p*v
      scale_add2(accv23, aa, mul2(pp, v_cur.y));
      if (want_abar) {
        const ulonglong2 a4 = *reinterpret_cast<const ulonglong2*>(arow + aoff);
        scale_add2(acca01, aa, mul2(pp, a4.x));
        scale_add2(acca23, aa, mul2(pp, a4.y));
      }
    }
    const float zden = l_run + 1e-16f;
    const float inv = 1.f / zden;
    {
      const f32x2 ii = pack2(inv, inv);
      ulonglong2 o;
      o.x = mul2(accv01, ii); o.y = mul2(accv23, ii);
      *reinterpret_cast<ulonglong2*>(P.agg_v + row * kC + 4 * lane) = o;
      if (want_abar) {
        o.x = mul2(acca01, ii); o.y = mul2(acca23, ii);
        *reinterpret_cast<ulonglong2*>(P.abar + row * (kH * kF) + 4 * lane) = o;
      }
    }
    if ((lane & 7) == 0) {
      const int h = lane >> 3;
      P.m[row * kH + h] = (e > b) ? m_run * kLn2 : 0.f;    // natural-log units
      P.z[row * kH + h] = zden;
      P.sump[row * kH + h] = l_run * inv;
    }
    b = e;
  }
}

// ------------------------------------------------------------------ backward rows
struct BwdArgs {
  const float* q; int ldq;
  const float* k; int ldk;
  const float* v; int ldv;
  const float* a;
  const int32_t* rowptr; const int32_t* col;
  int64_t num_rows;
  const float* Wq; const float* bq; const float* Wk; const float* bk;
  int scale_mode; float scale_value;
  const float* m; const float* z;
  const float* agg_v; const float* abar;
  const float* d_agg_v; const float* d_abar;
  float* dq; int lddq;
  float* da;
  float* Pbuf;   // [E, H]
  float* G;      // [E, 2HD] = [dq_e | dk_e]
  int rows_per_warp;
};

// smem per warp: stream (kStages tiles) + g[32]
__global__ void __launch_bounds__(kWarps * kWarp, 4)
k_attn_bwd_rows_fast(BwdArgs P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* mine = smem_raw + (size_t)w * kWarpSmem;
  unsigned char* after = smem_raw + (size_t)kWarps * kWarpSmem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(after) + w * kStages;
  float* g_s = reinterpret_cast<float*>(after + kWarps * kStages * 8) + w * 32;

  const int64_t gw = (int64_t)blockIdx.x * kWarps + w;
  const int64_t row0 = gw * P.rows_per_warp;
  if (row0 >= P.num_rows) return;
  const int64_t row1 = min(row0 + (int64_t)P.rows_per_warp, P.num_rows);

  // lane o: row o of [Wq;Wk] (forward GEMV) ; lane f: column f (da GEMV); packed pairs
  f32x2 wrow[kF / 2], wcol[kHD];
  float bias = 0.f;
  {
    const float* W = (lane < kHD) ? P.Wq : P.Wk;
    const float* B = (lane < kHD) ? P.bq : P.bk;
    int o = lane & (kHD - 1);
#pragma unroll
    for (int f = 0; f < kF / 2; ++f)
      wrow[f] = W ? pack2(W[o * kF + 2 * f], W[o * kF + 2 * f + 1]) : 0ull;
    if (W && B) bias = B[o];
#pragma unroll
    for (int oo = 0; oo < kHD / 2; ++oo) {
      wcol[oo] = P.Wq ? pack2(P.Wq[(2 * oo) * kF + lane], P.Wq[(2 * oo + 1) * kF + lane]) : 0ull;
      wcol[kHD / 2 + oo] =
          P.Wk ? pack2(P.Wk[(2 * oo) * kF + lane], P.Wk[(2 * oo + 1) * kF + lane]) : 0ull;
    }
  }

  const bool is_k = lane >= kHD;
  const int hsel = (lane >> 3) << 2;
  const int myhead = lane >> 3;
  const int aoff = 4 * (lane & 7);
  const int dcsel = ((lane & 15) >> 2) << 3;     // lane group holding dc of head(o)
  const bool has_dab = P.d_abar != nullptr && P.abar != nullptr;
  const bool want_da = P.da != nullptr;
  const int64_t e_begin = P.rowptr[row0], e_end = P.rowptr[row1];

  TileStream ts;
  ts.a = P.a; ts.col = P.col;
  ts.kbase = P.k + (lane & (kHD - 1));
  ts.vbase = P.v + 4 * lane;
  ts.ldk = (unsigned)P.ldk; ts.ldv = (unsigned)P.ldv;
  ts.smem = mine; ts.bar = bars; ts.lane = lane; ts.is_k = is_k;
  ts.init();
  ts.open(e_begin, e_end);

  int b = P.rowptr[row0];
  for (int64_t row = row0; row < row1; ++row) {
    const int e = P.rowptr[row + 1];
    const float scale = qk_scale_fast(P.scale_mode, P.scale_value, e - b);
    const float qs = is_k ? 0.f : P.q[row * P.ldq + lane] * scale;
    const float m2 = P.m[row * kH + myhead] * kLog2e;
    const float zinv = 1.f / P.z[row * kH + myhead];
    const float4 dy = *reinterpret_cast<const float4*>(P.d_agg_v + row * kC + 4 * lane);
    float4 dab = make_float4(0.f, 0.f, 0.f, 0.f);   // d_abar[row][myhead][4(l%8)..]
    float dabf[kH] = {0.f, 0.f, 0.f, 0.f};          // d_abar[row][h][lane]   (lane = f)
    // delta_h = <dY_h, agg_v_h> + <dAbar_h, abar_h>   (= sum_e p_e dp_e)
    float delta;
    {
      const float4 ag = *reinterpret_cast<const float4*>(P.agg_v + row * kC + 4 * lane);
      float part = dy.x * ag.x + dy.y * ag.y + dy.z * ag.z + dy.w * ag.w;
      if (has_dab) {
        dab = *reinterpret_cast<const float4*>(P.d_abar + row * (kH * kF) + 4 * lane);
        const float4 ab = *reinterpret_cast<const float4*>(P.abar + row * (kH * kF) + 4 * lane);
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

    for (int j = b; j < e; ++j) {
      ts.next();
      const float* arow = ts.a_row();
      const float r = gemv32(wrow, bias, arow);
      const float k_cur = is_k ? ts.k() : 0.f;
      const float4 v_cur = ts.v();
      const float val = (is_k ? k_cur : qs) + r;
      const float other = __shfl_xor_sync(kFull, val, 16);    // k_e for q lanes, q_e for k lanes
      float prod = val * other;
      prod += __shfl_xor_sync(kFull, prod, 1);
      prod += __shfl_xor_sync(kFull, prod, 2);
      const float c2 = __shfl_sync(kFull, prod, hsel) * kLog2e;
      const float p = ex2(c2 - m2) * zinv;                     // softmax weight of my head
      // dp_h = <dY_h, v_h> + <dAbar_h, a>
      float part = dy.x * v_cur.x + dy.y * v_cur.y + dy.z * v_cur.z + dy.w * v_cur.w;
      if (has_dab) {
        const float4 a4 = *reinterpret_cast<const float4*>(arow + aoff);
        part += dab.x * a4.x + dab.y * a4.y + dab.z * a4.z + dab.w * a4.w;
      }
      part += __shfl_xor_sync(kFull, part, 1);
      part += __shfl_xor_sync(kFull, part, 2);
      part += __shfl_xor_sync(kFull, part, 4);
      const float dc = p * (part - delta);                     // d logit of my head
      // g_o = dc_{h(o)} * other ; h(o) = (o % 16) / 4 lives in lanes 8*h(o)..
      const float g = __shfl_sync(kFull, dc, dcsel) * other;   // dq_e (lanes<16) | dk_e
      if (!is_k) dq_acc += g;
      P.G[(size_t)j * (2 * kHD) + lane] = g;
      if ((lane & 7) == 0) P.Pbuf[(size_t)j * kH + myhead] = p;
      if (want_da) {
        g_s[lane] = g;
        __syncwarp();
        f32x2 s01 = 0ull, s23 = 0ull;
        if (has_dab) {
          // sum_h p_h * dAbar[row][h][f=lane]
          s01 = pack2(__shfl_sync(kFull, p, 0) * dabf[0], __shfl_sync(kFull, p, 8) * dabf[1]);
          s23 = pack2(__shfl_sync(kFull, p, 16) * dabf[2], __shfl_sync(kFull, p, 24) * dabf[3]);
        }
        // sum_o W[o][f] g_o : g broadcast from shared memory as packed pairs (o, o+1)
#pragma unroll
        for (int c4 = 0; c4 < (2 * kHD) / 4; ++c4) {
          const ulonglong2 gg = *reinterpret_cast<const ulonglong2*>(g_s + 4 * c4);
          fma2(s01, wcol[2 * c4], gg.x);
          fma2(s23, wcol[2 * c4 + 1], gg.y);
        }
        P.da[(size_t)j * kF + lane] = hsum2(s01) + hsum2(s23);
        __syncwarp();
      }
    }
    if (!is_k) P.dq[row * P.lddq + lane] = dq_acc * scale;
    b = e;
  }
}

// ------------------------------------------------------------------ backward targets
// warp per target t: dv[t] = sum_in p * dY[src] (4 channels per lane),
// dk[t] = sum_in dk_e (lanes 16-31 read the dk_e half of G, coalesced 64 B).
struct TgtArgs {
  const int32_t* csc_ptr; const int32_t* csc_src; const int32_t* csc2csr;
  int64_t num_targets;
  const float* Pbuf; const float* G; const float* d_agg_v;
  float* dk; int lddk; float* dv; int lddv;
};

// (Round 2 tried 4 edges in flight + L2 policies (dY evict-last, P/G evict-first): 0.164 ms vs
// 0.155 ms for this version at E = 1.7 M — reverted.)
__global__ void __launch_bounds__(256)
k_attn_bwd_targets_fast(TgtArgs P) {
  const int lane = threadIdx.x & 31;
  const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (t >= P.num_targets) return;
  const int b = P.csc_ptr[t], e = P.csc_ptr[t + 1];
  const int myhead = lane >> 3;
  const bool is_k = lane >= kHD;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float dk = 0.f;
  for (int base = b; base < e; base += 32) {
    // coalesced fetch of up to 32 (slot, source) pairs, then broadcast by shuffle
    const int n = min(32, e - base);
    int jj = 0, ss = 0;
    if (lane < n) {
      jj = P.csc2csr[base + lane];
      ss = P.csc_src[base + lane];
    }
    int i = 0;
    for (; i + 1 < n; i += 2) {   // two edges in flight
      const unsigned j0 = (unsigned)__shfl_sync(kFull, jj, i);
      const unsigned j1 = (unsigned)__shfl_sync(kFull, jj, i + 1);
      const unsigned s0 = (unsigned)__shfl_sync(kFull, ss, i);
      const unsigned s1 = (unsigned)__shfl_sync(kFull, ss, i + 1);
      const float p0 = P.Pbuf[(size_t)j0 * kH + myhead];
      const float p1 = P.Pbuf[(size_t)j1 * kH + myhead];
      const float4 y0 = *reinterpret_cast<const float4*>(P.d_agg_v + (size_t)s0 * kC + 4 * lane);
      const float4 y1 = *reinterpret_cast<const float4*>(P.d_agg_v + (size_t)s1 * kC + 4 * lane);
      float g0 = 0.f, g1 = 0.f;
      if (is_k) {
        g0 = P.G[(size_t)j0 * (2 * kHD) + lane];
        g1 = P.G[(size_t)j1 * (2 * kHD) + lane];
      }
      acc.x = fmaf(p0, y0.x, acc.x); acc.y = fmaf(p0, y0.y, acc.y);
      acc.z = fmaf(p0, y0.z, acc.z); acc.w = fmaf(p0, y0.w, acc.w);
      acc.x = fmaf(p1, y1.x, acc.x); acc.y = fmaf(p1, y1.y, acc.y);
      acc.z = fmaf(p1, y1.z, acc.z); acc.w = fmaf(p1, y1.w, acc.w);
      dk += g0;
      dk += g1;
    }
    if (i < n) {
      const unsigned j0 = (unsigned)__shfl_sync(kFull, jj, i);
      const unsigned s0 = (unsigned)__shfl_sync(kFull, ss, i);
      const float p0 = P.Pbuf[(size_t)j0 * kH + myhead];
      const float4 y0 = *reinterpret_cast<const float4*>(P.d_agg_v + (size_t)s0 * kC + 4 * lane);
      if (is_k) dk += P.G[(size_t)j0 * (2 * kHD) + lane];
      acc.x = fmaf(p0, y0.x, acc.x); acc.y = fmaf(p0, y0.y, acc.y);
      acc.z = fmaf(p0, y0.z, acc.z); acc.w = fmaf(p0, y0.w, acc.w);
    }
  }
  *reinterpret_cast<float4*>(P.dv + t * P.lddv + 4 * lane) = acc;
  if (is_k) P.dk[t * P.lddk + (lane - kHD)] = dk;
}

// ------------------------------------------------------------------ backward weights
// d[Wq;Wk] (32x32) += G^T a over an edge slab per CTA, db += colsum(G).  Each warp
// takes every 8th edge of the slab; lane (oi = lane/4, fi = lane%4) owns the
// register tile o in {4oi..4oi+3} x f in {8fi..8fi+7} (3 LDG.128 : 32 FFMA).
struct DwArgs {
  const float* G; const float* a; int64_t E;
  float* dWq; float* dbq; float* dWk; float* dbk;
  int64_t edges_per_cta;
};

__global__ void __launch_bounds__(256)
k_attn_bwd_dw_fast(DwArgs P) {
  __shared__ float red[32 * 33];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int oi = lane >> 2, fi = lane & 3;
  f32x2 acc[4][4];            // [o][f pair]: o = 4oi + i, f = 8fi + 2jp, +1
  float accb[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0ull;
  const int64_t e0 = (int64_t)blockIdx.x * P.edges_per_cta;
  const int64_t e1 = min(e0 + P.edges_per_cta, P.E);
  // 4 edges in flight per warp iteration (12 independent 16-byte streaming loads)
  for (int64_t j0 = e0 + 4 * w; j0 < e1; j0 += 32) {
    float4 g[4];
    ulonglong2 a0[4], a1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t j = j0 + u;
      if (j < e1) {
        g[u] = ldg_stream4(P.G + j * 32 + 4 * oi);
        const float4 t0 = ldg_stream4(P.a + j * 32 + 8 * fi);
        const float4 t1 = ldg_stream4(P.a + j * 32 + 8 * fi + 4);
        a0[u].x = pack2(t0.x, t0.y); a0[u].y = pack2(t0.z, t0.w);
        a1[u].x = pack2(t1.x, t1.y); a1[u].y = pack2(t1.z, t1.w);
      } else {
        g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        a0[u].x = a0[u].y = a1[u].x = a1[u].y = 0ull;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float gv[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const f32x2 gg = pack2(gv[i], gv[i]);
        fma2(acc[i][0], gg, a0[u].x);
        fma2(acc[i][1], gg, a0[u].y);
        fma2(acc[i][2], gg, a1[u].x);
        fma2(acc[i][3], gg, a1[u].y);
        accb[i] += gv[i];
      }
    }
  }
  for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int jp = 0; jp < 4; ++jp) {
      float lo, hi;
      unpack2(acc[i][jp], lo, hi);
      atomicAdd(&red[(4 * oi + i) * 33 + 8 * fi + 2 * jp], lo);
      atomicAdd(&red[(4 * oi + i) * 33 + 8 * fi + 2 * jp + 1], hi);
    }
    if (fi == 0) atomicAdd(&red[(4 * oi + i) * 33 + 32], accb[i]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * 33; i += blockDim.x) {
    const int o = i / 33, f = i - o * 33;
    const float val = red[i];
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

inline size_t fwd_smem_bytes() { return (size_t)kWarps * kWarpSmem + kWarps * kStages * 8; }
inline size_t bwd_smem_bytes() {
  return (size_t)kWarps * kWarpSmem + kWarps * kStages * 8 + kWarps * 32 * 4;
}

}  // namespace fast
}  // namespace spt
