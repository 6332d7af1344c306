// attention_tile.cuh — row-tile attention kernels for the SPT "4 heads" family
// (H*D = 16 -> 2*H*D = 32 RPE outputs, F = 32 edge features, C = H*Dv = 128):
// BASELINE.json cfg 2/3/5 (C=128, H=4, qk_dim=4, in_rpe_dim=32).
//
// Replaces the per-edge dependent chain of attention_fast.cuh (144 / 242 warp
// instructions per edge) by a per-ROW-TILE schedule.  One warp owns a block of
// consecutive CSR rows; one row = one tile of <= 32 CSR slots (longer rows take
// several tiles, merged with an online softmax):
//
//   A  the tile's edge features a[tb : tb+n, 0:32] arrive as ONE 2-D TMA copy (four tensor
//      maps with boxes of 8 / 16 / 24 / 32 rows, SWIZZLE_128B, L2 evict-first) at the start
//      of a warp-private 4 KB stage; two stages per warp, the next row's tile is in flight
//      while the current one is consumed (struct TilePipe);
//   B  the RPE product of the whole tile, R[32 x 32] = A_tile [32 x 32] . [Wq;Wk]^T, runs on
//      the tensor cores: mma.sync m16n8k8 TF32 with the 3xTF32 split (fp32-accurate),
//      A fragments by ldmatrix from the swizzled tile (conflict-free), weight fragments
//      pre-split (hi/lo) once per CTA in fragment order;
//   C  in the accumulator fragment a thread holds r_q[h,d] and r_k[h,d] of the same
//      (edge, head, d in {2t,2t+1}): head logit = 2 FMAs + one shfl.xor(1);
//      two-pass softmax over the tile in registers (3 shfl.xor over the row lanes);
//   D  p goes through a [32 x 4] shared-memory tile of (p, p) pairs to the accumulation
//      layout (lane = 4 value channels + 4 abar entries): per edge one LDG.128 of the gathered
//      v row (L2 evict-last, 8 in flight, no predicates: n = 8 + ... + 4 + 2 + 1), one LDS.64
//      of (p, p), one LDS.128 of the staged feature row and 4 packed FMAs.
//
// The backward rows kernel has the same front end, recomputes p from the saved (m, z),
// evaluates dp = <dY, v> + <dAbar, a> with a butterfly transpose-reduce (7 shuffles per 8
// edges), forms G = [dq_e | dk_e] directly in the accumulator fragment, and feeds that
// fragment — with a permuted k order, no shuffles — as the A operand of the second
// tensor-core product da = G . [Wq;Wk] + P . dAbar.
//
// template <bool BF>: BF = true reads q / k / v / the features as bf16 (cfg 3): 64-byte feature
// rows (SWIZZLE_64B), single bf16 m16n8k16 MMAs with fp32 accumulation instead of the 3xTF32
// triple; the accumulator layout — and with it everything after the products — is shared.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "attention_fast.cuh"

namespace spt {
namespace tile {

using fast::smem_u32;
using fast::mbar_init;
using fast::mbar_fence_init;
using fast::mbar_expect_tx;
using fast::mbar_wait;
using fast::f32x2;
using fast::pack2;
using fast::unpack2;
using fast::fma2;
using fast::mul2;
using fast::ex2;
using fast::kLog2e;
using fast::kLn2;

constexpr int kH = 4, kD = 4, kDv = 32, kF = 32;
constexpr int kHD = kH * kD;          // 16
constexpr int kHD2 = 2 * kHD;         // 32
constexpr int kC = kH * kDv;          // 128
constexpr int kChunk = 32;            // CSR slots per tile stage = 4 TMA boxes of 8 rows
constexpr int kChunkBytes = kChunk * kF * 4;     // 4096
constexpr int kFragBytes = 16 * 32 * 16;         // one pre-split weight fragment set (8 KB)

// 2 CTAs per SM: forward 8 warps x 128 registers, backward 6 warps x 168 registers.  The
// kernels are bound by instruction latency (dependent ALU / shuffle / LDS chains per row), so
// resident warps per scheduler (4 / 3) are what hides it; shared memory per warp is 9-10 KB.
#ifndef SPT_FWD_WARPS
#define SPT_FWD_WARPS 8
#endif
#ifndef SPT_BWD_WARPS
#define SPT_BWD_WARPS 6
#endif
constexpr int kFwdWarps = SPT_FWD_WARPS;
constexpr int kBwdWarps = SPT_BWD_WARPS;

__host__ __device__ inline bool shape_ok(int H, int D, int Dv, int F) {
  return H == kH && D == kD && Dv == kDv && F == kF;
}

// one 2-D TMA box of the edge-feature matrix (rows `row`.. of the map's box height), L2
// evict-first: the features are streamed once per kernel and must not push the gathered node
// rows (k / v, re-read ~17 times) out of the L2
__device__ __forceinline__ void tma_box(void* dst, const CUtensorMap* tm, int row, uint64_t* bar,
                                        uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(0), "r"(row), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// gathered node rows: read-only path, no L1 allocation, L2 evict-last
__device__ __forceinline__ ulonglong2 ldg_row16(const char* p, uint64_t policy) {
  ulonglong2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;"
               : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ float2 ldg_row8(const char* p, uint64_t policy) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;"
               : "=f"(v.x), "=f"(v.y) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ uint32_t ldg_row4(const char* p, uint64_t policy) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;"
               : "=r"(v) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ uint2 ldg_row8u(const char* p, uint64_t policy) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.b32 {%0,%1}, [%2], %3;"
               : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(policy));
  return v;
}
// bf16 pair (low half = first element) -> two fp32
__device__ __forceinline__ float2 bf2_to_f2(uint32_t x) {
  return make_float2(__uint_as_float(x << 16), __uint_as_float(x & 0xffff0000u));
}
__device__ __forceinline__ f32x2 bf2_to_f32x2(uint32_t x) {
  return pack2(__uint_as_float(x << 16), __uint_as_float(x & 0xffff0000u));
}
// two fp32 -> packed bf16 pair (round to nearest even), low half = first element
__device__ __forceinline__ uint32_t f2_to_bf2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// the four tensor maps of the edge-feature matrix: boxes of 8 / 16 / 24 / 32 rows, so that a
// tile of any size is ONE TMA instruction
struct TileMaps {
  CUtensorMap m[4];
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                        uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}

// x = hi + lo, hi = x rounded to tf32 (integer round-to-nearest: full-rate ALU ops, the
// cvt.rna.tf32 instruction is quarter rate).  lo is passed as is: the tensor core reads the
// upper 19 bits of a tf32 operand, i.e. truncates lo at 2^-11 of an already 2^-11-small term.
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t& hi, uint32_t& lo) {
  hi = (x + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(__uint_as_float(x) - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += (ahi + alo) . (bhi + blo) without the lo.lo term (small terms first)
__device__ __forceinline__ void mma_3x(float (&c)[4], const uint32_t (&ahi)[4],
                                       const uint32_t (&alo)[4], const uint4& b) {
  mma_tf32(c, alo, b.x, b.y);
  mma_tf32(c, ahi, b.z, b.w);
  mma_tf32(c, ahi, b.x, b.y);
}

__device__ __forceinline__ float2 ldg_stream2(const float* p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];"
               : "=f"(r.x), "=f"(r.y)
               : "l"(p));
  return r;
}

// ---------------------------------------------------------------------------------------
// Warp-private double buffer of edge-feature tiles.  A tile = the <= 32 CSR slots
// [tb, tb + n) of one row; its rows arrive as ceil(n / 8) 2-D TMA boxes (8 rows x 128 B,
// SWIZZLE_128B) at the START of a 4 KB stage (1024-byte aligned), so the 16-byte chunk c of
// tile row r always sits at r * 128 + ((c ^ (r & 7)) << 4): every ldmatrix / LDS address of
// the consumers is a per-lane constant plus the stage base.  While tile i is consumed the
// TMA of tile i + 1 (normally the next row) is in flight in the other stage.  The boxes
// over-fetch up to 7 slots past the tile's end (other rows' features, L2-resident; slots
// past E are zero-filled by the TMA).
// ---------------------------------------------------------------------------------------
struct TilePipe {
  unsigned char* buf;       // 2 stages x 4 KB
  uint64_t* bar;            // 2 mbarriers
  const TileMaps* tm;
  uint64_t policy;
  uint32_t box_bytes;       // bytes of one 8-row box: 1024 (fp32 features) / 512 (bf16)
  int lane;
  uint32_t uses0, uses1;    // completed uses of each stage (phase parity)
  int staged_tb;            // first slot of the tile in flight in stage `next`, -1 if none
  int cur;                  // stage holding the current tile

  __device__ __forceinline__ void init() {
    if (lane == 0) {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      mbar_fence_init();
    }
    uses0 = uses1 = 0;
    staged_tb = -1;
    cur = 1;
    __syncwarp();
  }
  __device__ __forceinline__ void issue(int stage, int tb, int n) {
    if (lane == 0) {
      const int nb = (n + 7) >> 3;     // 1..4 boxes of 8 rows = one box of the nb-th map
      unsigned char* dst = buf + stage * kChunkBytes;
      mbar_expect_tx(&bar[stage], (uint32_t)nb * box_bytes);
      tma_box(dst, &tm->m[nb - 1], tb, &bar[stage], policy);
    }
  }
  // start fetching the tile that will be consumed after the current one
  __device__ __forceinline__ void prefetch(int tb, int n) {
    issue(cur ^ 1, tb, n);
    staged_tb = tb;
  }
  // make tile [tb, tb + n) current; returns the shared-memory address of its first row.
  // All lanes must have finished reading the stage that gets recycled (callers __syncwarp).
  __device__ __forceinline__ uint32_t acquire(int tb, int n) {
    const int st = cur ^ 1;
    if (staged_tb != tb) issue(st, tb, n);     // not prefetched (first tile / after empty rows)
    staged_tb = -1;
    cur = st;
    const uint32_t par = (st ? uses1 : uses0) & 1u;
    mbar_wait(&bar[st], par, tb, n);
    if (st) ++uses1; else ++uses0;
    __syncwarp();
    return smem_u32(buf + st * kChunkBytes);
  }
};

// weight fragments in shared memory.
//   frag1[(kk*4 + nn)*32 + lane] = {b0.hi, b1.hi, b0.lo, b1.lo} of  B1[k=f][n=o] = W[o][f]:
//       b0 = W[8nn + g][8kk + t], b1 = W[8nn + g][8kk + t + 4]           (GEMV 1: r = a W^T)
//   frag2[(ks*4 + nf)*32 + lane] of  B2[k=o][n=f] = W[o][f] with the k order of the
//       accumulator fragment (slot s < 4 <-> o = 8ks + 2s, slot 4 + s <-> o = 8ks + 2s + 1):
//       b0 = W[8ks + 2t][8nf + g], b1 = W[8ks + 2t + 1][8nf + g]          (GEMV 2: da = G W)
// W = [Wq; Wk] (rows 0-15 / 16-31); an absent encoder is a zero block.
__device__ __forceinline__ float w_at(const float* Wq, const float* Wk, int o, int f) {
  const float* W = (o < kHD) ? Wq : Wk;
  return W ? W[(o & (kHD - 1)) * kF + f] : 0.f;
}
__device__ __forceinline__ uint4 split_pair(float w0, float w1) {
  uint32_t h0, l0, h1, l1;
  split_tf32(__float_as_uint(w0), h0, l0);
  split_tf32(__float_as_uint(w1), h1, l1);
  // lo rounded as well (done once per CTA)
  l0 = (l0 + 0x1000u) & 0xffffe000u;
  l1 = (l1 + 0x1000u) & 0xffffe000u;
  return make_uint4(h0, h1, l0, l1);
}
__device__ __forceinline__ void build_frag1(uint4* frag, const float* Wq, const float* Wk) {
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) {
    const int ln = i & 31, fr = i >> 5, kk = fr >> 2, nn = fr & 3, g = ln >> 2, t = ln & 3;
    frag[i] = split_pair(w_at(Wq, Wk, 8 * nn + g, 8 * kk + t),
                         w_at(Wq, Wk, 8 * nn + g, 8 * kk + t + 4));
  }
}
__device__ __forceinline__ void build_frag2(uint4* frag, const float* Wq, const float* Wk) {
  for (int i = threadIdx.x; i < 16 * 32; i += blockDim.x) {
    const int ln = i & 31, fr = i >> 5, ks = fr >> 2, nf = fr & 3, g = ln >> 2, t = ln & 3;
    frag[i] = split_pair(w_at(Wq, Wk, 8 * ks + 2 * t, 8 * nf + g),
                         w_at(Wq, Wk, 8 * ks + 2 * t + 1, 8 * nf + g));
  }
}
__device__ __forceinline__ void build_bias(float* bias_s, const float* Wq, const float* bq,
                                           const float* Wk, const float* bk) {
  for (int o = threadIdx.x; o < kHD2; o += blockDim.x) {
    float b = 0.f;
    if (o < kHD) { if (Wq && bq) b = bq[o]; }
    else { if (Wk && bk) b = bk[o - kHD]; }
    bias_s[o] = b;
  }
}

// bf16 storage variant (cfg 3): the features arrive as bf16 [E, 32] (64-byte rows, SWIZZLE_64B:
// the 16-byte chunk c of tile row r sits at r*64 + ((c ^ ((r >> 1) & 3)) << 4)), the products run
// as single mma.sync.m16n8k16 bf16 with fp32 accumulation (no split), weights rounded to bf16
// once per CTA:
//   frag1b[(ks*4 + nn)*32 + lane] = {b0, b1}: b0 = (W[8nn+g][16ks+2t], W[8nn+g][16ks+2t+1]),
//                                            b1 = the same at k + 8            (r = a W^T)
//   frag2b[(ks*4 + nf)*32 + lane]: b0 = (W[16ks+2t][8nf+g], W[16ks+2t+1][8nf+g]), b1 at o + 8
//                                  (da = G W; the accumulator fragment IS the A fragment)
__device__ __forceinline__ void build_frag1_bf16(uint2* frag, const float* Wq, const float* Wk) {
  for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) {
    const int ln = i & 31, fr = i >> 5, ks = fr >> 2, nn = fr & 3, g = ln >> 2, t = ln & 3;
    const int o = 8 * nn + g, f = 16 * ks + 2 * t;
    frag[i] = make_uint2(f2_to_bf2(w_at(Wq, Wk, o, f), w_at(Wq, Wk, o, f + 1)),
                         f2_to_bf2(w_at(Wq, Wk, o, f + 8), w_at(Wq, Wk, o, f + 9)));
  }
}
__device__ __forceinline__ void build_frag2_bf16(uint2* frag, const float* Wq, const float* Wk) {
  for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) {
    const int ln = i & 31, fr = i >> 5, ks = fr >> 2, nf = fr & 3, g = ln >> 2, t = ln & 3;
    const int o = 16 * ks + 2 * t, f = 8 * nf + g;
    frag[i] = make_uint2(f2_to_bf2(w_at(Wq, Wk, o, f), w_at(Wq, Wk, o + 1, f)),
                         f2_to_bf2(w_at(Wq, Wk, o + 8, f), w_at(Wq, Wk, o + 9, f)));
  }
}
__device__ __forceinline__ void rpe_tile_bf16(float (&acc)[2][4][4], uint32_t tile, bool two,
                                              const uint2* frag1, const float* bias_s, int lane) {
  const int t = lane & 3;
#pragma unroll
  for (int nn = 0; nn < 4; ++nn) {
    const float2 b = *reinterpret_cast<const float2*>(bias_s + 8 * nn + 2 * t);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      acc[m][nn][0] = b.x; acc[m][nn][1] = b.y; acc[m][nn][2] = b.x; acc[m][nn][3] = b.y;
    }
  }
  // ldmatrix: lane L feeds tile row 16m + (L&7) + 8*((L>>3)&1), 16-byte chunk 2ks + (L>>4)
  const int r0 = (lane & 7) + ((lane >> 3) & 1) * 8;
  const uint32_t row0 = tile + (uint32_t)(r0 * 64);
  const uint32_t rsw = (uint32_t)((r0 >> 1) & 3), csel = (uint32_t)(lane >> 4);
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const uint32_t coff = ((2 * ks + csel) ^ rsw) << 4;   // (r0 + 16) >> 1 & 3 == rsw as well
    uint2 b[4];
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) b[nn] = frag1[(ks * 4 + nn) * 32 + lane];
    {
      uint32_t a[4];
      ldsm_x4(row0 + coff, a[0], a[1], a[2], a[3]);
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) mma_bf16(acc[0][nn], a, b[nn].x, b[nn].y);
    }
    if (two) {
      uint32_t a[4];
      ldsm_x4(row0 + 16 * 64 + coff, a[0], a[1], a[2], a[3]);
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) mma_bf16(acc[1][nn], a, b[nn].x, b[nn].y);
    }
  }
}

// R = A_tile . W^T + bias for the tile at shared-memory address `tile`:
// acc[m][nn][.] = accumulator fragments (m-tile m = edges 16m..16m+15, n-tile nn = outputs
// 8nn..8nn+7): c0,c1 = (edge 16m+g, outputs 8nn+2t, +1), c2,c3 = (edge 16m+8+g, same outputs).
__device__ __forceinline__ void rpe_tile(float (&acc)[2][4][4], uint32_t tile, bool two,
                                         const uint4* frag1, const float* bias_s, int lane) {
  const int t = lane & 3;
#pragma unroll
  for (int nn = 0; nn < 4; ++nn) {
    const float2 b = *reinterpret_cast<const float2*>(bias_s + 8 * nn + 2 * t);
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      acc[m][nn][0] = b.x; acc[m][nn][1] = b.y; acc[m][nn][2] = b.x; acc[m][nn][3] = b.y;
    }
  }
  // ldmatrix: lane L feeds tile row 16m + (L&7) + 8*((L>>3)&1) of matrix L>>3; (row & 7) = L & 7
  const uint32_t row0 = tile + (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * 128);
  const uint32_t rsw = (uint32_t)(lane & 7), csel = (uint32_t)(lane >> 4);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const uint32_t coff = ((2 * kk + csel) ^ rsw) << 4;
    uint4 b[4];
#pragma unroll
    for (int nn = 0; nn < 4; ++nn) b[nn] = frag1[(kk * 4 + nn) * 32 + lane];
    {
      uint32_t a[4], ahi[4], alo[4];
      ldsm_x4(row0 + coff, a[0], a[1], a[2], a[3]);
#pragma unroll
      for (int i = 0; i < 4; ++i) split_tf32(a[i], ahi[i], alo[i]);
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) mma_3x(acc[0][nn], ahi, alo, b[nn]);
    }
    if (two) {   // warp-uniform: rows 16..31 of the tile hold edges
      uint32_t a[4], ahi[4], alo[4];
      ldsm_x4(row0 + 16 * 128 + coff, a[0], a[1], a[2], a[3]);
#pragma unroll
      for (int i = 0; i < 4; ++i) split_tf32(a[i], ahi[i], alo[i]);
#pragma unroll
      for (int nn = 0; nn < 4; ++nn) mma_3x(acc[1][nn], ahi, alo, b[nn]);
    }
  }
}

__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// my 4 features 4*(lane&7).. of tile row r as two packed fp32 pairs.
//   fp32 tiles: 16-byte chunk (lane&7) of the 128-byte row, at ((l7 ^ (r&7)) << 4)
//   bf16 tiles: 8 bytes in the 64-byte row: chunk l7>>1 at ((l7>>1) ^ ((r>>1)&3)) << 4, half l7&1
template <bool BF, bool ALIGNED>
__device__ __forceinline__ ulonglong2 lds_a_chunk(uint32_t tile, int e0, int u, uint32_t l7s) {
  ulonglong2 a4;
  if (BF) {
    const uint32_t l7 = l7s >> 4;
    const uint32_t r = (uint32_t)(e0 + u);
    const uint32_t sw = ALIGNED ? (uint32_t)((u >> 1) & 3) : ((r >> 1) & 3u);
    const uint32_t ad = tile + r * 64u + ((((l7 >> 1) ^ sw)) << 4) + ((l7 & 1u) << 3);
    uint2 w;
    asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(w.x), "=r"(w.y) : "r"(ad));
    a4.x = bf2_to_f32x2(w.x);
    a4.y = bf2_to_f32x2(w.y);
    return a4;
  }
  uint32_t ad;
  if (ALIGNED) {   // e0 % 8 == 0: (e0 + u) & 7 == u, a compile-time constant
    ad = tile + (uint32_t)(e0 * 128) + (uint32_t)(u * 128) + (l7s ^ (uint32_t)((u & 7) << 4));
  } else {
    const uint32_t r = (uint32_t)(e0 + u);
    ad = tile + r * 128u + (l7s ^ ((r & 7u) << 4));
  }
  asm volatile("ld.shared.v2.u64 {%0,%1}, [%2];" : "=l"(a4.x), "=l"(a4.y) : "r"(ad));
  return a4;
}

// my 4 value channels of the gathered row `tc` as two packed fp32 pairs
template <bool BF>
__device__ __forceinline__ ulonglong2 gather_v(const char* vbase, unsigned tc, unsigned ldvb,
                                               uint64_t keep) {
  if (BF) {
    const uint2 w = ldg_row8u(vbase + (uint64_t)tc * (uint64_t)ldvb, keep);
    ulonglong2 v;
    v.x = bf2_to_f32x2(w.x);
    v.y = bf2_to_f32x2(w.y);
    return v;
  }
  return ldg_row16(vbase + (uint64_t)tc * (uint64_t)ldvb, keep);
}
// k[tc][2t, 2t+1] and k[tc][8+2t, 8+2t+1]
template <bool BF>
__device__ __forceinline__ void gather_k(const char* kbase, unsigned tc, unsigned ldkb,
                                         uint64_t keep, float2& kA, float2& kB) {
  const char* kp = kbase + (uint64_t)tc * (uint64_t)ldkb;
  if (BF) {
    kA = bf2_to_f2(ldg_row4(kp, keep));
    kB = bf2_to_f2(ldg_row4(kp + 16, keep));
  } else {
    kA = ldg_row8(kp, keep);
    kB = ldg_row8(kp + 32, keep);
  }
}
template <bool BF>
__device__ __forceinline__ void load_q(const void* q, int64_t row, int ldq, int t, float2& qA,
                                       float2& qB) {
  if (BF) {
    const uint16_t* p = reinterpret_cast<const uint16_t*>(q) + row * ldq + 2 * t;
    qA = bf2_to_f2(*reinterpret_cast<const uint32_t*>(p));
    qB = bf2_to_f2(*reinterpret_cast<const uint32_t*>(p + 8));
  } else {
    const float* p = reinterpret_cast<const float*>(q) + row * ldq + 2 * t;
    qA = *reinterpret_cast<const float2*>(p);
    qB = *reinterpret_cast<const float2*>(p + 8);
  }
}

// accumulation phase of the forward: CNT consecutive edges of the tile starting at e0, no
// per-edge predicates (the caller decomposes n into 8 + 4 + 2 + 1): CNT gathered v rows in
// flight, then per edge one LDS.64 (p, p), one LDS.128 of the staged feature row, 4 packed FMAs
template <bool BF, int CNT>
__device__ __forceinline__ void fwd_gather(ulonglong2 (&vv)[CNT], int e0, int mycol,
                                           const char* vbase, unsigned ldvb, uint64_t keep) {
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    const unsigned tc = (unsigned)__shfl_sync(kFull, mycol, e0 + u);
    vv[u] = gather_v<BF>(vbase, tc, ldvb, keep);
  }
}
template <bool BF, int CNT, bool ALIGNED>
__device__ __forceinline__ void fwd_consume(const ulonglong2 (&vv)[CNT], int e0,
                                            const f32x2* pcol, uint32_t tile, uint32_t l7s,
                                            bool want_abar, f32x2& accv01, f32x2& accv23,
                                            f32x2& acca01, f32x2& acca23) {
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    const f32x2 pp = pcol[(e0 + u) * kH];
    fma2(accv01, pp, vv[u].x);
    fma2(accv23, pp, vv[u].y);
    if (want_abar) {
      const ulonglong2 a4 = lds_a_chunk<BF, ALIGNED>(tile, e0, u, l7s);
      fma2(acca01, pp, a4.x);
      fma2(acca23, pp, a4.y);
    }
  }
}
template <bool BF, int CNT, bool ALIGNED>
__device__ __forceinline__ void fwd_accumulate(int e0, int mycol, const char* vbase,
                                               unsigned ldvb, uint64_t keep, const f32x2* pcol,
                                               uint32_t tile, uint32_t l7s, bool want_abar,
                                               f32x2& accv01, f32x2& accv23, f32x2& acca01,
                                               f32x2& acca23) {
  ulonglong2 vv[CNT];
  fwd_gather<BF, CNT>(vv, e0, mycol, vbase, ldvb, keep);
  fwd_consume<BF, CNT, ALIGNED>(vv, e0, pcol, tile, l7s, want_abar, accv01, accv23, acca01,
                                acca23);
}

// backward: dp = <dY, v> + <dAbar, a> partial sums of CNT (<= 8) edges -> s[0..CNT)
template <bool BF, int CNT, bool ALIGNED, int SN = 8>
__device__ __forceinline__ void bwd_partials(float (&s)[SN], int e0, int mycol, const char* vbase,
                                             unsigned ldvb, uint64_t keep, uint32_t tile,
                                             uint32_t l7s, bool has_dab, f32x2 dy01, f32x2 dy23,
                                             f32x2 dab01, f32x2 dab23) {
  ulonglong2 vv[CNT];
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    const unsigned tc = (unsigned)__shfl_sync(kFull, mycol, e0 + u);
    vv[u] = gather_v<BF>(vbase, tc, ldvb, keep);
  }
#pragma unroll
  for (int u = 0; u < CNT; ++u) {
    f32x2 d2 = mul2(dy01, vv[u].x);
    fma2(d2, dy23, vv[u].y);
    if (has_dab) {
      const ulonglong2 a4 = lds_a_chunk<BF, ALIGNED>(tile, e0, u, l7s);
      fma2(d2, dab01, a4.x);
      fma2(d2, dab23, a4.y);
    }
    s[u] = fast::hsum2(d2);
  }
}

struct FwdArgs {
  const void* q; int ldq;     // fp32 or bf16 (template parameter BF), leading dims in elements
  const void* k; int ldk;
  const void* v; int ldv;
  const int32_t* rowptr; const int32_t* col;
  int64_t num_rows;
  const float* Wq; const float* bq; const float* Wk; const float* bk;
  int scale_mode; float scale_value;
  float* agg_v; float* abar; float* sump; float* m; float* z;
  int rows_per_warp;
};

// shared memory (after 1024-byte alignment):
//   [kWarps x 2 tile stages 8 KB][weight fragments NFRAG x 8 KB][bias 128 B]
//   [mbarriers 2 per warp][kWarps x p tile 1 KB]
template <int kWarps, int NFRAG>
struct Smem {
  static constexpr int tile_off = 0;
  static constexpr int frag_off = kWarps * 2 * kChunkBytes;
  static constexpr int bias_off = frag_off + NFRAG * kFragBytes;
  static constexpr int bar_off = bias_off + 128;
  static constexpr int p_off = bar_off + kWarps * 2 * 8;
  static constexpr int total = p_off + kWarps * 32 * kH * 8;
};
using FwdSmem = Smem<kFwdWarps, 1>;
using BwdSmem = Smem<kBwdWarps, 2>;

// the warp's position in its list of tiles + the operands fetched one tile / one row ahead
// (rowptr -> col -> gathered rows is otherwise a chain of dependent latencies per row)
struct Cursor {
  int64_t row, row1;
  int b, e;          // CSR range of `row`
  int e_next;        // rowptr[row + 2] (end of the next row), valid while row + 1 < row1
  int e_end;         // end of the warp's slab
  const int32_t* rowptr;
  const int32_t* col;
  int lane;
  int col_next;      // col[first tile of the next tile][lane]
  int tb_next, n_next;   // the tile after the current one (n_next = 0: none / not known)

  // called when starting tile [tb, tb + n) of the current row: what comes after it?
  __device__ __forceinline__ void look_ahead(int tb, int n) {
    if (tb + 32 < e) { tb_next = tb + 32; n_next = min(32, e - tb_next); }
    else if (row + 1 < row1) { tb_next = e; n_next = min(32, e_next - e); }
    else { tb_next = e; n_next = 0; }
    col_next = (n_next > 0 && lane < n_next) ? col[tb_next + lane] : 0;
  }
};

template <bool BF>
__global__ void __launch_bounds__(kFwdWarps * 32, 2)
k_attn_fwd_tile(const __grid_constant__ TileMaps tmA, const FwdArgs P) {
  constexpr int kElt = BF ? 2 : 4;
  using L = FwdSmem;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem =
      smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);   // SWIZZLE_128B atoms
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint4* frag1 = reinterpret_cast<const uint4*>(smem + L::frag_off);
  const float* bias_s = reinterpret_cast<const float*>(smem + L::bias_off);
  float2* p_s = reinterpret_cast<float2*>(smem + L::p_off) + w * 32 * kH;   // (p, p) pairs

  if (BF) build_frag1_bf16(reinterpret_cast<uint2*>(smem + L::frag_off), P.Wq, P.Wk);
  else build_frag1(reinterpret_cast<uint4*>(smem + L::frag_off), P.Wq, P.Wk);
  build_bias(reinterpret_cast<float*>(smem + L::bias_off), P.Wq, P.bq, P.Wk, P.bk);
  __syncthreads();

  const int64_t gw = (int64_t)blockIdx.x * kFwdWarps + w;
  const int64_t row0 = gw * P.rows_per_warp;
  if (row0 >= P.num_rows) return;

  TilePipe pipe;
  pipe.buf = smem + L::tile_off + w * (2 * kChunkBytes);
  pipe.bar = reinterpret_cast<uint64_t*>(smem + L::bar_off) + w * 2;
  pipe.tm = &tmA;
  pipe.policy = policy_evict_first();
  pipe.box_bytes = BF ? 512u : 1024u;
  pipe.lane = lane;
  pipe.init();
  const uint64_t keep = policy_evict_last();

  Cursor cu;
  cu.row = row0;
  cu.row1 = min(row0 + (int64_t)P.rows_per_warp, P.num_rows);
  cu.rowptr = P.rowptr; cu.col = P.col; cu.lane = lane;
  cu.b = P.rowptr[row0];
  cu.e = P.rowptr[row0 + 1];
  cu.e_next = (row0 + 1 < cu.row1) ? P.rowptr[row0 + 2] : cu.e;
  cu.e_end = P.rowptr[cu.row1];

  const bool want_abar = P.abar != nullptr;
  const int hb = lane >> 3;                     // head of my 4 value channels
  const int hsrc = 2 * (hb & 1);                // a lane holding head hb in the fragment layout
  const int pw = ((hb & 1) << 1) | (hb >> 1);   // word of head hb in a p-tile row [h0 h2 h1 h3]
  const char* kbase = reinterpret_cast<const char*>(P.k) + 2 * t * kElt;
  const char* vbase = reinterpret_cast<const char*>(P.v) + 4 * lane * kElt;
  const unsigned ldkb = (unsigned)P.ldk * kElt, ldvb = (unsigned)P.ldv * kElt;   // row strides, bytes
  // byte offset of my 16-byte chunk in tile row u of a group of 8: u * 128 + ((l7 ^ u) << 4)
  const uint32_t l7s = (uint32_t)((lane & 7) << 4);

  int mycol = (lane < min(32, cu.e - cu.b)) ? P.col[cu.b + lane] : 0;
  float2 qA, qB;
  load_q<BF>(P.q, row0, P.ldq, t, qA, qB);

  for (; cu.row < cu.row1; ++cu.row) {
    const int64_t row = cu.row;
    const int b = cu.b, e = cu.e;
    // q of the next row
    float2 qA_n = qA, qB_n = qB;
    if (row + 1 < cu.row1) {
      load_q<BF>(P.q, row + 1, P.ldq, t, qA_n, qB_n);
    }
    const float scale = fast::qk_scale_fast(P.scale_mode, P.scale_value, e - b);
    qA.x *= scale; qA.y *= scale; qB.x *= scale; qB.y *= scale;
    float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;   // heads t>>1 and 2 + (t>>1)
    f32x2 accv01 = 0ull, accv23 = 0ull, acca01 = 0ull, acca23 = 0ull;

    for (int tb = b; tb < e; tb += 32) {
      const int n = min(32, e - tb);
      const bool two = n > 16;
      __syncwarp();                              // previous tile's stage / p tile are free
      const uint32_t tile = pipe.acquire(tb, n);
      cu.look_ahead(tb, n);
      if (cu.n_next > 0) pipe.prefetch(cu.tb_next, cu.n_next);
      // gathered k rows of my 4 edges (in flight during the tensor-core phase)
      // (slots past n read node 0: their logits are masked below)
      float2 kA[4], kB[4];
#pragma unroll
      for (int idx = 0; idx < 4; ++idx) {
        if (idx < 2 || two) {
          const unsigned tc = (unsigned)__shfl_sync(kFull, mycol, 8 * idx + g);
          gather_k<BF>(kbase, tc, ldkb, keep, kA[idx], kB[idx]);
        } else {
          kA[idx] = make_float2(0.f, 0.f); kB[idx] = kA[idx];
        }
      }
      float acc[2][4][4];
      if (BF) rpe_tile_bf16(acc, tile, two, reinterpret_cast<const uint2*>(frag1), bias_s, lane);
      else rpe_tile(acc, tile, two, frag1, bias_s, lane);

      // logits (base 2) of my 4 edges x 2 heads
      float cA[4], cB[4];
#pragma unroll
      for (int idx = 0; idx < 4; ++idx) {
        const int m = idx >> 1, hf = (idx & 1) * 2;
        float pa = (qA.x + acc[m][0][hf]) * (kA[idx].x + acc[m][2][hf]);
        pa = fmaf(qA.y + acc[m][0][hf + 1], kA[idx].y + acc[m][2][hf + 1], pa);
        float pb = (qB.x + acc[m][1][hf]) * (kB[idx].x + acc[m][3][hf]);
        pb = fmaf(qB.y + acc[m][1][hf + 1], kB[idx].y + acc[m][3][hf + 1], pb);
        pa += __shfl_xor_sync(kFull, pa, 1);
        pb += __shfl_xor_sync(kFull, pb, 1);
        const bool valid = 8 * idx + g < n;
        cA[idx] = valid ? pa * kLog2e : -INFINITY;
        cB[idx] = valid ? pb * kLog2e : -INFINITY;
      }
      // Requesting the first 8 gathered v rows here, so that their latency overlaps the softmax,
      // was measured SLOWER (0.418 vs 0.297 ms): at the 128-register budget of 16 warps/SM the 32
      // extra live registers spill (272 B).  Kept behind SPT_FWD_PREFETCH for the record.
      ulonglong2 vpre[8];
#ifdef SPT_FWD_PREFETCH
      const bool pre = n >= 8;
#else
      const bool pre = false;
#endif
      if (pre) fwd_gather<BF, 8>(vpre, 0, mycol, vbase, ldvb, keep);
      float tA = fmaxf(fmaxf(cA[0], cA[1]), fmaxf(cA[2], cA[3]));
      float tB = fmaxf(fmaxf(cB[0], cB[1]), fmaxf(cB[2], cB[3]));
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        tA = fmaxf(tA, __shfl_xor_sync(kFull, tA, o));
        tB = fmaxf(tB, __shfl_xor_sync(kFull, tB, o));
      }
      const float mA_new = fmaxf(mA, tA), mB_new = fmaxf(mB, tB);
      const float alA = ex2(mA - mA_new), alB = ex2(mB - mB_new);   // 0 on the first tile
      float sA = 0.f, sB = 0.f;
#pragma unroll
      for (int idx = 0; idx < 4; ++idx) {
        const float pA = ex2(cA[idx] - mA_new), pB = ex2(cB[idx] - mB_new);   // 0 if invalid
        sA += pA; sB += pB;
        // p tile row of edge 8idx+g: 4 (p, p) pairs in the word order [h0 h2 h1 h3]; lanes
        // t = 0 / 2 own heads (0, 2) / (1, 3): one 16-byte store each
        if ((t & 1) == 0)
          *reinterpret_cast<float4*>(p_s + (8 * idx + g) * kH + t) = make_float4(pA, pA, pB, pB);
      }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) {
        sA += __shfl_xor_sync(kFull, sA, o);
        sB += __shfl_xor_sync(kFull, sB, o);
      }
      lA = fmaf(lA, alA, sA); lB = fmaf(lB, alB, sB);
      mA = mA_new; mB = mB_new;
      __syncwarp();

      // accumulation layout: lane = value channels 4*lane.. (head hb) + abar[hb][4*(lane&7)..]
      if (tb != b) {   // later tiles of a long row: rescale the running sums
        const float a0 = __shfl_sync(kFull, alA, hsrc), a1 = __shfl_sync(kFull, alB, hsrc);
        const float al = hb < 2 ? a0 : a1;
        const f32x2 aa = pack2(al, al);
        accv01 = mul2(accv01, aa); accv23 = mul2(accv23, aa);
        acca01 = mul2(acca01, aa); acca23 = mul2(acca23, aa);
      }
      const f32x2* pcol = reinterpret_cast<const f32x2*>(p_s) + pw;
      {
        int e0 = 0;
        if (pre) {
          fwd_consume<BF, 8, true>(vpre, 0, pcol, tile, l7s, want_abar, accv01, accv23, acca01,
                               acca23);
          e0 = 8;
        }
#ifdef SPT_FWD_GATHER16
        // 16 gathered rows in flight for the rows that have them (the accumulator fragments
        // are dead here, so the registers exist)
        for (; e0 + 16 <= n; e0 += 16)
          fwd_accumulate<BF, 16, true>(e0, mycol, vbase, ldvb, keep, pcol, tile, l7s, want_abar,
                                       accv01, accv23, acca01, acca23);
#endif
        for (; e0 + 8 <= n; e0 += 8)
          fwd_accumulate<BF, 8, true>(e0, mycol, vbase, ldvb, keep, pcol, tile, l7s, want_abar,
                                  accv01, accv23, acca01, acca23);
        if (n & 4) {
          fwd_accumulate<BF, 4, false>(e0, mycol, vbase, ldvb, keep, pcol, tile, l7s, want_abar,
                                   accv01, accv23, acca01, acca23);
          e0 += 4;
        }
        if (n & 2) {
          fwd_accumulate<BF, 2, false>(e0, mycol, vbase, ldvb, keep, pcol, tile, l7s, want_abar,
                                   accv01, accv23, acca01, acca23);
          e0 += 2;
        }
        if (n & 1)
          fwd_accumulate<BF, 1, false>(e0, mycol, vbase, ldvb, keep, pcol, tile, l7s, want_abar,
                                   accv01, accv23, acca01, acca23);
      }
      mycol = cu.col_next;
    }

    // epilogue of the row
    const float zA = lA + 1e-16f, zB = lB + 1e-16f;     // PyG softmax: + 1e-16 after the sum
    const float iA = fast_rcp(zA), iB = fast_rcp(zB);
    {
      const float z0 = __shfl_sync(kFull, iA, hsrc), z1 = __shfl_sync(kFull, iB, hsrc);
      const float inv = hb < 2 ? z0 : z1;
      const f32x2 ii = pack2(inv, inv);
      ulonglong2 o;
      o.x = mul2(accv01, ii); o.y = mul2(accv23, ii);
      *reinterpret_cast<ulonglong2*>(P.agg_v + row * kC + 4 * lane) = o;
      if (want_abar) {
        o.x = mul2(acca01, ii); o.y = mul2(acca23, ii);
        *reinterpret_cast<ulonglong2*>(P.abar + row * (kH * kF) + 4 * lane) = o;
      }
    }
    if (g == 0 && (t & 1) == 0) {
      const int h0 = t >> 1, h1 = 2 + (t >> 1);
      const bool any = e > b;
      P.m[row * kH + h0] = any ? mA * kLn2 : 0.f;    // natural-log units
      P.m[row * kH + h1] = any ? mB * kLn2 : 0.f;
      P.z[row * kH + h0] = zA;
      P.z[row * kH + h1] = zB;
      P.sump[row * kH + h0] = lA * iA;
      P.sump[row * kH + h1] = lB * iB;
    }
    // advance to the next row
    if (e == b) {   // empty row: nothing was prefetched through the tile loop
      cu.look_ahead(b, 0);
      mycol = cu.col_next;
    }
    cu.b = e;
    cu.e = cu.e_next;
    if (row + 2 < cu.row1) cu.e_next = P.rowptr[row + 3];
    qA = qA_n; qB = qB_n;
  }
}

// transpose-reduce of 8 per-edge partial sums over the 8 lanes of a head (7 shuffles): lane j8
// ends up with the full sum of edge j8
__device__ __forceinline__ float butterfly8(const float* s, int j8) {
  float r4[4], r2[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = (j8 & 4) ? s[i] : s[i + 4];
    const float keep = (j8 & 4) ? s[i + 4] : s[i];
    r4[i] = keep + __shfl_xor_sync(kFull, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = (j8 & 2) ? r4[i] : r4[i + 2];
    const float keep = (j8 & 2) ? r4[i + 2] : r4[i];
    r2[i] = keep + __shfl_xor_sync(kFull, send, 2);
  }
  const float send = (j8 & 1) ? r2[0] : r2[1];
  const float keep = (j8 & 1) ? r2[1] : r2[0];
  return keep + __shfl_xor_sync(kFull, send, 1);
}

// ------------------------------------------------------------------ backward rows
struct BwdArgs {
  const void* q; int ldq;
  const void* k; int ldk;
  const void* v; int ldv;
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

template <bool BF>
__global__ void __launch_bounds__(kBwdWarps * 32, 2)
k_attn_bwd_tile(const __grid_constant__ TileMaps tmA, const BwdArgs P) {
  constexpr int kElt = BF ? 2 : 4;
  using L = BwdSmem;
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const uint4* frag1 = reinterpret_cast<const uint4*>(smem + L::frag_off);
  const uint4* frag2 = reinterpret_cast<const uint4*>(smem + L::frag_off + kFragBytes);
  const float* bias_s = reinterpret_cast<const float*>(smem + L::bias_off);
  float* dp_s = reinterpret_cast<float*>(smem + L::p_off) + w * 32 * kH * 2;

  if (BF) {
    build_frag1_bf16(reinterpret_cast<uint2*>(smem + L::frag_off), P.Wq, P.Wk);
    build_frag2_bf16(reinterpret_cast<uint2*>(smem + L::frag_off + kFragBytes), P.Wq, P.Wk);
  } else {
    build_frag1(reinterpret_cast<uint4*>(smem + L::frag_off), P.Wq, P.Wk);
    build_frag2(reinterpret_cast<uint4*>(smem + L::frag_off + kFragBytes), P.Wq, P.Wk);
  }
  build_bias(reinterpret_cast<float*>(smem + L::bias_off), P.Wq, P.bq, P.Wk, P.bk);
  __syncthreads();

  const int64_t gw = (int64_t)blockIdx.x * kBwdWarps + w;
  const int64_t row0 = gw * P.rows_per_warp;
  if (row0 >= P.num_rows) return;

  TilePipe pipe;
  pipe.buf = smem + L::tile_off + w * (2 * kChunkBytes);
  pipe.bar = reinterpret_cast<uint64_t*>(smem + L::bar_off) + w * 2;
  pipe.tm = &tmA;
  pipe.policy = policy_evict_first();
  pipe.box_bytes = BF ? 512u : 1024u;
  pipe.lane = lane;
  pipe.init();
  const uint64_t keep = policy_evict_last();

  Cursor cu;
  cu.row = row0;
  cu.row1 = min(row0 + (int64_t)P.rows_per_warp, P.num_rows);
  cu.rowptr = P.rowptr; cu.col = P.col; cu.lane = lane;
  cu.b = P.rowptr[row0];
  cu.e = P.rowptr[row0 + 1];
  cu.e_next = (row0 + 1 < cu.row1) ? P.rowptr[row0 + 2] : cu.e;
  cu.e_end = P.rowptr[cu.row1];

  const bool has_dab = P.d_abar != nullptr && P.abar != nullptr;
  const bool want_da = P.da != nullptr;
  const int hb = lane >> 3;
  const int j8 = lane & 7;
  const char* kbase = reinterpret_cast<const char*>(P.k) + 2 * t * kElt;
  const char* vbase = reinterpret_cast<const char*>(P.v) + 4 * lane * kElt;
  const unsigned ldkb = (unsigned)P.ldk * kElt, ldvb = (unsigned)P.ldv * kElt;   // row strides, bytes
  const int hA = t >> 1, hB = 2 + (t >> 1);
  const int hsl = (t & 1) * 2 + (t >> 1);   // head fed through k-slot t of the P . dAbar step
  const uint32_t l7s = (uint32_t)((lane & 7) << 4);

  int mycol = (lane < min(32, cu.e - cu.b)) ? P.col[cu.b + lane] : 0;
  float2 qA, qB;
  load_q<BF>(P.q, row0, P.ldq, t, qA, qB);

  for (; cu.row < cu.row1; ++cu.row) {
    const int64_t row = cu.row;
    const int b = cu.b, e = cu.e;
    float2 qA_n = qA, qB_n = qB;
    if (row + 1 < cu.row1) {
      load_q<BF>(P.q, row + 1, P.ldq, t, qA_n, qB_n);
    }
    const float scale = fast::qk_scale_fast(P.scale_mode, P.scale_value, e - b);
    qA.x *= scale; qA.y *= scale; qB.x *= scale; qB.y *= scale;
    const float m2A = P.m[row * kH + hA] * kLog2e, m2B = P.m[row * kH + hB] * kLog2e;
    const float ziA = fast_rcp(P.z[row * kH + hA]), ziB = fast_rcp(P.z[row * kH + hB]);
    // accumulation layout operands of the row
    const float4 dy = *reinterpret_cast<const float4*>(P.d_agg_v + row * kC + 4 * lane);
    float4 dab = make_float4(0.f, 0.f, 0.f, 0.f);
    float delta;   // <dY_h, agg_h> + <dAbar_h, abar_h> of head hb (= sum_e p_e dp_e)
    {
      const float4 ag = *reinterpret_cast<const float4*>(P.agg_v + row * kC + 4 * lane);
      float part = dy.x * ag.x + dy.y * ag.y + dy.z * ag.z + dy.w * ag.w;
      if (has_dab) {
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
    // B fragment of the P . dAbar k-step: b0 = dAbar[row][head(k-slot t)][8nf + g], b1 = 0
    uint32_t dbhi[4], dblo[4];
#pragma unroll
    for (int nf = 0; nf < 4; ++nf) {
      if (BF) {
        // bf16 k-step of 16 slots: slots (2t, 2t+1) of the even lanes t = 0 / 2 carry heads
        // (0, 2) / (1, 3) — the pair a lane already holds; odd lanes and slots 8..15 are zero
        const float x0 = has_dab ? P.d_abar[row * (kH * kF) + hA * kF + 8 * nf + g] : 0.f;
        const float x1 = has_dab ? P.d_abar[row * (kH * kF) + hB * kF + 8 * nf + g] : 0.f;
        dbhi[nf] = (t & 1) ? 0u : f2_to_bf2(x0, x1);
        dblo[nf] = 0u;
      } else {
        const float x = has_dab ? P.d_abar[row * (kH * kF) + hsl * kF + 8 * nf + g] : 0.f;
        split_tf32(__float_as_uint(x), dbhi[nf], dblo[nf]);
      }
    }
    float dqacc[4] = {0.f, 0.f, 0.f, 0.f};   // dq[8nn + 2t + j], nn = 0,1

    for (int tb = b; tb < e; tb += 32) {
      const int n = min(32, e - tb);
      const bool two = n > 16;
      __syncwarp();
      const uint32_t tile = pipe.acquire(tb, n);
      cu.look_ahead(tb, n);
      if (cu.n_next > 0) pipe.prefetch(cu.tb_next, cu.n_next);
      // (slots past n read node 0: their logits are masked below)
      float2 kA[4], kB[4];
#pragma unroll
      for (int idx = 0; idx < 4; ++idx) {
        if (idx < 2 || two) {
          const unsigned tc = (unsigned)__shfl_sync(kFull, mycol, 8 * idx + g);
          gather_k<BF>(kbase, tc, ldkb, keep, kA[idx], kB[idx]);
        } else {
          kA[idx] = make_float2(0.f, 0.f); kB[idx] = kA[idx];
        }
      }
      float acc[2][4][4];
      if (BF) rpe_tile_bf16(acc, tile, two, reinterpret_cast<const uint2*>(frag1), bias_s, lane);
      else rpe_tile(acc, tile, two, frag1, bias_s, lane);

      // q_e, k_e in place (acc[m][0/1] = q_e heads A/B, acc[m][2/3] = k_e), p of my edges
      float pA[4], pB[4];
#pragma unroll
      for (int idx = 0; idx < 4; ++idx) {
        const int m = idx >> 1, hf = (idx & 1) * 2;
        acc[m][0][hf] += qA.x; acc[m][0][hf + 1] += qA.y;
        acc[m][1][hf] += qB.x; acc[m][1][hf + 1] += qB.y;
        acc[m][2][hf] += kA[idx].x; acc[m][2][hf + 1] += kA[idx].y;
        acc[m][3][hf] += kB[idx].x; acc[m][3][hf + 1] += kB[idx].y;
        float pa = acc[m][0][hf] * acc[m][2][hf];
        pa = fmaf(acc[m][0][hf + 1], acc[m][2][hf + 1], pa);
        float pb = acc[m][1][hf] * acc[m][3][hf];
        pb = fmaf(acc[m][1][hf + 1], acc[m][3][hf + 1], pb);
        pa += __shfl_xor_sync(kFull, pa, 1);
        pb += __shfl_xor_sync(kFull, pb, 1);
        const bool valid = 8 * idx + g < n;
        pA[idx] = valid ? ex2(fmaf(pa, kLog2e, -m2A)) * ziA : 0.f;
        pB[idx] = valid ? ex2(fmaf(pb, kLog2e, -m2B)) * ziB : 0.f;
        if ((t & 1) == 0 && valid) {
          P.Pbuf[(size_t)(tb + 8 * idx + g) * kH + hA] = pA[idx];
          P.Pbuf[(size_t)(tb + 8 * idx + g) * kH + hB] = pB[idx];
        }
      }

      // dp - delta of every (edge, head): accumulation layout, 8 edges per butterfly
      for (int e0 = 0; e0 < n; e0 += 8) {
        const int cnt = n - e0;              // edges of this group (the last one may be short)
#ifdef SPT_BWD_GATHER16
        if (cnt >= 16) {   // 16 gathered rows in flight, two butterflies
          float s16[16];
          bwd_partials<BF, 16, true, 16>(s16, e0, mycol, vbase, ldvb, keep, tile, l7s, has_dab,
                                         dy01, dy23, dab01, dab23);
          dp_s[(e0 + j8) * kH + hb] = butterfly8(s16, j8) - delta;
          dp_s[(e0 + 8 + j8) * kH + hb] = butterfly8(s16 + 8, j8) - delta;
          e0 += 8;
          continue;
        }
#endif
        float s[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] = 0.f;
        if (cnt >= 8) {
          bwd_partials<BF, 8, true>(s, e0, mycol, vbase, ldvb, keep, tile, l7s, has_dab, dy01, dy23,
                                dab01, dab23);
        } else {
          float s4[8], s2[8], s1[8];
          int o = 0;
          if (cnt & 4) {
            bwd_partials<BF, 4, true>(s4, e0, mycol, vbase, ldvb, keep, tile, l7s, has_dab, dy01,
                                  dy23, dab01, dab23);
            s[0] = s4[0]; s[1] = s4[1]; s[2] = s4[2]; s[3] = s4[3];
            o = 4;
          }
          if (cnt & 2) {
            bwd_partials<BF, 2, false>(s2, e0 + o, mycol, vbase, ldvb, keep, tile, l7s, has_dab, dy01,
                                   dy23, dab01, dab23);
            if (o) { s[4] = s2[0]; s[5] = s2[1]; } else { s[0] = s2[0]; s[1] = s2[1]; }
            o += 2;
          }
          if (cnt & 1) {
            bwd_partials<BF, 1, false>(s1, e0 + o, mycol, vbase, ldvb, keep, tile, l7s, has_dab, dy01,
                                   dy23, dab01, dab23);
            // o in {0, 2, 4, 6}
            if (o == 0) s[0] = s1[0]; else if (o == 2) s[2] = s1[0];
            else if (o == 4) s[4] = s1[0]; else s[6] = s1[0];
          }
        }
        // transpose-reduce over the 8 lanes of a head: lane j8 ends with edge e0 + j8
        const float dp = butterfly8(s, j8);
        dp_s[(e0 + j8) * kH + hb] = dp - delta;
      }
      __syncwarp();

      // G = [dq_e | dk_e] in place: acc[m][0/1] <- dc * k_e, acc[m][2/3] <- dc * q_e
      float pk[4];   // p of k-slot t of the P . dAbar step, per edge
#pragma unroll
      for (int idx = 0; idx < 4; ++idx) {
        const int m = idx >> 1, hf = (idx & 1) * 2;
        const bool valid = 8 * idx + g < n;
        const float dA = valid ? pA[idx] * dp_s[(8 * idx + g) * kH + hA] : 0.f;
        const float dB = valid ? pB[idx] * dp_s[(8 * idx + g) * kH + hB] : 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float qa = acc[m][0][hf + j], ka = acc[m][2][hf + j];
          const float qb = acc[m][1][hf + j], kb = acc[m][3][hf + j];
          acc[m][0][hf + j] = valid ? dA * ka : 0.f;
          acc[m][2][hf + j] = valid ? dA * qa : 0.f;
          acc[m][1][hf + j] = valid ? dB * kb : 0.f;
          acc[m][3][hf + j] = valid ? dB * qb : 0.f;
        }
        pk[idx] = (t & 1) ? pB[idx] : pA[idx];
        if (valid) {
          float* gp = P.G + (size_t)(tb + 8 * idx + g) * kHD2 + 2 * t;
#pragma unroll
          for (int nn = 0; nn < 4; ++nn)
            *reinterpret_cast<float2*>(gp + 8 * nn) =
                make_float2(acc[m][nn][hf], acc[m][nn][hf + 1]);
        }
        dqacc[0] += acc[m][0][hf]; dqacc[1] += acc[m][0][hf + 1];
        dqacc[2] += acc[m][1][hf]; dqacc[3] += acc[m][1][hf + 1];
      }

      if (want_da) {
        // da = G . W + P . dAbar  (second tensor-core product; A = accumulator fragment)
        float dacc[2][4][4];
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int nf = 0; nf < 4; ++nf)
#pragma unroll
            for (int i = 0; i < 4; ++i) dacc[m][nf][i] = 0.f;
        if (BF) {
          if (has_dab) {
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              if (m == 0 || two) {
                uint32_t a[4];
                a[0] = (t & 1) ? 0u : f2_to_bf2(pA[2 * m], pB[2 * m]);
                a[1] = (t & 1) ? 0u : f2_to_bf2(pA[2 * m + 1], pB[2 * m + 1]);
                a[2] = a[3] = 0u;
#pragma unroll
                for (int nf = 0; nf < 4; ++nf) mma_bf16(dacc[m][nf], a, dbhi[nf], 0u);
              }
            }
          }
          const uint2* frag2b = reinterpret_cast<const uint2*>(frag2);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            uint2 bfr[4];
#pragma unroll
            for (int nf = 0; nf < 4; ++nf) bfr[nf] = frag2b[(ks * 4 + nf) * 32 + lane];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
              if (m == 0 || two) {
                // the accumulator fragments of n-tiles 2ks, 2ks+1 ARE the A fragment (k = output)
                uint32_t a[4];
                a[0] = f2_to_bf2(acc[m][2 * ks][0], acc[m][2 * ks][1]);
                a[1] = f2_to_bf2(acc[m][2 * ks][2], acc[m][2 * ks][3]);
                a[2] = f2_to_bf2(acc[m][2 * ks + 1][0], acc[m][2 * ks + 1][1]);
                a[3] = f2_to_bf2(acc[m][2 * ks + 1][2], acc[m][2 * ks + 1][3]);
#pragma unroll
                for (int nf = 0; nf < 4; ++nf) mma_bf16(dacc[m][nf], a, bfr[nf].x, bfr[nf].y);
              }
            }
          }
        } else {
        if (has_dab) {
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            if (m == 0 || two) {
              uint32_t ahi[4], alo[4];
              split_tf32(__float_as_uint(pk[2 * m]), ahi[0], alo[0]);
              split_tf32(__float_as_uint(pk[2 * m + 1]), ahi[1], alo[1]);
              ahi[2] = ahi[3] = alo[2] = alo[3] = 0u;
#pragma unroll
              for (int nf = 0; nf < 4; ++nf) {
                mma_tf32(dacc[m][nf], alo, dbhi[nf], 0u);
                mma_tf32(dacc[m][nf], ahi, dblo[nf], 0u);
                mma_tf32(dacc[m][nf], ahi, dbhi[nf], 0u);
              }
            }
          }
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint4 bfr[4];
#pragma unroll
          for (int nf = 0; nf < 4; ++nf) bfr[nf] = frag2[(ks * 4 + nf) * 32 + lane];
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            if (m == 0 || two) {
              // k-slot s <-> output 8ks + 2s, slot 4 + s <-> 8ks + 2s + 1
              uint32_t ahi[4], alo[4];
              split_tf32(__float_as_uint(acc[m][ks][0]), ahi[0], alo[0]);
              split_tf32(__float_as_uint(acc[m][ks][2]), ahi[1], alo[1]);
              split_tf32(__float_as_uint(acc[m][ks][1]), ahi[2], alo[2]);
              split_tf32(__float_as_uint(acc[m][ks][3]), ahi[3], alo[3]);
#pragma unroll
              for (int nf = 0; nf < 4; ++nf) mma_3x(dacc[m][nf], ahi, alo, bfr[nf]);
            }
          }
        }
        }
#pragma unroll
        for (int idx = 0; idx < 4; ++idx) {
          if (8 * idx + g < n) {
            const int m = idx >> 1, hf = (idx & 1) * 2;
            float* dp_ = P.da + (size_t)(tb + 8 * idx + g) * kF + 2 * t;
#pragma unroll
            for (int nf = 0; nf < 4; ++nf)
              *reinterpret_cast<float2*>(dp_ + 8 * nf) =
                  make_float2(dacc[m][nf][hf], dacc[m][nf][hf + 1]);
          }
        }
      }
      mycol = cu.col_next;
    }

    // dq of the row: reduce my 4 partial sums over the 8 row lanes (g)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float x = dqacc[i];
      x += __shfl_xor_sync(kFull, x, 4);
      x += __shfl_xor_sync(kFull, x, 8);
      x += __shfl_xor_sync(kFull, x, 16);
      dqacc[i] = x * scale;
    }
    if (g == 0) {
      *reinterpret_cast<float2*>(P.dq + row * P.lddq + 2 * t) = make_float2(dqacc[0], dqacc[1]);
      *reinterpret_cast<float2*>(P.dq + row * P.lddq + 8 + 2 * t) =
          make_float2(dqacc[2], dqacc[3]);
    }
    if (e == b) {
      cu.look_ahead(b, 0);
      mycol = cu.col_next;
    }
    cu.b = e;
    cu.e = cu.e_next;
    if (row + 2 < cu.row1) cu.e_next = P.rowptr[row + 3];
    qA = qA_n; qB = qB_n;
  }
}

}  // namespace tile
}  // namespace spt
