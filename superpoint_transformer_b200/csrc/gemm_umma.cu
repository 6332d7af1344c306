// gemm_umma.cu — nn.Linear forward / dX on the 5th-generation tensor cores.
//
//   C[M,N] = A[M,K] . B[N,K]^T + bias[N]          (fp32 in, fp32 out, fp32-accurate)
//
// Replaces the cuBLAS calls behind src/nn/attention.py:191,318 and src/nn/mlp.py:45 of the
// reference.  fp32 accuracy on TF32 tensor cores comes from the 3xTF32 split
//   x = hi + lo,  hi = tf32_rna(x),  lo = x - hi (stored as is: the tensor core reads only the
//   top 19 bits of a tf32 operand, i.e. truncates it),   A.B ~= Alo.Bhi + Ahi.Blo + Ahi.Bhi
// (~2^-21 relative), accumulated in fp32 in TMEM.
//
// k_gemm_nt_umma: one persistent CTA per SM, warp-specialised, tile = 128 rows x BN (<=192):
//   warp 0      TMA producer A: 128 x 32 fp32 chunks (SWIZZLE_128B, OOB zero-filled) into a
//                              ring of raw landing slots (2-8 x 16 KB in flight per SM)
//   warp 3      TMA producer B: the [BN, K] weights; RESIDENT in shared memory for the whole
//                              kernel when they fit (split once per CTA), else streamed per
//                              32-wide K chunk through 2-3 slots
//   warps 8-11  splitter A   : A row -> registers (ld.shared) -> hi / lo -> tcgen05.st into TMEM
//                              (row = lane, K along columns); the landing slot is released
//                              only after that tcgen05.st has consumed the values
//   warps 12-15 splitter B   : B chunk rewritten in place as `hi` + twin `lo` buffer
//                              (element-wise, so the TMA swizzle is preserved), fence.proxy.async;
//                              resident mode: done once, helped by warps 4-7, after which warps
//                              12-15 become the second epilogue group
//   warp 1      MMA issuer   : converged warp, elect.sync around the issue (uniform datapath):
//                              3 x tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) per 8-wide k-step,
//                              A operand from TMEM, B from K-major SWIZZLE_128B smem descriptors;
//                              tcgen05.commit releases the operand stage / publishes the
//                              accumulator
//   warps 4-7   epilogue     : tcgen05.ld 32 lanes x 32 columns (next slab in flight), + bias
//                              from smem, swizzled st.shared, fence.proxy.async, TMA store
//                              (clipped at M, N) issued by a FIXED lane (bulk groups are per
//                              thread); two TMEM accumulators so the epilogue of tile i overlaps
//                              the MMAs of tile i+1
//   warp 2      TMEM allocate / free (512 columns)
// HBM traffic per tile: A once, C once; B from L2 (once per CTA when resident).
// k_gemm_tn_umma (dW / dbias) is described where it is defined, further down.
#include <cuda.h>  // CUtensorMap types; the encoder is fetched through the runtime API
#include "umma_common.cuh"

#include "common.cuh"

namespace spt {
namespace umma {

constexpr int BM = 128;         // UMMA_M
constexpr int BK = 32;          // fp32 per K chunk: one 128-byte swizzle row
constexpr int UK = 8;           // UMMA_K of kind::tf32 (32 bytes)
constexpr int kThreads = 512;   // 16 warps
constexpr int kEpiWarp0 = 4;    // epilogue warps 4..7  (TMEM lane quadrant = warp % 4)
constexpr int kSplitWarp0 = 8;   // A splitter warps 8..11 (TMEM lane quadrant = warp % 4)
constexpr int kBSplitWarp0 = 12; // B splitter warps 12..15
constexpr int kSlab = 32;       // epilogue column slab (32 fp32 = 128 B = one swizzle row)
constexpr int kMaxStages = 4;
constexpr uint32_t kABytes = BM * BK * 4;      // 16 KB
constexpr size_t kSmemBudget = 226 * 1024;  // 227 KB opt-in limit minus 1 KB the kernel uses statically

// one 32-column slab of one accumulator row: + bias (smem broadcast) -> swizzled staging row
__device__ __forceinline__ void stage_slab(const uint32_t (&v)[32], unsigned char* buf, int row,
                                           const float* bias_slab) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 o;
    o.x = __uint_as_float(v[4 * j + 0]);
    o.y = __uint_as_float(v[4 * j + 1]);
    o.z = __uint_as_float(v[4 * j + 2]);
    o.w = __uint_as_float(v[4 * j + 3]);
    if (bias_slab) {
      const float4 bb = lds128(smem_u32(bias_slab + 4 * j));
      o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
    }
    sts128(smem_u32(buf) + (uint32_t)(row * 128 + ((j ^ (row & 7)) << 4)), o);
  }
}

#define TRC(role, idx, slot)                                                     \
  do {                                                                          \
    if (P.trace && blockIdx.x == 0 && (idx) < 32)                               \
      P.trace[((role) * 32 + (idx)) * 4 + (slot)] = clock64();                  \
  } while (0)

constexpr uint32_t kTmemCols = 512;
constexpr int kMaxRing = 8;     // raw A landing slots
constexpr int kMaxAStages = 4;  // A operand stages in TMEM (64 columns each)
constexpr int kMaxBSlots = 8;   // B operand slots in smem
constexpr uint32_t kSlabBytes = BM * kSlab * 4;  // 16 KB epilogue staging buffer
constexpr int kBarWords = 2 * kMaxRing + 2 * kMaxAStages + 3 * kMaxBSlots + 4 + 2;  // + tmem slot

struct Params {
  const float* bias;  // nullable
  int64_t M;
  int N, K;
  int BN;             // columns per tile, multiple of 16, <= 192
  int n_blocks;       // ceil(N / BN)
  int64_t tiles;      // ceil(M / 128) * n_blocks
  int ring;           // raw A landing slots (16 KB each)
  int a_stages;       // A hi/lo operand stages in TMEM
  int b_slots;        // B hi/lo operand slots in smem
  int b_resident;     // 1: slot kc holds chunk kc for the whole kernel (split once per CTA)
  int epi_bufs;       // 16 KB staging buffers for the TMA store (1 or 2)
  int epi_groups;     // 1, or 2 when the B splitter warps are free to help (resident B)
  uint32_t acc_stride;
  long long* trace;   // debug: per-role clock64 stamps of CTA 0 (SPT_UMMA_TRACE)
};

// Epilogue role of one group of 4 warps (TMEM lane quadrant q = warp % 4).  Group g of
// P.epi_groups takes the 32-column slabs g, g + groups, ...: tcgen05.ld (next slab in
// flight) -> + bias -> swizzled staging tile -> TMA store (clipped at M, N).
__device__ __forceinline__ void epilogue_role(const Params& P, const CUtensorMap* tmC,
                                              unsigned char* epi, int g, int q, int lane,
                                              uint32_t tmem_base, uint64_t* tmem_full,
                                              uint64_t* tmem_empty, const float* bias_s) {
  const int row = q * 32 + lane;  // tile-local row == TMEM lane
  const int G = P.epi_groups;
  const int nb = P.epi_bufs / G;  // staging buffers of this group
  unsigned char* mybuf = epi + (size_t)g * nb * kSlabBytes;
  const bool tr = (g == 0 && row == 0);
  uint32_t tl = 0, sc = 0;
  for (int64_t t = blockIdx.x; t < P.tiles; t += gridDim.x, ++tl) {
    const int m_blk = (int)(t / P.n_blocks), n_blk = (int)(t % P.n_blocks);
    const uint32_t acc = tl & 1, accph = (tl >> 1) & 1;
    if (tr) TRC(5, tl, 0);
    mbar_wait(&tmem_full[acc], accph, 7);
    tc_fence_after();
    if (tr) TRC(5, tl, 1);
    const int ncols = min(P.BN, P.N - n_blk * P.BN);
    const int nslab = (ncols + kSlab - 1) / kSlab;
    const float* bs = P.bias ? bias_s + n_blk * (int)P.acc_stride : nullptr;
    const uint32_t tsrc = tmem_base + acc * P.acc_stride + ((uint32_t)(q * 32) << 16);
    uint32_t va[32], vb[32];
    if (g < nslab) tmem_ld32_nowait(tsrc + g * kSlab, va);
    for (int sl = g; sl < nslab; sl += 2 * G) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int cur = sl + h * G;
        if (cur >= nslab) break;
        tmem_wait_ld();  // slab `cur` is in registers
        if (tr && cur == 0) TRC(6, tl, 0);
        if (cur + G < nslab) {
          if (h == 0) tmem_ld32_nowait(tsrc + (cur + G) * kSlab, vb);
          else tmem_ld32_nowait(tsrc + (cur + G) * kSlab, va);
        }
        unsigned char* buf = mybuf + (sc % nb) * kSlabBytes;
        // bulk async-groups are per thread: the same lane must commit and wait, so a fixed
        // lane (not elect.sync, whose choice is not specified to be stable) drives the store
        if (q == 0 && lane == 0) {  // the store that last used `buf` has read it
          if (nb == 2) bulk_wait_read<1>();
          else bulk_wait_read<0>();
        }
        if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        if (tr && cur == 0) TRC(6, tl, 1);
        if (h == 0) stage_slab(va, buf, row, bs ? bs + cur * kSlab : nullptr);
        else stage_slab(vb, buf, row, bs ? bs + cur * kSlab : nullptr);
        fence_proxy_async();
        if (tr && cur == 0) TRC(6, tl, 2);
        if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        if (q == 0 && lane == 0) {
          tma_store_2d(tmC, buf, n_blk * P.BN + cur * kSlab, m_blk * BM);
          bulk_commit();
        }
        if (tr && cur == 0) TRC(6, tl, 3);
        ++sc;
      }
    }
    // every tcgen05.ld of this group on this accumulator has completed (last wait::ld)
    tc_fence_before();
    mbar_arrive(&tmem_empty[acc]);
    if (tr) TRC(5, tl, 2);
  }
  if (q == 0 && lane == 0) bulk_wait_read<0>();
}

__global__ void __launch_bounds__(kThreads, 1)
k_gemm_nt_umma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const Params P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem =
      (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // swizzle atoms
  const uint32_t b_bytes = (uint32_t)P.BN * BK * 4;
  unsigned char* ringbuf = smem;                                   // ring x 16 KB raw A
  unsigned char* bbuf = ringbuf + (size_t)P.ring * kABytes;        // b_slots x (Bhi | Blo)
  unsigned char* epi = bbuf + (size_t)P.b_slots * 2 * b_bytes;     // epi_bufs x 16 KB
  uint64_t* bars = (uint64_t*)(epi + (size_t)P.epi_bufs * kSlabBytes);
  uint64_t* a_full = bars;                      // [ring]     TMA landed a raw A chunk
  uint64_t* a_free = a_full + kMaxRing;         // [ring]     splitter holds it in registers
  uint64_t* a_ready = a_free + kMaxRing;        // [a_stages] hi/lo written to TMEM
  uint64_t* a_empty = a_ready + kMaxAStages;    // [a_stages] MMAs that read the stage retired
  uint64_t* b_full = a_empty + kMaxAStages;     // [b_slots]  TMA landed a raw B chunk
  uint64_t* b_ready = b_full + kMaxBSlots;      // [b_slots]  hi/lo in place
  uint64_t* b_empty = b_ready + kMaxBSlots;     // [b_slots]  (streaming B only)
  uint64_t* tmem_full = b_empty + kMaxBSlots;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;         // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);
  float* bias_s = (float*)(bars + kBarWords);   // [n_blocks * acc_stride], zero-padded

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KC = (P.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int r = 0; r < P.ring; ++r) {
      mbar_init(&a_full[r], 1);
      mbar_init(&a_free[r], 128);
    }
    for (int s = 0; s < P.a_stages; ++s) {
      mbar_init(&a_ready[s], 128);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < P.b_slots; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_ready[s], P.b_resident ? 256 : 128);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 128 * P.epi_groups);
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
  if (P.bias) {
    // tile n_blk, slab column c reads bias_s[n_blk * acc_stride + c]
    for (int i = threadIdx.x; i < P.n_blocks * (int)P.acc_stride; i += kThreads) {
      const int nb = i / (int)P.acc_stride, c = i - nb * (int)P.acc_stride;
      const int col = nb * P.BN + c;
      bias_s[i] = (c < P.BN && col < P.N) ? __ldg(P.bias + col) : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a = tmem_base + 2 * P.acc_stride;  // stage s: hi at +64 s, lo at +64 s + 32

  if (warp == 0) {
    // ---------------- TMA producer, A: deep ring of raw 128 x 32 chunks ----------------
    uint32_t it = 0, ph = 0;
    int r = 0;
    for (int64_t t = blockIdx.x; t < P.tiles; t += gridDim.x) {
      const int m_blk = (int)(t / P.n_blocks);
      for (int kc = 0; kc < KC; ++kc, ++it) {
        mbar_wait(&a_free[r], ph ^ 1, 0);
        if (elect_one()) {
          TRC(0, it, 0);
          mbar_expect_tx(&a_full[r], kABytes);
          tma_load_2d(ringbuf + (size_t)r * kABytes, &tmA, kc * BK, m_blk * BM, &a_full[r]);
        }
        __syncwarp();
        if (++r == P.ring) { r = 0; ph ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ---------------- TMA producer, B (L2-resident weights) ----------------
    if (P.b_resident) {
      if (elect_one()) {
        for (int kc = 0; kc < KC; ++kc) {
          mbar_expect_tx(&b_full[kc], b_bytes);
          tma_load_2d(bbuf + (size_t)kc * 2 * b_bytes, &tmB, kc * BK, 0, &b_full[kc]);
        }
      }
      __syncwarp();
    } else {
      uint32_t it = 0, ph = 0;
      int s = 0;
      for (int64_t t = blockIdx.x; t < P.tiles; t += gridDim.x) {
        const int n_blk = (int)(t % P.n_blocks);
        for (int kc = 0; kc < KC; ++kc, ++it) {
          mbar_wait(&b_empty[s], ph ^ 1, 1);
          if (elect_one()) {
            TRC(1, it, 0);
            mbar_expect_tx(&b_full[s], b_bytes);
            tma_load_2d(bbuf + (size_t)s * 2 * b_bytes, &tmB, kc * BK, n_blk * P.BN,
                        &b_full[s]);
          }
          __syncwarp();
          if (++s == P.b_slots) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (converged warp, one elected lane issues) ----------------
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P.BN >> 3) << 17) |
                           ((uint32_t)(BM >> 4) << 24);
    uint32_t it = 0, tl = 0;
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    const uint32_t bbase = smem_u32(bbuf);
    for (int64_t t = blockIdx.x; t < P.tiles; t += gridDim.x, ++tl) {
      const uint32_t acc = tl & 1, accph = (tl >> 1) & 1;
      mbar_wait(&tmem_empty[acc], accph ^ 1, 2);
      tc_fence_after();
      const uint32_t d = tmem_base + acc * P.acc_stride;
      for (int kc = 0; kc < KC; ++kc, ++it) {
        if (P.b_resident) sb = kc;
        if (lane == 0) TRC(4, it, 0);
        if (!P.b_resident)
          mbar_wait(&b_ready[sb], phb, 3);
        else if (tl == 0)
          mbar_wait(&b_ready[sb], 0, 3);
        mbar_wait(&a_ready[sa], pha, 3);
        tc_fence_after();
        if (lane == 0) TRC(4, it, 1);
        const uint32_t a_hi = tmem_a + 64 * sa, a_lo = a_hi + 32;
        const uint32_t sbase = bbase + (uint32_t)sb * 2 * b_bytes;
        const uint64_t b_hi = smem_desc_sw128(sbase), b_lo = smem_desc_sw128(sbase + b_bytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t o = (uint64_t)(k * UK * 4 >> 4);  // advance inside the swizzle row
            umma_tf32_ts(d, a_lo + k * UK, b_hi + o, idesc, (kc | k) != 0);
            umma_tf32_ts(d, a_hi + k * UK, b_lo + o, idesc, 1);
            umma_tf32_ts(d, a_hi + k * UK, b_hi + o, idesc, 1);
          }
          umma_commit(&a_empty[sa]);  // TMEM stage free once these MMAs have read it
          if (!P.b_resident) umma_commit(&b_empty[sb]);
          if (kc == KC - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (lane == 0) TRC(4, it, 2);
        if (++sa == P.a_stages) { sa = 0; pha ^= 1; }
        if (!P.b_resident && ++sb == P.b_slots) { sb = 0; phb ^= 1; }
      }
    }
  } else if (warp >= kBSplitWarp0) {
    // ---------------- splitter, B: in place -> hi, twin buffer -> lo ----------------
    const int ts = threadIdx.x - kBSplitWarp0 * 32;
    if (P.b_resident) {
      // resident weights: split once, together with the (still idle) epilogue warps
      for (int kc = 0; kc < KC; ++kc) {
        mbar_wait(&b_full[kc], 0, 6);
        unsigned char* bs = bbuf + (size_t)kc * 2 * b_bytes;
        split_chunk(bs, bs + b_bytes, b_bytes / 16, ts + 128, 256);
        fence_proxy_async();
        mbar_arrive(&b_ready[kc]);
        if (ts == 0) TRC(3, kc, 2);
      }
      if (P.epi_groups == 2)  // the weights are split once: become the second epilogue group
        epilogue_role(P, &tmC, epi, 1, warp - kBSplitWarp0, lane, tmem_base, tmem_full,
                      tmem_empty, bias_s);
    } else {
      uint32_t it = 0;
      for (int64_t t = blockIdx.x; t < P.tiles; t += gridDim.x) {
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % P.b_slots;
          if (ts == 0) TRC(3, it, 0);
          mbar_wait(&b_full[s], (it / P.b_slots) & 1, 6);
          if (ts == 0) TRC(3, it, 1);
          unsigned char* bs = bbuf + (size_t)s * 2 * b_bytes;
          split_chunk(bs, bs + b_bytes, b_bytes / 16, ts, 128);
          fence_proxy_async();
          mbar_arrive(&b_ready[s]);
          if (ts == 0) TRC(3, it, 2);
        }
      }
    }
  } else if (warp >= kSplitWarp0) {
    // ---------------- splitter, A: raw row -> registers -> hi / lo -> TMEM ----------------
    const int q = warp - kSplitWarp0;   // TMEM lane quadrant (warp % 4)
    const int row = q * 32 + lane;      // chunk row == TMEM lane
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < P.tiles; t += gridDim.x) {
      for (int kc = 0; kc < KC; ++kc, ++it) {
        const int r = it % P.ring, s = it % P.a_stages;
        if (row == 0) TRC(2, it, 0);
        mbar_wait(&a_full[r], (it / P.ring) & 1, 4);
        if (row == 0) TRC(2, it, 1);
        const uint32_t arow = smem_u32(ringbuf) + (uint32_t)r * kABytes + (uint32_t)row * 128u;
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
        mbar_wait(&a_empty[s], ((it / P.a_stages) & 1) ^ 1, 5);  // previous MMAs retired
        tc_fence_after();
        if (row == 0) TRC(2, it, 2);
        const uint32_t ta = tmem_a + 64 * s + ((uint32_t)(q * 32) << 16);
        tmem_st32(ta, hi);
        tmem_st32(ta + 32, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        // Free the landing slot only now: the tcgen05.st above consumed the loaded values, so
        // the slot reads have really completed (an arrive issued right behind the loads is
        // not ordered with them in hardware and let the TMA refill race the reads).
        mbar_arrive(&a_free[r]);
        tc_fence_before();
        mbar_arrive(&a_ready[s]);
        if (row == 0) TRC(2, it, 3);
      }
    }
  } else if (warp >= kEpiWarp0) {
    if (P.b_resident) {
      const int ts = threadIdx.x - kEpiWarp0 * 32;
      for (int kc = 0; kc < KC; ++kc) {
        mbar_wait(&b_full[kc], 0, 6);
        unsigned char* bs = bbuf + (size_t)kc * 2 * b_bytes;
        split_chunk(bs, bs + b_bytes, b_bytes / 16, ts, 256);
        fence_proxy_async();
        mbar_arrive(&b_ready[kc]);
      }
    }
    epilogue_role(P, &tmC, epi, 0, warp - kEpiWarp0, lane, tmem_base, tmem_full, tmem_empty,
                  bias_s);
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

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// row-major fp32 matrix [rows, cols], leading dimension ld; box = [box_rows, 32 cols]
static bool make_map(CUtensorMap* tm, const float* ptr, int64_t rows, int64_t cols, int64_t ld,
                     int box_rows, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = encoder();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides,
             box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

// exported for csrc/attention.cu: [rows, 32] fp32 matrix, boxes of `box_rows` x 32 columns
bool make_map_rows32(CUtensorMap* tm, const float* ptr, int64_t rows, int64_t ld, int box_rows) {
  return make_map(tm, ptr, rows, 32, ld, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}
// exported for csrc/attention_umma.cuh: unswizzled box of `box_cols` x `box_rows` over a row-major
// fp32 matrix [rows, cols] (column-offset stores into a wider buffer)
bool make_map_box(CUtensorMap* tm, const float* ptr, int64_t rows, int64_t cols, int64_t ld,
                  int box_cols, int box_rows) {
  EncodeTiledFn enc = encoder();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box,
             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// the same for a bf16 [rows, 32] matrix (64-byte rows, SWIZZLE_64B)
bool make_map_rows32_bf16(CUtensorMap* tm, const void* ptr, int64_t rows, int box_rows) {
  EncodeTiledFn enc = encoder();
  if (!enc) return false;
  cuuint64_t dims[2] = {32u, (cuuint64_t)rows};
  cuuint64_t strides[1] = {64u};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box,
             estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

// layout the tensor-core path needs; otherwise the caller uses the mma.sync kernel
bool shape_ok(const float* A, int64_t M, int64_t K, int64_t lda, const float* B, int64_t N,
              int64_t ldb, const float* C, int64_t ldc) {
  const uintptr_t al = (uintptr_t)A | (uintptr_t)B | (uintptr_t)C;
  return (al & 15) == 0 && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0 && M >= 1 &&
         M < (1ll << 31) - 256 && N >= 1 && N <= 4096 && K >= 1 && K < (1 << 30) &&
         encoder() != nullptr;
}

int launch(const float* A, int64_t M, int64_t K, int64_t lda, const float* B, int64_t N,
           int64_t ldb, const float* bias, float* C, int64_t ldc, cudaStream_t stream) {
  Params P;
  P.bias = bias;
  P.M = M;
  P.N = (int)N;
  P.K = (int)K;
  const int n16 = (int)((N + 15) / 16 * 16);
  P.BN = n16 <= 192 ? n16 : 128;  // 2 accumulators + >= 2 A-operand stages in 512 TMEM columns
  P.n_blocks = (int)((N + P.BN - 1) / P.BN);
  P.tiles = ((M + BM - 1) / BM) * P.n_blocks;
  P.acc_stride = (uint32_t)((P.BN + 31) / 32 * 32);
  int a_stages = (int)((kTmemCols - 2 * P.acc_stride) / 64);
  P.a_stages = a_stages > kMaxAStages ? kMaxAStages : a_stages;
  const int KC = (int)((K + BK - 1) / BK);
  const size_t b_slot = 2 * (size_t)P.BN * BK * 4;
  const size_t misc = 1024 /*alignment*/ + (size_t)kBarWords * 8 + 16 +
                      (size_t)P.n_blocks * P.acc_stride * 4;
  // B resident (split once per CTA) when the whole [BN, K] hi/lo pair fits beside >= 2
  // landing slots; otherwise B chunks stream through 2-3 slots like A.
  P.b_resident = 0;
  P.ring = 0;
  if (P.n_blocks == 1 && KC <= kMaxBSlots) {
    for (int eb = 2; eb >= 1 && !P.b_resident; --eb) {
      const size_t fixed = misc + (size_t)KC * b_slot + (size_t)eb * kSlabBytes;
      if (fixed + (eb == 2 ? 4 : 2) * (size_t)kABytes <= kSmemBudget) {
        P.b_resident = 1;
        P.b_slots = KC;
        P.epi_bufs = eb;
        P.ring = (int)((kSmemBudget - fixed) / kABytes);
      }
    }
  }
  if (!P.b_resident) {
    P.epi_bufs = 2;
    P.b_slots = 3;
    size_t fixed = misc + (size_t)P.b_slots * b_slot + (size_t)P.epi_bufs * kSlabBytes;
    if (fixed + 3 * (size_t)kABytes > kSmemBudget) {
      P.b_slots = 2;
      fixed = misc + (size_t)P.b_slots * b_slot + (size_t)P.epi_bufs * kSlabBytes;
    }
    SPT_REQUIRE(fixed + 2 * (size_t)kABytes <= kSmemBudget, SPT_E_UNSUPPORTED,
                "gemm_nt(umma): shared memory budget");
    P.ring = (int)((kSmemBudget - fixed) / kABytes);
  }
  P.epi_groups = (P.b_resident && P.epi_bufs == 2) ? 2 : 1;
  if (P.ring > kMaxRing) P.ring = kMaxRing;
  SPT_REQUIRE(P.a_stages >= 2 && P.ring >= 2, SPT_E_UNSUPPORTED, "gemm_nt(umma): on-chip budget");
  const size_t smem = misc + (size_t)P.b_slots * b_slot + (size_t)P.epi_bufs * kSlabBytes +
                      (size_t)P.ring * kABytes;

  CUtensorMap tmA, tmB, tmC;
  SPT_REQUIRE(make_map(&tmA, A, M, K, lda, BM) && make_map(&tmB, B, N, K, ldb, P.BN) &&
                  make_map(&tmC, C, M, N, ldc, BM),
              SPT_E_UNSUPPORTED, "gemm_nt(umma): cuTensorMapEncodeTiled failed");

  static unsigned long long attr_done = 0;
  ensure_dynamic_smem(k_gemm_nt_umma, (int)kSmemBudget, &attr_done);
  const int sm_count = device_sm_count();
  const unsigned grid = (unsigned)(P.tiles < sm_count ? P.tiles : sm_count);
  static long long* trace_dev = nullptr;
  static const bool tracing = getenv("SPT_UMMA_TRACE") != nullptr;  // diagnostic only
  if (tracing && !trace_dev) cudaMalloc(&trace_dev, 7 * 32 * 4 * sizeof(long long));
  if (tracing) cudaMemsetAsync(trace_dev, 0, 7 * 32 * 4 * sizeof(long long), stream);
  P.trace = tracing ? trace_dev : nullptr;
  k_gemm_nt_umma<<<grid, kThreads, smem, stream>>>(tmA, tmB, tmC, P);
  if (tracing) {
    static int dumped = 0;
    cudaStreamSynchronize(stream);
    if (dumped++ == 2) {
      static long long h[7 * 32 * 4];
      cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost);
      long long t0 = h[0];
      const char* names[7] = {"prodA", "prodB", "splitA", "splitB", "mma", "epi", "slab"};
      printf("TRACE cfg ring %d a_stages %d b_slots %d resident %d epi_bufs %d smem %zu\n", P.ring,
             P.a_stages, P.b_slots, P.b_resident, P.epi_bufs, smem);
      for (int r = 0; r < 7; ++r)
        for (int i = 0; i < 32; ++i) {
          long long* e = &h[(r * 32 + i) * 4];
          if (!e[0] && !e[1] && !e[2]) continue;
          printf("TRACE %s %d %lld %lld %lld %lld\n", names[r], i, e[0] ? e[0] - t0 : -1,
                 e[1] ? e[1] - t0 : -1, e[2] ? e[2] - t0 : -1, e[3] ? e[3] - t0 : -1);
        }
      fflush(stdout);
    }
  }
  return check_launch("gemm_nt(umma)");
}


// ==========================================================================================
//  C[N,K] += A[M,N]^T . B[M,K] ;  colsumA[N] += column sums of A        (dW / dbias of Linear)
// ==========================================================================================
// The reduction runs over the ROWS of both operands, so both are "MN-major" for the tensor
// core: a [rows x 32 columns] TMA box with SWIZZLE_128B_ATOM_32B is exactly one column block
// of the canonical MN-major tf32 layout (128-byte rows, column blocks LBO apart).  Per chunk of
// `bkm` rows: TMA lands the column blocks of A and B, 8 splitter warps rewrite them in place
// as `hi` and write the `lo` twins (and keep per-thread column sums of A for dbias), the MMA
// warp issues 3 x tcgen05.mma.kind::tf32 per 8-row k-step into TMEM accumulators that live
// for the whole kernel; at the end 4 warps add the [N, K] partial of this CTA to C with
// 16-byte atomics.  Each CTA owns a contiguous range of row chunks.
struct TnParams {
  float* C;
  int64_t ldc;
  float* colsum;      // nullable
  int64_t M;
  int N, K;
  int bkm;            // rows per chunk: 8, 16 or 32
  int ncbA, ncbB;     // 32-column blocks TMA fills for A (= ceil(N/32)) and B (= ceil(Kp/32))
  int n_blocks;       // 128-row blocks of the output (= ceil(N/128))
  int Kp;             // K rounded up to 16 (UMMA N); pieces of <= 256 columns
  int stages;
  int64_t chunks;     // ceil(M / bkm)
  uint32_t blk_bytes; // bkm * 128: one column block of one operand half
  uint32_t a_bytes;   // a_blocks * blk_bytes      (A region, hi or lo)
  int a_blocks;       // 4 per 128-row output block; 1 for a last block of <= 32 real columns
  int grouped;        // N, K <= 32: the four MN blocks hold four consecutive row groups
                      // (block-diagonal batching, 4x fewer MMAs); bkm = rows per group
  uint32_t b_bytes;   // ncbB * blk_bytes          (B region, hi or lo)
};

// MN-major 32-bit operand: the only layout the tensor core takes is SWIZZLE_128B_BASE32B
// (32-byte units XOR-ed with the row index mod 4 inside 128-byte rows; 4-row K atoms), which
// is what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
// LBO = byte stride between 32-column blocks, SBO = 512 (between 4-row groups).
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | (32ull << 32) |
         (1ull << 46) | (1ull << 61);
}

constexpr int kTnSplitWarp0 = 8;    // splitter warps 8..15 (256 threads)
constexpr int kTnMaxStages = 4;

__global__ void __launch_bounds__(kThreads, 1)
k_gemm_tn_umma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const TnParams P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem =
      (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // stage: [A hi | A lo | B hi | B lo]
  const uint32_t stage_bytes = 2 * (P.a_bytes + P.b_bytes);
  uint64_t* bars = (uint64_t*)(smem + (size_t)P.stages * stage_bytes);
  uint64_t* full = bars;                       // [stages] TMA landed the raw chunk
  uint64_t* split = full + kTnMaxStages;       // [stages] hi/lo written (256 arrivals)
  uint64_t* empty = split + kTnMaxStages;      // [stages] MMAs that read the stage retired
  uint64_t* acc_full = empty + kTnMaxStages;   // [1]
  uint32_t* tmem_slot = (uint32_t*)(acc_full + 1);
  float* colsum_s = (float*)(bars + 16);       // [n_blocks * 128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous chunk range of this CTA
  const int64_t per = (P.chunks + gridDim.x - 1) / gridDim.x;
  const int64_t c0 = (int64_t)blockIdx.x * per;
  const int64_t c1 = min(P.chunks, c0 + per);
  const int nchunks = (int)max((int64_t)0, c1 - c0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&split[s], 256);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < P.n_blocks * 128; i += kThreads) colsum_s[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int ksteps = P.bkm / UK;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    int s = 0;
    uint32_t ph = 0;
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(&empty[s], ph ^ 1, 0);
      if (elect_one()) {
        unsigned char* st = smem + (size_t)s * stage_bytes;
        const int rows_per_chunk = P.grouped ? 4 * P.bkm : P.bkm;
        const int row0 = (int)((c0 + c) * rows_per_chunk);
        mbar_expect_tx(&full[s], (uint32_t)(P.ncbA + P.ncbB) * P.blk_bytes);
        unsigned char* sb = st + 2 * P.a_bytes;
        if (P.grouped) {  // block g <- rows of group g, columns 0..31
          for (int g = 0; g < 4; ++g) {
            tma_load_2d(st + (size_t)g * P.blk_bytes, &tmA, 0, row0 + g * P.bkm, &full[s]);
            tma_load_2d(sb + (size_t)g * P.blk_bytes, &tmB, 0, row0 + g * P.bkm, &full[s]);
          }
        } else {
          for (int cb = 0; cb < P.ncbA; ++cb)
            tma_load_2d(st + (size_t)cb * P.blk_bytes, &tmA, cb * 32, row0, &full[s]);
          for (int cb = 0; cb < P.ncbB; ++cb)
            tma_load_2d(sb + (size_t)cb * P.blk_bytes, &tmB, cb * 32, row0, &full[s]);
        }
      }
      __syncwarp();
      if (++s == P.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    int s = 0;
    uint32_t ph = 0;
    const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                ((uint32_t)(BM >> 4) << 24);
    const uint32_t sbase0 = smem_u32(smem);
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(&split[s], ph, 1);
      tc_fence_after();
      const uint32_t sa = sbase0 + (uint32_t)s * stage_bytes;
      const uint32_t sb = sa + 2 * P.a_bytes;
      if (elect_one()) {
        for (int ks = 0; ks < ksteps; ++ks) {
          for (int nb = 0; nb < P.n_blocks; ++nb) {
            // a block of <= 32 real columns is stored once; LBO = 0 makes the other three
            // 32-row groups of the M=128 operand alias it (their output rows are never read)
            const uint32_t lbo_a = (P.ncbA - 4 * nb == 1) ? 0u : P.blk_bytes;
            const uint32_t a_hi_addr = sa + (uint32_t)nb * 4 * P.blk_bytes + ks * 1024;
            const uint64_t a_hi = smem_desc_mn_sw128(a_hi_addr, lbo_a);
            const uint64_t a_lo = smem_desc_mn_sw128(a_hi_addr + P.a_bytes, lbo_a);
            for (int k0 = 0; k0 < P.Kp; k0 += 256) {
              const int nw = min(256, P.Kp - k0);
              const uint32_t idesc = idesc_base | ((uint32_t)(nw >> 3) << 17);
              const uint32_t b_hi_addr = sb + (uint32_t)(k0 >> 5) * P.blk_bytes + ks * 1024;
              const uint64_t b_hi = smem_desc_mn_sw128(b_hi_addr, P.blk_bytes);
              const uint64_t b_lo = smem_desc_mn_sw128(b_hi_addr + P.b_bytes, P.blk_bytes);
              const uint32_t d = tmem_base + (uint32_t)(nb * P.Kp + k0);
              const uint32_t accum = (c | ks) != 0;
              umma_tf32(d, a_lo, b_hi, idesc, accum);
              umma_tf32(d, a_hi, b_lo, idesc, 1);
              umma_tf32(d, a_hi, b_hi, idesc, 1);
            }
          }
        }
        umma_commit(&empty[s]);
        if (c == nchunks - 1) umma_commit(acc_full);
      }
      __syncwarp();
      if (++s == P.stages) { s = 0; ph ^= 1; }
    }
  } else if (warp >= kTnSplitWarp0) {
    // ---------------- splitter (+ column sums of A) ----------------
    const int ts = threadIdx.x - kTnSplitWarp0 * 32;  // 0..255
    const int blk4 = P.bkm * 8;                       // float4 per column block
    const int nA4 = P.ncbA * blk4, nB4 = P.ncbB * blk4;
    // thread-fixed position inside a block: row (ts / 8) % bkm, physical 16-byte chunk ts % 8
    const int prow = (ts >> 3) % P.bkm;
    // logical 4-column group of this thread (32-byte units swizzled with row % 4)
    const int lchunk = ((((ts & 7) >> 1) ^ (prow & 3)) << 1) | (ts & 1);
    float4 cs[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) cs[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
    uint32_t ph = 0;
    for (int c = 0; c < nchunks; ++c) {
      mbar_wait(&full[s], ph, 2);
      unsigned char* st = smem + (size_t)s * stage_bytes;
      const uint32_t ahi = smem_u32(st), alo = ahi + P.a_bytes;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = ts + 256 * u;
        if (i < nA4) {
          const float4 x = lds128(ahi + 16u * (uint32_t)i);
          float4 h, l;
          h.x = tf32_rna(x.x); h.y = tf32_rna(x.y); h.z = tf32_rna(x.z); h.w = tf32_rna(x.w);
          l.x = x.x - h.x; l.y = x.y - h.y;      // raw: the tensor core ignores the low 13 bits
          l.z = x.z - h.z; l.w = x.w - h.w;
          sts128(ahi + 16u * (uint32_t)i, h);
          sts128(alo + 16u * (uint32_t)i, l);
          cs[u].x += x.x; cs[u].y += x.y; cs[u].z += x.z; cs[u].w += x.w;
        }
      }
      split_chunk(st + 2 * P.a_bytes, st + 2 * P.a_bytes + P.b_bytes, nB4, ts, 256);
      fence_proxy_async();
      mbar_arrive(&split[s]);
      if (++s == P.stages) { s = 0; ph ^= 1; }
    }
    if (P.colsum) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = ts + 256 * u;
        if (i < nA4) {
          const int col = (P.grouped ? 0 : (i / blk4) * 32) + lchunk * 4;
          atomicAdd(&colsum_s[col + 0], cs[u].x);
          atomicAdd(&colsum_s[col + 1], cs[u].y);
          atomicAdd(&colsum_s[col + 2], cs[u].z);
          atomicAdd(&colsum_s[col + 3], cs[u].w);
        }
      }
    }
  } else if (warp >= kEpiWarp0 && nchunks > 0) {
    // ---------------- epilogue: TMEM partial -> atomics on C ----------------
    const int q = warp - kEpiWarp0;
    const int row = q * 32 + lane;
    mbar_wait(acc_full, 0, 3);
    tc_fence_after();
    for (int nb = 0; nb < P.n_blocks; ++nb) {
      // grouped: output row `lane` of row group q sits on the diagonal block (q, q)
      const int n = P.grouped ? lane : nb * 128 + row;
      for (int kk = 0; kk < (P.grouped ? 32 : P.Kp); kk += 32) {
        const int k0 = kk;  // column of C
        const int tcol = P.grouped ? 32 * q : nb * P.Kp + kk;
        uint32_t v[32];
        tmem_ld32_nowait(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)tcol, v);
        tmem_wait_ld();
        if (n < P.N) {
          float* crow = P.C + (int64_t)n * P.ldc + k0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (k0 + j + 3 < P.K && (P.ldc & 3) == 0) {  // 16-byte vector reduction
              atomicAdd((float4*)(crow + j),
                        make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                    __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (k0 + j + e < P.K) atomicAdd(crow + j + e, __uint_as_float(v[j + e]));
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (P.colsum && nchunks > 0)
    for (int i = threadIdx.x; i < P.N; i += kThreads) {
      const float v = colsum_s[i];
      if (v != 0.f) atomicAdd(P.colsum + i, v);
    }
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(kTmemCols)
                 : "memory");
  }
}

bool tn_shape_ok(const float* A, int64_t M, int64_t N, int64_t lda, const float* B, int64_t K,
                 int64_t ldb) {
  const uintptr_t al = (uintptr_t)A | (uintptr_t)B;
  if ((al & 15) != 0 || lda % 4 != 0 || ldb % 4 != 0 || M < 1 || M >= (1ll << 31) - 64 ||
      N < 1 || K < 1 || encoder() == nullptr)
    return false;
  const int64_t n_blocks = (N + 127) / 128, Kp = (K + 15) / 16 * 16;
  return n_blocks * Kp <= (int64_t)kTmemCols && N <= 256;  // accumulators + 8 colsum registers
}

int tn_launch(const float* A, int64_t M, int64_t N, int64_t lda, const float* B, int64_t K,
              int64_t ldb, float* C, int64_t ldc, float* colsum, cudaStream_t stream) {
  TnParams P;
  P.C = C;
  P.ldc = ldc;
  P.colsum = colsum;
  P.M = M;
  P.N = (int)N;
  P.K = (int)K;
  P.n_blocks = (int)((N + 127) / 128);
  P.Kp = (int)((K + 15) / 16 * 16);
  P.ncbA = (int)((N + 31) / 32);
  P.ncbB = (P.Kp + 31) / 32;
  P.grouped = (N <= 32 && K <= 32) ? 1 : 0;
  if (P.grouped) {
    P.Kp = 128;
    P.ncbA = P.ncbB = 4;
  }
  const size_t misc = 1024 + 16 * 8 + (size_t)P.n_blocks * 128 * 4;
  P.a_blocks = (P.ncbA - 4 * (P.n_blocks - 1) == 1) ? 4 * (P.n_blocks - 1) + 1 : 4 * P.n_blocks;
  P.bkm = 0;
  // rows per chunk: as many as fit 3 stages (narrow operands take up to 256-row chunks so that
  // the per-chunk barrier round trips amortise); bkm * ncbA <= 256 keeps the per-thread
  // column sums in 8 registers
  for (int bkm = 256; bkm >= 8 && !P.bkm; bkm >>= 1) {
    if (bkm * P.ncbA > 256 && bkm > 8) continue;
    const size_t blk = (size_t)bkm * 128;
    const size_t stage = 2 * ((size_t)P.a_blocks + P.ncbB) * blk;
    const int st = (int)((kSmemBudget - misc) / stage);
    if (st >= 3 || (bkm == 8 && st >= 2)) {
      P.bkm = bkm;
      P.stages = st > kTnMaxStages ? kTnMaxStages : st;
    }
  }
  SPT_REQUIRE(P.bkm != 0, SPT_E_UNSUPPORTED, "gemm_tn_acc(umma): shared memory budget");
  P.blk_bytes = (uint32_t)P.bkm * 128;
  P.a_bytes = (uint32_t)P.a_blocks * P.blk_bytes;
  P.b_bytes = (uint32_t)P.ncbB * P.blk_bytes;
  P.chunks = P.grouped ? (M + 4 * P.bkm - 1) / (4 * P.bkm) : (M + P.bkm - 1) / P.bkm;
  const size_t smem = misc + (size_t)P.stages * 2 * (P.a_bytes + P.b_bytes);

  CUtensorMap tmA, tmB;
  SPT_REQUIRE(make_map(&tmA, A, M, N, lda, P.bkm, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B) &&
                  make_map(&tmB, B, M, K, ldb, P.bkm, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B),
              SPT_E_UNSUPPORTED, "gemm_tn_acc(umma): cuTensorMapEncodeTiled failed");
  static unsigned long long attr_done = 0;
  ensure_dynamic_smem(k_gemm_tn_umma, (int)kSmemBudget, &attr_done);
  const int sm_count = device_sm_count();
  // >= 8 chunks per CTA so that the atomics of the [N, K] partials stay a small fraction
  int64_t grid = P.chunks / 8;
  if (grid < 1) grid = 1;
  if (grid > sm_count) grid = sm_count;
  k_gemm_tn_umma<<<(unsigned)grid, kThreads, smem, stream>>>(tmA, tmB, P);
  return check_launch("gemm_tn_acc(umma)");
}

}  // namespace umma
}  // namespace spt
