// gemm_umma.cu — nn.Linear forward / dX on the 5th-generation tensor cores.
//
//   C[M,N] = A[M,K] . B[N,K]^T + bias[N]          (fp32 in, fp32 out, fp32-accurate)
//
// Replaces the cuBLAS calls behind src/nn/attention.py:191,318 and src/nn/mlp.py:45 of the
// reference.  fp32 accuracy on TF32 tensor cores comes from the 3xTF32 split
//   x = hi + lo,  hi = tf32(x),  lo = tf32(x - hi),   A.B ~= Alo.Bhi + Ahi.Blo + Ahi.Bhi
// (~2^-21 relative), accumulated in fp32 in TMEM.
//
// One persistent CTA per SM, warp-specialised, tile = 128 rows x BN (<=256) columns:
//   warp 0      TMA producer : per 32-wide K chunk, one 2-D tiled copy of the A rows and one
//                              of the B rows (SWIZZLE_128B, OOB rows/cols zero-filled)
//   warps 8-11  splitter     : rewrite the raw fp32 chunk in place as `hi`, write `lo` to the
//                              twin buffer (element-wise, so the TMA swizzle is preserved),
//                              fence.proxy.async, arrive
//   warp 1      MMA issuer   : 3 x tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) per 8-wide
//                              k-step out of K-major SWIZZLE_128B smem descriptors;
//                              tcgen05.commit releases the stage / publishes the accumulator
//   warps 4-7   epilogue     : tcgen05.ld 32 lanes x 32 columns, + bias, swizzled st.shared,
//                              TMA store (clipped at M, N); two TMEM accumulators so the
//                              epilogue of tile i overlaps the MMAs of tile i+1
//   warp 2      TMEM allocate / free
// HBM traffic per tile: A once, C once; B chunks are re-read from L2.
#include <cuda.h>  // CUtensorMap types; the encoder is fetched through the runtime API

#include "common.cuh"

namespace spt {
namespace umma {

constexpr int BM = 128;         // UMMA_M
constexpr int BK = 32;          // fp32 per K chunk: one 128-byte swizzle row
constexpr int UK = 8;           // UMMA_K of kind::tf32 (32 bytes)
constexpr int kThreads = 384;   // 12 warps
constexpr int kEpiWarp0 = 4;    // epilogue warps 4..7  (TMEM lane quadrant = warp % 4)
constexpr int kSplitWarp0 = 8;  // splitter warps 8..11
constexpr int kSlab = 32;       // epilogue column slab (32 fp32 = 128 B = one swizzle row)
constexpr int kMaxStages = 4;
constexpr uint32_t kABytes = BM * BK * 4;      // 16 KB
constexpr uint32_t kSlabBytes = BM * kSlab * 4;  // 16 KB
constexpr size_t kSmemBudget = 227 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
               : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag) {
  uint32_t done;
#ifdef SPT_WATCHDOG
  long long t0 = clock64();
#endif
  do {
#ifdef SPT_WATCHDOG
    if (clock64() - t0 > 1000000000LL) {
      printf("umma mbar_wait stuck: block %d thread %d tag %d parity %u\n", blockIdx.x,
             threadIdx.x, tag, parity);
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* src, int c0,
                                             int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::
                   "l"(tm),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                   "r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, kind::tf32, issued by one thread for the whole CTA
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand, SWIZZLE_128B: rows of 128 B, 8-row atoms of 1024 B (SBO), version 1
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) |
         (2ull << 61);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// in place: raw -> hi; twin buffer: lo
__device__ __forceinline__ void split_chunk(float4* hi, float4* lo, int n4, int t) {
  for (int i = t; i < n4; i += 128) {
    const float4 x = hi[i];
    float4 h, l;
    h.x = tf32_rna(x.x); h.y = tf32_rna(x.y); h.z = tf32_rna(x.z); h.w = tf32_rna(x.w);
    l.x = tf32_rna(x.x - h.x); l.y = tf32_rna(x.y - h.y);
    l.z = tf32_rna(x.z - h.z); l.w = tf32_rna(x.w - h.w);
    hi[i] = h;
    lo[i] = l;
  }
}

struct Params {
  const float* bias;  // nullable
  int64_t M;
  int N, K;
  int BN;             // columns per tile, multiple of 16, <= 256
  int n_blocks;       // ceil(N / BN)
  int64_t tiles;      // ceil(M / 128) * n_blocks
  int stages;
  uint32_t tmem_cols; // power of two >= 2 * acc_stride
  uint32_t acc_stride;
};

__global__ void __launch_bounds__(kThreads, 1)
k_gemm_nt_umma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const Params P) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem =
      (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // swizzle atoms
  const uint32_t b_bytes = (uint32_t)P.BN * BK * 4;
  const uint32_t stage_bytes = 2 * kABytes + 2 * b_bytes;
  unsigned char* epi = smem + (size_t)P.stages * stage_bytes;  // 2 x 16 KB, 1024-aligned
  uint64_t* bars = (uint64_t*)(epi + 2 * kSlabBytes);
  uint64_t* full_raw = bars;
  uint64_t* full_split = bars + kMaxStages;
  uint64_t* empty = bars + 2 * kMaxStages;
  uint64_t* tmem_full = bars + 3 * kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KC = (P.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&full_raw[s], 1);
      mbar_init(&full_split[s], 128);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(P.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t t = blockIdx.x; t < P.tiles; t += gridDim.x) {
        const int m_blk = (int)(t / P.n_blocks), n_blk = (int)(t % P.n_blocks);
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % P.stages;
          const uint32_t ph = (it / P.stages) & 1;
          mbar_wait(&empty[s], ph ^ 1, 0);
          unsigned char* st = smem + (size_t)s * stage_bytes;
          mbar_expect_tx(&full_raw[s], kABytes + b_bytes);
          tma_load_2d(st, &tmA, kc * BK, m_blk * BM, &full_raw[s]);
          tma_load_2d(st + 2 * kABytes, &tmB, kc * BK, n_blk * P.BN, &full_raw[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(P.BN >> 3) << 17) |
                             ((uint32_t)(BM >> 4) << 24);
      uint32_t it = 0, tl = 0;
      for (int64_t t = blockIdx.x; t < P.tiles; t += gridDim.x, ++tl) {
        const uint32_t acc = tl & 1, accph = (tl >> 1) & 1;
        mbar_wait(&tmem_empty[acc], accph ^ 1, 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * P.acc_stride;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const int s = it % P.stages;
          const uint32_t ph = (it / P.stages) & 1;
          mbar_wait(&full_split[s], ph, 2);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
          const uint64_t a_hi = smem_desc_sw128(sa), a_lo = smem_desc_sw128(sa + kABytes);
          const uint64_t b_hi = smem_desc_sw128(sa + 2 * kABytes);
          const uint64_t b_lo = smem_desc_sw128(sa + 2 * kABytes + b_bytes);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t o = (uint64_t)(k * UK * 4 >> 4);  // advance inside the swizzle row
            umma_tf32(d, a_lo + o, b_hi + o, idesc, (kc | k) != 0);
            umma_tf32(d, a_hi + o, b_lo + o, idesc, 1);
            umma_tf32(d, a_hi + o, b_hi + o, idesc, 1);
          }
          umma_commit(&empty[s]);  // stage free once these MMAs have read it
        }
        umma_commit(&tmem_full[acc]);
      }
    }
  } else if (warp >= kSplitWarp0) {
    // ---------------- splitter ----------------
    const int ts = threadIdx.x - kSplitWarp0 * 32;
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < P.tiles; t += gridDim.x) {
      for (int kc = 0; kc < KC; ++kc, ++it) {
        const int s = it % P.stages;
        const uint32_t ph = (it / P.stages) & 1;
        mbar_wait(&full_raw[s], ph, 3);
        unsigned char* st = smem + (size_t)s * stage_bytes;
        split_chunk((float4*)st, (float4*)(st + kABytes), kABytes / 16, ts);
        split_chunk((float4*)(st + 2 * kABytes), (float4*)(st + 2 * kABytes + b_bytes),
                    b_bytes / 16, ts);
        fence_proxy_async();
        mbar_arrive(&full_split[s]);
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ---------------- epilogue ----------------
    const int q = warp - kEpiWarp0;        // TMEM lane quadrant
    const int et = threadIdx.x - kEpiWarp0 * 32;
    const int row = q * 32 + lane;         // tile-local row == TMEM lane
    uint32_t tl = 0, sc = 0;
    for (int64_t t = blockIdx.x; t < P.tiles; t += gridDim.x, ++tl) {
      const int m_blk = (int)(t / P.n_blocks), n_blk = (int)(t % P.n_blocks);
      const uint32_t acc = tl & 1, accph = (tl >> 1) & 1;
      mbar_wait(&tmem_full[acc], accph, 4);
      tc_fence_after();
      const int ncols = min(P.BN, P.N - n_blk * P.BN);
      const int nslab = (ncols + kSlab - 1) / kSlab;
      for (int sl = 0; sl < nslab; ++sl, ++sc) {
        unsigned char* buf = epi + (sc & 1) * kSlabBytes;
        if (et == 0) bulk_wait_read<1>();  // the store that last used `buf` has read it
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint32_t v[32];
        tmem_ld32(tmem_base + acc * P.acc_stride + ((uint32_t)(q * 32) << 16) + sl * kSlab, v);
        const int c0 = n_blk * P.BN + sl * kSlab;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 o;
          o.x = __uint_as_float(v[4 * j + 0]);
          o.y = __uint_as_float(v[4 * j + 1]);
          o.z = __uint_as_float(v[4 * j + 2]);
          o.w = __uint_as_float(v[4 * j + 3]);
          if (P.bias) {
            const int c = c0 + 4 * j;
            if (c + 0 < P.N) o.x += __ldg(P.bias + c + 0);
            if (c + 1 < P.N) o.y += __ldg(P.bias + c + 1);
            if (c + 2 < P.N) o.z += __ldg(P.bias + c + 2);
            if (c + 3 < P.N) o.w += __ldg(P.bias + c + 3);
          }
          *(float4*)(buf + row * 128 + ((j ^ (row & 7)) << 4)) = o;
        }
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          tma_store_2d(&tmC, buf, c0, m_blk * BM);
          bulk_commit();
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
    if (et == 0) bulk_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(P.tmem_cols)
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
                     int box_rows) {
  EncodeTiledFn enc = encoder();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides,
             box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

// layout the tensor-core path needs; otherwise the caller uses the mma.sync kernel
bool shape_ok(const float* A, int64_t M, int64_t K, int64_t lda, const float* B, int64_t N,
              int64_t ldb, const float* C, int64_t ldc) {
  const uintptr_t al = (uintptr_t)A | (uintptr_t)B | (uintptr_t)C;
  return (al & 15) == 0 && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0 && M >= 1 &&
         M < (1ll << 31) - 256 && N >= 1 && N < (1 << 30) && K >= 1 && K < (1 << 30) &&
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
  P.BN = n16 <= 256 ? n16 : 128;
  P.n_blocks = (int)((N + P.BN - 1) / P.BN);
  P.tiles = ((M + BM - 1) / BM) * P.n_blocks;
  P.acc_stride = (uint32_t)((P.BN + 31) / 32 * 32);
  uint32_t cols = 32;
  while (cols < 2 * P.acc_stride) cols <<= 1;
  P.tmem_cols = cols;
  const size_t stage_bytes = 2 * (size_t)kABytes + 2 * (size_t)P.BN * BK * 4;
  const size_t fixed = 1024 + 2 * (size_t)kSlabBytes + 256;
  int stages = (int)((kSmemBudget - fixed) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  SPT_REQUIRE(stages >= 2, SPT_E_UNSUPPORTED, "gemm_nt(umma): shared memory budget");
  P.stages = stages;
  const size_t smem = fixed + (size_t)stages * stage_bytes;

  CUtensorMap tmA, tmB, tmC;
  SPT_REQUIRE(make_map(&tmA, A, M, K, lda, BM) && make_map(&tmB, B, N, K, ldb, P.BN) &&
                  make_map(&tmC, C, M, N, ldc, BM),
              SPT_E_UNSUPPORTED, "gemm_nt(umma): cuTensorMapEncodeTiled failed");

  static int sm_count = 0;
  static bool attr = false;
  if (!attr) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(k_gemm_nt_umma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)kSmemBudget);
    attr = true;
  }
  const unsigned grid = (unsigned)(P.tiles < sm_count ? P.tiles : sm_count);
  k_gemm_nt_umma<<<grid, kThreads, smem, stream>>>(tmA, tmB, tmC, P);
  return check_launch("gemm_nt(umma)");
}

}  // namespace umma
}  // namespace spt
