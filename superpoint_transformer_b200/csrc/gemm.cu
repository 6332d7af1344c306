// gemm.cu — dense projections of the hot path on the tensor cores, fp32-accurate.
//
// The reference runs qkv / out_proj / RPE / MLP Linears as cuBLAS GEMMs
// (src/nn/attention.py:191,318; src/nn/mlp.py:45), in TF32 in production
// (src/train.py:93-94).  Here every GEMM is "3xTF32": each fp32 operand x is split
// in registers into hi = tf32(x), lo = tf32(x - hi) and three tensor-core MMAs
// (lo*hi + hi*lo + hi*hi, fp32 accumulate) reproduce the fp32 product to ~2^-21
// relative — the 1e-4 parity bar holds while the FLOPs leave the CUDA cores.
// All shapes on this path are skinny (rows = nodes/edges up to millions, K and N
// <= ~300), i.e. HBM-bound: the operand is streamed ONCE (the library alternative of
// three TF32 GEMMs re-streams it three times and measured slower than SGEMM).
//
//   spt_gemm_nt     : C[M,N]  = A[M,K] · B[N,K]^T + bias      (Linear fwd, dX)
//   spt_gemm_tn_acc : C[N,K] += A[M,N]^T · B[M,K], colsum(A)   (dW, dbias)
//
// MMA: legacy mma.sync.m16n8k8 tf32 (HMMA path).  A tcgen05/TMEM version needs the
// hi/lo split materialised in shared memory for the UMMA descriptors; with K <= 300
// these GEMMs are bandwidth-bound either way.
#include "common.cuh"

namespace spt {
namespace gemm {

__device__ __forceinline__ uint32_t tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
  hi = tf32_hi(x);
  lo = tf32_hi(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4],
                                         const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  int bytes = valid ? 16 : 0;   // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------ NT kernel
constexpr int BM = 128, BN = 64, BK = 32, LDS_ = BK + 4;   // +4 floats: conflict-free frags
constexpr int kThreads = 256;

struct NtSmem {
  float A[2][BM][LDS_];    // raw fp32 (split at fragment load)
  float Bh[2][BN][LDS_];   // pre-split weights
  float Bl[2][BN][LDS_];
};

__device__ __forceinline__ void nt_stage(NtSmem& s, int st, const float* __restrict__ A,
                                         int64_t M, int K, int64_t lda,
                                         const float* __restrict__ B, int N, int64_t ldb,
                                         int64_t m0, int n0, int k0) {
  // A tile: 128 rows x 8 chunks of 16 B
#pragma unroll
  for (int i = 0; i < (BM * BK / 4) / kThreads; ++i) {
    int c = threadIdx.x + kThreads * i;
    int row = c >> 3, cc = c & 7;
    int64_t gr = m0 + row;
    int gk = k0 + cc * 4;
    bool ok = (gr < M) && (gk < K);
    const float* src = A + (ok ? gr * lda + gk : 0);
    cp_async16(&s.A[st][row][cc * 4], src, ok);
  }
  cp_async_commit();
  // B tile: 64 rows (n) x 8 chunks, split into hi/lo
#pragma unroll
  for (int i = 0; i < (BN * BK / 4) / kThreads; ++i) {
    int c = threadIdx.x + kThreads * i;
    int row = c >> 3, cc = c & 7;
    int gn = n0 + row, gk = k0 + cc * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gn < N && gk < K) v = *reinterpret_cast<const float4*>(B + (int64_t)gn * ldb + gk);
    uint32_t h[4], l[4];
    split(v.x, h[0], l[0]); split(v.y, h[1], l[1]); split(v.z, h[2], l[2]); split(v.w, h[3], l[3]);
    *reinterpret_cast<uint4*>(&s.Bh[st][row][cc * 4]) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(&s.Bl[st][row][cc * 4]) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

__global__ void __launch_bounds__(kThreads)
k_gemm_nt(const float* __restrict__ A, int64_t M, int K, int64_t lda,
          const float* __restrict__ B, int N, int64_t ldb, const float* __restrict__ bias,
          float* __restrict__ C, int64_t ldc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  NtSmem& s = *reinterpret_cast<NtSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp >> 1, wn = warp & 1;     // 4 x 2 warps -> 32 x 32 warp tiles
  const int g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * BN;
  const int64_t m0 = (int64_t)blockIdx.y * BM;

  float acc[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

  const int nk = (K + BK - 1) / BK;
  nt_stage(s, 0, A, M, K, lda, B, N, ldb, m0, n0, 0);
  for (int kc = 0; kc < nk; ++kc) {
    const int st = kc & 1;
    if (kc + 1 < nk) {
      nt_stage(s, st ^ 1, A, M, K, lda, B, N, ldb, m0, n0, (kc + 1) * BK);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < BK / 8; ++ks) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r = wm * 32 + mt * 16 + g;
        const int c = ks * 8 + t;
        split(s.A[st][r][c], ah[mt][0], al[mt][0]);
        split(s.A[st][r + 8][c], ah[mt][1], al[mt][1]);
        split(s.A[st][r][c + 4], ah[mt][2], al[mt][2]);
        split(s.A[st][r + 8][c + 4], ah[mt][3], al[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int n = wn * 32 + nt * 8 + g;
        const int c = ks * 8 + t;
        uint32_t bh[2], bl[2];
        bh[0] = __float_as_uint(s.Bh[st][n][c]);
        bh[1] = __float_as_uint(s.Bh[st][n][c + 4]);
        bl[0] = __float_as_uint(s.Bl[st][n][c]);
        bl[1] = __float_as_uint(s.Bl[st][n][c + 4]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_tf32(acc[mt][nt], al[mt], bh);   // small terms first
          mma_tf32(acc[mt][nt], ah[mt], bl);
          mma_tf32(acc[mt][nt], ah[mt], bh);
        }
      }
    }
    __syncthreads();
  }
  // epilogue: + bias, store
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int col = n0 + wn * 32 + nt * 8 + 2 * t;
      const float b0 = (bias && col < N) ? bias[col] : 0.f;
      const float b1 = (bias && col + 1 < N) ? bias[col + 1] : 0.f;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int64_t row = m0 + wm * 32 + mt * 16 + g + half * 8;
        if (row >= M) continue;
        const float v0 = acc[mt][nt][half * 2] + b0, v1 = acc[mt][nt][half * 2 + 1] + b1;
        float* dst = C + row * ldc + col;
        if (col + 1 < N && ((ldc & 1) == 0)) {
          *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
        } else {
          if (col < N) dst[0] = v0;
          if (col + 1 < N) dst[1] = v1;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ TN kernel
// C[N,K] += A[M,N]^T B[M,K] ; one CTA = one 64x64 output tile over a slab of rows.
constexpr int TM = 64, TK = 64, TR = 32, LDT = 72;   // 72 mod 32 == 8: conflict-free ^T frags

struct TnSmem {
  float Ah[TR][LDT], Al[TR][LDT];
  float Bh[TR][LDT], Bl[TR][LDT];
  float colsum[TM];
};

__global__ void __launch_bounds__(kThreads)
k_gemm_tn_acc(const float* __restrict__ A, int64_t M, int N, int64_t lda,
              const float* __restrict__ B, int K, int64_t ldb, float* __restrict__ C,
              int64_t ldc, float* __restrict__ colsumA, int ktiles, int64_t rows_per_cta) {
  __shared__ TnSmem s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = warp >> 2, wn = warp & 3;     // 2 x 4 warps -> 32 (n) x 16 (k) warp tiles
  const int g = lane >> 2, t = lane & 3;
  const int tile_n = blockIdx.x / ktiles, tile_k = blockIdx.x - tile_n * ktiles;
  const int n0 = tile_n * TM, k0 = tile_k * TK;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r_end = min(r_begin + rows_per_cta, M);
  const bool do_colsum = (colsumA != nullptr) && (tile_k == 0);

  float acc[2][2][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;
  float cs[4] = {0.f, 0.f, 0.f, 0.f};   // column sums of my 4 A columns

  // staging map: 32 rows x 16 float4 = 512 float4 -> 2 per thread; column chunk fixed
  const int c4 = threadIdx.x & 15;       // float4 column of the 64-wide tile
  const int rr = threadIdx.x >> 4;       // 0..15 -> rows rr and rr + 16

  for (int64_t r0 = r_begin; r0 < r_end; r0 += TR) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int row = rr + 16 * i;
      const int64_t gr = r0 + row;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (gr < r_end) {
        const int gn = n0 + c4 * 4, gk = k0 + c4 * 4;
        if (gn < N) va = ldg_stream4(A + gr * lda + gn);
        if (gk < K) vb = ldg_stream4(B + gr * ldb + gk);
      }
      cs[0] += va.x; cs[1] += va.y; cs[2] += va.z; cs[3] += va.w;
      uint32_t h[4], l[4];
      split(va.x, h[0], l[0]); split(va.y, h[1], l[1]); split(va.z, h[2], l[2]); split(va.w, h[3], l[3]);
      *reinterpret_cast<uint4*>(&s.Ah[row][c4 * 4]) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(&s.Al[row][c4 * 4]) = make_uint4(l[0], l[1], l[2], l[3]);
      split(vb.x, h[0], l[0]); split(vb.y, h[1], l[1]); split(vb.z, h[2], l[2]); split(vb.w, h[3], l[3]);
      *reinterpret_cast<uint4*>(&s.Bh[row][c4 * 4]) = make_uint4(h[0], h[1], h[2], h[3]);
      *reinterpret_cast<uint4*>(&s.Bl[row][c4 * 4]) = make_uint4(l[0], l[1], l[2], l[3]);
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < TR / 8; ++ks) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        // A^T fragment: element (m = n_out, k = row) = A_s[row][n_out]
        const int n = wm * 32 + mt * 16 + g;
        const int r = ks * 8 + t;
        ah[mt][0] = __float_as_uint(s.Ah[r][n]);      al[mt][0] = __float_as_uint(s.Al[r][n]);
        ah[mt][1] = __float_as_uint(s.Ah[r][n + 8]);  al[mt][1] = __float_as_uint(s.Al[r][n + 8]);
        ah[mt][2] = __float_as_uint(s.Ah[r + 4][n]);  al[mt][2] = __float_as_uint(s.Al[r + 4][n]);
        ah[mt][3] = __float_as_uint(s.Ah[r + 4][n + 8]); al[mt][3] = __float_as_uint(s.Al[r + 4][n + 8]);
      }
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int kk = wn * 16 + nt * 8 + g;
        const int r = ks * 8 + t;
        uint32_t bh[2], bl[2];
        bh[0] = __float_as_uint(s.Bh[r][kk]);     bl[0] = __float_as_uint(s.Bl[r][kk]);
        bh[1] = __float_as_uint(s.Bh[r + 4][kk]); bl[1] = __float_as_uint(s.Bl[r + 4][kk]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_tf32(acc[mt][nt], al[mt], bh);
          mma_tf32(acc[mt][nt], ah[mt], bl);
          mma_tf32(acc[mt][nt], ah[mt], bh);
        }
      }
    }
  }
  // flush the 64x64 tile (fp32 atomics: one per output per CTA)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = n0 + wm * 32 + mt * 16 + g + ((r >> 1) << 3);
        const int k = k0 + wn * 16 + nt * 8 + 2 * t + (r & 1);
        if (n < N && k < K) atomicAdd(&C[(int64_t)n * ldc + k], acc[mt][nt][r]);
      }
  if (do_colsum) {
    __syncthreads();
    if (threadIdx.x < TM) s.colsum[threadIdx.x] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(&s.colsum[c4 * 4 + i], cs[i]);
    __syncthreads();
    if (threadIdx.x < TM && n0 + threadIdx.x < N)
      atomicAdd(&colsumA[n0 + threadIdx.x], s.colsum[threadIdx.x]);
  }
}

}  // namespace gemm
}  // namespace spt

using namespace spt;

namespace spt {
namespace umma {  // gemm_umma.cu: tcgen05 / TMEM / TMA path
bool shape_ok(const float* A, int64_t M, int64_t K, int64_t lda, const float* B, int64_t N,
              int64_t ldb, const float* C, int64_t ldc);
int launch(const float* A, int64_t M, int64_t K, int64_t lda, const float* B, int64_t N,
           int64_t ldb, const float* bias, float* C, int64_t ldc, cudaStream_t stream);
bool tn_shape_ok(const float* A, int64_t M, int64_t N, int64_t lda, const float* B, int64_t K,
                 int64_t ldb);
int tn_launch(const float* A, int64_t M, int64_t N, int64_t lda, const float* B, int64_t K,
              int64_t ldb, float* C, int64_t ldc, float* colsum, cudaStream_t stream);
}  // namespace umma
}  // namespace spt

// SPT_GEMM_MMA_SYNC=1 keeps the mma.sync kernel for A/B measurements (debug only)
static bool use_umma() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SPT_GEMM_MMA_SYNC");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

// rows below which the mma.sync kernels of this file run instead of the tcgen05 ones (whose
// prologue — TMEM allocation, barrier set-up, weight split, first TMA round trip — is a fixed
// ~5-15 us); tunable for experiments
static int64_t umma_min_rows(const char* env, int64_t dflt) {
  const char* e = getenv(env);
  if (!e) return dflt;
  const long long v = atoll(e);
  return v > 0 ? (int64_t)v : dflt;
}
static int64_t nt_umma_min_rows() {
  static int64_t v = umma_min_rows("SPT_GEMM_NT_UMMA_MIN", 8192);   // cfg 2: 13.67 -> 13.45 ms vs 512
  return v;
}
static int64_t tn_umma_min_rows() {
  static int64_t v = umma_min_rows("SPT_GEMM_TN_UMMA_MIN", 2048);
  return v;
}

extern "C" {

int spt_gemm_nt(const float* A, int64_t M, int64_t K, int64_t lda, const float* B, int64_t N,
                int64_t ldb, const float* bias, float* C, int64_t ldc, void* stream_) {
  SPT_REQUIRE(M >= 0 && K > 0 && N > 0, SPT_E_INVALID, "gemm_nt: bad sizes");
  if (M == 0) return SPT_OK;
  SPT_REQUIRE(A && B && C, SPT_E_INVALID, "gemm_nt: null pointer");
  SPT_REQUIRE(K % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 &&
                  ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0,
              SPT_E_UNSUPPORTED, "gemm_nt: K, lda, ldb must be multiples of 4 and A, B 16-byte aligned");
  SPT_REQUIRE(K < (1 << 24) && N < (1 << 24), SPT_E_TOO_LARGE, "gemm_nt: K/N too large");
  if (M >= nt_umma_min_rows() && use_umma() && umma::shape_ok(A, M, K, lda, B, N, ldb, C, ldc))
    return umma::launch(A, M, K, lda, B, N, ldb, bias, C, ldc, (cudaStream_t)stream_);
  static unsigned long long attr_done = 0;
  ensure_dynamic_smem(gemm::k_gemm_nt, (int)sizeof(gemm::NtSmem), &attr_done);
  dim3 grid((unsigned)ceil_div(N, gemm::BN), (unsigned)ceil_div(M, gemm::BM));
  SPT_REQUIRE(grid.y <= 65535u * 1024u, SPT_E_TOO_LARGE, "gemm_nt: M too large");
  if (grid.y > 65535) {
    // split the row range so that gridDim.y stays legal
    int64_t rows_per = (int64_t)65535 * gemm::BM;
    for (int64_t off = 0; off < M; off += rows_per) {
      int64_t m = (M - off < rows_per) ? M - off : rows_per;
      dim3 g2(grid.x, (unsigned)ceil_div(m, gemm::BM));
      gemm::k_gemm_nt<<<g2, gemm::kThreads, sizeof(gemm::NtSmem), (cudaStream_t)stream_>>>(
          A + off * lda, m, (int)K, lda, B, (int)N, ldb, bias, C + off * ldc, ldc);
    }
  } else {
    gemm::k_gemm_nt<<<grid, gemm::kThreads, sizeof(gemm::NtSmem), (cudaStream_t)stream_>>>(
        A, M, (int)K, lda, B, (int)N, ldb, bias, C, ldc);
  }
  return check_launch("gemm_nt");
}

int spt_gemm_tn_acc(const float* A, int64_t M, int64_t N, int64_t lda, const float* B,
                    int64_t K, int64_t ldb, float* C, int64_t ldc, float* colsumA,
                    void* stream_) {
  SPT_REQUIRE(M >= 0 && K > 0 && N > 0, SPT_E_INVALID, "gemm_tn_acc: bad sizes");
  if (M == 0) return SPT_OK;
  SPT_REQUIRE(A && B && C, SPT_E_INVALID, "gemm_tn_acc: null pointer");
  SPT_REQUIRE(N % 4 == 0 && K % 4 == 0 && lda % 4 == 0 && ldb % 4 == 0 &&
                  ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0,
              SPT_E_UNSUPPORTED,
              "gemm_tn_acc: N, K, lda, ldb must be multiples of 4 and A, B 16-byte aligned");
  if (M >= tn_umma_min_rows() && use_umma() && umma::tn_shape_ok(A, M, N, lda, B, K, ldb))
    return umma::tn_launch(A, M, N, lda, B, K, ldb, C, ldc, colsumA, (cudaStream_t)stream_);
  int ntiles = (int)ceil_div(N, gemm::TM), ktiles = (int)ceil_div(K, gemm::TK);
  int64_t tiles = (int64_t)ntiles * ktiles;
  // enough row slabs to fill the machine (~4 CTAs per SM), slabs multiple of TR rows
  int64_t slabs = ceil_div((int64_t)device_sm_count() * 4, tiles);
  int64_t rows_per = ceil_div(ceil_div(M, slabs), gemm::TR) * gemm::TR;
  if (rows_per < 256) rows_per = 256;
  slabs = ceil_div(M, rows_per);
  SPT_REQUIRE(slabs <= 65535, SPT_E_TOO_LARGE, "gemm_tn_acc: too many slabs");
  dim3 grid((unsigned)tiles, (unsigned)slabs);
  gemm::k_gemm_tn_acc<<<grid, gemm::kThreads, 0, (cudaStream_t)stream_>>>(
      A, M, (int)N, lda, B, (int)K, ldb, C, ldc, colsumA, ktiles, rows_per);
  return check_launch("gemm_tn_acc");
}

}  // extern "C"
