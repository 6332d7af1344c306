// vrpe.cu — glue kernels of the value-RPE path of SelfAttentionBlock (reference
// src/nn/attention.py:294-301).  The attention kernels emit abar = sum_e p_e a_e [N, H, F] and
// sump = sum_e p_e [N, H]; the value RPE is then applied algebraically,
//     y[n, h*Dv + d] = agg[n, h*Dv + d] + sum_f Wv[h*Dv + d, f] abar[n, h, f] + bv[h*Dv + d] sump[n, h],
// i.e. ONE dense [N, H*F] x [H*F, C] product with a block-diagonal weight (tcgen05 gemm_nt)
// plus the kernels below, which replace ~25 framework launches per block and step
// (block_diag / repeat_interleave / mul / add / slice / sum) by 3 (forward) + 3 (backward).
#include "common.cuh"

namespace spt {

// Wbd [C, H*F] (transpose = 0) or its transpose [H*F, C] (transpose = 1) from Wv:
//   Wbd[h*Dv + d, h'*F + f] = (h == h') ? Wv[(share ? d : h*Dv + d), f] : 0
__global__ void k_vrpe_blockdiag(const float* __restrict__ Wv, int H, int Dv, int F, int share,
                                 int transpose, float* __restrict__ out) {
  const int C = H * Dv, HF = H * F;
  const int total = C * HF;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int c, j;
    if (transpose) { j = i / C; c = i - j * C; } else { c = i / HF; j = i - c * HF; }
    const int h = c / Dv, hp = j / F, f = j - hp * F;
    out[i] = (h == hp) ? Wv[(size_t)(share ? (c - h * Dv) : c) * F + f] : 0.f;
  }
}

// y = agg + rv + sump (x) bv     (bv may be null; rv may be null)
__global__ void __launch_bounds__(256)
k_vrpe_epilogue(const float* __restrict__ agg, const float* __restrict__ rv,
                const float* __restrict__ sump, const float* __restrict__ bv, int64_t N, int C,
                int H, int Dv, int share, float* __restrict__ y) {
  const int64_t total4 = N * (C / 4);
  const int c4n = C / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / c4n;
    const int c = (int)(i - n * c4n) * 4;
    float4 a = *reinterpret_cast<const float4*>(agg + n * C + c);
    if (rv) {
      const float4 r = *reinterpret_cast<const float4*>(rv + n * C + c);
      a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
    }
    if (bv) {
      float o[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cc = c + j, h = cc / Dv;
        o[j] = fmaf(sump[n * H + h], bv[share ? cc - h * Dv : cc], o[j]);
      }
      a = make_float4(o[0], o[1], o[2], o[3]);
    }
    *reinterpret_cast<float4*>(y + n * C + c) = a;
  }
}

// dbv[c'] += sum_n dy[n, c] sump[n, h(c)]  (c' = c, or d when the heads share the encoder):
// persistent CTAs, thread = channel, one fp32 atomic per channel per CTA.
__global__ void __launch_bounds__(256)
k_vrpe_dbias(const float* __restrict__ dy, const float* __restrict__ sump, int64_t N, int C, int H,
             int Dv, int share, float* __restrict__ dbv) {
  const int c = threadIdx.x % C;           // blockDim.x is a multiple of C (C <= 256)
  const int rl = threadIdx.x / C, nrl = blockDim.x / C;
  const int h = c / Dv;
  float acc = 0.f;
  for (int64_t n = (int64_t)blockIdx.x * nrl + rl; n < N; n += (int64_t)gridDim.x * nrl)
    acc = fmaf(dy[n * C + c], sump[n * H + h], acc);
  atomicAdd(&dbv[share ? c - h * Dv : c], acc);
}

// dWv[(share ? d : h*Dv + d), f] += dWbd[h*Dv + d, h*F + f]   (diagonal blocks of the dense grad)
__global__ void k_vrpe_dweight(const float* __restrict__ dWbd, int H, int Dv, int F, int share,
                               float* __restrict__ dWv) {
  const int C = H * Dv, HF = H * F;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C * F; i += gridDim.x * blockDim.x) {
    const int c = i / F, f = i - c * F, h = c / Dv;
    const float v = dWbd[(size_t)c * HF + h * F + f];
    if (share) atomicAdd(&dWv[(size_t)(c - h * Dv) * F + f], v);
    else dWv[i] += v;
  }
}

// fp32 -> bf16 (round to nearest even), 2 elements per thread
__global__ void k_cast_bf16(const float* __restrict__ x, int64_t n2, int64_t n,
                            uint32_t* __restrict__ out2, uint16_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n2; i += stride) {
    const float2 v = reinterpret_cast<const float2*>(x)[i];
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(v.y), "f"(v.x));
    out2[i] = r;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && (n & 1)) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(0.f), "f"(x[n - 1]));
    out[n - 1] = (uint16_t)(r & 0xffffu);
  }
}

}  // namespace spt

using namespace spt;

extern "C" {

int spt_vrpe_blockdiag(const float* Wv, int H, int Dv, int F, int heads_share, int transpose,
                       float* out, void* stream_) {
  SPT_REQUIRE(H >= 1 && Dv >= 1 && F >= 1, SPT_E_INVALID, "vrpe_blockdiag: bad dims");
  SPT_REQUIRE(Wv && out, SPT_E_INVALID, "vrpe_blockdiag: null pointer");
  const int total = H * Dv * H * F;
  k_vrpe_blockdiag<<<(unsigned)ceil_div(total, 256), 256, 0, (cudaStream_t)stream_>>>(
      Wv, H, Dv, F, heads_share, transpose, out);
  return check_launch("vrpe_blockdiag");
}

int spt_vrpe_epilogue(const float* agg, const float* rv, const float* sump, const float* bv,
                      int64_t N, int H, int Dv, int heads_share, float* y, void* stream_) {
  SPT_REQUIRE(N >= 0 && H >= 1 && Dv >= 1, SPT_E_INVALID, "vrpe_epilogue: bad dims");
  if (N == 0) return SPT_OK;
  const int C = H * Dv;
  SPT_REQUIRE(C % 4 == 0, SPT_E_UNSUPPORTED, "vrpe_epilogue: C must be a multiple of 4");
  SPT_REQUIRE(agg && y && (!bv || sump), SPT_E_INVALID, "vrpe_epilogue: null pointer");
  int64_t blocks = ceil_div(N * (C / 4), 256);
  if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
  k_vrpe_epilogue<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(agg, rv, sump, bv, N, C, H,
                                                                        Dv, heads_share, y);
  return check_launch("vrpe_epilogue");
}

int spt_vrpe_bwd_params(const float* dy, const float* sump, const float* dWbd, int64_t N, int H,
                        int Dv, int F, int heads_share, float* dWv, float* dbv, void* stream_) {
  SPT_REQUIRE(N >= 0 && H >= 1 && Dv >= 1 && F >= 1, SPT_E_INVALID, "vrpe_bwd_params: bad dims");
  const int C = H * Dv;
  SPT_REQUIRE(C <= 256, SPT_E_UNSUPPORTED, "vrpe_bwd_params: C <= 256");
  cudaStream_t st = (cudaStream_t)stream_;
  if (dbv && N > 0) {
    SPT_REQUIRE(dy && sump, SPT_E_INVALID, "vrpe_bwd_params: null pointer");
    const int threads = (256 / C) * C;
    int64_t blocks = ceil_div(N, (int64_t)(threads / C) * 64);
    if (blocks > device_sm_count() * 4) blocks = device_sm_count() * 4;
    if (blocks < 1) blocks = 1;
    k_vrpe_dbias<<<(unsigned)blocks, threads, 0, st>>>(dy, sump, N, C, H, Dv, heads_share, dbv);
  }
  if (dWv) {
    SPT_REQUIRE(dWbd, SPT_E_INVALID, "vrpe_bwd_params: null pointer");
    k_vrpe_dweight<<<(unsigned)ceil_div(C * F, 256), 256, 0, st>>>(dWbd, H, Dv, F, heads_share,
                                                                   dWv);
  }
  return check_launch("vrpe_bwd_params");
}

int spt_cast_bf16(const float* x, int64_t n, uint16_t* out, void* stream_) {
  SPT_REQUIRE(n >= 0, SPT_E_INVALID, "cast_bf16: negative size");
  if (n == 0) return SPT_OK;
  SPT_REQUIRE(x && out, SPT_E_INVALID, "cast_bf16: null pointer");
  SPT_REQUIRE(((uintptr_t)x & 7) == 0 && ((uintptr_t)out & 3) == 0, SPT_E_INVALID,
              "cast_bf16: x must be 8-byte and out 4-byte aligned");
  const int64_t n2 = n >> 1;
  int64_t blocks = ceil_div(n2 > 0 ? n2 : 1, 256);
  if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
  k_cast_bf16<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(
      x, n2, n, reinterpret_cast<uint32_t*>(out), out);
  return check_launch("cast_bf16");
}

}  // extern "C"
