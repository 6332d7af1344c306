// norm.cu — GraphNorm forward/backward (torch_geometric.nn.norm.GraphNorm semantics,
// SURVEY.md Appendix A) for few, huge segments (one per cloud in the batch).
//
//   mean[b]  = sum_{i in b} x_i / max(n_b, 1)
//   out_i    = x_i - mean_scale * mean[b]
//   var[b]   = sum out_i^2 / max(n_b, 1)
//   y_i      = weight * out_i / sqrt(var[b] + eps) + bias
//
// Two-pass statistics (mean, then centred second moment) like the reference, so
// there is no E[x^2]-E[x]^2 cancellation.  Each CTA reduces a slab of rows in
// registers/shared memory and flushes per-column partials with fp64 atomics into
// a [B, C] accumulator (B is tiny; the atomics are ~C per CTA).  `batch` may be
// unsorted (edge-level norms use norm_index[edge_index[0]], src/models/components/
// spt.py:829-835): threads flush whenever the segment id of their row changes.
#include "common.cuh"
#include <stdlib.h>

namespace spt {

constexpr int kNormThreads = 256;
constexpr int kNormRows = 128;  // minimum rows per CTA slab (the launch picks the slab size)

struct ColMap {
  int tx;  // threads along columns (each VEC wide)
  int ty;  // row lanes
};

// Column-sliced slab reduction shared by the three statistics passes.
//   MODE 0: acc0 += x                         (+ row count)
//   MODE 1: acc0 += (x - mean_scale*mean)^2
//   MODE 2: acc0 += dy*xhat ; acc1 += dy      (backward)
//   MODE 3: acc0 += d ; acc1 += d*d, d = x - pivot[c]   (single-pass forward statistics:
//           `mean_scale` carries the per-channel pivot, a sample mean of the first rows, so
//           the fp32 partial sums hold O(sigma) quantities and the shifted-moment algebra of
//           the finalise kernel has nothing to cancel)
// Thread (cx, ry) owns VEC columns and every ty-th row of the slab.  When the
// whole slab belongs to one segment (the common case: `batch` sorted, slabs much
// smaller than graphs) the ty row-lanes are reduced through shared memory and only
// tx*VEC fp64 atomics leave the CTA; otherwise each thread flushes on every
// segment change.
template <int MODE, int VEC>
__global__ void __launch_bounds__(kNormThreads)
k_graphnorm_stats(const float* __restrict__ x, const float* __restrict__ dy,
                  const float* __restrict__ yact, float slope,
                  const int64_t* __restrict__ batch, int64_t N, int64_t C, int64_t B,
                  const float* __restrict__ mean_scale,
                  const double* __restrict__ sum_x, const double* __restrict__ count,
                  const float* __restrict__ mean, const float* __restrict__ rstd,
                  double* __restrict__ acc0 /*[B,C]*/, double* __restrict__ acc1 /*[B,C]*/,
                  double* __restrict__ cnt_out /*[B]*/, int tx, int ty, int slab_rows,
                  float* __restrict__ pivot_out = nullptr, int pivot_rows = 0) {
  constexpr int NACC = (MODE >= 2) ? 2 : 1;
  __shared__ float red[NACC][kNormThreads * VEC];
  int cx = threadIdx.x % tx;
  int ry = threadIdx.x / tx;
  // persistent CTA: slabs blockIdx.x, blockIdx.x + gridDim.x, ... ; partial sums
  // stay in registers across slabs and leave the CTA once (few fp64 atomics)
  const int64_t nslabs = (N + slab_rows - 1) / slab_rows;
  const int64_t first_row = (int64_t)blockIdx.x * slab_rows;
  int64_t first_b = batch ? batch[first_row < N ? first_row : 0] : 0;
  for (int64_t ct = 0; ct < C; ct += (int64_t)tx * VEC) {
    int64_t c0 = ct + (int64_t)cx * VEC;
    bool active = c0 < C;
    float a[NACC][VEC], mu[VEC], rs[VEC], ms[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      a[0][v] = 0.f;
      if (NACC == 2) a[NACC - 1][v] = 0.f;
      mu[v] = 0.f;
      rs[v] = 1.f;
      ms[v] = (MODE != 0 && active) ? (mean_scale ? mean_scale[c0 + v] : 1.f) : 0.f;
    }
    if (MODE == 3 && pivot_out) {
      // the per-channel pivot (mean of the first rows) computed by every CTA for itself — the
      // same loads in the same order everywhere, so all CTAs hold identical bits — instead of
      // a single-CTA kernel in front of this one; CTA 0 publishes it for the finalisation
      float part[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) part[v] = 0.f;
      if (active) {
        for (int r = ry; r < pivot_rows; r += ty) {
          if (VEC == 4) {
            const float4 t = *reinterpret_cast<const float4*>(x + (int64_t)r * C + c0);
            part[0] += t.x; part[1 % VEC] += t.y; part[2 % VEC] += t.z; part[3 % VEC] += t.w;
          } else {
            part[0] += x[(int64_t)r * C + c0];
          }
        }
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) red[0][threadIdx.x * VEC + v] = part[v];
      __syncthreads();
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float t = 0.f;
        for (int yy = 0; yy < ty; ++yy) t += red[0][(yy * tx + cx) * VEC + v];
        ms[v] = active ? t / (float)pivot_rows : 0.f;
      }
      __syncthreads();
      if (blockIdx.x == 0 && ry == 0 && active) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) pivot_out[c0 + v] = ms[v];
      }
    }
    int64_t cur = -1;
    int nrows = 0;
    bool uniform = true;
    int64_t rows_seen = 0;
    if (active) {
     for (int64_t slab = blockIdx.x; slab < nslabs; slab += gridDim.x) {
      const int64_t r0 = slab * slab_rows;
      const int64_t r1 = min(r0 + (int64_t)slab_rows, N);
      rows_seen += r1 - r0;
      constexpr int U = 4;   // rows in flight per thread (independent 16-byte loads)
      for (int64_t rb = r0 + ry; rb < r1; rb += (int64_t)ty * U) {
        int64_t bq[U];
        float xq[U][VEC], gq[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t r = rb + (int64_t)u * ty;
          bq[u] = -1;
          if (r < r1) {
            bq[u] = batch ? batch[r] : 0;
            if (VEC == 4) {
              float4 t = *reinterpret_cast<const float4*>(x + r * C + c0);
              xq[u][0] = t.x; xq[u][1 % VEC] = t.y; xq[u][2 % VEC] = t.z; xq[u][3 % VEC] = t.w;
              if (MODE == 2) {
                float4 g = *reinterpret_cast<const float4*>(dy + r * C + c0);
                if (yact) {   // fused LeakyReLU: d(pre-activation) = dy * (y > 0 ? 1 : slope)
                  float4 yy = *reinterpret_cast<const float4*>(yact + r * C + c0);
                  g.x *= (yy.x > 0.f) ? 1.f : slope; g.y *= (yy.y > 0.f) ? 1.f : slope;
                  g.z *= (yy.z > 0.f) ? 1.f : slope; g.w *= (yy.w > 0.f) ? 1.f : slope;
                }
                gq[u][0] = g.x; gq[u][1 % VEC] = g.y; gq[u][2 % VEC] = g.z; gq[u][3 % VEC] = g.w;
              }
            } else {
              xq[u][0] = x[r * C + c0];
              if (MODE == 2) {
                float g = dy[r * C + c0];
                if (yact) g *= (yact[r * C + c0] > 0.f) ? 1.f : slope;
                gq[u][0] = g;
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
        const int64_t b = bq[u];
        if (b < 0 || b >= B) continue;
        if (b != first_b) uniform = false;
        if (b != cur) {
          if (cur >= 0) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              atomicAdd(&acc0[cur * C + c0 + v], (double)a[0][v]);
              if (NACC == 2) atomicAdd(&acc1[cur * C + c0 + v], (double)a[NACC - 1][v]);
            }
            if (MODE != 1 && cnt_out && c0 == 0) atomicAdd(&cnt_out[cur], (double)nrows);
          }
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            a[0][v] = 0.f;
            if (NACC == 2) a[NACC - 1][v] = 0.f;
          }
          nrows = 0;
          cur = b;
          if (MODE == 1) {
            double n = fmax(count[b], 1.0);
#pragma unroll
            for (int v = 0; v < VEC; ++v)
              mu[v] = ms[v] * (float)(sum_x[b * C + c0 + v] / n);
          } else if (MODE == 2) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
              mu[v] = ms[v] * mean[b * C + c0 + v];
              rs[v] = rstd[b * C + c0 + v];
            }
          }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          if (MODE == 0) {
            a[0][v] += xq[u][v];
          } else if (MODE == 3) {
            float d = xq[u][v] - ms[v];
            a[0][v] += d;
            a[NACC - 1][v] = fmaf(d, d, a[NACC - 1][v]);
          } else if (MODE == 1) {
            float d = xq[u][v] - mu[v];
            a[0][v] = fmaf(d, d, a[0][v]);
          } else {
            float xhat = (xq[u][v] - mu[v]) * rs[v];
            a[0][v] = fmaf(gq[u][v], xhat, a[0][v]);
            a[NACC - 1][v] += gq[u][v];
          }
        }
        ++nrows;
        }
      }
     }
    }
    // CTA-uniform decision: every row of this CTA's slabs in segment first_b?
    int all_uniform = __syncthreads_and(uniform ? 1 : 0);
    if (all_uniform && ty > 1) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        red[0][threadIdx.x * VEC + v] = a[0][v];
        if (NACC == 2) red[NACC - 1][threadIdx.x * VEC + v] = a[NACC - 1][v];
      }
      __syncthreads();
      if (ry == 0 && active && first_b >= 0 && first_b < B) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          float s0 = 0.f, s1 = 0.f;
          for (int y = 0; y < ty; ++y) {
            s0 += red[0][(y * tx + cx) * VEC + v];
            if (NACC == 2) s1 += red[NACC - 1][(y * tx + cx) * VEC + v];
          }
          atomicAdd(&acc0[first_b * C + c0 + v], (double)s0);
          if (NACC == 2) atomicAdd(&acc1[first_b * C + c0 + v], (double)s1);
        }
        if (MODE != 1 && cnt_out && c0 == 0) atomicAdd(&cnt_out[first_b], (double)rows_seen);
      }
      __syncthreads();
    } else if (active && cur >= 0) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        atomicAdd(&acc0[cur * C + c0 + v], (double)a[0][v]);
        if (NACC == 2) atomicAdd(&acc1[cur * C + c0 + v], (double)a[NACC - 1][v]);
      }
      if (MODE != 1 && cnt_out && c0 == 0) atomicAdd(&cnt_out[cur], (double)nrows);
    }
  }
}

// The statistics passes for ONE graph (batch == nullptr, C % 4 == 0): no segment bookkeeping per
// row, 8 (forward) / 4 (backward, three streams) independent 16-byte loads in flight per thread.
//   MODE 3: acc0 += d, acc1 += d*d, d = x - pivot (pivot = mean of the first rows, computed here)
//   MODE 2: acc0 += g*xhat, acc1 += g, g = dy * act'(y)
template <int MODE>
__global__ void __launch_bounds__(kNormThreads)
k_graphnorm_stats_single(const float* __restrict__ x, const float* __restrict__ dy,
                         const float* __restrict__ yact, float slope, int64_t N, int C,
                         const float* __restrict__ mean_scale, const float* __restrict__ mean,
                         const float* __restrict__ rstd, double* __restrict__ acc0,
                         double* __restrict__ acc1, double* __restrict__ cnt_out, int tx, int ty,
                         int slab_rows, float* __restrict__ pivot_out, int pivot_rows) {
  __shared__ float red[2][kNormThreads * 4];
  const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
  const int64_t nslabs = (N + slab_rows - 1) / slab_rows;
  constexpr int U = (MODE == 3) ? 8 : 4;
  for (int ct = 0; ct < C; ct += tx * 4) {
    const int c0 = ct + cx * 4;
    const bool active = c0 < C;
    float4 sh = make_float4(0.f, 0.f, 0.f, 0.f), rs = make_float4(1.f, 1.f, 1.f, 1.f);
    if (MODE == 3) {
      float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
      if (active)
        for (int r = ry; r < pivot_rows; r += ty) {
          const float4 t = *reinterpret_cast<const float4*>(x + (int64_t)r * C + c0);
          part.x += t.x; part.y += t.y; part.z += t.z; part.w += t.w;
        }
      *reinterpret_cast<float4*>(&red[0][threadIdx.x * 4]) = part;
      __syncthreads();
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int yy = 0; yy < ty; ++yy) {
        const float4 q = *reinterpret_cast<const float4*>(&red[0][(yy * tx + cx) * 4]);
        t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w;
      }
      const float inv = 1.f / (float)pivot_rows;
      sh = make_float4(t.x * inv, t.y * inv, t.z * inv, t.w * inv);
      __syncthreads();
      if (blockIdx.x == 0 && ry == 0 && active)
        *reinterpret_cast<float4*>(pivot_out + c0) = sh;
    } else if (active) {
      const float4 ms = *reinterpret_cast<const float4*>(mean_scale + c0);
      const float4 mu = *reinterpret_cast<const float4*>(mean + c0);
      rs = *reinterpret_cast<const float4*>(rstd + c0);
      sh = make_float4(ms.x * mu.x, ms.y * mu.y, ms.z * mu.z, ms.w * mu.w);
    }
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
    int64_t rows_seen = 0;
    if (active) {
      for (int64_t slab = blockIdx.x; slab < nslabs; slab += gridDim.x) {
        const int64_t r0 = slab * slab_rows, r1 = min(r0 + (int64_t)slab_rows, N);
        rows_seen += r1 - r0;
        for (int64_t rb = r0 + ry; rb < r1; rb += (int64_t)ty * U) {
          float4 xq[U], gq[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int64_t r = rb + (int64_t)u * ty;
            // rows past the slab contribute zeros: x = shift, g = 0
            xq[u] = sh;
            gq[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < r1) {
              xq[u] = *reinterpret_cast<const float4*>(x + r * C + c0);
              if (MODE == 2) {
                gq[u] = *reinterpret_cast<const float4*>(dy + r * C + c0);
                if (yact) {
                  const float4 yy = *reinterpret_cast<const float4*>(yact + r * C + c0);
                  gq[u].x *= (yy.x > 0.f) ? 1.f : slope; gq[u].y *= (yy.y > 0.f) ? 1.f : slope;
                  gq[u].z *= (yy.z > 0.f) ? 1.f : slope; gq[u].w *= (yy.w > 0.f) ? 1.f : slope;
                }
              }
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            if (MODE == 3) {
              const float dx_ = xq[u].x - sh.x, dy_ = xq[u].y - sh.y, dz_ = xq[u].z - sh.z,
                          dw_ = xq[u].w - sh.w;
              a0.x += dx_; a0.y += dy_; a0.z += dz_; a0.w += dw_;
              a1.x = fmaf(dx_, dx_, a1.x); a1.y = fmaf(dy_, dy_, a1.y);
              a1.z = fmaf(dz_, dz_, a1.z); a1.w = fmaf(dw_, dw_, a1.w);
            } else {
              a0.x = fmaf(gq[u].x, (xq[u].x - sh.x) * rs.x, a0.x);
              a0.y = fmaf(gq[u].y, (xq[u].y - sh.y) * rs.y, a0.y);
              a0.z = fmaf(gq[u].z, (xq[u].z - sh.z) * rs.z, a0.z);
              a0.w = fmaf(gq[u].w, (xq[u].w - sh.w) * rs.w, a0.w);
              a1.x += gq[u].x; a1.y += gq[u].y; a1.z += gq[u].z; a1.w += gq[u].w;
            }
          }
        }
      }
    }
    *reinterpret_cast<float4*>(&red[0][threadIdx.x * 4]) = a0;
    *reinterpret_cast<float4*>(&red[1][threadIdx.x * 4]) = a1;
    __syncthreads();
    if (ry == 0 && active) {
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
      for (int yy = 0; yy < ty; ++yy) {
        const float4 p0 = *reinterpret_cast<const float4*>(&red[0][(yy * tx + cx) * 4]);
        const float4 p1 = *reinterpret_cast<const float4*>(&red[1][(yy * tx + cx) * 4]);
        s0[0] += p0.x; s0[1] += p0.y; s0[2] += p0.z; s0[3] += p0.w;
        s1[0] += p1.x; s1[1] += p1.y; s1[2] += p1.z; s1[3] += p1.w;
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        atomicAdd(&acc0[c0 + v], (double)s0[v]);
        atomicAdd(&acc1[c0 + v], (double)s1[v]);
      }
      if (cnt_out && c0 == 0) atomicAdd(&cnt_out[0], (double)rows_seen);
    }
    __syncthreads();
  }
}

// per-channel pivot = mean of the first `rows` rows (any graph): one thread per channel
__global__ void k_graphnorm_pivot(const float* __restrict__ x, int64_t rows, int64_t C,
                                  float* __restrict__ pivot) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int64_t r = 0; r < rows; ++r) s += x[r * C + c];
  pivot[c] = rows > 0 ? s / (float)rows : 0.f;
}

// [B,C] finalisation from shifted moments S1 = sum(x - p), S2 = sum((x - p)^2):
//   mean = p + S1/n ;  sum((x - ms*mean)^2)/n = S2/n + 2 q S1/n + q^2,  q = p - ms*mean
__global__ void k_graphnorm_finalize_shifted(const double* __restrict__ s1,
                                             const double* __restrict__ s2,
                                             const double* __restrict__ count,
                                             const float* __restrict__ pivot,
                                             const float* __restrict__ mean_scale, int64_t B,
                                             int64_t C, float eps, float* __restrict__ mean,
                                             float* __restrict__ rstd) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int64_t b = i / C, c = i - b * C;
  const double n = fmax(count[b], 1.0);
  const double p = pivot[c], m1 = s1[i] / n, m2 = s2[i] / n;
  const double mu = p + m1;
  const double q = p - (double)mean_scale[c] * mu;
  const double var = fmax(m2 + 2.0 * q * m1 + q * q, 0.0);
  mean[i] = (float)mu;
  rstd[i] = (float)(1.0 / sqrt(var + (double)eps));
}

template <int VEC>
__global__ void __launch_bounds__(kNormThreads)
k_graphnorm_apply(const float* __restrict__ x, const int64_t* __restrict__ batch,
                  int64_t N, int64_t C, int64_t B, const float* __restrict__ weight,
                  const float* __restrict__ bias, const float* __restrict__ mean_scale,
                  const float* __restrict__ mean, const float* __restrict__ rstd,
                  float slope, float* __restrict__ y) {
  int64_t cv = C / VEC;
  int64_t total = N * cv;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int64_t r = i / cv;
    int64_t c0 = (i - r * cv) * VEC;
    int64_t b = batch ? batch[r] : 0;
    if (b < 0 || b >= B) continue;
    float xv[VEC], o[VEC];
    if (VEC == 4) {
      float4 t = *reinterpret_cast<const float4*>(x + r * C + c0);
      xv[0] = t.x; xv[1 % VEC] = t.y; xv[2 % VEC] = t.z; xv[3 % VEC] = t.w;
    } else {
      xv[0] = x[r * C + c0];
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      int64_t c = c0 + v;
      float out = xv[v] - (mean_scale ? mean_scale[c] : 1.f) * mean[b * C + c];
      float w = weight ? weight[c] : 1.f;
      float bb = bias ? bias[c] : 0.f;
      float val = fmaf(w * out, rstd[b * C + c], bb);
      o[v] = (val > 0.f) ? val : val * slope;   // slope == 1: identity
    }
    if (VEC == 4)
      *reinterpret_cast<float4*>(y + r * C + c0) =
          make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]);
    else
      y[r * C + c0] = o[0];
  }
}

// [B,C] coefficients so that dx_i = k1 * dy_i - k2 * xhat_i - k3
//   k1 = w*rstd ; k2 = w*rstd*s1/n ; k3 = mean_scale * T / n,
//   T  = rstd*w*(s2 - s1*rstd*(1-mean_scale)*mean)   (= sum_j dout_j)
// and parameter grads: dweight = sum_b s1, dbias = sum_b s2,
//   dmean_scale = -sum_b mean*T.  One thread per column, loops over b.
__global__ void k_graphnorm_bwd_coef(const double* __restrict__ s1,
                                     const double* __restrict__ s2,
                                     const double* __restrict__ count, int64_t B,
                                     int64_t C, const float* __restrict__ weight,
                                     const float* __restrict__ mean_scale,
                                     const float* __restrict__ mean,
                                     const float* __restrict__ rstd,
                                     float* __restrict__ k2, float* __restrict__ k3,
                                     float* __restrict__ dweight,
                                     float* __restrict__ dbias,
                                     float* __restrict__ dmean_scale) {
  int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double w = weight ? (double)weight[c] : 1.0;
  double ms = mean_scale[c];
  double dw = 0, db = 0, dms = 0;
  for (int64_t b = 0; b < B; ++b) {
    double n = fmax(count[b], 1.0);
    double S1 = s1[b * C + c], S2 = s2[b * C + c];
    double rs = rstd[b * C + c], mu = mean[b * C + c];
    double T = rs * w * (S2 - S1 * rs * (1.0 - ms) * mu);
    k2[b * C + c] = (float)(w * rs * S1 / n);
    k3[b * C + c] = (float)(ms * T / n);
    dw += S1;
    db += S2;
    dms -= mu * T;
  }
  if (dweight) dweight[c] = (float)dw;
  if (dbias) dbias[c] = (float)db;
  if (dmean_scale) dmean_scale[c] = (float)dms;
}

template <int VEC>
__global__ void __launch_bounds__(kNormThreads)
k_graphnorm_bwd_apply(const float* __restrict__ x, const float* __restrict__ dy,
                      const int64_t* __restrict__ batch, int64_t N, int64_t C, int64_t B,
                      const float* __restrict__ weight, const float* __restrict__ mean_scale,
                      const float* __restrict__ mean, const float* __restrict__ rstd,
                      const float* __restrict__ k2, const float* __restrict__ k3,
                      const float* __restrict__ yact, float slope, float* __restrict__ dx) {
  int64_t cv = C / VEC;
  int64_t total = N * cv;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    int64_t r = i / cv;
    int64_t c0 = (i - r * cv) * VEC;
    int64_t b = batch ? batch[r] : 0;
    float o[VEC];
    if (b < 0 || b >= B) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) o[v] = 0.f;
    } else {
      float xv[VEC], gv[VEC];
      if (VEC == 4) {
        float4 t = *reinterpret_cast<const float4*>(x + r * C + c0);
        float4 g = *reinterpret_cast<const float4*>(dy + r * C + c0);
        xv[0] = t.x; xv[1 % VEC] = t.y; xv[2 % VEC] = t.z; xv[3 % VEC] = t.w;
        gv[0] = g.x; gv[1 % VEC] = g.y; gv[2 % VEC] = g.z; gv[3 % VEC] = g.w;
        if (yact) {
          float4 yy = *reinterpret_cast<const float4*>(yact + r * C + c0);
          gv[0] *= (yy.x > 0.f) ? 1.f : slope; gv[1 % VEC] *= (yy.y > 0.f) ? 1.f : slope;
          gv[2 % VEC] *= (yy.z > 0.f) ? 1.f : slope; gv[3 % VEC] *= (yy.w > 0.f) ? 1.f : slope;
        }
      } else {
        xv[0] = x[r * C + c0];
        gv[0] = dy[r * C + c0];
        if (yact) gv[0] *= (yact[r * C + c0] > 0.f) ? 1.f : slope;
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        int64_t c = c0 + v;
        float rs = rstd[b * C + c];
        float xhat = (xv[v] - (mean_scale ? mean_scale[c] : 1.f) * mean[b * C + c]) * rs;
        float w = weight ? weight[c] : 1.f;
        o[v] = w * rs * gv[v] - k2[b * C + c] * xhat - k3[b * C + c];
      }
    }
    if (VEC == 4)
      *reinterpret_cast<float4*>(dx + r * C + c0) =
          make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]);
    else
      dx[r * C + c0] = o[0];
  }
}


// ---------------------------------------------------------------------------------------
// Fused finalise + apply (round 2).  The element-wise passes above spend ~25 loads and a 64-bit
// division per 16 bytes of data (per-element parameter lookups), and each norm paid two
// single-CTA kernels (finalise / coefficients) of ~4 us.  Here every CTA first builds the
// [B, C] coefficient table in shared memory from the fp64 sums (a few hundred values), then
// threads with FIXED columns stream their slab of rows: one FMA chain per element, 4 rows in
// flight per thread, no divisions.  Used when B*C <= kNormTableMax and C % 4 == 0.
// ---------------------------------------------------------------------------------------
constexpr int kNormTableMax = 2048;

// y = (x - M) * A + Bb  (then the fused LeakyReLU);  M = mean_scale*mean, A = weight*rstd
__global__ void __launch_bounds__(kNormThreads)
k_graphnorm_apply_fused(const float* __restrict__ x, const int64_t* __restrict__ batch,
                        int64_t N, int C, int B, const float* __restrict__ weight,
                        const float* __restrict__ bias, const float* __restrict__ mean_scale,
                        const double* __restrict__ s1, const double* __restrict__ s2,
                        const double* __restrict__ count, const float* __restrict__ pivot,
                        float eps, float slope, float* __restrict__ y,
                        float* __restrict__ mean_out, float* __restrict__ rstd_out, int tx,
                        int ty, int slab_rows) {
  extern __shared__ __align__(16) float norm_tab[];
  float* tM = norm_tab;
  float* tA = norm_tab + B * C;
  float* tB = norm_tab + 2 * B * C;
  for (int i = threadIdx.x; i < B * C; i += kNormThreads) {
    const int b = i / C, c = i - b * C;
    const double n = fmax(count[b], 1.0);
    const double p = pivot[c], m1 = s1[i] / n, m2 = s2[i] / n;
    const double mu = p + m1;
    const double ms = mean_scale[c];
    const double q = p - ms * mu;
    const double var = fmax(m2 + 2.0 * q * m1 + q * q, 0.0);
    const float rs = (float)(1.0 / sqrt(var + (double)eps));
    const float muf = (float)mu;
    tM[i] = (float)ms * muf;
    tA[i] = (weight ? weight[c] : 1.f) * rs;
    tB[i] = bias ? bias[c] : 0.f;
    if (blockIdx.x == 0) { mean_out[i] = muf; rstd_out[i] = rs; }
  }
  __syncthreads();
  const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
  const int64_t nslabs = (N + slab_rows - 1) / slab_rows;
  constexpr int U = 4;
  for (int ct = 0; ct < C; ct += tx * 4) {
    const int c0 = ct + cx * 4;
    if (c0 >= C) continue;
    float4 M0, A0, B0;   // B == 1: the coefficients live in registers
    if (B == 1) {
      M0 = *reinterpret_cast<const float4*>(tM + c0);
      A0 = *reinterpret_cast<const float4*>(tA + c0);
      B0 = *reinterpret_cast<const float4*>(tB + c0);
    }
    for (int64_t slab = blockIdx.x; slab < nslabs; slab += gridDim.x) {
      const int64_t r0 = slab * slab_rows, r1 = min(r0 + (int64_t)slab_rows, N);
      for (int64_t rb = r0 + ry; rb < r1; rb += (int64_t)ty * U) {
        float4 xq[U];
        int bq[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t r = rb + (int64_t)u * ty;
          bq[u] = -1;
          if (r < r1) {
            bq[u] = batch ? (int)batch[r] : 0;
            xq[u] = *reinterpret_cast<const float4*>(x + r * C + c0);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int b = bq[u];
          if (b < 0 || b >= B) continue;
          float4 M = M0, A = A0, Bb = B0;
          if (B != 1) {
            M = *reinterpret_cast<const float4*>(tM + b * C + c0);
            A = *reinterpret_cast<const float4*>(tA + b * C + c0);
            Bb = *reinterpret_cast<const float4*>(tB + b * C + c0);
          }
          float4 o;
          o.x = fmaf(xq[u].x - M.x, A.x, Bb.x); o.y = fmaf(xq[u].y - M.y, A.y, Bb.y);
          o.z = fmaf(xq[u].z - M.z, A.z, Bb.z); o.w = fmaf(xq[u].w - M.w, A.w, Bb.w);
          o.x = o.x > 0.f ? o.x : o.x * slope; o.y = o.y > 0.f ? o.y : o.y * slope;
          o.z = o.z > 0.f ? o.z : o.z * slope; o.w = o.w > 0.f ? o.w : o.w * slope;
          *reinterpret_cast<float4*>(y + (rb + (int64_t)u * ty) * C + c0) = o;
        }
      }
    }
  }
}

// dx = k1 * g - k2 * xhat - k3,  xhat = (x - M) * rs,  g = dy * act'(y)   (see k_graphnorm_bwd_coef)
__global__ void __launch_bounds__(kNormThreads)
k_graphnorm_bwd_apply_fused(const float* __restrict__ x, const float* __restrict__ dy,
                            const int64_t* __restrict__ batch, int64_t N, int C, int B,
                            const float* __restrict__ weight,
                            const float* __restrict__ mean_scale, const float* __restrict__ mean,
                            const float* __restrict__ rstd, const double* __restrict__ s1,
                            const double* __restrict__ s2, const double* __restrict__ count,
                            const float* __restrict__ yact, float slope, float* __restrict__ dx,
                            float* __restrict__ dweight, float* __restrict__ dbias,
                            float* __restrict__ dmean_scale, int tx, int ty, int slab_rows) {
  extern __shared__ __align__(16) float norm_tab[];
  float* tM = norm_tab;                 // mean_scale * mean
  float* tR = norm_tab + B * C;         // rstd
  float* t1 = norm_tab + 2 * B * C;     // k1
  float* t2 = norm_tab + 3 * B * C;     // k2
  float* t3 = norm_tab + 4 * B * C;     // k3
  for (int i = threadIdx.x; i < B * C; i += kNormThreads) {
    const int b = i / C, c = i - b * C;
    const double w = weight ? (double)weight[c] : 1.0;
    const double ms = mean_scale[c];
    const double n = fmax(count[b], 1.0);
    const double S1 = s1[i], S2 = s2[i], rs = rstd[i], mu = mean[i];
    const double T = rs * w * (S2 - S1 * rs * (1.0 - ms) * mu);
    tM[i] = mean_scale[c] * mean[i];
    tR[i] = rstd[i];
    t1[i] = (float)(w * rs);
    t2[i] = (float)(w * rs * S1 / n);
    t3[i] = (float)(ms * T / n);
  }
  if (blockIdx.x == 0 && (dweight || dbias || dmean_scale)) {
    // parameter gradients: one thread per column, sums over the graphs of the batch
    for (int c = threadIdx.x; c < C; c += kNormThreads) {
      const double w = weight ? (double)weight[c] : 1.0;
      const double ms = mean_scale[c];
      double dw = 0, db = 0, dms = 0;
      for (int b = 0; b < B; ++b) {
        const double S1 = s1[b * C + c], S2 = s2[b * C + c];
        const double rs = rstd[b * C + c], mu = mean[b * C + c];
        const double T = rs * w * (S2 - S1 * rs * (1.0 - ms) * mu);
        dw += S1;
        db += S2;
        dms -= mu * T;
      }
      if (dweight) dweight[c] = (float)dw;
      if (dbias) dbias[c] = (float)db;
      if (dmean_scale) dmean_scale[c] = (float)dms;
    }
  }
  __syncthreads();
  const int cx = threadIdx.x % tx, ry = threadIdx.x / tx;
  const int64_t nslabs = (N + slab_rows - 1) / slab_rows;
  constexpr int U = 4;
  for (int ct = 0; ct < C; ct += tx * 4) {
    const int c0 = ct + cx * 4;
    if (c0 >= C) continue;
    float4 M0, R0, K1, K2, K3;
    if (B == 1) {
      M0 = *reinterpret_cast<const float4*>(tM + c0);
      R0 = *reinterpret_cast<const float4*>(tR + c0);
      K1 = *reinterpret_cast<const float4*>(t1 + c0);
      K2 = *reinterpret_cast<const float4*>(t2 + c0);
      K3 = *reinterpret_cast<const float4*>(t3 + c0);
    }
    for (int64_t slab = blockIdx.x; slab < nslabs; slab += gridDim.x) {
      const int64_t r0 = slab * slab_rows, r1 = min(r0 + (int64_t)slab_rows, N);
      for (int64_t rb = r0 + ry; rb < r1; rb += (int64_t)ty * U) {
        float4 xq[U], gq[U];
        int bq[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t r = rb + (int64_t)u * ty;
          bq[u] = -1;
          if (r < r1) {
            bq[u] = batch ? (int)batch[r] : 0;
            xq[u] = *reinterpret_cast<const float4*>(x + r * C + c0);
            gq[u] = *reinterpret_cast<const float4*>(dy + r * C + c0);
            if (yact) {
              const float4 yy = *reinterpret_cast<const float4*>(yact + r * C + c0);
              gq[u].x *= (yy.x > 0.f) ? 1.f : slope; gq[u].y *= (yy.y > 0.f) ? 1.f : slope;
              gq[u].z *= (yy.z > 0.f) ? 1.f : slope; gq[u].w *= (yy.w > 0.f) ? 1.f : slope;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int b = bq[u];
          if (rb + (int64_t)u * ty >= r1) continue;   // past the slab
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (b >= 0 && b < B) {
            float4 M = M0, R = R0, k1 = K1, k2 = K2, k3 = K3;
            if (B != 1) {
              M = *reinterpret_cast<const float4*>(tM + b * C + c0);
              R = *reinterpret_cast<const float4*>(tR + b * C + c0);
              k1 = *reinterpret_cast<const float4*>(t1 + b * C + c0);
              k2 = *reinterpret_cast<const float4*>(t2 + b * C + c0);
              k3 = *reinterpret_cast<const float4*>(t3 + b * C + c0);
            }
            o.x = k1.x * gq[u].x - k2.x * ((xq[u].x - M.x) * R.x) - k3.x;
            o.y = k1.y * gq[u].y - k2.y * ((xq[u].y - M.y) * R.y) - k3.y;
            o.z = k1.z * gq[u].z - k2.z * ((xq[u].z - M.z) * R.z) - k3.z;
            o.w = k1.w * gq[u].w - k2.w * ((xq[u].w - M.w) * R.w) - k3.w;
          }
          *reinterpret_cast<float4*>(dx + (rb + (int64_t)u * ty) * C + c0) = o;
        }
      }
    }
  }
}


// ---------------------------------------------------------------------------------------
// Graph-wise GroupNorm / LayerNorm(mode='graph'): statistics per (graph b, channel group g)
// over nodes x group channels (reference src/nn/norm.py:181-218; PyG LayerNorm 'graph' is
// the num_groups = 1 case).  The N x C passes are the GraphNorm kernels above run with
// mean_scale == 1; these [B, G]-sized kernels turn their per-channel sums into group values.
// ---------------------------------------------------------------------------------------
// sum_x[b, c] <- n_b * mean_{b, g(c)}  (so that the MODE 1 pass centres on the group mean)
__global__ void k_groupnorm_mean(double* __restrict__ sum_x, const double* __restrict__ count,
                                 int64_t B, int64_t C, int64_t G) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * G) return;
  const int64_t b = t / G, g = t - b * G, gc = C / G;
  const double n = fmax(count[b], 1.0);
  double s = 0;
  for (int64_t j = 0; j < gc; ++j) s += sum_x[b * C + g * gc + j];
  const double m = s / (n * (double)gc);
  for (int64_t j = 0; j < gc; ++j) sum_x[b * C + g * gc + j] = m * n;
}

__global__ void k_groupnorm_finalize(const double* __restrict__ sum_x /* n * mean */,
                                     const double* __restrict__ sum_sq,
                                     const double* __restrict__ count, int64_t B, int64_t C,
                                     int64_t G, float eps, int eps_outside,
                                     float* __restrict__ mean, float* __restrict__ rstd) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * G) return;
  const int64_t b = t / G, g = t - b * G, gc = C / G;
  const double n = fmax(count[b], 1.0);
  double s = 0;
  for (int64_t j = 0; j < gc; ++j) s += sum_sq[b * C + g * gc + j];
  const float var = (float)(s / (n * (double)gc));
  // PyG LayerNorm without `batch`: x / (std + eps); everything else: x / sqrt(var + eps)
  const float rs = eps_outside ? 1.f / (sqrtf(var) + eps) : 1.f / sqrtf(var + eps);
  for (int64_t j = 0; j < gc; ++j) {
    mean[b * C + g * gc + j] = (float)(sum_x[b * C + g * gc + j] / n);
    rstd[b * C + g * gc + j] = rs;
  }
}

// dx_i = w rs dy_i - k2 xhat_i - k3 with group-wide k2 = f * mean_g(w dy xhat),
// k3 = rs * mean_g(w dy); f = rs (eps inside the sqrt) or 1/std (eps outside)
__global__ void k_groupnorm_bwd_coef(const double* __restrict__ s1, const double* __restrict__ s2,
                                     const double* __restrict__ count, int64_t B, int64_t C,
                                     int64_t G, const float* __restrict__ weight,
                                     const float* __restrict__ rstd, float eps, int eps_outside,
                                     float* __restrict__ k2, float* __restrict__ k3,
                                     float* __restrict__ dweight, float* __restrict__ dbias) {
  int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < B * G) {
    const int64_t b = t / G, g = t - b * G, gc = C / G;
    const double n = fmax(count[b], 1.0) * (double)gc;
    double a1 = 0, a2 = 0;
    for (int64_t j = 0; j < gc; ++j) {
      const int64_t c = g * gc + j;
      const double w = weight ? (double)weight[c] : 1.0;
      a1 += w * s1[b * C + c];
      a2 += w * s2[b * C + c];
    }
    const double rs = rstd[b * C + g * gc];
    const double f = eps_outside ? 1.0 / fmax(1.0 / rs - (double)eps, 1e-30) : rs;
    for (int64_t j = 0; j < gc; ++j) {
      k2[b * C + g * gc + j] = (float)(f * a1 / n);
      k3[b * C + g * gc + j] = (float)(rs * a2 / n);
    }
  }
  if (t < C) {
    double dw = 0, db = 0;
    for (int64_t b = 0; b < B; ++b) {
      dw += s1[b * C + t];
      db += s2[b * C + t];
    }
    if (dweight) dweight[t] = (float)dw;
    if (dbias) dbias[t] = (float)db;
  }
}

// Statistics passes: ONE contiguous slab of rows per CTA, ~4 CTAs per SM.  Contiguous matters
// for batches: a CTA that strides over the whole array (the round-1 persistent loop over 128-row
// slabs) touches every graph of the batch and falls off the "whole CTA in one segment" fast
// path — GraphNorm on an 8-scene NAGBatch was 2.2x slower than with contiguous slabs (cfg 4).
// The CTA count stays at ~4 per SM so that the fp64 atomics on the [B, C] accumulators do not
// grow (8 smaller CTAs per SM measured 3 % slower on the single-scene cfg 2).
static inline int norm_slab_rows(int64_t N, int ty) {
  const int64_t per_iter = (int64_t)ty * 4;
  const int64_t ctas = (int64_t)device_sm_count() * 4;
  int64_t want = (N + ctas - 1) / ctas;
  want = (want + per_iter - 1) / per_iter * per_iter;
  if (want < kNormRows) want = kNormRows;
  if (want > (1 << 22)) want = (1 << 22);
  return (int)want;
}

// element-wise passes: contiguous slabs too (coalesced streams per CTA), ~8 CTAs per SM
static inline int norm_slab_rows_apply(int64_t N, int ty) {
  const int64_t per_iter = (int64_t)ty * 4;
  const int64_t ctas = (int64_t)device_sm_count() * 8;
  int64_t want = (N + ctas - 1) / ctas;
  want = (want + per_iter - 1) / per_iter * per_iter;
  if (want < per_iter) want = per_iter;
  if (want > (1 << 22)) want = (1 << 22);
  return (int)want;
}

static inline ColMap col_map(int64_t C, int vec) {
  int64_t cols = C / vec;
  int tx = 1;
  while (tx < cols && tx < kNormThreads) tx <<= 1;  // pow2 >= cols (cap 256)
  if (tx > kNormThreads) tx = kNormThreads;
  ColMap m;
  m.tx = tx;
  m.ty = kNormThreads / tx;
  return m;
}

}  // namespace spt

using namespace spt;

namespace {
inline int64_t imin(int64_t a, int64_t b) { return a < b ? a : b; }

struct NormWs {
  double* acc0;
  double* acc1;
  double* count;
  float* k2;
  float* k3;
  size_t zero_bytes;
};
inline NormWs carve(void* ws, int64_t B, int64_t C) {
  size_t bc = (size_t)B * (size_t)C;
  char* p = (char*)ws;
  NormWs w;
  w.acc0 = (double*)p;
  w.acc1 = (double*)(p + align_up(bc * 8, 256));
  w.count = (double*)(p + 2 * align_up(bc * 8, 256));
  w.k2 = (float*)(p + 2 * align_up(bc * 8, 256) + align_up((size_t)B * 8, 256));
  w.k3 = (float*)((char*)w.k2 + align_up(bc * 4, 256));
  w.zero_bytes = 2 * align_up(bc * 8, 256) + align_up((size_t)B * 8, 256);
  return w;
}
}  // namespace

extern "C" {

// ws: acc0[B*C] f64 | acc1[B*C] f64 | count[B] f64 | k2[B*C] f32 | k3[B*C] f32
size_t spt_graphnorm_workspace_bytes(int64_t B, int64_t C) {
  if (B < 0 || C < 0) return 0;
  size_t bc = (size_t)B * (size_t)C;
  return align_up(bc * 8, 256) * 2 + align_up((size_t)B * 8, 256) +
         align_up(bc * 4, 256) * 2;
}

int spt_graphnorm_fwd(const float* x, const int64_t* batch, int64_t N, int64_t C,
                      int64_t B, const float* weight, const float* bias,
                      const float* mean_scale, float eps, float act_slope, float* y,
                      float* mean, float* rstd, void* ws, size_t ws_bytes, void* stream_) {
  SPT_REQUIRE(N >= 0 && C > 0 && B > 0, SPT_E_INVALID, "graphnorm_fwd: bad sizes");
  SPT_REQUIRE(mean_scale && mean && rstd && ws && (N == 0 || (x && y)), SPT_E_INVALID,
              "graphnorm_fwd: null pointer");
  SPT_REQUIRE(ws_bytes >= spt_graphnorm_workspace_bytes(B, C), SPT_E_WORKSPACE,
              "graphnorm_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream_;
  // one graph: every row is in segment 0, the per-row id loads (8 bytes per 16 bytes of data
  // per thread) are skipped
  if (B == 1) batch = nullptr;
  NormWs w = carve(ws, B, C);
  cudaError_t ce = cudaMemsetAsync(ws, 0, w.zero_bytes, st);
  if (ce != cudaSuccess) {
    set_error("graphnorm_fwd memset: %s", cudaGetErrorString(ce));
    return (int)ce;
  }
  int vec = (C % 4 == 0) ? 4 : 1;
  ColMap cm = col_map(C, vec);
  const int slab_rows = norm_slab_rows(N, cm.ty);
  unsigned slabs = (unsigned)imin(ceil_div(N > 0 ? N : 1, slab_rows), device_sm_count() * 4);
  int64_t total = N * (C / vec);
  int agrid = (int)imin(ceil_div(total > 0 ? total : 1, kNormThreads), device_sm_count() * 16);
  // single pass over x: shifted first and second moments around a per-channel pivot
  float* pivot = w.k2;  // [C] floats of the (forward-unused) k2 area
  if (N > 0) {
    if (vec == 4 && !batch && !getenv("SPT_NORM_NO_FUSED")) {
      k_graphnorm_stats_single<3><<<slabs, kNormThreads, 0, st>>>(
          x, nullptr, nullptr, 1.f, N, (int)C, nullptr, nullptr, nullptr, w.acc0, w.acc1, w.count,
          cm.tx, cm.ty, slab_rows, pivot, (int)imin(N, 64));
    } else if (vec == 4 && !getenv("SPT_NORM_NO_FUSED")) {
      // pivot computed inside the statistics kernel (one launch less)
      k_graphnorm_stats<3, 4><<<slabs, kNormThreads, 0, st>>>(
          x, nullptr, nullptr, 1.f, batch, N, C, B, nullptr, nullptr, nullptr, nullptr, nullptr,
          w.acc0, w.acc1, w.count, cm.tx, cm.ty, slab_rows, pivot, (int)imin(N, 64));
    } else {
    k_graphnorm_pivot<<<(unsigned)ceil_div(C, 128), 128, 0, st>>>(x, imin(N, 64), C, pivot);
    if (vec == 4)
      k_graphnorm_stats<3, 4><<<slabs, kNormThreads, 0, st>>>(
          x, nullptr, nullptr, 1.f, batch, N, C, B, pivot, nullptr, nullptr, nullptr, nullptr,
          w.acc0, w.acc1, w.count, cm.tx, cm.ty, slab_rows);
    else
      k_graphnorm_stats<3, 1><<<slabs, kNormThreads, 0, st>>>(
          x, nullptr, nullptr, 1.f, batch, N, C, B, pivot, nullptr, nullptr, nullptr, nullptr,
          w.acc0, w.acc1, w.count, cm.tx, cm.ty, slab_rows);
    }
  } else {
    cudaMemsetAsync(pivot, 0, (size_t)C * 4, st);
  }
  if (N > 0 && vec == 4 && B * C <= kNormTableMax && !getenv("SPT_NORM_NO_FUSED")) {
    // finalise + apply in one launch (coefficient table in shared memory)
    const int slab_a = norm_slab_rows_apply(N, cm.ty);
    const unsigned ga = (unsigned)imin(ceil_div(N, slab_a), device_sm_count() * 8);
    k_graphnorm_apply_fused<<<ga, kNormThreads, (size_t)3 * B * C * 4, st>>>(
        x, batch, N, (int)C, (int)B, weight, bias, mean_scale, w.acc0, w.acc1, w.count, pivot,
        eps, act_slope, y, mean, rstd, cm.tx, cm.ty, slab_a);
    return check_launch("graphnorm_fwd");
  }
  k_graphnorm_finalize_shifted<<<(unsigned)ceil_div(B * C, 256), 256, 0, st>>>(
      w.acc0, w.acc1, w.count, pivot, mean_scale, B, C, eps, mean, rstd);
  if (N > 0) {
    if (vec == 4)
      k_graphnorm_apply<4><<<agrid, kNormThreads, 0, st>>>(x, batch, N, C, B, weight, bias,
                                                          mean_scale, mean, rstd, act_slope, y);
    else
      k_graphnorm_apply<1><<<agrid, kNormThreads, 0, st>>>(x, batch, N, C, B, weight, bias,
                                                          mean_scale, mean, rstd, act_slope, y);
  }
  return check_launch("graphnorm_fwd");
}

int spt_graphnorm_bwd(const float* x, const float* dy, const int64_t* batch, int64_t N,
                      int64_t C, int64_t B, const float* weight, const float* mean_scale,
                      const float* mean, const float* rstd, const float* y_act,
                      float act_slope, float* dx, float* dweight, float* dbias,
                      float* dmean_scale, void* ws, size_t ws_bytes, void* stream_) {
  const float* yact = (act_slope != 1.f) ? y_act : nullptr;
  SPT_REQUIRE(act_slope == 1.f || y_act, SPT_E_INVALID,
              "graphnorm_bwd: fused activation needs the forward output");
  SPT_REQUIRE(N >= 0 && C > 0 && B > 0, SPT_E_INVALID, "graphnorm_bwd: bad sizes");
  SPT_REQUIRE(mean_scale && mean && rstd && ws && (N == 0 || (x && dy && dx)),
              SPT_E_INVALID, "graphnorm_bwd: null pointer");
  SPT_REQUIRE(ws_bytes >= spt_graphnorm_workspace_bytes(B, C), SPT_E_WORKSPACE,
              "graphnorm_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream_;
  if (B == 1) batch = nullptr;
  NormWs w = carve(ws, B, C);
  cudaError_t ce = cudaMemsetAsync(ws, 0, w.zero_bytes, st);
  if (ce != cudaSuccess) {
    set_error("graphnorm_bwd memset: %s", cudaGetErrorString(ce));
    return (int)ce;
  }
  int vec = (C % 4 == 0) ? 4 : 1;
  ColMap cm = col_map(C, vec);
  const int slab_rows = norm_slab_rows(N, cm.ty);
  unsigned slabs = (unsigned)imin(ceil_div(N > 0 ? N : 1, slab_rows), device_sm_count() * 4);
  int64_t total = N * (C / vec);
  int agrid = (int)imin(ceil_div(total > 0 ? total : 1, kNormThreads), device_sm_count() * 16);
  if (N > 0) {
    if (vec == 4 && !batch && !getenv("SPT_NORM_NO_FUSED"))
      k_graphnorm_stats_single<2><<<slabs, kNormThreads, 0, st>>>(
          x, dy, yact, act_slope, N, (int)C, mean_scale, mean, rstd, w.acc0, w.acc1, w.count, cm.tx,
          cm.ty, slab_rows, nullptr, 0);
    else if (vec == 4)
      k_graphnorm_stats<2, 4><<<slabs, kNormThreads, 0, st>>>(
          x, dy, yact, act_slope, batch, N, C, B, mean_scale, nullptr, nullptr, mean, rstd, w.acc0,
          w.acc1, w.count, cm.tx, cm.ty, slab_rows);   // also counts the rows per graph
    else
      k_graphnorm_stats<2, 1><<<slabs, kNormThreads, 0, st>>>(
          x, dy, yact, act_slope, batch, N, C, B, mean_scale, nullptr, nullptr, mean, rstd, w.acc0,
          w.acc1, w.count, cm.tx, cm.ty, slab_rows);
  }
  if (N > 0 && vec == 4 && B * C <= kNormTableMax && !getenv("SPT_NORM_NO_FUSED")) {
    // coefficients + parameter gradients + apply in one launch
    const int slab_a = norm_slab_rows_apply(N, cm.ty);
    const unsigned ga = (unsigned)imin(ceil_div(N, slab_a), device_sm_count() * 8);
    k_graphnorm_bwd_apply_fused<<<ga, kNormThreads, (size_t)5 * B * C * 4, st>>>(
        x, dy, batch, N, (int)C, (int)B, weight, mean_scale, mean, rstd, w.acc0, w.acc1, w.count,
        yact, act_slope, dx, dweight, dbias, dmean_scale, cm.tx, cm.ty, slab_a);
    return check_launch("graphnorm_bwd");
  }
  k_graphnorm_bwd_coef<<<(unsigned)ceil_div(C, 128), 128, 0, st>>>(
      w.acc0, w.acc1, w.count, B, C, weight, mean_scale, mean, rstd, w.k2, w.k3, dweight,
      dbias, dmean_scale);
  if (N > 0) {
    if (vec == 4)
      k_graphnorm_bwd_apply<4><<<agrid, kNormThreads, 0, st>>>(
          x, dy, batch, N, C, B, weight, mean_scale, mean, rstd, w.k2, w.k3, yact, act_slope,
          dx);
    else
      k_graphnorm_bwd_apply<1><<<agrid, kNormThreads, 0, st>>>(
          x, dy, batch, N, C, B, weight, mean_scale, mean, rstd, w.k2, w.k3, yact, act_slope,
          dx);
  }
  return check_launch("graphnorm_bwd");
}

int spt_groupnorm_fwd(const float* x, const int64_t* batch, int64_t N, int64_t C, int64_t B,
                      int64_t num_groups, const float* weight, const float* bias, float eps,
                      int eps_outside, float* y, float* mean, float* rstd, void* ws,
                      size_t ws_bytes, void* stream_) {
  SPT_REQUIRE(N >= 0 && C > 0 && B > 0 && num_groups > 0 && C % num_groups == 0, SPT_E_INVALID,
              "groupnorm_fwd: bad sizes");
  SPT_REQUIRE(mean && rstd && ws && (N == 0 || (x && y)), SPT_E_INVALID,
              "groupnorm_fwd: null pointer");
  SPT_REQUIRE(ws_bytes >= spt_graphnorm_workspace_bytes(B, C), SPT_E_WORKSPACE,
              "groupnorm_fwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream_;
  NormWs w = carve(ws, B, C);
  cudaError_t ce = cudaMemsetAsync(ws, 0, w.zero_bytes, st);
  if (ce != cudaSuccess) {
    set_error("groupnorm_fwd memset: %s", cudaGetErrorString(ce));
    return (int)ce;
  }
  const int64_t G = num_groups;
  int vec = (C % 4 == 0) ? 4 : 1;
  ColMap cm = col_map(C, vec);
  const int slab_rows = norm_slab_rows(N, cm.ty);
  unsigned slabs = (unsigned)imin(ceil_div(N > 0 ? N : 1, slab_rows), device_sm_count() * 4);
  int64_t total = N * (C / vec);
  int agrid = (int)imin(ceil_div(total > 0 ? total : 1, kNormThreads), device_sm_count() * 16);
  const unsigned ggrid = (unsigned)ceil_div(B * G, 128);
  if (N > 0) {
    if (vec == 4)
      k_graphnorm_stats<0, 4><<<slabs, kNormThreads, 0, st>>>(
          x, nullptr, nullptr, 1.f, batch, N, C, B, nullptr, nullptr, nullptr, nullptr, nullptr,
          w.acc0, nullptr, w.count, cm.tx, cm.ty, slab_rows);
    else
      k_graphnorm_stats<0, 1><<<slabs, kNormThreads, 0, st>>>(
          x, nullptr, nullptr, 1.f, batch, N, C, B, nullptr, nullptr, nullptr, nullptr, nullptr,
          w.acc0, nullptr, w.count, cm.tx, cm.ty, slab_rows);
  }
  k_groupnorm_mean<<<ggrid, 128, 0, st>>>(w.acc0, w.count, B, C, G);
  if (N > 0) {
    if (vec == 4)
      k_graphnorm_stats<1, 4><<<slabs, kNormThreads, 0, st>>>(
          x, nullptr, nullptr, 1.f, batch, N, C, B, nullptr, w.acc0, w.count, nullptr, nullptr,
          w.acc1, nullptr, nullptr, cm.tx, cm.ty, slab_rows);
    else
      k_graphnorm_stats<1, 1><<<slabs, kNormThreads, 0, st>>>(
          x, nullptr, nullptr, 1.f, batch, N, C, B, nullptr, w.acc0, w.count, nullptr, nullptr,
          w.acc1, nullptr, nullptr, cm.tx, cm.ty, slab_rows);
  }
  k_groupnorm_finalize<<<ggrid, 128, 0, st>>>(w.acc0, w.acc1, w.count, B, C, G, eps, eps_outside,
                                              mean, rstd);
  if (N > 0) {
    if (vec == 4)
      k_graphnorm_apply<4><<<agrid, kNormThreads, 0, st>>>(x, batch, N, C, B, weight, bias,
                                                          nullptr, mean, rstd, 1.f, y);
    else
      k_graphnorm_apply<1><<<agrid, kNormThreads, 0, st>>>(x, batch, N, C, B, weight, bias,
                                                          nullptr, mean, rstd, 1.f, y);
  }
  return check_launch("groupnorm_fwd");
}

int spt_groupnorm_bwd(const float* x, const float* dy, const int64_t* batch, int64_t N,
                      int64_t C, int64_t B, int64_t num_groups, const float* weight,
                      const float* mean, const float* rstd, float eps, int eps_outside,
                      float* dx, float* dweight, float* dbias, void* ws, size_t ws_bytes,
                      void* stream_) {
  SPT_REQUIRE(N >= 0 && C > 0 && B > 0 && num_groups > 0 && C % num_groups == 0, SPT_E_INVALID,
              "groupnorm_bwd: bad sizes");
  SPT_REQUIRE(mean && rstd && ws && (N == 0 || (x && dy && dx)), SPT_E_INVALID,
              "groupnorm_bwd: null pointer");
  SPT_REQUIRE(ws_bytes >= spt_graphnorm_workspace_bytes(B, C), SPT_E_WORKSPACE,
              "groupnorm_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream_;
  NormWs w = carve(ws, B, C);
  cudaError_t ce = cudaMemsetAsync(ws, 0, w.zero_bytes, st);
  if (ce != cudaSuccess) {
    set_error("groupnorm_bwd memset: %s", cudaGetErrorString(ce));
    return (int)ce;
  }
  const int64_t G = num_groups;
  int vec = (C % 4 == 0) ? 4 : 1;
  ColMap cm = col_map(C, vec);
  const int slab_rows = norm_slab_rows(N, cm.ty);
  unsigned slabs = (unsigned)imin(ceil_div(N > 0 ? N : 1, slab_rows), device_sm_count() * 4);
  int64_t total = N * (C / vec);
  int agrid = (int)imin(ceil_div(total > 0 ? total : 1, kNormThreads), device_sm_count() * 16);
  if (N > 0) {
    if (vec == 4)
      k_graphnorm_stats<2, 4><<<slabs, kNormThreads, 0, st>>>(
          x, dy, nullptr, 1.f, batch, N, C, B, nullptr, nullptr, nullptr, mean, rstd, w.acc0,
          w.acc1, w.count, cm.tx, cm.ty, slab_rows);
    else
      k_graphnorm_stats<2, 1><<<slabs, kNormThreads, 0, st>>>(
          x, dy, nullptr, 1.f, batch, N, C, B, nullptr, nullptr, nullptr, mean, rstd, w.acc0,
          w.acc1, w.count, cm.tx, cm.ty, slab_rows);
  }
  const int64_t work = B * G > C ? B * G : C;
  k_groupnorm_bwd_coef<<<(unsigned)ceil_div(work, 128), 128, 0, st>>>(
      w.acc0, w.acc1, w.count, B, C, G, weight, rstd, eps, eps_outside, w.k2, w.k3, dweight,
      dbias);
  if (N > 0) {
    if (vec == 4)
      k_graphnorm_bwd_apply<4><<<agrid, kNormThreads, 0, st>>>(
          x, dy, batch, N, C, B, weight, nullptr, mean, rstd, w.k2, w.k3, nullptr, 1.f, dx);
    else
      k_graphnorm_bwd_apply<1><<<agrid, kNormThreads, 0, st>>>(
          x, dy, batch, N, C, B, weight, nullptr, mean, rstd, w.k2, w.k3, nullptr, 1.f, dx);
  }
  return check_launch("groupnorm_bwd");
}

}  // extern "C"
