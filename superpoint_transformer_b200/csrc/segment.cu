// segment.cu — HBM-bound segment kernels: row gathers, CSR segment pooling
// (sum/mean/max/min, fwd + gather-form bwd), UnitSphereNorm.
//
// Layout: features are fp32 row-major [rows, C].  A warp owns one output row and
// moves it as 16-byte vectors (C % 4 == 0 fast path; scalar path otherwise), so
// every gathered child row is one or a few full 128-byte lines.
#include "common.cuh"

namespace spt {

// ------------------------------------------------------------------ row gather
template <typename IdxT, int VEC>
__global__ void k_gather_rows(const float* __restrict__ x, const IdxT* __restrict__ idx,
                              int64_t n_out, int64_t C, float* __restrict__ out) {
  // one warp per output row, rows grid-strided
  int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_out; r += nwarps) {
    int64_t s = (int64_t)idx[r];
    const float* src = x + s * C;
    float* dst = out + r * C;
    if (VEC == 4) {
      for (int64_t c = lane * 4; c < C; c += 128)
        *reinterpret_cast<float4*>(dst + c) = *reinterpret_cast<const float4*>(src + c);
    } else {
      for (int64_t c = lane; c < C; c += 32) dst[c] = src[c];
    }
  }
}

// ------------------------------------------------------------------ pool fwd
template <int REDUCE, int VEC>
__global__ void k_segment_pool_fwd(const float* __restrict__ x,
                                   const int32_t* __restrict__ ptr,
                                   const int32_t* __restrict__ points,
                                   int64_t num_parents, int64_t C,
                                   float* __restrict__ out, int32_t* __restrict__ arg) {
  int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < num_parents; p += nwarps) {
    int b = ptr[p], e = ptr[p + 1];
    float inv = 1.f;
    if (REDUCE == SPT_REDUCE_MEAN) inv = 1.f / (float)max(e - b, 1);
    for (int64_t c0 = (int64_t)lane * VEC; c0 < C; c0 += 32 * VEC) {
      float acc[VEC];
      int32_t am[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        acc[v] = (REDUCE == SPT_REDUCE_MAX) ? -FLT_MAX
                 : (REDUCE == SPT_REDUCE_MIN) ? FLT_MAX : 0.f;
        am[v] = -1;
      }
      // two children in flight to overlap the dependent index->row loads
      int i = b;
      for (; i + 1 < e; i += 2) {
        int ch0 = points ? points[i] : i;
        int ch1 = points ? points[i + 1] : i + 1;
        float r0[VEC], r1[VEC];
        if (VEC == 4) {
          float4 t0 = *reinterpret_cast<const float4*>(x + (int64_t)ch0 * C + c0);
          float4 t1 = *reinterpret_cast<const float4*>(x + (int64_t)ch1 * C + c0);
          r0[0] = t0.x; r0[1 % VEC] = t0.y; r0[2 % VEC] = t0.z; r0[3 % VEC] = t0.w;
          r1[0] = t1.x; r1[1 % VEC] = t1.y; r1[2 % VEC] = t1.z; r1[3 % VEC] = t1.w;
        } else {
          r0[0] = x[(int64_t)ch0 * C + c0];
          r1[0] = x[(int64_t)ch1 * C + c0];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          if (REDUCE == SPT_REDUCE_MAX) {
            if (r0[v] > acc[v] || am[v] < 0) { acc[v] = r0[v]; am[v] = ch0; }
            if (r1[v] > acc[v]) { acc[v] = r1[v]; am[v] = ch1; }
          } else if (REDUCE == SPT_REDUCE_MIN) {
            if (r0[v] < acc[v] || am[v] < 0) { acc[v] = r0[v]; am[v] = ch0; }
            if (r1[v] < acc[v]) { acc[v] = r1[v]; am[v] = ch1; }
          } else {
            acc[v] += r0[v];
            acc[v] += r1[v];
          }
        }
      }
      if (i < e) {
        int ch0 = points ? points[i] : i;
        float r0[VEC];
        if (VEC == 4) {
          float4 t0 = *reinterpret_cast<const float4*>(x + (int64_t)ch0 * C + c0);
          r0[0] = t0.x; r0[1 % VEC] = t0.y; r0[2 % VEC] = t0.z; r0[3 % VEC] = t0.w;
        } else {
          r0[0] = x[(int64_t)ch0 * C + c0];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          if (REDUCE == SPT_REDUCE_MAX) {
            if (r0[v] > acc[v] || am[v] < 0) { acc[v] = r0[v]; am[v] = ch0; }
          } else if (REDUCE == SPT_REDUCE_MIN) {
            if (r0[v] < acc[v] || am[v] < 0) { acc[v] = r0[v]; am[v] = ch0; }
          } else {
            acc[v] += r0[v];
          }
        }
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        if (REDUCE == SPT_REDUCE_MAX || REDUCE == SPT_REDUCE_MIN) {
          if (am[v] < 0) acc[v] = 0.f;  // empty segment -> 0
        } else if (REDUCE == SPT_REDUCE_MEAN) {
          acc[v] *= inv;
        }
      }
      if (VEC == 4) {
        *reinterpret_cast<float4*>(out + p * C + c0) =
            make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
        if (arg && (REDUCE == SPT_REDUCE_MAX || REDUCE == SPT_REDUCE_MIN))
          *reinterpret_cast<int4*>(arg + p * C + c0) =
              make_int4(am[0], am[1 % VEC], am[2 % VEC], am[3 % VEC]);
      } else {
        out[p * C + c0] = acc[0];
        if (arg && (REDUCE == SPT_REDUCE_MAX || REDUCE == SPT_REDUCE_MIN))
          arg[p * C + c0] = am[0];
      }
    }
  }
}

// ------------------------------------------------------------------ pool bwd
template <int REDUCE, int VEC>
__global__ void k_segment_pool_bwd(const float* __restrict__ dout,
                                   const int64_t* __restrict__ parent,
                                   const int32_t* __restrict__ ptr,
                                   const int32_t* __restrict__ arg,
                                   int64_t num_children, int64_t C,
                                   float* __restrict__ dx) {
  int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < num_children; i += nwarps) {
    int64_t p = parent[i];
    float scale = 1.f;
    if (REDUCE == SPT_REDUCE_MEAN) scale = 1.f / (float)max(ptr[p + 1] - ptr[p], 1);
    for (int64_t c0 = (int64_t)lane * VEC; c0 < C; c0 += 32 * VEC) {
      if (VEC == 4) {
        float4 g = *reinterpret_cast<const float4*>(dout + p * C + c0);
        if (REDUCE == SPT_REDUCE_MAX || REDUCE == SPT_REDUCE_MIN) {
          int4 a = *reinterpret_cast<const int4*>(arg + p * C + c0);
          g.x = (a.x == (int)i) ? g.x : 0.f;
          g.y = (a.y == (int)i) ? g.y : 0.f;
          g.z = (a.z == (int)i) ? g.z : 0.f;
          g.w = (a.w == (int)i) ? g.w : 0.f;
        } else if (REDUCE == SPT_REDUCE_MEAN) {
          g.x *= scale; g.y *= scale; g.z *= scale; g.w *= scale;
        }
        *reinterpret_cast<float4*>(dx + i * C + c0) = g;
      } else {
        float g = dout[p * C + c0];
        if (REDUCE == SPT_REDUCE_MAX || REDUCE == SPT_REDUCE_MIN)
          g = (arg[p * C + c0] == (int)i) ? g : 0.f;
        else if (REDUCE == SPT_REDUCE_MEAN)
          g *= scale;
        dx[i * C + c0] = g;
      }
    }
  }
}

// ------------------------------------------------------------------ unit sphere
// one warp per parent: bbox, weighted centroid -> center[3], diameter
__global__ void k_unitsphere_stats(const float* __restrict__ pos,
                                   const int32_t* __restrict__ ptr,
                                   const int32_t* __restrict__ points,
                                   const float* __restrict__ w, int64_t num_parents,
                                   float* __restrict__ center /*[Np,3]*/,
                                   float* __restrict__ diameter /*[Np]*/) {
  int lane = threadIdx.x & 31;
  int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (p >= num_parents) return;
  int b = ptr[p], e = ptr[p + 1];
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  float mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  float sw = 0.f, sx[3] = {0.f, 0.f, 0.f};
  for (int i = b + lane; i < e; i += 32) {
    int ch = points ? points[i] : i;
    float wi = w ? w[ch] : 1.f;
    sw += wi;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float v = pos[(int64_t)ch * 3 + d];
      mn[d] = fminf(mn[d], v);
      mx[d] = fmaxf(mx[d], v);
      sx[d] += v * wi;
    }
  }
  sw = warp_sum(sw);
  float diam = 0.f;
  float cen[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    mn[d] = warp_min(mn[d]);
    mx[d] = warp_max(mx[d]);
    sx[d] = warp_sum(sx[d]);
    // empty segment: scatter min/max leave 0 (torch_scatter) -> span 0
    float span = (e > b) ? (mx[d] - mn[d]) : 0.f;
    diam = fmaxf(diam, span);
    // scatter_mean_weighted: w_segment == 0 -> 1 ; unweighted: clamp(count,1)
    float den = (sw == 0.f) ? 1.f : sw;
    cen[d] = sx[d] / den;
  }
  if (lane == 0) {
    center[p * 3 + 0] = cen[0];
    center[p * 3 + 1] = cen[1];
    center[p * 3 + 2] = cen[2];
    diameter[p] = diam;
  }
}

__global__ void k_unitsphere_apply(const float* __restrict__ pos,
                                   const int64_t* __restrict__ parent,
                                   const float* __restrict__ center,
                                   const float* __restrict__ diameter, int64_t N,
                                   float* __restrict__ pos_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int64_t p = parent ? parent[i] : 0;
  float inv = 1.f / (diameter[p] + 1e-2f);
#pragma unroll
  for (int d = 0; d < 3; ++d)
    pos_out[i * 3 + d] = (pos[i * 3 + d] - center[p * 3 + d]) * inv;
}

// ------------------------------------------------------------------ segment mean / std
// torch_scatter semantics: mean = sum / max(count, 1);
// std (unbiased) = sqrt(sum (x - mean)^2 / (max(count - 1, 1) + 1e-6)).  Warp per parent,
// two passes over the parent's children (the second one hits L1/L2), lane = channel.
__global__ void k_segment_mean_std(const float* __restrict__ x, const int32_t* __restrict__ ptr,
                                   const int32_t* __restrict__ points, int64_t num_parents,
                                   int64_t C, float* __restrict__ mean_out,
                                   float* __restrict__ std_out) {
  int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = warp; p < num_parents; p += nwarps) {
    const int b = ptr[p], e = ptr[p + 1];
    const float inv = 1.f / (float)max(e - b, 1);
    const float den = (float)max(e - b - 1, 1) + 1e-6f;
    for (int64_t c = lane; c < C; c += 32) {
      float s = 0.f;
      for (int i = b; i < e; ++i) s += x[(int64_t)(points ? points[i] : i) * C + c];
      const float mu = s * inv;
      if (mean_out) mean_out[p * C + c] = mu;
      if (std_out) {
        float q = 0.f;
        for (int i = b; i < e; ++i) {
          const float d = x[(int64_t)(points ? points[i] : i) * C + c] - mu;
          q = fmaf(d, d, q);
        }
        std_out[p * C + c] = sqrtf(q / den);
      }
    }
  }
}

// ------------------------------------------------------------------ superedge features
// _minimalistic_horizontal_edge_features (reference src/transforms/graph.py:950-1060): one
// warp per superedge over its sub-edges (CSR of se_id): mean offset, std of the offsets in an
// orthonormal base built around the mean offset (src/utils/geometry.py:42-77, incl. its
// (1,0,0) / (2,1,-1) fall-backs and the in-place overwrite of a zero mean offset), clipped to
// [-2, 2], and sqrt of the mean sub-edge length.  out [num_se, 7].
__global__ void k_superedge_features(const float* __restrict__ points,
                                     const int64_t* __restrict__ sp_src,
                                     const int64_t* __restrict__ sp_dst,
                                     const int32_t* __restrict__ ptr,
                                     const int32_t* __restrict__ perm, int64_t num_se,
                                     float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t se = warp; se < num_se; se += nwarps) {
    const int b = ptr[se], e = ptr[se + 1];
    float sx = 0.f, sy = 0.f, sz = 0.f, sd = 0.f;
    for (int i = b + lane; i < e; i += 32) {
      const int64_t j = perm[i];
      const int64_t ps = sp_src[j], pt = sp_dst[j];
      const float ox = points[pt * 3] - points[ps * 3];
      const float oy = points[pt * 3 + 1] - points[ps * 3 + 1];
      const float oz = points[pt * 3 + 2] - points[ps * 3 + 2];
      sx += ox; sy += oy; sz += oz;
      sd += sqrtf(ox * ox + oy * oy + oz * oz);
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); sd = warp_sum(sd);
    const float inv = 1.f / (float)max(e - b, 1);
    float mx = sx * inv, my = sy * inv, mz = sz * inv;
    const float md = sd * inv;
    // base vectors (the reference overwrites a zero mean offset with (1,0,0) in place)
    float na = sqrtf(mx * mx + my * my + mz * mz);
    if (na == 0.f) { mx = 1.f; my = 0.f; mz = 0.f; na = 1.f; }
    const float ax = mx / na, ay = my / na, az = mz / na;
    float bx = ay - az, by = az - ax, bz = ax - ay;
    float nb = sqrtf(bx * bx + by * by + bz * bz);
    if (nb == 0.f) { bx = 2.f; by = 1.f; bz = -1.f; nb = sqrtf(6.f); }
    bx /= nb; by /= nb; bz /= nb;
    const float cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
    // mean of the projections, then their unbiased std (two passes like scatter_std)
    float pu = 0.f, pv = 0.f, pw = 0.f;
    for (int i = b + lane; i < e; i += 32) {
      const int64_t j = perm[i];
      const int64_t ps = sp_src[j], pt = sp_dst[j];
      const float ox = points[pt * 3] - points[ps * 3];
      const float oy = points[pt * 3 + 1] - points[ps * 3 + 1];
      const float oz = points[pt * 3 + 2] - points[ps * 3 + 2];
      pu += ox * ax + oy * ay + oz * az;
      pv += ox * bx + oy * by + oz * bz;
      pw += ox * cx + oy * cy + oz * cz;
    }
    pu = warp_sum(pu) * inv; pv = warp_sum(pv) * inv; pw = warp_sum(pw) * inv;
    float qu = 0.f, qv = 0.f, qw = 0.f;
    for (int i = b + lane; i < e; i += 32) {
      const int64_t j = perm[i];
      const int64_t ps = sp_src[j], pt = sp_dst[j];
      const float ox = points[pt * 3] - points[ps * 3];
      const float oy = points[pt * 3 + 1] - points[ps * 3 + 1];
      const float oz = points[pt * 3 + 2] - points[ps * 3 + 2];
      const float du = ox * ax + oy * ay + oz * az - pu;
      const float dv = ox * bx + oy * by + oz * bz - pv;
      const float dw = ox * cx + oy * cy + oz * cz - pw;
      qu = fmaf(du, du, qu); qv = fmaf(dv, dv, qv); qw = fmaf(dw, dw, qw);
    }
    qu = warp_sum(qu); qv = warp_sum(qv); qw = warp_sum(qw);
    if (lane == 0) {
      const float den = (float)max(e - b - 1, 1) + 1e-6f;
      float* o = out + se * 7;
      o[0] = mx; o[1] = my; o[2] = mz;
      o[3] = fminf(fmaxf(sqrtf(qu / den), -2.f), 2.f);
      o[4] = fminf(fmaxf(sqrtf(qv / den), -2.f), 2.f);
      o[5] = fminf(fmaxf(sqrtf(qw / den), -2.f), 2.f);
      o[6] = sqrtf(md);
    }
  }
}

static inline int warps_grid(int64_t rows, int threads, int64_t cap_blocks) {
  int64_t wpb = threads / 32;
  int64_t b = ceil_div(rows, wpb);
  if (b < 1) b = 1;
  if (b > cap_blocks) b = cap_blocks;
  return (int)b;
}

}  // namespace spt

using namespace spt;

extern "C" {

int spt_gather_rows_i64(const float* x, const int64_t* idx, int64_t n_out,
                        int64_t C, float* out, void* stream_) {
  SPT_REQUIRE(n_out >= 0 && C >= 0, SPT_E_INVALID, "gather_rows: negative size");
  if (n_out == 0 || C == 0) return SPT_OK;
  SPT_REQUIRE(x && idx && out, SPT_E_INVALID, "gather_rows: null pointer");
  cudaStream_t st = (cudaStream_t)stream_;
  int grid = warps_grid(n_out, 256, device_sm_count() * 32);
  if (C % 4 == 0)
    k_gather_rows<int64_t, 4><<<grid, 256, 0, st>>>(x, idx, n_out, C, out);
  else
    k_gather_rows<int64_t, 1><<<grid, 256, 0, st>>>(x, idx, n_out, C, out);
  return check_launch("gather_rows_i64");
}

int spt_gather_rows_i32(const float* x, const int32_t* idx, int64_t n_out,
                        int64_t C, float* out, void* stream_) {
  SPT_REQUIRE(n_out >= 0 && C >= 0, SPT_E_INVALID, "gather_rows: negative size");
  if (n_out == 0 || C == 0) return SPT_OK;
  SPT_REQUIRE(x && idx && out, SPT_E_INVALID, "gather_rows: null pointer");
  cudaStream_t st = (cudaStream_t)stream_;
  int grid = warps_grid(n_out, 256, device_sm_count() * 32);
  if (C % 4 == 0)
    k_gather_rows<int32_t, 4><<<grid, 256, 0, st>>>(x, idx, n_out, C, out);
  else
    k_gather_rows<int32_t, 1><<<grid, 256, 0, st>>>(x, idx, n_out, C, out);
  return check_launch("gather_rows_i32");
}

#define SPT_DISPATCH_REDUCE(KERNEL, ...)                                        \
  do {                                                                          \
    if (C % 4 == 0) {                                                           \
      switch (reduce) {                                                         \
        case SPT_REDUCE_SUM: KERNEL<SPT_REDUCE_SUM, 4> __VA_ARGS__; break;      \
        case SPT_REDUCE_MEAN: KERNEL<SPT_REDUCE_MEAN, 4> __VA_ARGS__; break;    \
        case SPT_REDUCE_MAX: KERNEL<SPT_REDUCE_MAX, 4> __VA_ARGS__; break;      \
        default: KERNEL<SPT_REDUCE_MIN, 4> __VA_ARGS__; break;                  \
      }                                                                         \
    } else {                                                                    \
      switch (reduce) {                                                         \
        case SPT_REDUCE_SUM: KERNEL<SPT_REDUCE_SUM, 1> __VA_ARGS__; break;      \
        case SPT_REDUCE_MEAN: KERNEL<SPT_REDUCE_MEAN, 1> __VA_ARGS__; break;    \
        case SPT_REDUCE_MAX: KERNEL<SPT_REDUCE_MAX, 1> __VA_ARGS__; break;      \
        default: KERNEL<SPT_REDUCE_MIN, 1> __VA_ARGS__; break;                  \
      }                                                                         \
    }                                                                           \
  } while (0)

int spt_segment_pool_fwd(const float* x, const int32_t* ptr, const int32_t* points,
                         int64_t num_parents, int64_t C, int reduce, float* out,
                         int32_t* arg, void* stream_) {
  SPT_REQUIRE(num_parents >= 0 && C >= 0, SPT_E_INVALID, "segment_pool_fwd: negative size");
  SPT_REQUIRE(reduce >= SPT_REDUCE_SUM && reduce <= SPT_REDUCE_MIN, SPT_E_INVALID,
              "segment_pool_fwd: bad reduce %d", reduce);
  if (num_parents == 0 || C == 0) return SPT_OK;
  SPT_REQUIRE(ptr && out, SPT_E_INVALID, "segment_pool_fwd: null pointer");
  cudaStream_t st = (cudaStream_t)stream_;
  int grid = warps_grid(num_parents, 128, device_sm_count() * 64);
  SPT_DISPATCH_REDUCE(k_segment_pool_fwd,
                      <<<grid, 128, 0, st>>>(x, ptr, points, num_parents, C, out, arg));
  return check_launch("segment_pool_fwd");
}

int spt_segment_pool_bwd(const float* dout, const int64_t* parent, const int32_t* ptr,
                         const int32_t* arg, int64_t num_children, int64_t C,
                         int reduce, float* dx, void* stream_) {
  SPT_REQUIRE(num_children >= 0 && C >= 0, SPT_E_INVALID, "segment_pool_bwd: negative size");
  SPT_REQUIRE(reduce >= SPT_REDUCE_SUM && reduce <= SPT_REDUCE_MIN, SPT_E_INVALID,
              "segment_pool_bwd: bad reduce %d", reduce);
  if (num_children == 0 || C == 0) return SPT_OK;
  SPT_REQUIRE(dout && parent && dx, SPT_E_INVALID, "segment_pool_bwd: null pointer");
  SPT_REQUIRE(reduce != SPT_REDUCE_MEAN || ptr, SPT_E_INVALID,
              "segment_pool_bwd: mean needs ptr");
  SPT_REQUIRE((reduce != SPT_REDUCE_MAX && reduce != SPT_REDUCE_MIN) || arg,
              SPT_E_INVALID, "segment_pool_bwd: max/min need arg");
  cudaStream_t st = (cudaStream_t)stream_;
  int grid = warps_grid(num_children, 256, device_sm_count() * 32);
  SPT_DISPATCH_REDUCE(k_segment_pool_bwd,
                      <<<grid, 256, 0, st>>>(dout, parent, ptr, arg, num_children, C, dx));
  return check_launch("segment_pool_bwd");
}

int spt_segment_mean_std_fwd(const float* x, const int32_t* ptr, const int32_t* points,
                             int64_t num_parents, int64_t C, float* mean_out, float* std_out,
                             void* stream_) {
  SPT_REQUIRE(num_parents >= 0 && C >= 0, SPT_E_INVALID, "segment_mean_std: negative size");
  if (num_parents == 0 || C == 0) return SPT_OK;
  SPT_REQUIRE(x && ptr && (mean_out || std_out), SPT_E_INVALID, "segment_mean_std: null pointer");
  int grid = warps_grid(num_parents, 128, device_sm_count() * 64);
  k_segment_mean_std<<<grid, 128, 0, (cudaStream_t)stream_>>>(x, ptr, points, num_parents, C,
                                                               mean_out, std_out);
  return check_launch("segment_mean_std_fwd");
}

int spt_superedge_features_fwd(const float* points, const int64_t* sp_src, const int64_t* sp_dst,
                               const int32_t* ptr, const int32_t* perm, int64_t num_se,
                               float* out, void* stream_) {
  SPT_REQUIRE(num_se >= 0, SPT_E_INVALID, "superedge_features: negative size");
  if (num_se == 0) return SPT_OK;
  SPT_REQUIRE(points && sp_src && sp_dst && ptr && perm && out, SPT_E_INVALID,
              "superedge_features: null pointer");
  int grid = warps_grid(num_se, 128, device_sm_count() * 64);
  k_superedge_features<<<grid, 128, 0, (cudaStream_t)stream_>>>(points, sp_src, sp_dst, ptr,
                                                                 perm, num_se, out);
  return check_launch("superedge_features_fwd");
}

size_t spt_unitsphere_workspace_bytes(int64_t num_parents) {
  if (num_parents < 0) return 0;
  return align_up((size_t)num_parents * 3 * sizeof(float), 256);
}

int spt_unitsphere_fwd(const float* pos, const int64_t* parent, const int32_t* ptr,
                       const int32_t* points, const float* w, int64_t N,
                       int64_t num_parents, float* pos_out, float* diameter,
                       void* ws, size_t ws_bytes, void* stream_) {
  SPT_REQUIRE(N >= 0 && num_parents >= 0, SPT_E_INVALID, "unitsphere: negative size");
  if (num_parents == 0) return SPT_OK;
  SPT_REQUIRE(ptr && diameter && ws && (N == 0 || (pos && pos_out)), SPT_E_INVALID,
              "unitsphere: null pointer");
  SPT_REQUIRE(parent || num_parents == 1, SPT_E_INVALID,
              "unitsphere: parent index required when num_parents > 1");
  SPT_REQUIRE(ws_bytes >= spt_unitsphere_workspace_bytes(num_parents), SPT_E_WORKSPACE,
              "unitsphere: workspace too small");
  cudaStream_t st = (cudaStream_t)stream_;
  float* center = (float*)ws;
  k_unitsphere_stats<<<(unsigned)ceil_div(num_parents * 32, 256), 256, 0, st>>>(
      pos, ptr, points, w, num_parents, center, diameter);
  if (N > 0)
    k_unitsphere_apply<<<(unsigned)ceil_div(N, 256), 256, 0, st>>>(pos, parent, center,
                                                                  diameter, N, pos_out);
  return check_launch("unitsphere_fwd");
}

}  // extern "C"

// ------------------------------------------------------------------ tf32 split
// x = hi + lo exactly, hi = x with the 13 low mantissa bits cleared (a tf32 number),
// lo = x - hi.  Feeding (hi,lo) pairs to three tensor-core TF32 GEMMs
// (hi*hi + lo*hi + hi*lo, fp32 accumulate) reproduces an fp32 GEMM to ~2^-21
// relative error ("3xTF32"); used by the dense projections (qkv, out_proj, MLPs).
namespace spt {
__global__ void k_split_tf32(const float* __restrict__ x, int64_t n4, int64_t n,
                             float* __restrict__ hi, float* __restrict__ lo) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 v = reinterpret_cast<const float4*>(x)[i];
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
    reinterpret_cast<float4*>(hi)[i] = h;
    reinterpret_cast<float4*>(lo)[i] = l;
  }
  // scalar tail
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    int64_t j = (n4 << 2) + threadIdx.x;
    float v = x[j];
    float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    hi[j] = h;
    lo[j] = v - h;
  }
}
}  // namespace spt

extern "C" int spt_split_tf32(const float* x, int64_t n, float* hi, float* lo, void* stream_) {
  SPT_REQUIRE(n >= 0, SPT_E_INVALID, "split_tf32: negative size");
  if (n == 0) return SPT_OK;
  SPT_REQUIRE(x && hi && lo, SPT_E_INVALID, "split_tf32: null pointer");
  SPT_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(hi) |
                reinterpret_cast<uintptr_t>(lo)) & 15) == 0,
              SPT_E_INVALID, "split_tf32: pointers must be 16-byte aligned");
  int64_t n4 = n >> 2;
  int64_t blocks = spt::ceil_div(n4 > 0 ? n4 : 1, 256);
  if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
  spt::k_split_tf32<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(x, n4, n, hi, lo);
  return spt::check_launch("split_tf32");
}
