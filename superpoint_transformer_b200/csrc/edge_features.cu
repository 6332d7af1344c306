// edge_features.cu — on-the-fly horizontal edge features + self-loops in one pass.
//
// Reference: _on_the_fly_horizontal_edge_features (src/transforms/graph.py:1137-1277,
// ~40 tiny elementwise/gather launches per level) followed by NAGAddSelfLoops
// (src/transforms/graph.py:1419-1452, PyG add_self_loops fill_value=0).
// One thread per trimmed edge writes both directions (rows e and Eh+e); a second
// grid range writes the N zero-feature self-loops.  Output column order follows the
// reference's f_list assembly exactly (mean_off is *prepended*, graph.py:1216-1220).
#include "common.cuh"

namespace spt {

constexpr int kEF = 18;

__device__ __forceinline__ float nan_to_zero_clip(float x) {
  // se_direction[isnan] = 0 ; clip(-1, 1)   (graph.py:1211-1212, :1256-1257)
  if (isnan(x)) return 0.f;
  return fminf(fmaxf(x, -1.f), 1.f);
}

__global__ void k_edge_features(const int64_t* __restrict__ se, const float* __restrict__ ea,
                                const float* __restrict__ pos, const float* __restrict__ normal,
                                const float* __restrict__ log_length,
                                const float* __restrict__ log_surface,
                                const float* __restrict__ log_volume,
                                const float* __restrict__ log_size, int64_t Eh, int64_t N,
                                int64_t E_out, int add_self_loops,
                                int64_t* __restrict__ ei_out, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Eh) {
    const int64_t s = se[i], t = se[Eh + i];
    float f[kEF], g[kEF];  // forward edge (s->t) and flipped edge (t->s)
    // mean_off (3): +/-
    float mo[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      mo[d] = ea[i * 7 + d];
      f[d] = mo[d];
      g[d] = -mo[d];
    }
    // std_off (3), mean_dist (1): same both ways
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      f[3 + d] = ea[i * 7 + 3 + d];
      g[3 + d] = f[3 + d];
    }
    // direction of the mean offset
    float nrm = sqrtf(mo[0] * mo[0] + mo[1] * mo[1] + mo[2] * mo[2]);
    float dir[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) dir[d] = nan_to_zero_clip(mo[d] / nrm);
    float ns[3], nt[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      ns[d] = normal[s * 3 + d];
      nt[d] = normal[t * 3 + d];
    }
    float as = fabsf(dir[0] * ns[0] + dir[1] * ns[1] + dir[2] * ns[2]);
    float at = fabsf(dir[0] * nt[0] + dir[1] * nt[1] + dir[2] * nt[2]);
    f[7] = as; g[7] = as;   // angle_source: cat((f, f))  (graph.py:1225)
    f[8] = at; g[8] = at;   // angle_target: cat((f, f))  (graph.py:1230)
    float na = fabsf(ns[0] * nt[0] + ns[1] * nt[1] + ns[2] * nt[2]);
    f[9] = na; g[9] = na;   // normal_angle
    float v;
    v = log_length[s] - log_length[t];   f[10] = v; g[10] = -v;
    v = log_surface[s] - log_surface[t]; f[11] = v; g[11] = -v;
    v = log_volume[s] - log_volume[t];   f[12] = v; g[12] = -v;
    v = log_size[s] - log_size[t];       f[13] = v; g[13] = -v;
    // centroid direction / sqrt-distance (graph.py:1249-1263)
    float cd[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) cd[d] = pos[t * 3 + d] - pos[s * 3 + d];
    float dist = sqrtf(cd[0] * cd[0] + cd[1] * cd[1] + cd[2] * cd[2]);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float c = nan_to_zero_clip(cd[d] / dist);
      f[14 + d] = c;
      g[14 + d] = -c;
    }
    float sd = sqrtf(dist);
    f[17] = sd; g[17] = sd;

    float* of = out + i * kEF;
    float* og = out + (Eh + i) * kEF;
#pragma unroll
    for (int c = 0; c < kEF; ++c) {
      of[c] = f[c];
      og[c] = g[c];
    }
    ei_out[i] = s;
    ei_out[E_out + i] = t;
    ei_out[Eh + i] = t;
    ei_out[E_out + Eh + i] = s;
  } else if (add_self_loops && i < Eh + N) {
    int64_t n = i - Eh;
    int64_t r = 2 * Eh + n;
    float* o = out + r * kEF;
#pragma unroll
    for (int c = 0; c < kEF; ++c) o[c] = 0.f;
    ei_out[r] = n;
    ei_out[E_out + r] = n;
  }
}

}  // namespace spt

using namespace spt;

extern "C" int spt_edge_features_fwd(const int64_t* se, const float* ea, const float* pos,
                                     const float* normal, const float* log_length,
                                     const float* log_surface, const float* log_volume,
                                     const float* log_size, int64_t Eh, int64_t N,
                                     int add_self_loops, int64_t* edge_index_out,
                                     float* edge_attr_out, void* stream_) {
  SPT_REQUIRE(Eh >= 0 && N >= 0, SPT_E_INVALID, "edge_features: negative size");
  int64_t E_out = 2 * Eh + (add_self_loops ? N : 0);
  int64_t work = Eh + (add_self_loops ? N : 0);
  if (work == 0) return SPT_OK;
  SPT_REQUIRE(edge_index_out && edge_attr_out, SPT_E_INVALID, "edge_features: null output");
  SPT_REQUIRE(Eh == 0 || (se && ea && pos && normal && log_length && log_surface &&
                          log_volume && log_size),
              SPT_E_INVALID, "edge_features: null input");
  k_edge_features<<<(unsigned)ceil_div(work, 256), 256, 0, (cudaStream_t)stream_>>>(
      se, ea, pos, normal, log_length, log_surface, log_volume, log_size, Eh, N, E_out,
      add_self_loops, edge_index_out, edge_attr_out);
  return check_launch("edge_features_fwd");
}

// ---------------------------------------------------------------------------
// On-the-fly VERTICAL (child -> parent) edge features, default key set
// (src/transforms/graph.py:1336-1416): centroid_dir(3) centroid_dist normal_angle
// log_length log_surface log_volume log_size  -> v_edge_attr [Nc, 9]
// ---------------------------------------------------------------------------
namespace spt {
constexpr int kVF = 9;
__global__ void k_vertical_edge_features(
    const float* __restrict__ cpos, const float* __restrict__ ppos,
    const float* __restrict__ cnrm, const float* __restrict__ pnrm,
    const float* __restrict__ cl0, const float* __restrict__ cl1, const float* __restrict__ cl2,
    const float* __restrict__ cl3, const float* __restrict__ pl0, const float* __restrict__ pl1,
    const float* __restrict__ pl2, const float* __restrict__ pl3,
    const int64_t* __restrict__ parent, int64_t Nc, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Nc) return;
  const int64_t p = parent[i];
  float d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) d[k] = ppos[p * 3 + k] - cpos[i * 3 + k];
  const float dist = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  float f[kVF];
#pragma unroll
  for (int k = 0; k < 3; ++k) f[k] = nan_to_zero_clip(d[k] / dist);
  f[3] = sqrtf(dist);
  f[4] = fabsf(cnrm[i * 3] * pnrm[p * 3] + cnrm[i * 3 + 1] * pnrm[p * 3 + 1] +
               cnrm[i * 3 + 2] * pnrm[p * 3 + 2]);
  f[5] = pl0[p] - cl0[i];
  f[6] = pl1[p] - cl1[i];
  f[7] = pl2[p] - cl2[i];
  f[8] = pl3[p] - cl3[i];
#pragma unroll
  for (int k = 0; k < kVF; ++k) out[i * kVF + k] = f[k];
}
}  // namespace spt

extern "C" int spt_vertical_edge_features_fwd(
    const float* child_pos, const float* parent_pos, const float* child_normal,
    const float* parent_normal, const float* child_logs /*4 arrays, see below*/,
    const float* parent_logs, const int64_t* parent, int64_t Nc, int64_t Np, float* out,
    void* stream_) {
  // child_logs / parent_logs: [4, Nc] / [4, Np] row-major = (log_length, log_surface,
  // log_volume, log_size)
  SPT_REQUIRE(Nc >= 0 && Np >= 0, SPT_E_INVALID, "vertical_edge_features: negative size");
  if (Nc == 0) return SPT_OK;
  SPT_REQUIRE(child_pos && parent_pos && child_normal && parent_normal && child_logs &&
                  parent_logs && parent && out,
              SPT_E_INVALID, "vertical_edge_features: null pointer");
  spt::k_vertical_edge_features<<<(unsigned)spt::ceil_div(Nc, 256), 256, 0,
                                  (cudaStream_t)stream_>>>(
      child_pos, parent_pos, child_normal, parent_normal, child_logs, child_logs + Nc,
      child_logs + 2 * Nc, child_logs + 3 * Nc, parent_logs, parent_logs + Np,
      parent_logs + 2 * Np, parent_logs + 3 * Np, parent, Nc, out);
  return spt::check_launch("vertical_edge_features_fwd");
}
