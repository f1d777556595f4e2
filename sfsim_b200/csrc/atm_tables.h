// Host-visible declarations of the table kernels' launchers (atm_tables.cu).
#pragma once

#include <cuda_runtime.h>

#include "atm_device.cuh"

namespace atm {

constexpr int kMaxDirs = 1536;  // sphere quadrature directions the point-scatter kernel keeps in shared memory

constexpr int kMaxPeers = 8;    // GPUs of one NVSwitch box

// Which (height, elevation) pairs a launch integrates: CTA i owns pair begin + (i / run) * stride + i % run.
//   {0, 1, 1}               every pair (single GPU)
//   {first, 1, 1}           contiguous slab (NCCL mode: one all-gather reassembles the table)
//   {r * E, world * E, E}   whole height rows r, r + world, ... (peer-to-peer mode: finished texels are stored into
//                           every GPU's table, so there is no layout constraint; whole rows keep the pre-blended
//                           direction tiles of the point-scatter kernel local to the rank that needs them)
struct Shard {
  int begin, stride, run;
};

// One logical output table: the local copy (p[0]) and the same table on the peer GPUs, mapped through
// CUDA IPC.  Kernels store each finished texel to all of them (16-byte posted writes over NVLink), which
// overlaps the "all-gather" with the integration itself.
struct PeerOut {
  float4 *p[kMaxPeers];
  int n;   // 0 = output not wanted
};

inline PeerOut local_out(float4 *table) {
  PeerOut o = {};
  o.p[0] = table;
  o.n = table ? 1 : 0;
  return o;
}

// one output of the first-order kernel: ray-scatter of point-scatter-component (strength = 0) or
// strength-component (strength = 1) of scatter[component]
struct FirstOrderOut {
  PeerOut out;
  int component;
  int strength;
};

// S source of point-scatter / surface-radiance: tab_a, or tab_a + tab_b * phase(phase_g, v.l)
// (the closure of atmosphere_lut.clj:79-84)
struct SSource {
  const float4 *tab_a;
  const float4 *tab_b;
  double phase_g;
};

// per (height index, sphere direction) constants of in-scatter-from-direction (atmosphere.clj:208-222)
struct DirInfo {
  int surface;        // surface-point? of the ray extremity
  int eu, ev;         // elevation axis corners of S(x, omega, l, not surface)
  float es;
  float tb[3];        // T(x -> point) * brightness / pi
  int ehu, ehv;       // height axis corners of E(point, l)
  float ehs;
  double nx, ny, nz;  // the ray extremity `point` ...
  double nmag;        // ... and |point|
};

struct HalfDirInfo {
  int eu, ev;
  float es;
};

cudaError_t launch_transmittance_table(const Params &P, float4 *out, cudaStream_t st);
cudaError_t launch_surface_radiance_base(const Params &P, float4 *out, cudaStream_t st);
// The view ray of every (height, elevation) pair of a shard -- outer sample points, column densities x -> p_k, densities
// at p_k -- written once per build (view_pack_total_bytes(P) bytes, slot = global pair index) and read by the
// first-order kernel and by launch_ray_prepare instead of working it out per CTA.  view_packs may be NULL there: the
// kernels then integrate the view ray themselves.
size_t view_pack_total_bytes(const Params &P);
cudaError_t launch_view_prepare(const Params &P, Shard shard, int he_count, void *packs, unsigned long long *counter,
                                cudaStream_t st);
cudaError_t launch_first_order(const Params &P, Shard shard, int he_count, FirstOrderOut oa, FirstOrderOut ob,
                               unsigned long long *counter, const void *view_packs, cudaStream_t st);
// exp_table: device array of kExpTabSize doubles, exp(i/64) for i = kExpTabLo .. kExpTabHi (see exp_tab)
// The ray-scatter kernel reads per-(pair, outer sample) records that do not depend on the scattering order:
// launch_ray_prepare writes them once per build (ray_sample_bytes(P, he_count) bytes for the same shard and
// he_count, slot i = the i-th pair of the shard); counter receives the overall-extinction samples it evaluated.
bool ray_scatter_uses_samples(const Params &P);
size_t ray_sample_bytes(const Params &P, int he_count);
cudaError_t launch_ray_prepare(const Params &P, Shard shard, int he_count, void *samples, unsigned long long *counter,
                               const void *view_packs, cudaStream_t st);
cudaError_t launch_ray_scatter(const Params &P, Shard shard, int he_count, const void *samples, const float4 *dj,
                               const double *exp_table, PeerOut out, unsigned long long *counter, cudaStream_t st);
// cross-GPU barrier over peer-mapped flag words: signal the next epoch to every peer, wait for every peer's signal.
// flag_set 0 / 1: independent barriers for the main and the side stream (kMaxPeers words each).  The epoch lives
// in device memory (epochs[flag_set], incremented by the kernel), so a captured CUDA graph can be replayed.
// timeout_ms: a peer that never arrives raises *error_flag instead of hanging the box.
cudaError_t launch_peer_barrier(unsigned *local_flags, unsigned *const *peer_flags, int flag_set, int rank, int world,
                                unsigned *epochs, int *error_flag, int timeout_ms, cudaStream_t st);
cudaError_t launch_point_scatter_prepare(const Params &P, const double *dirs, int ndirs, DirInfo *info,
                                         cudaStream_t st);
// blends the tiles of height indices h_first, h_first + h_stride, ... (h_count of them)
cudaError_t launch_blend_dir_tiles(const Params &P, const float4 *tab, const DirInfo *info, int ndirs, int h_first,
                                   int h_stride, int h_count, float4 *tiles, cudaStream_t st);
// tiles_a / tiles_b: blended [height][direction][light-elevation][heading] tiles of the S source
cudaError_t launch_point_scatter(const Params &P, Shard shard, int he_count, const float4 *tiles_a,
                                 const float4 *tiles_b, double phase_g, const float4 *de, const double *dirs,
                                 const double *weights, int ndirs, const DirInfo *info, const double *exp_table,
                                 PeerOut out, cudaStream_t st);
size_t ray_scatter_smem(const Params &P);
cudaError_t launch_surface_radiance_prepare(const Params &P, const double *dirs, int ndirs, HalfDirInfo *info,
                                            cudaStream_t st);
// texels first, first + stride, ... of the surface-radiance table (every texel for {0, 1})
cudaError_t launch_surface_radiance(const Params &P, SSource src, const double *dirs, const double *weights,
                                    int ndirs, const HalfDirInfo *info, int first, int stride, PeerOut out,
                                    cudaStream_t st);
// re-tabulates the pairs shard.begin + n * shard.stride, n < pair_count
cudaError_t launch_resample_4d(const Params &P, Shard shard, int pair_count, const float4 *a, const float4 *b,
                               PeerOut out, float *file_out, cudaStream_t st);
cudaError_t launch_resample_2d(const Params &P, int which, const float4 *a, const float4 *b, float4 *out,
                               float *file_out, cudaStream_t st);
cudaError_t launch_rgb_to_float4(const float *in, float4 *out, long long n, cudaStream_t st);
cudaError_t launch_float4_to_rgb(const float4 *in, float *out, long long n, cudaStream_t st);

}  // namespace atm
