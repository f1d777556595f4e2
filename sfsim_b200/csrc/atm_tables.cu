// Table kernels of the atmosphere-LUT build (one per reference loop nest) for sm_100a.
//
//   K1 k_transmittance_table        make-lookup-table of transmittance          (atmosphere.clj:114-128)
//   K2 k_surface_radiance_base      make-lookup-table of surface-radiance-base  (atmosphere.clj:131-137)
//   K3 k_first_order                ray-scatter of the first-order sources      (atmosphere.clj:170-200)
//   K4 k_point_scatter              point-scatter (dJ)                          (atmosphere.clj:203-222)
//   K5 k_surface_radiance           surface-radiance (dE)                       (atmosphere.clj:225-230)
//   K6 k_ray_scatter                ray-scatter from the dJ table (dS)          (atmosphere.clj:192-200)
//   K7 k_resample_*                 re-tabulation through forward o backward    (atmosphere_lut.clj:94-101)
//
// Sample positions are exactly the reference's.  What changes is who evaluates them: all inner
// transmittance samples of one view ray lie on that ray, which is a function of the (height,
// elevation) texel pair only, so one CTA owns one (height, elevation) pair, integrates the view
// ray once into shared memory and then serves all (light-elevation, heading) texels from it.
#include <algorithm>
#include <cstdlib>

#include "atm_kernel_common.cuh"

#ifndef ATMLUT_K4_UNROLL
#define ATMLUT_K4_UNROLL 1
#endif
#ifndef ATMLUT_FO_MIN_BLOCKS
#define ATMLUT_FO_MIN_BLOCKS 4
#endif

namespace atm {

constexpr int kPointScatterUnroll = ATMLUT_K4_UNROLL;   // directions per trip of the point-scatter loop

// ------------------------------------------------------------------ K1 / K2: 2-D tables in double

// make-lookup-table of transmittance over transmittance-space (atmosphere_lut.clj:74)
__global__ void k_transmittance_table(Params P, float4 *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int n = P.shapes.st[0] * P.shapes.st[1];
  if (i >= n) return;
  V3 x, v;
  bool above;
  transmittance_backward(P.planet, P.shapes.st, (double)(i / P.shapes.st[1]), (double)(i % P.shapes.st[1]), x, v,
                         above);
  double t[3];
  transmittance_dir(P.planet, P.medium, P.shapes.ray_steps, x, v, above, t);
  out[i] = make_float4((float)t[0], (float)t[1], (float)t[2], 0.0f);
}

// make-lookup-table of surface-radiance-base over surface-radiance-space (atmosphere_lut.clj:75)
__global__ void k_surface_radiance_base(Params P, float4 *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int n = P.shapes.se[0] * P.shapes.se[1];
  if (i >= n) return;
  V3 x, l;
  surface_radiance_backward(P.planet, P.shapes.se, (double)(i / P.shapes.se[1]), (double)(i % P.shapes.se[1]), x, l);
  double t[3];
  transmittance_dir(P.planet, P.medium, P.shapes.ray_steps, x, l, true, t);
  double m = mag(x);
  V3 normal = v3(x.x / m, x.y / m, x.z / m);
  double f = fmax(0.0, dot(normal, l));
  out[i] = make_float4((float)((t[0] * P.intensity[0]) * f), (float)((t[1] * P.intensity[1]) * f),
                       (float)((t[2] * P.intensity[2]) * f), 0.0f);
}

// ------------------------------------------------------------------ K3: first-order ray scatter

// acc_c[ch] = sum over outer samples k in [k0, k1) of
//   exp(-h(p_k)/scale_c) T(x->p_k)[ch] T(p_k->sun)[ch] [sun visible from p_k]
// for the texel with light direction l (atmosphere.clj:192-200 with the first-order sources :140-182).
__device__ __forceinline__ void first_order_texel(const Params &P, const ViewSmem &vs, V3 l, int k0, int k1, int kstride,
                                                  float acc0[3], float acc1[3], unsigned &esamples) {
  const int steps = P.shapes.ray_steps;
  const double rt2 = sqr(P.planet.radius + P.planet.height);
  const double r2 = sqr(P.planet.radius);
  const double inv_steps = 1.0 / (double)steps;
  const double ll = dot(l, l);
  const double llen = sqrt(ll);
  const double inv_ll = 1.0 / ll;   // the end point of the sun ray only places samples; an ulp there is harmless
  const double quad_b = (2.0 * inv_steps) * P.fast.inv_rp2, quad_c = ((ll * inv_steps) * inv_steps) * P.fast.inv_rp2;
  for (int k = k0; k < k1; k += kstride) {
    const double pkx = vs.pkx[k], pky = vs.pky[k], rk2 = vs.rk2[k];
    const double pl = l.x * pkx + l.y * pky;
    // filtered-sun-light (atmosphere.clj:154-160): is-above-horizon? (p_k, l)
    if (!(pl >= 0 || pl * pl <= rk2 - r2)) continue;
    // atmosphere-intersection of (p_k, l) (atmosphere.clj:70-77, sphere.clj:46-59)
    const double disc = pl * pl - ll * (rk2 - rt2);
    const double middle = -(pl * inv_ll);
    double t;
    if (disc > 0) {
      double length2 = sqrt(disc) * inv_ll;
      t = (middle < length2) ? fmax(0.0, middle + length2) : (middle - length2) + 2 * length2;
    } else {
      t = fmax(0.0, middle);
    }
    // transmittance p_k -> end point: samples p_k + (l t)(j + 1/2)/steps
    // make_quad(P.fast, rk2, pl t, ll t^2, steps) with the factors that do not depend on k worked out once per texel
    Quad q;
    q.A = (rk2 - P.fast.rp2) * P.fast.inv_rp2;
    q.B = quad_b * (pl * t);
    q.C = quad_c * (t * t);
    float s0, s1;
    density_sums_seq(P, q, steps, s0, s1);
    esamples += steps;
    const float seg = (float)(t * llen * inv_steps);
    float tr[3];
    transmittance_rgb(P.fast, fmaf(s0, seg, vs.cv0[k]), fmaf(s1, seg, vs.cv1[k]), tr);
    const float d0 = vs.dens0[k], d1 = vs.dens1[k];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      acc0[ch] = fmaf(d0, tr[ch], acc0[ch]);
      acc1[ch] = fmaf(d1, tr[ch], acc1[ch]);
    }
  }
}

// ray.clj:19-30: every term carries a = stepsize * |direction|; the source's constant factors
// (scatter-base, phase, intensity) are applied once per texel
__device__ __forceinline__ void first_order_store(const Params &P, const ViewRay &ray, V3 v, V3 l, size_t idx,
                                                  const float acc0[3], const float acc1[3], FirstOrderOut oa,
                                                  FirstOrderOut ob) {
  const double a = ray.dlen / (double)P.shapes.ray_steps;
  const double mu = dot(v, l);
  FirstOrderOut outs[2] = {oa, ob};
#pragma unroll
  for (int o = 0; o < 2; o++) {
    if (outs[o].out.n == 0) continue;
    const int c = outs[o].component;
    const float *acc = c ? acc1 : acc0;
    const double ph = outs[o].strength ? 1.0 : phase(P.medium.g[c], mu);
    float4 val;
    val.x = (float)(P.medium.base[c][0] * ph * P.intensity[0] * a * (double)acc[0]);
    val.y = (float)(P.medium.base[c][1] * ph * P.intensity[1] * a * (double)acc[1]);
    val.z = (float)(P.medium.base[c][2] * ph * P.intensity[2] * a * (double)acc[2]);
    val.w = 0.0f;
    store_all(outs[o].out, idx, val);
  }
}

// The view rays of a shard's pairs, once per build (see atm_kernel_common.cuh): slot = the pair's global index.
__global__ void __launch_bounds__(128) k_view_prepare(Params P, Shard shard, unsigned char *packs, unsigned long long *counter) {
  __shared__ ViewSmem vs;
  const int he = shard_pair(shard, blockIdx.x);
  const int E = P.shapes.s4[1];
  unsigned esamples = 0;
  setup_view_ray(P, he / E, he % E, vs, esamples);
  store_view_ray(vs, P.shapes.ray_steps, packs + (size_t)he * view_pack_bytes(P.shapes.ray_steps));
  count_esamples(counter, esamples);
}

// First-order ray scatter of both sources in one pass.  A (height, elevation) pair is served by `nchunks` CTAs; each
// reads the pair's view ray into shared memory (load_view_ray; without packs it integrates it itself) and owns
// `passes` x warps groups of 32 (light-elevation, heading) texels.  Texels are light-elevation major and low sun rows
// skip most samples (sun below the local horizon), so group cost grows with the group index: the static deal folds the
// groups over the chunks (g, 2T-1-g, 2T+g, ...) so that all CTAs of a pair get equal work, each group staying a
// coherent row block (no extra divergence); inside a CTA the warps take groups from a counter, expensive ones first.
// kparts > 1 (few texels per pair): one pass, the outer sample range is split over `kparts` thread groups.
__global__ void __launch_bounds__(256, ATMLUT_FO_MIN_BLOCKS) k_first_order(Params P, Shard shard, int he_count, int kparts, int passes, int nchunks,
                                                     FirstOrderOut oa, FirstOrderOut ob,
                                                     unsigned long long *counter, const unsigned char *view_packs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ViewSmem &vs = *reinterpret_cast<ViewSmem *>(smem_raw);
  float *partial = reinterpret_cast<float *>(smem_raw + sizeof(ViewSmem));  // [6][blockDim] when kparts > 1
  const int E = P.shapes.s4[1], S = P.shapes.s4[2], A = P.shapes.s4[3];
  const int ntex = S * A;
  // The CTAs of a pair are dealt out chunk-major, the highest light-elevation rows first: those rows see the sun from
  // most outer samples and cost the most, so the cheap CTAs are the ones that fill the tail of the launch.
  const int he = shard_pair(shard, blockIdx.x % he_count);
  const int chunk = nchunks - 1 - blockIdx.x / he_count;
  const int h = he / E, e = he % E;
  const int steps = P.shapes.ray_steps;
  unsigned esamples = 0;
  __shared__ int s_next_group;
  if (threadIdx.x == 0) s_next_group = 0;   // published by the barrier that ends the view-ray set-up
  if (view_packs)
    load_view_ray(vs, steps, view_packs + (size_t)he * view_pack_bytes(steps));
  else
    setup_view_ray(P, h, e, vs, esamples);
  const ViewRay ray = vs.ray;
  const V3 v = v3(ray.vx, ray.vy, 0.0);

  if (kparts == 1) {
    const int nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int total_warps = nchunks * nwarps;            // warps serving this pair
    // Lane layout.  Whether the sun is visible from p_k depends mostly on the light-elevation row and on k, and
    // visible k form an interval.  With narrow rows (A headings, A a power of two below 32) a warp therefore takes
    // ONE row and spreads the outer samples over 32 / A lane groups, k interleaved (k = q, q + 32/A, ...): lanes
    // then agree on visibility almost always.  Wide or odd rows fall back to 32 consecutive texels per warp.
    const bool row_layout = A < 32 && (A & (A - 1)) == 0;
    const int kq = row_layout ? 32 / A : 1;
    const int ngroups = row_layout ? S : (ntex + 31) >> 5;
    // The CTA's row groups (the same set as a static deal of `passes` groups per warp) are handed out through a counter
    // in shared memory, the expensive high-sun rows first: a warp that finishes early takes the next group instead of
    // idling until its CTA-mates are done (13 % of the warp slots of an SM were empty that way, ncu).
    for (;;) {
      int turn = 0;
      if (lane == 0) turn = atomicAdd(&s_next_group, 1);
      turn = __shfl_sync(0xffffffffu, turn, 0);
      if (turn >= passes * nwarps) break;
      const int pass = passes - 1 - turn / nwarps, slot = chunk * nwarps + turn % nwarps;
      const int group = pass * total_warps + ((pass & 1) ? total_warps - 1 - slot : slot);
      const int texel = row_layout ? group * A + (lane & (A - 1)) : group * 32 + lane;
      if (group >= ngroups || texel >= ntex) continue;
      const int si = texel / A, ai = texel % A;
      const double ss = index_to_sin_sun_elevation(S, (double)si);
      const V3 l = index_to_sun_direction(A, v, ss, (double)ai);
      float acc0[3] = {0.f, 0.f, 0.f}, acc1[3] = {0.f, 0.f, 0.f};
      first_order_texel(P, vs, l, row_layout ? lane / A : 0, steps, kq, acc0, acc1, esamples);
      if (kq > 1) {
        for (int o = A; o < 32; o <<= 1) {
#pragma unroll
          for (int ch = 0; ch < 3; ch++) {
            acc0[ch] += __shfl_xor_sync(0xffffffffu, acc0[ch], o);
            acc1[ch] += __shfl_xor_sync(0xffffffffu, acc1[ch], o);
          }
        }
        if (lane >= A) continue;
      }
      first_order_store(P, ray, v, l, (size_t)he * ntex + texel, acc0, acc1, oa, ob);
    }
  } else {
    const int items = ntex * kparts;
    const int item = threadIdx.x;
    const bool active = item < items;
    const int texel = active ? item % ntex : 0;
    const int part = active ? item / ntex : 0;
    float acc0[3] = {0.f, 0.f, 0.f}, acc1[3] = {0.f, 0.f, 0.f};
    V3 l = v3(0, 0, 0);
    if (active) {
      const int si = texel / A, ai = texel % A;
      const double ss = index_to_sin_sun_elevation(S, (double)si);
      l = index_to_sun_direction(A, v, ss, (double)ai);
      const int k0 = (int)(((long long)steps * part) / kparts), k1 = (int)(((long long)steps * (part + 1)) / kparts);
      first_order_texel(P, vs, l, k0, k1, 1, acc0, acc1, esamples);
    }
    __syncthreads();
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
      partial[ch * blockDim.x + threadIdx.x] = acc0[ch];
      partial[(3 + ch) * blockDim.x + threadIdx.x] = acc1[ch];
    }
    __syncthreads();
    if (active && part == 0) {
      for (int p = 1; p < kparts; p++) {
        const int other = p * ntex + texel;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
          acc0[ch] += partial[ch * blockDim.x + other];
          acc1[ch] += partial[(3 + ch) * blockDim.x + other];
        }
      }
      first_order_store(P, ray, v, l, (size_t)he * ntex + texel, acc0, acc1, oa, ob);
    }
  }
  count_esamples(counter, esamples);
}

// ------------------------------------------------------------------ K5: surface radiance

// per (surface height index, half-sphere direction): elevation coordinate of (x, omega, above = true)
__global__ void k_surface_radiance_prepare(Params P, const double *__restrict__ dirs, int ndirs, HalfDirInfo *info) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.shapes.se[0] * ndirs) return;
  const int h = i / ndirs, d = i % ndirs;
  V3 x = index_to_height(P.planet, P.shapes.se[0], (double)h);
  V3 omega = v3(dirs[3 * d], dirs[3 * d + 1], dirs[3 * d + 2]);
  Axis ae = axis_from(elevation_to_index(P.planet, P.shapes.s4[1], x, omega, true), P.shapes.s4[1]);
  HalfDirInfo r;
  r.eu = ae.u;
  r.ev = ae.v;
  r.es = ae.s;
  info[i] = r;
}

// dE[i] = integral-half-sphere of S(x, omega, l, true) (omega . n)   (atmosphere.clj:225-230)
__global__ void __launch_bounds__(256) k_surface_radiance(Params P, SSource src, const double *__restrict__ dirs,
                                                          const double *__restrict__ weights, int ndirs,
                                                          const HalfDirInfo *__restrict__ info, int first, int stride,
                                                          PeerOut out) {
  __shared__ float red[3][8];
  const int i = first + blockIdx.x * stride;
  const int hi = i / P.shapes.se[1], si = i % P.shapes.se[1];
  V3 x, l;
  surface_radiance_backward(P.planet, P.shapes.se, (double)hi, (double)si, x, l);
  const int H = P.shapes.s4[0], S = P.shapes.s4[2], A = P.shapes.s4[3];
  const Axis ah = axis_from(height_to_index(P.planet, H, x), H);
  const Axis as = axis_from(sun_elevation_to_index(S, x, l), S);
  const double m = mag(x);
  const V3 normal = v3(x.x / m, x.y / m, x.z / m);
  const HalfDirInfo *hinfo = info + (size_t)hi * ndirs;
  float acc[3] = {0.f, 0.f, 0.f};
  for (int d = threadIdx.x; d < ndirs; d += blockDim.x) {
    V3 omega = v3(dirs[3 * d], dirs[3 * d + 1], dirs[3 * d + 2]);
    const double mu = dot(omega, l);
    const Axis aa = axis_from((double)(A - 1) * ((1 + mu) / 2), A);
    Axis ae;
    ae.u = hinfo[d].eu;
    ae.v = hinfo[d].ev;
    ae.s = hinfo[d].es;
    float4 s = lookup4(src.tab_a, P.shapes.s4, ah, ae, as, aa);
    if (src.tab_b) {
      float4 mm = lookup4(src.tab_b, P.shapes.s4, ah, ae, as, aa);
      const float base = (float)((1.0 + src.phase_g * src.phase_g) - 2.0 * src.phase_g * mu);
      const float ph = (float)((3.0 * (1.0 - src.phase_g * src.phase_g)) / (8.0 * kPi * (2.0 + src.phase_g * src.phase_g))) *
                       (float)(1.0 + mu * mu) / (base * sqrtf(base));
      s.x = fmaf(mm.x, ph, s.x);
      s.y = fmaf(mm.y, ph, s.y);
      s.z = fmaf(mm.z, ph, s.z);
    }
    const float f = (float)(dot(omega, normal) * weights[d]);
    acc[0] = fmaf(s.x, f, acc[0]);
    acc[1] = fmaf(s.y, f, acc[1]);
    acc[2] = fmaf(s.z, f, acc[2]);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int ch = 0; ch < 3; ch++) {
    float v = warp_sum(acc[ch]);
    if (lane == 0) red[ch][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float r[3];
    for (int ch = 0; ch < 3; ch++) {
      float v = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += red[ch][w];
      r[ch] = v;
    }
    store_all(out, (size_t)i, make_float4(r[0], r[1], r[2], 0.0f));
  }
}

// ------------------------------------------------------------------ K7: re-tabulation

// out[i] = lookup(a, g(i)) [+ lookup(b, g(i))], g = ray-scatter-forward o ray-scatter-backward.
// file_layout != 0: write packed RGB float in convert-4d-to-2d order (image.clj:299-312).
__global__ void k_resample_4d(Params P, Shard shard, long long count, const float4 *__restrict__ a,
                              const float4 *__restrict__ b, PeerOut out, float *file_out) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // j-th texel of this launch
  if (j >= count) return;
  const int H = P.shapes.s4[0], E = P.shapes.s4[1], S = P.shapes.s4[2], A = P.shapes.s4[3];
  // texels are taken pair by pair: the n-th pair of the shard, all of its S*A texels
  const int ntex = S * A;
  const long long i = (long long)shard_pair(shard, (int)(j / ntex)) * ntex + j % ntex;
  const int ai = (int)(i % A), si = (int)((i / A) % S), ei = (int)((i / ((long long)A * S)) % E),
            hi = (int)(i / ((long long)A * S * E));
  V3 x, v, l;
  bool above;
  ray_scatter_backward(P.planet, P.shapes.s4, (double)hi, (double)ei, (double)si, (double)ai, x, v, l, above);
  double idx[4];
  ray_scatter_forward(P.planet, P.shapes.s4, x, v, l, above, idx);
  const Axis ah = axis_from(idx[0], H), ae = axis_from(idx[1], E), as = axis_from(idx[2], S), aa = axis_from(idx[3], A);
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a) r = lookup4(a, P.shapes.s4, ah, ae, as, aa);
  if (b) {
    float4 t = lookup4(b, P.shapes.s4, ah, ae, as, aa);
    r.x += t.x;
    r.y += t.y;
    r.z += t.z;
  }
  store_all(out, (size_t)i, r);
  if (file_out) {
    const long long y = (long long)hi * S + si, xx = (long long)ei * A + ai;
    float *o = file_out + (y * ((long long)E * A) + xx) * 3;
    o[0] = r.x;
    o[1] = r.y;
    o[2] = r.z;
  }
}

// which = 1: surface-radiance-space, which = 2: transmittance-space
__global__ void k_resample_2d(Params P, int which, const float4 *__restrict__ a, const float4 *__restrict__ b,
                              float4 *out, float *file_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int *shape = which == 1 ? P.shapes.se : P.shapes.st;
  if (i >= shape[0] * shape[1]) return;
  double idx[2];
  if (which == 1) {
    V3 x, l;
    surface_radiance_backward(P.planet, shape, (double)(i / shape[1]), (double)(i % shape[1]), x, l);
    surface_radiance_forward(P.planet, shape, x, l, idx);
  } else {
    V3 x, v;
    bool above;
    transmittance_backward(P.planet, shape, (double)(i / shape[1]), (double)(i % shape[1]), x, v, above);
    transmittance_forward(P.planet, shape, x, v, above, idx);
  }
  const Axis ar = axis_from(idx[0], shape[0]), ac = axis_from(idx[1], shape[1]);
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a) r = lookup2(a, shape[1], ar, ac);
  if (b) {
    float4 t = lookup2(b, shape[1], ar, ac);
    r.x += t.x;
    r.y += t.y;
    r.z += t.z;
  }
  if (out) out[i] = r;
  if (file_out) {
    file_out[3 * i] = r.x;
    file_out[3 * i + 1] = r.y;
    file_out[3 * i + 2] = r.z;
  }
}

// RGB float <-> padded float4 conversion for the host-facing table entry points
__global__ void k_rgb_to_float4(const float *in, float4 *out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float4(in[3 * i], in[3 * i + 1], in[3 * i + 2], 0.f);
}

__global__ void k_float4_to_rgb(const float4 *in, float *out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float4 v = in[i];
    out[3 * i] = v.x;
    out[3 * i + 1] = v.y;
    out[3 * i + 2] = v.z;
  }
}

// ------------------------------------------------------------------ cross-GPU barrier (peer-to-peer mode)

struct PeerFlags {
  unsigned *p[kMaxPeers];
};

// Thread q signals the next epoch into flag word [rank] of GPU q and then waits until GPU q has signalled this
// GPU.  The kernels whose stores must be visible ran earlier on the same stream; the system-scope fence
// orders them before the flag.  Epochs only grow, so a GPU that runs ahead never confuses a slower one; the
// counter lives in device memory so that the launch can be replayed from a CUDA graph.
// A peer that never arrives (crashed process) trips the clock-based timeout instead of hanging the box.
__global__ void k_peer_barrier(unsigned *local_flags, PeerFlags peers, int offset, int set, int rank, int world,
                               unsigned *epochs, int *error_flag, long long timeout_clocks) {
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) {
    s_epoch = epochs[set] + 1;
    epochs[set] = s_epoch;
  }
  __syncthreads();
  const unsigned epoch = s_epoch;
  const int q = threadIdx.x;
  if (q >= world) return;
  local_flags += offset;   // independent flag sets for the two streams
  __threadfence_system();
  volatile unsigned *remote = peers.p[q] + offset + rank;
  *remote = epoch;
  __threadfence_system();
  volatile unsigned *mine = local_flags + q;
  const long long start = clock64();
  while ((int)(*mine - epoch) < 0) {
    if (clock64() - start > timeout_clocks) {
      atomicExch(error_flag, 1);
      break;
    }
    __nanosleep(100);
  }
  __threadfence_system();
}

// ------------------------------------------------------------------ launchers

static int div_up(long long a, int b) { return (int)((a + b - 1) / b); }

cudaError_t launch_transmittance_table(const Params &P, float4 *out, cudaStream_t st) {
  int n = P.shapes.st[0] * P.shapes.st[1];
  k_transmittance_table<<<div_up(n, 64), 64, 0, st>>>(P, out);
  return cudaGetLastError();
}

cudaError_t launch_surface_radiance_base(const Params &P, float4 *out, cudaStream_t st) {
  int n = P.shapes.se[0] * P.shapes.se[1];
  k_surface_radiance_base<<<div_up(n, 64), 64, 0, st>>>(P, out);
  return cudaGetLastError();
}

static int env_int(const char *name, int fallback);

// CTAs a launch should at least have before the work of a (height, elevation) pair is left in one CTA:
// six waves of 4 CTAs on every SM of the current device
static int small_grid_ctas() {
  static const int forced = env_int("ATMLUT_MIN_CTAS", 0);
  if (forced > 0) return forced;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return 6 * 4 * sms;
}

static int env_int(const char *name, int fallback) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : fallback;
}

size_t view_pack_total_bytes(const Params &P) {
  return (size_t)P.shapes.s4[0] * P.shapes.s4[1] * view_pack_bytes(P.shapes.ray_steps);
}

cudaError_t launch_view_prepare(const Params &P, Shard shard, int he_count, void *packs, unsigned long long *counter,
                                cudaStream_t st) {
  if (he_count <= 0) return cudaSuccess;
  k_view_prepare<<<he_count, 128, 0, st>>>(P, shard, (unsigned char *)packs, counter);
  return cudaGetLastError();
}

cudaError_t launch_first_order(const Params &P, Shard shard, int he_count, FirstOrderOut oa, FirstOrderOut ob,
                               unsigned long long *counter, const void *view_packs, cudaStream_t st) {
  if (he_count <= 0) return cudaSuccess;
  // tuning knobs (defaults chosen on B200, see profiles/): warps per CTA and texel groups per warp
  static const int max_warps = std::min(8, std::max(1, env_int("ATMLUT_FIRST_ORDER_WARPS", 4)));
  static const int want_passes = std::max(1, env_int("ATMLUT_FIRST_ORDER_PASSES", 4));
  int warps = max_warps;
  const int ntex = P.shapes.s4[2] * P.shapes.s4[3];
  int kparts = 1, passes = 1, nchunks = 1;
  if (ntex < warps * 32) {
    kparts = std::min(P.shapes.ray_steps, warps * 32 / ntex);
  } else {
    const int A = P.shapes.s4[3];
    const bool row_layout = A < 32 && (A & (A - 1)) == 0;     // must match the kernel
    const int ngroups = row_layout ? P.shapes.s4[2] : (ntex + 31) / 32;
    passes = std::min(want_passes, (ngroups + warps - 1) / warps);
    // Small grids (one shard of a multi-GPU build, small tables): fewer row groups per warp, then fewer warps per
    // CTA, i.e. more and smaller CTAs per pair, until the launch spans several waves -- the repeated view-ray
    // set-up (steps^2 samples per CTA) is cheaper than SMs idling behind the last, expensive CTAs of a short launch.
    const int min_ctas = small_grid_ctas();
    for (;;) {
      const int total_warps = (ngroups + passes - 1) / passes;
      nchunks = (total_warps + warps - 1) / warps;
      if ((long long)he_count * nchunks >= min_ctas) break;
      if (passes > 1)
        passes = (passes + 1) / 2;
      else if (warps > 4)
        warps /= 2;
      else
        break;
    }
  }
  const int threads = warps * 32;
  size_t smem = sizeof(ViewSmem) + (kparts > 1 ? 6 * threads * sizeof(float) : 0);
  k_first_order<<<he_count * nchunks, threads, smem, st>>>(P, shard, he_count, kparts, passes, nchunks, oa, ob, counter,
                                                          (const unsigned char *)view_packs);
  return cudaGetLastError();
}

cudaError_t launch_surface_radiance_prepare(const Params &P, const double *dirs, int ndirs, HalfDirInfo *info,
                                            cudaStream_t st) {
  int n = P.shapes.se[0] * ndirs;
  k_surface_radiance_prepare<<<div_up(n, 128), 128, 0, st>>>(P, dirs, ndirs, info);
  return cudaGetLastError();
}

cudaError_t launch_surface_radiance(const Params &P, SSource src, const double *dirs, const double *weights,
                                    int ndirs, const HalfDirInfo *info, int first, int stride, PeerOut out,
                                    cudaStream_t st) {
  const int n = P.shapes.se[0] * P.shapes.se[1];
  const int count = n > first ? (n - first + stride - 1) / stride : 0;
  if (count <= 0) return cudaSuccess;
  k_surface_radiance<<<count, 256, 0, st>>>(P, src, dirs, weights, ndirs, info, first, stride, out);
  return cudaGetLastError();
}

cudaError_t launch_resample_4d(const Params &P, Shard shard, int pair_count, const float4 *a, const float4 *b,
                               PeerOut out, float *file_out, cudaStream_t st) {
  const long long count = (long long)pair_count * P.shapes.s4[2] * P.shapes.s4[3];
  if (count <= 0) return cudaSuccess;
  k_resample_4d<<<div_up(count, 128), 128, 0, st>>>(P, shard, count, a, b, out, file_out);
  return cudaGetLastError();
}

cudaError_t launch_resample_2d(const Params &P, int which, const float4 *a, const float4 *b, float4 *out,
                               float *file_out, cudaStream_t st) {
  const int *shape = which == 1 ? P.shapes.se : P.shapes.st;
  int n = shape[0] * shape[1];
  k_resample_2d<<<div_up(n, 128), 128, 0, st>>>(P, which, a, b, out, file_out);
  return cudaGetLastError();
}

cudaError_t launch_peer_barrier(unsigned *local_flags, unsigned *const *peer_flags, int flag_set, int rank, int world,
                                unsigned *epochs, int *error_flag, int timeout_ms, cudaStream_t st) {
  PeerFlags f = {};
  for (int q = 0; q < world && q < kMaxPeers; q++) f.p[q] = peer_flags[q];
  int dev = 0, khz = 2000000;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const long long timeout_clocks = (long long)khz * (long long)(timeout_ms > 0 ? timeout_ms : 10000);
  k_peer_barrier<<<1, 32, 0, st>>>(local_flags, f, flag_set * kMaxPeers, flag_set, rank, world, epochs, error_flag,
                                   timeout_clocks);
  return cudaGetLastError();
}

cudaError_t launch_rgb_to_float4(const float *in, float4 *out, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_rgb_to_float4<<<div_up(n, 256), 256, 0, st>>>(in, out, n);
  return cudaGetLastError();
}

cudaError_t launch_float4_to_rgb(const float4 *in, float *out, long long n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_float4_to_rgb<<<div_up(n, 256), 256, 0, st>>>(in, out, n);
  return cudaGetLastError();
}

}  // namespace atm
