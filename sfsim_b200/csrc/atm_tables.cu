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

#include "atm_device.cuh"
#include "atm_tables.h"

#ifndef ATMLUT_K4_UNROLL
#define ATMLUT_K4_UNROLL 1
#endif
#ifndef ATMLUT_FO_MIN_BLOCKS
#define ATMLUT_FO_MIN_BLOCKS 4
#endif

namespace atm {

constexpr int kPointScatterUnroll = ATMLUT_K4_UNROLL;   // directions per trip of the point-scatter loop

// ------------------------------------------------------------------ view ray shared by one CTA

struct ViewRay {
  double r;        // |x|, x = (r, 0, 0)
  double vx, vy;   // view direction (atmosphere.clj:256-270)
  double dx, dy;   // ray end point minus x (atmosphere.clj:196-198)
  double dlen;     // |d|
  int above;
};

struct alignas(16) ViewSmem {
  ViewRay ray;
  double pkx[kMaxSteps], pky[kMaxSteps];  // outer sample points p_k
  double rk2[kMaxSteps];                  // |p_k|^2
  float cv0[kMaxSteps], cv1[kMaxSteps];   // column densities x -> p_k per component
  float dens0[kMaxSteps], dens1[kMaxSteps];  // exp(-h(p_k)/scale_c)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Fills ViewSmem for texel pair (h, e) of the 4-D space: the view ray (ray-scatter-backward,
// atmosphere.clj:401-412; ray end point :196-198), its `steps` outer sample points (ray.clj:19-30)
// and the transmittance integral x -> p_k for every k (atmosphere.clj:118-125, :199).
__device__ void setup_view_ray(const Params &P, int h, int e, ViewSmem &vs, unsigned &esamples) {
  const int steps = P.shapes.ray_steps;
  if (threadIdx.x == 0) {
    V3 x = index_to_height(P.planet, P.shapes.s4[0], (double)h);
    V3 v;
    bool above;
    index_to_elevation(P.planet, P.shapes.s4[1], x.x, (double)e, v, above);
    V3 end = above ? atmosphere_intersection(P.planet, x, v) : surface_intersection(P.planet, x, v);
    V3 d = end - x;
    vs.ray.r = x.x;
    vs.ray.vx = v.x;
    vs.ray.vy = v.y;
    vs.ray.dx = d.x;
    vs.ray.dy = d.y;
    vs.ray.dlen = mag(d);
    vs.ray.above = above ? 1 : 0;
  }
  __syncthreads();
  const ViewRay ray = vs.ray;
  const double stepsize = 1.0 / (double)steps;
  for (int k = threadIdx.x; k < steps; k += blockDim.x) {
    double s = (0.5 + (double)k) * stepsize;
    double px = ray.r + ray.dx * s, py = ray.dy * s;
    double r2 = px * px + py * py;
    double hk = sqrt(r2) - P.planet.radius;
    vs.pkx[k] = px;
    vs.pky[k] = py;
    vs.rk2[k] = r2;
    vs.dens0[k] = (float)exp(-(hk / P.medium.scale[0]));
    vs.dens1[k] = (float)exp(-(hk / P.medium.scale[1]));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int k = warp; k < steps; k += nwarps) {
    // segment x -> p_k, direction d_k = p_k - x
    double dkx = vs.pkx[k] - ray.r, dky = vs.pky[k];
    double dd = dkx * dkx + dky * dky;
    Quad q = make_quad(P.fast, ray.r * ray.r, ray.r * dkx, dd, steps);
    float s0, s1;
    density_sums_strided(P, q, steps, lane, 32, s0, s1, esamples);
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    if (lane == 0) {
      double seg = stepsize * sqrt(dd);
      vs.cv0[k] = (float)((double)s0 * seg);
      vs.cv1[k] = (float)((double)s1 * seg);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void store_all(const PeerOut &o, size_t idx, float4 v) {
#pragma unroll 1
  for (int q = 0; q < o.n; q++) o.p[q][idx] = v;
}

__device__ __forceinline__ void count_esamples(unsigned long long *counter, unsigned n) {
  if (!counter) return;
  n = (unsigned)__reduce_add_sync(0xffffffffu, n);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(counter, (unsigned long long)n);
}

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
    Quad q = make_quad(P.fast, rk2, pl * t, ll * t * t, steps);
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

// First-order ray scatter of both sources in one pass.  A (height, elevation) pair is served by `nchunks`
// CTAs; each integrates the view ray once into shared memory (setup_view_ray) and then every warp takes
// `passes` groups of 32 (light-elevation, heading) texels.  Texels are light-elevation major and low sun
// rows skip most samples (sun below the local horizon), so group cost grows with the group index: groups
// are dealt to warps in folded order (g, 2T-1-g, 2T+g, ...) so that all warps of a pair finish together
// while each group stays a coherent row block (no extra divergence).
// kparts > 1 (few texels per pair): one pass, the outer sample range is split over `kparts` thread groups.
__global__ void __launch_bounds__(256, ATMLUT_FO_MIN_BLOCKS) k_first_order(Params P, Shard shard, int kparts, int passes, int nchunks,
                                                     FirstOrderOut oa, FirstOrderOut ob,
                                                     unsigned long long *counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ViewSmem &vs = *reinterpret_cast<ViewSmem *>(smem_raw);
  float *partial = reinterpret_cast<float *>(smem_raw + sizeof(ViewSmem));  // [6][blockDim] when kparts > 1
  const int E = P.shapes.s4[1], S = P.shapes.s4[2], A = P.shapes.s4[3];
  const int ntex = S * A;
  const int he = shard.begin + (blockIdx.x / nchunks) * shard.stride;
  const int chunk = blockIdx.x % nchunks;
  const int h = he / E, e = he % E;
  const int steps = P.shapes.ray_steps;
  unsigned esamples = 0;
  setup_view_ray(P, h, e, vs, esamples);
  const ViewRay ray = vs.ray;
  const V3 v = v3(ray.vx, ray.vy, 0.0);

  if (kparts == 1) {
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_warps = nchunks * nwarps;            // warps serving this pair
    const int wg = chunk * nwarps + warp;
    // Lane layout.  Whether the sun is visible from p_k depends mostly on the light-elevation row and on k, and
    // visible k form an interval.  With narrow rows (A headings, A a power of two below 32) a warp therefore takes
    // ONE row and spreads the outer samples over 32 / A lane groups, k interleaved (k = q, q + 32/A, ...): lanes
    // then agree on visibility almost always.  Wide or odd rows fall back to 32 consecutive texels per warp.
    const bool row_layout = A < 32 && (A & (A - 1)) == 0;
    const int kq = row_layout ? 32 / A : 1;
    const int ngroups = row_layout ? S : (ntex + 31) >> 5;
    for (int pass = 0; pass < passes; pass++) {
      const int group = pass * total_warps + ((pass & 1) ? total_warps - 1 - wg : wg);
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

// ------------------------------------------------------------------ K6: ray scatter from the dJ table

// Per-sample constants of the ray-scatter kernel, packed so that every thread fetches them with as few
// (broadcast) shared-memory instructions as possible -- the kernel is bound by shared-memory wavefronts.
struct alignas(16) LookupSmem {
  longlong2 row01[kMaxSteps];   // byte offsets of the (height, elevation) corner tiles (hu,eu), (hu,ev)
  longlong2 row23[kMaxSteps];   //                                                       (hv,eu), (hv,ev)
  float4 trw[kMaxSteps];        // T(x -> p_k) rgb, elevation weight es
  double2 nxy[kMaxSteps];       // p_k / |p_k|
  float hs[kMaxSteps];          // height weight
  double rk[kMaxSteps];         // |p_k| (exact fallback next to the clamp)
  double exp_table[kExpTabSize + 3];     // exp(i/64), i = -256 .. 0 (exp_tab)
};

// dS[i] = integral-ray over p_k of T(x, p_k) * dJ(p_k, v, l, above)   (atmosphere.clj:192-200 with
// point-scatter = the interpolation-table of dJ, interpolate.clj:101-104).
//
// All texels of the CTA look dJ up at the same height and elevation coordinates for a given outer
// sample k (they depend on p_k and v only), so the CTA first blends the four (height, elevation)
// corner tiles of dJ into one [light-elevation][heading] tile in shared memory (coalesced float4
// loads, double buffered), and each texel then interpolates inside that tile: 4 shared-memory loads
// per lookup instead of 16 scattered global ones.
__global__ void __launch_bounds__(1024) k_ray_scatter(Params P, Shard shard, const float4 *__restrict__ dj,
                                                      const double *__restrict__ exp_table, PeerOut out,
                                                      unsigned long long *counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ViewSmem &vs = *reinterpret_cast<ViewSmem *>(smem_raw);
  LookupSmem &ls = *reinterpret_cast<LookupSmem *>(smem_raw + sizeof(ViewSmem));
  float4 *tiles = reinterpret_cast<float4 *>(smem_raw + sizeof(ViewSmem) + sizeof(LookupSmem));
  const int H = P.shapes.s4[0], E = P.shapes.s4[1], S = P.shapes.s4[2], A = P.shapes.s4[3];
  const int he = shard.begin + blockIdx.x * shard.stride;
  const int h = he / E, e = he % E;
  const int steps = P.shapes.ray_steps;
  unsigned esamples = 0;
  fill_exp_tab(ls.exp_table, exp_table);
  setup_view_ray(P, h, e, vs, esamples);
  const ViewRay ray = vs.ray;
  const V3 v = v3(ray.vx, ray.vy, 0.0);
  // per outer sample: lookup coordinates that do not depend on the light direction
  for (int k = threadIdx.x; k < steps; k += blockDim.x) {
    V3 p = v3(vs.pkx[k], vs.pky[k], 0.0);
    Axis ah = axis_from(height_to_index(P.planet, H, p), H);
    Axis ae = axis_from(elevation_to_index(P.planet, E, p, v, ray.above != 0), E);
    const long long tile_bytes = (long long)S * A * sizeof(float4);
    ls.row01[k] = make_longlong2((ah.u * E + ae.u) * tile_bytes, (ah.u * E + ae.v) * tile_bytes);
    ls.row23[k] = make_longlong2((ah.v * E + ae.u) * tile_bytes, (ah.v * E + ae.v) * tile_bytes);
    ls.hs[k] = ah.s;
    float tr[3];
    transmittance_rgb(P.fast, vs.cv0[k], vs.cv1[k], tr);
    ls.trw[k] = make_float4(tr[0], tr[1], tr[2], ae.s);
    ls.rk[k] = sqrt(vs.rk2[k]);
    ls.nxy[k] = make_double2(vs.pkx[k] / ls.rk[k], vs.pky[k] / ls.rk[k]);
  }
  __syncthreads();
  const int ntex = S * A;
  const float a = (float)(ray.dlen / (double)steps);
  for (int chunk = 0; chunk < ntex; chunk += blockDim.x) {
    const int texel = chunk + threadIdx.x;
    const bool active = texel < ntex;
    const int si = active ? texel / A : 0, ai = active ? texel % A : 0;
    const double ss = index_to_sin_sun_elevation(S, (double)si);
    const V3 l = index_to_sun_direction(A, v, ss, (double)ai);
    const Axis aa = axis_from(sun_angle_to_index(A, v, l), A);
    float acc[3] = {0.f, 0.f, 0.f};
    // When the tile has at most one element per thread its four corner loads for sample k + 1 are issued
    // before the lookups of sample k (software pipelining: the L2 latency hides behind the FP64 coordinate
    // math instead of being exposed in front of every barrier).
    const bool one_per_thread = ntex <= (int)blockDim.x;
    const bool loader = (int)threadIdx.x < ntex;
    float4 c00 = make_float4(0.f, 0.f, 0.f, 0.f), c01 = c00, c10 = c00, c11 = c00;
    const char *dj_mine = reinterpret_cast<const char *>(dj + threadIdx.x);
    if (one_per_thread && loader) {
      const longlong2 r01 = ls.row01[0], r23 = ls.row23[0];
      c00 = ldg4(reinterpret_cast<const float4 *>(dj_mine + r01.x));
      c01 = ldg4(reinterpret_cast<const float4 *>(dj_mine + r01.y));
      c10 = ldg4(reinterpret_cast<const float4 *>(dj_mine + r23.x));
      c11 = ldg4(reinterpret_cast<const float4 *>(dj_mine + r23.y));
    }
    const double sun_scale = sun_elevation_scale(S), s_max = (double)(S - 1);
    for (int k = 0; k < steps; k++) {
      float4 *tile = tiles + (size_t)(k & 1) * ntex;
      const float4 trw = ls.trw[k];
      {
        const float es = trw.w, hs = ls.hs[k];
        if (one_per_thread) {
          if (loader) tile[threadIdx.x] = mix4(mix4(c00, c01, es), mix4(c10, c11, es), hs);
        } else {
          const char *base = reinterpret_cast<const char *>(dj);
          const longlong2 r01 = ls.row01[k], r23 = ls.row23[k];
          const float4 *t00 = reinterpret_cast<const float4 *>(base + r01.x);
          const float4 *t01 = reinterpret_cast<const float4 *>(base + r01.y);
          const float4 *t10 = reinterpret_cast<const float4 *>(base + r23.x);
          const float4 *t11 = reinterpret_cast<const float4 *>(base + r23.y);
          for (int idx = threadIdx.x; idx < ntex; idx += blockDim.x)
            tile[idx] = mix4(mix4(ldg4(t00 + idx), ldg4(t01 + idx), es), mix4(ldg4(t10 + idx), ldg4(t11 + idx), es), hs);
        }
      }
      __syncthreads();   // one barrier per sample: the other buffer was last read before the previous barrier
      if (one_per_thread && loader && k + 1 < steps) {
        const longlong2 r01 = ls.row01[k + 1], r23 = ls.row23[k + 1];
        c00 = ldg4(reinterpret_cast<const float4 *>(dj_mine + r01.x));
        c01 = ldg4(reinterpret_cast<const float4 *>(dj_mine + r01.y));
        c10 = ldg4(reinterpret_cast<const float4 *>(dj_mine + r23.x));
        c11 = ldg4(reinterpret_cast<const float4 *>(dj_mine + r23.y));
      }
      if (active) {
        // The sun-elevation coordinate stays in double: dJ falls by decades across the terminator, so a
        // float32 coordinate (about 1e-5 index units) shows up as 1e-3 relative error in dim texels.
        // sin = l . (p_k / |p_k|) with the unit vector precomputed per sample; only next to the lower clamp
        // (sin = -0.2, coordinate 0) it is recomputed as the reference writes it, (dot p l) / (mag p), so
        // that rows the reference clamps to exactly 0 are clamped here as well.
        const double2 n = ls.nxy[k];
        double sin_elev = l.x * n.x + l.y * n.y;
        if (sin_elev < -0.2 + 1e-9) sin_elev = (l.x * vs.pkx[k] + l.y * vs.pky[k]) / ls.rk[k];
        const Axis as = axis_from_nonneg(sun_elevation_coord(ls.exp_table, sun_scale, sin_elev), S, s_max);
        const float4 j = lookup2_smem(tile, A, as, aa);
        acc[0] = fmaf(trw.x, j.x, acc[0]);
        acc[1] = fmaf(trw.y, j.y, acc[1]);
        acc[2] = fmaf(trw.z, j.z, acc[2]);
      }
    }
    if (active) store_all(out, (size_t)he * ntex + texel, make_float4(acc[0] * a, acc[1] * a, acc[2] * a, 0.0f));
    __syncthreads();
  }
  count_esamples(counter, esamples);
}

// ------------------------------------------------------------------ K4: point scatter

// Per (height index, sphere direction): everything of in-scatter-from-direction
// (atmosphere.clj:208-222) that does not depend on the view or light direction.
__global__ void k_point_scatter_prepare(Params P, const double *__restrict__ dirs, int ndirs, DirInfo *info) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int H = P.shapes.s4[0];
  if (i >= H * ndirs) return;
  const int h = i / ndirs, d = i % ndirs;
  V3 x = index_to_height(P.planet, H, (double)h);
  V3 omega = v3(dirs[3 * d], dirs[3 * d + 1], dirs[3 * d + 2]);
  V3 point = ray_extremity(P.planet, x, omega);
  bool surface = surface_point(P.planet, point);
  DirInfo r;
  r.surface = surface ? 1 : 0;
  Axis ae = axis_from(elevation_to_index(P.planet, P.shapes.s4[1], x, omega, !surface), P.shapes.s4[1]);
  r.eu = ae.u;
  r.ev = ae.v;
  r.es = ae.s;
  r.tb[0] = r.tb[1] = r.tb[2] = 0.f;
  r.ehu = r.ehv = 0;
  r.ehs = 0.f;
  r.nx = r.ny = r.nz = 0.0;
  r.nmag = 1.0;
  if (surface) {
    double t[3];
    transmittance_points(P.planet, P.medium, P.shapes.ray_steps, x, point, t);
    for (int ch = 0; ch < 3; ch++) r.tb[ch] = (float)(t[ch] * (P.planet.brightness[ch] / kPi));
    Axis ah = axis_from(height_to_index(P.planet, P.shapes.se[0], point), P.shapes.se[0]);
    r.ehu = ah.u;
    r.ehv = ah.v;
    r.ehs = ah.s;
    r.nx = point.x;
    r.ny = point.y;
    r.nz = point.z;
    r.nmag = mag(point);
  }
  info[i] = r;
}

// S(x, omega_d, l, not surface) is looked up at height and elevation coordinates that depend on the
// height index h and the direction d only (atmosphere.clj:217: point x = (r_h, 0, 0), direction omega_d),
// never on the view or light direction.  Blend those two axes once per (h, d):
// tiles[(h * ndirs + d)][s][a] = mix_h(mix_e(tab)); the point-scatter kernel then interpolates the two
// remaining axes inside one 4 KB tile (4 loads per lookup instead of 16, shared by 127 CTAs).
__global__ void __launch_bounds__(256) k_blend_dir_tiles(Params P, const float4 *__restrict__ tab,
                                                         const DirInfo *__restrict__ info, int ndirs, int h_first,
                                                         float4 *tiles) {
  const int H = P.shapes.s4[0], E = P.shapes.s4[1];
  const int ntex = P.shapes.s4[2] * P.shapes.s4[3];
  const int hd = h_first * ndirs + blockIdx.x;
  const int h = hd / ndirs;
  const V3 x = index_to_height(P.planet, H, (double)h);
  const Axis ah = axis_from(height_to_index(P.planet, H, x), H);
  const DirInfo di = info[hd];
  const size_t r00 = ((size_t)ah.u * E + di.eu) * ntex, r01 = ((size_t)ah.u * E + di.ev) * ntex;
  const size_t r10 = ((size_t)ah.v * E + di.eu) * ntex, r11 = ((size_t)ah.v * E + di.ev) * ntex;
  float4 *tile = tiles + (size_t)hd * ntex;
  for (int idx = threadIdx.x; idx < ntex; idx += blockDim.x)
    tile[idx] = mix4(mix4(ldg4(tab + r00 + idx), ldg4(tab + r01 + idx), di.es),
                     mix4(ldg4(tab + r10 + idx), ldg4(tab + r11 + idx), di.es), ah.s);
}

struct alignas(16) PointDir {
  double ox, oy, oz;      // omega_d
  double px, py, pz, pm;  // ray extremity and its norm (surface directions)
  double ux, uy, uz;      // ray extremity / norm
  float sc[3];            // weight_d * sum_c scattering_c(h(x)) phase_c(v . omega_d)
  float tb[3];            // T(x -> point) * brightness / pi
  int surface;
  int ehu, ehv;
  float ehs;
};

// dJ[i] = integral-sphere of overall-in-scattering * (S(x, omega, l, not surface) + surface term).
// 64 registers (4 CTAs per SM): neutral on a full grid, but the 508 CTAs of an 8-GPU slab then fit in one wave
// (every thread's serial work is the same 71 directions, so smaller CTAs would not shorten the wave).
__global__ void __launch_bounds__(256, 4) k_point_scatter(Params P, Shard shard, const float4 *__restrict__ tiles_a,
                                                       const float4 *__restrict__ tiles_b, double phase_g,
                                                       const float4 *__restrict__ de,
                                                       const double *__restrict__ dirs,
                                                       const double *__restrict__ weights, int ndirs,
                                                       const DirInfo *__restrict__ info,
                                                       const double *__restrict__ exp_table, PeerOut out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  PointDir *pd = reinterpret_cast<PointDir *>(smem_raw);
  __shared__ double s_geom[4];
  __shared__ double s_exp[kExpTabSize];
  fill_exp_tab(s_exp, exp_table);
  const int H = P.shapes.s4[0], E = P.shapes.s4[1], S = P.shapes.s4[2], A = P.shapes.s4[3];
  const int he = shard.begin + blockIdx.x * shard.stride;
  const int h = he / E, e = he % E;
  if (threadIdx.x == 0) {
    V3 x = index_to_height(P.planet, H, (double)h);
    V3 v;
    bool above;
    index_to_elevation(P.planet, E, x.x, (double)e, v, above);
    s_geom[0] = x.x;
    s_geom[1] = v.x;
    s_geom[2] = v.y;
  }
  __syncthreads();
  const V3 x = v3(s_geom[0], 0.0, 0.0);
  const V3 v = v3(s_geom[1], s_geom[2], 0.0);
  const double hx = height(P.planet, x);
  for (int d = threadIdx.x; d < ndirs; d += blockDim.x) {
    V3 omega = v3(dirs[3 * d], dirs[3 * d + 1], dirs[3 * d + 2]);
    double mu = dot(v, omega);
    PointDir r;
    r.ox = omega.x;
    r.oy = omega.y;
    r.oz = omega.z;
    // overall-in-scattering (atmosphere.clj:147-151)
    for (int ch = 0; ch < 3; ch++) {
      double sum = 0.0;
      for (int c = 0; c < P.medium.n; c++) {
        double term = scattering(P.medium, c, ch, hx) * phase(P.medium.g[c], mu);
        sum = (c == 0) ? term : sum + term;
      }
      r.sc[ch] = (float)(sum * weights[d]);
    }
    const DirInfo di = info[(size_t)h * ndirs + d];
    r.surface = di.surface;
    r.tb[0] = di.tb[0];
    r.tb[1] = di.tb[1];
    r.tb[2] = di.tb[2];
    r.ehu = di.ehu;
    r.ehv = di.ehv;
    r.ehs = di.ehs;
    r.px = di.nx;
    r.py = di.ny;
    r.pz = di.nz;
    r.pm = di.nmag;
    r.ux = di.nx / di.nmag;
    r.uy = di.ny / di.nmag;
    r.uz = di.nz / di.nmag;
    pd[d] = r;
  }
  __syncthreads();
  const int ntex = S * A;
  const size_t tile_base = (size_t)h * ndirs * ntex;
  const float phase_c0 = (float)((3.0 * (1.0 - phase_g * phase_g)) / (8.0 * kPi * (2.0 + phase_g * phase_g)));
  const double e_scale = sun_elevation_scale(P.shapes.se[1]), e_max = (double)(P.shapes.se[1] - 1);
  const double a_half = 0.5 * (double)(A - 1), a_max = (double)(A - 1);
  for (int texel = threadIdx.x; texel < ntex; texel += blockDim.x) {
    const int si = texel / A, ai = texel % A;
    double ss = index_to_sin_sun_elevation(S, (double)si);
    V3 l = index_to_sun_direction(A, v, ss, (double)ai);
    const Axis as = axis_from(sun_elevation_to_index(S, x, l), S);
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll kPointScatterUnroll
    for (int d = 0; d < ndirs; d++) {
      const PointDir &r = pd[d];
      const double mu = r.ox * l.x + r.oy * l.y + r.oz * l.z;
      const double ca = a_half * (1 + mu);   // sun-angle-to-index (atmosphere.clj:368-372); continuous coordinate
      const Axis aa = axis_from_nonneg(ca < 0.0 ? 0.0 : ca, A, a_max);
      float4 s = lookup2(tiles_a + tile_base + (size_t)d * ntex, A, as, aa);
      if (tiles_b) {
        float4 m = lookup2(tiles_b + tile_base + (size_t)d * ntex, A, as, aa);
        // phase (atmosphere.clj:56-61) with the cancellation-prone base in double and the rest in float
        const float base = (float)((1.0 + phase_g * phase_g) - 2.0 * phase_g * mu);
        const float ph = phase_c0 * (float)(1.0 + mu * mu) / (base * sqrtf(base));
        s.x = fmaf(m.x, ph, s.x);
        s.y = fmaf(m.y, ph, s.y);
        s.z = fmaf(m.z, ph, s.z);
      }
      if (r.surface) {
        // surface-radiance (point, l): interpolation-table of dE over surface-radiance-space
        Axis eh;
        eh.u = r.ehu;
        eh.v = r.ehv;
        eh.s = r.ehs;
        // sine of the sun elevation at the ground point: unit vector precomputed per direction; next to the
        // lower clamp it is recomputed as the reference writes it, (dot point l) / (mag point)
        double sin_elev = r.ux * l.x + r.uy * l.y + r.uz * l.z;
        if (sin_elev < -0.2 + 1e-9) sin_elev = (r.px * l.x + r.py * l.y + r.pz * l.z) / r.pm;
        Axis es = axis_from_nonneg(sun_elevation_coord(s_exp, e_scale, sin_elev), P.shapes.se[1], e_max);
        float4 ev = lookup2(de, P.shapes.se[1], eh, es);
        s.x = fmaf(r.tb[0], ev.x, s.x);
        s.y = fmaf(r.tb[1], ev.y, s.y);
        s.z = fmaf(r.tb[2], ev.z, s.z);
      }
      acc[0] = fmaf(r.sc[0], s.x, acc[0]);
      acc[1] = fmaf(r.sc[1], s.y, acc[1]);
      acc[2] = fmaf(r.sc[2], s.z, acc[2]);
    }
    store_all(out, (size_t)he * ntex + texel, make_float4(acc[0], acc[1], acc[2], 0.0f));
  }
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
                                                          const HalfDirInfo *__restrict__ info, float4 *out) {
  __shared__ float red[3][8];
  const int i = blockIdx.x;
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
    out[i] = make_float4(r[0], r[1], r[2], 0.0f);
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
  // texels are taken pair by pair: pair shard.begin + n * shard.stride, all of its S*A texels
  const int ntex = S * A;
  const long long i = ((long long)shard.begin + (j / ntex) * shard.stride) * ntex + j % ntex;
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

// Thread q signals `epoch` into flag word [rank] of GPU q and then waits until GPU q has signalled this
// GPU.  The kernels whose stores must be visible ran earlier on the same stream; the system-scope fence
// orders them before the flag.  Epochs only grow, so a GPU that runs ahead never confuses a slower one.
// A peer that never arrives (crashed process) trips the clock-based timeout instead of hanging the box.
__global__ void k_peer_barrier(unsigned *local_flags, PeerFlags peers, int offset, int rank, int world, unsigned epoch,
                               int *error_flag) {
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
    if (clock64() - start > 20000000000LL) {   // about 10 s
      atomicExch(error_flag, 1);
      break;
    }
    __nanosleep(200);
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

cudaError_t launch_first_order(const Params &P, Shard shard, int he_count, FirstOrderOut oa, FirstOrderOut ob,
                               unsigned long long *counter, cudaStream_t st) {
  if (he_count <= 0) return cudaSuccess;
  // tuning knobs (defaults chosen on B200, see profiles/): warps per CTA and texel groups per warp
  static const int warps = std::min(8, std::max(1, env_int("ATMLUT_FIRST_ORDER_WARPS", 8)));
  static const int want_passes = std::max(1, env_int("ATMLUT_FIRST_ORDER_PASSES", 4));
  const int threads = warps * 32;
  const int ntex = P.shapes.s4[2] * P.shapes.s4[3];
  int kparts = 1, passes = 1, nchunks = 1;
  if (ntex < threads) {
    kparts = std::min(P.shapes.ray_steps, threads / ntex);
  } else {
    const int A = P.shapes.s4[3];
    const bool row_layout = A < 32 && (A & (A - 1)) == 0;     // must match the kernel
    const int ngroups = row_layout ? P.shapes.s4[2] : (ntex + 31) / 32;
    passes = std::min(want_passes, (ngroups + warps - 1) / warps);
    // Small grids (one slab of a multi-GPU build, small tables): fewer row groups per warp, i.e. more and smaller
    // CTAs per pair, until the launch spans several waves -- the repeated view-ray set-up (steps^2 samples per
    // CTA) is cheaper than SMs idling behind the slowest CTA of a single wave.
    const int min_ctas = small_grid_ctas();
    for (;;) {
      const int total_warps = (ngroups + passes - 1) / passes;
      nchunks = (total_warps + warps - 1) / warps;
      if ((long long)he_count * nchunks >= min_ctas || passes == 1) break;
      passes = (passes + 1) / 2;
    }
  }
  size_t smem = sizeof(ViewSmem) + (kparts > 1 ? 6 * threads * sizeof(float) : 0);
  k_first_order<<<he_count * nchunks, threads, smem, st>>>(P, shard, kparts, passes, nchunks, oa, ob, counter);
  return cudaGetLastError();
}

static int ray_scatter_threads(const Params &P) {
  int ntex = P.shapes.s4[2] * P.shapes.s4[3];
  int t = (ntex + 31) / 32 * 32;
  return t < 128 ? 128 : (t > 1024 ? 1024 : t);
}

size_t ray_scatter_smem(const Params &P) {
  return sizeof(ViewSmem) + sizeof(LookupSmem) + 2 * (size_t)P.shapes.s4[2] * P.shapes.s4[3] * sizeof(float4);
}

cudaError_t launch_ray_scatter(const Params &P, Shard shard, int he_count, const float4 *dj, const double *exp_table,
                               PeerOut out, unsigned long long *counter, cudaStream_t st) {
  if (he_count <= 0) return cudaSuccess;
  size_t smem = ray_scatter_smem(P);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;   // light-elevation x heading tile too large
  cudaError_t e = cudaFuncSetAttribute(k_ray_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_ray_scatter<<<he_count, ray_scatter_threads(P), smem, st>>>(P, shard, dj, exp_table, out, counter);
  return cudaGetLastError();
}

cudaError_t launch_point_scatter_prepare(const Params &P, const double *dirs, int ndirs, DirInfo *info,
                                         cudaStream_t st) {
  int n = P.shapes.s4[0] * ndirs;
  k_point_scatter_prepare<<<div_up(n, 64), 64, 0, st>>>(P, dirs, ndirs, info);
  return cudaGetLastError();
}

cudaError_t launch_blend_dir_tiles(const Params &P, const float4 *tab, const DirInfo *info, int ndirs, int h_first,
                                   int h_count, float4 *tiles, cudaStream_t st) {
  if (h_count <= 0) return cudaSuccess;
  k_blend_dir_tiles<<<h_count * ndirs, 256, 0, st>>>(P, tab, info, ndirs, h_first, tiles);
  return cudaGetLastError();
}

cudaError_t launch_point_scatter(const Params &P, Shard shard, int he_count, const float4 *tiles_a,
                                 const float4 *tiles_b, double phase_g, const float4 *de, const double *dirs,
                                 const double *weights, int ndirs, const DirInfo *info, const double *exp_table,
                                 PeerOut out, cudaStream_t st) {
  if (he_count <= 0) return cudaSuccess;
  size_t smem = (size_t)ndirs * sizeof(PointDir);
  cudaError_t e = cudaFuncSetAttribute(k_point_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_point_scatter<<<he_count, 256, smem, st>>>(P, shard, tiles_a, tiles_b, phase_g, de, dirs, weights, ndirs, info,
                                               exp_table, out);
  return cudaGetLastError();
}

cudaError_t launch_surface_radiance_prepare(const Params &P, const double *dirs, int ndirs, HalfDirInfo *info,
                                            cudaStream_t st) {
  int n = P.shapes.se[0] * ndirs;
  k_surface_radiance_prepare<<<div_up(n, 128), 128, 0, st>>>(P, dirs, ndirs, info);
  return cudaGetLastError();
}

cudaError_t launch_surface_radiance(const Params &P, SSource src, const double *dirs, const double *weights,
                                    int ndirs, const HalfDirInfo *info, float4 *out, cudaStream_t st) {
  int n = P.shapes.se[0] * P.shapes.se[1];
  k_surface_radiance<<<n, 256, 0, st>>>(P, src, dirs, weights, ndirs, info, out);
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
                                unsigned epoch, int *error_flag, cudaStream_t st) {
  PeerFlags f = {};
  for (int q = 0; q < world && q < kMaxPeers; q++) f.p[q] = peer_flags[q];
  k_peer_barrier<<<1, 32, 0, st>>>(local_flags, f, flag_set * kMaxPeers, rank, world, epoch, error_flag);
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
