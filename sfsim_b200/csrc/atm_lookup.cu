// Lookup kernels of the atmosphere-LUT build for sm_100a: the two integrals whose integrand is a 4-D table lookup.
//
//   K4 k_point_scatter   point-scatter (dJ): sphere quadrature over the previous order's dS   (atmosphere.clj:203-222)
//   K6 k_ray_scatter     ray-scatter from the dJ table (dS)                                   (atmosphere.clj:192-200)
//
// Both are bound by shared-memory / L1 wavefronts and issue slots, not by arithmetic: a 4-D multilinear lookup
// (interpolate.clj:87-98) is 16 texel reads.  Two of the four coordinates are shared by all texels of a CTA, so
// those two axes are blended ONCE into a [light-elevation][heading] tile (k_blend_dir_tiles for K4, in the kernel
// per outer sample for K6) and every texel then reads 4 entries of that tile from shared memory.
#include <algorithm>
#include <cstdlib>

#include "atm_kernel_common.cuh"

namespace atm {

// sines of the sun elevation this close to the lower end of the table (-0.2) are recomputed exactly as the
// reference writes them, so that rows the reference clamps to exactly 0 are clamped here as well
__device__ constexpr double kLowGuard = -0.2 + 1e-9;

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
__global__ void __launch_bounds__(1024) k_ray_scatter_v1(Params P, Shard shard, const float4 *__restrict__ dj,
                                                      const double *__restrict__ exp_table, PeerOut out,
                                                      unsigned long long *counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ViewSmem &vs = *reinterpret_cast<ViewSmem *>(smem_raw);
  LookupSmem &ls = *reinterpret_cast<LookupSmem *>(smem_raw + sizeof(ViewSmem));
  float4 *tiles = reinterpret_cast<float4 *>(smem_raw + sizeof(ViewSmem) + sizeof(LookupSmem));
  const int H = P.shapes.s4[0], E = P.shapes.s4[1], S = P.shapes.s4[2], A = P.shapes.s4[3];
  const int he = shard_pair(shard, blockIdx.x);
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
                                                         int h_stride, float4 *tiles) {
  const int H = P.shapes.s4[0], E = P.shapes.s4[1];
  const int ntex = P.shapes.s4[2] * P.shapes.s4[3];
  const int h = h_first + (blockIdx.x / ndirs) * h_stride;
  const int hd = h * ndirs + blockIdx.x % ndirs;
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
__global__ void __launch_bounds__(256, 4) k_point_scatter_v1(Params P, Shard shard, const float4 *__restrict__ tiles_a,
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
  const int he = shard_pair(shard, blockIdx.x);
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


// ================================================================== bulk copies (TMA unit) and mbarriers

// ---- bulk copies by the TMA unit, completion on an mbarrier

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// global -> shared copy of `bytes` (multiple of 16, both addresses 16-byte aligned) issued by ONE thread
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// a copy that never lands would hang the box: trap instead (seconds of polling)
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 28)) __trap();
}


// coefficients of exp(r) on |r| <= 1/128 (see exp_tab); operands from the constant bank instead of immediates that
// the compiler rebuilds with uniform moves inside the loops
__constant__ double kExpPoly[6] = {1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0};

// scale * (1 - exp(y)) from a table that already holds scale * exp(i/64); y in [-4, 2.5]; <= 0 for y >= 0
__device__ __forceinline__ double coord_from_exponent(const double *scaled_tab, double scale, double y) {
  const double magic = 6755399441055744.0;            // 1.5 * 2^52: the low word of y*64 + magic is rint(y*64)
  const double t = fma(y, 64.0, magic);
  const int i = __double2loint(t);
  const double r = fma(t - magic, -1.0 / 64.0, y);    // y - i/64
  double p = fma(r, kExpPoly[0], kExpPoly[1]);
  p = fma(p, r, kExpPoly[2]);
  p = fma(p, r, kExpPoly[3]);
  p = fma(p, r, kExpPoly[4]);
  p = fma(p, r, kExpPoly[5]);
  p = fma(p, r, kExpPoly[5]);
  return fma(-scaled_tab[i - kExpTabLo], p, scale);
}

// ================================================================== K6, current version

constexpr unsigned kOffMask = (1u << 28) - 1;   // corner-tile offsets (float4 units) fit 28 bits: the table has <= 2^28 texels

// Everything the ray-scatter kernel needs per outer sample p_k that does not depend on the light direction.  The
// records depend on the (height, elevation) pair only -- not on the scattering order -- so k_ray_prepare writes
// them once per build and every ray-scatter pass fetches its pair's records with one bulk copy.
struct alignas(16) RaySample {
  uint4 rows;       // float4 offsets of the four (height, elevation) corner tiles, in REGISTER-SLOT order
  float4 cw;        // bilinear weight of the corner held in each slot
  float4 tr;        // T(x -> p_k) rgb * (ray length / steps); .w: bit mask of the slots the NEXT sample reloads
  double2 n3;       // -3 p_k / |p_k|: y = n3 . l - 0.6 is the exponent of sun-elevation-to-index
  double2 pk;       // p_k and ...
  double rk;        // ... |p_k| for the exact evaluation next to the lower clamp
  double unused;
};

// One CTA per (height, elevation) pair: the view ray (setup_view_ray) and, per outer sample, the corner tiles and
// weights of the dJ lookup's height and elevation axes, T(x -> p_k), and the unit vector of p_k.
//
// Corner tiles live in four register slots of the ray-scatter kernel.  While the sample stays inside one
// (height, elevation) cell nothing is loaded; a step into a neighbouring cell loads the two new corners and the two
// surviving ones keep their slots and change ROLE (a height step swaps the roles 00<->10, 01<->11, an elevation
// step 00<->01, 10<->11; the swaps commute, so the slot of role r at sample k is r xor the parity of the steps so
// far).  Offsets and weights are therefore stored in slot order, with a mask of the slots to load.
__global__ void __launch_bounds__(256) k_ray_prepare(Params P, Shard shard, RaySample *__restrict__ samples,
                                                     unsigned long long *counter, const unsigned char *view_packs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned char s_code[kMaxSteps];
  __shared__ unsigned char s_mask[kMaxSteps];
  ViewSmem &vs = *reinterpret_cast<ViewSmem *>(smem_raw);
  RaySample *sample = reinterpret_cast<RaySample *>(smem_raw + sizeof(ViewSmem));
  const int steps = P.shapes.ray_steps;
  const int H = P.shapes.s4[0], E = P.shapes.s4[1];
  const int ntex = P.shapes.s4[2] * P.shapes.s4[3];
  const int he = shard_pair(shard, blockIdx.x);
  const int h = he / E, e = he % E;
  const int tid = threadIdx.x;
  unsigned esamples = 0;
  if (view_packs)
    load_view_ray(vs, steps, view_packs + (size_t)he * view_pack_bytes(steps));
  else
    setup_view_ray(P, h, e, vs, esamples);
  const ViewRay ray = vs.ray;
  const V3 v = v3(ray.vx, ray.vy, 0.0);
  const float a = (float)(ray.dlen / (double)steps);      // ray.clj:19-30: stepsize * |direction|
  // ---- role order first
  for (int k = tid; k < steps; k += blockDim.x) {
    const V3 p = v3(vs.pkx[k], vs.pky[k], 0.0);
    const Axis ah = axis_from(height_to_index(P.planet, H, p), H);
    const Axis ae = axis_from(elevation_to_index(P.planet, E, p, v, ray.above != 0), E);
    RaySample &r = sample[k];
    r.rows = make_uint4((unsigned)((ah.u * E + ae.u) * ntex), (unsigned)((ah.u * E + ae.v) * ntex),
                        (unsigned)((ah.v * E + ae.u) * ntex), (unsigned)((ah.v * E + ae.v) * ntex));
    const float wh1 = ah.s, wh0 = 1.0f - ah.s, we1 = ae.s, we0 = 1.0f - ae.s;
    r.cw = make_float4(wh0 * we0, wh0 * we1, wh1 * we0, wh1 * we1);
    float tr[3];
    transmittance_rgb(P.fast, vs.cv0[k], vs.cv1[k], tr);
    r.tr = make_float4(tr[0] * a, tr[1] * a, tr[2] * a, 0.0f);
    const double rk = sqrt(vs.rk2[k]);
    r.rk = rk;
    r.pk = make_double2(vs.pkx[k], vs.pky[k]);
    r.n3 = make_double2(-3.0 * (vs.pkx[k] / rk), -3.0 * (vs.pky[k] / rk));
    r.unused = 0.0;
  }
  __syncthreads();
  // ---- how the corner tiles of sample k follow from those of sample k - 1
  for (int k = tid; k < steps; k += blockDim.x) {
    int code = 5;                                           // all four are new
    if (k > 0) {
      const uint4 n = sample[k].rows, o = sample[k - 1].rows;
      if (n.x == o.x && n.y == o.y && n.z == o.z && n.w == o.w) code = 0;       // same cell
      else if (n.x == o.z && n.y == o.w) code = 1;          // one height row up: roles 10, 11 are new
      else if (n.z == o.x && n.w == o.y) code = 2;          // one height row down: roles 00, 01 are new
      else if (n.x == o.y && n.z == o.w) code = 3;          // one elevation column up: roles 01, 11 are new
      else if (n.y == o.x && n.w == o.z) code = 4;          // one elevation column down: roles 00, 10 are new
    }
    s_code[k] = (unsigned char)code;
  }
  __syncthreads();
  // ---- role order -> slot order
  for (int k = tid; k < steps; k += blockDim.x) {
    int perm = 0;                                           // slot of role r at sample k = r ^ perm
    for (int j = 1; j <= k; j++) {
      const int c = s_code[j];
      perm ^= (c == 1 || c == 2) ? 2 : ((c == 3 || c == 4) ? 1 : 0);
    }
    const int c = s_code[k];
    const unsigned new_roles = c == 0 ? 0u : c == 1 ? 0xcu : c == 2 ? 0x3u : c == 3 ? 0xau : c == 4 ? 0x5u : 0xfu;
    RaySample &r = sample[k];
    const uint4 rows = r.rows;
    const float4 cw = r.cw;
    const unsigned ro[4] = {rows.x, rows.y, rows.z, rows.w};
    const float wr[4] = {cw.x, cw.y, cw.z, cw.w};
    unsigned so[4], mask = 0;
    float sw[4];
#pragma unroll
    for (int slot = 0; slot < 4; slot++) {
      const int role = slot ^ perm;
      so[slot] = ro[role];
      sw[slot] = wr[role];
      mask |= ((new_roles >> role) & 1u) << slot;
    }
    r.rows = make_uint4(so[0], so[1], so[2], so[3]);
    r.cw = make_float4(sw[0], sw[1], sw[2], sw[3]);
    s_mask[k] = (unsigned char)mask;
  }
  __syncthreads();
  for (int k = tid; k < steps; k += blockDim.x) {
    sample[k].tr.w = __uint_as_float(k + 1 < steps ? (unsigned)s_mask[k + 1] : 0u);
  }
  __syncthreads();
  // ---- out: 6 x 16 bytes per sample, coalesced
  const uint4 *src = reinterpret_cast<const uint4 *>(sample);
  uint4 *dst = reinterpret_cast<uint4 *>(samples + (size_t)blockIdx.x * steps);
  const int n16 = steps * (int)(sizeof(RaySample) / 16);
  for (int i = tid; i < n16; i += blockDim.x) dst[i] = src[i];
  count_esamples(counter, esamples);
}

// dS[i] = integral-ray over p_k of T(x, p_k) * dJ(p_k, v, l, above)   (atmosphere.clj:192-200 with point-scatter =
// the interpolation-table of dJ, interpolate.clj:101-104).  One CTA per (height, elevation) pair, one thread per
// (light-elevation, heading) texel; the pair's RaySample records arrive by one bulk copy.
//
// Per outer sample k the CTA blends the four (height, elevation) corner tiles of dJ, times T(x -> p_k), into one
// tile in shared memory (double buffered, one barrier per sample); each texel then interpolates the remaining two
// axes there.  The kernel is bound by L1/shared wavefronts and issue slots:
//  * corner tiles stay in registers across samples (see k_ray_prepare); a cell change loads two new corners;
//  * per-sample constants are 4 broadcast LDS.128 (a 128-bit broadcast costs two wavefronts), the offsets of the
//    next sample are only fetched when its reload mask is not empty;
//  * tile rows are NOT padded: the eight threads of a 128-bit shared-memory wavefront are the eight headings of one
//    light-elevation row and read eight different columns, which unpadded rows of 8 x 16 bytes map to different
//    banks whatever row each thread reads (measured: 4.97 wavefronts per LDS.128 against 5.46 with a pad entry;
//    the excess over 4 comes from texels whose clamped headings share a column);
//  * the exponent y and the scaled exponential come from one fused multiply-add each (table pre-multiplied by the
//    coordinate scale, polynomial coefficients from the constant bank), floor / fraction from the 1.5 * 2^52
//    rounding constant instead of conversion instructions.
//  * kSplitTile: the tile is kept as a float2 (red, green) plane and a float (blue) plane: a corner read is an LDS.64
//    and an LDS.32 = 3 wavefronts instead of the 4 of an LDS.128 whose fourth lane is padding.
template <bool kSplitTile>
__global__ void __launch_bounds__(1024) k_ray_scatter(Params P, Shard shard, const RaySample *__restrict__ samples,
                                                      const float4 *__restrict__ dj,
                                                      const double *__restrict__ exp_table, PeerOut out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int steps = P.shapes.ray_steps;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw);
  RaySample *sample = reinterpret_cast<RaySample *>(smem_raw + 16);
  double *s_exp = reinterpret_cast<double *>(sample + steps);                     // exp(i/64) * coordinate scale
  float4 *tiles = reinterpret_cast<float4 *>(s_exp + ((kExpTabSize + 1) & ~1));
  const int E = P.shapes.s4[1], S = P.shapes.s4[2], A = P.shapes.s4[3];
  const int ntex = S * A;
  const int row_pitch = A;                                // entries per tile row (see below: no padding)
  const int tile_entries = S * row_pitch;
  float2 *tiles_rg = reinterpret_cast<float2 *>(tiles);   // kSplitTile: [2][entries] float2, then [2][entries] float
  float *tiles_b = reinterpret_cast<float *>(tiles_rg + 2 * tile_entries);
  const int he = shard_pair(shard, blockIdx.x);
  const int h = he / E, e = he % E;
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const unsigned bytes = (unsigned)steps * (unsigned)sizeof(RaySample);
    mbar_expect_tx(full, bytes);
    bulk_copy_g2s(sample, samples + (size_t)blockIdx.x * steps, bytes, full);
  }
  const double sun_scale = sun_elevation_scale(S);
  for (int i = tid; i < kExpTabSize; i += blockDim.x) s_exp[i] = exp_table[i] * sun_scale;
  // this thread's texel: light direction (ray-scatter-backward, atmosphere.clj:401-412) and its heading lookup axis
  const bool active = tid < ntex;
  const int si = active ? tid / A : 0, ai = active ? tid % A : 0;
  V3 x = index_to_height(P.planet, P.shapes.s4[0], (double)h);
  V3 v;
  bool above;
  index_to_elevation(P.planet, E, x.x, (double)e, v, above);
  const double ss = index_to_sin_sun_elevation(S, (double)si);
  const V3 l = index_to_sun_direction(A, v, ss, (double)ai);
  const Axis aa = axis_from(sun_angle_to_index(A, v, l), A);
  const float wa1 = aa.s, wa0 = 1.0f - aa.s;
  const double lx = l.x, ly = l.y;
  const int s_last = S - 1;
  const unsigned utid = (unsigned)tid;   // corner offset + texel stays below 2^29: one 32-bit add, one wide multiply-add
  const int my_entry = si * row_pitch + ai;
  __syncthreads();                       // the mbarrier is initialised, the exponential table is filled
  mbar_wait(full, 0);
  float2 acc01 = make_float2(0.f, 0.f);
  float acc2 = 0.f;
  float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0, c2 = c0, c3 = c0;
  if (active) {
    const uint4 r0 = sample[0].rows;
    c0 = ldg4(dj + (r0.x + utid));
    c1 = ldg4(dj + (r0.y + utid));
    c2 = ldg4(dj + (r0.z + utid));
    c3 = ldg4(dj + (r0.w + utid));
  }
#pragma unroll 2
  for (int k = 0; k < steps; k++) {
    float4 *tile = tiles + (size_t)(k & 1) * tile_entries;
    const RaySample &r = sample[k];
    unsigned next_mask = 0;
    if (active) {
      const float4 cw = r.cw, tr = r.tr;
      next_mask = __float_as_uint(tr.w);
      // red and green in one packed instruction per term (FFMA2: the same two IEEE operations, one issue slot)
      float2 brg = __fmul2_rn(make_float2(cw.x, cw.x), make_float2(c0.x, c0.y));
      brg = __ffma2_rn(make_float2(cw.y, cw.y), make_float2(c1.x, c1.y), brg);
      brg = __ffma2_rn(make_float2(cw.z, cw.z), make_float2(c2.x, c2.y), brg);
      brg = __ffma2_rn(make_float2(cw.w, cw.w), make_float2(c3.x, c3.y), brg);
      brg = __fmul2_rn(make_float2(tr.x, tr.y), brg);
      const float bb = tr.z * fmaf(cw.w, c3.z, fmaf(cw.z, c2.z, fmaf(cw.y, c1.z, cw.x * c0.z)));
      if (kSplitTile) {
        tiles_rg[(size_t)(k & 1) * tile_entries + my_entry] = brg;
        tiles_b[(size_t)(k & 1) * tile_entries + my_entry] = bb;
      } else {
        tile[my_entry] = make_float4(brg.x, brg.y, bb, 0.0f);
      }
    }
    __syncthreads();   // one barrier per sample: the other buffer was last read before the previous barrier
    if (active) {
      if (next_mask) {
        // corner tiles of the next sample; the loads fly while this sample's lookup is computed
        const uint4 rn = sample[k + 1].rows;
        if (next_mask & 1u) c0 = ldg4(dj + (rn.x + utid));
        if (next_mask & 2u) c1 = ldg4(dj + (rn.y + utid));
        if (next_mask & 4u) c2 = ldg4(dj + (rn.z + utid));
        if (next_mask & 8u) c3 = ldg4(dj + (rn.w + utid));
      }
      // The sun-elevation coordinate stays in double: dJ falls by decades across the terminator, so a float32
      // coordinate (about 1e-5 index units) shows up as 1e-3 relative error in dim texels.
      // y = -3 sin - 0.6 with sin = l . (p_k / |p_k|) (atmosphere.clj:322-326); where the coordinate is about to be
      // clamped to 0 (sin = -0.2, y = 0) it is recomputed as the reference writes it, from (dot p l) / (mag p), so
      // that samples the reference clamps to exactly 0 are clamped here as well.
      const double2 n3 = r.n3;
      double y = fma(lx, n3.x, fma(ly, n3.y, -0.6));
      if (fabs(y) < 3e-9) y = (0 - 3 * ((lx * r.pk.x + ly * r.pk.y) / r.rk)) - 0.6;
      // coordinate = scale (1 - exp(y)); <= 0 where the reference's max(0, .) acts
      FloorFrac fs = floor_frac(coord_from_exponent(s_exp, sun_scale, y));
      if (fs.u < 0) {
        fs.u = 0;
        fs.s = 0.0f;
      }
      const int sv = min(fs.u + 1, s_last);
      float2 a00, a01, a10, a11;     // red, green of the four corners
      float z00, z01, z10, z11;      // blue
      if (kSplitTile) {
        const float2 *g0 = tiles_rg + (size_t)(k & 1) * tile_entries + fs.u * row_pitch, *g1 = g0 + (sv - fs.u) * row_pitch;
        const float *b0 = tiles_b + (size_t)(k & 1) * tile_entries + fs.u * row_pitch, *b1 = b0 + (sv - fs.u) * row_pitch;
        a00 = g0[aa.u], a01 = g0[aa.v], a10 = g1[aa.u], a11 = g1[aa.v];
        z00 = b0[aa.u], z01 = b0[aa.v], z10 = b1[aa.u], z11 = b1[aa.v];
      } else {
        const float4 *r0 = tile + fs.u * row_pitch, *r1 = tile + sv * row_pitch;
        const float4 v00 = r0[aa.u], v01 = r0[aa.v], v10 = r1[aa.u], v11 = r1[aa.v];
        a00 = make_float2(v00.x, v00.y), a01 = make_float2(v01.x, v01.y), a10 = make_float2(v10.x, v10.y);
        a11 = make_float2(v11.x, v11.y);
        z00 = v00.z, z01 = v01.z, z10 = v10.z, z11 = v11.z;
      }
      const float ws1 = fs.s, ws0 = 1.0f - fs.s;
      const float w00 = ws0 * wa0, w01 = ws0 * wa1, w10 = ws1 * wa0, w11 = ws1 * wa1;
      acc01 = __ffma2_rn(make_float2(w00, w00), a00, acc01);
      acc01 = __ffma2_rn(make_float2(w01, w01), a01, acc01);
      acc01 = __ffma2_rn(make_float2(w10, w10), a10, acc01);
      acc01 = __ffma2_rn(make_float2(w11, w11), a11, acc01);
      acc2 = fmaf(w11, z11, fmaf(w10, z10, fmaf(w01, z01, fmaf(w00, z00, acc2))));
    }
  }
  if (active) store_all(out, (size_t)he * ntex + tid, make_float4(acc01.x, acc01.y, acc2, 0.0f));
}

// ================================================================== K4, current version

constexpr int kTileStages = 6;           // ring of direction tiles per CTA
constexpr int kPointScatterThreads = 256;                         // consumer threads = texels per CTA (at most) ...
constexpr int kPointScatterBlock = kPointScatterThreads + 32;     // ... plus one producer warp

struct alignas(16) PointDir2 {
  double ox, oy;            // omega_d
  double oz, q1;            // q1 = -3 |x| / |point|
  double q2, pm;            // q2 = -3 |point - x| / |point|: -3 sin(sun elevation at point) = l.x q1 + (omega . l) q2
  double px, py;            // the ray extremity `point` (exact evaluation next to the lower clamp)
  double pz, unused;
  float sc[3];              // weight_d * sum_c scattering_c(h(x)) phase_c(v . omega_d)
  int surface;
  float tb[3];              // T(x -> point) * brightness / pi
  float ehs;                // height axis of E(point, l): weight and corners
  int ehu, ehv, pad[2];
};

__device__ __forceinline__ float4 blend4(float4 v00, float4 v01, float4 v10, float4 v11, float w00, float w01, float w10,
                                         float w11) {
  return make_float4(fmaf(w11, v11.x, fmaf(w10, v10.x, fmaf(w01, v01.x, w00 * v00.x))),
                     fmaf(w11, v11.y, fmaf(w10, v10.y, fmaf(w01, v01.y, w00 * v00.y))),
                     fmaf(w11, v11.z, fmaf(w10, v10.z, fmaf(w01, v01.z, w00 * v00.z))), 0.0f);
}

// dJ[i] = integral-sphere of overall-in-scattering * (S(x, omega, l, not surface) + surface term)
// (atmosphere.clj:203-222).  One CTA per (height, elevation) pair and chunk of 256 texels, one thread per texel,
// the directions in sequence.  Every thread of the CTA reads the SAME pre-blended [light-elevation][heading] tile
// for a direction (k_blend_dir_tiles), so the rows of that tile the chunk can touch are streamed into a ring of
// shared-memory buffers by the TMA unit (cp.async.bulk, one instruction per tile, completion on a "full" mbarrier):
// a lookup is 4 shared-memory reads with no exposed L2 latency.  A ninth warp is the producer; the eight consumer
// warps hand buffers back through "empty" mbarriers, so no warp ever waits for another one's arithmetic -- the
// kernel is latency-bound (double-precision coordinate chains at 24 warps per SM), CTA-wide barriers per direction
// cost more than anything else.
// (A variant without the producer warp -- thread 0 requesting four directions ahead, 64 registers, four CTAs per SM
// so that the 508 CTAs of an 8-GPU shard are resident at once -- was measured slower everywhere: 3.3 ms against
// 2.7 ms per build on one GPU, 0.73 ms against 0.67 ms on eight.)
template <bool kTwoTables, bool kDeShared>
__global__ void __launch_bounds__(kPointScatterBlock, 3)
    k_point_scatter(Params P, Shard shard, int chunks, int rows_max, const float4 *__restrict__ tiles_a,
                    const float4 *__restrict__ tiles_b, double phase_g, const float4 *__restrict__ de,
                    const double *__restrict__ dirs, const double *__restrict__ weights, int ndirs,
                    const DirInfo *__restrict__ info, const double *__restrict__ exp_table, PeerOut out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double s_geom[4];
  __shared__ int s_rows[2];
  const int H = P.shapes.s4[0], E = P.shapes.s4[1], S = P.shapes.s4[2], A = P.shapes.s4[3];
  const int Eh = P.shapes.se[0], Es = P.shapes.se[1];
  const int ntex = S * A;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw);   // [kTileStages]
  unsigned long long *empty = full + kTileStages;                                 // [kTileStages]
  PointDir2 *pd = reinterpret_cast<PointDir2 *>(smem_raw + 128);
  double *s_exp = reinterpret_cast<double *>(pd + ndirs);
  float4 *s_de = reinterpret_cast<float4 *>(s_exp + ((kExpTabSize + 1) & ~1));
  float4 *stages = s_de + (kDeShared ? Eh * Es : 0);
  const int stage_texels = rows_max * A * (kTwoTables ? 2 : 1);
  const int he = shard_pair(shard, blockIdx.x / chunks);
  const int chunk = blockIdx.x % chunks;
  const int h = he / E, e = he % E;
  const int tid = threadIdx.x;
  const int consumers = kPointScatterThreads;             // texels per CTA
  const bool producer = tid >= consumers;
  if (tid == 0) {
    V3 x = index_to_height(P.planet, H, (double)h);
    V3 v;
    bool above;
    index_to_elevation(P.planet, E, x.x, (double)e, v, above);
    s_geom[0] = x.x;
    s_geom[1] = v.x;
    s_geom[2] = v.y;
    s_rows[0] = S;
    s_rows[1] = -1;
    for (int s = 0; s < kTileStages; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], consumers);          // every consumer thread releases a buffer itself
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const V3 x = v3(s_geom[0], 0.0, 0.0);
  const V3 v = v3(s_geom[1], s_geom[2], 0.0);
  // this thread's texel: light direction and the sun-elevation rows of its lookups (the same for every direction)
  const int texel = chunk * consumers + tid;
  const bool active = !producer && texel < ntex;
  const int si = active ? texel / A : 0, ai = active ? texel % A : 0;
  const double ss = index_to_sin_sun_elevation(S, (double)si);
  const V3 l = index_to_sun_direction(A, v, ss, (double)ai);
  const Axis as = axis_from(sun_elevation_to_index(S, x, l), S);
  if (active) {
    atomicMin(&s_rows[0], as.u);
    atomicMax(&s_rows[1], as.v);
  }
  __syncthreads();
  const int row_lo = s_rows[0];
  const int nrows = min(s_rows[1] - row_lo + 1, rows_max);   // rows_max bounds it by construction (see the launcher)

  // ---- the tiles stream into ring position d % kTileStages; the first requests go out before the per-direction
  // constants below are worked out, so that their latency is not exposed at the head of the direction loop
  const unsigned tile_bytes = (unsigned)(nrows * A) * (unsigned)sizeof(float4);
  const size_t tile_first = (size_t)h * ndirs * ntex + (size_t)row_lo * A;   // float4 offset of direction 0's rows
  auto request = [&](int d, int stage) {
    const int use = d / kTileStages;
    if (use > 0) mbar_wait(&empty[stage], (unsigned)(use - 1) & 1u);             // every consumer warp has released it
    float4 *dst = stages + (size_t)stage * stage_texels;
    mbar_expect_tx(&full[stage], kTwoTables ? 2 * tile_bytes : tile_bytes);
    bulk_copy_g2s(dst, tiles_a + tile_first + (size_t)d * ntex, tile_bytes, &full[stage]);
    if (kTwoTables) bulk_copy_g2s(dst + nrows * A, tiles_b + tile_first + (size_t)d * ntex, tile_bytes, &full[stage]);
  };
  if (producer) {
    // producer warp: one lane runs up to kTileStages directions ahead of the slowest consumer warp
    if (tid == consumers)
      for (int d = 0; d < ndirs; d++) request(d, d % kTileStages);
    return;
  }
  // ---- per-direction constants, the exponential table and dE (consumer threads only from here on)
  const double e_scale = sun_elevation_scale(Es);
  for (int i = tid; i < kExpTabSize; i += consumers) s_exp[i] = exp_table[i] * e_scale;   // see coord_from_exponent
  if (kDeShared)
    for (int i = tid; i < Eh * Es; i += consumers) s_de[i] = ldg4(de + i);
  const double hx = height(P.planet, x);
  // scattering (atmosphere.clj:42-47) = base * density: the density does not depend on the direction or the channel
  double density[2] = {0.0, 0.0};
  for (int c = 0; c < P.medium.n; c++) density[c] = exp(-(hx / P.medium.scale[c]));
  for (int d = tid; d < ndirs; d += consumers) {
    const V3 omega = v3(dirs[3 * d], dirs[3 * d + 1], dirs[3 * d + 2]);
    const double mu = dot(v, omega);
    PointDir2 r;
    r.ox = omega.x;
    r.oy = omega.y;
    r.oz = omega.z;
    // overall-in-scattering (atmosphere.clj:147-151).  phase (atmosphere.clj:56-61) with b^1.5 as b sqrt(b): one
    // rounding more than pow, far below the float32 the sum is stored in, and a tenth of pow's instructions.
    double ph[2] = {0.0, 0.0};
    for (int c = 0; c < P.medium.n; c++) {
      const double g = P.medium.g[c], g2 = g * g, base = (1.0 + g2) - 2.0 * g * mu;
      ph[c] = (3.0 * (1.0 - g2) * (1.0 + mu * mu)) / (8.0 * kPi * (2.0 + g2) * (base * sqrt(base)));
    }
    for (int ch = 0; ch < 3; ch++) {
      double sum = 0.0;
      for (int c = 0; c < P.medium.n; c++) {
        double term = (P.medium.base[c][ch] * density[c]) * ph[c];
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
    r.q1 = -3.0 * (x.x / di.nmag);
    r.q2 = -3.0 * (((di.nx - x.x) * omega.x + di.ny * omega.y + di.nz * omega.z) / di.nmag);
    r.unused = 0.0;
    r.pad[0] = r.pad[1] = 0;
    pd[d] = r;
  }
  // the consumer threads meet here (the producer warp is on its own way): named barrier 1
  asm volatile("bar.sync 1, %0;" ::"n"(kPointScatterThreads) : "memory");

  const float phase_c0 = (float)((3.0 * (1.0 - phase_g * phase_g)) / (8.0 * kPi * (2.0 + phase_g * phase_g)));
  const double a_half = 0.5 * (double)(A - 1);
  const int a_last = A - 1, es_last = Es - 1;
  const float ws1 = as.s, ws0 = 1.0f - as.s;
  const int ru = (as.u - row_lo) * A, rv = (as.v - row_lo) * A;
  const double lx = l.x, ly = l.y, lz = l.z;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
  for (int d0 = 0; d0 < ndirs; d0 += kTileStages) {
    const unsigned parity = (unsigned)(d0 / kTileStages) & 1u;
#pragma unroll
    for (int j = 0; j < kTileStages; j++) {        // the ring position is the unrolled index: static addresses
      const int d = d0 + j;
      if (d >= ndirs) break;
      mbar_wait(&full[j], parity);
      if (active) {
        const float4 *tile = stages + (size_t)j * stage_texels;
        const PointDir2 &r = pd[d];
        const double mu = fma(r.oz, lz, fma(r.oy, ly, r.ox * lx));
        // sun-angle-to-index (atmosphere.clj:368-372); continuous coordinate
        FloorFrac fa = floor_frac(fma(a_half, mu, a_half));
        if (fa.u < 0) {
          fa.u = 0;
          fa.s = 0.0f;
        }
        const int au = min(fa.u, a_last), av = min(fa.u + 1, a_last);
        const float wa1 = fa.s, wa0 = 1.0f - fa.s;
        const float w00 = ws0 * wa0, w01 = ws0 * wa1, w10 = ws1 * wa0, w11 = ws1 * wa1;
        float4 sv = blend4(tile[ru + au], tile[ru + av], tile[rv + au], tile[rv + av], w00, w01, w10, w11);
        if (kTwoTables) {
          const float4 *tile_b = tile + nrows * A;
          const float4 m = blend4(tile_b[ru + au], tile_b[ru + av], tile_b[rv + au], tile_b[rv + av], w00, w01, w10, w11);
          // phase (atmosphere.clj:56-61) with the cancellation-prone base in double and the rest in float
          const float base = (float)((1.0 + phase_g * phase_g) - 2.0 * phase_g * mu);
          const float ph = phase_c0 * (float)(1.0 + mu * mu) / (base * sqrtf(base));
          sv.x = fmaf(m.x, ph, sv.x);
          sv.y = fmaf(m.y, ph, sv.y);
          sv.z = fmaf(m.z, ph, sv.z);
        }
        if (r.surface) {
          // surface-radiance (point, l): interpolation-table of dE over surface-radiance-space.  The sine of the sun
          // elevation at the ground point follows from point = x + t omega: y = -3 sin - 0.6 in two fused
          // multiply-adds; where the coordinate is about to be clamped to 0 it is recomputed as the reference
          // writes it, from (dot point l) / (mag point).
          double y = fma(mu, r.q2, fma(lx, r.q1, -0.6));
          if (fabs(y) < 3e-9) y = (0 - 3 * ((r.px * lx + r.py * ly + r.pz * lz) / r.pm)) - 0.6;
          FloorFrac fe = floor_frac(coord_from_exponent(s_exp, e_scale, y));
          if (fe.u < 0) {
            fe.u = 0;
            fe.s = 0.0f;
          }
          const int eu = min(fe.u, es_last), ev = min(fe.u + 1, es_last);
          const float4 *det = kDeShared ? s_de : de;
          const float4 *e0 = det + r.ehu * Es, *e1 = det + r.ehv * Es;
          const float we1 = fe.s, we0 = 1.0f - fe.s, wh1 = r.ehs, wh0 = 1.0f - r.ehs;
          float4 ev4;
          if (kDeShared)
            ev4 = blend4(e0[eu], e0[ev], e1[eu], e1[ev], wh0 * we0, wh0 * we1, wh1 * we0, wh1 * we1);
          else
            ev4 = blend4(ldg4(e0 + eu), ldg4(e0 + ev), ldg4(e1 + eu), ldg4(e1 + ev), wh0 * we0, wh0 * we1, wh1 * we0,
                         wh1 * we1);
          sv.x = fmaf(r.tb[0], ev4.x, sv.x);
          sv.y = fmaf(r.tb[1], ev4.y, sv.y);
          sv.z = fmaf(r.tb[2], ev4.z, sv.z);
        }
        acc0 = fmaf(r.sc[0], sv.x, acc0);
        acc1 = fmaf(r.sc[1], sv.y, acc1);
        acc2 = fmaf(r.sc[2], sv.z, acc2);
      }
      // Every thread releases the buffer itself.  (One arrive per warp after __syncwarp() is correct as well, but
      // measured 2 % slower, and compute-sanitizer's racecheck does not follow the ordering __syncwarp() gives the
      // other lanes' reads: 4 reported hazards against 0 this way, profiles/r2/sanitizer_*.txt.)
      mbar_arrive(&empty[j]);
    }
  }
  if (active) store_all(out, (size_t)he * ntex + texel, make_float4(acc0, acc1, acc2, 0.0f));
}

// ================================================================== launchers

static int env_variant(const char *name, int fallback) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : fallback;
}

static int ray_scatter_threads(const Params &P) {
  int ntex = P.shapes.s4[2] * P.shapes.s4[3];
  int t = (ntex + 31) / 32 * 32;
  return t < 128 ? 128 : (t > 1024 ? 1024 : t);
}

static size_t ray_scatter_smem_v1(const Params &P) {
  return sizeof(ViewSmem) + sizeof(LookupSmem) + 2 * (size_t)P.shapes.s4[2] * P.shapes.s4[3] * sizeof(float4);
}

static int ray_tile_entries(const Params &P) {
  const int S = P.shapes.s4[2], A = P.shapes.s4[3];
  return S * A;                               // must match k_ray_scatter
}

static size_t ray_scatter_smem_v2(const Params &P) {
  return 16 + (size_t)P.shapes.ray_steps * sizeof(RaySample) + (size_t)((kExpTabSize + 1) & ~1) * sizeof(double) +
         2 * (size_t)ray_tile_entries(P) * sizeof(float4);
}

size_t ray_scatter_smem(const Params &P) { return std::max(ray_scatter_smem_v1(P), ray_scatter_smem_v2(P)); }

static int ray_scatter_variant() {
  static const int variant = env_variant("ATMLUT_K6", 2);
  return variant;
}

// the per-sample records live outside the kernel only for tiles of at most 1024 texels (one texel per thread)
bool ray_scatter_uses_samples(const Params &P) {
  return ray_scatter_variant() >= 2 && P.shapes.s4[2] * P.shapes.s4[3] <= 1024;
}

size_t ray_sample_bytes(const Params &P, int he_count) {
  return (size_t)std::max(he_count, 1) * P.shapes.ray_steps * sizeof(RaySample);
}

cudaError_t launch_ray_prepare(const Params &P, Shard shard, int he_count, void *samples, unsigned long long *counter,
                               const void *view_packs, cudaStream_t st) {
  if (he_count <= 0 || !ray_scatter_uses_samples(P)) return cudaSuccess;
  const size_t smem = sizeof(ViewSmem) + (size_t)P.shapes.ray_steps * sizeof(RaySample);
  cudaError_t e = cudaFuncSetAttribute(k_ray_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_ray_prepare<<<he_count, 256, smem, st>>>(P, shard, (RaySample *)samples, counter, (const unsigned char *)view_packs);
  return cudaGetLastError();
}

cudaError_t launch_ray_scatter(const Params &P, Shard shard, int he_count, const void *samples, const float4 *dj,
                               const double *exp_table, PeerOut out, unsigned long long *counter, cudaStream_t st) {
  if (he_count <= 0) return cudaSuccess;
  const bool v2 = ray_scatter_uses_samples(P);
  const size_t smem = v2 ? ray_scatter_smem_v2(P) : ray_scatter_smem_v1(P);
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;   // light-elevation x heading tile too large
  static const bool split = env_variant("ATMLUT_K6_SPLIT", 1) != 0;
  cudaError_t e = !v2 ? cudaFuncSetAttribute(k_ray_scatter_v1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                  : split ? cudaFuncSetAttribute(k_ray_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                          : cudaFuncSetAttribute(k_ray_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (v2 && split)
    k_ray_scatter<true><<<he_count, ray_scatter_threads(P), smem, st>>>(P, shard, (const RaySample *)samples, dj, exp_table, out);
  else if (v2)
    k_ray_scatter<false><<<he_count, ray_scatter_threads(P), smem, st>>>(P, shard, (const RaySample *)samples, dj, exp_table, out);
  else
    k_ray_scatter_v1<<<he_count, ray_scatter_threads(P), smem, st>>>(P, shard, dj, exp_table, out, counter);
  return cudaGetLastError();
}

cudaError_t launch_point_scatter_prepare(const Params &P, const double *dirs, int ndirs, DirInfo *info,
                                         cudaStream_t st) {
  int n = P.shapes.s4[0] * ndirs;
  k_point_scatter_prepare<<<(n + 63) / 64, 64, 0, st>>>(P, dirs, ndirs, info);
  return cudaGetLastError();
}

cudaError_t launch_blend_dir_tiles(const Params &P, const float4 *tab, const DirInfo *info, int ndirs, int h_first,
                                   int h_stride, int h_count, float4 *tiles, cudaStream_t st) {
  if (h_count <= 0) return cudaSuccess;
  k_blend_dir_tiles<<<h_count * ndirs, 256, 0, st>>>(P, tab, info, ndirs, h_first, h_stride, tiles);
  return cudaGetLastError();
}

template <bool kTwoTables, bool kDeShared>
static cudaError_t launch_point_scatter_v2(const Params &P, Shard shard, int he_count, int chunks, int rows_max,
                                           size_t smem, const float4 *tiles_a, const float4 *tiles_b, double phase_g,
                                           const float4 *de, const double *dirs, const double *weights, int ndirs,
                                           const DirInfo *info, const double *exp_table, PeerOut out, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(k_point_scatter<kTwoTables, kDeShared>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_point_scatter<kTwoTables, kDeShared><<<he_count * chunks, kPointScatterBlock, smem, st>>>(
      P, shard, chunks, rows_max, tiles_a, tiles_b, phase_g, de, dirs, weights, ndirs, info, exp_table, out);
  return cudaGetLastError();
}

cudaError_t launch_point_scatter(const Params &P, Shard shard, int he_count, const float4 *tiles_a,
                                 const float4 *tiles_b, double phase_g, const float4 *de, const double *dirs,
                                 const double *weights, int ndirs, const DirInfo *info, const double *exp_table,
                                 PeerOut out, cudaStream_t st) {
  if (he_count <= 0) return cudaSuccess;
  static const int variant = env_variant("ATMLUT_K4", 2);
  const int S = P.shapes.s4[2], A = P.shapes.s4[3], ntex = S * A;
  if (variant >= 2) {
    // rows of a direction tile one chunk of 256 consecutive texels can touch: its own light-elevation rows, one
    // more on either side (forward o backward moves that coordinate by < 1e-11), and the chunk may start mid-row
    const int chunks = (ntex + kPointScatterThreads - 1) / kPointScatterThreads;
    const int rows_max = std::min(S, (kPointScatterThreads + A - 1) / A + 3);
    const size_t ne = (size_t)P.shapes.se[0] * P.shapes.se[1];
    const size_t fixed = 128 + (size_t)ndirs * sizeof(PointDir2) + (size_t)((kExpTabSize + 1) & ~1) * sizeof(double);
    const size_t stages = (size_t)kTileStages * rows_max * A * sizeof(float4) * (tiles_b ? 2 : 1);
    const bool de_shared = ne * sizeof(float4) <= 32 * 1024 && fixed + stages + ne * sizeof(float4) <= 75 * 1024;
    const size_t smem = fixed + stages + (de_shared ? ne * sizeof(float4) : 0);
    if (smem <= 227 * 1024) {
      if (tiles_b)
        return de_shared ? launch_point_scatter_v2<true, true>(P, shard, he_count, chunks, rows_max, smem, tiles_a,
                                                               tiles_b, phase_g, de, dirs, weights, ndirs, info, exp_table, out, st)
                         : launch_point_scatter_v2<true, false>(P, shard, he_count, chunks, rows_max, smem, tiles_a,
                                                                tiles_b, phase_g, de, dirs, weights, ndirs, info, exp_table, out, st);
      return de_shared ? launch_point_scatter_v2<false, true>(P, shard, he_count, chunks, rows_max, smem, tiles_a,
                                                              tiles_b, phase_g, de, dirs, weights, ndirs, info, exp_table, out, st)
                       : launch_point_scatter_v2<false, false>(P, shard, he_count, chunks, rows_max, smem, tiles_a,
                                                               tiles_b, phase_g, de, dirs, weights, ndirs, info, exp_table, out, st);
    }
  }
  size_t smem = (size_t)ndirs * sizeof(PointDir);
  cudaError_t e = cudaFuncSetAttribute(k_point_scatter_v1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_point_scatter_v1<<<he_count, 256, smem, st>>>(P, shard, tiles_a, tiles_b, phase_g, de, dirs, weights, ndirs, info,
                                                  exp_table, out);
  return cudaGetLastError();
}

}  // namespace atm
