// Device-side building blocks shared by the table kernels: the fast overall-extinction sampler
// (FP64 position polynomial -> FP32 height polynomial -> MUFU.EX2), float4 table lookups, and the
// parameter block every kernel receives.
#pragma once

#include "atm_math.cuh"

#ifndef ATMLUT_SAMPLER_UNROLL
#define ATMLUT_SAMPLER_UNROLL 4
#endif

namespace atm {

constexpr int kSamplerUnroll = ATMLUT_SAMPLER_UNROLL;   // 4-sample groups per loop trip of the hot sampler
constexpr int kMaxSteps = 256;       // ray_steps limit of the table kernels (shared-memory arrays)
constexpr float kLog2e = 1.4426950408889634f;

struct Shapes {
  int s4[4];  // height, elevation, light-elevation, heading
  int st[2];  // transmittance height, elevation
  int se[2];  // surface height, sun elevation
  int ray_steps, sphere_steps;
};

// Constants of the fast sampler.  Heights are measured from R' = R - delta so that
// r^2 - R'^2 stays strictly positive (its float image keeps full relative precision), and
// h' = r - R' = h + delta is folded back into the exponent as an additive constant.
struct Fast {
  int poly;          // 1: (Rt^2 - R'^2) / R'^2 small enough for the series of sqrt(1+u) - 1
  double rp2;        // R'^2
  double inv_rp2;    // 1 / R'^2
  float k[2];        // -R' * log2(e) / scale_c     (exponent slope in units of h'/R')
  float b[2];        // delta * log2(e) / scale_c   (exponent offset)
  float ext[2][3];   // extinction at h = 0: base_c / quotient_c
  // Commensurable scale heights (Earth: 1200 m and 8000 m = 24000 m / 20 and / 3): both densities are integer powers
  // of ONE exponential t = exp(-h / 24000 m), t^20 and t^3 -- six packed multiplies on the FMA pipe instead of a
  // second MUFU.EX2, the pipe the first-order kernel is bound by.  pow_mode 1: component 0 = t^20, component 1 = t^3;
  // 2: the other way round; 0: scale heights in no such ratio, two exponentials.
  int pow_mode;
  float kt, bt;      // exponent slope and offset of t, like k[] and b[]
  // pow_alt 1: only every second pair of samples takes the one-exponential form, which balances the MUFU and the FMA pipe
  // (tools/pipe_peaks.cu: 35.7 clocks per pair with two exponentials, 31.0 with one, 29.7 alternating).
  int pow_alt;
  // degree of the series of sqrt(1+u) - 1 in u: 2 where the atmosphere is so thin against the radius that the u^3 term
  // moves no exponent by more than 4e-6 (Earth: u <= 0.0115), else 4
  int degree;
};

struct Params {
  Planet planet;
  Medium medium;
  Shapes shapes;
  Fast fast;
  double intensity[3];
};

// float image of a positive normal double by truncation: two integer instructions instead of an
// F2F.F32.F64 (which shares the 16/clk/SM conversion/SFU rate with MUFU.EX2, tools/pipe_peaks.cu).
__device__ __forceinline__ float trunc_d2f(double d) {
  int hi = __double2hiint(d), lo = __double2loint(d);
  return __int_as_float(__funnelshift_l(lo, hi - 0x38000000, 3));
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp(-h/scale_c) for both components from u = (|q|^2 - R'^2) / R'^2  (fast path, u > 0 small)
// sqrt(1+u) - 1 = u (1/2 - u/8 + u^2/16 - 5u^3/128 + 7u^4/256 - ...)
// ONE: both densities as powers of one exponential (component 0 = t^20, component 1 = t^3; the caller swaps for the
// other order); k0 .. b1: exponent slopes and offsets of the two-exponential form.
template <bool ONE, int DEG>
__device__ __forceinline__ void densities_from_u(const Fast &f, float k0, float b0, float k1, float b1, float u, float &e0,
                                                 float &e1) {
  float q;
  if (DEG == 2) {
    q = fmaf(u, 0.0625f, -0.125f);
  } else {
    q = fmaf(u, 0.02734375f, -0.0390625f);
    q = fmaf(q, u, 0.0625f);
    q = fmaf(q, u, -0.125f);
  }
  q = fmaf(q, u, 0.5f);
  float hq = u * q;  // h' / R'
  if (!ONE) {
    e0 = ex2_approx(fmaf(hq, k0, b0));
    e1 = ex2_approx(fmaf(hq, k1, b1));
  } else {
    const float t = ex2_approx(fmaf(hq, f.kt, f.bt));
    const float p2 = t * t, p4 = p2 * p2, p8 = p4 * p4, p16 = p8 * p8;
    e0 = p4 * p16;
    e1 = t * p2;
  }
}

// Coefficients of u(m) = A + B m + C m^2, m = j + 1/2, for the samples q_j = o + dir (j + 1/2) / steps
// given oo = |o|^2, od = o.dir, dd = |dir|^2.
struct Quad {
  double A, B, C;
};

__device__ __forceinline__ Quad make_quad(const Fast &f, double oo, double od, double dd, int steps) {
  double inv = 1.0 / (double)steps;
  Quad q;
  q.A = (oo - f.rp2) * f.inv_rp2;
  q.B = (2.0 * od * inv) * f.inv_rp2;
  q.C = (dd * inv * inv) * f.inv_rp2;
  return q;
}

// slow but general: exp(-h/scale) in double (planets whose atmosphere is not thin)
__device__ __forceinline__ void densities_general(const Params &P, double rq2, float &e0, float &e1) {
  double h = sqrt(rq2) - P.planet.radius;
  e0 = (float)exp(-(h / P.medium.scale[0]));
  e1 = (float)exp(-(h / P.medium.scale[1]));
}

// acc_c += exp(-h/scale_c) for TWO samples at once with Blackwell's packed FP32 instructions (FFMA2 / FMUL2: one issue
// slot for two lanes of work).  Two exponentials per sample: four MUFU.EX2 per pair; ONE: two, and the powers
// t^3 = t t^2 and t^20 = t^4 t^16 fused with the accumulation (six packed instructions).
template <bool ONE, int DEG>
__device__ __forceinline__ void accumulate_densities2(const Fast &f, float k0, float b0, float k1, float b1, float2 u,
                                                      float2 &acc0, float2 &acc1) {
  float2 q;
  if (DEG == 2) {
    q = __ffma2_rn(u, make_float2(0.0625f, 0.0625f), make_float2(-0.125f, -0.125f));
  } else {
    q = __ffma2_rn(u, make_float2(0.02734375f, 0.02734375f), make_float2(-0.0390625f, -0.0390625f));
    q = __ffma2_rn(q, u, make_float2(0.0625f, 0.0625f));
    q = __ffma2_rn(q, u, make_float2(-0.125f, -0.125f));
  }
  q = __ffma2_rn(q, u, make_float2(0.5f, 0.5f));
  const float2 hq = __fmul2_rn(u, q);  // h' / R'
  if (!ONE) {
    const float2 a0 = __ffma2_rn(hq, make_float2(k0, k0), make_float2(b0, b0));
    const float2 a1 = __ffma2_rn(hq, make_float2(k1, k1), make_float2(b1, b1));
    acc0 = __fadd2_rn(acc0, make_float2(ex2_approx(a0.x), ex2_approx(a0.y)));
    acc1 = __fadd2_rn(acc1, make_float2(ex2_approx(a1.x), ex2_approx(a1.y)));
  } else {
    const float2 a = __ffma2_rn(hq, make_float2(f.kt, f.kt), make_float2(f.bt, f.bt));
    const float2 t = make_float2(ex2_approx(a.x), ex2_approx(a.y));
    const float2 p2 = __fmul2_rn(t, t);
    const float2 p4 = __fmul2_rn(p2, p2);
    const float2 p8 = __fmul2_rn(p4, p4);
    const float2 p16 = __fmul2_rn(p8, p8);
    acc0 = __ffma2_rn(p4, p16, acc0);
    acc1 = __ffma2_rn(t, p2, acc1);
  }
}

// Sum over the `steps` samples of a segment of exp(-h(q_j)/scale_c), one thread per segment: u(m) advances by forward
// differences in double, four samples per step (2 DADD); the other three u are the float image of the first plus float
// differences.  Packed accumulators, two independent pairs in flight (pair A: ONE_A, pair B: ONE_B).
// `swap`: the one-exponential form yields (t^20, t^3); the medium lists its components the other way round.
template <bool ONE_A, bool ONE_B, int DEG>
__device__ __forceinline__ void density_sums_fast(const Fast &f, bool swap, const Quad &q, int steps, float &s0, float &s1) {
  const float k0 = f.k[swap ? 1 : 0], b0 = f.b[swap ? 1 : 0], k1 = f.k[swap ? 0 : 1], b1 = f.b[swap ? 0 : 1];
  double u = fma(fma(q.C, 0.5, q.B), 0.5, q.A);   // u at m = 1/2
  const double d1 = q.B + 2.0 * q.C;               // u(m+1) - u(m) at m = 1/2
  const double d2 = 2.0 * q.C;                     // second difference
  // four samples per trip: u(m) in double (2 DADD per trip), u(m+1..3) = its float image plus float differences
  // e_i(m) = u(m+i) - u(m) = i d1(m) + (i (i-1) / 2) d2, which grow by 4 i d2 per trip
  double step4 = 4.0 * d1 + 6.0 * d2;              // u(m+4) - u(m)
  const double step4_inc = 16.0 * d2;
  // the float differences from two conversions (they share the MUFU pipe): multiples of d2 by powers of two are exact
  const float d1f = (float)d1, d2f = (float)d2;
  float e1 = d1f;
  float2 e23 = make_float2(fmaf(2.0f, d1f, d2f), fmaf(3.0f, d1f, 3.0f * d2f));
  const float e1_inc = 4.0f * d2f;
  const float2 e23_inc = make_float2(8.0f * d2f, 12.0f * d2f);
  float2 acc0a = make_float2(0.f, 0.f), acc1a = make_float2(0.f, 0.f);
  float2 acc0b = make_float2(0.f, 0.f), acc1b = make_float2(0.f, 0.f);
  int j = 0;
#pragma unroll kSamplerUnroll
  for (; j + 4 <= steps; j += 4) {
    const float ue = trunc_d2f(u);
    accumulate_densities2<ONE_A, DEG>(f, k0, b0, k1, b1, make_float2(ue, ue + e1), acc0a, acc1a);
    accumulate_densities2<ONE_B, DEG>(f, k0, b0, k1, b1, __fadd2_rn(make_float2(ue, ue), e23), acc0b, acc1b);
    u += step4;
    step4 += step4_inc;
    e1 += e1_inc;
    e23 = __fadd2_rn(e23, e23_inc);
  }
  float t0 = 0.f, t1 = 0.f;
  if (j < steps) {
    // tail of up to 3 samples, one at a time: u(m+1) = u(m) + d1(m)
    double d1m = d1 + (double)j * d2;
    for (; j < steps; j++) {
      float x0, x1;
      densities_from_u<ONE_B, DEG>(f, k0, b0, k1, b1, trunc_d2f(u), x0, x1);
      t0 += x0;
      t1 += x1;
      u += d1m;
      d1m += d2;
    }
  }
  const float r0 = ((acc0a.x + acc0b.x) + (acc0a.y + acc0b.y)) + t0;
  const float r1 = ((acc1a.x + acc1b.x) + (acc1a.y + acc1b.y)) + t1;
  s0 = swap ? r1 : r0;
  s1 = swap ? r0 : r1;
}

__device__ __forceinline__ void density_sums_seq(const Params &P, const Quad &q, int steps, float &s0, float &s1) {
  const Fast &f = P.fast;
  if (f.poly) {
    if (f.pow_mode && f.degree == 2) {
      if (f.pow_alt)
        density_sums_fast<false, true, 2>(f, f.pow_mode == 2, q, steps, s0, s1);
      else
        density_sums_fast<true, true, 2>(f, f.pow_mode == 2, q, steps, s0, s1);
    } else if (f.degree == 2) {
      density_sums_fast<false, false, 2>(f, false, q, steps, s0, s1);
    } else {
      density_sums_fast<false, false, 4>(f, false, q, steps, s0, s1);
    }
  } else {
    s0 = 0.f;
    s1 = 0.f;
    for (int j = 0; j < steps; j++) {
      double m = (double)j + 0.5;
      double u = fma(fma(q.C, m, q.B), m, q.A);
      float e0, e1;
      densities_general(P, fma(u, P.fast.rp2, P.fast.rp2), e0, e1);
      s0 += e0;
      s1 += e1;
    }
  }
}

// exp(-(ext0 * c0 + ext1 * c1)) per colour channel, c = column densities (metres of sea-level medium)
__device__ __forceinline__ void transmittance_rgb(const Fast &f, float c0, float c1, float t[3]) {
#pragma unroll
  for (int ch = 0; ch < 3; ch++) {
    float tau = fmaf(f.ext[1][ch], c1, f.ext[0][ch] * c0);
    t[ch] = ex2_approx(-tau * kLog2e);
  }
}

// ------------------------------------------------------------------ sun-elevation lookup coordinate

// exp(y) for y in [-4, 2.5] to double accuracy from a table of exp(i/64) (in shared memory, filled with the
// host's exp) and a degree-6 polynomial on the remainder |y - i/64| <= 1/128 (truncation 4e-19 relative, i.e. the
// result is as good as the table entry: lookups that sit exactly on a table node then carry the same one-ulp
// weights on the neighbouring rows as the reference's own arithmetic, not 1e-12 ones).
// The library exp costs ~60 issue slots here (11 polynomial constants reloaded through the uniform datapath);
// this one costs 14.  Used only for lookup coordinates, which are continuous in it.
constexpr int kExpTabLo = -4 * 64;                     // i = -256 .. 160
constexpr int kExpTabHi = 160;
constexpr int kExpTabSize = kExpTabHi - kExpTabLo + 1;

__device__ __forceinline__ double exp_tab(const double *tab, double y) {
  const double magic = 6755399441055744.0;            // 1.5 * 2^52: the low word of y*64 + magic is rint(y*64)
  const double t = fma(y, 64.0, magic);
  const int i = __double2loint(t);                    // -256 .. 160
  const double r = fma(t - magic, -1.0 / 64.0, y);    // y - i/64
  double p = fma(r, 1.0 / 720.0, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return tab[i - kExpTabLo] * p;
}

// scale * (1 - exp(y)) from a table that already holds scale * exp(i/64): one fused multiply-add after the
// polynomial.  y in [-4, 2.5]; the result is <= 0 for y >= 0.
__device__ __forceinline__ double scaled_one_minus_exp(const double *scaled_tab, double scale, double y) {
  const double magic = 6755399441055744.0;
  const double t = fma(y, 64.0, magic);
  const int i = __double2loint(t);
  const double r = fma(t - magic, -1.0 / 64.0, y);
  double p = fma(r, 1.0 / 720.0, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return fma(-scaled_tab[i - kExpTabLo], p, scale);
}

__device__ __forceinline__ void fill_exp_tab(double *tab, const double *global_tab) {
  for (int i = threadIdx.x; i < kExpTabSize; i += blockDim.x) tab[i] = global_tab[i];
}

// atmosphere.clj:322-326 for lookups inside the integration kernels, from the sine of the sun elevation.
// sin <= -0.2 is below the table's range: the reference's max(0, .) gives exactly 0 there.
// `scale` = (size - 1) / (1 - exp(-3.6)), hoisted by the caller.  Plain compares instead of fmax/fmin: the
// operands are never NaN here and the IEEE min/max sequences cost twice as many instructions in double.
__device__ __forceinline__ double sun_elevation_scale(int size) { return (double)(size - 1) / (1 - 0.02732372244729256); }

__device__ __forceinline__ double sun_elevation_coord(const double *exp_table, double scale, double sin_elevation) {
  double y = 0 - 3 * sin_elevation - 0.6;
  if (y >= 0.0) return 0.0;
  y = y < -4.0 ? -4.0 : y;                 // table range (|sin| <= 1 gives y >= -3.6)
  const double c = (1 - exp_tab(exp_table, y)) * scale;
  return c < 0.0 ? 0.0 : c;
}

// ------------------------------------------------------------------ float4 table lookups

// interpolate.clj:75-98: clip to [0, n-1], u = floor, v = min(u + 1, n - 1), s = i - u
struct Axis {
  int u, v;
  float s;
};

__device__ __forceinline__ Axis axis_from(float c, int n) {
  float i = fminf(fmaxf(c, 0.0f), (float)(n - 1));
  float u = floorf(i);
  Axis a;
  a.u = (int)u;
  a.v = min(a.u + 1, n - 1);
  a.s = i - u;
  return a;
}

__device__ __forceinline__ Axis axis_from(double c, int n) {
  double i = fmin(fmax(c, 0.0), (double)(n - 1));
  double u = floor(i);
  Axis a;
  a.u = (int)u;
  a.v = min(a.u + 1, n - 1);
  a.s = (float)(i - u);
  return a;
}

// interpolate.clj:75-98 for a coordinate known to be >= 0 and not NaN: clip to n - 1, floor, fraction
__device__ __forceinline__ Axis axis_from_nonneg(double c, int n, double nmax) {
  const double i = c < nmax ? c : nmax;
  const double u = floor(i);
  Axis a;
  a.u = (int)u;
  a.v = min(a.u + 1, n - 1);
  a.s = (float)(i - u);
  return a;
}

// floor and fraction of a lookup coordinate |c| < 2^31 with the 1.5 * 2^52 rounding constant: three DADD and ONE
// conversion (of the fraction, which keeps its relative precision) instead of F2I + FRND + F2F, all of which
// share the 16 / clk / SM conversion pipe.  c < 0 gives u < 0: the caller clamps (interpolate.clj:75-78).
struct FloorFrac {
  int u;
  float s;
};

__device__ __forceinline__ FloorFrac floor_frac(double c) {
  const double magic = 6755399441055744.0;
  const double t = c + magic;
  const int nearest = __double2loint(t);        // rint(c)
  const double d = c - (t - magic);             // in [-1/2, 1/2], exact
  const float df = (float)d;
  const bool below = __double2hiint(d) < 0;
  FloorFrac r;
  r.u = below ? nearest - 1 : nearest;
  r.s = below ? df + 1.0f : df;
  return r;
}

// interpolate.clj:81-84 mix: a (1 - s) + b s
__device__ __forceinline__ float4 mix4(float4 a, float4 b, float s) {
  float t = 1.0f - s;
  return make_float4(fmaf(b.x, s, a.x * t), fmaf(b.y, s, a.y * t), fmaf(b.z, s, a.z * t), 0.0f);
}

__device__ __forceinline__ float4 ldg4(const float4 *p) { return __ldg(p); }

// 4-D multilinear lookup, first axis outermost (interpolate.clj:87-98)
__device__ __forceinline__ float4 lookup4(const float4 *__restrict__ tab, const int s4[4], Axis h, Axis e, Axis s,
                                          Axis a) {
  const int E = s4[1], S = s4[2], A = s4[3];
  float4 he[2][2];
#pragma unroll
  for (int ih = 0; ih < 2; ih++) {
#pragma unroll
    for (int ie = 0; ie < 2; ie++) {
      const float4 *base = tab + ((size_t)((ih ? h.v : h.u) * E + (ie ? e.v : e.u)) * S) * A;
      const float4 *r0 = base + (size_t)s.u * A;
      const float4 *r1 = base + (size_t)s.v * A;
      float4 x0 = mix4(ldg4(r0 + a.u), ldg4(r0 + a.v), a.s);
      float4 x1 = mix4(ldg4(r1 + a.u), ldg4(r1 + a.v), a.s);
      he[ih][ie] = mix4(x0, x1, s.s);
    }
  }
  return mix4(mix4(he[0][0], he[0][1], e.s), mix4(he[1][0], he[1][1], e.s), h.s);
}

// 2-D bilinear lookup in a shared-memory tile
__device__ __forceinline__ float4 lookup2_smem(const float4 *tab, int cols, Axis r, Axis c) {
  const float4 *r0 = tab + r.u * cols;
  const float4 *r1 = tab + r.v * cols;
  return mix4(mix4(r0[c.u], r0[c.v], c.s), mix4(r1[c.u], r1[c.v], c.s), r.s);
}

// 2-D bilinear lookup
__device__ __forceinline__ float4 lookup2(const float4 *__restrict__ tab, int cols, Axis r, Axis c) {
  const float4 *r0 = tab + (size_t)r.u * cols;
  const float4 *r1 = tab + (size_t)r.v * cols;
  return mix4(mix4(ldg4(r0 + c.u), ldg4(r0 + c.v), c.s), mix4(ldg4(r1 + c.u), ldg4(r1 + c.v), c.s), r.s);
}

}  // namespace atm
