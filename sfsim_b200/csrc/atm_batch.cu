// Batch point evaluators: the public functions of sfsim.atmosphere at arbitrary arguments, one
// thread per item, in double precision on the device.  These are the entry points the host-side
// mirror (sfsim_b200/atmosphere.py, or the Clojure shim) calls for single evaluations and for the
// known-answer tests; the table kernels in atm_tables.cu are the throughput path.
#include <vector>

#include "atm_api_internal.h"
#include "atm_math.cuh"

namespace atm {

namespace {

__device__ __forceinline__ V3 load3(const double *p, int i) { return v3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
__device__ __forceinline__ void store3(double *p, int i, const double v[3]) {
  p[3 * i] = v[0];
  p[3 * i + 1] = v[1];
  p[3 * i + 2] = v[2];
}

__global__ void k_transmittance_batch(Planet pl, Medium md, int steps, int count, const double *x, const double *x0,
                                      double *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double t[3];
  transmittance_points(pl, md, steps, load3(x, i), load3(x0, i), t);
  store3(out, i, t);
}

__global__ void k_transmittance_dir_batch(Planet pl, Medium md, int steps, int count, const double *x,
                                          const double *v, const int *above, double *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double t[3];
  transmittance_dir(pl, md, steps, load3(x, i), load3(v, i), above[i] != 0, t);
  store3(out, i, t);
}

// atmosphere.clj:131-137
__global__ void k_surface_radiance_base_batch(Planet pl, Medium md, int steps, V3 intensity, int count,
                                              const double *x, const double *l, double *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  V3 xi = load3(x, i), li = load3(l, i);
  double m = mag(xi);
  V3 normal = v3(xi.x / m, xi.y / m, xi.z / m);
  double t[3];
  transmittance_dir(pl, md, steps, xi, li, true, t);
  double f = fmax(0.0, dot(normal, li));
  double r[3] = {(t[0] * intensity.x) * f, (t[1] * intensity.y) * f, (t[2] * intensity.z) * f};
  store3(out, i, r);
}

// atmosphere.clj:140-189: kind 0 point-scatter-component, 1 strength-component, 2 point-scatter-base
__device__ void point_scatter_first_order(const Planet &pl, const Medium &md, int kind, int component, int steps,
                                          V3 intensity, V3 x, V3 v, V3 l, double out[3]) {
  // filtered-sun-light (atmosphere.clj:154-160)
  double sun[3] = {0.0, 0.0, 0.0};
  if (is_above_horizon(pl, x, l)) {
    double t[3];
    transmittance_dir(pl, md, steps, x, l, true, t);
    sun[0] = intensity.x * t[0];
    sun[1] = intensity.y * t[1];
    sun[2] = intensity.z * t[2];
  }
  const double h = height(pl, x);
  const double mu = dot(v, l);
  for (int ch = 0; ch < 3; ch++) {
    double s;
    if (kind == 1) {
      s = scattering(md, component, ch, h);
    } else if (kind == 0) {
      s = scattering(md, component, ch, h) * phase(md.g[component], mu);
    } else {
      s = 0.0;
      for (int c = 0; c < md.n; c++) {
        double term = scattering(md, c, ch, h) * phase(md.g[c], mu);
        s = (c == 0) ? term : s + term;
      }
    }
    out[ch] = s * sun[ch];
  }
}

__global__ void k_point_scatter_first_order_batch(Planet pl, Medium md, int kind, int component, int steps,
                                                  V3 intensity, int count, const double *x, const double *v,
                                                  const double *l, double *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double r[3];
  point_scatter_first_order(pl, md, kind, component, steps, intensity, load3(x, i), load3(v, i), load3(l, i), r);
  store3(out, i, r);
}

// atmosphere.clj:192-200 ray-scatter with a first-order point-scatter function, plain restatement:
// one warp per item, lanes over the outer samples
__global__ void k_ray_scatter_first_order_batch(Planet pl, Medium md, int kind, int component, int steps,
                                                V3 intensity, int count, const double *x, const double *v,
                                                const double *l, const int *above, double *out) {
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (i >= count) return;
  V3 xi = load3(x, i), vi = load3(v, i), li = load3(l, i);
  V3 point = above[i] ? atmosphere_intersection(pl, xi, vi) : surface_intersection(pl, xi, vi);
  V3 d = point - xi;
  double stepsize = 1.0 / (double)steps;
  double a = stepsize * mag(d);
  double acc[3] = {0.0, 0.0, 0.0};
  for (int k = lane; k < steps; k += 32) {
    double s = (0.5 + (double)k) * stepsize;
    V3 p = xi + d * s;
    double t[3], j[3];
    transmittance_points(pl, md, steps, xi, p, t);
    point_scatter_first_order(pl, md, kind, component, steps, intensity, p, vi, li, j);
    for (int ch = 0; ch < 3; ch++) acc[ch] += (t[ch] * j[ch]) * a;
  }
  for (int ch = 0; ch < 3; ch++)
    for (int o = 16; o > 0; o >>= 1) acc[ch] += __shfl_xor_sync(0xffffffffu, acc[ch], o);
  if (lane == 0) store3(out, i, acc);
}

__global__ void k_index_forward(Planet pl, int which, int s0, int s1, int s2, int s3, int count,
                                const double *point, const double *direction, const double *light, const int *above,
                                double *indices) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int shape[4] = {s0, s1, s2, s3};
  if (which == 0) {
    double idx[4];
    ray_scatter_forward(pl, shape, load3(point, i), load3(direction, i), load3(light, i), above[i] != 0, idx);
    for (int k = 0; k < 4; k++) indices[4 * i + k] = idx[k];
  } else if (which == 1) {
    double idx[2];
    surface_radiance_forward(pl, shape, load3(point, i), load3(light, i), idx);
    indices[2 * i] = idx[0];
    indices[2 * i + 1] = idx[1];
  } else {
    double idx[2];
    transmittance_forward(pl, shape, load3(point, i), load3(direction, i), above[i] != 0, idx);
    indices[2 * i] = idx[0];
    indices[2 * i + 1] = idx[1];
  }
}

__global__ void k_index_backward(Planet pl, int which, int s0, int s1, int s2, int s3, int count,
                                 const double *indices, double *point, double *direction, double *light,
                                 int *above) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int shape[4] = {s0, s1, s2, s3};
  V3 p = v3(0, 0, 0), d = v3(0, 0, 0), l = v3(0, 0, 0);
  bool ab = false;
  if (which == 0)
    ray_scatter_backward(pl, shape, indices[4 * i], indices[4 * i + 1], indices[4 * i + 2], indices[4 * i + 3], p, d,
                         l, ab);
  else if (which == 1)
    surface_radiance_backward(pl, shape, indices[2 * i], indices[2 * i + 1], p, l);
  else
    transmittance_backward(pl, shape, indices[2 * i], indices[2 * i + 1], p, d, ab);
  if (point) {
    point[3 * i] = p.x;
    point[3 * i + 1] = p.y;
    point[3 * i + 2] = p.z;
  }
  if (direction) {
    direction[3 * i] = d.x;
    direction[3 * i + 1] = d.y;
    direction[3 * i + 2] = d.z;
  }
  if (light) {
    light[3 * i] = l.x;
    light[3 * i + 1] = l.y;
    light[3 * i + 2] = l.z;
  }
  if (above) above[i] = ab ? 1 : 0;
}

// the scalar index maps of atmosphere.clj:233-384, one item per thread (fn codes in sfsim_atmosphere.h)
__global__ void k_index_map(Planet pl, int fn, int size, int count, const double *a, const double *b, const int *flag,
                            double *out, int *out_flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  V3 va = a ? load3(a, i) : v3(0, 0, 0), vb = b ? load3(b, i) : v3(0, 0, 0);
  V3 r = v3(0, 0, 0);
  bool ab = false;
  switch (fn) {
    case 0: r.x = elevation_to_index(pl, size, va, vb, flag[i] != 0); break;
    case 1: index_to_elevation(pl, size, va.x, va.y, r, ab); break;
    case 2: r.x = height_to_index(pl, size, va); break;
    case 3: r = index_to_height(pl, size, va.x); break;
    case 4: r.x = sun_elevation_to_index(size, va, vb); break;
    case 5: r.x = index_to_sin_sun_elevation(size, va.x); break;
    case 6: r.x = sun_angle_to_index(size, va, vb); break;
    case 7: r = index_to_sun_direction(size, va, vb.x, vb.y); break;
    default: r.x = horizon_distance(pl, va.x); break;
  }
  out[3 * i] = r.x;
  out[3 * i + 1] = r.y;
  out[3 * i + 2] = r.z;
  if (out_flag) out_flag[i] = ab ? 1 : 0;
}

// interpolate.clj:87-98 interpolate-value, first axis outermost, mix = a (1 - s) + b s
__global__ void k_interpolate(const float *table, int dims, int n0, int n1, int n2, int n3, int ncomp, int count,
                              const double *coords, float *out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const int shape[4] = {n0, n1, n2, n3};
  int lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
  float frac[4] = {0.f, 0.f, 0.f, 0.f};
  long long stride[4];
  long long s = ncomp;
  for (int d = dims - 1; d >= 0; d--) {
    stride[d] = s;
    s *= shape[d];
  }
  for (int d = 0; d < dims; d++) {
    double c = fmin(fmax(coords[(long long)i * dims + d], 0.0), (double)(shape[d] - 1));
    double u = floor(c);
    lo[d] = (int)u;
    hi[d] = min(lo[d] + 1, shape[d] - 1);
    frac[d] = (float)(c - u);
  }
  for (int k = 0; k < ncomp; k++) {
    // evaluate the recursion bottom-up over the 2^dims corners
    float vals[16];
    const int corners = 1 << dims;
    for (int c = 0; c < corners; c++) {
      long long off = k;
      for (int d = 0; d < dims; d++) off += (long long)(((c >> (dims - 1 - d)) & 1) ? hi[d] : lo[d]) * stride[d];
      vals[c] = table[off];
    }
    for (int d = dims - 1; d >= 0; d--) {
      const int half = 1 << d;
      for (int c = 0; c < half; c++) vals[c] = vals[2 * c] * (1.0f - frac[d]) + vals[2 * c + 1] * frac[d];
    }
    out[(long long)i * ncomp + k] = vals[0];
  }
}

// ------------------------------------------------------------------ table-source evaluators at arbitrary points

// interpolate-value (interpolate.clj:87-98) on a packed RGB float table, coordinates in double
__device__ void lookup_rgb(const float *table, const int *shape, int dims, const double *coords, double out[3]) {
  int lo[4], hi[4];
  double frac[4];
  long long stride[4];
  long long s = 3;
  for (int d = dims - 1; d >= 0; d--) {
    stride[d] = s;
    s *= shape[d];
  }
  for (int d = 0; d < dims; d++) {
    double c = fmin(fmax(coords[d], 0.0), (double)(shape[d] - 1));
    double u = floor(c);
    lo[d] = (int)u;
    hi[d] = min(lo[d] + 1, shape[d] - 1);
    frac[d] = c - u;
  }
  for (int ch = 0; ch < 3; ch++) {
    double vals[16];
    const int corners = 1 << dims;
    for (int c = 0; c < corners; c++) {
      long long off = ch;
      for (int d = 0; d < dims; d++) off += (long long)(((c >> (dims - 1 - d)) & 1) ? hi[d] : lo[d]) * stride[d];
      vals[c] = (double)table[off];
    }
    for (int d = dims - 1; d >= 0; d--)
      for (int c = 0; c < (1 << d); c++) vals[c] = vals[2 * c] * (1 - frac[d]) + vals[2 * c + 1] * frac[d];
    out[ch] = vals[0];
  }
}

struct TableSource {   // S source: tab_a, or tab_a + tab_b * phase(g, v.l) (atmosphere_lut.clj:79-84)
  const float *tab_a, *tab_b;
  double phase_g;
  int shape[4];
};

__device__ void s_source(const Planet &pl, const TableSource &src, V3 x, V3 v, V3 l, bool above, double out[3]) {
  double idx[4];
  ray_scatter_forward(pl, src.shape, x, v, l, above, idx);
  lookup_rgb(src.tab_a, src.shape, 4, idx, out);
  if (src.tab_b) {
    double m[3];
    lookup_rgb(src.tab_b, src.shape, 4, idx, m);
    double ph = phase(src.phase_g, dot(v, l));
    for (int ch = 0; ch < 3; ch++) out[ch] = out[ch] + m[ch] * ph;
  }
}

__device__ __forceinline__ void warp_sum3(double acc[3]) {
  for (int ch = 0; ch < 3; ch++)
    for (int o = 16; o > 0; o >>= 1) acc[ch] += __shfl_xor_sync(0xffffffffu, acc[ch], o);
}

// ray-scatter (atmosphere.clj:192-200) with point-scatter = interpolation-table of dJ; one warp per item
__global__ void k_ray_scatter_table_batch(Planet pl, Medium md, int steps, TableSource src, int count, const double *x,
                                          const double *v, const double *l, const int *above, double *out) {
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (i >= count) return;
  V3 xi = load3(x, i), vi = load3(v, i), li = load3(l, i);
  const bool ab = above[i] != 0;
  V3 point = ab ? atmosphere_intersection(pl, xi, vi) : surface_intersection(pl, xi, vi);
  V3 d = point - xi;
  double stepsize = 1.0 / (double)steps;
  double a = stepsize * mag(d);
  double acc[3] = {0.0, 0.0, 0.0};
  for (int k = lane; k < steps; k += 32) {
    V3 p = xi + d * ((0.5 + (double)k) * stepsize);
    double t[3], j[3];
    transmittance_points(pl, md, steps, xi, p, t);
    s_source(pl, src, p, vi, li, ab, j);
    for (int ch = 0; ch < 3; ch++) acc[ch] += (t[ch] * j[ch]) * a;
  }
  warp_sum3(acc);
  if (lane == 0) store3(out, i, acc);
}

// quaternion.clj:166-171 orthogonal, matrix.clj:207-213 oriented-matrix (rows n, o1, o2)
__device__ void oriented_matrix(V3 n, V3 &o1, V3 &o2) {
  double ax = fabs(n.x), ay = fabs(n.y), az = fabs(n.z);
  V3 b = v3(1, 0, 0);
  double best = ax;
  if (ay < best) {
    b = v3(0, 1, 0);
    best = ay;
  }
  if (az < best) b = v3(0, 0, 1);
  V3 c = v3(n.y * b.z - n.z * b.y, n.z * b.x - n.x * b.z, n.x * b.y - n.y * b.x);
  double m = mag(c);
  o1 = v3(c.x / m, c.y / m, c.z / m);
  o2 = v3(n.y * o1.z - n.z * o1.y, n.z * o1.x - n.x * o1.z, n.x * o1.y - n.y * o1.x);
}

// rings of spherical-integral (sphere.clj:70-93), computed on the host: cos/sin theta, ring size, weight
struct Ring {
  double cos_theta, sin_theta, weight;   // weight = factor * 2 pi / ringsteps
  int ringsteps, first;
};

__device__ V3 ring_direction(const Ring &r, int j, V3 n, V3 o1, V3 o2) {
  double phi = 2 * kPi * ((0.5 + (double)j) / (double)r.ringsteps);
  double xx = r.cos_theta, yy = r.sin_theta * cos(phi), zz = r.sin_theta * sin(phi);
  return v3(n.x * xx + o1.x * yy + o2.x * zz, n.y * xx + o1.y * yy + o2.y * zz, n.z * xx + o1.z * yy + o2.z * zz);
}

// point-scatter (atmosphere.clj:203-222) with table sources; one warp per item, lanes over directions
__global__ void k_point_scatter_batch(Planet pl, Medium md, int ray_steps, TableSource src, const float *de,
                                      int e_rows, int e_cols, const Ring *rings, int nrings, int ndirs, int count,
                                      const double *x, const double *v, const double *l, double *out) {
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (i >= count) return;
  V3 xi = load3(x, i), vi = load3(v, i), li = load3(l, i);
  double m = mag(xi);
  V3 n = v3(xi.x / m, xi.y / m, xi.z / m), o1, o2;
  oriented_matrix(n, o1, o2);
  const double hx = height(pl, xi);
  const int e_shape[2] = {e_rows, e_cols};
  double acc[3] = {0.0, 0.0, 0.0};
  for (int d = lane; d < ndirs; d += 32) {
    int r = 0;
    while (r + 1 < nrings && rings[r + 1].first <= d) r++;
    V3 omega = ring_direction(rings[r], d - rings[r].first, n, o1, o2);
    V3 point = ray_extremity(pl, xi, omega);
    bool surface = surface_point(pl, point);
    double s[3];
    s_source(pl, src, xi, omega, li, !surface, s);
    if (surface) {
      double idx[2], e[3], t[3];
      surface_radiance_forward(pl, e_shape, point, li, idx);
      lookup_rgb(de, e_shape, 2, idx, e);
      transmittance_points(pl, md, ray_steps, xi, point, t);
      for (int ch = 0; ch < 3; ch++) s[ch] = s[ch] + t[ch] * ((pl.brightness[ch] / kPi) * e[ch]);
    }
    double mu = dot(vi, omega);
    for (int ch = 0; ch < 3; ch++) {
      double sum = 0.0;
      for (int c = 0; c < md.n; c++) {
        double term = scattering(md, c, ch, hx) * phase(md.g[c], mu);
        sum = (c == 0) ? term : sum + term;
      }
      acc[ch] += (sum * s[ch]) * rings[r].weight;
    }
  }
  warp_sum3(acc);
  if (lane == 0) store3(out, i, acc);
}

// surface-radiance (atmosphere.clj:225-230) with a table source; one warp per item
__global__ void k_surface_radiance_batch(Planet pl, TableSource src, const Ring *rings, int nrings, int ndirs,
                                         int count, const double *x, const double *l, double *out) {
  int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (i >= count) return;
  V3 xi = load3(x, i), li = load3(l, i);
  double m = mag(xi);
  V3 n = v3(xi.x / m, xi.y / m, xi.z / m), o1, o2;
  oriented_matrix(n, o1, o2);
  double acc[3] = {0.0, 0.0, 0.0};
  for (int d = lane; d < ndirs; d += 32) {
    int r = 0;
    while (r + 1 < nrings && rings[r + 1].first <= d) r++;
    V3 omega = ring_direction(rings[r], d - rings[r].first, n, o1, o2);
    double s[3];
    s_source(pl, src, xi, omega, li, true, s);
    double f = dot(omega, n) * rings[r].weight;
    for (int ch = 0; ch < 3; ch++) acc[ch] += s[ch] * f;
  }
  warp_sum3(acc);
  if (lane == 0) store3(out, i, acc);
}

// host: ring table of spherical-integral (sphere.clj:70-93)
std::vector<Ring> make_rings(int theta_steps, int phi_steps, double theta_range, int &ndirs) {
  std::vector<Ring> rings;
  const double delta2 = theta_range / (double)theta_steps / 2;
  ndirs = 0;
  for (int k = 0; k < theta_steps; k++) {
    double theta = theta_range * ((0.5 + (double)k) / (double)theta_steps);
    double factor = cos(theta - delta2) - cos(theta + delta2);
    Ring r;
    r.ringsteps = (int)ceil(sin(theta) * (double)phi_steps);
    r.cos_theta = cos(theta);
    r.sin_theta = sin(theta);
    r.weight = factor * ((2 * kPi) / (double)r.ringsteps);
    r.first = ndirs;
    ndirs += r.ringsteps;
    rings.push_back(r);
  }
  return rings;
}

// atmosphere.clj:42-61 at arbitrary arguments: fn 0 scattering (height -> rgb), 1 extinction (height -> rgb),
// 2 phase (mu -> value, stored in out[3 i])
__global__ void k_medium_batch(Medium m, int fn, int count, const double *__restrict__ arg, double *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  if (fn == 2) {
    out[3 * i] = phase(m.g[0], arg[i]);
    out[3 * i + 1] = out[3 * i + 2] = 0.0;
    return;
  }
  for (int ch = 0; ch < 3; ch++) {
    const double sc = scattering(m, 0, ch, arg[i]);
    out[3 * i + ch] = fn == 0 ? sc : sc / m.quotient[0];
  }
}

struct Bufs {
  std::vector<void *> ptrs;
  ~Bufs() {
    for (void *p : ptrs) cudaFree(p);
  }
  template <typename T>
  int up(const T *host, size_t count, T *&dev) {
    dev = nullptr;
    CUDA_TRY(cudaMalloc((void **)&dev, (count ? count : 1) * sizeof(T)));
    ptrs.push_back(dev);
    if (host && count)
      CUDA_TRY(cudaMemcpyAsync(dev, host, count * sizeof(T), cudaMemcpyHostToDevice, stream()));
    return 0;
  }
  template <typename T>
  int down(const T *dev, size_t count, T *host) {
    if (count) CUDA_TRY(cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, stream()));
    CUDA_TRY(cudaStreamSynchronize(stream()));
    return 0;
  }
};

// subtract the planet centre from an array of points
std::vector<double> centred(const atmlut_planet *planet, const double *x, int count) {
  std::vector<double> r((size_t)count * 3);
  for (int i = 0; i < count; i++)
    for (int k = 0; k < 3; k++) r[3 * i + k] = x[3 * i + k] - planet->centre[k];
  return r;
}

int blocks(long long n, int threads) { return (int)((n + threads - 1) / threads); }

}  // namespace
}  // namespace atm

using namespace atm;

#define CHECK_ARGS(cond, msg) \
  do {                        \
    if (!(cond)) return fail(msg); \
  } while (0)

extern "C" int atmlut_transmittance_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n, int steps,
                                          int count, const double *x, const double *x0, double *out) {
  Params P;
  if (ensure_init() || make_planet_medium(planet, scatter, n, P)) return 1;
  CHECK_ARGS(steps >= 1 && count >= 0 && x && x0 && out, "invalid argument");
  if (count == 0) return 0;
  Bufs b;
  std::vector<double> cx = centred(planet, x, count), cx0 = centred(planet, x0, count);
  double *dx, *dx0, *dout;
  if (b.up(cx.data(), cx.size(), dx) || b.up(cx0.data(), cx0.size(), dx0) || b.up<double>(nullptr, (size_t)count * 3, dout))
    return 1;
  k_transmittance_batch<<<blocks(count, 64), 64, 0, stream()>>>(P.planet, P.medium, steps, count, dx, dx0, dout);
  CUDA_TRY(cudaGetLastError());
  return b.down(dout, (size_t)count * 3, out);
}

extern "C" int atmlut_transmittance_dir_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                              int steps, int count, const double *x, const double *v,
                                              const int *above, double *out) {
  Params P;
  if (ensure_init() || make_planet_medium(planet, scatter, n, P)) return 1;
  CHECK_ARGS(steps >= 1 && count >= 0 && x && v && above && out, "invalid argument");
  if (count == 0) return 0;
  Bufs b;
  std::vector<double> cx = centred(planet, x, count);
  double *dx, *dv, *dout;
  int *dab;
  if (b.up(cx.data(), cx.size(), dx) || b.up(v, (size_t)count * 3, dv) || b.up(above, (size_t)count, dab) ||
      b.up<double>(nullptr, (size_t)count * 3, dout))
    return 1;
  k_transmittance_dir_batch<<<blocks(count, 64), 64, 0, stream()>>>(P.planet, P.medium, steps, count, dx, dv, dab,
                                                                    dout);
  CUDA_TRY(cudaGetLastError());
  return b.down(dout, (size_t)count * 3, out);
}

extern "C" int atmlut_surface_radiance_base_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                                  int steps, const double *intensity, int count, const double *x,
                                                  const double *l, double *out) {
  Params P;
  if (ensure_init() || make_planet_medium(planet, scatter, n, P)) return 1;
  CHECK_ARGS(steps >= 1 && count >= 0 && intensity && x && l && out, "invalid argument");
  if (count == 0) return 0;
  Bufs b;
  std::vector<double> cx = centred(planet, x, count);
  double *dx, *dl, *dout;
  if (b.up(cx.data(), cx.size(), dx) || b.up(l, (size_t)count * 3, dl) || b.up<double>(nullptr, (size_t)count * 3, dout))
    return 1;
  k_surface_radiance_base_batch<<<blocks(count, 64), 64, 0, stream()>>>(
      P.planet, P.medium, steps, v3(intensity[0], intensity[1], intensity[2]), count, dx, dl, dout);
  CUDA_TRY(cudaGetLastError());
  return b.down(dout, (size_t)count * 3, out);
}

extern "C" int atmlut_point_scatter_first_order_batch(const atmlut_planet *planet, const atmlut_scatter *scatter,
                                                      int n, int kind, int component, int steps,
                                                      const double *intensity, int count, const double *x,
                                                      const double *v, const double *l, double *out) {
  Params P;
  if (ensure_init() || make_planet_medium(planet, scatter, n, P)) return 1;
  CHECK_ARGS(steps >= 1 && count >= 0 && intensity && x && v && l && out, "invalid argument");
  CHECK_ARGS(kind >= 0 && kind <= 2, "kind must be 0, 1 or 2");
  CHECK_ARGS(kind == 2 || (component >= 0 && component < n), "component index out of range");
  if (count == 0) return 0;
  Bufs b;
  std::vector<double> cx = centred(planet, x, count);
  double *dx, *dv, *dl, *dout;
  if (b.up(cx.data(), cx.size(), dx) || b.up(v, (size_t)count * 3, dv) || b.up(l, (size_t)count * 3, dl) ||
      b.up<double>(nullptr, (size_t)count * 3, dout))
    return 1;
  k_point_scatter_first_order_batch<<<blocks(count, 64), 64, 0, stream()>>>(
      P.planet, P.medium, kind, component, steps, v3(intensity[0], intensity[1], intensity[2]), count, dx, dv, dl,
      dout);
  CUDA_TRY(cudaGetLastError());
  return b.down(dout, (size_t)count * 3, out);
}

extern "C" int atmlut_ray_scatter_first_order_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                                    int kind, int component, int steps, const double *intensity,
                                                    int count, const double *x, const double *v, const double *l,
                                                    const int *above, double *out) {
  Params P;
  if (ensure_init() || make_planet_medium(planet, scatter, n, P)) return 1;
  CHECK_ARGS(steps >= 1 && count >= 0 && intensity && x && v && l && above && out, "invalid argument");
  CHECK_ARGS(kind >= 0 && kind <= 2, "kind must be 0, 1 or 2");
  CHECK_ARGS(kind == 2 || (component >= 0 && component < n), "component index out of range");
  if (count == 0) return 0;
  Bufs b;
  std::vector<double> cx = centred(planet, x, count);
  double *dx, *dv, *dl, *dout;
  int *dab;
  if (b.up(cx.data(), cx.size(), dx) || b.up(v, (size_t)count * 3, dv) || b.up(l, (size_t)count * 3, dl) ||
      b.up(above, (size_t)count, dab) || b.up<double>(nullptr, (size_t)count * 3, dout))
    return 1;
  k_ray_scatter_first_order_batch<<<blocks((long long)count * 32, 128), 128, 0, stream()>>>(
      P.planet, P.medium, kind, component, steps, v3(intensity[0], intensity[1], intensity[2]), count, dx, dv, dl, dab,
      dout);
  CUDA_TRY(cudaGetLastError());
  return b.down(dout, (size_t)count * 3, out);
}

static int make_table_source(Bufs &b, const Params &P, const int *shape4, const float *ds_a, const float *ds_b,
                             double phase_g, TableSource &src) {
  size_t n = 3;
  for (int i = 0; i < 4; i++) {
    if (shape4[i] < 2) return fail("every table axis needs at least 2 entries");
    src.shape[i] = shape4[i];
    n *= (size_t)shape4[i];
  }
  float *da = nullptr, *db = nullptr;
  if (b.up(ds_a, n, da) || (ds_b && b.up(ds_b, n, db))) return 1;
  src.tab_a = da;
  src.tab_b = db;
  src.phase_g = phase_g;
  (void)P;
  return 0;
}

extern "C" int atmlut_ray_scatter_table_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                              int steps, const int *shape4, const float *dj, int count,
                                              const double *x, const double *v, const double *l, const int *above,
                                              double *out) {
  Params P;
  if (ensure_init() || make_planet_medium(planet, scatter, n, P)) return 1;
  CHECK_ARGS(steps >= 1 && count >= 0 && shape4 && dj && x && v && l && above && out, "invalid argument");
  if (count == 0) return 0;
  Bufs b;
  TableSource src = {};
  if (make_table_source(b, P, shape4, dj, nullptr, 0.0, src)) return 1;
  std::vector<double> cx = centred(planet, x, count);
  double *dx, *dv, *dl, *dout;
  int *dab;
  if (b.up(cx.data(), cx.size(), dx) || b.up(v, (size_t)count * 3, dv) || b.up(l, (size_t)count * 3, dl) ||
      b.up(above, (size_t)count, dab) || b.up<double>(nullptr, (size_t)count * 3, dout))
    return 1;
  k_ray_scatter_table_batch<<<blocks((long long)count * 32, 128), 128, 0, stream()>>>(P.planet, P.medium, steps, src,
                                                                                      count, dx, dv, dl, dab, dout);
  CUDA_TRY(cudaGetLastError());
  return b.down(dout, (size_t)count * 3, out);
}

extern "C" int atmlut_point_scatter_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                          int sphere_steps, int ray_steps, const int *shape4, const float *ds_a,
                                          const float *ds_b, int phase_component, const int *shape_e,
                                          const float *de, int count, const double *x, const double *v,
                                          const double *l, double *out) {
  Params P;
  if (ensure_init() || make_planet_medium(planet, scatter, n, P)) return 1;
  CHECK_ARGS(sphere_steps >= 2 && ray_steps >= 1 && count >= 0 && shape4 && ds_a && shape_e && de && x && v && l && out,
             "invalid argument");
  CHECK_ARGS(!ds_b || (phase_component >= 0 && phase_component < n), "phase component out of range");
  CHECK_ARGS(shape_e[0] >= 2 && shape_e[1] >= 2, "every table axis needs at least 2 entries");
  if (count == 0) return 0;
  Bufs b;
  TableSource src = {};
  if (make_table_source(b, P, shape4, ds_a, ds_b, ds_b ? P.medium.g[phase_component] : 0.0, src)) return 1;
  int ndirs = 0;
  std::vector<Ring> rings = make_rings(sphere_steps >> 1, sphere_steps, kPi, ndirs);   // sphere.clj:102-105
  std::vector<double> cx = centred(planet, x, count);
  double *dx, *dv, *dl, *dout;
  float *dde;
  Ring *drings;
  if (b.up(cx.data(), cx.size(), dx) || b.up(v, (size_t)count * 3, dv) || b.up(l, (size_t)count * 3, dl) ||
      b.up(de, (size_t)shape_e[0] * shape_e[1] * 3, dde) || b.up(rings.data(), rings.size(), drings) ||
      b.up<double>(nullptr, (size_t)count * 3, dout))
    return 1;
  k_point_scatter_batch<<<blocks((long long)count * 32, 128), 128, 0, stream()>>>(
      P.planet, P.medium, ray_steps, src, dde, shape_e[0], shape_e[1], drings, (int)rings.size(), ndirs, count, dx, dv,
      dl, dout);
  CUDA_TRY(cudaGetLastError());
  return b.down(dout, (size_t)count * 3, out);
}

extern "C" int atmlut_surface_radiance_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                             int steps, const int *shape4, const float *ds_a, const float *ds_b,
                                             int phase_component, int count, const double *x, const double *l,
                                             double *out) {
  Params P;
  if (ensure_init() || make_planet_medium(planet, scatter, n, P)) return 1;
  CHECK_ARGS(steps >= 4 && count >= 0 && shape4 && ds_a && x && l && out, "invalid argument (steps must be >= 4)");
  CHECK_ARGS(!ds_b || (phase_component >= 0 && phase_component < n), "phase component out of range");
  if (count == 0) return 0;
  Bufs b;
  TableSource src = {};
  if (make_table_source(b, P, shape4, ds_a, ds_b, ds_b ? P.medium.g[phase_component] : 0.0, src)) return 1;
  int ndirs = 0;
  std::vector<Ring> rings = make_rings(steps >> 2, steps, kPi / 2, ndirs);             // sphere.clj:96-99
  std::vector<double> cx = centred(planet, x, count);
  double *dx, *dl, *dout;
  Ring *drings;
  if (b.up(cx.data(), cx.size(), dx) || b.up(l, (size_t)count * 3, dl) || b.up(rings.data(), rings.size(), drings) ||
      b.up<double>(nullptr, (size_t)count * 3, dout))
    return 1;
  k_surface_radiance_batch<<<blocks((long long)count * 32, 128), 128, 0, stream()>>>(
      P.planet, src, drings, (int)rings.size(), ndirs, count, dx, dl, dout);
  CUDA_TRY(cudaGetLastError());
  return b.down(dout, (size_t)count * 3, out);
}

extern "C" int atmlut_index_forward_batch(const atmlut_planet *planet, int which, const int *shape, int count,
                                          const double *point, const double *direction, const double *light,
                                          const int *above, double *indices) {
  Params P;
  if (ensure_init() || make_planet_medium(planet, nullptr, 0, P)) return 1;
  CHECK_ARGS(which >= 0 && which <= 2 && shape && count >= 0 && point && indices, "invalid argument");
  CHECK_ARGS(which == 1 || (direction && above), "direction and above are required for this space");
  CHECK_ARGS(which == 2 || light, "light is required for this space");
  if (count == 0) return 0;
  const int dims = which == 0 ? 4 : 2;
  int s[4] = {2, 2, 2, 2};
  for (int i = 0; i < dims; i++) s[i] = shape[i];
  Bufs b;
  std::vector<double> cp = centred(planet, point, count);
  double *dp, *dd = nullptr, *dl = nullptr, *didx;
  int *dab = nullptr;
  if (b.up(cp.data(), cp.size(), dp) || (direction && b.up(direction, (size_t)count * 3, dd)) ||
      (light && b.up(light, (size_t)count * 3, dl)) || (above && b.up(above, (size_t)count, dab)) ||
      b.up<double>(nullptr, (size_t)count * dims, didx))
    return 1;
  k_index_forward<<<blocks(count, 64), 64, 0, stream()>>>(P.planet, which, s[0], s[1], s[2], s[3], count, dp, dd, dl,
                                                          dab, didx);
  CUDA_TRY(cudaGetLastError());
  return b.down(didx, (size_t)count * dims, indices);
}

extern "C" int atmlut_index_backward_batch(const atmlut_planet *planet, int which, const int *shape, int count,
                                           const double *indices, double *point, double *direction, double *light,
                                           int *above) {
  Params P;
  if (ensure_init() || make_planet_medium(planet, nullptr, 0, P)) return 1;
  CHECK_ARGS(which >= 0 && which <= 2 && shape && count >= 0 && indices, "invalid argument");
  if (count == 0) return 0;
  const int dims = which == 0 ? 4 : 2;
  int s[4] = {2, 2, 2, 2};
  for (int i = 0; i < dims; i++) s[i] = shape[i];
  Bufs b;
  double *didx, *dp, *dd, *dl;
  int *dab;
  if (b.up(indices, (size_t)count * dims, didx) || b.up<double>(nullptr, (size_t)count * 3, dp) ||
      b.up<double>(nullptr, (size_t)count * 3, dd) || b.up<double>(nullptr, (size_t)count * 3, dl) ||
      b.up<int>(nullptr, (size_t)count, dab))
    return 1;
  k_index_backward<<<blocks(count, 64), 64, 0, stream()>>>(P.planet, which, s[0], s[1], s[2], s[3], count, didx, dp,
                                                           dd, dl, dab);
  CUDA_TRY(cudaGetLastError());
  if (point) {
    if (b.down(dp, (size_t)count * 3, point)) return 1;
    for (int i = 0; i < count; i++)
      for (int k = 0; k < 3; k++) point[3 * i + k] += planet->centre[k];
  }
  if (direction && b.down(dd, (size_t)count * 3, direction)) return 1;
  if (light && b.down(dl, (size_t)count * 3, light)) return 1;
  if (above && b.down(dab, (size_t)count, above)) return 1;
  CUDA_TRY(cudaStreamSynchronize(stream()));
  return 0;
}

extern "C" int atmlut_index_map_batch(const atmlut_planet *planet, int fn, int size, int count, const double *a,
                                      const double *b, const int *flag, double *out, int *out_flag) {
  Params P;
  if (ensure_init() || make_planet_medium(planet, nullptr, 0, P)) return 1;
  CHECK_ARGS(fn >= 0 && fn <= 8 && count >= 0 && a && out, "invalid argument");
  CHECK_ARGS(fn != 0 || flag, "elevation-to-index needs the above-horizon flags");
  CHECK_ARGS(!(fn == 0 || fn == 4 || fn == 6 || fn == 7) || b, "this map needs a second argument");
  CHECK_ARGS(fn == 8 || fn == 4 || fn == 6 || size >= 2 || fn == 5, "size must be at least 2");
  if (count == 0) return 0;
  Bufs bufs;
  std::vector<double> ca;
  const bool point_arg = (fn == 0 || fn == 2 || fn == 4);
  if (point_arg) ca = centred(planet, a, count);
  double *da, *db = nullptr, *dout;
  int *dflag = nullptr, *doflag = nullptr;
  if (bufs.up(point_arg ? ca.data() : a, (size_t)count * 3, da) || (b && bufs.up(b, (size_t)count * 3, db)) ||
      (flag && bufs.up(flag, (size_t)count, dflag)) || bufs.up<double>(nullptr, (size_t)count * 3, dout) ||
      bufs.up<int>(nullptr, (size_t)count, doflag))
    return 1;
  k_index_map<<<blocks(count, 64), 64, 0, stream()>>>(P.planet, fn, size, count, da, db, dflag, dout, doflag);
  CUDA_TRY(cudaGetLastError());
  if (bufs.down(dout, (size_t)count * 3, out)) return 1;
  if (fn == 3)
    for (int i = 0; i < count; i++)
      for (int k = 0; k < 3; k++) out[3 * i + k] += planet->centre[k];
  if (out_flag && bufs.down(doflag, (size_t)count, out_flag)) return 1;
  return 0;
}

extern "C" int atmlut_medium_batch(const atmlut_scatter *component, int fn, int count, const double *arg, double *out) {
  atmlut_planet unit = {{0, 0, 0}, 1.0, 1.0, {0, 0, 0}};
  Params P;
  if (ensure_init() || make_planet_medium(&unit, component, component ? 1 : 0, P)) return 1;
  CHECK_ARGS(component && fn >= 0 && fn <= 2 && count >= 0 && arg && out, "invalid argument");
  if (count == 0) return 0;
  Bufs bufs;
  double *da, *dout;
  if (bufs.up(arg, (size_t)count, da) || bufs.up<double>(nullptr, (size_t)count * 3, dout)) return 1;
  k_medium_batch<<<blocks(count, 64), 64, 0, stream()>>>(P.medium, fn, count, da, dout);
  CUDA_TRY(cudaGetLastError());
  return bufs.down(dout, (size_t)count * 3, out);
}

extern "C" int atmlut_interpolate_batch(const float *table, const int *shape, int dims, int ncomp, int count,
                                        const double *coords, float *out) {
  if (ensure_init()) return 1;
  CHECK_ARGS(table && shape && coords && out && dims >= 1 && dims <= 4 && ncomp >= 1 && count >= 0,
             "invalid argument");
  if (count == 0) return 0;
  int s[4] = {1, 1, 1, 1};
  size_t total = ncomp;
  for (int i = 0; i < dims; i++) {
    CHECK_ARGS(shape[i] >= 1, "invalid shape");
    s[i] = shape[i];
    total *= (size_t)shape[i];
  }
  Bufs b;
  float *dt, *dout;
  double *dc;
  if (b.up(table, total, dt) || b.up(coords, (size_t)count * dims, dc) ||
      b.up<float>(nullptr, (size_t)count * ncomp, dout))
    return 1;
  k_interpolate<<<blocks(count, 64), 64, 0, stream()>>>(dt, dims, s[0], s[1], s[2], s[3], ncomp, count, dc, dout);
  CUDA_TRY(cudaGetLastError());
  return b.down(dout, (size_t)count * ncomp, out);
}
