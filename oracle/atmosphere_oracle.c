/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the sfsim atmosphere-LUT hot path.
 * See atmosphere_oracle.h.  Compile with -ffp-contract=off so every operation is a
 * separately rounded IEEE double operation like the JVM's.
 *
 * file:line citations are relative to the reference tree (wedesoft/sfsim).
 */
#include "atmosphere_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ counters */
static _Thread_local long long tl_esamples, tl_lookups4d, tl_lookups2d;
static long long g_esamples, g_lookups4d, g_lookups2d;

static void counters_flush(void) {
#pragma omp atomic
  g_esamples += tl_esamples;
#pragma omp atomic
  g_lookups4d += tl_lookups4d;
#pragma omp atomic
  g_lookups2d += tl_lookups2d;
  tl_esamples = tl_lookups4d = tl_lookups2d = 0;
}

void orc_counters_reset(void) {
  counters_flush();
  g_esamples = g_lookups4d = g_lookups2d = 0;
}

void orc_counters_get(long long *esamples, long long *lookups4d, long long *lookups2d) {
  counters_flush();
  *esamples = g_esamples;
  *lookups4d = g_lookups4d;
  *lookups2d = g_lookups2d;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------ Vec3 helpers (fastmath.vector) */
static inline double sqr(double x) { return x * x; } /* util.clj:334-337 */
static inline double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline double mag3(const double a[3]) { return sqrt(dot3(a, a)); }
static inline void sub3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[0] - b[0];
  o[1] = a[1] - b[1];
  o[2] = a[2] - b[2];
}
static inline void cross3(const double a[3], const double b[3], double o[3]) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x;
  o[1] = y;
  o[2] = z;
}
static inline void normalize3(const double a[3], double o[3]) {
  double m = mag3(a);
  o[0] = a[0] / m;
  o[1] = a[1] / m;
  o[2] = a[2] / m;
}

/* ------------------------------------------------------------------ util.clj */

/* util.clj:403-416 limit-quot */
double orc_limit_quot(double a, double b, double lower, double upper) {
  if (a == 0.0) return a;
  if (b < 0) return orc_limit_quot(-a, -b, lower, upper);
  if (a < b * upper) {
    if (a > b * lower) return a / b;
    return lower;
  }
  return upper;
}

double orc_limit_quot3(double a, double b, double limit) { return orc_limit_quot(a, b, -limit, limit); }

/* ------------------------------------------------------------------ sphere.clj, ray.clj */

/* sphere.clj:28-31 height */
double orc_height(const orc_planet *planet, const double p[3]) {
  double d[3];
  sub3(p, planet->centre, d);
  return mag3(d) - planet->radius;
}

/* sphere.clj:34-59 ray-sphere-determinant, ray-sphere-intersection */
void orc_ray_sphere_intersection(const double centre[3], double radius, const double origin[3],
                                 const double direction[3], double *distance, double *length) {
  double offset[3];
  sub3(origin, centre, offset);
  double direction_sqr = dot3(direction, direction);
  double discriminant = sqr(dot3(direction, offset)) - direction_sqr * (dot3(offset, offset) - sqr(radius));
  double middle = -(dot3(direction, offset) / direction_sqr);
  if (discriminant > 0) {
    double length2 = sqrt(discriminant) / direction_sqr;
    if (middle < length2) {
      *distance = 0.0;
      *length = fmax(0.0, middle + length2);
    } else {
      *distance = middle - length2;
      *length = 2 * length2;
    }
  } else {
    *distance = fmax(0.0, middle);
    *length = 0.0;
  }
}

/* ray.clj:19-30 integral-ray */
void orc_integral_ray(const double origin[3], const double direction[3], long steps, double distance, orc_vec_fn fun,
                      void *ctx, int dims, double *out) {
  double stepsize = distance / (double)steps;
  double direction_len = mag3(direction);
  double a = stepsize * direction_len;
  double val[8];
  for (int c = 0; c < dims; c++) out[c] = 0.0;
  for (long n = 0; n < steps; n++) {
    double s = (0.5 + (double)n) * stepsize;
    double p[3] = {origin[0] + direction[0] * s, origin[1] + direction[1] * s, origin[2] + direction[2] * s};
    fun(ctx, p, val);
    for (int c = 0; c < dims; c++) out[c] = out[c] + val[c] * a;
  }
}

/* sphere.clj:62-67 integrate-circle */
void orc_integrate_circle(long steps, orc_angle_fn fun, void *ctx, int dims, double *out) {
  double weight = (2 * M_PI) / (double)steps;
  double acc[8], val[8];
  for (long j = 0; j < steps; j++) {
    double phi = 2 * M_PI * ((0.5 + (double)j) / (double)steps);
    fun(ctx, phi, val);
    if (j == 0)
      for (int c = 0; c < dims; c++) acc[c] = val[c];
    else
      for (int c = 0; c < dims; c++) acc[c] = acc[c] + val[c];
  }
  for (int c = 0; c < dims; c++) out[c] = acc[c] * weight;
}

/* quaternion.clj:166-171 orthogonal: b = unit axis with the smallest |n.e_i| (stable sort keeps the first on ties) */
void orc_orthogonal(const double n[3], double out[3]) {
  int best = 0;
  for (int i = 1; i < 3; i++)
    if (fabs(n[i]) < fabs(n[best])) best = i;
  double b[3] = {0, 0, 0};
  b[best] = 1.0;
  double c[3];
  cross3(n, b, c);
  normalize3(c, out);
}

/* matrix.clj:207-213 oriented-matrix: rows n, o1, o2 */
void orc_oriented_matrix(const double n[3], double out[9]) {
  double o1[3], o2[3];
  orc_orthogonal(n, o1);
  cross3(n, o1, o2);
  for (int i = 0; i < 3; i++) {
    out[i] = n[i];
    out[3 + i] = o1[i];
    out[6 + i] = o2[i];
  }
}

typedef struct {
  orc_vec_fn fun;
  void *ctx;
  const double *mat; /* oriented matrix (rows n, o1, o2); we apply its transpose */
  double cos_theta, sin_theta;
} circle_ctx;

static void sample_point(void *vctx, double phi, double *out) {
  circle_ctx *c = (circle_ctx *)vctx;
  double x = c->cos_theta, y = c->sin_theta * cos(phi), z = c->sin_theta * sin(phi);
  const double *m = c->mat;
  /* mulv (transpose mat) (x y z) */
  double d[3] = {m[0] * x + m[3] * y + m[6] * z, m[1] * x + m[4] * y + m[7] * z, m[2] * x + m[5] * y + m[8] * z};
  c->fun(c->ctx, d, out);
}

/* sphere.clj:70-93 spherical-integral */
static void spherical_integral(long theta_steps, long phi_steps, double theta_range, const double normal[3],
                               orc_vec_fn fun, void *ctx, int dims, double *out) {
  double delta2 = theta_range / (double)theta_steps / 2;
  double mat[9];
  orc_oriented_matrix(normal, mat);
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ring[8]; /* (reduce add []) of zero rings: the reference throws; we return 0 */
  for (long k = 0; k < theta_steps; k++) {
    double theta = theta_range * ((0.5 + (double)k) / (double)theta_steps);
    double factor = cos(theta - delta2) - cos(theta + delta2);
    long ringsteps = (long)(int)ceil(sin(theta) * (double)phi_steps);
    circle_ctx c = {fun, ctx, mat, cos(theta), sin(theta)};
    orc_integrate_circle(ringsteps, sample_point, &c, dims, ring);
    if (k == 0)
      for (int i = 0; i < dims; i++) acc[i] = ring[i] * factor;
    else
      for (int i = 0; i < dims; i++) acc[i] = acc[i] + ring[i] * factor;
  }
  for (int i = 0; i < dims; i++) out[i] = acc[i];
}

/* sphere.clj:96-99 integral-half-sphere */
void orc_integral_half_sphere(long steps, const double normal[3], orc_vec_fn fun, void *ctx, int dims, double *out) {
  spherical_integral(steps >> 2, steps, M_PI / 2, normal, fun, ctx, dims, out);
}

/* sphere.clj:102-105 integral-sphere */
void orc_integral_sphere(long steps, const double normal[3], orc_vec_fn fun, void *ctx, int dims, double *out) {
  spherical_integral(steps >> 1, steps, M_PI, normal, fun, ctx, dims, out);
}

/* The direction/weight list spherical-integral walks through (sphere.clj:70-93), for test fixtures. */
long orc_sphere_directions(long theta_steps, long phi_steps, double theta_range, const double normal[3], double *dirs,
                           double *weights) {
  double delta2 = theta_range / (double)theta_steps / 2;
  double m[9];
  orc_oriented_matrix(normal, m);
  long n = 0;
  for (long k = 0; k < theta_steps; k++) {
    double theta = theta_range * ((0.5 + (double)k) / (double)theta_steps);
    double factor = cos(theta - delta2) - cos(theta + delta2);
    long ringsteps = (long)(int)ceil(sin(theta) * (double)phi_steps);
    double weight = (2 * M_PI) / (double)ringsteps;
    for (long j = 0; j < ringsteps; j++, n++) {
      if (!dirs) continue;
      double phi = 2 * M_PI * ((0.5 + (double)j) / (double)ringsteps);
      double x = cos(theta), y = sin(theta) * cos(phi), z = sin(theta) * sin(phi);
      dirs[3 * n + 0] = m[0] * x + m[3] * y + m[6] * z;
      dirs[3 * n + 1] = m[1] * x + m[4] * y + m[7] * z;
      dirs[3 * n + 2] = m[2] * x + m[5] * y + m[8] * z;
      weights[n] = factor * weight;
    }
  }
  return n;
}

/* ------------------------------------------------------------------ atmosphere.clj: medium */

/* atmosphere.clj:42-47 scattering */
void orc_scattering(const orc_scatter *s, double height, double out[3]) {
  double e = exp(-(height / s->scale));
  out[0] = s->base[0] * e;
  out[1] = s->base[1] * e;
  out[2] = s->base[2] * e;
}

/* atmosphere.clj:50-53 extinction */
void orc_extinction(const orc_scatter *s, double height, double out[3]) {
  orc_scattering(s, height, out);
  out[0] = out[0] / s->quotient;
  out[1] = out[1] / s->quotient;
  out[2] = out[2] / s->quotient;
}

/* atmosphere.clj:56-61 phase */
double orc_phase(const orc_scatter *s, double mu) {
  double g = s ? s->g : 0.0;
  double g2 = sqr(g);
  return (3.0 * (1.0 - g2) * (1.0 + sqr(mu))) / (8.0 * M_PI * (2.0 + g2) * pow((1.0 + g2) - 2.0 * g * mu, 1.5));
}

/* atmosphere.clj:70-77 atmosphere-intersection */
void orc_atmosphere_intersection(const orc_planet *planet, const double origin[3], const double direction[3],
                                 double out[3]) {
  double distance, length;
  orc_ray_sphere_intersection(planet->centre, planet->radius + planet->height, origin, direction, &distance, &length);
  double t = distance + length;
  for (int i = 0; i < 3; i++) out[i] = origin[i] + direction[i] * t;
}

/* atmosphere.clj:80-85 surface-intersection */
void orc_surface_intersection(const orc_planet *planet, const double origin[3], const double direction[3],
                              double out[3]) {
  double distance, length;
  orc_ray_sphere_intersection(planet->centre, planet->radius, origin, direction, &distance, &length);
  for (int i = 0; i < 3; i++) out[i] = origin[i] + direction[i] * distance;
}

/* atmosphere.clj:88-92 surface-point? */
int orc_surface_point(const orc_planet *planet, const double p[3]) {
  return 2.0 * orc_height(planet, p) < planet->height;
}

/* atmosphere.clj:95-102 is-above-horizon? (assumes centre = origin, like the reference) */
int orc_is_above_horizon(const orc_planet *planet, const double p[3], const double direction[3]) {
  double norm_point = mag3(p);
  double sin_elevation_radius = dot3(direction, p);
  double horizon_distance_sqr = sqr(norm_point) - sqr(planet->radius);
  return sin_elevation_radius >= 0 || sqr(sin_elevation_radius) <= horizon_distance_sqr;
}

/* atmosphere.clj:105-111 ray-extremity */
void orc_ray_extremity(const orc_planet *planet, const double origin[3], const double direction[3], double out[3]) {
  if (orc_is_above_horizon(planet, origin, direction))
    orc_atmosphere_intersection(planet, origin, direction, out);
  else
    orc_surface_intersection(planet, origin, direction, out);
}

/* ------------------------------------------------------------------ atmosphere.clj: radiative quantities */

typedef struct {
  const orc_planet *planet;
  const orc_scatter *scatter;
  int n;
} ext_ctx;

/* atmosphere.clj:119-124 overall-extinction (1 or 2 components; more are summed in order) */
static void overall_extinction(void *vctx, const double p[3], double *out) {
  ext_ctx *c = (ext_ctx *)vctx;
  tl_esamples++;
  out[0] = out[1] = out[2] = 0.0;
  for (int i = 0; i < c->n; i++) {
    double e[3];
    orc_extinction(&c->scatter[i], orc_height(c->planet, p), e);
    if (i == 0) {
      out[0] = e[0];
      out[1] = e[1];
      out[2] = e[2];
    } else {
      out[0] = out[0] + e[0];
      out[1] = out[1] + e[1];
      out[2] = out[2] + e[2];
    }
  }
}

/* atmosphere.clj:118-125 transmittance, 5-arity */
void orc_transmittance(const orc_planet *planet, const orc_scatter *scatter, int n, long steps, const double x[3],
                       const double x0[3], double out[3]) {
  ext_ctx c = {planet, scatter, n};
  double d[3], integral[3];
  sub3(x0, x, d);
  orc_integral_ray(x, d, steps, 1.0, overall_extinction, &c, 3, integral);
  for (int i = 0; i < 3; i++) out[i] = exp(-integral[i]);
}

/* atmosphere.clj:126-128 transmittance, 6-arity */
void orc_transmittance_dir(const orc_planet *planet, const orc_scatter *scatter, int n, long steps,
                           const double x[3], const double v[3], int above, double out[3]) {
  double x0[3];
  if (above)
    orc_atmosphere_intersection(planet, x, v, x0);
  else
    orc_surface_intersection(planet, x, v, x0);
  orc_transmittance(planet, scatter, n, steps, x, x0, out);
}

/* atmosphere.clj:131-137 surface-radiance-base */
void orc_surface_radiance_base(const orc_planet *planet, const orc_scatter *scatter, int n, long steps,
                               const double intensity[3], const double x[3], const double l[3], double out[3]) {
  double d[3], normal[3], t[3];
  sub3(x, planet->centre, d);
  normalize3(d, normal);
  orc_transmittance_dir(planet, scatter, n, steps, x, l, 1, t);
  double f = fmax(0.0, dot3(normal, l));
  for (int i = 0; i < 3; i++) out[i] = (t[i] * intensity[i]) * f;
}

/* atmosphere.clj:154-160 filtered-sun-light */
static void filtered_sun_light(const orc_planet *planet, const orc_scatter *scatter, int n, long steps,
                               const double x[3], const double l[3], const double intensity[3], double out[3]) {
  if (orc_is_above_horizon(planet, x, l)) {
    double t[3];
    orc_transmittance_dir(planet, scatter, n, steps, x, l, 1, t);
    for (int i = 0; i < 3; i++) out[i] = intensity[i] * t[i];
  } else {
    out[0] = out[1] = out[2] = 0.0;
  }
}

/* atmosphere.clj:140-151 in-scattering-component, overall-in-scattering */
static void overall_in_scattering(const orc_planet *planet, const orc_scatter *components, int n, const double x[3],
                                  const double v[3], const double l[3], double out[3]) {
  for (int i = 0; i < n; i++) {
    double s[3];
    orc_scattering(&components[i], orc_height(planet, x), s);
    double ph = orc_phase(&components[i], dot3(v, l));
    for (int c = 0; c < 3; c++) {
      double term = s[c] * ph;
      out[c] = (i == 0) ? term : out[c] + term;
    }
  }
}

/* atmosphere.clj:162-167 overall-point-scatter */
static void overall_point_scatter(const orc_planet *planet, const orc_scatter *scatter, int n,
                                  const orc_scatter *components, int ncomp, long steps, const double intensity[3],
                                  const double x[3], const double v[3], const double l[3], double out[3]) {
  double a[3], b[3];
  overall_in_scattering(planet, components, ncomp, x, v, l, a);
  filtered_sun_light(planet, scatter, n, steps, x, l, intensity, b);
  for (int i = 0; i < 3; i++) out[i] = a[i] * b[i];
}

/* atmosphere.clj:170-174 point-scatter-component */
void orc_point_scatter_component(const orc_planet *planet, const orc_scatter *scatter, int n,
                                 const orc_scatter *component, long steps, const double intensity[3],
                                 const double x[3], const double v[3], const double l[3], int above, double out[3]) {
  (void)above;
  overall_point_scatter(planet, scatter, n, component, 1, steps, intensity, x, v, l, out);
}

/* atmosphere.clj:177-182 strength-component */
void orc_strength_component(const orc_planet *planet, const orc_scatter *scatter, int n,
                            const orc_scatter *component, long steps, const double intensity[3], const double x[3],
                            const double v[3], const double l[3], int above, double out[3]) {
  (void)above;
  (void)v;
  double s[3], f[3];
  orc_scattering(component, orc_height(planet, x), s);
  filtered_sun_light(planet, scatter, n, steps, x, l, intensity, f);
  for (int i = 0; i < 3; i++) out[i] = s[i] * f[i];
}

/* atmosphere.clj:185-189 point-scatter-base */
void orc_point_scatter_base(const orc_planet *planet, const orc_scatter *scatter, int n, long steps,
                            const double intensity[3], const double x[3], const double v[3], const double l[3],
                            int above, double out[3]) {
  (void)above;
  overall_point_scatter(planet, scatter, n, scatter, n, steps, intensity, x, v, l, out);
}

typedef struct {
  const orc_planet *planet;
  const orc_scatter *scatter;
  int n;
  long steps;
  orc_point_fn point_scatter;
  void *ctx;
  const double *x, *v, *l;
  int above;
} rs_ctx;

static void ray_scatter_integrand(void *vctx, const double p[3], double *out) {
  rs_ctx *c = (rs_ctx *)vctx;
  double t[3], j[3];
  orc_transmittance(c->planet, c->scatter, c->n, c->steps, c->x, p, t);
  c->point_scatter(c->ctx, p, c->v, c->l, c->above, j);
  for (int i = 0; i < 3; i++) out[i] = t[i] * j[i];
}

/* atmosphere.clj:192-200 ray-scatter */
void orc_ray_scatter(const orc_planet *planet, const orc_scatter *scatter, int n, long steps,
                     orc_point_fn point_scatter, void *ctx, const double x[3], const double v[3], const double l[3],
                     int above, double out[3]) {
  double point[3], d[3];
  if (above)
    orc_atmosphere_intersection(planet, x, v, point);
  else
    orc_surface_intersection(planet, x, v, point);
  sub3(point, x, d);
  rs_ctx c = {planet, scatter, n, steps, point_scatter, ctx, x, v, l, above};
  orc_integral_ray(x, d, steps, 1.0, ray_scatter_integrand, &c, 3, out);
}

typedef struct {
  const orc_planet *planet;
  const orc_scatter *scatter;
  int n;
  orc_point_fn ray_scatter;
  void *rs_ctx;
  orc_surface_fn surface_radiance;
  void *sr_ctx;
  long ray_steps;
  const double *x, *v, *l;
} ps_ctx;

/* Test hooks: the analogue of the reference tests' with-redefs (t_atmosphere.clj:330-361 rebinds atmosphere/phase,
 * atmosphere/ray-extremity and atmosphere/transmittance around point-scatter).  NULL = the real function.  Only the
 * point-scatter integrand consults them, and only the known-answer tests set them. */
static orc_test_hooks g_hooks = {0, 0, 0};

void orc_set_test_hooks(const orc_test_hooks *hooks) {
  if (hooks)
    g_hooks = *hooks;
  else
    memset(&g_hooks, 0, sizeof g_hooks);
}

static void in_scatter_from_direction(void *vctx, const double omega[3], double *out) {
  ps_ctx *c = (ps_ctx *)vctx;
  double point[3], overall[3], rs[3], extra[3] = {0, 0, 0};
  if (g_hooks.ray_extremity)
    g_hooks.ray_extremity(c->planet, c->x, omega, point);
  else
    orc_ray_extremity(c->planet, c->x, omega, point);
  int surface = orc_surface_point(c->planet, point);
  if (g_hooks.phase) {
    /* overall-in-scattering (atmosphere.clj:147-151) with the rebound phase */
    for (int i = 0; i < c->n; i++) {
      double s[3];
      orc_scattering(&c->scatter[i], orc_height(c->planet, c->x), s);
      double ph = g_hooks.phase(&c->scatter[i], dot3(c->v, omega));
      for (int k = 0; k < 3; k++) {
        double term = s[k] * ph;
        overall[k] = (i == 0) ? term : overall[k] + term;
      }
    }
  } else {
    overall_in_scattering(c->planet, c->scatter, c->n, c->x, c->v, omega, overall);
  }
  c->ray_scatter(c->rs_ctx, c->x, omega, c->l, !surface, rs);
  if (surface) {
    double e[3], t[3];
    c->surface_radiance(c->sr_ctx, point, c->l, e);
    if (g_hooks.transmittance)
      g_hooks.transmittance(c->planet, c->scatter, c->n, c->ray_steps, c->x, point, t);
    else
      orc_transmittance(c->planet, c->scatter, c->n, c->ray_steps, c->x, point, t);
    for (int i = 0; i < 3; i++) {
      double surface_brightness = (c->planet->brightness[i] / M_PI) * e[i];
      extra[i] = t[i] * surface_brightness;
    }
  }
  for (int i = 0; i < 3; i++) out[i] = overall[i] * (rs[i] + extra[i]);
}

/* atmosphere.clj:203-222 point-scatter */
void orc_point_scatter(const orc_planet *planet, const orc_scatter *scatter, int n, orc_point_fn ray_scatter,
                       void *rsc, orc_surface_fn surface_radiance, void *src, const double intensity[3],
                       long sphere_steps, long ray_steps, const double x[3], const double v[3], const double l[3],
                       int above, double out[3]) {
  (void)intensity;
  (void)above;
  double d[3], normal[3];
  sub3(x, planet->centre, d);
  normalize3(d, normal);
  ps_ctx c = {planet, scatter, n, ray_scatter, rsc, surface_radiance, src, ray_steps, x, v, l};
  orc_integral_sphere(sphere_steps, normal, in_scatter_from_direction, &c, 3, out);
}

/* the integrand of point-scatter for ONE direction omega (the `fun` the reference hands to integral-sphere,
 * atmosphere.clj:208-222): what t_atmosphere.clj:330-361 probes through its integral-sphere mock */
void orc_in_scatter_from_direction(const orc_planet *planet, const orc_scatter *scatter, int n,
                                   orc_point_fn ray_scatter, void *rsc, orc_surface_fn surface_radiance, void *src,
                                   long ray_steps, const double x[3], const double v[3], const double l[3],
                                   const double omega[3], double out[3]) {
  ps_ctx c = {planet, scatter, n, ray_scatter, rsc, surface_radiance, src, ray_steps, x, v, l};
  in_scatter_from_direction(&c, omega, out);
}

typedef struct {
  orc_point_fn ray_scatter;
  void *rs_ctx;
  const double *x, *l, *normal;
} sr_ctx;

static void surface_radiance_integrand(void *vctx, const double omega[3], double *out) {
  sr_ctx *c = (sr_ctx *)vctx;
  double rs[3];
  c->ray_scatter(c->rs_ctx, c->x, omega, c->l, 1, rs);
  double f = dot3(omega, c->normal);
  for (int i = 0; i < 3; i++) out[i] = rs[i] * f;
}

/* atmosphere.clj:225-230 surface-radiance */
void orc_surface_radiance(const orc_planet *planet, orc_point_fn ray_scatter, void *rsc, long steps,
                          const double x[3], const double l[3], double out[3]) {
  double d[3], normal[3];
  sub3(x, planet->centre, d);
  normalize3(d, normal);
  sr_ctx c = {ray_scatter, rsc, x, l, normal};
  orc_integral_half_sphere(steps, normal, surface_radiance_integrand, &c, 3, out);
}

/* ------------------------------------------------------------------ atmosphere.clj: index maps */

/* atmosphere.clj:233-236 horizon-distance */
double orc_horizon_distance(const orc_planet *planet, double radius) {
  return sqrt(fmax(0.0, sqr(radius) - sqr(planet->radius)));
}

/* atmosphere.clj:239-253 elevation-to-index */
double orc_elevation_to_index(const orc_planet *planet, long size, const double point[3], const double direction[3],
                              int above) {
  double radius = mag3(point);
  double ground_radius = planet->radius;
  double top_radius = ground_radius + planet->height;
  double sin_elevation = dot3(point, direction) / radius;
  double rho = orc_horizon_distance(planet, radius);
  double Delta = sqr(radius * sin_elevation) - sqr(rho);
  double H = sqrt(sqr(top_radius) - sqr(ground_radius));
  double q;
  if (above)
    q = 0.5 - orc_limit_quot(radius * sin_elevation - sqrt(fmax(0.0, Delta + sqr(H))), rho + rho + 2 * H, -0.5, 0.0);
  else
    q = 0.5 + orc_limit_quot(radius * sin_elevation + sqrt(fmax(0.0, Delta)), rho + rho, -0.5, 0.0);
  return (double)(size - 1) * q;
}

/* atmosphere.clj:256-270 index-to-elevation */
void orc_index_to_elevation(const orc_planet *planet, long size, double radius, double index, double dir[3],
                            int *above) {
  double ground_radius = planet->radius;
  double top_radius = ground_radius + planet->height;
  double horizon_dist = orc_horizon_distance(planet, radius);
  double H = sqrt(sqr(top_radius) - sqr(ground_radius));
  double scaled_index = index / (double)(size - 1);
  double sin_elevation;
  if (scaled_index < 0.5 || (scaled_index == 0.5 && 2.0 * radius < ground_radius + top_radius)) {
    double ground_dist = horizon_dist * (1 - 2 * scaled_index);
    sin_elevation =
        orc_limit_quot3(sqr(ground_radius) - sqr(radius) - sqr(ground_dist), 2 * radius * ground_dist, 1.0);
    *above = 0;
  } else {
    double sky_dist = (horizon_dist + H) * (2 * scaled_index - 1);
    sin_elevation = orc_limit_quot(sqr(top_radius) - sqr(radius) - sqr(sky_dist), 2 * radius * sky_dist, -1.0, 1.0);
    *above = 1;
  }
  dir[0] = sin_elevation;
  dir[1] = sqrt(1 - sqr(sin_elevation));
  dir[2] = 0.0;
}

/* atmosphere.clj:273-278 height-to-index */
double orc_height_to_index(const orc_planet *planet, long size, const double point[3]) {
  double radius = planet->radius;
  double max_height = planet->height;
  return (double)(size - 1) * (orc_horizon_distance(planet, mag3(point)) / orc_horizon_distance(planet, radius + max_height));
}

/* atmosphere.clj:281-288 index-to-height */
void orc_index_to_height(const orc_planet *planet, long size, double index, double point[3]) {
  double radius = planet->radius;
  double max_height = planet->height;
  double max_horizon = sqrt(sqr(radius + max_height) - sqr(radius));
  double horizon_dist = (index / (double)(size - 1)) * max_horizon;
  point[0] = sqrt(sqr(radius) + sqr(horizon_dist));
  point[1] = 0.0;
  point[2] = 0.0;
}

/* atmosphere.clj:322-326 sun-elevation-to-index */
double orc_sun_elevation_to_index(long size, const double point[3], const double l[3]) {
  double sin_elevation = dot3(point, l) / mag3(point);
  return (double)(size - 1) * fmax(0.0, (1 - exp(0 - 3 * sin_elevation - 0.6)) / (1 - exp(-3.6)));
}

/* atmosphere.clj:329-332 index-to-sin-sun-elevation */
double orc_index_to_sin_sun_elevation(long size, double index) {
  return (log(1 - (index / (double)(size - 1)) * (1 - exp(-3.6))) + 0.6) / -3;
}

/* atmosphere.clj:368-372 sun-angle-to-index */
double orc_sun_angle_to_index(long size, const double direction[3], const double l[3]) {
  return (double)(size - 1) * ((1 + dot3(direction, l)) / 2);
}

/* atmosphere.clj:375-384 index-to-sun-direction */
void orc_index_to_sun_direction(long size, const double direction[3], double sin_sun_elevation, double index,
                                double out[3]) {
  double dot_view_sun = 2.0 * (index / (double)(size - 1)) - 1.0;
  double max_sun_1 = sqrt(fmax(0.0, 1.0 - sqr(sin_sun_elevation)));
  double sun_1 = orc_limit_quot3(dot_view_sun - direction[0] * sin_sun_elevation, direction[1], max_sun_1);
  double sun_2 = sqrt(fmax(0.0, 1.0 - sqr(sun_1) - sqr(sin_sun_elevation)));
  out[0] = sin_sun_elevation;
  out[1] = sun_1;
  out[2] = sun_2;
}

/* atmosphere.clj:291-299 transmittance-forward */
void orc_transmittance_forward(const orc_planet *planet, const long shape[2], const double point[3],
                               const double direction[3], int above, double idx[2]) {
  idx[0] = orc_height_to_index(planet, shape[0], point);
  idx[1] = orc_elevation_to_index(planet, shape[1], point, direction, above);
}

/* atmosphere.clj:302-310 transmittance-backward */
void orc_transmittance_backward(const orc_planet *planet, const long shape[2], double hi, double ei, double point[3],
                                double direction[3], int *above) {
  orc_index_to_height(planet, shape[0], hi, point);
  orc_index_to_elevation(planet, shape[1], point[0], ei, direction, above);
}

/* atmosphere.clj:335-343 surface-radiance-forward */
void orc_surface_radiance_forward(const orc_planet *planet, const long shape[2], const double point[3],
                                  const double l[3], double idx[2]) {
  idx[0] = orc_height_to_index(planet, shape[0], point);
  idx[1] = orc_sun_elevation_to_index(shape[1], point, l);
}

/* atmosphere.clj:346-356 surface-radiance-backward */
void orc_surface_radiance_backward(const orc_planet *planet, const long shape[2], double hi, double si,
                                   double point[3], double l[3]) {
  orc_index_to_height(planet, shape[0], hi, point);
  double sin_sun_elevation = orc_index_to_sin_sun_elevation(shape[1], si);
  double cos_sun_elevation = sqrt(fmax(0.0, 1 - sqr(sin_sun_elevation)));
  l[0] = sin_sun_elevation;
  l[1] = cos_sun_elevation;
  l[2] = 0.0;
}

/* atmosphere.clj:387-398 ray-scatter-forward */
void orc_ray_scatter_forward(const orc_planet *planet, const long shape[4], const double point[3],
                             const double direction[3], const double l[3], int above, double idx[4]) {
  idx[0] = orc_height_to_index(planet, shape[0], point);
  idx[1] = orc_elevation_to_index(planet, shape[1], point, direction, above);
  idx[2] = orc_sun_elevation_to_index(shape[2], point, l);
  idx[3] = orc_sun_angle_to_index(shape[3], direction, l);
}

/* atmosphere.clj:401-412 ray-scatter-backward */
void orc_ray_scatter_backward(const orc_planet *planet, const long shape[4], double hi, double ei, double si,
                              double ai, double point[3], double direction[3], double l[3], int *above) {
  orc_index_to_height(planet, shape[0], hi, point);
  orc_index_to_elevation(planet, shape[1], point[0], ei, direction, above);
  double sin_sun_elevation = orc_index_to_sin_sun_elevation(shape[2], si);
  orc_index_to_sun_direction(shape[3], direction, sin_sun_elevation, ai, l);
}

/* ------------------------------------------------------------------ interpolate.clj */

/* interpolate.clj:75-98 clip, mix, interpolate-value (first axis outermost) */
static void interpolate_value(const double *table, const long *shape, int dims, int ncomp, const double *coords,
                              double *out) {
  if (dims == 0) {
    for (int c = 0; c < ncomp; c++) out[c] = table[c];
    return;
  }
  long size = shape[0];
  long stride = ncomp;
  for (int d = 1; d < dims; d++) stride *= shape[d];
  double i = fmin(fmax(coords[0], 0.0), (double)(size - 1));
  double u = floor(i);
  double v = fmin(fmax(u + 1, 0.0), (double)(size - 1));
  double s = i - u;
  double a[8], b[8];
  interpolate_value(table + (long)u * stride, shape + 1, dims - 1, ncomp, coords + 1, a);
  interpolate_value(table + (long)v * stride, shape + 1, dims - 1, ncomp, coords + 1, b);
  for (int c = 0; c < ncomp; c++) out[c] = a[c] * (1 - s) + b[c] * s;
}

void orc_interpolate(const double *table, const long *shape, int dims, int ncomp, const double *coords, double *out) {
  if (dims == 4)
    tl_lookups4d++;
  else if (dims == 2)
    tl_lookups2d++;
  interpolate_value(table, shape, dims, ncomp, coords, out);
}

/* ------------------------------------------------------------------ output packing */

/* matrix.clj:130-134 pack-matrices: doubles -> float32 */
void orc_pack_floats(const double *in, long count, float *out) {
  for (long i = 0; i < count; i++) out[i] = (float)in[i];
}

/* image.clj:299-312 convert-4d-to-2d: [d c b a] -> rows y = d_i*b + b_i, cols x = c_i*a + a_i */
void orc_convert_4d_to_2d(const double *in, const long shape[4], int ncomp, double *out) {
  long d = shape[0], c = shape[1], b = shape[2], a = shape[3];
  long h = d * b, w = c * a;
  for (long y = 0; y < h; y++)
    for (long x = 0; x < w; x++) {
      long src = (((y / b) * c + (x / a)) * b + (y % b)) * a + (x % a);
      for (int k = 0; k < ncomp; k++) out[(y * w + x) * ncomp + k] = in[src * ncomp + k];
    }
}

/* util.clj:227-240 floats->bytes / spit-floats: headerless little-endian float32 */
int orc_spit_floats(const char *path, const float *data, long count) {
  FILE *f = fopen(path, "wb");
  if (!f) return -1;
  for (long i = 0; i < count; i++) {
    unsigned int bits;
    memcpy(&bits, &data[i], 4);
    unsigned char b[4] = {(unsigned char)(bits & 255), (unsigned char)((bits >> 8) & 255),
                          (unsigned char)((bits >> 16) & 255), (unsigned char)((bits >> 24) & 255)};
    if (fwrite(b, 1, 4, f) != 4) {
      fclose(f);
      return -1;
    }
  }
  return fclose(f);
}

/* util.clj:188-203 bytes->floats / slurp-floats */
long orc_slurp_floats(const char *path, float *data, long max_count) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  long n = 0;
  unsigned char b[4];
  while (n < max_count && fread(b, 1, 4, f) == 4) {
    unsigned int bits = (unsigned)b[0] | ((unsigned)b[1] << 8) | ((unsigned)b[2] << 16) | ((unsigned)b[3] << 24);
    memcpy(&data[n++], &bits, 4);
  }
  fclose(f);
  return n;
}

/* ------------------------------------------------------------------ atmosphere_lut.clj: table builders */

static long prod(const long *shape, int dims) {
  long n = 1;
  for (int i = 0; i < dims; i++) n *= shape[i];
  return n;
}

/* interpolate.clj:54-72 make-lookup-table: the integer grid index, as doubles, goes through `backward` */
static void unravel(long flat, const long *shape, int dims, double *idx) {
  for (int d = dims - 1; d >= 0; d--) {
    idx[d] = (double)(flat % shape[d]);
    flat /= shape[d];
  }
}

/* atmosphere_lut.clj:74 T = (interpolate-function transmittance-planet transmittance-space-planet) */
void orc_table_transmittance(const orc_planet *planet, const orc_scatter *scatter, int n, const orc_config *cfg,
                             const long *indices, long count, double *out) {
  if (!indices) count = prod(cfg->shape_t, 2);
#pragma omp parallel
  {
#pragma omp for schedule(dynamic, 16)
    for (long i = 0; i < count; i++) {
      double idx[2], x[3], v[3];
      int above;
      unravel(indices ? indices[i] : i, cfg->shape_t, 2, idx);
      orc_transmittance_backward(planet, cfg->shape_t, idx[0], idx[1], x, v, &above);
      orc_transmittance_dir(planet, scatter, n, cfg->ray_steps, x, v, above, out + 3 * i);
    }
    counters_flush();
  }
}

/* atmosphere_lut.clj:75 dE = (interpolate-function surface-radiance-base-planet surface-radiance-space-planet) */
void orc_table_surface_radiance_base(const orc_planet *planet, const orc_scatter *scatter, int n,
                                     const orc_config *cfg, const long *indices, long count, double *out) {
  if (!indices) count = prod(cfg->shape_e, 2);
#pragma omp parallel
  {
#pragma omp for schedule(dynamic, 16)
    for (long i = 0; i < count; i++) {
      double idx[2], x[3], l[3];
      unravel(indices ? indices[i] : i, cfg->shape_e, 2, idx);
      orc_surface_radiance_backward(planet, cfg->shape_e, idx[0], idx[1], x, l);
      orc_surface_radiance_base(planet, scatter, n, cfg->ray_steps, cfg->intensity, x, l, out + 3 * i);
    }
    counters_flush();
  }
}

typedef struct {
  const orc_planet *planet;
  const orc_scatter *scatter;
  int n;
  const orc_scatter *component;
  long steps;
  const double *intensity;
  int strength;
} first_order_ctx;

static void first_order_point(void *vctx, const double p[3], const double v[3], const double l[3], int above,
                              double out[3]) {
  first_order_ctx *c = (first_order_ctx *)vctx;
  if (c->strength)
    orc_strength_component(c->planet, c->scatter, c->n, c->component, c->steps, c->intensity, p, v, l, above, out);
  else
    orc_point_scatter_component(c->planet, c->scatter, c->n, c->component, c->steps, c->intensity, p, v, l, above, out);
}

/* atmosphere_lut.clj:68-72,77-78 first-order-rayleigh / first-order-mie-strength */
void orc_table_first_order(const orc_planet *planet, const orc_scatter *scatter, int n, const orc_config *cfg,
                           const orc_scatter *component, int strength, const long *indices, long count, double *out) {
  if (!indices) count = prod(cfg->shape4, 4);
  first_order_ctx c = {planet, scatter, n, component, cfg->ray_steps, cfg->intensity, strength};
#pragma omp parallel
  {
#pragma omp for schedule(dynamic, 4)
    for (long i = 0; i < count; i++) {
      double idx[4], x[3], v[3], l[3];
      int above;
      unravel(indices ? indices[i] : i, cfg->shape4, 4, idx);
      orc_ray_scatter_backward(planet, cfg->shape4, idx[0], idx[1], idx[2], idx[3], x, v, l, &above);
      orc_ray_scatter(planet, scatter, n, cfg->ray_steps, first_order_point, &c, x, v, l, above, out + 3 * i);
    }
    counters_flush();
  }
}

typedef struct {
  const orc_planet *planet;
  const orc_config *cfg;
  const orc_s_source *src;
} s_lookup_ctx;

/* interpolation-table over ray-scatter-space (interpolate.clj:101-104), plus the iteration-1 closure
 * atmosphere_lut.clj:79-84: first-order-rayleigh + first-order-mie-strength * phase(mie, v.l) */
static void s_lookup(void *vctx, const double p[3], const double v[3], const double l[3], int above, double out[3]) {
  s_lookup_ctx *c = (s_lookup_ctx *)vctx;
  double idx[4];
  orc_ray_scatter_forward(c->planet, c->cfg->shape4, p, v, l, above, idx);
  orc_interpolate(c->src->tab_a, c->cfg->shape4, 4, 3, idx, out);
  if (c->src->kind == 1) {
    double m[3];
    orc_ray_scatter_forward(c->planet, c->cfg->shape4, p, v, l, above, idx);
    orc_interpolate(c->src->tab_b, c->cfg->shape4, 4, 3, idx, m);
    double ph = orc_phase(&c->src->phase_component, dot3(v, l));
    for (int i = 0; i < 3; i++) out[i] = out[i] + m[i] * ph;
  }
}

typedef struct {
  const orc_planet *planet;
  const orc_config *cfg;
  const double *tab;
} e_lookup_ctx;

static void e_lookup(void *vctx, const double p[3], const double l[3], double out[3]) {
  e_lookup_ctx *c = (e_lookup_ctx *)vctx;
  double idx[2];
  orc_surface_radiance_forward(c->planet, c->cfg->shape_e, p, l, idx);
  orc_interpolate(c->tab, c->cfg->shape_e, 2, 3, idx, out);
}

/* atmosphere_lut.clj:88,90 dJ = (interpolate-function point-scatter-planet point-scatter-space-planet) */
void orc_table_point_scatter(const orc_planet *planet, const orc_scatter *scatter, int n, const orc_config *cfg,
                             const orc_s_source *ds, const double *de, const long *indices, long count, double *out) {
  if (!indices) count = prod(cfg->shape4, 4);
  s_lookup_ctx sc = {planet, cfg, ds};
  e_lookup_ctx ec = {planet, cfg, de};
#pragma omp parallel
  {
#pragma omp for schedule(dynamic, 4)
    for (long i = 0; i < count; i++) {
      double idx[4], x[3], v[3], l[3];
      int above;
      unravel(indices ? indices[i] : i, cfg->shape4, 4, idx);
      orc_ray_scatter_backward(planet, cfg->shape4, idx[0], idx[1], idx[2], idx[3], x, v, l, &above);
      orc_point_scatter(planet, scatter, n, s_lookup, &sc, e_lookup, &ec, cfg->intensity, cfg->sphere_steps,
                        cfg->ray_steps, x, v, l, above, out + 3 * i);
    }
    counters_flush();
  }
}

/* atmosphere_lut.clj:89,92 dE = (interpolate-function (partial surface-radiance earth @dS ray-steps) ...) */
void orc_table_surface_radiance(const orc_planet *planet, const orc_config *cfg, const orc_s_source *ds,
                                const long *indices, long count, double *out) {
  if (!indices) count = prod(cfg->shape_e, 2);
  s_lookup_ctx sc = {planet, cfg, ds};
#pragma omp parallel
  {
#pragma omp for schedule(dynamic, 1)
    for (long i = 0; i < count; i++) {
      double idx[2], x[3], l[3];
      unravel(indices ? indices[i] : i, cfg->shape_e, 2, idx);
      orc_surface_radiance_backward(planet, cfg->shape_e, idx[0], idx[1], x, l);
      orc_surface_radiance(planet, s_lookup, &sc, cfg->ray_steps, x, l, out + 3 * i);
    }
    counters_flush();
  }
}

/* atmosphere_lut.clj:91,93 dS = (interpolate-function (partial ray-scatter earth scatter ray-steps dJ) ...) */
void orc_table_ray_scatter(const orc_planet *planet, const orc_scatter *scatter, int n, const orc_config *cfg,
                           const double *dj, const long *indices, long count, double *out) {
  if (!indices) count = prod(cfg->shape4, 4);
  orc_s_source src = {0, dj, NULL, {{0, 0, 0}, 1, 0, 1}};
  s_lookup_ctx sc = {planet, cfg, &src};
#pragma omp parallel
  {
#pragma omp for schedule(dynamic, 4)
    for (long i = 0; i < count; i++) {
      double idx[4], x[3], v[3], l[3];
      int above;
      unravel(indices ? indices[i] : i, cfg->shape4, 4, idx);
      orc_ray_scatter_backward(planet, cfg->shape4, idx[0], idx[1], idx[2], idx[3], x, v, l, &above);
      orc_ray_scatter(planet, scatter, n, cfg->ray_steps, s_lookup, &sc, x, v, l, above, out + 3 * i);
    }
    counters_flush();
  }
}

/* atmosphere_lut.clj:94-101: re-tabulation of closures: new[i] = sum_k lookup(tab_k, forward(backward(i))) */
void orc_table_resample_sum_4d(const orc_planet *planet, const orc_config *cfg, const double *const *tabs, int ntabs,
                               const long *indices, long count, double *out) {
  if (!indices) count = prod(cfg->shape4, 4);
#pragma omp parallel
  {
#pragma omp for schedule(static)
    for (long i = 0; i < count; i++) {
      double idx[4], x[3], v[3], l[3], acc[3] = {0, 0, 0};
      int above, first = 1;
      unravel(indices ? indices[i] : i, cfg->shape4, 4, idx);
      orc_ray_scatter_backward(planet, cfg->shape4, idx[0], idx[1], idx[2], idx[3], x, v, l, &above);
      for (int k = 0; k < ntabs; k++) {
        double val[3] = {0, 0, 0}, f[4];
        if (tabs[k]) {
          orc_ray_scatter_forward(planet, cfg->shape4, x, v, l, above, f);
          orc_interpolate(tabs[k], cfg->shape4, 4, 3, f, val);
        }
        for (int c = 0; c < 3; c++) acc[c] = first ? val[c] : acc[c] + val[c];
        first = 0;
      }
      for (int c = 0; c < 3; c++) out[3 * i + c] = acc[c];
    }
    counters_flush();
  }
}

void orc_table_resample_sum_e(const orc_planet *planet, const orc_config *cfg, const double *const *tabs, int ntabs,
                              const long *indices, long count, double *out) {
  if (!indices) count = prod(cfg->shape_e, 2);
  for (long i = 0; i < count; i++) {
    double idx[2], x[3], l[3], acc[3] = {0, 0, 0};
    int first = 1;
    unravel(indices ? indices[i] : i, cfg->shape_e, 2, idx);
    orc_surface_radiance_backward(planet, cfg->shape_e, idx[0], idx[1], x, l);
    for (int k = 0; k < ntabs; k++) {
      double val[3] = {0, 0, 0}, f[2];
      if (tabs[k]) {
        orc_surface_radiance_forward(planet, cfg->shape_e, x, l, f);
        orc_interpolate(tabs[k], cfg->shape_e, 2, 3, f, val);
      }
      for (int c = 0; c < 3; c++) acc[c] = first ? val[c] : acc[c] + val[c];
      first = 0;
    }
    for (int c = 0; c < 3; c++) out[3 * i + c] = acc[c];
  }
  counters_flush();
}

void orc_table_resample_sum_t(const orc_planet *planet, const orc_config *cfg, const double *const *tabs, int ntabs,
                              const long *indices, long count, double *out) {
  if (!indices) count = prod(cfg->shape_t, 2);
  for (long i = 0; i < count; i++) {
    double idx[2], x[3], v[3], acc[3] = {0, 0, 0};
    int above, first = 1;
    unravel(indices ? indices[i] : i, cfg->shape_t, 2, idx);
    orc_transmittance_backward(planet, cfg->shape_t, idx[0], idx[1], x, v, &above);
    for (int k = 0; k < ntabs; k++) {
      double val[3] = {0, 0, 0}, f[2];
      if (tabs[k]) {
        orc_transmittance_forward(planet, cfg->shape_t, x, v, above, f);
        orc_interpolate(tabs[k], cfg->shape_t, 2, 3, f, val);
      }
      for (int c = 0; c < 3; c++) acc[c] = first ? val[c] : acc[c] + val[c];
      first = 0;
    }
    for (int c = 0; c < 3; c++) out[3 * i + c] = acc[c];
  }
  counters_flush();
}

/* g(i) = forward(backward(i)) of the three spaces: the map every re-tabulation goes through.
 * out holds `dims` doubles per texel. */
void orc_roundtrip_4d(const orc_planet *planet, const orc_config *cfg, double *out) {
  long count = prod(cfg->shape4, 4);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < count; i++) {
    double idx[4], x[3], v[3], l[3];
    int above;
    unravel(i, cfg->shape4, 4, idx);
    orc_ray_scatter_backward(planet, cfg->shape4, idx[0], idx[1], idx[2], idx[3], x, v, l, &above);
    orc_ray_scatter_forward(planet, cfg->shape4, x, v, l, above, out + 4 * i);
  }
}

void orc_roundtrip_t(const orc_planet *planet, const orc_config *cfg, double *out) {
  long count = prod(cfg->shape_t, 2);
  for (long i = 0; i < count; i++) {
    double idx[2], x[3], v[3];
    int above;
    unravel(i, cfg->shape_t, 2, idx);
    orc_transmittance_backward(planet, cfg->shape_t, idx[0], idx[1], x, v, &above);
    orc_transmittance_forward(planet, cfg->shape_t, x, v, above, out + 2 * i);
  }
}

void orc_roundtrip_e(const orc_planet *planet, const orc_config *cfg, double *out) {
  long count = prod(cfg->shape_e, 2);
  for (long i = 0; i < count; i++) {
    double idx[2], x[3], l[3];
    unravel(i, cfg->shape_e, 2, idx);
    orc_surface_radiance_backward(planet, cfg->shape_e, idx[0], idx[1], x, l);
    orc_surface_radiance_forward(planet, cfg->shape_e, x, l, out + 2 * i);
  }
}

/* backward(i) of every integer texel of the three spaces (atmosphere.clj:302-310, 348-356, 401-412): the inputs each
 * table entry is integrated for.  Used by the full-grid index-map parity test.  Unused outputs may be NULL. */
void orc_backward_all(const orc_planet *planet, const orc_config *cfg, int which, double *point, double *direction,
                      double *light, int *above) {
  const long *shape = which == 0 ? cfg->shape4 : which == 1 ? cfg->shape_e : cfg->shape_t;
  const int dims = which == 0 ? 4 : 2;
  long count = prod(shape, dims);
#pragma omp parallel for schedule(static)
  for (long i = 0; i < count; i++) {
    double idx[4], x[3], v[3] = {0, 0, 0}, l[3] = {0, 0, 0};
    int ab = 0;
    unravel(i, shape, dims, idx);
    if (which == 0)
      orc_ray_scatter_backward(planet, shape, idx[0], idx[1], idx[2], idx[3], x, v, l, &ab);
    else if (which == 1)
      orc_surface_radiance_backward(planet, shape, idx[0], idx[1], x, l);
    else
      orc_transmittance_backward(planet, shape, idx[0], idx[1], x, v, &ab);
    for (int c = 0; c < 3; c++) {
      if (point) point[3 * i + c] = x[c];
      if (direction) direction[3 * i + c] = v[c];
      if (light) light[3 * i + c] = l[c];
    }
    if (above) above[i] = ab;
  }
}
