/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the sfsim atmosphere-LUT hot path.
 *
 * Double-precision C restatement of the reference's Clojure algorithm
 * (wedesoft/sfsim: src/clj/sfsim/atmosphere.clj, atmosphere_lut.clj, interpolate.clj,
 * ray.clj, sphere.clj, matrix.clj, quaternion.clj, util.clj, image.clj).  Each function
 * cites the reference file:line it follows.  Nothing under sfsim_b200/ may link, import
 * or call this code; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs use it, and only as the checker.
 *
 * Third-party arithmetic the reference relies on but does not vendor:
 * generateme/fastmath 2.4.0 (deps.edn:10) -- Vec3 add/sub/mult/div/emult/dot/mag/
 * normalize/cross/exp and 3x3 transpose/mulv, all component-wise IEEE double
 * operations -- and clojure.math (java.lang.Math).  They are restated here with the C
 * operators and libm.
 *
 * Parity pin: checked against the reference's own known-answer tests (see
 * tests/test_oracle_known_answers.py, which cites test/clj/sfsim/t_atmosphere.clj etc.).
 */
#ifndef SFSIM_ATMOSPHERE_ORACLE_H
#define SFSIM_ATMOSPHERE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  double centre[3];
  double radius;
  double height;         /* :sfsim.atmosphere/height */
  double brightness[3];  /* :sfsim.atmosphere/brightness */
} orc_planet;

typedef struct {
  double base[3];   /* ::scatter-base */
  double scale;     /* ::scatter-scale */
  double g;         /* ::scatter-g, 0 when absent */
  double quotient;  /* ::scatter-quotient, 1 when absent */
} orc_scatter;

/* point-scatter style callback: J(p, v, l, above) -> rgb */
typedef void (*orc_point_fn)(void *ctx, const double p[3], const double v[3], const double l[3], int above,
                             double out[3]);
/* surface-radiance style callback: E(p, l) -> rgb */
typedef void (*orc_surface_fn)(void *ctx, const double p[3], const double l[3], double out[3]);
/* generic vector integrand */
typedef void (*orc_vec_fn)(void *ctx, const double p[3], double *out);
typedef void (*orc_angle_fn)(void *ctx, double phi, double *out);

/* ---- util.clj ---- */
double orc_limit_quot(double a, double b, double lower, double upper);
double orc_limit_quot3(double a, double b, double limit);

/* ---- sphere.clj / ray.clj / matrix.clj / quaternion.clj ---- */
double orc_height(const orc_planet *planet, const double p[3]);
void orc_ray_sphere_intersection(const double centre[3], double radius, const double origin[3],
                                 const double direction[3], double *distance, double *length);
void orc_integral_ray(const double origin[3], const double direction[3], long steps, double distance,
                      orc_vec_fn fun, void *ctx, int dims, double *out);
void orc_integrate_circle(long steps, orc_angle_fn fun, void *ctx, int dims, double *out);
void orc_integral_half_sphere(long steps, const double normal[3], orc_vec_fn fun, void *ctx, int dims, double *out);
void orc_integral_sphere(long steps, const double normal[3], orc_vec_fn fun, void *ctx, int dims, double *out);
void orc_orthogonal(const double n[3], double out[3]);
void orc_oriented_matrix(const double n[3], double out[9]);
/* direction list of spherical-integral in evaluation order; returns count (call with NULLs to size) */
long orc_sphere_directions(long theta_steps, long phi_steps, double theta_range, const double normal[3],
                           double *dirs /* [n][3] */, double *weights /* [n] = factor * 2pi / ringsteps */);

/* ---- atmosphere.clj: medium and geometry ---- */
void orc_scattering(const orc_scatter *s, double height, double out[3]);
void orc_extinction(const orc_scatter *s, double height, double out[3]);
double orc_phase(const orc_scatter *s, double mu);
void orc_atmosphere_intersection(const orc_planet *planet, const double origin[3], const double direction[3],
                                 double out[3]);
void orc_surface_intersection(const orc_planet *planet, const double origin[3], const double direction[3],
                              double out[3]);
int orc_surface_point(const orc_planet *planet, const double p[3]);
int orc_is_above_horizon(const orc_planet *planet, const double p[3], const double direction[3]);
void orc_ray_extremity(const orc_planet *planet, const double origin[3], const double direction[3], double out[3]);

/* ---- atmosphere.clj: radiative quantities ---- */
void orc_transmittance(const orc_planet *planet, const orc_scatter *scatter, int n, long steps, const double x[3],
                       const double x0[3], double out[3]);
void orc_transmittance_dir(const orc_planet *planet, const orc_scatter *scatter, int n, long steps,
                           const double x[3], const double v[3], int above, double out[3]);
void orc_surface_radiance_base(const orc_planet *planet, const orc_scatter *scatter, int n, long steps,
                               const double intensity[3], const double x[3], const double l[3], double out[3]);
void orc_point_scatter_component(const orc_planet *planet, const orc_scatter *scatter, int n,
                                 const orc_scatter *component, long steps, const double intensity[3],
                                 const double x[3], const double v[3], const double l[3], int above, double out[3]);
void orc_strength_component(const orc_planet *planet, const orc_scatter *scatter, int n,
                            const orc_scatter *component, long steps, const double intensity[3], const double x[3],
                            const double v[3], const double l[3], int above, double out[3]);
void orc_point_scatter_base(const orc_planet *planet, const orc_scatter *scatter, int n, long steps,
                            const double intensity[3], const double x[3], const double v[3], const double l[3],
                            int above, double out[3]);
void orc_ray_scatter(const orc_planet *planet, const orc_scatter *scatter, int n, long steps, orc_point_fn point_scatter,
                     void *ctx, const double x[3], const double v[3], const double l[3], int above, double out[3]);
void orc_point_scatter(const orc_planet *planet, const orc_scatter *scatter, int n, orc_point_fn ray_scatter,
                       void *rs_ctx, orc_surface_fn surface_radiance, void *sr_ctx, const double intensity[3],
                       long sphere_steps, long ray_steps, const double x[3], const double v[3], const double l[3],
                       int above, double out[3]);
/* test hooks (with-redefs analogue) and the single-direction integrand of point-scatter */
typedef struct {
  double (*phase)(const orc_scatter *s, double mu);
  void (*ray_extremity)(const orc_planet *planet, const double origin[3], const double direction[3], double out[3]);
  void (*transmittance)(const orc_planet *planet, const orc_scatter *scatter, int n, long steps, const double x[3],
                        const double x0[3], double out[3]);
} orc_test_hooks;
void orc_set_test_hooks(const orc_test_hooks *hooks);
void orc_in_scatter_from_direction(const orc_planet *planet, const orc_scatter *scatter, int n,
                                   orc_point_fn ray_scatter, void *rs_ctx, orc_surface_fn surface_radiance,
                                   void *sr_ctx, long ray_steps, const double x[3], const double v[3],
                                   const double l[3], const double omega[3], double out[3]);
void orc_surface_radiance(const orc_planet *planet, orc_point_fn ray_scatter, void *rs_ctx, long steps,
                          const double x[3], const double l[3], double out[3]);

/* ---- atmosphere.clj: index maps ---- */
double orc_horizon_distance(const orc_planet *planet, double radius);
double orc_elevation_to_index(const orc_planet *planet, long size, const double point[3], const double direction[3],
                              int above);
void orc_index_to_elevation(const orc_planet *planet, long size, double radius, double index, double dir[3],
                            int *above);
double orc_height_to_index(const orc_planet *planet, long size, const double point[3]);
void orc_index_to_height(const orc_planet *planet, long size, double index, double point[3]);
double orc_sun_elevation_to_index(long size, const double point[3], const double l[3]);
double orc_index_to_sin_sun_elevation(long size, double index);
double orc_sun_angle_to_index(long size, const double direction[3], const double l[3]);
void orc_index_to_sun_direction(long size, const double direction[3], double sin_sun_elevation, double index,
                                double out[3]);
void orc_transmittance_forward(const orc_planet *planet, const long shape[2], const double point[3],
                               const double direction[3], int above, double idx[2]);
void orc_transmittance_backward(const orc_planet *planet, const long shape[2], double hi, double ei, double point[3],
                                double direction[3], int *above);
void orc_surface_radiance_forward(const orc_planet *planet, const long shape[2], const double point[3],
                                  const double l[3], double idx[2]);
void orc_surface_radiance_backward(const orc_planet *planet, const long shape[2], double hi, double si,
                                   double point[3], double l[3]);
void orc_ray_scatter_forward(const orc_planet *planet, const long shape[4], const double point[3],
                             const double direction[3], const double l[3], int above, double idx[4]);
void orc_ray_scatter_backward(const orc_planet *planet, const long shape[4], double hi, double ei, double si,
                              double ai, double point[3], double direction[3], double l[3], int *above);

/* ---- interpolate.clj ---- */
/* multilinear lookup in a row-major table of `ncomp`-vectors (interpolate-value) */
void orc_interpolate(const double *table, const long *shape, int dims, int ncomp, const double *coords, double *out);

/* ---- matrix.clj pack-matrices, image.clj convert-4d-to-2d, util.clj spit-floats ---- */
void orc_pack_floats(const double *in, long count, float *out);
void orc_convert_4d_to_2d(const double *in, const long shape[4], int ncomp, double *out);
int orc_spit_floats(const char *path, const float *data, long count);
long orc_slurp_floats(const char *path, float *data, long max_count);

/* ---- atmosphere_lut.clj: table builders (make-lookup-table of each stage) ----
 * `indices` selects flat row-major texel indices to evaluate (NULL = all `count` = product of shape);
 * out[i*3..] receives texel indices[i].  All tables are double RGB, row-major, first axis outermost. */
typedef struct {
  long shape4[4];       /* height, elevation, light-elevation, heading */
  long shape_t[2];      /* transmittance height, elevation */
  long shape_e[2];      /* surface height, sun elevation */
  long ray_steps, sphere_steps;
  double intensity[3];
} orc_config;

void orc_table_transmittance(const orc_planet *planet, const orc_scatter *scatter, int n, const orc_config *cfg,
                             const long *indices, long count, double *out);
void orc_table_surface_radiance_base(const orc_planet *planet, const orc_scatter *scatter, int n,
                                     const orc_config *cfg, const long *indices, long count, double *out);
/* first-order ray-scatter of `component` with (strength=0) or without (strength=1) phase function */
void orc_table_first_order(const orc_planet *planet, const orc_scatter *scatter, int n, const orc_config *cfg,
                           const orc_scatter *component, int strength, const long *indices, long count, double *out);
/* S source: kind 0 = single table tab_a; kind 1 = tab_a + tab_b * phase(phase_component, v.l) (atmosphere_lut.clj:79-84) */
typedef struct {
  int kind;
  const double *tab_a, *tab_b;
  orc_scatter phase_component;
} orc_s_source;
void orc_table_point_scatter(const orc_planet *planet, const orc_scatter *scatter, int n, const orc_config *cfg,
                             const orc_s_source *ds, const double *de, const long *indices, long count, double *out);
void orc_table_surface_radiance(const orc_planet *planet, const orc_config *cfg, const orc_s_source *ds,
                                const long *indices, long count, double *out);
void orc_table_ray_scatter(const orc_planet *planet, const orc_scatter *scatter, int n, const orc_config *cfg,
                           const double *dj, const long *indices, long count, double *out);
/* out[i] = sum_k lookup(tabs[k], forward(backward(i))) ; tabs may hold NULL entries (the constantly-zero E) */
void orc_table_resample_sum_4d(const orc_planet *planet, const orc_config *cfg, const double *const *tabs, int ntabs,
                               const long *indices, long count, double *out);
void orc_table_resample_sum_e(const orc_planet *planet, const orc_config *cfg, const double *const *tabs, int ntabs,
                              const long *indices, long count, double *out);
void orc_table_resample_sum_t(const orc_planet *planet, const orc_config *cfg, const double *const *tabs, int ntabs,
                              const long *indices, long count, double *out);

/* g(i) = forward(backward(i)) for every texel of a space (dims doubles per texel) */
void orc_roundtrip_4d(const orc_planet *planet, const orc_config *cfg, double *out);
void orc_roundtrip_t(const orc_planet *planet, const orc_config *cfg, double *out);
void orc_roundtrip_e(const orc_planet *planet, const orc_config *cfg, double *out);
/* backward(i) for every integer texel; which: 0 = ray-scatter-space, 1 = surface-radiance-space, 2 = transmittance-space */
void orc_backward_all(const orc_planet *planet, const orc_config *cfg, int which, double *point, double *direction,
                      double *light, int *above);

/* counters: overall-extinction evaluations ("E-samples") and table lookups since the last reset */
void orc_counters_reset(void);
void orc_counters_get(long long *esamples, long long *lookups4d, long long *lookups2d);
int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
