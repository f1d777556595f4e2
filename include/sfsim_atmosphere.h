/*
 * libsfsim_atmosphere.so -- C ABI of the B200 atmosphere-LUT accelerator for sfsim.
 *
 * Drop-in boundary for the `clj -T:build atmosphere-lut` path of wedesoft/sfsim
 * (build.clj:84-87 -> src/clj/sfsim/atmosphere_lut.clj:43-105).  The reference has no
 * native interface for this path; the conventions follow its only native precedent, the
 * Jolt wrapper (src/c/sfsim/jolt.hh:1-77: extern "C", POD structs of doubles, arrays as
 * pointer + count, no C++ types), with one deliberate difference: structs are passed by
 * pointer and every entry point returns an int status (0 = ok) because CUDA can fail;
 * atmlut_last_error() gives the message.  INTEGRATION.md shows the coffi/FFM binding.
 *
 * All vectors are double[3] (x, y, z) / RGB triples.  Tables are float32, RGB, row-major,
 * first axis outermost -- the layout `pack-matrices` (matrix.clj:130-134) produces from the
 * nested vectors of `make-lookup-table` (interpolate.clj:68-72).
 *
 * The library is not re-entrant: one call at a time from one thread (like the reference's
 * single JVM thread driving libjolt).  There is no CPU fallback: without a CUDA device every
 * compute entry point returns an error.
 */
#ifndef SFSIM_ATMOSPHERE_H
#define SFSIM_ATMOSPHERE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* planet map of atmosphere.clj:64-67 / atmosphere_lut.clj:24-28 */
typedef struct {
  double centre[3];     /* :sfsim.sphere/centre (index maps assume the origin, atmosphere.clj:95-102) */
  double radius;        /* :sfsim.sphere/radius */
  double height;        /* :sfsim.atmosphere/height */
  double brightness[3]; /* :sfsim.atmosphere/brightness */
} atmlut_planet;

/* scatter map of atmosphere.clj:35-39 */
typedef struct {
  double base[3]; /* ::scatter-base */
  double scale;   /* ::scatter-scale */
  double g;       /* ::scatter-g, 0.0 when the key is absent */
  double quotient;/* ::scatter-quotient, 1.0 when the key is absent */
} atmlut_scatter;

/* the let-bindings of generate-atmosphere-luts, atmosphere_lut.clj:47-63 */
typedef struct {
  int height_size, elevation_size, light_elevation_size, heading_size; /* ray-scatter-shape */
  int transmittance_height_size, transmittance_elevation_size;          /* transmittance-shape */
  int surface_height_size, surface_sun_elevation_size;                  /* surface-radiance-shape */
  int ray_steps, sphere_steps, iterations;
  double intensity[3];
} atmlut_config;

/* ---- lifetime (jolt_init / jolt_destroy, jolt.cc:137-182) ---- */
int atmlut_init(int device);                 /* select the CUDA device, create the stream */
void atmlut_destroy(void);
void *atmlut_stream(void);                   /* the cudaStream_t of the per-table and batch entry points */
const char *atmlut_last_error(void);
int atmlut_device_count(void);
void atmlut_default_config(atmlut_config *cfg); /* shipped constants, atmosphere_lut.clj:47-63 */

/* ---- the whole path: generate-atmosphere-luts (atmosphere_lut.clj:43-105) ----
 * scatter = [mie rayleigh] (atmosphere_lut.clj:64); first-order tables use scatter[1] with its
 * phase function and scatter[0] without (atmosphere_lut.clj:66-69).  Outputs are host buffers
 * in FILE layout, ready for spit-floats (util.clj:236-240):
 *   transmittance    float[Th][Te][3]
 *   surface_radiance float[Sh][Ss][3]
 *   ray_scatter, mie_strength float[H*S][E*A][3]   (convert-4d-to-2d, image.clj:299-312) */
int atmlut_generate(const atmlut_planet *planet, const atmlut_scatter *scatter, int n, const atmlut_config *cfg,
                    float *transmittance, float *surface_radiance, float *ray_scatter, float *mie_strength);

/* The same call spread over the first `num_gpus` GPUs of the box (<= 8, peer access required) from ONE host
 * process and thread -- the form the Clojure host uses.  Results are byte-identical to atmlut_generate. */
int atmlut_generate_multi(const atmlut_planet *planet, const atmlut_scatter *scatter, int n, const atmlut_config *cfg,
                          int num_gpus, float *transmittance, float *surface_radiance, float *ray_scatter,
                          float *mie_strength);

/* ---- device-resident builder: same computation, split so a host can shard it over GPUs ----
 * All-gather mode: rank r of `world` computes a contiguous slab of every 4-D table; after each table the
 * `allgather` callback must make buf[0 .. world*bytes_per_rank) identical on all ranks,
 * given that this rank filled buf[rank*bytes_per_rank ..+bytes_per_rank).  `stream` is the
 * cudaStream_t the slab was produced on; the callback must order its work after it. */
typedef int (*atmlut_allgather_fn)(void *user, void *device_buf, size_t bytes_per_rank, void *stream);
int atmlut_builder_create(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                          const atmlut_config *cfg, int rank, int world, void **builder);
/* host-only helper: the quadrature directions (double[n][3]) and weights the sphere kernels use for table points
 * (normal (1,0,0)); half = 0: integral-sphere(steps) sphere.clj:102-105, half = 1: integral-half-sphere(steps)
 * sphere.clj:96-99.  Returns the number of directions (call with NULL pointers to size), -1 on error. */
int atmlut_sphere_directions(int steps, int half, double *dirs, double *weights, int capacity);
/* the partition: rank owns (height, elevation) pairs [begin, begin + count) of n_pairs = height_size *
 * elevation_size, padded to per_rank pairs per rank (host-only helper, needs no device) */
int atmlut_slab(int n_pairs, int rank, int world, int *begin, int *count, int *per_rank);
/* Peer-to-peer mode (alternative to the all-gather callback; one process per GPU, all on one NVSwitch box,
 * at most 8): every rank exports CUDA IPC handles of its sharded tables (atmlut_builder_ipc_handle_bytes() bytes:
 * 64 per table), the host exchanges them, and every rank imports all of them in rank order (world times that many
 * bytes).  The kernels then store each finished texel into every GPU's table over NVLink, whole height rows are
 * dealt to the ranks in turn (rank r owns rows r, r + world, ...), a flag barrier replaces each all-gather, and the
 * file-layout tables are assembled on rank 0 (atmlut_builder_download is valid there only). */
int atmlut_builder_ipc_handle_bytes(void);
int atmlut_builder_ipc_export(void *builder, unsigned char *handles, int capacity_bytes);
int atmlut_builder_ipc_import(void *builder, const unsigned char *all_handles, int world);
int atmlut_builder_set_allgather(void *builder, atmlut_allgather_fn fn, void *user);
/* options.  ATMLUT_OPT_GRAPH (default 1; ignored in all-gather-callback mode): atmlut_builder_run replays the whole
 * two-stream build as one CUDA graph, captured on the first run.  ATMLUT_OPT_BARRIER_TIMEOUT_MS (default 10000): how
 * long a peer-to-peer barrier waits for a peer before the build is failed; all ranks must call run within this time
 * of each other. */
#define ATMLUT_OPT_GRAPH 1
#define ATMLUT_OPT_BARRIER_TIMEOUT_MS 2
int atmlut_builder_set_option(void *builder, int option, int value);
int atmlut_builder_run(void *builder);        /* asynchronous on the builder's stream */
int atmlut_builder_run_timed(void *builder);  /* the same build with per-stage events (see atmlut_builder_stage_ms) */
void *atmlut_builder_stream(void *builder);   /* the builder's cudaStream_t (for timing / ordering by the host) */
int atmlut_builder_sync(void *builder);       /* wait for the stream */
/* file layout, as atmlut_generate.  Destinations may be pageable (staged through library-owned pinned buffers) or
 * page-locked (written by the copy engine directly; see atmlut_host_alloc). */
int atmlut_builder_download(void *builder, float *transmittance, float *surface_radiance, float *ray_scatter,
                            float *mie_strength);
/* page-locked host memory for output tables (a JVM host allocates its output segments here: FFM arenas are pageable) */
void *atmlut_host_alloc(size_t bytes);
void atmlut_host_free(void *p);
/* per-stage device time of the last atmlut_builder_run_timed in milliseconds; names via atmlut_builder_stage_name */
int atmlut_builder_stage_count(void *builder);
const char *atmlut_builder_stage_name(void *builder, int stage);
int atmlut_builder_stage_ms(void *builder, int stage, float *ms);
/* sample counts of the last run: overall-extinction evaluations, 4-D lookups, 2-D lookups */
int atmlut_builder_work(void *builder, double *esamples, double *lookups4d, double *lookups2d);
/* which: 0 = overall-extinction samples of the first-order kernel, 1 = of the ray-scatter kernels (all
 * iterations), both counted on the device for this rank's slab; 2 = kernels launched by the last run; 3 = MUFU.EX2
 * instructions the sampler issues per overall-extinction sample (2 exponentials each: 2, or 1.5 / 1 where the two scale
 * heights are commensurable and both densities come from one exponential) */
int atmlut_builder_counter(void *builder, int which, double *value);
int atmlut_builder_destroy(void *builder);

/* ---- per-table entry points: make-lookup-table of each public function over its space ----
 * (interpolate.clj:68-72 applied to atmosphere.clj:114-230).  Host float tables in and out,
 * logical layout [..][3]. */
int atmlut_transmittance_table(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                               const atmlut_config *cfg, float *out);
int atmlut_surface_radiance_base_table(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                       const atmlut_config *cfg, float *out);
/* ray-scatter of point-scatter-component (strength = 0) or strength-component (strength = 1) of
 * scatter[component]; either output may be NULL */
int atmlut_first_order_tables(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                              const atmlut_config *cfg, int component_a, int strength_a, float *out_a,
                              int component_b, int strength_b, float *out_b);
/* S source of point-scatter / surface-radiance: ds_a alone, or ds_a + ds_b * phase(scatter[phase_component], v.l)
 * when ds_b != NULL (the closure of atmosphere_lut.clj:79-84) */
int atmlut_point_scatter_table(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                               const atmlut_config *cfg, const float *ds_a, const float *ds_b, int phase_component,
                               const float *de, float *out);
int atmlut_surface_radiance_table(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                  const atmlut_config *cfg, const float *ds_a, const float *ds_b,
                                  int phase_component, float *out);
int atmlut_ray_scatter_table(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                             const atmlut_config *cfg, const float *dj, float *out);
/* re-tabulation of closures (atmosphere_lut.clj:94-101): out[i] = lookup(a, g(i)) [+ lookup(b, g(i))],
 * g = forward o backward of the space.  which: 0 = ray-scatter (4-D), 1 = surface-radiance, 2 = transmittance */
int atmlut_resample_table(const atmlut_planet *planet, const atmlut_config *cfg, int which, const float *a,
                          const float *b, float *out);

/* ---- batch point evaluators: the public functions of sfsim.atmosphere at arbitrary arguments ---- */
/* transmittance, 5-arity (atmosphere.clj:118-125): x -> x0 */
int atmlut_transmittance_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n, int steps,
                               int count, const double *x, const double *x0, double *out);
/* transmittance, 6-arity (atmosphere.clj:126-128): x along v to the shell (above != 0) or the ground */
int atmlut_transmittance_dir_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n, int steps,
                                   int count, const double *x, const double *v, const int *above, double *out);
/* surface-radiance-base (atmosphere.clj:131-137) */
int atmlut_surface_radiance_base_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n, int steps,
                                       const double *intensity, int count, const double *x, const double *l,
                                       double *out);
/* kind 0: point-scatter-component of scatter[component] (atmosphere.clj:170-174)
 * kind 1: strength-component of scatter[component]      (atmosphere.clj:177-182)
 * kind 2: point-scatter-base                            (atmosphere.clj:185-189) */
int atmlut_point_scatter_first_order_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                           int kind, int component, int steps, const double *intensity, int count,
                                           const double *x, const double *v, const double *l, double *out);
/* ray-scatter (atmosphere.clj:192-200) of a first-order point-scatter function (kind/component as above) */
int atmlut_ray_scatter_first_order_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n, int kind,
                                         int component, int steps, const double *intensity, int count,
                                         const double *x, const double *v, const double *l, const int *above,
                                         double *out);

/* The same three integrals with TABLE sources at arbitrary points: the functions the reference passes around
 * as closures are interpolation-tables here (host float RGB tables, logical layout [..][3]).
 * ray-scatter (atmosphere.clj:192-200) with point-scatter = interpolation-table of dj over point-scatter-space */
int atmlut_ray_scatter_table_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n, int steps,
                                   const int *shape4, const float *dj, int count, const double *x, const double *v,
                                   const double *l, const int *above, double *out);
/* point-scatter (atmosphere.clj:203-222); ray-scatter = ds_a [+ ds_b * phase(scatter[phase_component], v.l)],
 * surface-radiance = interpolation-table of de over surface-radiance-space of shape_e */
int atmlut_point_scatter_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n, int sphere_steps,
                               int ray_steps, const int *shape4, const float *ds_a, const float *ds_b,
                               int phase_component, const int *shape_e, const float *de, int count, const double *x,
                               const double *v, const double *l, double *out);
/* surface-radiance (atmosphere.clj:225-230) with the same kind of ray-scatter source */
int atmlut_surface_radiance_batch(const atmlut_planet *planet, const atmlut_scatter *scatter, int n, int steps,
                                  const int *shape4, const float *ds_a, const float *ds_b, int phase_component,
                                  int count, const double *x, const double *l, double *out);

/* ---- index maps (atmosphere.clj:233-422), evaluated on the device in double precision ---- */
/* which: 0 = ray-scatter-space [4 indices; point, direction, light, above]
 *        1 = surface-radiance-space [2 indices; point, light]
 *        2 = transmittance-space [2 indices; point, direction, above]
 * shape: the table shape of that space.  Unused inputs/outputs may be NULL. */
int atmlut_index_forward_batch(const atmlut_planet *planet, int which, const int *shape, int count,
                               const double *point, const double *direction, const double *light, const int *above,
                               double *indices);
int atmlut_index_backward_batch(const atmlut_planet *planet, int which, const int *shape, int count,
                                const double *indices, double *point, double *direction, double *light, int *above);

/* the scalar maps one by one; a, b are double[3] per item, out is double[3] per item (unused lanes 0):
 *   fn 0 elevation-to-index   (a = point, b = direction, flag = above-horizon)        -> out[0]     :239-253
 *   fn 1 index-to-elevation   (a = (radius, index, -))                                -> out = direction, out_flag = above :256-270
 *   fn 2 height-to-index      (a = point)                                             -> out[0]     :273-278
 *   fn 3 index-to-height      (a = (index, -, -))                                     -> out = point :281-288
 *   fn 4 sun-elevation-to-index (a = point, b = light)                                -> out[0]     :322-326
 *   fn 5 index-to-sin-sun-elevation (a = (index, -, -))                               -> out[0]     :329-332
 *   fn 6 sun-angle-to-index   (a = direction, b = light)                              -> out[0]     :368-372
 *   fn 7 index-to-sun-direction (a = direction, b = (sin sun elevation, index, -))    -> out = light :375-384
 *   fn 8 horizon-distance     (a = (radius, -, -))                                    -> out[0]     :233-236 */
int atmlut_index_map_batch(const atmlut_planet *planet, int fn, int size, int count, const double *a,
                           const double *b, const int *flag, double *out, int *out_flag);

/* scattering (fn 0, atmosphere.clj:42-47), extinction (fn 1, :50-53): arg = heights, out = double[count][3];
 * phase (fn 2, :56-61): arg = cosines of the scattering angle, out[3 i] = value.  The device functions every table
 * kernel inlines, evaluated at arbitrary arguments. */
int atmlut_medium_batch(const atmlut_scatter *component, int fn, int count, const double *arg, double *out);

/* interpolate-value (interpolate.clj:87-98) on a float table of `ncomp`-vectors, dims <= 4 */
int atmlut_interpolate_batch(const float *table, const int *shape, int dims, int ncomp, int count,
                             const double *coords, float *out);

/* ---- output (util.clj:227-240, image.clj:299-312) ---- */
int atmlut_convert_4d_to_2d(const float *in, const int *shape, int ncomp, float *out);
int atmlut_write_floats(const char *path, const float *data, long count);
long atmlut_read_floats(const char *path, float *data, long max_count);

#ifdef __cplusplus
}
#endif
#endif
