/*
 * Cube-map tile generation of sfsim on the GPU -- part of libsfsim_atmosphere.so.
 *
 * Drop-in for the per-pixel loops of `clj -T:build cube-maps` (build.clj:294-310 -> src/clj/sfsim/globe.clj:29-80
 * make-cube-map, over the point-wise functions of src/clj/sfsim/cubemap.clj).  For every tile (face, row b, column a)
 * of an output level the reference fills five arrays pixel by pixel and then hands them to its image codecs
 * (spit-jpg, spit-bytes-gz, spit-floats-gz, spit-normals); this library fills the same five arrays.  The codecs and
 * the tar step (globe.clj:83-96) stay on the host.
 *
 * The world rasters (Mercator colour tiles tmp/day, tmp/night and elevation tiles tmp/elevation, produced by
 * build.clj:130-292) live in device memory.  The host hands them over tile by tile, or a level at a time in the
 * tile-major order the reference keeps them in on disk (util.clj:286-290 tile-path <prefix>/<level>/<x>/<y>): level L
 * has 2n x 4n tiles of width x width pixels, n = 2^L;
 *   elevation: int16  [2n][4n][width][width]      (slurp-shorts, cubemap.clj:240-248)
 *   colours  : uint8  [2n][4n][width][width][4]   (slurp-image RGBA, cubemap.clj:232-237)
 * (on the device every tile is copied into its place in one row-major raster per level).
 * A whole level-5 colour raster is 14.9 GB and the level-4 elevation raster 1.9 GB: all levels the path reads fit
 * the 180 GB of one B200 together, so the LRU tile cache of the reference (128 tiles) has no counterpart here.
 *
 * All arithmetic is IEEE double in the reference's operation order.  Conventions as in sfsim_atmosphere.h: int status
 * (0 = ok), atmlut_last_error() for the message, no CPU fallback.  Faces are 0..5 (::face0..::face5, util.clj
 * index->face).
 */
#ifndef SFSIM_CUBEMAP_H
#define SFSIM_CUBEMAP_H

#ifdef __cplusplus
extern "C" {
#endif

/* the constants of make-cube-map (globe.clj:32-40) */
typedef struct {
  int in_level;          /* level of the map tiles to read (build.clj:300-310: out_level - 3)          */
  int out_level;         /* level of the cube-map pyramid to write: 2^out_level x 2^out_level tiles    */
  int width;             /* 675: pixels per map tile edge                                               */
  int surface_tilesize;  /* 65;  colour tiles have 2^sublevel (surface_tilesize - 1) + 1 = 129 pixels   */
  int sublevel;          /* 1                                                                           */
  int max_surface_level; /* 4                                                                           */
  int max_color_level;   /* 5                                                                           */
  double radius;         /* 6378000.0                                                                   */
} sfsim_cubemap_config;

void sfsim_cubemap_default_config(sfsim_cubemap_config *cfg);

/* ---- world rasters in device memory ---- */
int sfsim_cubemap_world_create(int width, void **world);
void sfsim_cubemap_world_destroy(void *world);
/* a whole level at once (tile-major, see above) ... */
int sfsim_cubemap_world_set_elevation(void *world, int level, const short *tiles);
int sfsim_cubemap_world_set_color(void *world, int night, int level, const unsigned char *tiles);
/* ... or tile by tile, as the host reads the files (elevation-tile / world-map-tile, cubemap.clj:232-248) */
int sfsim_cubemap_world_set_elevation_tile(void *world, int level, int ty, int tx, const short *tile);
int sfsim_cubemap_world_set_color_tile(void *world, int night, int level, int ty, int tx, const unsigned char *rgba);

/* ---- make-cube-map for a batch of tiles ----
 * tiles: ntiles triples (face, b, a).  Per tile, with st = surface_tilesize, ct = 2^sublevel (st - 1) + 1:
 *   day, night   uint8 [ct][ct][4]        set-pixel! image.clj:191-199 (alpha 255)            -> spit-jpg
 *   water        uint8 [ct][align4(ct)]   set-byte!  image.clj:249-252, globe.clj:46          -> spit-bytes-gz
 *   surface      float [st][st][3]        project-onto-globe - tile-center, globe.clj:50-55   -> spit-floats-gz
 *   normals      float [ct][ct][3]        normal-for-point, globe.clj:62                      -> spit-normals
 *   normal_bytes int8  [ct][ct][3]        the bytes spit-normals (image.clj:126-136) hands to the PNG encoder
 * Any output pointer may be NULL.  Host pointers; page-locked ones (atmlut_host_alloc) are filled by direct DMA. */
int sfsim_cubemap_tiles(void *world, const sfsim_cubemap_config *cfg, int ntiles, const int *tiles, unsigned char *day,
                        unsigned char *night, unsigned char *water, float *surface, float *normals,
                        signed char *normal_bytes);
/* the same kernels into the library's device buffers only; *ms = device time of the batch (CUDA events) */
int sfsim_cubemap_tiles_timed(void *world, const sfsim_cubemap_config *cfg, int ntiles, const int *tiles, float *ms);
/* rank r of `world_size` takes tiles r, r + world_size, ... of the 6 * 4^out_level tiles of a level, in the order of
 * globe.clj:41 (face, b, a); writes up to `capacity` triples and returns the count through *ntiles */
int sfsim_cubemap_tile_shard(int out_level, int rank, int world_size, int capacity, int *tiles, int *ntiles);

/* ---- make-cube-map for a whole level, streamed (the call `clj -T:build cube-map` would make) ----
 * Generates the tiles rank `rank` of `world_size` owns, `batch_tiles` at a time, and hands every finished tile to `fn`
 * in the order of sfsim_cubemap_tile_shard.  The pointers address page-locked host memory owned by the library and are
 * valid during the call only: the callback encodes / writes the five files of the tile (globe.clj:74-78).  While the
 * callback works on one batch, the next is in flight over PCIe and the one after is being computed.  A non-zero
 * return value of `fn` aborts the level.  The call is bound by the bus (450 KB per tile with all six arrays): a host
 * that encodes the normals from `normal_bytes` itself leaves SFSIM_CUBEMAP_NORMALS out and moves 44 % fewer bytes. */
enum {   /* the arrays a level call computes, brings back and hands to the callback (the others arrive as NULL) */
  SFSIM_CUBEMAP_DAY = 1, SFSIM_CUBEMAP_NIGHT = 2, SFSIM_CUBEMAP_WATER = 4, SFSIM_CUBEMAP_SURFACE = 8,
  SFSIM_CUBEMAP_NORMALS = 16, SFSIM_CUBEMAP_NORMAL_BYTES = 32, SFSIM_CUBEMAP_ALL = 63
};
typedef int (*sfsim_cubemap_tile_fn)(void *user, int face, int b, int a, const unsigned char *day,
                                     const unsigned char *night, const unsigned char *water, const float *surface,
                                     const float *normals, const signed char *normal_bytes);
int sfsim_cubemap_level(void *world, const sfsim_cubemap_config *cfg, int rank, int world_size, int batch_tiles,
                        int outputs, sfsim_cubemap_tile_fn fn, void *user);
/* a ready-made callback that reads the first and last byte of every array and counts the tiles (user = long long[2]:
 * tiles, byte sum): measures the pipeline without a consumer */
int sfsim_cubemap_tile_counter(void *user, int face, int b, int a, const unsigned char *day, const unsigned char *night,
                               const unsigned char *water, const float *surface, const float *normals,
                               const signed char *normal_bytes);

/* ---- the point-wise functions of sfsim.cubemap at arbitrary arguments (known-answer and parity tests) ---- */
/* project-onto-globe (cubemap.clj:336-342): p double[n][3] -> out double[n][3] */
int sfsim_cubemap_project_onto_globe_batch(void *world, int in_level, double radius, int n, const double *p, double *out);
/* normal-for-point (cubemap.clj:357-366) */
int sfsim_cubemap_normal_for_point_batch(void *world, int in_level, int out_level, int tilesize, double radius, int n,
                                         const double *p, double *out);
/* kind 0: elevation-geodetic -> double[n]; 1: water-geodetic -> double[n]; 2 / 3: color-geodetic-day / -night -> double[n][3] */
int sfsim_cubemap_geodetic_batch(void *world, int kind, int in_level, int n, const double *lon, const double *lat,
                                 double *out);

#ifdef __cplusplus
}
#endif
#endif
