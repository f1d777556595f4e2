// Cube-map tile generation (include/sfsim_cubemap.h): globe.clj:29-80 make-cube-map over cubemap.clj.
//
// One thread per output pixel.  A colour pixel costs ten project-onto-globe evaluations (itself and the nine points of
// the normal estimate), each two atan2, two sin/cos pairs, three square roots and one bilinear read of the elevation
// raster: the work is double-precision arithmetic, not memory -- neighbouring pixels read neighbouring raster pixels
// and a 129 x 129 tile touches a few hundred KB of a raster that stays in HBM / L2 for the whole level.
// Device layout of a raster level: ROW-MAJOR [2n width][4n width] -- the host hands tiles over in the reference's
// tile-major file order and every tile lands in place with one 2-D copy, so a pixel address is dy * cols + dx (the
// tile-major address needs two integer divisions by the runtime tile width per read, 104 per colour pixel: measured
// a third of all instructions).  The per-level factors size / (2 pi) and size / pi of map-x / map-y are worked out
// once on the host with the same IEEE division.
// The translation unit is compiled with -fmad=false: every operation is rounded like the reference's JVM arithmetic.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "sfsim_cubemap.h"
#include "atm_api_internal.h"

namespace atm {
namespace cube {

constexpr double kPi = 3.141592653589793;   // clojure.math/PI
constexpr int kLevels = 8;

struct WorldDev {
  const short *elevation[kLevels];
  const uchar4 *day[kLevels];
  const uchar4 *night[kLevels];
  double xscale[kLevels];   // (4 n width) / (2 pi)   map-x, cubemap.clj:171-175
  double yscale[kLevels];   // (2 n width) / pi       map-y, :178-182
  int width;
};

struct D3 {
  double x, y, z;
};

// ------------------------------------------------------------------ cubemap.clj:29-67, 107-111

__device__ __forceinline__ D3 cube_map(int face, double j, double i) {
  switch (face) {
    case 0: return {-1.0 + 2.0 * i, 1.0 - 2.0 * j, 1.0};
    case 1: return {-1.0 + 2.0 * i, -1.0, 1.0 - 2.0 * j};
    case 2: return {1.0, -1.0 + 2.0 * i, 1.0 - 2.0 * j};
    case 3: return {1.0 - 2.0 * i, 1.0, 1.0 - 2.0 * j};
    case 4: return {-1.0, 1.0 - 2.0 * i, 1.0 - 2.0 * j};
    default: return {-1.0 + 2.0 * i, -1.0 + 2.0 * j, -1.0};
  }
}

__host__ __device__ __forceinline__ double cube_coordinate(int level, int tilesize, int tile, double pixel) {
  const int tiles = 1 << level;
  return ((double)tile + pixel / (double)(tilesize - 1)) / (double)tiles;
}

// ------------------------------------------------------------------ cubemap.clj:123-168

__device__ __forceinline__ double mag(D3 p) { return sqrt(p.x * p.x + p.y * p.y + p.z * p.z); }

__device__ __forceinline__ D3 project_onto_sphere(D3 p, double radius) {
  const double m = mag(p);
  return {p.x / m * radius, p.y / m * radius, p.z / m * radius};
}

__device__ __forceinline__ void cartesian_to_geodetic(D3 p, double &lon, double &lat) {
  lon = atan2(p.y, p.x);
  const double q = sqrt(p.x * p.x + p.y * p.y);
  lat = atan2(p.z, q);
}

__device__ __forceinline__ D3 geodetic_to_cartesian(double lon, double lat, double height, double radius) {
  const double distance = height + radius;
  double sin_lat, cos_lat, sin_lon, cos_lon;
  sincos(lat, &sin_lat, &cos_lat);
  sincos(lon, &sin_lon, &cos_lon);
  return {distance * cos_lat * cos_lon, distance * cos_lat * sin_lon, distance * sin_lat};
}

// ------------------------------------------------------------------ cubemap.clj:171-207, 289-299

struct Pixels {
  int i0, i1;
  double f0, f1;
};

__device__ __forceinline__ Pixels map_pixels_x(double lon, int size, double xscale) {
  const double x = (kPi + lon) * xscale;
  const int x0 = (int)floor(x);
  const int x1 = x0 + 1;
  const double frac1 = x - (double)x0;
  // (mod x size): x >= 0 because lon >= -pi, so only x0 = size (lon = pi exactly) and x1 = size wrap
  const int m0 = x0 >= size ? x0 - size : (x0 < 0 ? x0 + size : x0);
  const int m1 = x1 >= size ? x1 - size : (x1 < 0 ? x1 + size : x1);
  return {m0, m1, 1 - frac1, frac1};
}

__device__ __forceinline__ Pixels map_pixels_y(double lat, int size, double yscale) {
  const double y = (kPi / 2 - lat) * yscale;
  const int y0 = (int)floor(y);
  const int y1 = y0 + 1;
  const double frac1 = y - (double)y0;
  return {min(y0, size - 1), min(y1, size - 1), 1 - frac1, frac1};
}

__device__ __forceinline__ double interpolate4(double v0, double v1, double v2, double v3, const Pixels &px, const Pixels &py) {
  return ((v0 * (py.f0 * px.f0) + v1 * (py.f0 * px.f1)) + v2 * (py.f1 * px.f0)) + v3 * (py.f1 * px.f1);
}

// elevation-geodetic (cubemap.clj:323-326)
__device__ __forceinline__ double elevation_geodetic(const WorldDev &w, int level, double lon, double lat) {
  const int rows = (2 << level) * w.width, cols = 2 * rows;
  const Pixels px = map_pixels_x(lon, cols, w.xscale[level]), py = map_pixels_y(lat, rows, w.yscale[level]);
  const short *r0 = w.elevation[level] + (size_t)py.i0 * cols;
  const short *r1 = w.elevation[level] + (size_t)py.i1 * cols;
  return interpolate4((double)__ldg(r0 + px.i0), (double)__ldg(r0 + px.i1), (double)__ldg(r1 + px.i0),
                      (double)__ldg(r1 + px.i1), px, py);
}

// color-geodetic-day / -night (cubemap.clj:311-320)
__device__ __forceinline__ D3 color_geodetic(const WorldDev &w, const uchar4 *__restrict__ img, int level, double lon,
                                             double lat) {
  const int rows = (2 << level) * w.width, cols = 2 * rows;
  const Pixels px = map_pixels_x(lon, cols, w.xscale[level]), py = map_pixels_y(lat, rows, w.yscale[level]);
  const uchar4 *r0 = img + (size_t)py.i0 * cols;
  const uchar4 *r1 = img + (size_t)py.i1 * cols;
  const uchar4 v0 = __ldg(r0 + px.i0), v1 = __ldg(r0 + px.i1), v2 = __ldg(r1 + px.i0), v3 = __ldg(r1 + px.i1);
  return {interpolate4(v0.x, v1.x, v2.x, v3.x, px, py), interpolate4(v0.y, v1.y, v2.y, v3.y, px, py),
          interpolate4(v0.z, v1.z, v2.z, v3.z, px, py)};
}

// water-geodetic (cubemap.clj:329-333)
__device__ __forceinline__ int water_from_height(double height) { return height < 0 ? (int)((height * 255) / -500) : 0; }

// project-onto-globe (cubemap.clj:336-342)
__device__ __forceinline__ D3 project_onto_globe(const WorldDev &w, D3 p, int level, double radius) {
  const D3 sp = project_onto_sphere(p, radius);
  double lon, lat;
  cartesian_to_geodetic(sp, lon, lat);
  const double height = fmax(elevation_geodetic(w, level, lon, lat), 0.0);
  return geodetic_to_cartesian(lon, lat, height, radius);
}

// ------------------------------------------------------------------ cubemap.clj:210-229, 345-366

// fastmath mulv: row . vector, left to right
__device__ __forceinline__ D3 mulv(const double m[9], D3 v) {
  return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z};
}

// lon = (longitude p) = atan2(p.y, p.x), lat = (latitude p) = atan2(p.z, sqrt(p.x^2 + p.y^2))  (cubemap.clj:123-133):
// the tile kernel already holds both from cartesian->geodetic of the same point, the same two expressions
__device__ __forceinline__ void offsets(D3 p, double lon, double lat, int level, int tilesize, D3 &d1, D3 &d2) {
  const double norm = mag(p);
  const double d = (norm * kPi) / (double)(2 * tilesize * (1 << level));
  double s, c;
  sincos(lon, &s, &c);
  const double rz[9] = {c, -s, 0, s, c, 0, 0, 0, 1};              // rotation-matrix-3d-z
  d1 = mulv(rz, D3{0, d, 0});                                     // offset-longitude :210-215
  double sy, cy;
  sincos(-lat, &sy, &cy);
  const double ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};          // rotation-matrix-3d-y
  d2 = mulv(rz, mulv(ry, D3{0, 0, d}));                           // offset-latitude :218-229
}

// The centre point (dj = di = 0) carries weight 0 in both Sobel masks (cubemap.clj:361-362): its projection adds +-0 to
// sums that already hold a +0 or a non-zero value, which changes no bit of them, so it is not evaluated.
__device__ D3 normal_for_point(const WorldDev &w, D3 p, double lon, double lat, int in_level, int out_level, int tilesize,
                               double radius) {
  D3 d1, d2;
  offsets(p, lon, lat, out_level, tilesize, d1, d2);
  const double sx[9] = {-0.25, 0, 0.25, -0.5, 0, 0.5, -0.25, 0, 0.25};
  const double sy[9] = {-0.25, -0.5, -0.25, 0, 0, 0, 0.25, 0.5, 0.25};
  D3 n1 = {0, 0, 0}, n2 = {0, 0, 0};
  int k = 0;
#pragma unroll 1
  for (int dj = -1; dj <= 1; dj++)
#pragma unroll 1
    for (int di = -1; di <= 1; di++, k++) {
      if (k == 4) continue;
      const D3 ps = {p.x + (d2.x * dj + d1.x * di), p.y + (d2.y * dj + d1.y * di), p.z + (d2.z * dj + d1.z * di)};
      const D3 g = project_onto_globe(w, ps, in_level, radius);
      if (k == 0) {
        n1 = {g.x * sx[0], g.y * sx[0], g.z * sx[0]};
        n2 = {g.x * sy[0], g.y * sy[0], g.z * sy[0]};
      } else {
        n1 = {n1.x + g.x * sx[k], n1.y + g.y * sx[k], n1.z + g.z * sx[k]};
        n2 = {n2.x + g.x * sy[k], n2.y + g.y * sy[k], n2.z + g.z * sy[k]};
      }
    }
  const D3 cr = {n1.y * n2.z - n1.z * n2.y, n1.z * n2.x - n1.x * n2.z, n1.x * n2.y - n1.y * n2.x};
  const double m = mag(cr);
  return {cr.x / m, cr.y / m, cr.z / m};
}

// ------------------------------------------------------------------ make-cube-map (globe.clj:29-80)

struct TileJob {
  sfsim_cubemap_config cfg;
  int ls, lc, lw;       // clamped levels: surface / normals, colours, water
  int st, ct, wpitch;
};

struct TileOut {
  uchar4 *day, *night;
  unsigned char *water;
  float *surface, *normals;
  signed char *normal_bytes;
};

// globe.clj:50-55
__global__ void __launch_bounds__(128) k_cube_surface(WorldDev w, TileJob job, const int *__restrict__ tiles, TileOut out) {
  const int t = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= job.st * job.st) return;
  const int face = tiles[3 * t], b = tiles[3 * t + 1], a = tiles[3 * t + 2];
  const int v = pix / job.st, u = pix - v * job.st;
  const double radius = job.cfg.radius;
  // tile-center (cubemap.clj:302-308)
  const D3 center = project_onto_sphere(cube_map(face, cube_coordinate(job.cfg.out_level, 3, b, 1.0),
                                                 cube_coordinate(job.cfg.out_level, 3, a, 1.0)), radius);
  const double j = cube_coordinate(job.cfg.out_level, job.st, b, (double)v);
  const double i = cube_coordinate(job.cfg.out_level, job.st, a, (double)u);
  const D3 point = project_onto_globe(w, cube_map(face, j, i), job.ls, radius);
  float *o = out.surface + ((size_t)t * job.st * job.st + pix) * 3;
  o[0] = (float)(point.x - center.x);
  o[1] = (float)(point.y - center.y);
  o[2] = (float)(point.z - center.z);
}

// globe.clj:56-72
__global__ void __launch_bounds__(128) k_cube_color(WorldDev w, TileJob job, const int *__restrict__ tiles, TileOut out) {
  const int t = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  const int ct = job.ct;
  if (pix >= ct * ct) return;
  const int face = tiles[3 * t], b = tiles[3 * t + 1], a = tiles[3 * t + 2];
  const int v = pix / ct, u = pix - v * ct;
  const double radius = job.cfg.radius;
  const double j = cube_coordinate(job.cfg.out_level, ct, b, (double)v);
  const double i = cube_coordinate(job.cfg.out_level, ct, a, (double)u);
  const D3 point = project_onto_globe(w, cube_map(face, j, i), job.ls, radius);
  double lon, lat;
  cartesian_to_geodetic(point, lon, lat);
  const size_t p = (size_t)t * ct * ct + pix;
  if (out.normals || out.normal_bytes) {
    const D3 n = normal_for_point(w, point, lon, lat, job.ls, job.cfg.out_level, ct, radius);
    const float nx = (float)n.x, ny = (float)n.y, nz = (float)n.z;
    if (out.normals) {
      out.normals[p * 3] = nx;
      out.normals[p * 3 + 1] = ny;
      out.normals[p * 3 + 2] = nz;
    }
    if (out.normal_bytes) {   // spit-normals (image.clj:126-136): round(x 127.5 - 0.5), Math.round = floor(. + 1/2)
      out.normal_bytes[p * 3] = (signed char)(int)floor(((double)nx * 127.5 - 0.5) + 0.5);
      out.normal_bytes[p * 3 + 1] = (signed char)(int)floor(((double)ny * 127.5 - 0.5) + 0.5);
      out.normal_bytes[p * 3 + 2] = (signed char)(int)floor(((double)nz * 127.5 - 0.5) + 0.5);
    }
  }
  if (out.day) {
    const D3 c = color_geodetic(w, w.day[job.lc], job.lc, lon, lat);
    out.day[p] = make_uchar4((unsigned char)(int)c.x, (unsigned char)(int)c.y, (unsigned char)(int)c.z, 255);
  }
  if (out.night) {
    const D3 c = color_geodetic(w, w.night[job.lc], job.lc, lon, lat);
    out.night[p] = make_uchar4((unsigned char)(int)c.x, (unsigned char)(int)c.y, (unsigned char)(int)c.z, 255);
  }
  if (out.water)
    out.water[((size_t)t * ct + v) * job.wpitch + u] =
        (unsigned char)water_from_height(elevation_geodetic(w, job.lw, lon, lat));
}

// ------------------------------------------------------------------ point-wise batch evaluators

__global__ void k_project_onto_globe_batch(WorldDev w, int level, double radius, int n, const double *__restrict__ p,
                                           double *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const D3 g = project_onto_globe(w, D3{p[3 * t], p[3 * t + 1], p[3 * t + 2]}, level, radius);
  out[3 * t] = g.x;
  out[3 * t + 1] = g.y;
  out[3 * t + 2] = g.z;
}

__global__ void k_normal_for_point_batch(WorldDev w, int in_level, int out_level, int tilesize, double radius, int n,
                                         const double *__restrict__ p, double *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const D3 q = {p[3 * t], p[3 * t + 1], p[3 * t + 2]};
  const D3 g = normal_for_point(w, q, atan2(q.y, q.x), atan2(q.z, sqrt(q.x * q.x + q.y * q.y)), in_level, out_level, tilesize,
                                radius);
  out[3 * t] = g.x;
  out[3 * t + 1] = g.y;
  out[3 * t + 2] = g.z;
}

__global__ void k_geodetic_batch(WorldDev w, int kind, int level, int n, const double *__restrict__ lon,
                                 const double *__restrict__ lat, double *__restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  if (kind == 0) {
    out[t] = elevation_geodetic(w, level, lon[t], lat[t]);
  } else if (kind == 1) {
    out[t] = (double)water_from_height(elevation_geodetic(w, level, lon[t], lat[t]));
  } else {
    const D3 c = color_geodetic(w, kind == 2 ? w.day[level] : w.night[level], level, lon[t], lat[t]);
    out[3 * t] = c.x;
    out[3 * t + 1] = c.y;
    out[3 * t + 2] = c.z;
  }
}

// ------------------------------------------------------------------ host side

struct World {
  int width = 0;
  int device = 0;
  short *elevation[kLevels] = {};
  unsigned char *day[kLevels] = {};
  unsigned char *night[kLevels] = {};
  // two sets of device outputs (grown on demand) with their tile lists: the one-batch calls use set 0, the level
  // pipeline alternates; page-locked host images of both sets and a copy stream for the pipeline
  struct OutSet {
    void *dev[6] = {};
    size_t bytes[6] = {};
    int *tiles = nullptr;
    int tiles_capacity = 0;
    void *host[6] = {};
    size_t host_bytes[6] = {};
    cudaEvent_t computed = nullptr, copied = nullptr;
  } sets[2];
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
};

static size_t level_pixels(const World &w, int level) {
  const size_t n = (size_t)1 << level;
  return 2 * n * 4 * n * (size_t)w.width * w.width;
}

static int check_level(const World *w, int level) {
  if (!w) return fail("world must not be NULL");
  if (level < 0 || level >= kLevels) return fail("level must be in [0, 7]");
  return 0;
}

static WorldDev device_view(const World &w) {
  WorldDev d;
  for (int l = 0; l < kLevels; l++) {
    d.elevation[l] = w.elevation[l];
    d.day[l] = reinterpret_cast<const uchar4 *>(w.day[l]);
    d.night[l] = reinterpret_cast<const uchar4 *>(w.night[l]);
    const long long n = 1ll << l;
    d.xscale[l] = (double)(4 * n * w.width) / (2 * kPi);
    d.yscale[l] = (double)(2 * n * w.width) / kPi;
  }
  d.width = w.width;
  return d;
}

static int ensure_level(World *w, int kind, int level) {
  void **slot = kind == 0 ? (void **)&w->elevation[level] : (kind == 1 ? (void **)&w->day[level] : (void **)&w->night[level]);
  if (*slot) return 0;
  const size_t bytes = level_pixels(*w, level) * (kind == 0 ? sizeof(short) : 4);
  CUDA_TRY(cudaMalloc(slot, bytes));
  CUDA_TRY(cudaMemsetAsync(*slot, 0, bytes, stream()));
  return 0;
}

static int make_job(const World &w, const sfsim_cubemap_config *cfg, TileJob &job) {
  if (!cfg) return fail("config must not be NULL");
  if (cfg->width != w.width) return fail("config width differs from the world's tile width");
  if (cfg->out_level < 0 || cfg->out_level > 12) return fail("out_level must be in [0, 12]");
  if (cfg->surface_tilesize < 2 || cfg->surface_tilesize > 1025) return fail("surface_tilesize must be in [2, 1025]");
  if (cfg->sublevel < 0 || cfg->sublevel > 3) return fail("sublevel must be in [0, 3]");
  if (cfg->max_surface_level < 0 || cfg->max_surface_level >= kLevels || cfg->max_color_level < 0 ||
      cfg->max_color_level >= kLevels)
    return fail("max levels must be in [0, 7]");
  if (!(cfg->radius > 0)) return fail("radius must be positive");
  job.cfg = *cfg;
  // globe.clj:53,59-65: (max 0 (min max-level (+ in-level sublevel)))
  job.ls = std::max(0, std::min(cfg->max_surface_level, cfg->in_level));
  job.lc = std::max(0, std::min(cfg->max_color_level, cfg->in_level + cfg->sublevel));
  job.lw = std::max(0, std::min(cfg->max_surface_level, cfg->in_level + cfg->sublevel));
  job.st = cfg->surface_tilesize;
  job.ct = (1 << cfg->sublevel) * (cfg->surface_tilesize - 1) + 1;   // globe.clj:38
  job.wpitch = (job.ct + 3) & ~3;                                     // align-address, globe.clj:46
  return 0;
}

static int run_tiles(World *w, const sfsim_cubemap_config *cfg, int ntiles, const int *tiles, const bool want[6],
                     float *ms, int set_index = 0) {
  if (ensure_init()) return 1;
  if (!w) return fail("world must not be NULL");
  World::OutSet &set = w->sets[set_index];
  TileJob job;
  if (make_job(*w, cfg, job)) return 1;
  if (ntiles < 0 || (ntiles > 0 && !tiles)) return fail("tiles must not be NULL");
  if (ntiles > 65535) return fail("at most 65535 tiles per batch");
  const int n = 1 << cfg->out_level;
  for (int t = 0; t < ntiles; t++)
    if (tiles[3 * t] < 0 || tiles[3 * t] > 5 || tiles[3 * t + 1] < 0 || tiles[3 * t + 1] >= n || tiles[3 * t + 2] < 0 ||
        tiles[3 * t + 2] >= n)
      return fail("tile (face, b, a) out of range for the output level");
  if ((want[3] || want[4] || want[5] || want[0] || want[1] || want[2]) && !w->elevation[job.ls])
    return fail("the elevation raster of level " + std::to_string(job.ls) + " has not been loaded");
  if (want[2] && !w->elevation[job.lw])
    return fail("the elevation raster of level " + std::to_string(job.lw) + " has not been loaded");
  if (want[0] && !w->day[job.lc]) return fail("the day raster of level " + std::to_string(job.lc) + " has not been loaded");
  if (want[1] && !w->night[job.lc])
    return fail("the night raster of level " + std::to_string(job.lc) + " has not been loaded");
  if (ntiles == 0) {
    if (ms) *ms = 0.f;
    return 0;
  }
  const size_t cpix = (size_t)job.ct * job.ct, spix = (size_t)job.st * job.st;
  const size_t need[6] = {cpix * 4 * ntiles, cpix * 4 * ntiles, (size_t)job.ct * job.wpitch * ntiles,
                          spix * 12 * ntiles, cpix * 12 * ntiles, cpix * 3 * ntiles};
  for (int k = 0; k < 6; k++)
    if (want[k] && set.bytes[k] < need[k]) {
      CUDA_TRY(cudaDeviceSynchronize());
      cudaFree(set.dev[k]);
      set.dev[k] = nullptr;
      set.bytes[k] = 0;
      CUDA_TRY(cudaMalloc(&set.dev[k], need[k]));
      set.bytes[k] = need[k];
    }
  if (set.tiles_capacity < ntiles) {
    CUDA_TRY(cudaDeviceSynchronize());
    cudaFree(set.tiles);
    set.tiles = nullptr;
    set.tiles_capacity = 0;
    CUDA_TRY(cudaMalloc((void **)&set.tiles, (size_t)ntiles * 3 * sizeof(int)));
    set.tiles_capacity = ntiles;
  }
  cudaStream_t st = stream();
  CUDA_TRY(cudaMemcpyAsync(set.tiles, tiles, (size_t)ntiles * 3 * sizeof(int), cudaMemcpyHostToDevice, st));
  TileOut out;
  out.day = want[0] ? (uchar4 *)set.dev[0] : nullptr;
  out.night = want[1] ? (uchar4 *)set.dev[1] : nullptr;
  out.water = want[2] ? (unsigned char *)set.dev[2] : nullptr;
  out.surface = want[3] ? (float *)set.dev[3] : nullptr;
  out.normals = want[4] ? (float *)set.dev[4] : nullptr;
  out.normal_bytes = want[5] ? (signed char *)set.dev[5] : nullptr;
  const WorldDev dev = device_view(*w);
  if (ms) CUDA_TRY(cudaEventRecord(w->ev[0], st));
  if (out.water) CUDA_TRY(cudaMemsetAsync(out.water, 0, need[2], st));   // the pad columns of make-byte-image
  if (out.surface)
    k_cube_surface<<<dim3((unsigned)((spix + 127) / 128), ntiles), 128, 0, st>>>(dev, job, set.tiles, out);
  if (out.day || out.night || out.water || out.normals || out.normal_bytes)
    k_cube_color<<<dim3((unsigned)((cpix + 127) / 128), ntiles), 128, 0, st>>>(dev, job, set.tiles, out);
  CUDA_TRY(cudaGetLastError());
  if (ms) {
    CUDA_TRY(cudaEventRecord(w->ev[1], st));
    CUDA_TRY(cudaEventSynchronize(w->ev[1]));
    CUDA_TRY(cudaEventElapsedTime(ms, w->ev[0], w->ev[1]));
  }
  return 0;
}

template <typename F>
static int run_points(World *w, int n, const double *in_a, size_t in_a_count, const double *in_b, size_t in_b_count,
                      double *out, size_t out_count, F launch) {
  if (ensure_init()) return 1;
  if (!w) return fail("world must not be NULL");
  if (n < 0) return fail("count must not be negative");
  if (n == 0) return 0;
  if (!in_a || !out || (in_b_count && !in_b)) return fail("argument arrays must not be NULL");
  double *d_a = nullptr, *d_b = nullptr, *d_o = nullptr;
  cudaStream_t st = stream();
  int rc = 0;
  cudaError_t e = cudaMalloc((void **)&d_a, in_a_count * sizeof(double));
  if (e == cudaSuccess && in_b_count) e = cudaMalloc((void **)&d_b, in_b_count * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc((void **)&d_o, out_count * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_a, in_a, in_a_count * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && in_b_count) e = cudaMemcpyAsync(d_b, in_b, in_b_count * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    launch(d_a, d_b, d_o, st);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_o, out_count * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) rc = fail_cuda(e, "cube-map batch evaluation");
  cudaFree(d_a);
  cudaFree(d_b);
  cudaFree(d_o);
  return rc;
}

}  // namespace cube
}  // namespace atm

using namespace atm;
using namespace atm::cube;

extern "C" int sfsim_cubemap_tile_shard(int out_level, int rank, int world_size, int capacity, int *tiles, int *ntiles);

extern "C" void sfsim_cubemap_default_config(sfsim_cubemap_config *cfg) {
  if (!cfg) return;
  cfg->in_level = -3;   // build.clj:300-302: (cube-map {:in-level -3 :out-level 0})
  cfg->out_level = 0;
  cfg->width = 675;
  cfg->surface_tilesize = 65;
  cfg->sublevel = 1;
  cfg->max_surface_level = 4;
  cfg->max_color_level = 5;
  cfg->radius = 6378000.0;
}

extern "C" int sfsim_cubemap_world_create(int width, void **world) {
  if (ensure_init()) return 1;
  if (!world) return fail("world must not be NULL");
  if (width < 1 || width > 8192) return fail("tile width must be in [1, 8192]");
  World *w = new World();
  w->width = width;
  cudaGetDevice(&w->device);
  for (auto &e : w->ev)
    if (cudaEventCreate(&e) != cudaSuccess) {
      delete w;
      return fail("cudaEventCreate failed");
    }
  *world = w;
  return 0;
}

extern "C" void sfsim_cubemap_world_destroy(void *world) {
  World *w = (World *)world;
  if (!w) return;
  if (stream()) cudaStreamSynchronize(stream());
  for (int l = 0; l < kLevels; l++) {
    cudaFree(w->elevation[l]);
    cudaFree(w->day[l]);
    cudaFree(w->night[l]);
  }
  cudaDeviceSynchronize();
  for (auto &set : w->sets) {
    for (auto &o : set.dev) cudaFree(o);
    for (auto &h : set.host) cudaFreeHost(h);
    cudaFree(set.tiles);
    if (set.computed) cudaEventDestroy(set.computed);
    if (set.copied) cudaEventDestroy(set.copied);
  }
  if (w->copy_stream) cudaStreamDestroy(w->copy_stream);
  for (auto &e : w->ev)
    if (e) cudaEventDestroy(e);
  delete w;
}

static int check_tile(int level, int ty, int tx) {
  const int n = 1 << level;
  if (ty < 0 || ty >= 2 * n || tx < 0 || tx >= 4 * n) return fail("map tile index out of range for the level");
  return 0;
}

// One map tile (width x width elements of `elt` bytes, contiguous on the host) into its place in the row-major raster.
static cudaError_t put_tile(const World &w, void *raster, int level, int ty, int tx, const void *tile, size_t elt) {
  const size_t cols = (size_t)(4 << level) * w.width;
  char *dst = (char *)raster + ((size_t)ty * w.width * cols + (size_t)tx * w.width) * elt;
  return cudaMemcpy2DAsync(dst, cols * elt, tile, (size_t)w.width * elt, (size_t)w.width * elt, (size_t)w.width,
                           cudaMemcpyHostToDevice, stream());
}

// kind 0: elevation, 1: day, 2: night; ty < 0: every tile of the level from a tile-major host buffer
static int upload(void *world, int kind, int level, int ty, int tx, const void *data) {
  World *w = (World *)world;
  if (ensure_init() || check_level(w, level)) return 1;
  if (!data) return fail("the tile data must not be NULL");
  if (ty >= 0 && check_tile(level, ty, tx)) return 1;
  if (ensure_level(w, kind, level)) return 1;
  void *raster = kind == 0 ? (void *)w->elevation[level] : (kind == 1 ? (void *)w->day[level] : (void *)w->night[level]);
  const size_t elt = kind == 0 ? sizeof(short) : 4;
  if (ty >= 0) {
    CUDA_TRY(put_tile(*w, raster, level, ty, tx, data, elt));
  } else {
    const size_t tile_bytes = (size_t)w->width * w->width * elt;
    for (int y = 0; y < (2 << level); y++)
      for (int x = 0; x < (4 << level); x++)
        CUDA_TRY(put_tile(*w, raster, level, y, x, (const char *)data + ((size_t)y * (4 << level) + x) * tile_bytes, elt));
  }
  CUDA_TRY(cudaStreamSynchronize(stream()));
  return 0;
}

extern "C" int sfsim_cubemap_world_set_elevation(void *world, int level, const short *tiles) {
  return upload(world, 0, level, -1, -1, tiles);
}

extern "C" int sfsim_cubemap_world_set_color(void *world, int night, int level, const unsigned char *tiles) {
  return upload(world, night ? 2 : 1, level, -1, -1, tiles);
}

extern "C" int sfsim_cubemap_world_set_elevation_tile(void *world, int level, int ty, int tx, const short *tile) {
  if (ty < 0) return fail("map tile index out of range for the level");
  return upload(world, 0, level, ty, tx, tile);
}

extern "C" int sfsim_cubemap_world_set_color_tile(void *world, int night, int level, int ty, int tx,
                                                  const unsigned char *rgba) {
  if (ty < 0) return fail("map tile index out of range for the level");
  return upload(world, night ? 2 : 1, level, ty, tx, rgba);
}

extern "C" int sfsim_cubemap_tiles(void *world, const sfsim_cubemap_config *cfg, int ntiles, const int *tiles,
                                   unsigned char *day, unsigned char *night, unsigned char *water, float *surface,
                                   float *normals, signed char *normal_bytes) {
  World *w = (World *)world;
  void *dst[6] = {day, night, water, surface, normals, normal_bytes};
  bool want[6];
  for (int k = 0; k < 6; k++) want[k] = dst[k] != nullptr;
  if (run_tiles(w, cfg, ntiles, tiles, want, nullptr)) return 1;
  if (ntiles == 0) return 0;
  TileJob job;
  if (make_job(*w, cfg, job)) return 1;
  const size_t cpix = (size_t)job.ct * job.ct, spix = (size_t)job.st * job.st;
  const size_t bytes[6] = {cpix * 4 * ntiles, cpix * 4 * ntiles, (size_t)job.ct * job.wpitch * ntiles,
                           spix * 12 * ntiles, cpix * 12 * ntiles, cpix * 3 * ntiles};
  for (int k = 0; k < 6; k++)
    if (want[k]) CUDA_TRY(cudaMemcpyAsync(dst[k], w->sets[0].dev[k], bytes[k], cudaMemcpyDeviceToHost, stream()));
  CUDA_TRY(cudaStreamSynchronize(stream()));
  return 0;
}

extern "C" int sfsim_cubemap_tiles_timed(void *world, const sfsim_cubemap_config *cfg, int ntiles, const int *tiles,
                                         float *ms) {
  if (!ms) return fail("ms must not be NULL");
  const bool want[6] = {true, true, true, true, true, true};
  return run_tiles((World *)world, cfg, ntiles, tiles, want, ms);
}

// make-cube-map for this rank's share of a level (globe.clj:41-72), streamed: while the host callback consumes batch
// i - 1 from page-locked memory, the copy engine brings batch i back and the SMs compute batch i + 1.
extern "C" int sfsim_cubemap_level(void *world, const sfsim_cubemap_config *cfg, int rank, int world_size, int batch_tiles,
                                   int outputs, sfsim_cubemap_tile_fn fn, void *user) {
  World *w = (World *)world;
  if (ensure_init()) return 1;
  if (!w) return fail("world must not be NULL");
  if (!fn) return fail("the tile callback must not be NULL");
  TileJob job;
  if (make_job(*w, cfg, job)) return 1;
  if (batch_tiles < 1 || batch_tiles > 65535) return fail("batch_tiles must be in [1, 65535]");
  if (outputs <= 0 || outputs > SFSIM_CUBEMAP_ALL) return fail("outputs must be a non-empty mask of SFSIM_CUBEMAP_* bits");
  int count = 0;
  if (sfsim_cubemap_tile_shard(cfg->out_level, rank, world_size, 0, nullptr, &count)) return 1;
  std::vector<int> tiles((size_t)std::max(count, 1) * 3);
  if (sfsim_cubemap_tile_shard(cfg->out_level, rank, world_size, count, tiles.data(), &count)) return 1;
  if (count == 0) return 0;
  batch_tiles = std::min(batch_tiles, count);
  const size_t cpix = (size_t)job.ct * job.ct, spix = (size_t)job.st * job.st;
  const size_t per_tile[6] = {cpix * 4, cpix * 4, (size_t)job.ct * job.wpitch, spix * 12, cpix * 12, cpix * 3};
  if (!w->copy_stream) CUDA_TRY(cudaStreamCreateWithFlags(&w->copy_stream, cudaStreamNonBlocking));
  for (auto &set : w->sets) {
    if (!set.computed) CUDA_TRY(cudaEventCreateWithFlags(&set.computed, cudaEventDisableTiming));
    if (!set.copied) CUDA_TRY(cudaEventCreateWithFlags(&set.copied, cudaEventDisableTiming));
    for (int k = 0; k < 6; k++)
      if ((outputs >> k & 1) && set.host_bytes[k] < per_tile[k] * batch_tiles) {
        CUDA_TRY(cudaDeviceSynchronize());
        cudaFreeHost(set.host[k]);
        set.host[k] = nullptr;
        set.host_bytes[k] = 0;
        CUDA_TRY(cudaMallocHost(&set.host[k], per_tile[k] * batch_tiles));
        set.host_bytes[k] = per_tile[k] * batch_tiles;
      }
  }
  bool want[6];
  for (int k = 0; k < 6; k++) want[k] = (outputs >> k & 1) != 0;
  const int nbatches = (count + batch_tiles - 1) / batch_tiles;
  auto batch_size = [&](int i) { return std::min(batch_tiles, count - i * batch_tiles); };
  auto deliver = [&](int i) -> int {       // hand the tiles of batch i to the host, in order
    World::OutSet &set = w->sets[i & 1];
    CUDA_TRY(cudaEventSynchronize(set.copied));
    for (int t = 0; t < batch_size(i); t++) {
      const int *tile = &tiles[(size_t)(i * batch_tiles + t) * 3];
      const char *p[6];
      for (int k = 0; k < 6; k++) p[k] = want[k] ? (const char *)set.host[k] + per_tile[k] * t : nullptr;
      if (fn(user, tile[0], tile[1], tile[2], (const unsigned char *)p[0], (const unsigned char *)p[1],
             (const unsigned char *)p[2], (const float *)p[3], (const float *)p[4], (const signed char *)p[5]))
        return fail("the tile callback reported an error");
    }
    return 0;
  };
  for (int i = 0; i < nbatches; i++) {
    World::OutSet &set = w->sets[i & 1];
    // the kernels of batch i overwrite the device set batch i - 2 used: its copy must have left the device
    if (i >= 2) CUDA_TRY(cudaStreamWaitEvent(stream(), set.copied, 0));
    if (run_tiles(w, cfg, batch_size(i), &tiles[(size_t)i * batch_tiles * 3], want, nullptr, i & 1)) return 1;
    CUDA_TRY(cudaEventRecord(set.computed, stream()));
    CUDA_TRY(cudaStreamWaitEvent(w->copy_stream, set.computed, 0));
    // the host image of this set was consumed by deliver(i - 2), which ran before this point
    for (int k = 0; k < 6; k++)
      if (want[k])
        CUDA_TRY(cudaMemcpyAsync(set.host[k], set.dev[k], per_tile[k] * batch_size(i), cudaMemcpyDeviceToHost, w->copy_stream));
    CUDA_TRY(cudaEventRecord(set.copied, w->copy_stream));
    if (i >= 1 && deliver(i - 1)) return 1;
  }
  return deliver(nbatches - 1);
}

extern "C" int sfsim_cubemap_tile_counter(void *user, int, int, int, const unsigned char *day, const unsigned char *night,
                                          const unsigned char *water, const float *surface, const float *normals,
                                          const signed char *normal_bytes) {
  long long *acc = (long long *)user;
  if (!acc) return 1;
  acc[0] += 1;
  acc[1] += (day ? day[0] : 0) + (night ? night[0] : 0) + (water ? water[0] : 0) + (surface && surface[0] != 0.f) +
            (normals && normals[0] != 0.f) + (normal_bytes ? normal_bytes[0] : 0);
  return 0;
}

extern "C" int sfsim_cubemap_tile_shard(int out_level, int rank, int world_size, int capacity, int *tiles, int *ntiles) {
  if (out_level < 0 || out_level > 12) return fail("out_level must be in [0, 12]");
  if (world_size < 1 || rank < 0 || rank >= world_size) return fail("rank must be in [0, world_size)");
  if (!ntiles || (capacity > 0 && !tiles)) return fail("tiles and ntiles must not be NULL");
  const long long n = 1ll << out_level, total = 6 * n * n;
  int count = 0;
  for (long long t = rank; t < total; t += world_size, count++)
    if (count < capacity) {   // globe.clj:41: [k (range 6) b (range n) a (range n)]
      tiles[3 * count] = (int)(t / (n * n));
      tiles[3 * count + 1] = (int)((t / n) % n);
      tiles[3 * count + 2] = (int)(t % n);
    }
  *ntiles = count;
  return 0;
}

extern "C" int sfsim_cubemap_project_onto_globe_batch(void *world, int in_level, double radius, int n, const double *p,
                                                      double *out) {
  World *w = (World *)world;
  if (check_level(w, in_level)) return 1;
  if (!w->elevation[in_level]) return fail("the elevation raster of that level has not been loaded");
  const WorldDev dev = device_view(*w);
  return run_points(w, n, p, (size_t)n * 3, nullptr, 0, out, (size_t)n * 3,
                    [&](const double *a, const double *, double *o, cudaStream_t st) {
                      k_project_onto_globe_batch<<<(n + 127) / 128, 128, 0, st>>>(dev, in_level, radius, n, a, o);
                    });
}

extern "C" int sfsim_cubemap_normal_for_point_batch(void *world, int in_level, int out_level, int tilesize, double radius,
                                                    int n, const double *p, double *out) {
  World *w = (World *)world;
  if (check_level(w, in_level)) return 1;
  if (!w->elevation[in_level]) return fail("the elevation raster of that level has not been loaded");
  if (out_level < 0 || out_level > 12 || tilesize < 1) return fail("out_level must be in [0, 12] and tilesize positive");
  const WorldDev dev = device_view(*w);
  return run_points(w, n, p, (size_t)n * 3, nullptr, 0, out, (size_t)n * 3,
                    [&](const double *a, const double *, double *o, cudaStream_t st) {
                      k_normal_for_point_batch<<<(n + 127) / 128, 128, 0, st>>>(dev, in_level, out_level, tilesize, radius,
                                                                               n, a, o);
                    });
}

extern "C" int sfsim_cubemap_geodetic_batch(void *world, int kind, int in_level, int n, const double *lon,
                                            const double *lat, double *out) {
  World *w = (World *)world;
  if (check_level(w, in_level)) return 1;
  if (kind < 0 || kind > 3) return fail("kind must be 0 (elevation), 1 (water), 2 (day) or 3 (night)");
  const void *raster = kind <= 1 ? (const void *)w->elevation[in_level]
                                 : (kind == 2 ? (const void *)w->day[in_level] : (const void *)w->night[in_level]);
  if (!raster) return fail("the raster of that level has not been loaded");
  const WorldDev dev = device_view(*w);
  return run_points(w, n, lon, (size_t)n, lat, (size_t)n, out, (size_t)n * (kind >= 2 ? 3 : 1),
                    [&](const double *a, const double *b, double *o, cudaStream_t st) {
                      k_geodetic_batch<<<(n + 127) / 128, 128, 0, st>>>(dev, kind, in_level, n, a, b, o);
                    });
}
