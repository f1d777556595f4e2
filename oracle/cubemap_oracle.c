/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle of sfsim's cube-map tile generation (`clj -T:build cube-maps`,
 * build.clj:294-310 -> src/clj/sfsim/globe.clj:29-80 over src/clj/sfsim/cubemap.clj).  Double-precision
 * restatement, one C function per reference function, same operation order (compiled with
 * -ffp-contract=off).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may load it.
 *
 * Third-party arithmetic not under /root/reference: generateme/fastmath 2.4.0 (deps.edn:10) supplies
 * vec3 add mult div cross mag normalize and rotation-matrix-3d-y/z, mulv.  They are plain component-wise
 * IEEE-double operations; the two rotation matrices are pinned through the reference's own
 * offset-longitude / offset-latitude facts (t_cubemap.clj:216-229).
 *
 * World rasters are passed tile-major, the way the reference keeps them on disk
 * (tmp/elevation/<level>/<x>/<y>.raw, tmp/day/<level>/<x>/<y>.png, util.clj:286-290):
 *   elevation: int16 [2n][4n][width][width], colours: uint8 [2n][4n][width][width][4], n = 2^level.
 * Parity status: PINNED to t_cubemap.clj (tests/test_cubemap_oracle.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PI 3.141592653589793 /* clojure.math/PI = java.lang.Math.PI */

typedef struct {
  const int16_t *elevation[8]; /* per level, tile-major; NULL = level not loaded */
  const uint8_t *day[8];
  const uint8_t *night[8];
  long width; /* pixels per map tile edge (675) */
} orc_world;

/* ------------------------------------------------------------------ fastmath.vector helpers */
static double sqr(double x) { return x * x; } /* util.clj:334-337 */
static double mag3(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
static void normalize3(const double a[3], double o[3]) {
  double m = mag3(a);
  o[0] = a[0] / m;
  o[1] = a[1] / m;
  o[2] = a[2] / m;
}
static void mulv3(const double m[9], const double v[3], double o[3]) {
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  double y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  double z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  o[0] = x;
  o[1] = y;
  o[2] = z;
}
static void rotation_z(double a, double m[9]) {
  double s = sin(a), c = cos(a);
  double r[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
  memcpy(m, r, sizeof r);
}
static void rotation_y(double a, double m[9]) {
  double s = sin(a), c = cos(a);
  double r[9] = {c, 0, s, 0, 1, 0, -s, 0, c};
  memcpy(m, r, sizeof r);
}

/* ------------------------------------------------------------------ cube faces (cubemap.clj:29-67) */
double orc_cm_cube_map_x(int face, double j, double i) {
  (void)j;
  switch (face) {
    case 0: return -1.0 + 2.0 * i;
    case 1: return -1.0 + 2.0 * i;
    case 2: return 1.0;
    case 3: return 1.0 - 2.0 * i;
    case 4: return -1.0;
    default: return -1.0 + 2.0 * i;
  }
}
double orc_cm_cube_map_y(int face, double j, double i) {
  switch (face) {
    case 0: return 1.0 - 2.0 * j;
    case 1: return -1.0;
    case 2: return -1.0 + 2.0 * i;
    case 3: return 1.0;
    case 4: return 1.0 - 2.0 * i;
    default: return -1.0 + 2.0 * j;
  }
}
double orc_cm_cube_map_z(int face, double j, double i) {
  (void)i;
  switch (face) {
    case 0: return 1.0;
    case 1: return 1.0 - 2.0 * j;
    case 2: return 1.0 - 2.0 * j;
    case 3: return 1.0 - 2.0 * j;
    case 4: return 1.0 - 2.0 * j;
    default: return -1.0;
  }
}
void orc_cm_cube_map(int face, double j, double i, double out[3]) {
  out[0] = orc_cm_cube_map_x(face, j, i);
  out[1] = orc_cm_cube_map_y(face, j, i);
  out[2] = orc_cm_cube_map_z(face, j, i);
}

/* cubemap.clj:70-78 */
int orc_cm_determine_face(const double p[3]) {
  double x = p[0], y = p[1], z = p[2];
  if (fabs(x) >= fmax(fabs(y), fabs(z))) return x >= 0 ? 2 : 4;
  if (fabs(y) >= fmax(fabs(x), fabs(z))) return y >= 0 ? 3 : 1;
  return z >= 0 ? 0 : 5;
}
/* cubemap.clj:81-91 */
double orc_cm_cube_i(int face, const double p[3]) {
  switch (face) {
    case 0: return 0.5 * (p[0] + 1.0);
    case 1: return 0.5 * (p[0] + 1.0);
    case 2: return 0.5 * (p[1] + 1.0);
    case 3: return 0.5 * (1.0 - p[0]);
    case 4: return 0.5 * (1.0 - p[1]);
    default: return 0.5 * (p[0] + 1.0);
  }
}
/* cubemap.clj:94-104 */
double orc_cm_cube_j(int face, const double p[3]) {
  switch (face) {
    case 0: return 0.5 * (1.0 - p[1]);
    case 1: return 0.5 * (1.0 - p[2]);
    case 2: return 0.5 * (1.0 - p[2]);
    case 3: return 0.5 * (1.0 - p[2]);
    case 4: return 0.5 * (1.0 - p[2]);
    default: return 0.5 * (p[1] + 1.0);
  }
}
/* cubemap.clj:107-111 */
double orc_cm_cube_coordinate(long level, long tilesize, long tile, double pixel) {
  long tiles = 1L << level;
  return ((double)tile + pixel / (double)(tilesize - 1)) / (double)tiles;
}
/* cubemap.clj:114-120 */
void orc_cm_cube_map_corners(int face, long level, long row, long column, double out[4][3]) {
  orc_cm_cube_map(face, orc_cm_cube_coordinate(level, 2, row, 0.0), orc_cm_cube_coordinate(level, 2, column, 0.0), out[0]);
  orc_cm_cube_map(face, orc_cm_cube_coordinate(level, 2, row, 0.0), orc_cm_cube_coordinate(level, 2, column, 1.0), out[1]);
  orc_cm_cube_map(face, orc_cm_cube_coordinate(level, 2, row, 1.0), orc_cm_cube_coordinate(level, 2, column, 0.0), out[2]);
  orc_cm_cube_map(face, orc_cm_cube_coordinate(level, 2, row, 1.0), orc_cm_cube_coordinate(level, 2, column, 1.0), out[3]);
}

/* ------------------------------------------------------------------ geodetic conversions (cubemap.clj:123-168) */
double orc_cm_longitude(const double p[3]) { return atan2(p[1], p[0]); }
double orc_cm_latitude(const double p[3]) {
  double q = sqrt(sqr(p[0]) + sqr(p[1]));
  return atan2(p[2], q);
}
void orc_cm_geodetic_to_cartesian(double longitude, double latitude, double height, double radius, double out[3]) {
  double distance = height + radius;
  double cos_lat = cos(latitude), sin_lat = sin(latitude);
  out[0] = distance * cos_lat * cos(longitude);
  out[1] = distance * cos_lat * sin(longitude);
  out[2] = distance * sin_lat;
}
void orc_cm_project_onto_sphere(const double p[3], double radius, double out[3]) {
  double n[3];
  normalize3(p, n);
  out[0] = n[0] * radius;
  out[1] = n[1] * radius;
  out[2] = n[2] * radius;
}
void orc_cm_project_onto_cube(const double p[3], double out[3]) {
  double ax = fabs(p[0]), ay = fabs(p[1]), az = fabs(p[2]);
  double d = ax >= fmax(ay, az) ? ax : (ay >= fmax(ax, az) ? ay : az);
  out[0] = p[0] / d;
  out[1] = p[1] / d;
  out[2] = p[2] / d;
}
/* returns [longitude latitude height] */
void orc_cm_cartesian_to_geodetic(const double p[3], double radius, double out[3]) {
  double height = mag3(p) - radius;
  double longitude = atan2(p[1], p[0]);
  double q = sqrt(sqr(p[0]) + sqr(p[1]));
  double latitude = atan2(p[2], q);
  out[0] = longitude;
  out[1] = latitude;
  out[2] = height;
}

/* ------------------------------------------------------------------ raster coordinates (cubemap.clj:171-207) */
double orc_cm_map_x(double longitude, long tilesize, long level) {
  long n = 1L << level;
  return (PI + longitude) * ((double)(4 * n * tilesize) / (2 * PI));
}
double orc_cm_map_y(double latitude, long tilesize, long level) {
  long n = 1L << level;
  return (PI / 2 - latitude) * ((double)(2 * n * tilesize) / PI);
}
static long floor_mod(long a, long b) {
  long m = a % b;
  return m < 0 ? m + b : m;
}
/* out = [x0 x1], frac = [frac0 frac1] */
void orc_cm_map_pixels_x(double longitude, long tilesize, long level, long out[2], double frac[2]) {
  long n = 1L << level;
  long size = 4 * n * tilesize;
  double x = orc_cm_map_x(longitude, tilesize, level);
  long x0 = (long)(int)floor(x);
  long x1 = x0 + 1;
  double frac1 = x - (double)x0;
  double frac0 = 1 - frac1;
  out[0] = floor_mod(x0, size);
  out[1] = floor_mod(x1, size);
  frac[0] = frac0;
  frac[1] = frac1;
}
void orc_cm_map_pixels_y(double latitude, long tilesize, long level, long out[2], double frac[2]) {
  long n = 1L << level;
  long size = 2 * n * tilesize;
  double y = orc_cm_map_y(latitude, tilesize, level);
  long y0 = (long)(int)floor(y);
  long y1 = y0 + 1;
  double frac1 = y - (double)y0;
  double frac0 = 1 - frac1;
  out[0] = y0 < size - 1 ? y0 : size - 1;
  out[1] = y1 < size - 1 ? y1 : size - 1;
  frac[0] = frac0;
  frac[1] = frac1;
}

/* ------------------------------------------------------------------ offsets for the normal estimate (cubemap.clj:210-229) */
void orc_cm_offset_longitude(const double p[3], long level, long tilesize, double out[3]) {
  double lon = orc_cm_longitude(p);
  double norm = mag3(p);
  double v[3] = {0, (norm * PI) / (double)(2 * tilesize * (1L << level)), 0};
  double m[9];
  rotation_z(lon, m);
  mulv3(m, v, out);
}
void orc_cm_offset_latitude(const double p[3], long level, long tilesize, double out[3]) {
  double lon = orc_cm_longitude(p);
  double lat = orc_cm_latitude(p);
  double norm = mag3(p);
  double v[3] = {0, 0, (norm * PI) / (double)(2 * tilesize * (1L << level))};
  double my[9], mz[9], t[3];
  rotation_y(-lat, my);
  mulv3(my, v, t);
  rotation_z(lon, mz);
  mulv3(mz, t, out);
}

/* ------------------------------------------------------------------ raster access (cubemap.clj:232-286) */
static size_t tile_offset(long dy, long dx, long level, long width) {
  long ty = dy / width, tx = dx / width, py = dy % width, px = dx % width;
  long tiles_x = 4 * (1L << level);
  return (((size_t)ty * tiles_x + tx) * width + py) * width + px;
}
/* world-map-pixel: RGB of pixel (dy, dx) of the level's raster (image.clj:182-188 get-pixel) */
void orc_cm_world_map_pixel(const orc_world *w, int night, long dy, long dx, long level, double out[3]) {
  const uint8_t *img = night ? w->night[level] : w->day[level];
  size_t o = 4 * tile_offset(dy, dx, level, w->width);
  out[0] = img[o];
  out[1] = img[o + 1];
  out[2] = img[o + 2];
}
long orc_cm_elevation_pixel(const orc_world *w, long dy, long dx, long level) {
  return w->elevation[level][tile_offset(dy, dx, level, w->width)];
}
/* the weighted sum of map-interpolation (cubemap.clj:289-299): v0 (yfrac0 xfrac0) + v1 (yfrac0 xfrac1) + ... */
double orc_cm_interpolate4(double v0, double v1, double v2, double v3, const double xfrac[2], const double yfrac[2]) {
  return ((v0 * (yfrac[0] * xfrac[0]) + v1 * (yfrac[0] * xfrac[1])) + v2 * (yfrac[1] * xfrac[0])) + v3 * (yfrac[1] * xfrac[1]);
}

/* tile-center (cubemap.clj:302-308) */
void orc_cm_tile_center(int face, long level, long row, long column, double radius, double out[3]) {
  double j = orc_cm_cube_coordinate(level, 3, row, 1.0);
  double i = orc_cm_cube_coordinate(level, 3, column, 1.0);
  double p[3];
  orc_cm_cube_map(face, j, i, p);
  orc_cm_project_onto_sphere(p, radius, out);
}

/* color-geodetic-day / -night (cubemap.clj:311-320) */
void orc_cm_color_geodetic(const orc_world *w, int night, long in_level, double lon, double lat, double out[3]) {
  long dx[2], dy[2];
  double xf[2], yf[2], v[4][3];
  orc_cm_map_pixels_x(lon, w->width, in_level, dx, xf);
  orc_cm_map_pixels_y(lat, w->width, in_level, dy, yf);
  orc_cm_world_map_pixel(w, night, dy[0], dx[0], in_level, v[0]);
  orc_cm_world_map_pixel(w, night, dy[0], dx[1], in_level, v[1]);
  orc_cm_world_map_pixel(w, night, dy[1], dx[0], in_level, v[2]);
  orc_cm_world_map_pixel(w, night, dy[1], dx[1], in_level, v[3]);
  for (int c = 0; c < 3; c++) out[c] = orc_cm_interpolate4(v[0][c], v[1][c], v[2][c], v[3][c], xf, yf);
}
/* elevation-geodetic (cubemap.clj:323-326) */
double orc_cm_elevation_geodetic(const orc_world *w, long in_level, double lon, double lat) {
  long dx[2], dy[2];
  double xf[2], yf[2];
  orc_cm_map_pixels_x(lon, w->width, in_level, dx, xf);
  orc_cm_map_pixels_y(lat, w->width, in_level, dy, yf);
  return orc_cm_interpolate4((double)orc_cm_elevation_pixel(w, dy[0], dx[0], in_level),
                             (double)orc_cm_elevation_pixel(w, dy[0], dx[1], in_level),
                             (double)orc_cm_elevation_pixel(w, dy[1], dx[0], in_level),
                             (double)orc_cm_elevation_pixel(w, dy[1], dx[1], in_level), xf, yf);
}
/* water-geodetic from a height (cubemap.clj:329-333) */
long orc_cm_water_from_height(double height) { return height < 0 ? (long)(int)((height * 255) / -500) : 0; }
long orc_cm_water_geodetic(const orc_world *w, long in_level, double lon, double lat) {
  return orc_cm_water_from_height(orc_cm_elevation_geodetic(w, in_level, lon, lat));
}
/* project-onto-globe (cubemap.clj:336-342) */
void orc_cm_project_onto_globe(const orc_world *w, const double p[3], long in_level, double radius, double out[3]) {
  double sp[3], g[3];
  orc_cm_project_onto_sphere(p, radius, sp);
  orc_cm_cartesian_to_geodetic(sp, radius, g);
  double height = fmax(orc_cm_elevation_geodetic(w, in_level, g[0], g[1]), 0.0);
  orc_cm_geodetic_to_cartesian(g[0], g[1], height, radius, out);
}
/* the nine unprojected points of surrounding-points (cubemap.clj:345-354): dj outer, di inner */
void orc_cm_surrounding_offsets(const double p[3], const double d1[3], const double d2[3], double out[9][3]) {
  int k = 0;
  for (int dj = -1; dj <= 1; dj++)
    for (int di = -1; di <= 1; di++, k++)
      for (int c = 0; c < 3; c++) out[k][c] = p[c] + (d2[c] * dj + d1[c] * di);
}
void orc_cm_surrounding_points(const orc_world *w, const double p[3], long in_level, long out_level, long tilesize,
                               double radius, double out[9][3]) {
  double d1[3], d2[3], ps[9][3];
  orc_cm_offset_longitude(p, out_level, tilesize, d1);
  orc_cm_offset_latitude(p, out_level, tilesize, d2);
  orc_cm_surrounding_offsets(p, d1, d2, ps);
  for (int k = 0; k < 9; k++) orc_cm_project_onto_globe(w, ps[k], in_level, radius, out[k]);
}
/* the Sobel part of normal-for-point (cubemap.clj:357-366) */
void orc_cm_normal_from_points(const double pc[9][3], double out[3]) {
  static const double sx[9] = {-0.25, 0, 0.25, -0.5, 0, 0.5, -0.25, 0, 0.25};
  static const double sy[9] = {-0.25, -0.5, -0.25, 0, 0, 0, 0.25, 0.5, 0.25};
  double n1[3], n2[3];
  for (int c = 0; c < 3; c++) {
    n1[c] = pc[0][c] * sx[0];
    n2[c] = pc[0][c] * sy[0];
    for (int k = 1; k < 9; k++) {
      n1[c] = n1[c] + pc[k][c] * sx[k];
      n2[c] = n2[c] + pc[k][c] * sy[k];
    }
  }
  double cr[3] = {n1[1] * n2[2] - n1[2] * n2[1], n1[2] * n2[0] - n1[0] * n2[2], n1[0] * n2[1] - n1[1] * n2[0]};
  normalize3(cr, out);
}
void orc_cm_normal_for_point(const orc_world *w, const double p[3], long in_level, long out_level, long tilesize,
                             double radius, double out[3]) {
  double pc[9][3];
  orc_cm_surrounding_points(w, p, in_level, out_level, tilesize, radius, pc);
  orc_cm_normal_from_points(pc, out);
}

/* ------------------------------------------------------------------ make-cube-map, one tile (globe.clj:29-80) */
static long clampl(long x, long lo, long hi) { return x < lo ? lo : (x > hi ? hi : x); }
static long align_address(long a, long alignment) { return (a + alignment - 1) & ~(alignment - 1); } /* util.clj:367-371 */
static uint8_t ubyte(long v) { return (uint8_t)(v & 255); }                                           /* ubyte->byte, as stored */

/* spit-normals (image.clj:126-136): round(x 127.5 - 0.5) stored as a signed byte */
int8_t orc_cm_normal_byte(float x) { return (int8_t)(long)floor(((double)x * 127.5 - 0.5) + 0.5); }

/* Outputs (the arrays the reference hands to spit-jpg / spit-bytes-gz / spit-floats-gz / spit-normals):
 *   day, night: uint8 [ct][ct][4] (alpha 255), water: uint8 [ct][align4(ct)], surface: float [st][st][3],
 *   normals: float [ct][ct][3]; ct = 2 (st - 1) + 1.
 * raw (optional, test diagnostics): double [ct][ct][7] = the day and night colours and the water value BEFORE their
 * truncation to integers -- where one of them sits on an integer (constant raster regions: v (w0 + w1 + w2 + w3) with
 * weights that sum to 1 - 1e-16 or to 1), the byte is rounding noise of whichever libm produced lon / lat. */
void orc_cm_make_cube_map_tile(const orc_world *w, int face, long in_level, long out_level, long b, long a,
                               long surface_tilesize, long max_surface_level, long max_color_level, double radius,
                               uint8_t *day, uint8_t *night, uint8_t *water, float *surface, float *normals, double *raw) {
  const long sublevel = 1; /* max_surface_level = 4, max_color_level = 5 in globe.clj:36-37 */
  const long subsample = 1L << sublevel;
  const long ct = subsample * (surface_tilesize - 1) + 1;
  const long wpitch = align_address(ct, 4);
  const long ls = clampl(in_level, 0, max_surface_level);
  const long lc = clampl(in_level + sublevel, 0, max_color_level);
  const long lw = clampl(in_level + sublevel, 0, max_surface_level);
  double center[3];
  orc_cm_tile_center(face, out_level, b, a, radius, center);
  memset(water, 0, (size_t)(wpitch * ct));
#pragma omp parallel for schedule(dynamic, 1)
  for (long v = 0; v < surface_tilesize; v++)
    for (long u = 0; u < surface_tilesize; u++) {
      double j = orc_cm_cube_coordinate(out_level, surface_tilesize, b, (double)v);
      double i = orc_cm_cube_coordinate(out_level, surface_tilesize, a, (double)u);
      double p[3], point[3];
      orc_cm_cube_map(face, j, i, p);
      orc_cm_project_onto_globe(w, p, ls, radius, point);
      for (int c = 0; c < 3; c++) surface[(v * surface_tilesize + u) * 3 + c] = (float)(point[c] - center[c]);
    }
#pragma omp parallel for schedule(dynamic, 1)
  for (long v = 0; v < ct; v++)
    for (long u = 0; u < ct; u++) {
      double j = orc_cm_cube_coordinate(out_level, ct, b, (double)v);
      double i = orc_cm_cube_coordinate(out_level, ct, a, (double)u);
      double p[3], point[3], g[3], normal[3], cd[3], cn[3];
      orc_cm_cube_map(face, j, i, p);
      orc_cm_project_onto_globe(w, p, ls, radius, point);
      orc_cm_cartesian_to_geodetic(point, radius, g);
      orc_cm_normal_for_point(w, point, ls, out_level, ct, radius, normal);
      orc_cm_color_geodetic(w, 0, lc, g[0], g[1], cd);
      orc_cm_color_geodetic(w, 1, lc, g[0], g[1], cn);
      double wet_height = orc_cm_elevation_geodetic(w, lw, g[0], g[1]);
      long wet = orc_cm_water_from_height(wet_height);
      if (raw) {
        double *r = raw + (v * ct + u) * 7;
        for (int c = 0; c < 3; c++) {
          r[c] = cd[c];
          r[3 + c] = cn[c];
        }
        r[6] = wet_height < 0 ? (wet_height * 255) / -500 : 0.0;
      }
      for (int c = 0; c < 3; c++) {
        normals[(v * ct + u) * 3 + c] = (float)normal[c];
        day[(v * ct + u) * 4 + c] = ubyte((long)cd[c]);     /* set-pixel! image.clj:191-199: (long (c k)) */
        night[(v * ct + u) * 4 + c] = ubyte((long)cn[c]);
      }
      day[(v * ct + u) * 4 + 3] = 255;
      night[(v * ct + u) * 4 + 3] = 255;
      water[v * wpitch + u] = ubyte(wet);
    }
}
