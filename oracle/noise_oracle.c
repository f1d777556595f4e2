/* TEST INFRASTRUCTURE ONLY -- CPU restatement of sfsim's 3-D noise texture precompute (`clj -T:build worley`,
 * `clj -T:build perlin`, build.clj:34-43) for checking the CUDA kernels.  Plain double-precision C, one function per
 * reference function, each citing the file:line it follows (wedesoft/sfsim, src/clj/sfsim/).  Pinned to the known
 * answers of test/clj/sfsim/t_worley.clj and t_perlin.clj by tests/test_noise_oracle.py.  Nothing in the product path
 * may include, link or call this file.
 *
 * The random grids (random-point-grid worley.clj:27-44, random-gradient-grid perlin.clj:38-50) are INPUTS here and in
 * the library, so that results can be compared value for value. */
#include <math.h>
#include <stdlib.h>

/* worley.clj:47-55 clipped-index-and-offset */
static void clipped_index_and_offset(long divisions, long size, long index, long *clipped, double *offset) {
  *clipped = index < divisions ? (index >= 0 ? index : index + divisions) : index - divisions;
  *offset = index < divisions ? (index >= 0 ? 0.0 : -(double)size) : (double)size;
}

/* worley.clj:58-66 extract-point-from-grid; grid is double[dims[0]][dims[1]][dims[2]][3], indexed [k][j][i]
 * (dimension-count of the nested vectors: the reference's tests use ragged 1 x 1 x 2 grids) */
void orc_extract_point_from_grid(const double *grid, const long dims[3], long size, long k, long j, long i, double out[3]) {
  long ic, jc, kc;
  double xo, yo, zo;
  clipped_index_and_offset(dims[2], size, i, &ic, &xo);
  clipped_index_and_offset(dims[1], size, j, &jc, &yo);
  clipped_index_and_offset(dims[0], size, k, &kc, &zo);
  const double *p = grid + ((kc * dims[1] + jc) * dims[2] + ic) * 3;
  out[0] = p[0] + xo;
  out[1] = p[1] + yo;
  out[2] = p[2] + zo;
}

/* worley.clj:69-80 closest-distance-to-point-in-grid */
double orc_closest_distance_to_point_in_grid(const double *grid, const long dims[3], long divisions, long size,
                                             const double point[3]) {
  double cellsize = (double)size / (double)divisions;
  long i = (long)trunc(point[0] / cellsize), j = (long)trunc(point[1] / cellsize), k = (long)trunc(point[2] / cellsize);
  double best = INFINITY;
  for (long dk = -1; dk <= 1; dk++)
    for (long dj = -1; dj <= 1; dj++)
      for (long di = -1; di <= 1; di++) {
        double p[3];
        orc_extract_point_from_grid(grid, dims, size, k + dk, j + dj, i + di, p);
        double dx = point[0] - p[0], dy = point[1] - p[1], dz = point[2] - p[2];
        double d = sqrt(dx * dx + dy * dy + dz * dz);
        if (d < best) best = d;
      }
  return best;
}

/* worley.clj:95-112 worley-noise: sample (k, j, i) sits at the point (k + 1/2, j + 1/2, i + 1/2) -- x runs with the
 * OUTERMOST index -- then normalize-vector (:83-88, divide by the maximum) and invert-vector (:91-92). */
void orc_worley_noise(const double *grid, long divisions, long size, double *out) {
  long n = size * size * size;
  const long dims[3] = {divisions, divisions, divisions};
  double maximum = -INFINITY;
  for (long k = 0; k < size; k++)
    for (long j = 0; j < size; j++)
      for (long i = 0; i < size; i++) {
        double point[3] = {(double)k + 0.5, (double)j + 0.5, (double)i + 0.5};
        double d = orc_closest_distance_to_point_in_grid(grid, dims, divisions, size, point);
        out[(k * size + j) * size + i] = d;
        if (d > maximum) maximum = d;
      }
  for (long t = 0; t < n; t++) out[t] = 1.0 - out[t] / maximum;
}

/* perlin.clj:83-86 ease-curve */
double orc_ease_curve(double t) { return ((((t * 6.0) - 15.0) * t) + 10.0) * t * t * t; }

/* perlin.clj:111-119 perlin-noise-sample with corner-vectors :60-66, corner-gradients :69-76, influence-values :79-82,
 * interpolation-weights :89-99; gradients is double[divisions][divisions][divisions][3], indexed [z][y][x] */
double orc_perlin_noise_sample(const double *gradients, long divisions, long size, const double cell[3]) {
  double scale = (double)divisions / (double)size;
  double point[3] = {cell[0] * scale, cell[1] * scale, cell[2] * scale};
  double division[3] = {floor(point[0]), floor(point[1]), floor(point[2])};
  long c[3] = {(long)division[0], (long)division[1], (long)division[2]};
  long cp[3];
  for (int a = 0; a < 3; a++) {
    cp[a] = (c[a] + 1) % divisions;
    if (cp[a] < 0) cp[a] += divisions;
  }
  double b[3] = {point[0] - division[0], point[1] - division[1], point[2] - division[2]};
  double a1[3] = {1.0 - b[0], 1.0 - b[1], 1.0 - b[2]};
  double sum = 0.0;
  int first = 1;
  for (int z = 0; z < 2; z++)
    for (int y = 0; y < 2; y++)
      for (int x = 0; x < 2; x++) {
        double corner[3] = {point[0] - (division[0] + x), point[1] - (division[1] + y), point[2] - (division[2] + z)};
        const double *g = gradients + (((z ? cp[2] : c[2]) * divisions + (y ? cp[1] : c[1])) * divisions + (x ? cp[0] : c[0])) * 3;
        double influence = g[0] * corner[0] + g[1] * corner[1] + g[2] * corner[2];
        double weight = orc_ease_curve(z ? b[2] : a1[2]) * orc_ease_curve(y ? b[1] : a1[1]) * orc_ease_curve(x ? b[0] : a1[0]);
        double term = weight * influence;
        sum = first ? term : sum + term;
        first = 0;
      }
  return sum;
}

/* perlin.clj:122-137 perlin-noise: sample (k, j, i) at the cell (i + 1/2, j + 1/2, k + 1/2), then normalize-vector
 * (:102-108: (v - min) / (max - min)) */
void orc_perlin_noise(const double *gradients, long divisions, long size, double *out) {
  long n = size * size * size;
  double minimum = INFINITY, maximum = -INFINITY;
  for (long k = 0; k < size; k++)
    for (long j = 0; j < size; j++)
      for (long i = 0; i < size; i++) {
        double cell[3] = {(double)i + 0.5, (double)j + 0.5, (double)k + 0.5};
        double v = orc_perlin_noise_sample(gradients, divisions, size, cell);
        out[(k * size + j) * size + i] = v;
        if (v < minimum) minimum = v;
        if (v > maximum) maximum = v;
      }
  for (long t = 0; t < n; t++) out[t] = (out[t] - minimum) / (maximum - minimum);
}

/* ------------------------------------------------------------------ blue noise (bluenoise.clj), void-and-cluster
 *
 * The density function f(dx, dy) = exp(-(dx^2 + dy^2) / (2 sigma^2)) (bluenoise.clj:53-56) is passed as a TABLE
 * ftab[(dy + m/2) * m + (dx + m/2)] over the wrapped offsets, and the random seed picks (pick-n, :35-40) as a list,
 * so that an implementation can be compared decision for decision: everything else is additions, subtractions and
 * comparisons of doubles. */

/* bluenoise.clj:73-78 wrap */
long orc_wrap(long x, long m) {
  long offset = m / 2;
  long r = (x + offset) % m;
  if (r < 0) r += m;
  return r - offset;
}

static double ftab_at(const double *ftab, long m, long dx, long dy) {
  long offset = m / 2;
  return ftab[(orc_wrap(dy, m) + offset) * m + (orc_wrap(dx, m) + offset)];
}

/* bluenoise.clj:59-63 argmax-with-mask: largest element whose mask is true; max-key keeps the LAST of equal maxima */
long orc_argmax_with_mask(const double *arr, const unsigned char *mask, long count) {
  long best = -1;
  for (long i = 0; i < count; i++)
    if (mask[i] && (best < 0 || arr[i] >= arr[best])) best = i;
  return best;
}

/* bluenoise.clj:66-70 argmin-with-mask: smallest element whose mask is false; min-key keeps the LAST of equal minima */
long orc_argmin_with_mask(const double *arr, const unsigned char *mask, long count) {
  long best = -1;
  for (long i = 0; i < count; i++)
    if (!mask[i] && (best < 0 || arr[i] <= arr[best])) best = i;
  return best;
}

/* bluenoise.clj:81-91 density-sample: sum in (y, x) order over the set mask entries */
double orc_density_sample(const unsigned char *mask, long m, const double *ftab, long cx, long cy) {
  double sum = 0.0;
  for (long y = 0; y < m; y++)
    for (long x = 0; x < m; x++)
      sum = sum + (mask[y * m + x] ? ftab_at(ftab, m, x - cx, y - cy) : 0.0);
  return sum;
}

/* bluenoise.clj:94-98 density-array */
void orc_density_array(const unsigned char *mask, long m, const double *ftab, double *out) {
  for (long cy = 0; cy < m; cy++)
    for (long cx = 0; cx < m; cx++) out[cy * m + cx] = orc_density_sample(mask, m, ftab, cx, cy);
}

/* bluenoise.clj:101-111 density-change, in place; sign = +1 / -1 for the reference's `+` / `-` */
void orc_density_change(double *density, long m, int sign, const double *ftab, long index) {
  long cy = index / m, cx = index % m;
  for (long y = 0; y < m; y++)
    for (long x = 0; x < m; x++) {
      double f = ftab_at(ftab, m, x - cx, y - cy);
      density[y * m + x] = sign > 0 ? density[y * m + x] + f : density[y * m + x] - f;
    }
}

/* bluenoise.clj:114-126 seed-pattern (mask modified in place) */
void orc_seed_pattern(unsigned char *mask, long m, const double *ftab) {
  double *density = (double *)malloc((size_t)(m * m) * sizeof(double));
  orc_density_array(mask, m, ftab, density);
  for (;;) {
    long cluster = orc_argmax_with_mask(density, mask, m * m);
    mask[cluster] = 0;
    orc_density_change(density, m, -1, ftab, cluster);
    long hole = orc_argmin_with_mask(density, mask, m * m);
    mask[hole] = 1;
    if (cluster == hole) break;
    orc_density_change(density, m, +1, ftab, hole);
  }
  free(density);
}

/* bluenoise.clj:129-143 dither-phase1: mask is NOT modified for the caller (the reference's is immutable) */
void orc_dither_phase1(const unsigned char *mask_in, long m, long n, const double *ftab, long *dither) {
  long count = m * m;
  unsigned char *mask = (unsigned char *)malloc((size_t)count);
  double *density = (double *)malloc((size_t)count * sizeof(double));
  for (long i = 0; i < count; i++) {
    mask[i] = mask_in[i];
    dither[i] = 0;
  }
  orc_density_array(mask, m, ftab, density);
  while (n > 0) {
    long cluster = orc_argmax_with_mask(density, mask, count);
    orc_density_change(density, m, -1, ftab, cluster);
    mask[cluster] = 0;
    n--;
    dither[cluster] = n;
  }
  free(mask);
  free(density);
}

/* bluenoise.clj:146-156 dither-phase2: fills the mask (in place) until half of it is set */
void orc_dither_phase2(unsigned char *mask, long m, long n, long *dither, const double *ftab) {
  long count = m * m;
  double *density = (double *)malloc((size_t)count * sizeof(double));
  orc_density_array(mask, m, ftab, density);
  while (n < count / 2) {
    long hole = orc_argmin_with_mask(density, mask, count);
    orc_density_change(density, m, +1, ftab, hole);
    mask[hole] = 1;
    dither[hole] = n;
    n++;
  }
  free(density);
}

/* bluenoise.clj:159-172 dither-phase3 */
void orc_dither_phase3(const unsigned char *mask, long m, long n, long *dither, const double *ftab) {
  long count = m * m;
  unsigned char *mask_not = (unsigned char *)malloc((size_t)count);
  double *density = (double *)malloc((size_t)count * sizeof(double));
  for (long i = 0; i < count; i++) mask_not[i] = !mask[i];
  orc_density_array(mask_not, m, ftab, density);
  while (n < count) {
    long cluster = orc_argmax_with_mask(density, mask_not, count);
    orc_density_change(density, m, -1, ftab, cluster);
    mask_not[cluster] = 0;
    dither[cluster] = n;
    n++;
  }
  free(mask_not);
  free(density);
}

/* bluenoise.clj:175-185 blue-noise: picks = the n indices pick-n drew */
void orc_blue_noise(long m, long n, const long *picks, const double *ftab, long *dither) {
  long count = m * m;
  unsigned char *seed = (unsigned char *)calloc((size_t)count, 1);
  for (long i = 0; i < n; i++) seed[picks[i]] = 1;               /* scatter-mask :46-49 */
  orc_seed_pattern(seed, m, ftab);
  orc_dither_phase1(seed, m, n, ftab, dither);
  unsigned char *half = (unsigned char *)malloc((size_t)count);
  for (long i = 0; i < count; i++) half[i] = seed[i];
  orc_dither_phase2(half, m, n, dither, ftab);
  orc_dither_phase3(half, m, count / 2, dither, ftab);
  free(seed);
  free(half);
}
