/*
 * 3-D noise textures of sfsim on the GPU -- part of libsfsim_atmosphere.so.
 *
 * Drop-in for the sampling loops of `clj -T:build worley`, `perlin` and `bluenoise` (build.clj:34-51 ->
 * src/clj/sfsim/worley.clj:95-112, src/clj/sfsim/perlin.clj:122-137, src/clj/sfsim/bluenoise.clj:175-185).  The reference draws its random point /
 * gradient grid with clojure.core/rand inside those functions; here the grid is an INPUT, so a host that wants the
 * reference's exact texture passes the grid it drew (the reference's own tests rebind random-point-grid and
 * random-gradient-grid the same way, t_worley.clj:70-74, t_perlin.clj:140-146).  All arithmetic is IEEE double in
 * the reference's operation order; outputs are float32 like `(float-array ...)` before spit-floats.
 *
 * Conventions as in sfsim_atmosphere.h: int status (0 = ok), atmlut_last_error() for the message, no CPU fallback.
 */
#ifndef SFSIM_NOISE_H
#define SFSIM_NOISE_H

#ifdef __cplusplus
extern "C" {
#endif

/* worley-noise (worley.clj:95-112).  grid: double[divisions][divisions][divisions][3], grid[k][j][i] = the random
 * point of cell (i, j, k) as random-point-grid (worley.clj:27-44) builds it.  out: float[size^3], element
 * (k * size + j) * size + i = 1 - d / max(d), d = distance from (k + 1/2, j + 1/2, i + 1/2) to the closest grid point
 * with periodic wrap-around.  size must be a multiple of divisions. */
int sfsim_worley_noise(const double *grid, int divisions, int size, float *out);

/* perlin-noise (perlin.clj:122-137).  gradients: double[divisions][divisions][divisions][3] as random-gradient-grid
 * (perlin.clj:38-50) builds it ([z][y][x]).  out: float[size^3], element (k * size + j) * size + i = the sample at cell
 * (i + 1/2, j + 1/2, k + 1/2), normalised to [0, 1] by (v - min) / (max - min). */
int sfsim_perlin_noise(const double *gradients, int divisions, int size, float *out);

/* the un-normalised samples (closest distances / Perlin sums) in double, for parity checks */
int sfsim_worley_distances(const double *grid, int divisions, int size, double *out);
int sfsim_perlin_samples(const double *gradients, int divisions, int size, double *out);

/* blue-noise (bluenoise.clj:175-185), the void-and-cluster dither array of `clj -T:build bluenoise` (build.clj:45-51).
 * picks: the n distinct indices pick-n (bluenoise.clj:35-40) drew from 0 .. size^2 - 1; ftab: the density function
 * over the wrapped offsets, ftab[(dy + size/2) * size + (dx + size/2)] = f(dx, dy) (density-function :53-56);
 * dither: int[size^2], every rank 0 .. size^2 - 1 exactly once.  The insertions are serial (each needs the arg-max or
 * arg-min of the density array the previous one left behind); each of them runs on 1024 threads of one CTA. */
int sfsim_blue_noise(const int *picks, int n, int size, const double *ftab, int *dither);
/* the float array build.clj:45-51 writes: dither / size / size, with density-function(sigma) tabulated by the host */
int sfsim_blue_noise_texture(const int *picks, int n, int size, double sigma, float *out);

#ifdef __cplusplus
}
#endif
#endif
