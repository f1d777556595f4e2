// 3-D Worley and Perlin noise textures (include/sfsim_noise.h): one thread per texture sample, double precision in
// the reference's operation order (the translation unit is compiled with -fmad=false), a device-wide min / max by
// atomics on order-preserving integer keys, and a second pass that normalises and casts to float32.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "sfsim_noise.h"
#include "atm_api_internal.h"

namespace atm {

// ------------------------------------------------------------------ order-preserving keys for atomic min / max

__device__ __forceinline__ unsigned long long ordered_key(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double from_ordered_key(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__device__ void block_min_max(double v, bool valid, unsigned long long *extrema) {
  __shared__ unsigned long long s_min, s_max;
  if (threadIdx.x == 0) {
    s_min = ~0ull;
    s_max = 0ull;
  }
  __syncthreads();
  if (valid) {
    const unsigned long long k = ordered_key(v);
    atomicMin(&s_min, k);
    atomicMax(&s_max, k);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicMin(&extrema[0], s_min);
    atomicMax(&extrema[1], s_max);
  }
}

// ------------------------------------------------------------------ Worley (worley.clj)

// worley.clj:47-55 clipped-index-and-offset
__device__ __forceinline__ void clipped_index_and_offset(int divisions, int size, int index, int &clipped, double &offset) {
  clipped = index < divisions ? (index >= 0 ? index : index + divisions) : index - divisions;
  offset = index < divisions ? (index >= 0 ? 0.0 : -(double)size) : (double)size;
}

// worley.clj:69-80 closest-distance-to-point-in-grid with extract-point-from-grid :58-66
__device__ double closest_distance(const double *__restrict__ grid, int divisions, int size, double x, double y, double z) {
  const double cellsize = (double)size / (double)divisions;
  const int i = (int)trunc(x / cellsize), j = (int)trunc(y / cellsize), k = (int)trunc(z / cellsize);
  double best = INFINITY;
  for (int dk = -1; dk <= 1; dk++)
    for (int dj = -1; dj <= 1; dj++)
      for (int di = -1; di <= 1; di++) {
        int ic, jc, kc;
        double xo, yo, zo;
        clipped_index_and_offset(divisions, size, i + di, ic, xo);
        clipped_index_and_offset(divisions, size, j + dj, jc, yo);
        clipped_index_and_offset(divisions, size, k + dk, kc, zo);
        const double *p = grid + ((size_t)(kc * divisions + jc) * divisions + ic) * 3;
        const double dx = x - (p[0] + xo), dy = y - (p[1] + yo), dz = z - (p[2] + zo);
        const double d = sqrt(dx * dx + dy * dy + dz * dz);
        best = d < best ? d : best;
      }
  return best;
}

// worley.clj:104-112: sample (k, j, i) sits at (k + 1/2, j + 1/2, i + 1/2) -- x runs with the outermost index
__global__ void k_worley_distances(const double *__restrict__ grid, int divisions, int size, double *raw,
                                   unsigned long long *extrema) {
  const long long n = (long long)size * size * size;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double d = 0.0;
  if (t < n) {
    const int i = (int)(t % size), j = (int)((t / size) % size), k = (int)(t / ((long long)size * size));
    d = closest_distance(grid, divisions, size, (double)k + 0.5, (double)j + 0.5, (double)i + 0.5);
    raw[t] = d;
  }
  block_min_max(d, t < n, extrema);
}

// normalize-vector (worley.clj:83-88) then invert-vector (:91-92)
__global__ void k_worley_finish(const double *__restrict__ raw, long long n, const unsigned long long *extrema, float *out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double maximum = from_ordered_key(extrema[1]);
  out[t] = (float)(1.0 - raw[t] / maximum);
}

// ------------------------------------------------------------------ Perlin (perlin.clj)

// perlin.clj:83-86 ease-curve
__device__ __forceinline__ double ease_curve(double t) { return ((((t * 6.0) - 15.0) * t) + 10.0) * t * t * t; }

// perlin.clj:111-119 perlin-noise-sample (corner-vectors :60-66, corner-gradients :69-76, influence-values :79-82,
// interpolation-weights :89-99)
__device__ double perlin_sample(const double *__restrict__ gradients, int divisions, int size, double cx, double cy, double cz) {
  const double scale = (double)divisions / (double)size;
  const double point[3] = {cx * scale, cy * scale, cz * scale};
  double division[3], b[3], a[3];
  int c[3], cp[3];
  for (int q = 0; q < 3; q++) {
    division[q] = floor(point[q]);
    c[q] = (int)division[q];
    cp[q] = (c[q] + 1) % divisions;
    if (cp[q] < 0) cp[q] += divisions;
    b[q] = point[q] - division[q];
    a[q] = 1.0 - b[q];
  }
  double sum = 0.0;
  for (int z = 0; z < 2; z++)
    for (int y = 0; y < 2; y++)
      for (int x = 0; x < 2; x++) {
        const double corner[3] = {point[0] - (division[0] + x), point[1] - (division[1] + y), point[2] - (division[2] + z)};
        const double *g = gradients + ((size_t)((z ? cp[2] : c[2]) * divisions + (y ? cp[1] : c[1])) * divisions +
                                       (x ? cp[0] : c[0])) * 3;
        const double influence = g[0] * corner[0] + g[1] * corner[1] + g[2] * corner[2];
        const double weight = ease_curve(z ? b[2] : a[2]) * ease_curve(y ? b[1] : a[1]) * ease_curve(x ? b[0] : a[0]);
        const double term = weight * influence;
        sum = (x | y | z) ? sum + term : term;
      }
  return sum;
}

// perlin.clj:130-137: sample (k, j, i) at the cell (i + 1/2, j + 1/2, k + 1/2)
__global__ void k_perlin_samples(const double *__restrict__ gradients, int divisions, int size, double *raw,
                                 unsigned long long *extrema) {
  const long long n = (long long)size * size * size;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double v = 0.0;
  if (t < n) {
    const int i = (int)(t % size), j = (int)((t / size) % size), k = (int)(t / ((long long)size * size));
    v = perlin_sample(gradients, divisions, size, (double)i + 0.5, (double)j + 0.5, (double)k + 0.5);
    raw[t] = v;
  }
  block_min_max(v, t < n, extrema);
}

// normalize-vector (perlin.clj:102-108)
__global__ void k_perlin_finish(const double *__restrict__ raw, long long n, const unsigned long long *extrema, float *out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const double minimum = from_ordered_key(extrema[0]), maximum = from_ordered_key(extrema[1]);
  out[t] = (float)((raw[t] - minimum) / (maximum - minimum));
}

// ------------------------------------------------------------------ blue noise (bluenoise.clj), void-and-cluster
//
// 4096 insertions at the shipped size 64 x 64, each an arg-max / arg-min over the whole density array followed by an
// update of the whole array: parallel inside a step (one element per thread and trip), strictly serial across steps.
// One persistent CTA of 1024 threads runs all phases; the density function arrives as a table over the wrapped
// offsets (host libm exp), so that every decision can be checked against a CPU restatement: everything here is additions,
// subtractions and comparisons of doubles in the reference's order.

struct BlueNoise {
  int m, n;
  const double *ftab;        // [(dy + m/2) * m + (dx + m/2)]
  double *density;           // [m * m]
  unsigned char *mask;       // seed pattern, then the half-filled mask
  unsigned char *work;       // working copy (phase 1) / negated mask (phase 3)
  int *dither;
};

// bluenoise.clj:73-78 wrap
__device__ __forceinline__ int blue_wrap(int x, int m) {
  const int offset = m / 2;
  int r = (x + offset) % m;
  if (r < 0) r += m;
  return r - offset;
}

// bluenoise.clj:94-98 density-array: every element the (y, x)-ordered sum over the set mask entries (:81-91)
__device__ void blue_density_array(const BlueNoise &b, const unsigned char *mask) {
  const int m = b.m, count = m * m, offset = m / 2;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const int cy = i / m, cx = i % m;
    double sum = 0.0;
    for (int y = 0; y < m; y++) {
      const int wy = (blue_wrap(y - cy, m) + offset) * m;
      for (int x = 0; x < m; x++)
        sum = sum + (mask[y * m + x] ? b.ftab[wy + blue_wrap(x - cx, m) + offset] : 0.0);
    }
    b.density[i] = sum;
  }
  __syncthreads();
}

// bluenoise.clj:101-111 density-change
__device__ void blue_density_change(const BlueNoise &b, int sign, int index) {
  const int m = b.m, count = m * m, offset = m / 2;
  const int cy = index / m, cx = index % m;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const int y = i / m, x = i % m;
    const double f = b.ftab[(blue_wrap(y - cy, m) + offset) * m + blue_wrap(x - cx, m) + offset];
    b.density[i] = sign > 0 ? b.density[i] + f : b.density[i] - f;
  }
  __syncthreads();
}

// argmax-with-mask (want_max, entries whose mask is true) / argmin-with-mask (entries whose mask is false),
// bluenoise.clj:59-70.  max-key / min-key keep the LAST of equal extrema: ties go to the larger index.
__device__ int blue_arg_extreme(const BlueNoise &b, const unsigned char *mask, bool want_max) {
  __shared__ double s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_result;
  const int count = b.m * b.m;
  double best = 0.0;
  int best_i = -1;
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    if ((mask[i] != 0) != want_max) continue;
    const double v = b.density[i];
    if (best_i < 0 || (want_max ? v >= best : v <= best)) {
      best = v;
      best_i = i;
    }
  }
  auto better = [&](double v, int i, double w, int j) {   // is (v, i) preferred over (w, j)?
    if (i < 0) return false;
    if (j < 0) return true;
    if (v != w) return want_max ? v > w : v < w;
    return i > j;
  };
  for (int o = 16; o > 0; o >>= 1) {
    const double v = __shfl_xor_sync(0xffffffffu, best, o);
    const int i = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (better(v, i, best, best_i)) {
      best = v;
      best_i = i;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_val[warp] = best;
    s_idx[warp] = best_i;
  }
  __syncthreads();
  if (warp == 0) {
    const int nwarps = blockDim.x >> 5;
    best = lane < nwarps ? s_val[lane] : 0.0;
    best_i = lane < nwarps ? s_idx[lane] : -1;
    for (int o = 16; o > 0; o >>= 1) {
      const double v = __shfl_xor_sync(0xffffffffu, best, o);
      const int i = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (better(v, i, best, best_i)) {
        best = v;
        best_i = i;
      }
    }
    if (lane == 0) s_result = best_i;
  }
  __syncthreads();
  const int result = s_result;
  __syncthreads();          // s_result is reused by the next call
  return result;
}

// blue-noise, bluenoise.clj:175-185; the mask arrives with the seed picks scattered (scatter-mask :46-49)
__global__ void __launch_bounds__(1024) k_blue_noise(BlueNoise b) {
  const int count = b.m * b.m;
  // ---- seed-pattern (:114-126)
  blue_density_array(b, b.mask);
  for (;;) {
    const int cluster = blue_arg_extreme(b, b.mask, true);
    if (threadIdx.x == 0) b.mask[cluster] = 0;
    __syncthreads();
    blue_density_change(b, -1, cluster);
    const int hole = blue_arg_extreme(b, b.mask, false);
    if (threadIdx.x == 0) b.mask[hole] = 1;
    __syncthreads();
    if (cluster == hole) break;
    blue_density_change(b, +1, hole);
  }
  // ---- phase 1 (:129-143): remove the seed samples one by one, densest first
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    b.work[i] = b.mask[i];
    b.dither[i] = 0;
  }
  __syncthreads();
  blue_density_array(b, b.work);
  for (int n = b.n; n > 0;) {
    const int cluster = blue_arg_extreme(b, b.work, true);
    blue_density_change(b, -1, cluster);
    n--;
    if (threadIdx.x == 0) {
      b.work[cluster] = 0;
      b.dither[cluster] = n;
    }
    __syncthreads();
  }
  // ---- phase 2 (:146-156): fill the largest voids until half of the mask is set
  blue_density_array(b, b.mask);
  for (int n = b.n; n < count / 2; n++) {
    const int hole = blue_arg_extreme(b, b.mask, false);
    blue_density_change(b, +1, hole);
    if (threadIdx.x == 0) {
      b.mask[hole] = 1;
      b.dither[hole] = n;
    }
    __syncthreads();
  }
  // ---- phase 3 (:159-172): negate the mask and remove its clusters
  for (int i = threadIdx.x; i < count; i += blockDim.x) b.work[i] = !b.mask[i];
  __syncthreads();
  blue_density_array(b, b.work);
  for (int n = count / 2; n < count; n++) {
    const int cluster = blue_arg_extreme(b, b.work, true);
    blue_density_change(b, -1, cluster);
    if (threadIdx.x == 0) {
      b.work[cluster] = 0;
      b.dither[cluster] = n;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ host

namespace {

struct NoiseBuffers {
  double *grid = nullptr, *raw = nullptr;
  unsigned long long *extrema = nullptr;
  float *out = nullptr;
  ~NoiseBuffers() {
    cudaFree(grid);
    cudaFree(raw);
    cudaFree(extrema);
    cudaFree(out);
  }
};

// which: 0 = Worley, 1 = Perlin; exactly one of out_f / out_raw is wanted
int run_noise(int which, const double *grid, int divisions, int size, float *out_f, double *out_raw) {
  if (ensure_init()) return 1;
  if (!grid || (!out_f && !out_raw)) return fail("grid and out must not be NULL");
  if (divisions < 1 || size < 1) return fail("divisions and size must be positive");
  if (which == 0 && size % divisions != 0) return fail("size must be a multiple of divisions (worley.clj:69-72 takes size / divisions as the cell size)");
  if ((long long)size * size * size > (1LL << 31)) return fail("size is limited to 1290");
  const long long n = (long long)size * size * size;
  const size_t grid_doubles = (size_t)divisions * divisions * divisions * 3;
  cudaStream_t st = stream();
  NoiseBuffers b;
  CUDA_TRY(cudaMalloc((void **)&b.grid, grid_doubles * sizeof(double)));
  CUDA_TRY(cudaMalloc((void **)&b.raw, (size_t)n * sizeof(double)));
  CUDA_TRY(cudaMalloc((void **)&b.extrema, 2 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc((void **)&b.out, (size_t)n * sizeof(float)));
  CUDA_TRY(cudaMemcpyAsync(b.grid, grid, grid_doubles * sizeof(double), cudaMemcpyHostToDevice, st));
  const unsigned long long init[2] = {~0ull, 0ull};
  CUDA_TRY(cudaMemcpyAsync(b.extrema, init, sizeof init, cudaMemcpyHostToDevice, st));
  const int threads = 128;
  const int blocks = (int)((n + threads - 1) / threads);
  if (which == 0)
    k_worley_distances<<<blocks, threads, 0, st>>>(b.grid, divisions, size, b.raw, b.extrema);
  else
    k_perlin_samples<<<blocks, threads, 0, st>>>(b.grid, divisions, size, b.raw, b.extrema);
  CUDA_TRY(cudaGetLastError());
  if (out_raw) {
    CUDA_TRY(cudaMemcpyAsync(out_raw, b.raw, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
  } else {
    if (which == 0)
      k_worley_finish<<<blocks, threads, 0, st>>>(b.raw, n, b.extrema, b.out);
    else
      k_perlin_finish<<<blocks, threads, 0, st>>>(b.raw, n, b.extrema, b.out);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_f, b.out, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace
}  // namespace atm

extern "C" int sfsim_worley_noise(const double *grid, int divisions, int size, float *out) {
  return atm::run_noise(0, grid, divisions, size, out, nullptr);
}

extern "C" int sfsim_perlin_noise(const double *gradients, int divisions, int size, float *out) {
  return atm::run_noise(1, gradients, divisions, size, out, nullptr);
}

extern "C" int sfsim_worley_distances(const double *grid, int divisions, int size, double *out) {
  return atm::run_noise(0, grid, divisions, size, nullptr, out);
}

extern "C" int sfsim_perlin_samples(const double *gradients, int divisions, int size, double *out) {
  return atm::run_noise(1, gradients, divisions, size, nullptr, out);
}

namespace atm {
namespace {

struct BlueBuffers {
  double *ftab = nullptr, *density = nullptr;
  unsigned char *mask = nullptr, *work = nullptr;
  int *dither = nullptr;
  ~BlueBuffers() {
    cudaFree(ftab);
    cudaFree(density);
    cudaFree(mask);
    cudaFree(work);
    cudaFree(dither);
  }
};

int run_blue_noise(const int *picks, int n, int m, const double *ftab, int *dither) {
  if (ensure_init()) return 1;
  if (!picks || !ftab || !dither) return fail("picks, ftab and dither must not be NULL");
  if (m < 2 || m > 1024) return fail("the dither array size must be in [2, 1024]");
  const int count = m * m;
  if (n < 1 || n > count / 2) return fail("the number of seed samples must be in [1, size^2 / 2]");
  std::vector<unsigned char> mask((size_t)count, 0);
  for (int i = 0; i < n; i++) {
    if (picks[i] < 0 || picks[i] >= count) return fail("seed index out of range");
    if (mask[picks[i]]) return fail("seed indices must be distinct (pick-n draws without replacement)");
    mask[picks[i]] = 1;                                  // scatter-mask, bluenoise.clj:46-49
  }
  cudaStream_t st = stream();
  BlueBuffers b;
  CUDA_TRY(cudaMalloc((void **)&b.ftab, (size_t)count * sizeof(double)));
  CUDA_TRY(cudaMalloc((void **)&b.density, (size_t)count * sizeof(double)));
  CUDA_TRY(cudaMalloc((void **)&b.mask, (size_t)count));
  CUDA_TRY(cudaMalloc((void **)&b.work, (size_t)count));
  CUDA_TRY(cudaMalloc((void **)&b.dither, (size_t)count * sizeof(int)));
  CUDA_TRY(cudaMemcpyAsync(b.ftab, ftab, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b.mask, mask.data(), (size_t)count, cudaMemcpyHostToDevice, st));
  BlueNoise args = {m, n, b.ftab, b.density, b.mask, b.work, b.dither};
  k_blue_noise<<<1, 1024, 0, st>>>(args);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(dither, b.dither, (size_t)count * sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace
}  // namespace atm

extern "C" int sfsim_blue_noise(const int *picks, int n, int size, const double *ftab, int *dither) {
  return atm::run_blue_noise(picks, n, size, ftab, dither);
}

// the array `build.clj:45-51` writes to data/bluenoise.raw: dither / size / size as float32, density-function(sigma)
extern "C" int sfsim_blue_noise_texture(const int *picks, int n, int size, double sigma, float *out) {
  if (!out) return atm::fail("out is NULL");
  if (size < 2 || size > 1024) return atm::fail("the dither array size must be in [2, 1024]");
  if (!(sigma > 0)) return atm::fail("sigma must be positive");
  const int offset = size / 2;
  std::vector<double> ftab((size_t)size * size);
  for (int dy = -offset; dy < size - offset; dy++)
    for (int dx = -offset; dx < size - offset; dx++)     // density-function, bluenoise.clj:53-56
      ftab[(size_t)(dy + offset) * size + (dx + offset)] = exp(-((double)(dx * dx + dy * dy) / (2.0 * sigma * sigma)));
  std::vector<int> dither((size_t)size * size);
  if (atm::run_blue_noise(picks, n, size, ftab.data(), dither.data())) return 1;
  for (size_t i = 0; i < dither.size(); i++) out[i] = (float)(((double)dither[i] / (double)size) / (double)size);
  return 0;
}
