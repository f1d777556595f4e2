// 3-D Worley and Perlin noise textures (include/sfsim_noise.h): one thread per texture sample, double precision in
// the reference's operation order (the translation unit is compiled with -fmad=false), a device-wide min / max by
// atomics on order-preserving integer keys, and a second pass that normalises and casts to float32.
#include <cstring>
#include <string>

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
