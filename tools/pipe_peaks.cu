// Micro-benchmark: measured per-pipe instruction peaks of the GPU this runs on.
// These are the roofline denominators for the atmosphere-LUT kernels, which are bound by the
// SFU (MUFU.EX2) and FP32/FP64 pipes rather than by HBM.  Prints one JSON object.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_peaks tools/pipe_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CHECK(x)                                                                         \
  do {                                                                                   \
    cudaError_t e = (x);                                                                 \
    if (e != cudaSuccess) {                                                              \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));          \
      return 1;                                                                          \
    }                                                                                    \
  } while (0)

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

__global__ void k_ffma(float *out, float a, float b) {
  float x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x * 1e-3f + c;
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = fmaf(x[c], a, b);
  }
  float s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dfma(double *out, double a, double b) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) x[c] = fma(x[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ex2(float *out, float a) {
  float x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) x[c] = -(threadIdx.x * 1e-3f + c);
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
      float y;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[c]));
      x[c] = y;
    }
  }
  float s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s * a;
}

__global__ void k_rsqrt(float *out, float a) {
  float x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) x[c] = threadIdx.x * 1e-3f + c + 1.0f;
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
      float y;
      asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x[c]));
      x[c] = y;
    }
  }
  float s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s * a;
}

// f64 -> f32 conversion (F2F.F32.F64); the f32 -> f64 direction is folded into a DADD chain
__global__ void k_cvt(float *out, double a) {
  double x[CHAINS];
  float acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) {
    x[c] = threadIdx.x * 1e-3 + c;
    acc[c] = 0;
  }
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
      float y;
      asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(y) : "d"(x[c]));
      acc[c] += y;          // FADD (fp32 pipe)
      x[c] = x[c] + a;      // DADD (fp64 pipe)
    }
  }
  float s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) s += acc[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the mixed inner loop the LUT kernels run: 2 DFMA + int repack + 7 FFMA + 2 EX2 + 2 FADD
__global__ void k_mix(float *out, double c0, double c1, double c2, float k1, float k2) {
  float s1 = 0, s2 = 0;
  double m = threadIdx.x * 1e-6;
  for (int i = 0; i < ITERS; i++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
      double num = fma(fma(c2, m, c1), m, c0);
      m += 1.0;
      int hi = __double2hiint(num), lo = __double2loint(num);
      float w = __int_as_float(__funnelshift_l(lo, hi - 0x38000000, 3));
      float q = fmaf(fmaf(fmaf(w, -0.0390625f, 0.0625f), w, -0.125f), w, 0.5f) * w;
      float e1, e2;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q * k1));
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(q * k2));
      s1 += e1;
      s2 += e2;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s1 + s2;
}

template <typename F>
static double time_ms(F launch) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int i = 0; i < 3; i++) launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p;
  CHECK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  int blocks = sms * 8, threads = 256;
  void *buf;
  CHECK(cudaMalloc(&buf, (size_t)blocks * threads * 8));
  double n = (double)blocks * threads * ITERS * CHAINS;
  double t_ffma = time_ms([&] { k_ffma<<<blocks, threads>>>((float *)buf, 1.0001f, 0.5f); });
  double t_dfma = time_ms([&] { k_dfma<<<blocks, threads>>>((double *)buf, 1.0001, 0.5); });
  double t_ex2 = time_ms([&] { k_ex2<<<blocks, threads>>>((float *)buf, 1.0f); });
  double t_rsq = time_ms([&] { k_rsqrt<<<blocks, threads>>>((float *)buf, 1.0f); });
  double t_cvt = time_ms([&] { k_cvt<<<blocks, threads>>>((float *)buf, 1.5); });
  double t_mix = time_ms([&] { k_mix<<<blocks, threads>>>((float *)buf, 1e-3, 1e-7, 1e-12, -3.0f, -20.0f); });
  CHECK(cudaGetLastError());
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, "
         "\"ffma_per_s\": %.4e, \"dfma_per_s\": %.4e, \"mufu_ex2_per_s\": %.4e, \"mufu_rsqrt_per_s\": %.4e, "
         "\"cvt_f64_f32_loop_per_s\": %.4e, \"mixed_esample_per_s\": %.4e, "
         "\"ffma_per_clk_sm\": %.2f, \"dfma_per_clk_sm\": %.2f, \"ex2_per_clk_sm\": %.2f}\n",
         p.name, sms, clk, n / (t_ffma * 1e-3), n / (t_dfma * 1e-3), n / (t_ex2 * 1e-3), n / (t_rsq * 1e-3),
         n / (t_cvt * 1e-3), n / (t_mix * 1e-3), n / (t_ffma * 1e-3) / (clk * 1e3) / sms,
         n / (t_dfma * 1e-3) / (clk * 1e3) / sms, n / (t_ex2 * 1e-3) / (clk * 1e3) / sms);
  return 0;
}
