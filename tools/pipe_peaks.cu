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

// the first-order kernel's hot loop as shipped (atm_device.cuh density_sums_seq): per PAIR of samples 2 DADD, an
// integer repack, 2 FADD, 7 packed FFMA2 / FMUL2, 4 MUFU.EX2 and 2 FADD2 -- 19 issue slots for 4 exponentials, all lanes
// active, no per-item prologue: the ceiling of this instruction mix
__global__ void k_mix_packed(float *out, double u0, double step2_0, double step2_inc, float d1f0, float d1f_inc, float k0,
                             float k1, float b0, float b1) {
  float2 acc0a = make_float2(0.f, 0.f), acc1a = make_float2(0.f, 0.f), acc0b = make_float2(0.f, 0.f), acc1b = make_float2(0.f, 0.f);
  for (int rep = 0; rep < ITERS / 16; rep++) {
    double u = u0 + threadIdx.x * 1e-9 + rep * 1e-9, step2 = step2_0;
    float d1f = d1f0;
#pragma unroll 4
    for (int j = 0; j < 128; j += 4) {
#pragma unroll
      for (int half = 0; half < 2; half++) {
        int hi = __double2hiint(u), lo = __double2loint(u);
        const float ue = __int_as_float(__funnelshift_l(lo, hi - 0x38000000, 3));
        const float2 uu = make_float2(ue, ue + d1f);
        float2 q = __ffma2_rn(uu, make_float2(0.02734375f, 0.02734375f), make_float2(-0.0390625f, -0.0390625f));
        q = __ffma2_rn(q, uu, make_float2(0.0625f, 0.0625f));
        q = __ffma2_rn(q, uu, make_float2(-0.125f, -0.125f));
        q = __ffma2_rn(q, uu, make_float2(0.5f, 0.5f));
        const float2 hq = __fmul2_rn(uu, q);
        const float2 a0 = __ffma2_rn(hq, make_float2(k0, k0), make_float2(b0, b0));
        const float2 a1 = __ffma2_rn(hq, make_float2(k1, k1), make_float2(b1, b1));
        float2 e0, e1;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0.x) : "f"(a0.x));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0.y) : "f"(a0.y));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1.x) : "f"(a1.x));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1.y) : "f"(a1.y));
        if (half == 0) {
          acc0a = __fadd2_rn(acc0a, e0);
          acc1a = __fadd2_rn(acc1a, e1);
        } else {
          acc0b = __fadd2_rn(acc0b, e0);
          acc1b = __fadd2_rn(acc1b, e1);
        }
        u += step2;
        step2 += step2_inc;
        d1f += d1f_inc;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (acc0a.x + acc0a.y) + (acc1a.x + acc1a.y) + (acc0b.x + acc0b.y) + (acc1b.x + acc1b.y);
}

// Variants of that loop, to find out what it is bound by (per pair of samples):
//  1: one exponential + powers t^3, t^20 (2 MUFU, 12 packed)   2: u by packed FP32 Horner in the sample index (no DADD,
//  no repack)   3: variant 2 with a degree-2 height series (2 packed fewer)   4: variant 0 with the degree-2 series
template <int V>
__global__ void k_mix_variant(float *out, double u0, double step2_0, double step2_inc, float d1f0, float d1f_inc, float k0,
                              float k1, float b0, float b1, float A, float B, float C) {
  float2 acc0a = make_float2(0.f, 0.f), acc1a = make_float2(0.f, 0.f), acc0b = make_float2(0.f, 0.f), acc1b = make_float2(0.f, 0.f);
  for (int rep = 0; rep < ITERS / 16; rep++) {
    double u = u0 + threadIdx.x * 1e-9 + rep * 1e-9, step2 = step2_0;
    float d1f = d1f0;
    float2 m2 = make_float2(0.5f + rep * 1e-3f, 1.5f + rep * 1e-3f);
#pragma unroll 4
    for (int j = 0; j < 128; j += 4) {
#pragma unroll
      for (int half = 0; half < 2; half++) {
        float2 uu;
        if (V == 2 || V == 3) {
          uu = __ffma2_rn(__ffma2_rn(make_float2(C, C), m2, make_float2(B, B)), m2, make_float2(A, A));
          m2 = __fadd2_rn(m2, make_float2(2.f, 2.f));
        } else {
          int hi = __double2hiint(u), lo = __double2loint(u);
          const float ue = __int_as_float(__funnelshift_l(lo, hi - 0x38000000, 3));
          uu = make_float2(ue, ue + d1f);
          u += step2;
          step2 += step2_inc;
          d1f += d1f_inc;
        }
        float2 q;
        if (V == 3 || V == 4) {
          q = __ffma2_rn(uu, make_float2(0.0625f, 0.0625f), make_float2(-0.125f, -0.125f));
          q = __ffma2_rn(q, uu, make_float2(0.5f, 0.5f));
        } else {
          q = __ffma2_rn(uu, make_float2(0.02734375f, 0.02734375f), make_float2(-0.0390625f, -0.0390625f));
          q = __ffma2_rn(q, uu, make_float2(0.0625f, 0.0625f));
          q = __ffma2_rn(q, uu, make_float2(-0.125f, -0.125f));
          q = __ffma2_rn(q, uu, make_float2(0.5f, 0.5f));
        }
        const float2 hq = __fmul2_rn(uu, q);
        float2 &acc0 = half ? acc0b : acc0a, &acc1 = half ? acc1b : acc1a;
        if (V == 1) {
          const float2 a = __ffma2_rn(hq, make_float2(k0, k0), make_float2(b0, b0));
          float2 t;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(a.x));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(a.y));
          const float2 p2 = __fmul2_rn(t, t), p4 = __fmul2_rn(p2, p2), p8 = __fmul2_rn(p4, p4), p16 = __fmul2_rn(p8, p8);
          acc0 = __ffma2_rn(p4, p16, acc0);
          acc1 = __ffma2_rn(t, p2, acc1);
        } else {
          const float2 a0 = __ffma2_rn(hq, make_float2(k0, k0), make_float2(b0, b0));
          const float2 a1 = __ffma2_rn(hq, make_float2(k1, k1), make_float2(b1, b1));
          float2 e0, e1;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0.x) : "f"(a0.x));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0.y) : "f"(a0.y));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1.x) : "f"(a1.x));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1.y) : "f"(a1.y));
          acc0 = __fadd2_rn(acc0, e0);
          acc1 = __fadd2_rn(acc1, e1);
        }
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (acc0a.x + acc0a.y) + (acc1a.x + acc1a.y) + (acc0b.x + acc0b.y) + (acc1b.x + acc1b.y);
}

// second round: 5: one exponential + powers with the degree-2 series (10 packed, 2 MUFU)  6: pairs alternate between
// the two-exponential and the one-exponential form (degree 2)  7: variant 5 with the powers in scalar FP32
// 8: variant 5 all scalar  9: two exponentials, degree 2, all scalar
template <int V>
__global__ void k_mix_variant2(float *out, double u0, double step2_0, double step2_inc, float d1f0, float d1f_inc, float k0,
                               float k1, float b0, float b1, float kt, float bt) {
  float2 acc0a = make_float2(0.f, 0.f), acc1a = make_float2(0.f, 0.f), acc0b = make_float2(0.f, 0.f), acc1b = make_float2(0.f, 0.f);
  for (int rep = 0; rep < ITERS / 16; rep++) {
    double u = u0 + threadIdx.x * 1e-9 + rep * 1e-9, step2 = step2_0;
    float d1f = d1f0;
#pragma unroll 4
    for (int j = 0; j < 128; j += 4) {
#pragma unroll
      for (int half = 0; half < 2; half++) {
        int hi = __double2hiint(u), lo = __double2loint(u);
        const float ue = __int_as_float(__funnelshift_l(lo, hi - 0x38000000, 3));
        const float2 uu = make_float2(ue, ue + d1f);
        u += step2;
        step2 += step2_inc;
        d1f += d1f_inc;
        float2 &acc0 = half ? acc0b : acc0a, &acc1 = half ? acc1b : acc1a;
        float2 hq;
        if (V == 8 || V == 9) {
          float qx = fmaf(fmaf(uu.x, 0.0625f, -0.125f), uu.x, 0.5f), qy = fmaf(fmaf(uu.y, 0.0625f, -0.125f), uu.y, 0.5f);
          hq = make_float2(uu.x * qx, uu.y * qy);
        } else {
          float2 q = __ffma2_rn(uu, make_float2(0.0625f, 0.0625f), make_float2(-0.125f, -0.125f));
          q = __ffma2_rn(q, uu, make_float2(0.5f, 0.5f));
          hq = __fmul2_rn(uu, q);
        }
        const bool one_exp = V == 5 || V == 7 || V == 8 || (V == 6 && half == 1);
        if (one_exp) {
          float2 t;
          if (V == 8) {
            const float ax = fmaf(hq.x, kt, bt), ay = fmaf(hq.y, kt, bt);
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(ax));
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(ay));
          } else {
            const float2 a = __ffma2_rn(hq, make_float2(kt, kt), make_float2(bt, bt));
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(a.x));
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(a.y));
          }
          if (V == 7 || V == 8) {
            const float p2x = t.x * t.x, p2y = t.y * t.y, p4x = p2x * p2x, p4y = p2y * p2y, p8x = p4x * p4x, p8y = p4y * p4y;
            const float p16x = p8x * p8x, p16y = p8y * p8y;
            acc0.x = fmaf(p4x, p16x, acc0.x);
            acc0.y = fmaf(p4y, p16y, acc0.y);
            acc1.x = fmaf(t.x, p2x, acc1.x);
            acc1.y = fmaf(t.y, p2y, acc1.y);
          } else {
            const float2 p2 = __fmul2_rn(t, t), p4 = __fmul2_rn(p2, p2), p8 = __fmul2_rn(p4, p4), p16 = __fmul2_rn(p8, p8);
            acc0 = __ffma2_rn(p4, p16, acc0);
            acc1 = __ffma2_rn(t, p2, acc1);
          }
        } else {
          float2 a0, a1;
          if (V == 9) {
            a0 = make_float2(fmaf(hq.x, k0, b0), fmaf(hq.y, k0, b0));
            a1 = make_float2(fmaf(hq.x, k1, b1), fmaf(hq.y, k1, b1));
          } else {
            a0 = __ffma2_rn(hq, make_float2(k0, k0), make_float2(b0, b0));
            a1 = __ffma2_rn(hq, make_float2(k1, k1), make_float2(b1, b1));
          }
          float2 e0, e1;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0.x) : "f"(a0.x));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0.y) : "f"(a0.y));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1.x) : "f"(a1.x));
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1.y) : "f"(a1.y));
          if (V == 9) {
            acc0.x += e0.x;
            acc0.y += e0.y;
            acc1.x += e1.x;
            acc1.y += e1.y;
          } else {
            acc0 = __fadd2_rn(acc0, e0);
            acc1 = __fadd2_rn(acc1, e1);
          }
        }
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (acc0a.x + acc0a.y) + (acc1a.x + acc1a.y) + (acc0b.x + acc0b.y) + (acc1b.x + acc1b.y);
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
  // 4 CTAs of 256 threads per SM like the shipped kernel (64 registers); ITERS / 16 * 128 samples per thread
  double t_mixp = time_ms([&] { k_mix_packed<<<sms * 4 * 4, threads>>>((float *)buf, 1e-3, 2e-6, 1e-9, 1e-6f, 1e-9f, -300.f, -2000.f, 0.1f, 0.7f); });
  double n_mixp = (double)sms * 16 * threads * (ITERS / 16) * 128;
  // the same with 48 KB of (unused) dynamic shared memory per CTA: 4 resident CTAs = 32 warps per SM, the shipped occupancy
  double t_mixp4 = time_ms([&] { k_mix_packed<<<sms * 4 * 4, threads, 48 * 1024>>>((float *)buf, 1e-3, 2e-6, 1e-9, 1e-6f, 1e-9f, -300.f, -2000.f, 0.1f, 0.7f); });
#define MIXV(V) time_ms([&] { k_mix_variant<V><<<sms * 4 * 4, threads, 48 * 1024>>>((float *)buf, 1e-3, 2e-6, 1e-9, 1e-6f, 1e-9f, -300.f, -2000.f, 0.1f, 0.7f, 1e-3f, 1e-6f, 1e-9f); })
  double t_v1 = MIXV(1), t_v2 = MIXV(2), t_v3 = MIXV(3), t_v4 = MIXV(4);
#define MIXW(V) time_ms([&] { k_mix_variant2<V><<<sms * 4 * 4, threads, 48 * 1024>>>((float *)buf, 1e-3, 2e-6, 1e-9, 1e-6f, 1e-9f, -300.f, -2000.f, 0.1f, 0.7f, -100.f, 0.03f); })
  double t_v5 = MIXW(5), t_v6 = MIXW(6), t_v7 = MIXW(7), t_v8 = MIXW(8), t_v9 = MIXW(9);
  CHECK(cudaGetLastError());
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, "
         "\"ffma_per_s\": %.4e, \"dfma_per_s\": %.4e, \"mufu_ex2_per_s\": %.4e, \"mufu_rsqrt_per_s\": %.4e, "
         "\"cvt_f64_f32_loop_per_s\": %.4e, \"mixed_esample_per_s\": %.4e, \"packed_hot_loop_exp_per_s\": %.4e, \"packed_hot_loop_exp_per_s_32_warps\": %.4e, "
         "\"ffma_per_clk_sm\": %.2f, \"dfma_per_clk_sm\": %.2f, \"ex2_per_clk_sm\": %.2f}\n",
         p.name, sms, clk, n / (t_ffma * 1e-3), n / (t_dfma * 1e-3), n / (t_ex2 * 1e-3), n / (t_rsq * 1e-3),
         n / (t_cvt * 1e-3), n / (t_mix * 1e-3), 2 * n_mixp / (t_mixp * 1e-3), 2 * n_mixp / (t_mixp4 * 1e-3), n / (t_ffma * 1e-3) / (clk * 1e3) / sms,
         n / (t_dfma * 1e-3) / (clk * 1e3) / sms, n / (t_ex2 * 1e-3) / (clk * 1e3) / sms);
  // clocks per pair of samples per SM sub-partition (4 per SM), 32 resident warps per SM
  double pairs = n_mixp / 2.0 / 32.0;   // warp-level pair iterations
  double per_smsp = pairs / (sms * 4.0);
  printf("{\"hot_loop_clk_per_pair\": {\"shipped\": %.2f, \"one_exponential_and_powers\": %.2f, \"fp32_horner_u\": %.2f, "
         "\"fp32_horner_u_degree2\": %.2f, \"degree2\": %.2f}}\n",
         t_mixp4 * 1e-3 * clk * 1e3 / per_smsp, t_v1 * 1e-3 * clk * 1e3 / per_smsp, t_v2 * 1e-3 * clk * 1e3 / per_smsp,
         t_v3 * 1e-3 * clk * 1e3 / per_smsp, t_v4 * 1e-3 * clk * 1e3 / per_smsp);
  printf("{\"hot_loop_clk_per_pair_round2\": {\"one_exp_degree2\": %.2f, \"alternating\": %.2f, \"one_exp_scalar_powers\": %.2f, "
         "\"one_exp_all_scalar\": %.2f, \"two_exp_all_scalar\": %.2f}}\n",
         t_v5 * 1e-3 * clk * 1e3 / per_smsp, t_v6 * 1e-3 * clk * 1e3 / per_smsp, t_v7 * 1e-3 * clk * 1e3 / per_smsp,
         t_v8 * 1e-3 * clk * 1e3 / per_smsp, t_v9 * 1e-3 * clk * 1e3 / per_smsp);
  return 0;
}
