// Device code shared by the table kernels (atm_tables.cu) and the lookup kernels (atm_lookup.cu): the view ray
// of a (height, elevation) texel pair in shared memory, peer stores, counters.
#pragma once

#include "atm_device.cuh"
#include "atm_tables.h"

namespace atm {

// pair index of the i-th CTA of a launch (see Shard)
__device__ __forceinline__ int shard_pair(const Shard &s, int i) {
  return s.run <= 1 ? s.begin + i * s.stride : s.begin + (i / s.run) * s.stride + (i % s.run);
}

// ------------------------------------------------------------------ view ray shared by one CTA

struct ViewRay {
  double r;        // |x|, x = (r, 0, 0)
  double vx, vy;   // view direction (atmosphere.clj:256-270)
  double dx, dy;   // ray end point minus x (atmosphere.clj:196-198)
  double dlen;     // |d|
  int above;
};

struct alignas(16) ViewSmem {
  ViewRay ray;
  double pkx[kMaxSteps], pky[kMaxSteps];  // outer sample points p_k
  double rk2[kMaxSteps];                  // |p_k|^2
  float cv0[kMaxSteps], cv1[kMaxSteps];   // column densities x -> p_k per component
  float dens0[kMaxSteps], dens1[kMaxSteps];  // exp(-h(p_k)/scale_c)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Fills ViewSmem for texel pair (h, e) of the 4-D space: the view ray (ray-scatter-backward,
// atmosphere.clj:401-412; ray end point :196-198), its `steps` outer sample points (ray.clj:19-30)
// and the transmittance integral x -> p_k for every k (atmosphere.clj:118-125, :199).
static __device__ void setup_view_ray(const Params &P, int h, int e, ViewSmem &vs, unsigned &esamples) {
  const int steps = P.shapes.ray_steps;
  if (threadIdx.x == 0) {
    V3 x = index_to_height(P.planet, P.shapes.s4[0], (double)h);
    V3 v;
    bool above;
    index_to_elevation(P.planet, P.shapes.s4[1], x.x, (double)e, v, above);
    V3 end = above ? atmosphere_intersection(P.planet, x, v) : surface_intersection(P.planet, x, v);
    V3 d = end - x;
    vs.ray.r = x.x;
    vs.ray.vx = v.x;
    vs.ray.vy = v.y;
    vs.ray.dx = d.x;
    vs.ray.dy = d.y;
    vs.ray.dlen = mag(d);
    vs.ray.above = above ? 1 : 0;
  }
  __syncthreads();
  const ViewRay ray = vs.ray;
  const double stepsize = 1.0 / (double)steps;
  for (int k = threadIdx.x; k < steps; k += blockDim.x) {
    double s = (0.5 + (double)k) * stepsize;
    double px = ray.r + ray.dx * s, py = ray.dy * s;
    double r2 = px * px + py * py;
    double hk = sqrt(r2) - P.planet.radius;
    vs.pkx[k] = px;
    vs.pky[k] = py;
    vs.rk2[k] = r2;
    vs.dens0[k] = (float)exp(-(hk / P.medium.scale[0]));
    vs.dens1[k] = (float)exp(-(hk / P.medium.scale[1]));
  }
  __syncthreads();
  // one thread per segment x -> p_k with the sequential sampler of the hot loops (two samples per packed instruction,
  // forward differences): a segment takes one thread ~4000 cycles, all `steps` of them run side by side
  for (int k = threadIdx.x; k < steps; k += blockDim.x) {
    double dkx = vs.pkx[k] - ray.r, dky = vs.pky[k];
    double dd = dkx * dkx + dky * dky;
    Quad q = make_quad(P.fast, ray.r * ray.r, ray.r * dkx, dd, steps);
    float s0, s1;
    density_sums_seq(P, q, steps, s0, s1);
    esamples += steps;
    double seg = stepsize * sqrt(dd);
    vs.cv0[k] = (float)((double)s0 * seg);
    vs.cv1[k] = (float)((double)s1 * seg);
  }
  __syncthreads();
}

// ------------------------------------------------------------------ view rays worked out once per build
//
// setup_view_ray is a latency-bound prologue (one thread's index maps, `steps` double-precision exponentials, one
// integral per outer sample) that three kernels used to repeat per CTA -- the first-order kernel once per CTA of a pair
// (eight times per pair on an 8-GPU shard), the ray-scatter records once more.  k_view_prepare runs it once per pair in
// small CTAs, all pairs in flight together, and leaves a compact image of ViewSmem in global memory; the consumers read
// it back with a few coalesced loads (measured on one B200: first order 4.28 -> 4.16 ms, 17 % fewer warp stall samples
// outside the hot loop).  Layout per pair: ViewRay (64 bytes reserved), then pkx, pky, rk2 (steps doubles each), then cv0,
// cv1, dens0, dens1 (steps floats each).
__host__ __device__ inline size_t view_pack_bytes(int steps) { return 64 + (size_t)steps * (3 * sizeof(double) + 4 * sizeof(float)); }

__device__ __forceinline__ void store_view_ray(const ViewSmem &vs, int steps, unsigned char *pack) {
  if (threadIdx.x == 0) *reinterpret_cast<ViewRay *>(pack) = vs.ray;
  double *d = reinterpret_cast<double *>(pack + 64);
  float *f = reinterpret_cast<float *>(pack + 64 + (size_t)steps * 3 * sizeof(double));
  for (int k = threadIdx.x; k < steps; k += blockDim.x) {
    d[k] = vs.pkx[k];
    d[steps + k] = vs.pky[k];
    d[2 * steps + k] = vs.rk2[k];
    f[k] = vs.cv0[k];
    f[steps + k] = vs.cv1[k];
    f[2 * steps + k] = vs.dens0[k];
    f[3 * steps + k] = vs.dens1[k];
  }
}

// ends with a CTA barrier, like setup_view_ray
__device__ __forceinline__ void load_view_ray(ViewSmem &vs, int steps, const unsigned char *__restrict__ pack) {
  if (threadIdx.x == 0) vs.ray = *reinterpret_cast<const ViewRay *>(pack);
  const double *d = reinterpret_cast<const double *>(pack + 64);
  const float *f = reinterpret_cast<const float *>(pack + 64 + (size_t)steps * 3 * sizeof(double));
  for (int k = threadIdx.x; k < steps; k += blockDim.x) {
    vs.pkx[k] = __ldg(d + k);
    vs.pky[k] = __ldg(d + steps + k);
    vs.rk2[k] = __ldg(d + 2 * steps + k);
    vs.cv0[k] = __ldg(f + k);
    vs.cv1[k] = __ldg(f + steps + k);
    vs.dens0[k] = __ldg(f + 2 * steps + k);
    vs.dens1[k] = __ldg(f + 3 * steps + k);
  }
  __syncthreads();
}

__device__ __forceinline__ void store_all(const PeerOut &o, size_t idx, float4 v) {
#pragma unroll 1
  for (int q = 0; q < o.n; q++) o.p[q][idx] = v;
}

__device__ __forceinline__ void count_esamples(unsigned long long *counter, unsigned n) {
  if (!counter) return;
  n = (unsigned)__reduce_add_sync(0xffffffffu, n);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(counter, (unsigned long long)n);
}


}  // namespace atm
