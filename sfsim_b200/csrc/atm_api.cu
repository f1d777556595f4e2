// C ABI of libsfsim_atmosphere.so (include/sfsim_atmosphere.h): parameter validation, device
// buffers, and the host orchestration of generate-atmosphere-luts (atmosphere_lut.clj:43-105).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "sfsim_atmosphere.h"
#include "atm_api_internal.h"
#include "atm_tables.h"

namespace atm {

// ------------------------------------------------------------------ global state (jolt_init-style singletons)

static thread_local std::string g_error;
static int g_device = -1;
static cudaStream_t g_stream = nullptr;

int fail(const std::string &msg) {
  g_error = msg;
  return 1;
}

int fail_cuda(cudaError_t e, const char *what) {
  g_error = std::string(what) + ": " + cudaGetErrorString(e);
  return 1;
}

int ensure_init() {
  if (g_device >= 0) return 0;
  return atmlut_init(0);
}

cudaStream_t stream() { return g_stream; }

// exp(i/64), i = kExpTabLo .. kExpTabHi, for the kernels' exp_tab (host libm values), one copy per device
static double *g_exp_table[64] = {};

int exp_table(const double *&table) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= 64) return fail("cannot identify the current device");
  if (!g_exp_table[dev]) {
    std::vector<double> host(kExpTabSize);
    for (int i = 0; i < kExpTabSize; i++) host[i] = exp((double)(i + kExpTabLo) / 64.0);
    e = cudaMalloc((void **)&g_exp_table[dev], kExpTabSize * sizeof(double));
    if (e != cudaSuccess) return fail_cuda(e, "cudaMalloc(exp table)");
    e = cudaMemcpy(g_exp_table[dev], host.data(), kExpTabSize * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return fail_cuda(e, "cudaMemcpy(exp table)");
  }
  table = g_exp_table[dev];
  return 0;
}

// ------------------------------------------------------------------ parameters

int make_planet_medium(const atmlut_planet *planet, const atmlut_scatter *scatter, int n, Params &P) {
  if (!planet) return fail("planet is NULL");
  if (n < 0 || n > 2) return fail("scatter count must be 0, 1 or 2");
  if (n > 0 && !scatter) return fail("scatter is NULL");
  if (!(planet->radius > 0) || !(planet->height > 0)) return fail("planet radius and height must be positive");
  memset(&P, 0, sizeof(P));
  P.planet.radius = planet->radius;
  P.planet.height = planet->height;
  for (int i = 0; i < 3; i++) P.planet.brightness[i] = planet->brightness[i];
  P.medium.n = n;
  for (int c = 0; c < 2; c++) {
    P.medium.scale[c] = 1.0;
    P.medium.quotient[c] = 1.0;
  }
  for (int c = 0; c < n; c++) {
    if (!(scatter[c].scale > 0)) return fail("scatter scale must be positive");
    if (scatter[c].quotient == 0) return fail("scatter quotient must not be zero");
    for (int i = 0; i < 3; i++) P.medium.base[c][i] = scatter[c].base[i];
    P.medium.scale[c] = scatter[c].scale;
    P.medium.g[c] = scatter[c].g;
    P.medium.quotient[c] = scatter[c].quotient;
  }
  // fast sampler constants (atm_device.cuh)
  const double R = planet->radius, Rt = planet->radius + planet->height;
  const double delta = R / 4096.0;
  const double Rp = R - delta;
  P.fast.rp2 = Rp * Rp;
  P.fast.inv_rp2 = 1.0 / (Rp * Rp);
  P.fast.poly = ((Rt * Rt - Rp * Rp) / (Rp * Rp) <= 0.05) ? 1 : 0;
  const double log2e = 1.4426950408889634;
  for (int c = 0; c < 2; c++) {
    P.fast.k[c] = (float)(-Rp * log2e / P.medium.scale[c]);
    P.fast.b[c] = (float)(delta * log2e / P.medium.scale[c]);
    for (int i = 0; i < 3; i++) P.fast.ext[c][i] = (float)(P.medium.base[c][i] / P.medium.quotient[c]);
  }
  // Sampler variants (atm_device.cuh).  Degree of the height series: the u^3 term of sqrt(1+u) - 1 is 0.078 u^3 of the
  // height; it is dropped where it moves the largest exponent (top of the atmosphere, shortest scale height) by < 4e-6.
  // Scale heights in the ratio 3 : 20 (exactly, in double): one exponential per sample on every second pair.
  static const int knob_pow = getenv("ATMLUT_POW") ? atoi(getenv("ATMLUT_POW")) : 2;       // A/B: 0 off, 1 every pair, 2 alternate
  static const int knob_degree = getenv("ATMLUT_DEGREE") ? atoi(getenv("ATMLUT_DEGREE")) : 0;   // A/B: 0 automatic
  const double u_max = (Rt * Rt - Rp * Rp) / (Rp * Rp);
  const double min_scale = n == 2 ? std::min(P.medium.scale[0], P.medium.scale[1]) : P.medium.scale[0];
  const bool thin = 0.078 * u_max * u_max * u_max * ((Rt - R) / min_scale) < 4e-6;
  P.fast.degree = knob_degree == 2 || knob_degree == 4 ? knob_degree : (thin ? 2 : 4);
  P.fast.pow_mode = 0;
  P.fast.pow_alt = knob_pow == 2 ? 1 : 0;
  if (n == 2 && knob_pow != 0 && P.fast.degree == 2) {
    const double s0 = P.medium.scale[0], s1 = P.medium.scale[1];
    const double common = s0 * 20.0 == s1 * 3.0 ? s0 * 20.0 : (s0 * 3.0 == s1 * 20.0 ? s0 * 3.0 : 0.0);
    if (common > 0.0) {
      P.fast.pow_mode = s0 * 20.0 == s1 * 3.0 ? 1 : 2;
      P.fast.kt = (float)(-Rp * log2e / common);
      P.fast.bt = (float)(delta * log2e / common);
    }
  }
  for (int i = 0; i < 3; i++) P.intensity[i] = 1.0;
  return 0;
}

int make_params(const atmlut_planet *planet, const atmlut_scatter *scatter, int n, const atmlut_config *cfg,
                Params &P) {
  if (make_planet_medium(planet, scatter, n, P)) return 1;
  if (!cfg) return fail("config is NULL");
  if (planet->centre[0] != 0 || planet->centre[1] != 0 || planet->centre[2] != 0)
    return fail("table builders need the planet centre at the origin (the index maps assume it, atmosphere.clj:95-102)");
  const int sizes[8] = {cfg->height_size, cfg->elevation_size, cfg->light_elevation_size, cfg->heading_size,
                        cfg->transmittance_height_size, cfg->transmittance_elevation_size, cfg->surface_height_size,
                        cfg->surface_sun_elevation_size};
  for (int i = 0; i < 8; i++)
    if (sizes[i] < 2) return fail("every table axis needs at least 2 entries");
  if (cfg->ray_steps < 1 || cfg->ray_steps > kMaxSteps) return fail("ray_steps must be in [1, 256]");
  if (cfg->sphere_steps < 2) return fail("sphere_steps must be at least 2");
  if ((long long)cfg->height_size * cfg->elevation_size * cfg->light_elevation_size * cfg->heading_size > (1LL << 28))
    return fail("the 4-D table must not exceed 2^28 texels");
  if ((long long)cfg->light_elevation_size * cfg->heading_size > 6144)
    return fail("light_elevation_size * heading_size must not exceed 6144 (shared-memory tile of the ray-scatter kernel)");
  for (int i = 0; i < 4; i++) P.shapes.s4[i] = sizes[i];
  P.shapes.st[0] = sizes[4];
  P.shapes.st[1] = sizes[5];
  P.shapes.se[0] = sizes[6];
  P.shapes.se[1] = sizes[7];
  P.shapes.ray_steps = cfg->ray_steps;
  P.shapes.sphere_steps = cfg->sphere_steps;
  for (int i = 0; i < 3; i++) P.intensity[i] = cfg->intensity[i];
  return 0;
}

// Direction list of spherical-integral (sphere.clj:70-93) for normal (1, 0, 0), in evaluation order.
// Every table point is x = (r, 0, 0), so oriented-matrix (matrix.clj:207-213) has rows n = (1,0,0),
// o1 = orthogonal(n) = normalize(n x e1) = (0,0,1) (quaternion.clj:166-171), o2 = n x o1 = (0,-1,0).
// Ring counts ceil(sin(theta) * phi_steps) are evaluated here in host double precision.
void sphere_directions(int theta_steps, int phi_steps, double theta_range, std::vector<double> &dirs,
                       std::vector<double> &weights) {
  dirs.clear();
  weights.clear();
  const double delta2 = theta_range / (double)theta_steps / 2;
  const double m[9] = {1, 0, 0, 0, 0, 1, 0, -1, 0};
  for (int k = 0; k < theta_steps; k++) {
    double theta = theta_range * ((0.5 + (double)k) / (double)theta_steps);
    double factor = cos(theta - delta2) - cos(theta + delta2);
    int ringsteps = (int)ceil(sin(theta) * (double)phi_steps);
    double weight = (2 * kPi) / (double)ringsteps;
    for (int j = 0; j < ringsteps; j++) {
      double phi = 2 * kPi * ((0.5 + (double)j) / (double)ringsteps);
      double x = cos(theta), y = sin(theta) * cos(phi), z = sin(theta) * sin(phi);
      dirs.push_back(m[0] * x + m[3] * y + m[6] * z);
      dirs.push_back(m[1] * x + m[4] * y + m[7] * z);
      dirs.push_back(m[2] * x + m[5] * y + m[8] * z);
      weights.push_back(factor * weight);
    }
  }
}

// ------------------------------------------------------------------ builder

struct Stage {
  std::string name;
  cudaEvent_t begin, end;
};

// number of sharded device tables a rank exposes to its peers (CUDA IPC or peer access), in this order:
// R1, M1, dJ, dS, dS2, S, S_new (4-D), dE, dE_new (2-D), file_S, file_M (rank 0's are written by everybody), flags
static const int kPeerTables = 12;

struct Builder {
  Params P;
  int iterations = 0;
  int rank = 0, world = 1;
  int n_he = 0, he_per_rank = 0;
  Shard shard = {0, 1, 1};                 // the (height, elevation) pairs this rank integrates
  int he_count = 0;
  int h_first = 0, h_stride = 1, h_count = 0;   // height rows whose direction tiles this rank needs
  long long ntex = 0, n4 = 0, n4_pad = 0, nt = 0, ne = 0;
  atmlut_allgather_fn allgather = nullptr;
  void *allgather_user = nullptr;
  // device tables (float4)
  float4 *T = nullptr, *dE = nullptr, *dE_new = nullptr, *Eacc = nullptr, *Eacc_new = nullptr;
  float4 *R1 = nullptr, *M1 = nullptr, *dS = nullptr, *dS2 = nullptr, *dJ = nullptr, *S = nullptr, *S_new = nullptr;
  int device = 0;                          // CUDA device of this builder
  cudaStream_t main = nullptr;             // main stream of the build DAG (owned)
  cudaStream_t side = nullptr;             // second stream of the build DAG
  // peer-to-peer mode: every sharded table also mapped on the peers, flag words for the barrier
  bool p2p = false;
  PeerOut peer_R1 = {}, peer_M1 = {}, peer_dJ = {}, peer_dS = {}, peer_dS2 = {}, peer_S = {}, peer_S_new = {};
  PeerOut peer_dE = {}, peer_dE_new = {};
  float *file_S_root = nullptr, *file_M_root = nullptr;   // rank 0's file-layout tables (every rank fills its pairs)
  unsigned *flags = nullptr;
  unsigned *epochs = nullptr;              // barrier epochs of the two streams (device memory: graph replay)
  int *error_flag = nullptr;
  unsigned *peer_flags[kMaxPeers] = {};
  int barrier_timeout_ms = 10000;
  std::vector<void *> ipc_opened;
  std::vector<cudaEvent_t> events;         // ordering events between the two streams
  float *file_T = nullptr, *file_E = nullptr, *file_S = nullptr, *file_M = nullptr;
  double *sphere_dirs = nullptr, *sphere_w = nullptr, *half_dirs = nullptr, *half_w = nullptr;
  int n_sphere = 0, n_half = 0;
  DirInfo *dir_info = nullptr;
  float4 *tiles_a = nullptr, *tiles_b = nullptr;   // blended S tiles per (height, sphere direction)
  HalfDirInfo *half_info = nullptr;
  void *ray_samples = nullptr;             // per (pair, outer sample) records of the ray-scatter kernel
  void *view_packs = nullptr;              // the view ray of every pair (launch_view_prepare)
  unsigned long long *counter = nullptr;
  const double *exp_tab = nullptr;
  std::vector<Stage> stages;
  size_t stage_cursor = 0;
  bool timed = false;                      // this enqueue records the per-stage events
  bool ran = false;
  long long launches = 0;                  // kernels launched by one build
  // the whole two-stream DAG as one CUDA graph (captured on the first plain run)
  bool use_graph = true;
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};   // [1]: the same DAG with the per-stage event records
  bool capturing = false;
  // page-locked host destinations of the four file tables (the one-shot call on one GPU): the downloads are then
  // nodes of the build graph, and the 12.5 MB Mie-strength file, finished after the first iteration, leaves the
  // device while the remaining iterations run
  float *host_out[4] = {nullptr, nullptr, nullptr, nullptr};
  // pinned staging for downloads into pageable host memory
  unsigned char *staging[2] = {nullptr, nullptr};
  cudaEvent_t staging_done[2] = {nullptr, nullptr};

  ~Builder() {
    cudaSetDevice(device);
    if (main) cudaStreamSynchronize(main);
    for (auto g : graph_exec)
      if (g) cudaGraphExecDestroy(g);
    void *ptrs[] = {T, dE, dE_new, Eacc, Eacc_new, R1, M1, dS, dS2, dJ, S, S_new, file_T, file_E, file_S, file_M,
                    sphere_dirs, sphere_w, half_dirs, half_w, dir_info, half_info, counter,
                    tiles_a, tiles_b, flags, epochs, error_flag, ray_samples, view_packs};
    for (void *p : ptrs)
      if (p) cudaFree(p);
    for (auto &s : stages) {
      cudaEventDestroy(s.begin);
      cudaEventDestroy(s.end);
    }
    for (auto e : events) cudaEventDestroy(e);
    for (void *p : ipc_opened) cudaIpcCloseMemHandle(p);
    for (int i = 0; i < 2; i++) {
      if (staging[i]) cudaFreeHost(staging[i]);
      if (staging_done[i]) cudaEventDestroy(staging_done[i]);
    }
    if (side) cudaStreamDestroy(side);
    if (main) cudaStreamDestroy(main);
  }
};

template <typename T>
static int dev_alloc(T *&p, size_t count) {
  CUDA_TRY(cudaMalloc((void **)&p, count * sizeof(T)));
  return 0;
}

static int upload(double *&dst, const std::vector<double> &src, cudaStream_t stream) {
  if (dev_alloc(dst, src.size())) return 1;
  CUDA_TRY(cudaMemcpyAsync(dst, src.data(), src.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  return 0;
}

// Slab of rank `rank`: (height, elevation) pairs [begin, begin + count); every rank owns `per_rank` slots
// of the (padded) table so that one equal-sized all-gather reassembles it.
static void slab_of(int n_pairs, int rank, int world, int &begin, int &count, int &per_rank) {
  per_rank = (n_pairs + world - 1) / world;
  begin = rank * per_rank;
  count = std::max(0, std::min(n_pairs, begin + per_rank) - begin);
}

// NCCL (all-gather) mode and single GPU: a contiguous slab of pairs
static void use_slab_sharding(Builder &b) {
  const int E = b.P.shapes.s4[1];
  int begin = 0;
  slab_of(b.n_he, b.rank, b.world, begin, b.he_count, b.he_per_rank);
  b.shard = Shard{begin, 1, 1};
  b.h_stride = 1;
  b.h_first = b.he_count > 0 ? begin / E : 0;
  b.h_count = b.he_count > 0 ? (begin + b.he_count - 1) / E - b.h_first + 1 : 0;
}

// peer-to-peer mode: whole height rows rank, rank + world, ... (low and high altitudes on every rank; the direction
// tiles of a row are needed by this rank only)
static void use_row_sharding(Builder &b) {
  const int H = b.P.shapes.s4[0], E = b.P.shapes.s4[1];
  b.h_first = b.rank;
  b.h_stride = b.world;
  b.h_count = H > b.rank ? (H - b.rank + b.world - 1) / b.world : 0;
  b.shard = Shard{b.rank * E, b.world * E, E};
  b.he_count = b.h_count * E;
}

static int builder_alloc(Builder &b) {
  const Params &P = b.P;
  b.n_he = P.shapes.s4[0] * P.shapes.s4[1];
  b.ntex = (long long)P.shapes.s4[2] * P.shapes.s4[3];
  use_slab_sharding(b);
  b.n4 = b.n_he * b.ntex;
  b.n4_pad = (long long)b.he_per_rank * b.world * b.ntex;
  b.nt = (long long)P.shapes.st[0] * P.shapes.st[1];
  b.ne = (long long)P.shapes.se[0] * P.shapes.se[1];
  CUDA_TRY(cudaStreamCreateWithFlags(&b.main, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&b.side, cudaStreamNonBlocking));
  float4 **four[] = {&b.R1, &b.M1, &b.dS, &b.dS2, &b.dJ, &b.S, &b.S_new};
  for (auto p : four) {
    if (dev_alloc(*p, (size_t)b.n4_pad)) return 1;
    CUDA_TRY(cudaMemsetAsync(*p, 0, (size_t)b.n4_pad * sizeof(float4), b.main));
  }
  float4 **two[] = {&b.dE, &b.dE_new, &b.Eacc, &b.Eacc_new};
  for (auto p : two)
    if (dev_alloc(*p, (size_t)b.ne)) return 1;
  if (dev_alloc(b.T, (size_t)b.nt)) return 1;
  if (dev_alloc(b.file_T, (size_t)b.nt * 3)) return 1;
  if (dev_alloc(b.file_E, (size_t)b.ne * 3)) return 1;
  if (dev_alloc(b.file_S, (size_t)b.n4 * 3)) return 1;
  if (dev_alloc(b.file_M, (size_t)b.n4 * 3)) return 1;
  b.file_S_root = b.file_S;
  b.file_M_root = b.file_M;
  if (dev_alloc(b.counter, 2)) return 1;
  if (dev_alloc(b.flags, 2 * kMaxPeers) || dev_alloc(b.epochs, 2) || dev_alloc(b.error_flag, 1)) return 1;
  CUDA_TRY(cudaMemsetAsync(b.flags, 0, 2 * kMaxPeers * sizeof(unsigned), b.main));
  CUDA_TRY(cudaMemsetAsync(b.epochs, 0, 2 * sizeof(unsigned), b.main));
  CUDA_TRY(cudaMemsetAsync(b.error_flag, 0, sizeof(int), b.main));
  b.peer_R1 = local_out(b.R1);
  b.peer_M1 = local_out(b.M1);
  b.peer_dJ = local_out(b.dJ);
  b.peer_dS = local_out(b.dS);
  b.peer_dS2 = local_out(b.dS2);
  b.peer_S = local_out(b.S);
  b.peer_S_new = local_out(b.S_new);
  b.peer_dE = local_out(b.dE);
  b.peer_dE_new = local_out(b.dE_new);
  std::vector<double> dirs, w;
  sphere_directions(P.shapes.sphere_steps >> 1, P.shapes.sphere_steps, kPi, dirs, w);  // sphere.clj:102-105
  b.n_sphere = (int)w.size();
  if (b.n_sphere > kMaxDirs) return fail("sphere_steps too large for the point-scatter kernel");
  if (upload(b.sphere_dirs, dirs, b.main) || upload(b.sphere_w, w, b.main)) return 1;
  // surface-radiance is called with ray-steps as its sphere steps (atmosphere_lut.clj:89)
  sphere_directions(P.shapes.ray_steps >> 2, P.shapes.ray_steps, kPi / 2, dirs, w);  // sphere.clj:96-99
  b.n_half = (int)w.size();
  if (b.n_half > 0 && (upload(b.half_dirs, dirs, b.main) || upload(b.half_w, w, b.main))) return 1;
  if (dev_alloc(b.dir_info, (size_t)P.shapes.s4[0] * b.n_sphere)) return 1;
  if (dev_alloc(b.tiles_a, (size_t)P.shapes.s4[0] * b.n_sphere * b.ntex)) return 1;
  if (dev_alloc(b.tiles_b, (size_t)P.shapes.s4[0] * b.n_sphere * b.ntex)) return 1;
  if (dev_alloc(b.half_info, (size_t)P.shapes.se[0] * std::max(1, b.n_half))) return 1;
  if (exp_table(b.exp_tab)) return 1;
  CUDA_TRY(cudaStreamSynchronize(b.main));
  return 0;
}

// the peer tables of builder `b` in the order of kPeerTables
static void peer_table_list(Builder &b, void *ptrs[kPeerTables]) {
  void *mine[kPeerTables] = {b.R1, b.M1, b.dJ, b.dS, b.dS2, b.S, b.S_new, b.dE, b.dE_new, b.file_S, b.file_M, b.flags};
  memcpy(ptrs, mine, sizeof mine);
}

// registers the tables of peer `q` (already mapped into this process / device) with builder `b`
static void add_peer(Builder &b, int q, void *const ptrs[kPeerTables]) {
  PeerOut *outs[9] = {&b.peer_R1, &b.peer_M1, &b.peer_dJ, &b.peer_dS, &b.peer_dS2, &b.peer_S, &b.peer_S_new,
                      &b.peer_dE, &b.peer_dE_new};
  if (q != b.rank)
    for (int i = 0; i < 9; i++) outs[i]->p[outs[i]->n++] = (float4 *)ptrs[i];
  if (q == 0) {
    b.file_S_root = (float *)ptrs[9];
    b.file_M_root = (float *)ptrs[10];
  }
  b.peer_flags[q] = (unsigned *)ptrs[11];
}

static int stage_begin(Builder &b, const std::string &name) {
  if (!b.timed) return 0;
  if (b.stage_cursor == b.stages.size()) {
    Stage s;
    s.name = name;
    CUDA_TRY(cudaEventCreate(&s.begin));
    CUDA_TRY(cudaEventCreate(&s.end));
    b.stages.push_back(s);
  }
  // inside a capture the record becomes an event-record node of the graph (timing stays valid on replay)
  if (b.capturing)
    CUDA_TRY(cudaEventRecordWithFlags(b.stages[b.stage_cursor].begin, b.main, cudaEventRecordExternal));
  else
    CUDA_TRY(cudaEventRecord(b.stages[b.stage_cursor].begin, b.main));
  return 0;
}

static int stage_end(Builder &b) {
  if (!b.timed) return 0;
  if (b.capturing)
    CUDA_TRY(cudaEventRecordWithFlags(b.stages[b.stage_cursor].end, b.main, cudaEventRecordExternal));
  else
    CUDA_TRY(cudaEventRecord(b.stages[b.stage_cursor].end, b.main));
  b.stage_cursor++;
  return 0;
}

static int peer_barrier(Builder &b, int flag_set, cudaStream_t stream) {
  CUDA_TRY(launch_peer_barrier(b.flags, b.peer_flags, flag_set, b.rank, b.world, b.epochs, b.error_flag,
                               b.barrier_timeout_ms, stream));
  b.launches++;
  return 0;
}

// closes a sharded table on the main stream: barrier (the kernels already stored this rank's texels on every GPU)
// or all-gather callback
static int gather(Builder &b, float4 *table) {
  if (b.world <= 1) return 0;
  if (b.p2p) return peer_barrier(b, 0, b.main);
  if (!b.allgather) return 0;
  size_t bytes = (size_t)b.he_per_rank * b.ntex * sizeof(float4);
  if (b.allgather(b.allgather_user, table, bytes, (void *)b.main)) return fail("allgather callback failed");
  return 0;
}

#define TRY(expr)          \
  do {                     \
    if (expr) return 1;    \
  } while (0)
#define LAUNCH(expr) \
  do {               \
    CUDA_TRY(expr);   \
    b.launches++;     \
  } while (0)

// event pool: events are created on first use and reused by later runs
static int event_at(Builder &b, size_t index, cudaEvent_t &ev) {
  while (b.events.size() <= index) {
    cudaEvent_t e;
    CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    b.events.push_back(e);
  }
  ev = b.events[index];
  return 0;
}

// record an event on `from` and make `to` wait for it
static int order_after(Builder &b, size_t &cursor, cudaStream_t from, cudaStream_t to) {
  cudaEvent_t ev;
  if (event_at(b, cursor++, ev)) return 1;
  CUDA_TRY(cudaEventRecord(ev, from));
  CUDA_TRY(cudaStreamWaitEvent(to, ev, 0));
  return 0;
}

// generate-atmosphere-luts, atmosphere_lut.clj:43-105 (line numbers in the comments below), enqueued on the
// builder's two streams.
//
// The MAIN stream carries the chain that bounds the build:
//     first-order ray scatter -> [ blend tiles -> point scatter (dJ) -> ray scatter (dS) ] x iterations
// with one exchange (flag barrier or all-gather) per sharded table.  Everything that merely hangs off that chain
// runs on the SIDE stream, concurrently with the heavy kernels: the 2-D tables and per-direction constants (needed
// by the first point-scatter only), surface radiance dE_n = f(dS_{n-1}) (needed by the NEXT point scatter), the
// re-tabulated sums E += dE and S += dS (needed only at the very end), and the final file-layout tables.
// dS and dE are double buffered so the side stream can still read order n-1 while order n is written.
//
// Peer-to-peer mode shards the side stream's kernels as well (surface radiance by texel, the S accumulation and the
// file-layout tables by pair, the latter stored into rank 0's buffers only); the exchange after each point-scatter
// closes them together with dJ, so all cross-GPU barriers stay on the main stream.
static int builder_enqueue(Builder &b) {
  const Params &P = b.P;
  cudaStream_t st = b.main, side = b.side;
  b.stage_cursor = 0;
  b.launches = 0;
  size_t ev = 0;
  const Shard shard = b.shard;
  const Shard everything = {0, 1, 1};
  const bool sharded_side = b.p2p && b.world > 1;
  float4 *dsbuf[2] = {b.dS, b.dS2};
  const PeerOut dsout[2] = {b.peer_dS, b.peer_dS2};
  float4 *debuf[2] = {b.dE, b.dE_new};
  const PeerOut deout[2] = {b.peer_dE, b.peer_dE_new};
  const int N = b.iterations;
  const double *etab = b.exp_tab;
  const int e_first = sharded_side ? b.rank : 0, e_stride = sharded_side ? b.world : 1;

  CUDA_TRY(cudaMemsetAsync(b.counter, 0, 2 * sizeof(unsigned long long), st));
  CUDA_TRY(cudaMemsetAsync(b.error_flag, 0, sizeof(int), st));
  // peer-to-peer mode: nobody may store into a peer's tables before that peer has finished its previous run
  if (sharded_side) TRY(peer_barrier(b, 0, st));
  // the view rays of this rank's pairs, once: the first-order kernel and the ray-scatter records both start from them
  static const bool use_view_packs = !(getenv("ATMLUT_VIEW_PACK") && atoi(getenv("ATMLUT_VIEW_PACK")) == 0);   // A/B knob
  const void *packs = nullptr;
  if (use_view_packs) {
    if (!b.view_packs) CUDA_TRY(cudaMalloc(&b.view_packs, view_pack_total_bytes(P)));
    LAUNCH(launch_view_prepare(P, shard, b.he_count, b.view_packs, b.counter, st));
    packs = b.view_packs;
  }
  TRY(order_after(b, ev, st, side));   // the side stream starts after whatever the main stream did before

  // ---- side: 2-D tables and per-direction constants
  LAUNCH(launch_transmittance_table(P, b.T, side));                                   // :74
  LAUNCH(launch_surface_radiance_base(P, debuf[0], side));                            // :75  dE_0
  LAUNCH(launch_point_scatter_prepare(P, b.sphere_dirs, b.n_sphere, b.dir_info, side));
  if (b.n_half > 0) LAUNCH(launch_surface_radiance_prepare(P, b.half_dirs, b.n_half, b.half_info, side));
  cudaEvent_t e_prepared;
  TRY(event_at(b, ev++, e_prepared));
  CUDA_TRY(cudaEventRecord(e_prepared, side));
  // the view rays of this rank's pairs with everything the ray-scatter passes need per outer sample
  cudaEvent_t e_rays = nullptr;
  if (N > 0 && ray_scatter_uses_samples(P)) {
    if (!b.ray_samples) CUDA_TRY(cudaMalloc(&b.ray_samples, ray_sample_bytes(P, b.he_count)));
    LAUNCH(launch_ray_prepare(P, shard, b.he_count, b.ray_samples, b.counter + 1, packs, side));
    TRY(event_at(b, ev++, e_rays));
    CUDA_TRY(cudaEventRecord(e_rays, side));
  }
  LAUNCH(launch_resample_2d(P, 2, b.T, nullptr, nullptr, b.file_T, side));            // :98,102

  // ---- main: first order (both tables come out of one kernel: one exchange closes both in peer-to-peer mode)
  TRY(stage_begin(b, "first_order"));
  FirstOrderOut rayleigh = {b.peer_R1, 1, 0};                                         // :68,71,77
  FirstOrderOut mie_strength = {b.peer_M1, 0, 1};                                     // :69,72,78
  LAUNCH(launch_first_order(P, shard, b.he_count, rayleigh, mie_strength, b.counter, packs, st));
  TRY(stage_end(b));
  TRY(stage_begin(b, "first_order_exchange"));
  TRY(gather(b, b.R1));
  if (!b.p2p) TRY(gather(b, b.M1));
  TRY(stage_end(b));

  SSource ds = {b.R1, b.M1, P.medium.g[0]};                                           // :79-84  dS_0
  const float4 *s_cur = b.R1;                                                         // :85     S_0
  const float4 *e_cur = nullptr;                                                      // :76     E_0 = 0
  std::vector<cudaEvent_t> side_done(N + 1, nullptr);   // side finished reading dS_n

  // S += dS (atmosphere_lut.clj:96-97): this rank's pairs into every GPU's copy, or the whole table
  auto accumulate_s = [&](const float4 *ds_tab) -> int {
    if (sharded_side)
      LAUNCH(launch_resample_4d(P, shard, b.he_count, s_cur, ds_tab, b.peer_S_new, nullptr, side));
    else
      LAUNCH(launch_resample_4d(P, everything, b.n_he, s_cur, ds_tab, local_out(b.S_new), nullptr, side));
    std::swap(b.S, b.S_new);
    std::swap(b.peer_S, b.peer_S_new);
    s_cur = b.S;
    return 0;
  };
  // make-lookup-table of a 4-D table in file layout (:100-101, :104-105): every rank its own pairs, into rank 0
  auto file_table = [&](const float4 *tab, float *local_file, float *root_file) -> int {
    if (sharded_side)
      LAUNCH(launch_resample_4d(P, shard, b.he_count, tab, nullptr, local_out(nullptr), root_file, side));
    else if (b.rank == 0)
      LAUNCH(launch_resample_4d(P, everything, b.n_he, tab, nullptr, local_out(nullptr), local_file, side));
    return 0;
  };

  for (int it = 0; it < N; it++) {                                                    // :86
    char name[64];
    // ---- side, order it: dE_{it+1} = surface-radiance(dS_it) and S += dS_it (both read dS_it; sharded in
    // peer-to-peer mode: each rank stores its texels into every GPU's copy)
    TRY(order_after(b, ev, st, side));                                                // dS_it is complete
    if (it == 0) CUDA_TRY(cudaStreamWaitEvent(side, e_prepared, 0));
    float4 *de_next = debuf[(it + 1) & 1];
    if (b.n_half > 0)
      LAUNCH(launch_surface_radiance(P, ds, b.half_dirs, b.half_w, b.n_half, b.half_info, e_first, e_stride,
                                     deout[(it + 1) & 1], side));                     // :89,92
    else
      CUDA_TRY(cudaMemsetAsync(de_next, 0, (size_t)b.ne * sizeof(float4), side));
    if (it == 0) {
      TRY(file_table(b.M1, b.file_M, b.file_M_root));                                 // :101,105
      if (b.host_out[3])
        CUDA_TRY(cudaMemcpyAsync(b.host_out[3], b.file_M, (size_t)b.n4 * 12, cudaMemcpyDeviceToHost, side));
    } else
      TRY(accumulate_s(ds.tab_a));                                                    // :96-97
    TRY(event_at(b, ev++, side_done[it]));
    CUDA_TRY(cudaEventRecord(side_done[it], side));

    // ---- main: dJ_{it+1} = point-scatter(dS_it, dE_it)
    snprintf(name, sizeof name, "iter%d_point_scatter", it + 1);
    TRY(stage_begin(b, name));
    if (it == 0) CUDA_TRY(cudaStreamWaitEvent(st, e_prepared, 0));   // per-direction constants and dE_0 come from the side stream
    if (b.h_count > 0) {
      LAUNCH(launch_blend_dir_tiles(P, ds.tab_a, b.dir_info, b.n_sphere, b.h_first, b.h_stride, b.h_count, b.tiles_a, st));
      if (ds.tab_b)
        LAUNCH(launch_blend_dir_tiles(P, ds.tab_b, b.dir_info, b.n_sphere, b.h_first, b.h_stride, b.h_count, b.tiles_b, st));
    }
    // dE_it (it >= 1) was closed by the exchange after the previous point-scatter, on this stream
    LAUNCH(launch_point_scatter(P, shard, b.he_count, b.tiles_a, ds.tab_b ? b.tiles_b : nullptr, ds.phase_g,
                                debuf[it & 1], b.sphere_dirs, b.sphere_w, b.n_sphere, b.dir_info, etab, b.peer_dJ, st));  // :88,90
    // The exchange below closes THREE things at once: dJ_{it+1}; the side stream's sharded dE_{it+1} and S (this
    // rank's part is done once side_done[it] has fired, everybody's once the barrier is passed); and the
    // permission to overwrite the buffer of dS_{it-1}, which the side streams of all ranks have finished reading.
    // Every cross-GPU barrier thus sits on the main stream: no two barriers of one GPU are ever in flight together.
    CUDA_TRY(cudaStreamWaitEvent(st, side_done[it], 0));
    TRY(stage_end(b));
    snprintf(name, sizeof name, "iter%d_point_scatter_exchange", it + 1);
    TRY(stage_begin(b, name));
    TRY(gather(b, b.dJ));
    TRY(stage_end(b));
    // ---- side: E += dE_{it+1} needs the whole dE_{it+1}
    TRY(order_after(b, ev, st, side));
    LAUNCH(launch_resample_2d(P, 1, e_cur, de_next, b.Eacc_new, nullptr, side));      // :94-95
    std::swap(b.Eacc, b.Eacc_new);
    e_cur = b.Eacc;

    // ---- main: dS_{it+1} = ray-scatter(dJ_{it+1}) into the buffer that held dS_{it-1}
    snprintf(name, sizeof name, "iter%d_ray_scatter", it + 1);
    TRY(stage_begin(b, name));
    float4 *ds_next = dsbuf[(it + 1) & 1];
    if (it == 0 && e_rays) CUDA_TRY(cudaStreamWaitEvent(st, e_rays, 0));
    LAUNCH(launch_ray_scatter(P, shard, b.he_count, b.ray_samples, b.dJ, etab, dsout[(it + 1) & 1], b.counter + 1, st));  // :91,93
    TRY(stage_end(b));
    snprintf(name, sizeof name, "iter%d_ray_scatter_exchange", it + 1);
    TRY(stage_begin(b, name));
    TRY(gather(b, ds_next));
    TRY(stage_end(b));
    ds = SSource{ds_next, nullptr, 0.0};
  }

  // ---- last accumulation and the file-layout tables (kernels on the side stream, barriers on the main stream)
  TRY(stage_begin(b, "final_accumulate_and_files"));
  TRY(order_after(b, ev, st, side));
  if (N == 0) {
    CUDA_TRY(cudaStreamWaitEvent(side, e_prepared, 0));
    CUDA_TRY(cudaMemsetAsync(b.Eacc, 0, (size_t)b.ne * sizeof(float4), side));       // :76 E = 0
    e_cur = b.Eacc;
    TRY(file_table(b.M1, b.file_M, b.file_M_root));                                   // :101,105
    if (b.host_out[3])
      CUDA_TRY(cudaMemcpyAsync(b.host_out[3], b.file_M, (size_t)b.n4 * 12, cudaMemcpyDeviceToHost, side));
  } else {
    TRY(accumulate_s(ds.tab_a));                                                      // :96-97 (last order)
    if (sharded_side) {
      TRY(order_after(b, ev, side, st));
      TRY(peer_barrier(b, 0, st));                                                    // S is whole on every GPU
      TRY(order_after(b, ev, st, side));
    }
  }
  LAUNCH(launch_resample_2d(P, 1, e_cur, nullptr, nullptr, b.file_E, side));          // :99,103
  TRY(file_table(s_cur, b.file_S, b.file_S_root));                                    // :100,104
  if (b.host_out[2]) CUDA_TRY(cudaMemcpyAsync(b.host_out[2], b.file_S, (size_t)b.n4 * 12, cudaMemcpyDeviceToHost, side));
  if (b.host_out[1]) CUDA_TRY(cudaMemcpyAsync(b.host_out[1], b.file_E, (size_t)b.ne * 12, cudaMemcpyDeviceToHost, side));
  if (b.host_out[0]) CUDA_TRY(cudaMemcpyAsync(b.host_out[0], b.file_T, (size_t)b.nt * 12, cudaMemcpyDeviceToHost, side));
  TRY(order_after(b, ev, side, st));
  if (sharded_side) TRY(peer_barrier(b, 0, st));                                      // rank 0 holds every pair's texels
  TRY(stage_end(b));
  return 0;
}

// A build leaves b.S / b.S_new (and Eacc / Eacc_new) swapped an odd or even number of times.  Every enqueue must start
// from the same assignment, or a captured graph would not match a later eager run: remember and restore.
struct BufferRoles {
  float4 *S, *S_new, *Eacc, *Eacc_new;
  PeerOut peer_S, peer_S_new;
};

static BufferRoles save_roles(const Builder &b) { return BufferRoles{b.S, b.S_new, b.Eacc, b.Eacc_new, b.peer_S, b.peer_S_new}; }

static void restore_roles(Builder &b, const BufferRoles &r) {
  b.S = r.S;
  b.S_new = r.S_new;
  b.Eacc = r.Eacc;
  b.Eacc_new = r.Eacc_new;
  b.peer_S = r.peer_S;
  b.peer_S_new = r.peer_S_new;
}

// The DAG is captured once into a CUDA graph and replayed (unless the all-gather callback mode or the option forbids
// it): one submission instead of ~50 launches and ~25 event edges per build, which matters when a build takes a few
// milliseconds.  timed = true: a second graph of the same DAG that also records the per-stage events.
static int builder_run(Builder &b, bool timed) {
  CUDA_TRY(cudaSetDevice(b.device));
  static const bool graph_env = !(getenv("ATMLUT_GRAPH") && atoi(getenv("ATMLUT_GRAPH")) == 0);   // debugging aid
  const bool graph = b.use_graph && graph_env && !(b.world > 1 && !b.p2p);
  const BufferRoles roles = save_roles(b);
  b.timed = timed;
  int rc = 0;
  if (!graph) {
    rc = builder_enqueue(b);
  } else {
    cudaGraphExec_t &exec = b.graph_exec[timed ? 1 : 0];
    if (!exec) {
      cudaGraph_t g = nullptr;
      CUDA_TRY(cudaStreamBeginCapture(b.main, cudaStreamCaptureModeRelaxed));
      b.capturing = true;
      rc = builder_enqueue(b);
      b.capturing = false;
      cudaError_t e = cudaStreamEndCapture(b.main, &g);
      if (!rc && e != cudaSuccess) rc = fail_cuda(e, "cudaStreamEndCapture");
      if (!rc) {
        e = cudaGraphInstantiate(&exec, g, 0);
        if (e != cudaSuccess) rc = fail_cuda(e, "cudaGraphInstantiate");
      }
      if (g) cudaGraphDestroy(g);
    }
    if (!rc) CUDA_TRY(cudaGraphLaunch(exec, b.main));
  }
  restore_roles(b, roles);
  b.timed = false;
  if (!rc) b.ran = true;
  return rc;
}

}  // namespace atm

using namespace atm;

// ------------------------------------------------------------------ C ABI: lifetime

extern "C" int atmlut_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

namespace {
void drop_generate_cache();
void drop_multi_cache();
}

extern "C" int atmlut_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(std::string("no CUDA device available (there is no CPU fallback)") +
                (e != cudaSuccess ? std::string(": ") + cudaGetErrorString(e) : std::string()));
  if (device < 0 || device >= n) return fail("invalid device index");
  if (g_stream && g_device != device) {
    // the cached one-shot builders live on the old device; builders a host created itself own their streams
    drop_generate_cache();
    drop_multi_cache();
    cudaSetDevice(g_device);
    cudaStreamSynchronize(g_stream);
    cudaStreamDestroy(g_stream);
    g_stream = nullptr;
  }
  CUDA_TRY(cudaSetDevice(device));
  if (!g_stream) CUDA_TRY(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
  g_device = device;
  return 0;
}

extern "C" void *atmlut_stream(void) { return (void *)g_stream; }

extern "C" void atmlut_destroy(void) {
  drop_generate_cache();
  drop_multi_cache();
  for (int d = 0; d < 64; d++)
    if (g_exp_table[d]) {
      cudaSetDevice(d);
      cudaFree(g_exp_table[d]);
      g_exp_table[d] = nullptr;
    }
  if (g_device >= 0) cudaSetDevice(g_device);
  if (g_stream) {
    cudaStreamSynchronize(g_stream);
    cudaStreamDestroy(g_stream);
    g_stream = nullptr;
  }
  g_device = -1;
}

extern "C" const char *atmlut_last_error(void) { return g_error.c_str(); }

extern "C" void atmlut_default_config(atmlut_config *cfg) {
  // atmosphere_lut.clj:47-63
  cfg->height_size = 32;
  cfg->elevation_size = 127;
  cfg->light_elevation_size = 32;
  cfg->heading_size = 8;
  cfg->transmittance_height_size = 64;
  cfg->transmittance_elevation_size = 255;
  cfg->surface_height_size = 16;
  cfg->surface_sun_elevation_size = 63;
  cfg->ray_steps = 100;
  cfg->sphere_steps = 15;
  cfg->iterations = 5;
  cfg->intensity[0] = cfg->intensity[1] = cfg->intensity[2] = 1.0;
}

// ------------------------------------------------------------------ C ABI: builder

static int new_builder(const atmlut_planet *planet, const atmlut_scatter *scatter, int n, const atmlut_config *cfg,
                       int rank, int world, int device, Builder **out) {
  *out = nullptr;
  if (n != 2) return fail("generate-atmosphere-luts needs scatter = [mie rayleigh] (atmosphere_lut.clj:64)");
  if (world < 1 || rank < 0 || rank >= world) return fail("invalid rank/world");
  if (!cfg) return fail("config is NULL");
  if (cfg->iterations < 0) return fail("iterations must not be negative");
  CUDA_TRY(cudaSetDevice(device));
  Builder *b = new Builder();
  b->device = device;
  if (make_params(planet, scatter, n, cfg, b->P)) {
    delete b;
    return 1;
  }
  b->iterations = cfg->iterations;
  b->rank = rank;
  b->world = world;
  if (builder_alloc(*b)) {
    delete b;
    return 1;
  }
  *out = b;
  return 0;
}

extern "C" int atmlut_builder_create(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                     const atmlut_config *cfg, int rank, int world, void **builder) {
  if (!builder) return fail("builder is NULL");
  *builder = nullptr;
  if (ensure_init()) return 1;
  Builder *b = nullptr;
  if (new_builder(planet, scatter, n, cfg, rank, world, g_device, &b)) return 1;
  *builder = b;
  return 0;
}

// Host-only: the direction/weight list of integral-sphere (half = 0) or integral-half-sphere (half = 1) for the
// normal (1, 0, 0) of every table point, exactly as the kernels receive it.  Call with dirs = NULL to size.
extern "C" int atmlut_sphere_directions(int steps, int half, double *dirs, double *weights, int capacity) {
  if (steps < 1) return -1;
  std::vector<double> d, w;
  if (half)
    sphere_directions(steps >> 2, steps, kPi / 2, d, w);   // sphere.clj:96-99
  else
    sphere_directions(steps >> 1, steps, kPi, d, w);       // sphere.clj:102-105
  const int n = (int)w.size();
  if (dirs && weights) {
    if (capacity < n) return -1;
    memcpy(dirs, d.data(), d.size() * sizeof(double));
    memcpy(weights, w.data(), w.size() * sizeof(double));
  }
  return n;
}

extern "C" int atmlut_slab(int n_pairs, int rank, int world, int *begin, int *count, int *per_rank) {
  if (n_pairs < 0 || world < 1 || rank < 0 || rank >= world || !begin || !count || !per_rank)
    return fail("invalid argument");
  slab_of(n_pairs, rank, world, *begin, *count, *per_rank);
  return 0;
}

// ---- peer-to-peer mode: CUDA IPC handles of the sharded tables and the barrier flags ----

extern "C" int atmlut_builder_ipc_handle_bytes(void) { return kPeerTables * (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int atmlut_builder_ipc_export(void *builder, unsigned char *handles, int capacity_bytes) {
  Builder *b = (Builder *)builder;
  if (!b || !handles) return fail("invalid argument");
  if (capacity_bytes < atmlut_builder_ipc_handle_bytes()) return fail("handle buffer too small");
  CUDA_TRY(cudaSetDevice(b->device));
  void *ptrs[kPeerTables];
  peer_table_list(*b, ptrs);
  for (int i = 0; i < kPeerTables; i++) {
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, ptrs[i]));
    memcpy(handles + i * sizeof h, &h, sizeof h);
  }
  return 0;
}

extern "C" int atmlut_builder_ipc_import(void *builder, const unsigned char *all_handles, int world) {
  Builder *b = (Builder *)builder;
  if (!b || !all_handles) return fail("invalid argument");
  if (world != b->world) return fail("world size mismatch");
  if (world > kMaxPeers) return fail("peer-to-peer mode supports at most 8 GPUs");
  if (b->p2p) return fail("peer handles were already imported");
  if (b->ran) return fail("import the peer handles before the first run");
  CUDA_TRY(cudaSetDevice(b->device));
  for (int q = 0; q < world; q++) {
    void *ptrs[kPeerTables];
    if (q == b->rank) {
      peer_table_list(*b, ptrs);
    } else {
      for (int i = 0; i < kPeerTables; i++) {
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + ((size_t)q * kPeerTables + i) * sizeof h, sizeof h);
        CUDA_TRY(cudaIpcOpenMemHandle(&ptrs[i], h, cudaIpcMemLazyEnablePeerAccess));
        b->ipc_opened.push_back(ptrs[i]);
      }
    }
    add_peer(*b, q, ptrs);
  }
  b->p2p = true;
  use_row_sharding(*b);
  return 0;
}

static int check_peer_error(Builder *b) {
  if (!b->p2p) return 0;
  int err = 0;
  CUDA_TRY(cudaMemcpy(&err, b->error_flag, sizeof err, cudaMemcpyDeviceToHost));
  if (err)
    return fail("peer barrier timed out: a peer GPU never arrived (all ranks must call run within the barrier timeout "
                "of each other, see ATMLUT_OPT_BARRIER_TIMEOUT_MS)");
  return 0;
}

extern "C" int atmlut_builder_set_allgather(void *builder, atmlut_allgather_fn fn, void *user) {
  if (!builder) return fail("builder is NULL");
  Builder *b = (Builder *)builder;
  b->allgather = fn;
  b->allgather_user = user;
  return 0;
}

extern "C" int atmlut_builder_set_option(void *builder, int option, int value) {
  if (!builder) return fail("builder is NULL");
  Builder *b = (Builder *)builder;
  if (option == ATMLUT_OPT_GRAPH) {
    b->use_graph = value != 0;
  } else if (option == ATMLUT_OPT_BARRIER_TIMEOUT_MS) {
    if (value < 1) return fail("the barrier timeout must be positive");
    if (b->graph_exec[0] || b->graph_exec[1]) return fail("set the barrier timeout before the first run");
    b->barrier_timeout_ms = value;
  } else {
    return fail("unknown option");
  }
  return 0;
}

static int run_checked(void *builder, bool timed) {
  if (!builder) return fail("builder is NULL");
  Builder *b = (Builder *)builder;
  if (b->world > 1 && !b->allgather && !b->p2p)
    return fail("world > 1 needs an allgather callback or imported peer handles");
  return builder_run(*b, timed);
}

extern "C" int atmlut_builder_run(void *builder) { return run_checked(builder, false); }

extern "C" int atmlut_builder_run_timed(void *builder) { return run_checked(builder, true); }

extern "C" void *atmlut_builder_stream(void *builder) { return builder ? (void *)((Builder *)builder)->main : nullptr; }

extern "C" int atmlut_builder_sync(void *builder) {
  if (!builder) return fail("builder is NULL");
  Builder *b = (Builder *)builder;
  CUDA_TRY(cudaSetDevice(b->device));
  CUDA_TRY(cudaStreamSynchronize(b->main));
  return check_peer_error(b);
}

// Device -> host copy of one file table.  Pinned or registered destinations (atmlut_host_alloc, cudaHostRegister,
// torch pinned tensors) are written by the copy engine directly.  Pageable ones -- what a JVM arena or malloc hands
// out -- would make the driver stage the transfer in small synchronous pieces; instead it is pipelined through two
// library-owned pinned buffers: the copy engine fills one while the host thread empties the other.
static const size_t kStagingBytes = 4u << 20;

// neither allocated by cudaMallocHost / atmlut_host_alloc nor registered with cudaHostRegister
static bool is_pageable(const void *p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return true;
  }
  return attr.type == cudaMemoryTypeUnregistered;
}

static int copy_out(Builder *b, void *dst, const void *src, size_t bytes) {
  if (!is_pageable(dst) || bytes <= (256u << 10)) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, b->main));
    return 0;
  }
  for (int i = 0; i < 2; i++) {
    if (!b->staging[i]) CUDA_TRY(cudaMallocHost((void **)&b->staging[i], kStagingBytes));
    if (!b->staging_done[i]) CUDA_TRY(cudaEventCreateWithFlags(&b->staging_done[i], cudaEventDisableTiming));
  }
  const size_t chunks = (bytes + kStagingBytes - 1) / kStagingBytes;
  auto chunk_bytes = [&](size_t c) { return std::min(kStagingBytes, bytes - c * kStagingBytes); };
  CUDA_TRY(cudaMemcpyAsync(b->staging[0], src, chunk_bytes(0), cudaMemcpyDeviceToHost, b->main));
  CUDA_TRY(cudaEventRecord(b->staging_done[0], b->main));
  for (size_t c = 0; c < chunks; c++) {
    if (c + 1 < chunks) {
      CUDA_TRY(cudaMemcpyAsync(b->staging[(c + 1) & 1], (const unsigned char *)src + (c + 1) * kStagingBytes,
                               chunk_bytes(c + 1), cudaMemcpyDeviceToHost, b->main));
      CUDA_TRY(cudaEventRecord(b->staging_done[(c + 1) & 1], b->main));
    }
    CUDA_TRY(cudaEventSynchronize(b->staging_done[c & 1]));
    memcpy((unsigned char *)dst + c * kStagingBytes, b->staging[c & 1], chunk_bytes(c));
  }
  return 0;
}

extern "C" int atmlut_builder_download(void *builder, float *transmittance, float *surface_radiance,
                                       float *ray_scatter, float *mie_strength) {
  if (!builder) return fail("builder is NULL");
  Builder *b = (Builder *)builder;
  if (!b->ran) return fail("builder has not run");
  if (b->world > 1 && b->rank != 0) return fail("the file-layout tables of a sharded build are assembled on rank 0");
  CUDA_TRY(cudaSetDevice(b->device));
  // the large tables first: their staged copies overlap with whatever is still running
  if (ray_scatter && copy_out(b, ray_scatter, b->file_S, (size_t)b->n4 * 12)) return 1;
  if (mie_strength && copy_out(b, mie_strength, b->file_M, (size_t)b->n4 * 12)) return 1;
  if (transmittance && copy_out(b, transmittance, b->file_T, (size_t)b->nt * 12)) return 1;
  if (surface_radiance && copy_out(b, surface_radiance, b->file_E, (size_t)b->ne * 12)) return 1;
  CUDA_TRY(cudaStreamSynchronize(b->main));
  return check_peer_error(b);
}

// page-locked host memory for the output tables: the copy engine then writes them directly (a host that cannot
// pin its own buffers -- the JVM's arenas -- allocates its output segments here)
extern "C" void *atmlut_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (ensure_init()) return nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    fail("cudaMallocHost failed");
    return nullptr;
  }
  return p;
}

extern "C" void atmlut_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

extern "C" int atmlut_builder_stage_count(void *builder) {
  if (!builder) return 0;
  return (int)((Builder *)builder)->stages.size();
}

extern "C" const char *atmlut_builder_stage_name(void *builder, int stage) {
  Builder *b = (Builder *)builder;
  if (!b || stage < 0 || stage >= (int)b->stages.size()) return "";
  return b->stages[stage].name.c_str();
}

extern "C" int atmlut_builder_stage_ms(void *builder, int stage, float *ms) {
  Builder *b = (Builder *)builder;
  if (!b || stage < 0 || stage >= (int)b->stages.size()) return fail("invalid stage (stages are recorded by run_timed)");
  CUDA_TRY(cudaEventSynchronize(b->stages[stage].end));
  CUDA_TRY(cudaEventElapsedTime(ms, b->stages[stage].begin, b->stages[stage].end));
  return 0;
}

extern "C" int atmlut_builder_work(void *builder, double *esamples, double *lookups4d, double *lookups2d) {
  Builder *b = (Builder *)builder;
  if (!b || !b->ran) return fail("builder has not run");
  CUDA_TRY(cudaSetDevice(b->device));
  unsigned long long c[2];
  CUDA_TRY(cudaMemcpyAsync(c, b->counter, sizeof c, cudaMemcpyDeviceToHost, b->main));
  CUDA_TRY(cudaStreamSynchronize(b->main));
  std::vector<DirInfo> info((size_t)b->P.shapes.s4[0] * b->n_sphere);
  CUDA_TRY(cudaMemcpy(info.data(), b->dir_info, info.size() * sizeof(DirInfo), cudaMemcpyDeviceToHost));
  const Params &P = b->P;
  const double steps = P.shapes.ray_steps;
  double surf_dirs = 0;  // surface-hitting directions summed over this rank's (height, elevation) pairs
  for (int i = 0; i < b->he_count; i++) {
    const Shard &sh = b->shard;
    const int he = sh.run <= 1 ? sh.begin + i * sh.stride : sh.begin + (i / sh.run) * sh.stride + i % sh.run;
    for (int d = 0; d < b->n_sphere; d++) surf_dirs += info[(size_t)(he / P.shapes.s4[1]) * b->n_sphere + d].surface;
  }
  double surf_hd = 0;
  for (auto &i : info) surf_hd += i.surface;
  const double slab = (double)b->he_count * b->ntex;
  const double it = b->iterations;
  if (esamples) *esamples = (double)c[0] + (double)c[1] + (b->nt + b->ne + surf_hd) * steps;
  if (lookups4d)
    *lookups4d = slab * b->n_sphere * (it > 0 ? it + 1 : 0) + (double)b->ne * b->n_half * (it > 0 ? it + 1 : 0) +
                 slab * steps * it + (double)b->n4 * (2 * it + 2);
  if (lookups2d) *lookups2d = surf_dirs * b->ntex * it + (double)b->ne * (2 * it - (it > 0 ? 1 : 0) + 1) + b->nt;
  return 0;
}

extern "C" int atmlut_builder_counter(void *builder, int which, double *value) {
  Builder *b = (Builder *)builder;
  if (!b || !b->ran || !value) return fail("builder has not run");
  if (which == 2) {
    *value = (double)b->launches;
    return 0;
  }
  if (which == 3) {   // MUFU.EX2 instructions per overall-extinction sample of the fast sampler (atm_device.cuh)
    const Fast &f = b->P.fast;
    *value = !f.poly ? 0.0 : (f.pow_mode && f.degree == 2 ? (f.pow_alt ? 1.5 : 1.0) : 2.0);
    return 0;
  }
  if (which < 0 || which > 3) return fail("which must be 0, 1, 2 or 3");
  CUDA_TRY(cudaSetDevice(b->device));
  unsigned long long c[2];
  CUDA_TRY(cudaMemcpyAsync(c, b->counter, sizeof c, cudaMemcpyDeviceToHost, b->main));
  CUDA_TRY(cudaStreamSynchronize(b->main));
  *value = (double)c[which];
  return 0;
}

extern "C" int atmlut_builder_destroy(void *builder) {
  if (!builder) return 0;
  Builder *b = (Builder *)builder;
  delete b;
  if (g_device >= 0) cudaSetDevice(g_device);
  return 0;
}

// field-wise equality of the parameter structs (memcmp would also compare padding bytes)
static bool same_inputs(const atmlut_planet &pa, const atmlut_scatter *sa, const atmlut_config &ca,
                        const atmlut_planet &pb, const atmlut_scatter *sb, const atmlut_config &cb) {
  for (int i = 0; i < 3; i++)
    if (pa.centre[i] != pb.centre[i] || pa.brightness[i] != pb.brightness[i] || ca.intensity[i] != cb.intensity[i])
      return false;
  if (pa.radius != pb.radius || pa.height != pb.height) return false;
  for (int c = 0; c < 2; c++) {
    for (int i = 0; i < 3; i++)
      if (sa[c].base[i] != sb[c].base[i]) return false;
    if (sa[c].scale != sb[c].scale || sa[c].g != sb[c].g || sa[c].quotient != sb[c].quotient) return false;
  }
  return ca.height_size == cb.height_size && ca.elevation_size == cb.elevation_size &&
         ca.light_elevation_size == cb.light_elevation_size && ca.heading_size == cb.heading_size &&
         ca.transmittance_height_size == cb.transmittance_height_size &&
         ca.transmittance_elevation_size == cb.transmittance_elevation_size &&
         ca.surface_height_size == cb.surface_height_size &&
         ca.surface_sun_elevation_size == cb.surface_sun_elevation_size && ca.ray_steps == cb.ray_steps &&
         ca.sphere_steps == cb.sphere_steps && ca.iterations == cb.iterations;
}

// The one-shot call keeps its builder (device tables, quadrature tables, events) between calls with
// identical parameters, so a repeated build pays for kernels and the result copy only.
namespace {
struct GenerateCache {
  void *builder = nullptr;
  atmlut_planet planet;
  atmlut_scatter scatter[2];
  atmlut_config cfg;
  int device = -1;
} g_cache;

void drop_generate_cache() {
  if (g_cache.builder) atmlut_builder_destroy(g_cache.builder);
  g_cache.builder = nullptr;
}
}  // namespace

extern "C" int atmlut_generate(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                               const atmlut_config *cfg, float *transmittance, float *surface_radiance,
                               float *ray_scatter, float *mie_strength) {
  if (ensure_init()) return 1;
  if (!planet || !scatter || !cfg) return fail("planet, scatter and config must not be NULL");
  if (n != 2) return fail("generate-atmosphere-luts needs scatter = [mie rayleigh] (atmosphere_lut.clj:64)");
  const bool hit = g_cache.builder && g_cache.device == g_device &&
                   same_inputs(g_cache.planet, g_cache.scatter, g_cache.cfg, *planet, scatter, *cfg);
  if (!hit) {
    drop_generate_cache();
    if (atmlut_builder_create(planet, scatter, n, cfg, 0, 1, &g_cache.builder)) return 1;
    g_cache.planet = *planet;
    g_cache.scatter[0] = scatter[0];
    g_cache.scatter[1] = scatter[1];
    g_cache.cfg = *cfg;
    g_cache.device = g_device;
  }
  // Page-locked destinations: the downloads become nodes of the build graph (re-captured when the pointers change).
  Builder *b = (Builder *)g_cache.builder;
  float *dst[4] = {transmittance, surface_radiance, ray_scatter, mie_strength};
  bool in_graph = true;
  for (float *p : dst) in_graph = in_graph && p && !is_pageable(p);
  int rc = 0;
  if (in_graph) {
    if (memcmp(b->host_out, dst, sizeof dst) != 0) {
      CUDA_TRY(cudaStreamSynchronize(b->main));
      for (auto &g : b->graph_exec)
        if (g) {
          cudaGraphExecDestroy(g);
          g = nullptr;
        }
      memcpy(b->host_out, dst, sizeof dst);
    }
    rc = atmlut_builder_run(g_cache.builder);
    if (!rc) rc = atmlut_builder_sync(g_cache.builder);
  } else {
    if (b->host_out[0]) {
      CUDA_TRY(cudaStreamSynchronize(b->main));
      for (auto &g : b->graph_exec)
        if (g) {
          cudaGraphExecDestroy(g);
          g = nullptr;
        }
      memset(b->host_out, 0, sizeof b->host_out);
    }
    rc = atmlut_builder_run(g_cache.builder);
    if (!rc) rc = atmlut_builder_download(g_cache.builder, transmittance, surface_radiance, ray_scatter, mie_strength);
  }
  if (rc) drop_generate_cache();
  return rc;
}

// ------------------------------------------------------------------ C ABI: one process, several GPUs

// A host that cannot run one process per GPU (the JVM behind `clj -T:build`) drives all GPUs of the box from
// this single call.  Same peer-to-peer scheme as the multi-process mode (interleaved pairs, every finished
// texel stored into every GPU's tables, flag barriers), but the peer pointers come from
// cudaDeviceEnablePeerAccess instead of CUDA IPC, and one host thread enqueues every GPU's DAG in turn --
// nothing blocks on the host until the final synchronisation.
namespace {

struct MultiCache {
  std::vector<Builder *> group;
  atmlut_planet planet;
  atmlut_scatter scatter[2];
  atmlut_config cfg;
} g_multi;

void drop_multi_cache() {
  for (Builder *b : g_multi.group) atmlut_builder_destroy(b);
  g_multi.group.clear();
}

int create_group(const atmlut_planet *planet, const atmlut_scatter *scatter, const atmlut_config *cfg, int num_gpus,
                 std::vector<Builder *> &group) {
  for (int d = 0; d < num_gpus; d++)
    for (int q = 0; q < num_gpus; q++) {
      if (q == d) continue;
      int can = 0;
      CUDA_TRY(cudaDeviceCanAccessPeer(&can, d, q));
      if (!can) return fail("GPUs without peer access cannot share one build");
    }
  for (int d = 0; d < num_gpus; d++) {
    Builder *b = nullptr;
    if (new_builder(planet, scatter, 2, cfg, d, num_gpus, d, &b)) return 1;
    group.push_back(b);
  }
  for (int d = 0; d < num_gpus; d++) {
    CUDA_TRY(cudaSetDevice(d));
    for (int q = 0; q < num_gpus; q++) {
      if (q == d) continue;
      cudaError_t e = cudaDeviceEnablePeerAccess(q, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled)
        cudaGetLastError();
      else if (e != cudaSuccess)
        return fail_cuda(e, "cudaDeviceEnablePeerAccess");
    }
  }
  for (int d = 0; d < num_gpus; d++) {
    Builder *b = group[d];
    for (int q = 0; q < num_gpus; q++) {
      void *ptrs[kPeerTables];
      peer_table_list(*group[q], ptrs);
      add_peer(*b, q, ptrs);
    }
    b->p2p = true;
    use_row_sharding(*b);
  }
  return 0;
}

}  // namespace

extern "C" int atmlut_generate_multi(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                     const atmlut_config *cfg, int num_gpus, float *transmittance,
                                     float *surface_radiance, float *ray_scatter, float *mie_strength) {
  if (num_gpus == 1)
    return atmlut_generate(planet, scatter, n, cfg, transmittance, surface_radiance, ray_scatter, mie_strength);
  if (!planet || !scatter || !cfg) return fail("planet, scatter and config must not be NULL");
  if (n != 2) return fail("generate-atmosphere-luts needs scatter = [mie rayleigh] (atmosphere_lut.clj:64)");
  if (num_gpus < 1 || num_gpus > kMaxPeers || num_gpus > atmlut_device_count())
    return fail("num_gpus must be between 1 and min(8, number of CUDA devices)");
  if (cfg->iterations < 0) return fail("iterations must not be negative");
  const bool hit = (int)g_multi.group.size() == num_gpus &&
                   same_inputs(g_multi.planet, g_multi.scatter, g_multi.cfg, *planet, scatter, *cfg);
  if (!hit) {
    drop_multi_cache();
    if (create_group(planet, scatter, cfg, num_gpus, g_multi.group)) {
      drop_multi_cache();
      return 1;
    }
    g_multi.planet = *planet;
    g_multi.scatter[0] = scatter[0];
    g_multi.scatter[1] = scatter[1];
    g_multi.cfg = *cfg;
  }
  // every GPU's build is ONE graph launch: the host thread never blocks while a GPU waits for a peer in a barrier
  int rc = 0;
  for (Builder *b : g_multi.group)
    if (!rc) rc = builder_run(*b, false);
  for (Builder *b : g_multi.group)
    if (!rc) rc = atmlut_builder_sync(b);
  if (!rc)
    rc = atmlut_builder_download(g_multi.group[0], transmittance, surface_radiance, ray_scatter, mie_strength);
  if (rc) drop_multi_cache();
  if (g_device >= 0) cudaSetDevice(g_device);
  return rc;
}

// ------------------------------------------------------------------ C ABI: per-table entry points

namespace {

struct DevTables {
  std::vector<void *> ptrs;
  ~DevTables() {
    for (void *p : ptrs) cudaFree(p);
  }
  template <typename T>
  int alloc(T *&p, size_t count) {
    if (dev_alloc(p, count ? count : 1)) return 1;
    ptrs.push_back(p);
    return 0;
  }
  // host RGB float table -> device float4 table
  int upload_rgb(const float *host, long long texels, float4 *&dev) {
    float *staging = nullptr;
    if (alloc(staging, (size_t)texels * 3) || alloc(dev, (size_t)texels)) return 1;
    CUDA_TRY(cudaMemcpyAsync(staging, host, (size_t)texels * 12, cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(launch_rgb_to_float4(staging, dev, texels, g_stream));
    return 0;
  }
  int download_rgb(const float4 *dev, long long texels, float *host) {
    float *staging = nullptr;
    if (alloc(staging, (size_t)texels * 3)) return 1;
    CUDA_TRY(launch_float4_to_rgb(dev, staging, texels, g_stream));
    CUDA_TRY(cudaMemcpyAsync(host, staging, (size_t)texels * 12, cudaMemcpyDeviceToHost, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return 0;
  }
  int upload_doubles(const std::vector<double> &src, double *&dev) {
    if (alloc(dev, src.size())) return 1;
    CUDA_TRY(cudaMemcpyAsync(dev, src.data(), src.size() * sizeof(double), cudaMemcpyHostToDevice, g_stream));
    CUDA_TRY(cudaStreamSynchronize(g_stream));
    return 0;
  }
};

long long n4_of(const Params &P) { return (long long)P.shapes.s4[0] * P.shapes.s4[1] * P.shapes.s4[2] * P.shapes.s4[3]; }

}  // namespace

extern "C" int atmlut_transmittance_table(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                          const atmlut_config *cfg, float *out) {
  Params P;
  if (ensure_init() || make_params(planet, scatter, n, cfg, P)) return 1;
  if (!out) return fail("out is NULL");
  DevTables d;
  float4 *t = nullptr;
  long long nt = (long long)P.shapes.st[0] * P.shapes.st[1];
  if (d.alloc(t, (size_t)nt)) return 1;
  CUDA_TRY(launch_transmittance_table(P, t, g_stream));
  return d.download_rgb(t, nt, out);
}

extern "C" int atmlut_surface_radiance_base_table(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                                  const atmlut_config *cfg, float *out) {
  Params P;
  if (ensure_init() || make_params(planet, scatter, n, cfg, P)) return 1;
  if (!out) return fail("out is NULL");
  DevTables d;
  float4 *t = nullptr;
  long long ne = (long long)P.shapes.se[0] * P.shapes.se[1];
  if (d.alloc(t, (size_t)ne)) return 1;
  CUDA_TRY(launch_surface_radiance_base(P, t, g_stream));
  return d.download_rgb(t, ne, out);
}

extern "C" int atmlut_first_order_tables(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                         const atmlut_config *cfg, int component_a, int strength_a, float *out_a,
                                         int component_b, int strength_b, float *out_b) {
  Params P;
  if (ensure_init() || make_params(planet, scatter, n, cfg, P)) return 1;
  if (n < 1) return fail("first-order tables need at least one scatter component");
  if ((out_a && (component_a < 0 || component_a >= n)) || (out_b && (component_b < 0 || component_b >= n)))
    return fail("component index out of range");
  DevTables d;
  float4 *ta = nullptr, *tb = nullptr;
  const long long n4 = n4_of(P);
  if (out_a && d.alloc(ta, (size_t)n4)) return 1;
  if (out_b && d.alloc(tb, (size_t)n4)) return 1;
  FirstOrderOut oa = {local_out(ta), component_a, strength_a}, ob = {local_out(tb), component_b, strength_b};
  CUDA_TRY(launch_first_order(P, Shard{0, 1, 1}, P.shapes.s4[0] * P.shapes.s4[1], oa, ob, nullptr, nullptr, g_stream));
  if (out_a && d.download_rgb(ta, n4, out_a)) return 1;
  if (out_b && d.download_rgb(tb, n4, out_b)) return 1;
  CUDA_TRY(cudaStreamSynchronize(g_stream));
  return 0;
}

extern "C" int atmlut_point_scatter_table(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                          const atmlut_config *cfg, const float *ds_a, const float *ds_b,
                                          int phase_component, const float *de, float *out) {
  Params P;
  if (ensure_init() || make_params(planet, scatter, n, cfg, P)) return 1;
  if (!ds_a || !de || !out) return fail("ds_a, de and out must not be NULL");
  if (ds_b && (phase_component < 0 || phase_component >= n)) return fail("phase component out of range");
  DevTables d;
  const long long n4 = n4_of(P), ne = (long long)P.shapes.se[0] * P.shapes.se[1];
  float4 *a = nullptr, *b = nullptr, *e = nullptr, *o = nullptr;
  if (d.upload_rgb(ds_a, n4, a) || (ds_b && d.upload_rgb(ds_b, n4, b)) || d.upload_rgb(de, ne, e) ||
      d.alloc(o, (size_t)n4))
    return 1;
  std::vector<double> dirs, w;
  sphere_directions(P.shapes.sphere_steps >> 1, P.shapes.sphere_steps, kPi, dirs, w);
  if ((int)w.size() > kMaxDirs) return fail("sphere_steps too large for the point-scatter kernel");
  double *ddirs = nullptr, *dw = nullptr;
  DirInfo *info = nullptr;
  if (d.upload_doubles(dirs, ddirs) || d.upload_doubles(w, dw) || d.alloc(info, (size_t)P.shapes.s4[0] * w.size()))
    return 1;
  CUDA_TRY(launch_point_scatter_prepare(P, ddirs, (int)w.size(), info, g_stream));
  const double *etab = nullptr;
  if (exp_table(etab)) return 1;
  float4 *ta = nullptr, *tb = nullptr;
  const size_t tile_count = (size_t)P.shapes.s4[0] * w.size() * P.shapes.s4[2] * P.shapes.s4[3];
  if (d.alloc(ta, tile_count) || (b && d.alloc(tb, tile_count))) return 1;
  CUDA_TRY(launch_blend_dir_tiles(P, a, info, (int)w.size(), 0, 1, P.shapes.s4[0], ta, g_stream));
  if (b) CUDA_TRY(launch_blend_dir_tiles(P, b, info, (int)w.size(), 0, 1, P.shapes.s4[0], tb, g_stream));
  CUDA_TRY(launch_point_scatter(P, Shard{0, 1, 1}, P.shapes.s4[0] * P.shapes.s4[1], ta, tb,
                                ds_b ? P.medium.g[phase_component] : 0.0, e, ddirs, dw, (int)w.size(), info, etab,
                                local_out(o), g_stream));
  return d.download_rgb(o, n4, out);
}

extern "C" int atmlut_surface_radiance_table(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                             const atmlut_config *cfg, const float *ds_a, const float *ds_b,
                                             int phase_component, float *out) {
  Params P;
  if (ensure_init() || make_params(planet, scatter, n, cfg, P)) return 1;
  if (!ds_a || !out) return fail("ds_a and out must not be NULL");
  if (ds_b && (phase_component < 0 || phase_component >= n)) return fail("phase component out of range");
  DevTables d;
  const long long n4 = n4_of(P), ne = (long long)P.shapes.se[0] * P.shapes.se[1];
  float4 *a = nullptr, *b = nullptr, *o = nullptr;
  if (d.upload_rgb(ds_a, n4, a) || (ds_b && d.upload_rgb(ds_b, n4, b)) || d.alloc(o, (size_t)ne)) return 1;
  std::vector<double> dirs, w;
  sphere_directions(P.shapes.ray_steps >> 2, P.shapes.ray_steps, kPi / 2, dirs, w);
  if (w.empty()) {
    CUDA_TRY(cudaMemsetAsync(o, 0, (size_t)ne * sizeof(float4), g_stream));
    return d.download_rgb(o, ne, out);
  }
  double *ddirs = nullptr, *dw = nullptr;
  HalfDirInfo *info = nullptr;
  if (d.upload_doubles(dirs, ddirs) || d.upload_doubles(w, dw) || d.alloc(info, (size_t)P.shapes.se[0] * w.size()))
    return 1;
  CUDA_TRY(launch_surface_radiance_prepare(P, ddirs, (int)w.size(), info, g_stream));
  SSource src = {a, b, ds_b ? P.medium.g[phase_component] : 0.0};
  CUDA_TRY(launch_surface_radiance(P, src, ddirs, dw, (int)w.size(), info, 0, 1, local_out(o), g_stream));
  return d.download_rgb(o, ne, out);
}

extern "C" int atmlut_ray_scatter_table(const atmlut_planet *planet, const atmlut_scatter *scatter, int n,
                                        const atmlut_config *cfg, const float *dj, float *out) {
  Params P;
  if (ensure_init() || make_params(planet, scatter, n, cfg, P)) return 1;
  if (!dj || !out) return fail("dj and out must not be NULL");
  DevTables d;
  const long long n4 = n4_of(P);
  float4 *j = nullptr, *o = nullptr;
  if (d.upload_rgb(dj, n4, j) || d.alloc(o, (size_t)n4)) return 1;
  const double *etab = nullptr;
  if (exp_table(etab)) return 1;
  const int n_he = P.shapes.s4[0] * P.shapes.s4[1];
  unsigned char *samples = nullptr;
  if (ray_scatter_uses_samples(P)) {
    if (d.alloc(samples, ray_sample_bytes(P, n_he))) return 1;
    CUDA_TRY(launch_ray_prepare(P, Shard{0, 1, 1}, n_he, samples, nullptr, nullptr, g_stream));
  }
  CUDA_TRY(launch_ray_scatter(P, Shard{0, 1, 1}, n_he, samples, j, etab, local_out(o), nullptr, g_stream));
  return d.download_rgb(o, n4, out);
}

extern "C" int atmlut_resample_table(const atmlut_planet *planet, const atmlut_config *cfg, int which,
                                     const float *a, const float *b, float *out) {
  Params P;
  if (ensure_init() || make_params(planet, nullptr, 0, cfg, P)) return 1;
  if (!out) return fail("out is NULL");
  if (which < 0 || which > 2) return fail("which must be 0 (ray-scatter), 1 (surface-radiance) or 2 (transmittance)");
  DevTables d;
  const long long n = which == 0 ? n4_of(P)
                                 : which == 1 ? (long long)P.shapes.se[0] * P.shapes.se[1]
                                              : (long long)P.shapes.st[0] * P.shapes.st[1];
  float4 *da = nullptr, *db = nullptr, *o = nullptr;
  if ((a && d.upload_rgb(a, n, da)) || (b && d.upload_rgb(b, n, db)) || d.alloc(o, (size_t)n)) return 1;
  if (which == 0)
    CUDA_TRY(launch_resample_4d(P, Shard{0, 1, 1}, P.shapes.s4[0] * P.shapes.s4[1], da, db, local_out(o), nullptr, g_stream));
  else
    CUDA_TRY(launch_resample_2d(P, which, da, db, o, nullptr, g_stream));
  return d.download_rgb(o, n, out);
}

// ------------------------------------------------------------------ C ABI: output helpers

// image.clj:299-312 convert-4d-to-2d
extern "C" int atmlut_convert_4d_to_2d(const float *in, const int *shape, int ncomp, float *out) {
  if (!in || !shape || !out || ncomp < 1) return fail("invalid argument");
  const long long d = shape[0], c = shape[1], b = shape[2], a = shape[3];
  const long long h = d * b, w = c * a;
  for (long long y = 0; y < h; y++)
    for (long long x = 0; x < w; x++) {
      long long src = (((y / b) * c + (x / a)) * b + (y % b)) * a + (x % a);
      for (int k = 0; k < ncomp; k++) out[(y * w + x) * ncomp + k] = in[src * ncomp + k];
    }
  return 0;
}

// util.clj:227-240 spit-floats: headerless little-endian float32
extern "C" int atmlut_write_floats(const char *path, const float *data, long count) {
  if (!path || (!data && count > 0)) return fail("invalid argument");
  FILE *f = fopen(path, "wb");
  if (!f) return fail(std::string("cannot open ") + path);
  std::vector<unsigned char> buf((size_t)count * 4);
  for (long i = 0; i < count; i++) {
    unsigned int bits;
    memcpy(&bits, &data[i], 4);
    buf[4 * i] = (unsigned char)(bits & 255);
    buf[4 * i + 1] = (unsigned char)((bits >> 8) & 255);
    buf[4 * i + 2] = (unsigned char)((bits >> 16) & 255);
    buf[4 * i + 3] = (unsigned char)((bits >> 24) & 255);
  }
  size_t written = fwrite(buf.data(), 1, buf.size(), f);
  if (fclose(f) != 0 || written != buf.size()) return fail(std::string("short write to ") + path);
  return 0;
}

// util.clj:188-203 slurp-floats
extern "C" long atmlut_read_floats(const char *path, float *data, long max_count) {
  if (!path || !data) return -1;
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  long n = 0;
  unsigned char b[4];
  while (n < max_count && fread(b, 1, 4, f) == 4) {
    unsigned int bits = (unsigned)b[0] | ((unsigned)b[1] << 8) | ((unsigned)b[2] << 16) | ((unsigned)b[3] << 24);
    memcpy(&data[n++], &bits, 4);
  }
  fclose(f);
  return n;
}
