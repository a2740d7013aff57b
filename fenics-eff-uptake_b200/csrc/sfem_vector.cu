// Vector kernels: axpby family, deterministic two-stage dot products, diagonal extraction, dense
// coarse-level gemv, concentration post-processing (solvers.py:86-105,154-173).
#include "sfem_common.cuh"
#include "sfem_internal.h"
#include "sfem_graph.h"
#include "sfem_dist.h"

#include <cmath>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace sfem {

// ------------------------------------------------------------------ global state
static thread_local std::string t_last_error;
std::atomic<long long> g_launches{0};

void set_error(const std::string& msg) { t_last_error = msg; }

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      sms = v;
    else
      sms = 148;
  }
  return sms;
}

// ------------------------------------------------------------------ per-launch profiler
namespace {
struct ProfRec { int cat; double bytes; cudaEvent_t e0, e1; };
std::atomic<bool> g_prof_on{false};
std::vector<ProfRec> g_prof;
size_t g_prof_max = 0;
}  // namespace

namespace { std::atomic<unsigned long long> g_graph_epoch{1}; }
unsigned long long graph_epoch() { return g_graph_epoch.load(std::memory_order_acquire); }
void graph_epoch_bump() { g_graph_epoch.fetch_add(1, std::memory_order_acq_rel); }

bool profiling_active() { return g_prof_on.load(std::memory_order_relaxed); }
bool graphs_enabled() {
  static int env = -1;
  if (env < 0) {
    const char* e = std::getenv("SFEM_GRAPHS");
    env = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return env == 1 && !profiling_active();
}

Prof::Prof(int cat, double bytes, cudaStream_t s) : idx(-1), st(s) {
  if (!g_prof_on.load(std::memory_order_relaxed) || g_prof.size() >= g_prof_max) return;
  ProfRec r;
  r.cat = cat; r.bytes = bytes;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, st);
  g_prof.push_back(r);
  idx = (int)g_prof.size() - 1;
}
Prof::~Prof() {
  if (idx >= 0) cudaEventRecord(g_prof[idx].e1, st);
}

// ------------------------------------------------------------------ kernels
__global__ void k_set(int n, double a, double* __restrict__ x) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = a;
}

__global__ void k_axpby(int n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    y[i] = (b == 0.0) ? a * x[i] : fma(a, x[i], b * y[i]);
}

__global__ void k_mul_scale(int n, double a, const double* __restrict__ d, const double* x,
                            double* y) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = a * d[i] * x[i];
}

template <bool VEC>
__global__ void __launch_bounds__(kThreads) k_dot_partial(int n, const double* __restrict__ x,
                                                          const double* __restrict__ y, double* __restrict__ partial) {
  __shared__ double sh[33];
  double acc = 0.0, acc1 = 0.0;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  if (VEC) {
    const int n2 = n >> 1;
    const double2* X2 = reinterpret_cast<const double2*>(x);
    const double2* Y2 = reinterpret_cast<const double2*>(y);
    for (int i = tid; i < n2; i += 2 * nt) {
      const int j = i + nt;
      const double2 x0 = X2[i], y0 = Y2[i];
      acc = fma(x0.x, y0.x, acc);
      acc = fma(x0.y, y0.y, acc);
      if (j < n2) {
        const double2 x1 = X2[j], y1 = Y2[j];
        acc1 = fma(x1.x, y1.x, acc1);
        acc1 = fma(x1.y, y1.y, acc1);
      }
    }
    if ((n & 1) && tid == 0) acc = fma(x[n - 1], y[n - 1], acc);
    acc += acc1;
  } else {
    for (int i = tid; i < n; i += nt) acc = fma(x[i], y[i], acc);
  }
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// out[0] = sum of the partials, summed over all ranks when a communicator is active
__global__ void __launch_bounds__(kThreads) k_sum_partials(const double* __restrict__ partial, int n,
                                                           double* __restrict__ out, DistDev D) {
  __shared__ double sh[33];
  double t = block_sum_array(partial, n, sh);
  if (threadIdx.x == 0) {
    dist_allreduce_scalars(D, &t, 1);
    out[0] = t;
  }
}

__global__ void k_diag_inv(int n, const int* __restrict__ rowptr, const int* __restrict__ cols,
                           const double* __restrict__ vals, double* __restrict__ dinv) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double d = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
      if (cols[k] == i) d = vals[k];
    dinv[i] = (d != 0.0) ? 1.0 / d : 1.0;
  }
}

// y = M x for a small dense row-major matrix (coarsest multigrid level): one warp per row.
template <int NB>
__global__ void __launch_bounds__(kThreads) k_dense_gemv(int n, const double* __restrict__ M,
                                                         const double* __restrict__ x, double* __restrict__ y) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < n; row += gridDim.x * warps_per_block) {
    const double* m = M + (size_t)row * n;
    double acc[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) acc[c] = 0.0;
    // four matrix entries (and their right-hand-side values) per round: the 18 column steps of a 556-row inverse are
    // otherwise a chain of dependent L2 round trips (the launch ran 12.6 us, ncu r02)
    for (int j0 = lane; j0 < n; j0 += 128) {
      double mj[4], xv[4][NB];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 32 * u;
        const bool ok = j < n;
        mj[u] = ok ? m[j] : 0.0;
#pragma unroll
        for (int c = 0; c < NB; ++c) xv[u][c] = ok ? x[(size_t)j * NB + c] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int c = 0; c < NB; ++c) acc[c] = fma(mj[u], xv[u][c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      const double t = warp_sum(acc[c]);
      if (lane == 0) y[(size_t)row * NB + c] = t;
    }
  }
}

// stats: [0] #nonfinite, [1] #negative, [2] min, [3] max, [4] sum   (per block partials, 5 each)
__global__ void __launch_bounds__(kThreads) k_conc_stats(int n, double* __restrict__ c, int fix_nonfinite,
                                                         double* __restrict__ partial) {
  __shared__ double sh[33];
  double nbad = 0.0, nneg = 0.0, mn = INFINITY, mx = -INFINITY, sum = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double v = c[i];
    if (!isfinite(v)) {
      nbad += 1.0;
      if (fix_nonfinite) { v = 0.0; c[i] = 0.0; } else continue;
    }
    if (v < 0.0) nneg += 1.0;
    mn = fmin(mn, v);
    mx = fmax(mx, v);
    sum += v;
  }
  const double a = block_sum(nbad, sh);
  const double bq = block_sum(nneg, sh);
  const double s = block_sum(sum, sh);
  // min / max through the same shared scratch
  for (int o = 16; o > 0; o >>= 1) {
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __shared__ double smn[32], smx[32];
  if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w) { mn = fmin(mn, smn[w]); mx = fmax(mx, smx[w]); }
    double* p = partial + 5 * blockIdx.x;
    p[0] = a; p[1] = bq; p[2] = mn; p[3] = mx; p[4] = s;
  }
}

__global__ void k_clamp_negative(int n, double* __restrict__ c) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (c[i] < 0.0) c[i] = 0.0;
}

// ------------------------------------------------------------------ host wrappers
int vec_set(int n, double a, double* x, cudaStream_t st) {
  if (n <= 0) return SFEM_OK;
  k_set<<<grid_for(n, kThreads * 4), kThreads, 0, st>>>(n, a, x);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int vec_copy(int n, const double* x, double* y, cudaStream_t st) {
  if (n <= 0) return SFEM_OK;
  Prof prof(PC_VEC, 16.0 * n, st);
  SFEM_CUDA(cudaMemcpyAsync(y, x, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  return SFEM_OK;
}

int vec_axpby(int n, double a, const double* x, double b, double* y, cudaStream_t st) {
  if (n <= 0) return SFEM_OK;
  Prof prof(PC_VEC, 8.0 * n * (b == 0.0 ? 2 : 3), st);
  k_axpby<<<grid_for(n, kThreads * 4), kThreads, 0, st>>>(n, a, x, b, y);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int vec_mul_scale(int n, double a, const double* d, const double* x, double* y, cudaStream_t st) {
  if (n <= 0) return SFEM_OK;
  Prof prof(PC_VEC, 24.0 * n, st);
  k_mul_scale<<<grid_for(n, kThreads * 4), kThreads, 0, st>>>(n, a, d, x, y);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int vec_dot_partial(int n, const double* x, const double* y, double* partial, int* nparts, cudaStream_t st) {
  const int grid = grid_for(n, kThreads * 4, 4);
  Prof prof(PC_VEC, 16.0 * n, st);
  if (aligned16(x, y)) k_dot_partial<true><<<grid, kThreads, 0, st>>>(n, x, y, partial);
  else k_dot_partial<false><<<grid, kThreads, 0, st>>>(n, x, y, partial);
  SFEM_LAUNCH_CHECK();
  *nparts = grid;
  return SFEM_OK;
}

DistDev dist_dev() {
  Dist* d = active_dist();
  return d ? d->dev : DistDev();
}

int vec_sum_partials(const double* partial, int n, double* out, cudaStream_t st) {
  k_sum_partials<<<1, kThreads, 0, st>>>(partial, n, out, dist_dev());
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

// scratch: kMaxPartials + 1 doubles
int vec_dot_host(int n, const double* x, const double* y, double* scratch, double* h_out, cudaStream_t st) {
  int np = 0;
  SFEM_TRY(vec_dot_partial(n, x, y, scratch, &np, st));
  k_sum_partials<<<1, kThreads, 0, st>>>(scratch, np, scratch + kMaxPartials, dist_dev());
  SFEM_LAUNCH_CHECK();
  SFEM_CUDA(cudaMemcpyAsync(h_out, scratch + kMaxPartials, sizeof(double), cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  return SFEM_OK;
}

int extract_diag_inv(const Csr& A, double* dinv, cudaStream_t st) {
  if (A.nrows <= 0) return SFEM_OK;
  k_diag_inv<<<grid_for(A.nrows, kThreads), kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, dinv);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int dense_gemv(int n, const double* M, const double* x, double* y, cudaStream_t st, int nb) {
  if (n <= 0) return SFEM_OK;
  Prof prof(PC_OTHER, 8.0 * n * n + 16.0 * n * nb, st);
  if (nb == 2) k_dense_gemv<2><<<grid_for(n, kThreads / 32), kThreads, 0, st>>>(n, M, x, y);
  else k_dense_gemv<1><<<grid_for(n, kThreads / 32), kThreads, 0, st>>>(n, M, x, y);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // namespace sfem

// ------------------------------------------------------------------ C ABI
using namespace sfem;

extern "C" {

const char* sfem_last_error(void) { return t_last_error.c_str(); }
int sfem_version(void) { return 100; }
int sfem_device_sms(void) { return num_sms(); }
long long sfem_launch_count(void) { return g_launches.load(); }
void sfem_launch_count_reset(void) { g_launches.store(0); }

int sfem_profile_start(int max_records) {
  for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
  g_prof_max = max_records > 0 ? (size_t)max_records : 0;
  g_prof.reserve(g_prof_max);
  g_prof_on.store(true);
  return SFEM_OK;
}

/* Stops profiling, synchronises the device and copies up to `cap` records out; returns the count. */
int sfem_profile_stop(int cap, int* h_cat, double* h_bytes, float* h_ms) {
  g_prof_on.store(false);
  SFEM_CUDA(cudaDeviceSynchronize());
  int n = 0;
  for (auto& r : g_prof) {
    if (n < cap) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) ms = -1.f;
      h_cat[n] = r.cat; h_bytes[n] = r.bytes; h_ms[n] = ms;
      ++n;
    }
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  return n;
}

namespace sfem {
__global__ void __launch_bounds__(kThreads) k_vec_select(int n, const unsigned char* __restrict__ flag, const double* a,
                                                         const double* b, double* out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[i] = flag[i] ? a[i] : (b ? b[i] : 0.0);
}
}  // namespace sfem

int sfem_vec_select(int n, const unsigned char* flag, const double* a, const double* b, double* out, void* stream) {
  if (n < 0 || (n > 0 && (!flag || !a || !out))) { set_error("vec_select: bad arguments"); return SFEM_ERR_ARG; }
  if (n == 0) return SFEM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  { Prof prof(PC_VEC, 25.0 * n, st);
  k_vec_select<<<grid_for(n, kThreads * 4), kThreads, 0, st>>>(n, flag, a, b, out); }
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_vec_copy(int n, const double* x, double* y, void* stream) { return vec_copy(n, x, y, (cudaStream_t)stream); }

int sfem_vec_axpby(int n, double a, const double* x, double b, double* y, void* stream) {
  return vec_axpby(n, a, x, b, y, (cudaStream_t)stream);
}

int sfem_vec_dot(int n, const double* x, const double* y, double* h_out, void* stream) {
  double* scratch = nullptr;
  SFEM_CUDA(cudaMalloc(&scratch, (kMaxPartials + 1) * sizeof(double)));
  int r = vec_dot_host(n, x, y, scratch, h_out, (cudaStream_t)stream);
  cudaFree(scratch);
  return r;
}

int sfem_extract_diag_inv(int n, const int* rowptr, const int* cols, const double* vals, double* dinv, void* stream) {
  Csr A;
  A.nrows = A.ncols = n;
  A.rowptr = rowptr; A.cols = cols; A.vals = vals;
  return extract_diag_inv(A, dinv, (cudaStream_t)stream);
}

int sfem_postprocess_concentration(int n, double* c, int fix_nonfinite, double* h_stats, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) { set_error("empty vector"); return SFEM_ERR_ARG; }
  const int grid = grid_for(n, kThreads * 4, 4);
  // per-thread scratch, allocated once (cudaMalloc / cudaFree per call would synchronise the whole device: a sweep calls
  // this once per case, possibly from several solver threads at once)
  static thread_local double* partial = nullptr;
  if (partial == nullptr) SFEM_CUDA(cudaMalloc(&partial, (size_t)kMaxPartials * 5 * sizeof(double)));
  k_conc_stats<<<grid, kThreads, 0, st>>>(n, c, fix_nonfinite, partial);
  g_launches.fetch_add(1);
  std::vector<double> h((size_t)grid * 5);
  cudaError_t e = cudaMemcpyAsync(h.data(), partial, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return SFEM_ERR_CUDA; }
  double nbad = 0, nneg = 0, mn = INFINITY, mx = -INFINITY, sum = 0;
  for (int b = 0; b < grid; ++b) {
    nbad += h[5 * b]; nneg += h[5 * b + 1];
    mn = std::fmin(mn, h[5 * b + 2]); mx = std::fmax(mx, h[5 * b + 3]); sum += h[5 * b + 4];
  }
  double clamped = 0.0;
  if (nneg > 0 && std::fabs(mn) < 1e-12) {     // solvers.py:161-164: only tiny negatives are clipped
    k_clamp_negative<<<grid_for(n, kThreads * 4), kThreads, 0, st>>>(n, c);
    g_launches.fetch_add(1);
    clamped = 1.0;
  }
  h_stats[0] = nbad; h_stats[1] = nneg; h_stats[2] = mn; h_stats[3] = mx; h_stats[4] = sum / n; h_stats[5] = clamped;
  return SFEM_OK;
}

}  // extern "C"
