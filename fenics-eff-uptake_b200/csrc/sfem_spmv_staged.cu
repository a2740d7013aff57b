// Shared-memory-staged CSR SpMV family for sm_100a: warp-specialised TMA pipeline, nnz-parallel
// products, row reduction out of shared memory.
//
// Why: with LANES lanes per row the vector engine (sfem_spmv.cu) spreads every warp-level load of
// vals / cols over ~8 rows, i.e. ~10 distinct 128-byte lines per instruction; the L1 tag stage
// (one line per cycle) then caps the kernel near 4 TB/s on B200 whatever the occupancy (measured).
// Here the matrix stream never touches the LSU/L1 path:
//   * a tile = consecutive rows with at most `cap` entries and at most 128 rows (host-built plan,
//     sfem_staged_plan); rows are stored back to back in CSR, so the tile's vals / cols / rowptr
//     slice are three contiguous spans that ONE elected producer thread streams HBM -> shared memory
//     with cp.async.bulk (TMA, UBLKCP in SASS) through a `stages`-deep full/empty mbarrier ring;
//   * phase 1 (256 consumer threads): entry-parallel -- thread t takes entries t, t+256, ... of the
//     tile, reads (col, val) from shared memory, gathers x (L2-resident) with several independent
//     loads in flight, and writes the product back over the value in shared memory.  Perfectly
//     balanced whatever the row lengths, coalesced/conflict-free shared-memory traffic;
//   * phase 2: one lane pair per row sums the row's products from shared memory in a fixed order
//     (bit-reproducible) and runs the fused epilogue (sfem_spmv_epi.cuh), whose operands were
//     requested before phase 1.
// NB = 2 interleaved right-hand sides share the matrix stream (second product array in the stage).
//
// Requirements (checked): rowptr/cols/vals 16-byte aligned and readable 4 elements past their end.
#include "sfem_common.cuh"
#include "sfem_internal.h"
#include "sfem_spmv_epi.cuh"
#include "sfem_tma.cuh"

#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace sfem {

namespace {

constexpr int kProducerThreads = 32;
constexpr int kMaxStages = 8;
constexpr int kMaxTileRows = 128;         // rows per tile = consumer lane pairs = CT / 2 (CT = 128 or 256)
constexpr int kPhase1Unroll = 6;

// barrier among the consumer warps only (the producer warp never joins)
template <int CT>
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(CT) : "memory"); }

template <int NB>
__host__ __device__ constexpr size_t stage_bytes(int cap, int rp_cap) {
  return (size_t)cap * (NB == 2 ? 20 : 12) + (size_t)rp_cap * 4;
}

template <int CT, int NB, class Epi, bool REDUCE>
__global__ void __launch_bounds__(CT + kProducerThreads)
    k_staged(int ntiles, const int* __restrict__ tile_row, int stages, int cap, int rp_cap,
             const int* __restrict__ rowptr, const int* __restrict__ cols, const double* __restrict__ vals,
             const double* __restrict__ x, Epi epi, double* __restrict__ partial) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];
  constexpr int kConsumerThreads = CT;
  __shared__ double red[kConsumerThreads / 32];

  const size_t sbytes = stage_bytes<NB>(cap, rp_cap);
  const int warp = threadIdx.x >> 5;
  const int lane32 = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerThreads / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0) {
    // ------------------------------------------------------------------ producer (one elected thread)
    if (lane32 == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % stages;
        const uint32_t ph = (uint32_t)((it / stages) & 1);
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const int r0 = tile_row[tile], r1 = tile_row[tile + 1];
        const int k0 = rowptr[r0] & ~3;
        const int k1 = (rowptr[r1] + 3) & ~3;
        const int cnt = k1 - k0;
        const int rb = r0 & ~3;                                  // 16-byte aligned start of the rowptr slice
        const int rcnt = ((r1 - rb + 1) + 3) & ~3;
        unsigned char* base = smem + (size_t)s * sbytes;
        double* sv = reinterpret_cast<double*>(base);
        int* sc = reinterpret_cast<int*>(base + (size_t)cap * (NB == 2 ? 16 : 8));
        int* sr = sc + cap;
        mbar_expect_tx(&full_bar[s], (uint32_t)(cnt * 12 + rcnt * 4));
        if (cnt > 0) {
          bulk_g2s(sv, vals + k0, (uint32_t)(cnt * 8), &full_bar[s]);
          bulk_g2s(sc, cols + k0, (uint32_t)(cnt * 4), &full_bar[s]);
        }
        bulk_g2s(sr, rowptr + rb, (uint32_t)(rcnt * 4), &full_bar[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ consumers
    const int ct = threadIdx.x - kProducerThreads;      // 0..255
    const int pair = ct >> 1, half = ct & 1;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it % stages;
      const uint32_t ph = (uint32_t)((it / stages) & 1);
      const int r0 = tile_row[tile];
      const int nr = tile_row[tile + 1] - r0;
      // epilogue operands of this thread's row: requested before the tile is even in shared memory
      const int row = r0 + pair;
      const bool active = pair < nr && half < NB;
      const typename Epi::Pre pre = epi.pre(row, half, active);
      mbar_wait(&full_bar[s], ph);
      unsigned char* base = smem + (size_t)s * sbytes;
      double* sv = reinterpret_cast<double*>(base);
      double* sp = sv + cap;                                         // second product array (NB == 2)
      const int* sc = reinterpret_cast<const int*>(base + (size_t)cap * (NB == 2 ? 16 : 8));
      const int* sr = sc + cap + (r0 & 3);                         // slice was loaded from r0 & ~3
      const int k0 = sr[0] & ~3;
      const int kb = sr[0] - k0, ke = sr[nr] - k0;
      // phase 1: entry-parallel products
      for (int k = kb + ct; k < ke; k += kPhase1Unroll * kConsumerThreads) {
        int c[kPhase1Unroll];
#pragma unroll
        for (int j = 0; j < kPhase1Unroll; ++j) {
          const int kk = k + j * kConsumerThreads;
          c[j] = (kk < ke) ? sc[kk] : -1;
        }
        if (NB == 1) {
          double xv[kPhase1Unroll];
#pragma unroll
          for (int j = 0; j < kPhase1Unroll; ++j) xv[j] = (c[j] >= 0) ? __ldg(x + c[j]) : 0.0;
#pragma unroll
          for (int j = 0; j < kPhase1Unroll; ++j) {
            const int kk = k + j * kConsumerThreads;
            if (c[j] >= 0) sv[kk] *= xv[j];
          }
        } else {
          double2 xv[kPhase1Unroll];
#pragma unroll
          for (int j = 0; j < kPhase1Unroll; ++j)
            xv[j] = (c[j] >= 0) ? __ldg(reinterpret_cast<const double2*>(x) + c[j]) : make_double2(0.0, 0.0);
#pragma unroll
          for (int j = 0; j < kPhase1Unroll; ++j) {
            const int kk = k + j * kConsumerThreads;
            if (c[j] >= 0) {
              const double v = sv[kk];
              sv[kk] = v * xv[j].x;
              sp[kk] = v * xv[j].y;
            }
          }
        }
      }
      consumer_sync<CT>();
      // phase 2: one lane pair per row, fixed summation order
      {
        double a0 = 0.0, a1 = 0.0;
        if (pair < nr) {
          const int ks = sr[pair] - k0, kend = sr[pair + 1] - k0;
          for (int k = ks + half; k < kend; k += 2) {
            a0 += sv[k];
            if (NB == 2) a1 += sp[k];
          }
        }
        a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
        if (NB == 2) a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
        if (active) epi.fin(row, half, (NB == 2 && half == 1) ? a1 : a0, pre);
      }
      __syncwarp();
      if (lane32 == 0) mbar_arrive(&empty_bar[s]);
    }
    if (REDUCE) {
      // deterministic sum of the consumers' dot contributions (fixed order)
      double a = warp_sum(epi_acc_of(epi));
      if (lane32 == 0) red[warp - 1] = a;
      consumer_sync<CT>();
      if (ct == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kConsumerThreads / 32; ++w) t += red[w];
        partial[blockIdx.x] = t;
      }
    }
  }
}

// ------------------------------------------------------------------ plan registry
// Tile plans are attached to a matrix by the device address of its rowptr array, so every call
// site that sees the same CSR (multigrid levels, Krylov drivers, the C ABI) picks the staged engine
// without carrying the plan around.
struct Plan {
  const int* tile_row = nullptr;   // device, [ntiles + 1]
  int ntiles = 0;
  int cap = 0;                     // entries per stage (multiple of 4), includes the 16-byte rounding slack
  int nrows = 0;
  int max_rows = 128;              // rows per tile (64 -> 128 consumer threads, 128 -> 256)
};
std::mutex g_plan_mu;
std::atomic<int> g_min_tiles{-2};          // -2: read SFEM_STAGED_MIN_TILES on first use; <= 0: default (SM count)
std::unordered_map<const int*, Plan> g_plans;

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

template <int CT, int NB, class Epi, bool REDUCE>
int launch_staged_ct(const Plan& P, const Csr& A, const double* x, const Epi& epi, double* partial, int* nparts,
                     cudaStream_t st) {
  constexpr int kConsumerThreads = CT;
  static const int stages_env = env_int("SFEM_STAGED_STAGES", 0);
  static const int ctas_env = env_int("SFEM_STAGED_CTAS_PER_SM", 0);
  const int rp_cap = ((CT / 2 + 1 + 3) + 3) & ~3;
  const size_t sb = stage_bytes<NB>(P.cap, rp_cap);
  int stages = stages_env > 0 ? stages_env : 2;      // measured: more resident CTAs beat deeper rings
  if (stages > kMaxStages) stages = kMaxStages;
  while (stages > 2 && (size_t)stages * sb > 100 * 1024) --stages;
  const size_t smem = (size_t)stages * sb;
  if (smem > 220 * 1024) { set_error("staged spmv: tile does not fit shared memory"); return SFEM_ERR_ARG; }
  int per_sm = (int)((220 * 1024) / (smem + 2048));
  const int thread_limit = 2048 / (CT + kProducerThreads);
  if (per_sm > thread_limit) per_sm = thread_limit;
  if (per_sm < 1) per_sm = 1;
  if (ctas_env > 0) per_sm = ctas_env;
  int grid = num_sms() * per_sm;
  if (grid > P.ntiles) grid = P.ntiles;
  if (REDUCE && grid > kMaxPartials) grid = kMaxPartials;
  auto kern = k_staged<CT, NB, Epi, REDUCE>;
  static thread_local size_t configured = 0;     // per instantiation (function-local static)
  if (configured < smem) {
    SFEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(220 * 1024)));
    configured = 220 * 1024;
  }
  kern<<<grid, kConsumerThreads + kProducerThreads, smem, st>>>(P.ntiles, P.tile_row, stages, P.cap, rp_cap, A.rowptr,
                                                                 A.cols, A.vals, x, epi, partial);
  SFEM_LAUNCH_CHECK();
  if (nparts) *nparts = grid;
  return SFEM_OK;
}

template <int NB, class Epi, bool REDUCE>
int launch_staged(const Plan& P, const Csr& A, const double* x, const Epi& epi, double* partial, int* nparts,
                  cudaStream_t st) {
  if (P.max_rows <= 64) return launch_staged_ct<128, NB, Epi, REDUCE>(P, A, x, epi, partial, nparts, st);
  return launch_staged_ct<256, NB, Epi, REDUCE>(P, A, x, epi, partial, nparts, st);
}

bool find_plan(const Csr& A, Plan* out, int nb) {
  static const int disabled = env_int("SFEM_NO_STAGED", 0);
  static const int nb2 = env_int("SFEM_STAGED_NB2", 0);         // measured: with two right-hand sides the vector engine wins
  static const long long min_nnz = env_int("SFEM_STAGED_MIN_NNZ", 8000000);   // below ~100 MB the matrix is L2-resident
  if (disabled || (nb == 2 && !nb2 && g_min_tiles.load(std::memory_order_relaxed) <= 0)) return false;
  if (g_min_tiles.load(std::memory_order_relaxed) == -2) g_min_tiles.store(env_int("SFEM_STAGED_MIN_TILES", 0));
  std::lock_guard<std::mutex> lk(g_plan_mu);
  auto it = g_plans.find(A.rowptr);
  if (it == g_plans.end() || it->second.nrows != A.nrows) return false;
  // small / L2-resident matrices: the vector engine wins (lower start-up latency).  An explicit
  // sfem_staged_set_min_tiles(n > 0) replaces both thresholds (tests force the staged engine with n = 1).
  const int min_tiles = g_min_tiles.load(std::memory_order_relaxed);
  if (min_tiles > 0) {
    if (it->second.ntiles < min_tiles) return false;
  } else if (it->second.ntiles < num_sms() || A.nnz < min_nnz) {
    return false;
  }
  if ((reinterpret_cast<uintptr_t>(A.rowptr) | reinterpret_cast<uintptr_t>(A.cols) | reinterpret_cast<uintptr_t>(A.vals)) & 15u)
    return false;
  *out = it->second;
  return true;
}

}  // namespace

// ---- engine entry points used by sfem_spmv.cu: return 1 when the staged engine took the launch
int staged_spmv(const Csr& A, const double* x, const double* b, double* y, int mode, int nb, cudaStream_t st) {
  Plan P;
  if (!find_plan(A, &P, nb)) return 0;
  Prof prof(PC_SPMV_STAGED, 12.0 * A.nnz + 4.0 * A.nrows + 8.0 * nb * ((double)A.ncols + (double)A.nrows * (mode == 0 ? 1 : 2)), st);
  int rc;
  if (nb == 1) {
    if (mode == 0) rc = launch_staged<1, EpiStore<1, 0>, false>(P, A, x, EpiStore<1, 0>{b, y}, nullptr, nullptr, st);
    else if (mode == 1) rc = launch_staged<1, EpiStore<1, 1>, false>(P, A, x, EpiStore<1, 1>{b, y}, nullptr, nullptr, st);
    else rc = launch_staged<1, EpiStore<1, 2>, false>(P, A, x, EpiStore<1, 2>{b, y}, nullptr, nullptr, st);
  } else {
    if (mode == 0) rc = launch_staged<2, EpiStore<2, 0>, false>(P, A, x, EpiStore<2, 0>{b, y}, nullptr, nullptr, st);
    else if (mode == 1) rc = launch_staged<2, EpiStore<2, 1>, false>(P, A, x, EpiStore<2, 1>{b, y}, nullptr, nullptr, st);
    else rc = launch_staged<2, EpiStore<2, 2>, false>(P, A, x, EpiStore<2, 2>{b, y}, nullptr, nullptr, st);
  }
  return rc == SFEM_OK ? 1 : rc;
}

int staged_spmv_dot(const Csr& A, const double* x, const double* dx, double* y, double* partial, int* nparts, int mode,
                    int nb, cudaStream_t st) {
  Plan P;
  if (!find_plan(A, &P, nb)) return 0;
  Prof prof(PC_SPMV_DOT, 12.0 * A.nnz + 4.0 * A.nrows + 8.0 * nb * ((double)A.ncols + 2.0 * A.nrows), st);
  int rc;
  if (nb == 1) {
    if (mode == 0) rc = launch_staged<1, EpiDot<1, 0>, true>(P, A, x, EpiDot<1, 0>{dx, y, 0.0}, partial, nparts, st);
    else rc = launch_staged<1, EpiDot<1, 2>, true>(P, A, x, EpiDot<1, 2>{dx, y, 0.0}, partial, nparts, st);
  } else {
    if (mode == 0) rc = launch_staged<2, EpiDot<2, 0>, true>(P, A, x, EpiDot<2, 0>{dx, y, 0.0}, partial, nparts, st);
    else rc = launch_staged<2, EpiDot<2, 2>, true>(P, A, x, EpiDot<2, 2>{dx, y, 0.0}, partial, nparts, st);
  }
  return rc == SFEM_OK ? 1 : rc;
}

int staged_cheb_step(const Csr& A, const double* dinv, const double* d_old, double* d_new, double* r, double* x,
                     const double* c12, int last, int nb, cudaStream_t st, const double* b0) {
  Plan P;
  if (!find_plan(A, &P, nb)) return 0;
  Prof prof(PC_CHEB, 12.0 * A.nnz + 12.0 * A.nrows + 48.0 * nb * A.nrows, st);
  int rc;
  if (nb == 1) rc = launch_staged<1, EpiChebPtr<1>, false>(P, A, d_old, EpiChebPtr<1>{dinv, d_old, d_new, r, x, c12, last, b0}, nullptr, nullptr, st);
  else rc = launch_staged<2, EpiChebPtr<2>, false>(P, A, d_old, EpiChebPtr<2>{dinv, d_old, d_new, r, x, c12, last, b0}, nullptr, nullptr, st);
  return rc == SFEM_OK ? 1 : rc;
}

int staged_resid_d0(const Csr& A, const double* dinv, const double* b, const double* x, double* r, double* d,
                    const double* c0, int nb, cudaStream_t st) {
  Plan P;
  if (!find_plan(A, &P, nb)) return 0;
  Prof prof(PC_RESID_D0, 12.0 * A.nnz + 12.0 * A.nrows + 32.0 * nb * A.nrows, st);
  int rc;
  if (nb == 1) rc = launch_staged<1, EpiResidD0Ptr<1>, false>(P, A, x, EpiResidD0Ptr<1>{dinv, b, r, d, c0}, nullptr, nullptr, st);
  else rc = launch_staged<2, EpiResidD0Ptr<2>, false>(P, A, x, EpiResidD0Ptr<2>{dinv, b, r, d, c0}, nullptr, nullptr, st);
  return rc == SFEM_OK ? 1 : rc;
}

}  // namespace sfem

using namespace sfem;

extern "C" {

// Host helper: greedy tile plan.  h_rowptr [nrows+1] (host); h_tile_row out, capacity nrows+1.
// A tile takes consecutive rows while it holds at most cap_nnz entries and max_rows (64 or 128) rows.  Returns the
// number of tiles, or -1 when a single row exceeds cap_nnz (use the vector engine for such matrices).
int sfem_staged_plan(int nrows, const int* h_rowptr, int cap_nnz, int max_rows, int* h_tile_row) {
  if (nrows < 0 || cap_nnz < 4 || (max_rows != 64 && max_rows != 128)) return -1;
  int nt = 0, r = 0;
  h_tile_row[0] = 0;
  while (r < nrows) {
    int r1 = r;
    const long long k0 = h_rowptr[r];
    while (r1 < nrows && r1 - r < max_rows && (long long)h_rowptr[r1 + 1] - k0 <= cap_nnz) ++r1;
    if (r1 == r) return -1;
    h_tile_row[++nt] = r1;
    r = r1;
  }
  return nt;
}

// Attach / detach a tile plan to the matrix whose rowptr lives at device address `rowptr`.
// tile_row: DEVICE array [ntiles+1] owned by the caller and kept alive until unregister.
int sfem_staged_register(const int* rowptr, int nrows, const int* tile_row, int ntiles, int cap_nnz, int max_rows) {
  if (!rowptr || !tile_row || ntiles < 0 || cap_nnz < 4 || (max_rows != 64 && max_rows != 128)) {
    set_error("staged register: bad arguments");
    return SFEM_ERR_ARG;
  }
  Plan P;
  P.tile_row = tile_row; P.ntiles = ntiles; P.nrows = nrows; P.max_rows = max_rows;
  P.cap = ((cap_nnz + 3) & ~3) + 8;          // slack for rounding the span to 16 bytes at both ends
  std::lock_guard<std::mutex> lk(g_plan_mu);
  g_plans[rowptr] = P;
  graph_epoch_bump();
  return SFEM_OK;
}

/* matrices with fewer tiles than this use the vector engine (<= 0: default = SM count); returns the old value */
int sfem_staged_set_min_tiles(int min_tiles) {
  int old = g_min_tiles.exchange(min_tiles < 0 ? 0 : min_tiles);
  graph_epoch_bump();
  return old == -2 ? 0 : old;
}

void sfem_staged_unregister(const int* rowptr) {
  std::lock_guard<std::mutex> lk(g_plan_mu);
  g_plans.erase(rowptr);
  graph_epoch_bump();
}

}  // extern "C"
