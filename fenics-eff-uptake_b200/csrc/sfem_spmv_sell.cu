// Sliced-ELL (SELL-32-sigma) mirror of a CSR matrix and the SpMV family on it.
//
// Why a third engine: with LANES lanes per row (sfem_spmv.cu) every warp-level load of vals / cols touches
// ~10 distinct 128-byte lines and the reduction costs shuffles per row; the TMA-staged engine
// (sfem_spmv_staged.cu) fixes the global side but is bound by the shared-memory pipe.  Finite-element rows are
// short and nearly uniform (P2: 19 entries on vertex rows, 9 on edge rows), which is the textbook case for a
// sliced ELLPACK layout:
//   * rows are sorted by length inside windows of sigma rows (stable, so neighbours stay neighbours) and cut
//     into slices of 32 rows; a slice stores its entries column-major, padded to its longest row:
//         entry j of the row held by lane l of slice s  ->  position  slice_ptr[s] + 32 j + l
//   * ONE lane owns one row: a warp-level load of vals is one contiguous 256-byte span, of cols one 128-byte
//     span -- 3 wavefronts instead of ~15, no shuffles, every lane active in the epilogue;
//   * every warp owns a CONTIGUOUS run of slices, i.e. one contiguous span of the mirror (balanced by entries on
//     the host), and walks it in chunks of U column-steps (a step = 32 consecutive entries) regardless of where
//     the slice boundaries fall: every chunk is full, so no memory round trip is spent on the short tail of a
//     9- or 19-entry row; the (cols, vals) of chunk c + 1 are requested before the gathers of chunk c are
//     consumed.  A slice ends where step g + 1 == slice_ptr[s + 1] / 32; the fused epilogue of its 32 rows runs
//     right there, its operands having been requested when the slice was opened.  All control flow is
//     warp-uniform.
// Measured on B200 (profiles/r01_spmv_microbench.md): 5.2-5.6 TB/s = 0.81-0.86 of the measured copy bandwidth on
// P2 matrices larger than L2, against 4.1-4.4 (lane groups) and 4.0-4.7 (TMA-staged).  Two earlier versions are
// recorded there as negative results: warp-strided slices with per-slice passes (tail passes cost a full round
// trip: 3.9-5.1 TB/s depending on how U divides the row lengths) and a per-warp TMA ring in shared memory
// (4.4-5.1 TB/s: the deeper prefetch does not pay for the shared-memory round trip and the smaller L1).
// The mirror holds its own value array; it is refreshed from the CSR values by k_sell_pack.  Every library entry
// that writes CSR values (sfem_gather_csr, sfem_apply_dirichlet, sfem_csr_extract) marks the mirror of that value
// array dirty; a dirty mirror is re-packed at the next solver entry / un-captured SpMV and is never read stale
// (inside a stream capture a dirty mirror makes the launch fall back to the CSR engines).
//
// The fused epilogues are the shared ones of sfem_spmv_epi.cuh, so all engines compute the same updates; the
// per-row summation order is the column order of the row split over two accumulators by chunk position (fixed for
// a given matrix and device, so results are bit-reproducible run to run).
#include "sfem_common.cuh"
#include "sfem_internal.h"
#include "sfem_row_engine.cuh"
#include "sfem_spmv_epi.cuh"

#include <cstdlib>
#include <mutex>
#include <thread>
#include <unordered_map>

namespace sfem {

namespace {

constexpr int kSellThreads = 256;
constexpr int kSellBlocksPerSm = 4;

// The warp ranges are balanced by steps on the host: parts[k] = first slice of fine part k (kSellFine per
// SM-resident warp slot), and a kernel whose occupancy is MINB blocks per SM merges 12 / MINB fine parts per warp.
constexpr int kSellFine = 12;      // fine parts per (SM x warp slot of a 256-thread block); divisible by 2, 3, 4, 6

// epilogue state of the row a lane currently owns; with two right-hand sides both are handled as one 16-byte access
template <int NB, class Epi>
struct RowEpi {
  typename Epi::Pre p[NB];
  __device__ __forceinline__ void open(const Epi& e, int row) {
#pragma unroll
    for (int c = 0; c < NB; ++c) p[c] = e.pre(row, c, row >= 0);
  }
  __device__ __forceinline__ void close(Epi& e, int row, const Acc<NB>& a0, const Acc<NB>& a1) const {
#pragma unroll
    for (int c = 0; c < NB; ++c) e.fin(row, c, a0.v[c] + a1.v[c], p[c]);
  }
};
template <class Epi>
struct RowEpi<2, Epi> {
  typename Epi::Pre2 p;
  __device__ __forceinline__ void open(const Epi& e, int row) { p = e.pre2(row, row >= 0); }
  __device__ __forceinline__ void close(Epi& e, int row, const Acc<2>& a0, const Acc<2>& a1) const {
    e.fin2(row, a0.v[0] + a1.v[0], a0.v[1] + a1.v[1], p);
  }
};

template <int NB, class Epi, bool REDUCE, int U, int MINB>
__global__ void __launch_bounds__(kSellThreads, MINB)
    k_sell_stream(int nwarps, const int* __restrict__ parts, const int* __restrict__ slice_ptr,
                  const int* __restrict__ perm, const int* __restrict__ scols, const double* __restrict__ svals,
                  const double* __restrict__ x, Epi epi, double* __restrict__ partial) {
  __shared__ double sh[33];
  const int lane = threadIdx.x & 31;
  constexpr int kWarps = kSellThreads / 32;
  constexpr int kPer = kSellFine / MINB;
  const int gw = blockIdx.x * kWarps + (threadIdx.x >> 5);
  int s = 0, s_end = 0;
  if (gw < nwarps) { s = parts[gw * kPer]; s_end = parts[(gw + 1) * kPer]; }
  if (s < s_end) {
    int g = slice_ptr[s] >> 5;
    const int g_end = slice_ptr[s_end] >> 5;
    int e = slice_ptr[s + 1] >> 5;
    int row = perm[(size_t)s * 32 + lane];
    int ne = e, nrow = -1;
    if (s + 1 < s_end) { ne = slice_ptr[s + 2] >> 5; nrow = perm[(size_t)(s + 1) * 32 + lane]; }
    RowEpi<NB, Epi> re;
    Acc<NB> a0, a1;
    re.open(epi, row);
#pragma unroll
    for (int c = 0; c < NB; ++c) { a0.v[c] = 0.0; a1.v[c] = 0.0; }
    // close the current slice and open the next one (also swallows slices without any entry)
    auto advance = [&](int gnext) {
      for (;;) {
        if (row >= 0) re.close(epi, row, a0, a1);
        ++s; row = nrow; e = ne;
#pragma unroll
        for (int c = 0; c < NB; ++c) { a0.v[c] = 0.0; a1.v[c] = 0.0; }
        if (s >= s_end) break;
        re.open(epi, row);
        if (s + 1 < s_end) { ne = slice_ptr[s + 2] >> 5; nrow = perm[(size_t)(s + 1) * 32 + lane]; }
        if (e != gnext) break;           // the usual case: the new slice has entries
      }
    };
    while (s < s_end && e == g) advance(g);          // leading slices without entries
    int cc[U];
    double vv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = g + u < g_end;
      const long long k = (long long)(g + u) * 32 + lane;
      cc[u] = ok ? __ldcs(scols + k) : -1;
      vv[u] = ok ? __ldcs(svals + k) : 0.0;
    }
    while (g < g_end) {
      XVal<NB, double> xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) xv[u].load(x, cc[u], cc[u] >= 0);
      int ncc[U];
      double nvv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool ok = g + U + u < g_end;
        const long long k = (long long)(g + U + u) * 32 + lane;
        ncc[u] = ok ? __ldcs(scols + k) : -1;
        nvv[u] = ok ? __ldcs(svals + k) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (g + u < g_end) {
          xv[u].fma_into(vv[u], (u & 1) ? a1 : a0);
          if (g + u + 1 == e) advance(g + u + 1);
        }
      }
      g += U;
#pragma unroll
      for (int u = 0; u < U; ++u) { cc[u] = ncc[u]; vv[u] = nvv[u]; }
    }
    while (s < s_end) advance(g_end);                 // trailing slices without entries
  }
  if (REDUCE) {
    const double t = block_sum(epi_acc_of(epi), sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(kThreads) k_sell_pack(long long padded, const int* __restrict__ src,
                                                        const double* __restrict__ csr_vals, double* __restrict__ svals) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += stride) {
    const int k = __ldcs(src + i);
    svals[i] = k >= 0 ? __ldg(csr_vals + k) : 0.0;
  }
}

// ------------------------------------------------------------------ registry
struct SellPlan {
  int nrows = 0;
  int nslices = 0;
  long long padded = 0;
  const int* slice_ptr = nullptr;    // device [nslices + 1], multiples of 32
  const int* perm = nullptr;         // device [nslices * 32], row of every lane (-1: none)
  const int* parts = nullptr;        // device [nparts + 1], first slice of every fine part (balanced by entries)
  int nparts = 0;                    // = kSellFine x SMs x warps per block
  const int* scols = nullptr;        // device [padded], -1 on padding
  const int* src = nullptr;          // device [padded], CSR slot of every position (-1 on padding)
  double* svals = nullptr;           // device [padded]
  const double* csr_vals = nullptr;  // the CSR value array this mirror follows
  bool dirty = true;
  std::thread::id dirty_by;          // thread whose entry wrote the CSR values last (it owns the stream they were written on)
};
std::mutex g_mu;
std::unordered_map<const int*, SellPlan> g_plans;              // by rowptr address
std::unordered_map<const double*, const int*> g_by_vals;       // CSR value array -> rowptr key
std::atomic<int> g_min_rows{-2};                               // -2: read SFEM_SELL_MIN_ROWS on first use
std::atomic<int> g_ndirty{0};

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

int min_rows() {
  int m = g_min_rows.load(std::memory_order_relaxed);
  if (m == -2) {
    m = env_int("SFEM_SELL_MIN_ROWS", 0);
    if (m < 0) m = 0;
    g_min_rows.store(m);
  }
  return m > 0 ? m : 250000;      // default: smaller (L2-resident, latency-bound) levels are faster with the lane-group engine (measured)
}

int pack_locked(SellPlan& P, cudaStream_t st) {
  if (P.padded > 0) {
    Prof prof(PC_GATHER, 20.0 * (double)P.padded, st);
    k_sell_pack<<<grid_for(P.padded, kThreads * 4, 8), kThreads, 0, st>>>(P.padded, P.src, P.csr_vals, P.svals);
    SFEM_LAUNCH_CHECK();
  }
  if (P.dirty) { P.dirty = false; g_ndirty.fetch_sub(1); }
  return SFEM_OK;
}

// returns 1 and fills *out when the mirror of A can serve the launch (packing it first if needed), 0 otherwise
int find_sell(const Csr& A, SellPlan* out, cudaStream_t st) {
  static const int disabled = env_int("SFEM_NO_SELL", 0);
  if (disabled) return 0;
  if (A.nrows < min_rows()) return 0;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_plans.find(A.rowptr);
  if (it == g_plans.end()) return 0;
  SellPlan& P = it->second;
  if (P.nrows != A.nrows || P.csr_vals != A.vals) return 0;
  if (P.dirty) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return 0;
    const int rc = pack_locked(P, st);
    if (rc != SFEM_OK) return rc;
  }
  *out = P;
  return 1;
}

template <int NB, class Epi, bool REDUCE, int U = (NB == 1 ? 4 : 3), int MINB = kSellBlocksPerSm>
int launch_sell(const SellPlan& P, const double* x, const Epi& epi, double* partial, int* nparts, cudaStream_t st) {
  constexpr int kWarps = kSellThreads / 32;
  static_assert(kSellFine % MINB == 0, "fine parts must merge evenly");
  const int nwarps = P.nparts / (kSellFine / MINB);
  int grid = (nwarps + kWarps - 1) / kWarps;
  if (grid < 1) grid = 1;
  k_sell_stream<NB, Epi, REDUCE, U, MINB><<<grid, kSellThreads, 0, st>>>(nwarps, P.parts, P.slice_ptr, P.perm, P.scols,
                                                                        P.svals, x, epi, partial);
  SFEM_LAUNCH_CHECK();
  if (nparts) *nparts = grid;
  return SFEM_OK;
}

inline double spmv_bytes(const Csr& A, int nb, int vec_passes) {
  return 12.0 * A.nnz + 4.0 * A.nrows + 8.0 * nb * ((double)A.ncols + (double)A.nrows * vec_passes);
}

}  // namespace

// ---- engine entry points used by sfem_spmv.cu: 1 = took the launch, 0 = not applicable, < 0 = error
int sell_spmv(const Csr& A, const double* x, const double* b, double* y, int mode, int nb, cudaStream_t st) {
  if (nb == 2 && !aligned16(x, b, y)) return 0;       // the two-RHS epilogues use 16-byte accesses
  SellPlan P;
  const int f = find_sell(A, &P, st);
  if (f <= 0) return f;
  Prof prof(PC_SPMV, spmv_bytes(A, nb, mode == 0 ? 1 : 2), st);
  int rc;
  if (nb == 1) {
    if (mode == 0) rc = launch_sell<1, EpiStore<1, 0>, false>(P, x, EpiStore<1, 0>{b, y}, nullptr, nullptr, st);
    else if (mode == 1) rc = launch_sell<1, EpiStore<1, 1>, false>(P, x, EpiStore<1, 1>{b, y}, nullptr, nullptr, st);
    else rc = launch_sell<1, EpiStore<1, 2>, false>(P, x, EpiStore<1, 2>{b, y}, nullptr, nullptr, st);
  } else {
    if (mode == 0) rc = launch_sell<2, EpiStore<2, 0>, false>(P, x, EpiStore<2, 0>{b, y}, nullptr, nullptr, st);
    else if (mode == 1) rc = launch_sell<2, EpiStore<2, 1>, false>(P, x, EpiStore<2, 1>{b, y}, nullptr, nullptr, st);
    else rc = launch_sell<2, EpiStore<2, 2>, false>(P, x, EpiStore<2, 2>{b, y}, nullptr, nullptr, st);
  }
  return rc == SFEM_OK ? 1 : rc;
}

int sell_spmv_dot(const Csr& A, const double* x, const double* dx, double* y, double* partial, int* nparts, int mode,
                  int nb, cudaStream_t st) {
  if (nb == 2 && !aligned16(x, dx, y)) return 0;
  SellPlan P;
  const int f = find_sell(A, &P, st);
  if (f <= 0) return f;
  Prof prof(PC_SPMV_DOT, spmv_bytes(A, nb, 2), st);
  int rc;
  if (nb == 1) {
    if (mode == 0) rc = launch_sell<1, EpiDot<1, 0>, true>(P, x, EpiDot<1, 0>{dx, y, 0.0}, partial, nparts, st);
    else rc = launch_sell<1, EpiDot<1, 2>, true>(P, x, EpiDot<1, 2>{dx, y, 0.0}, partial, nparts, st);
  } else {
    if (mode == 0) rc = launch_sell<2, EpiDot<2, 0>, true>(P, x, EpiDot<2, 0>{dx, y, 0.0}, partial, nparts, st);
    else rc = launch_sell<2, EpiDot<2, 2>, true>(P, x, EpiDot<2, 2>{dx, y, 0.0}, partial, nparts, st);
  }
  return rc == SFEM_OK ? 1 : rc;
}

int sell_cheb_step(const Csr& A, const double* dinv, const double* d_old, double* d_new, double* r, double* x,
                   const double* c12, int last, int nb, cudaStream_t st, const double* b0) {
  if (nb == 2 && (!aligned16(d_old, d_new, r, x) || (b0 && !aligned16(b0)))) return 0;
  SellPlan P;
  const int f = find_sell(A, &P, st);
  if (f <= 0) return f;
  Prof prof(PC_CHEB, 12.0 * A.nnz + 12.0 * A.nrows + (b0 ? 40.0 : 48.0) * nb * A.nrows, st);
  int rc;
  if (nb == 1) rc = launch_sell<1, EpiChebPtr<1>, false>(P, d_old, EpiChebPtr<1>{dinv, d_old, d_new, r, x, c12, last, b0}, nullptr, nullptr, st);
  // heaviest epilogue (7 vector streams x 2 right-hand sides): 4-step chunks at 3 blocks per SM beat 3-step chunks at 4
  // blocks (spills) and 2-step chunks at 4 blocks -- V-cycle at r = 2: 0.667 / 0.713 / 0.675 ms
  else rc = launch_sell<2, EpiChebPtr<2>, false, 4, 3>(P, d_old, EpiChebPtr<2>{dinv, d_old, d_new, r, x, c12, last, b0}, nullptr, nullptr, st);
  return rc == SFEM_OK ? 1 : rc;
}

int sell_resid_d0(const Csr& A, const double* dinv, const double* b, const double* x, double* r, double* d,
                  const double* c0, int nb, cudaStream_t st) {
  if (nb == 2 && !aligned16(b, x, r, d)) return 0;
  SellPlan P;
  const int f = find_sell(A, &P, st);
  if (f <= 0) return f;
  Prof prof(PC_RESID_D0, 12.0 * A.nnz + 12.0 * A.nrows + 32.0 * nb * A.nrows, st);
  int rc;
  if (nb == 1) rc = launch_sell<1, EpiResidD0Ptr<1>, false>(P, x, EpiResidD0Ptr<1>{dinv, b, r, d, c0}, nullptr, nullptr, st);
  else rc = launch_sell<2, EpiResidD0Ptr<2>, false>(P, x, EpiResidD0Ptr<2>{dinv, b, r, d, c0}, nullptr, nullptr, st);
  return rc == SFEM_OK ? 1 : rc;
}

void sell_mark_dirty(const double* csr_vals) {
  if (csr_vals == nullptr) return;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_by_vals.find(csr_vals);
  if (it == g_by_vals.end()) return;
  auto pit = g_plans.find(it->second);
  if (pit == g_plans.end()) return;
  if (!pit->second.dirty) { pit->second.dirty = true; g_ndirty.fetch_add(1); }
  pit->second.dirty_by = std::this_thread::get_id();
}

int sell_ensure_all(cudaStream_t st) {
  if (g_ndirty.load(std::memory_order_relaxed) <= 0) return SFEM_OK;
  // Only mirrors the engine would serve (>= min_rows; smaller ones stay dirty and cost nothing) and only those whose
  // values were written by THIS thread: with several solver threads (sweep.run_concurrent, one stream each) another
  // thread's matrix may be half-assembled on a stream this one is not ordered with.
  const int mr = min_rows();
  const std::thread::id me = std::this_thread::get_id();
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& kv : g_plans)
    if (kv.second.dirty && kv.second.nrows >= mr && kv.second.dirty_by == me) SFEM_TRY(pack_locked(kv.second, st));
  return SFEM_OK;
}

}  // namespace sfem

using namespace sfem;

extern "C" {

/* number of fine parts the partition array of a mirror must have on this device */
int sfem_sell_parts(void) { return kSellFine * num_sms() * (kSellThreads / 32); }

int sfem_sell_register(const int* rowptr, const double* csr_vals, int nrows, int nslices, const int* slice_ptr,
                       const int* perm, const int* scols, const int* src, double* svals, long long padded,
                       const int* parts, int nparts) {
  if (!rowptr || !csr_vals || nrows < 0 || nslices < 0 || padded < 0 || (nslices > 0 && (!slice_ptr || !perm)) ||
      (padded > 0 && (!scols || !src || !svals)) || (long long)nslices * 32 < nrows || !parts ||
      nparts != sfem_sell_parts()) {
    set_error("sell register: bad arguments");
    return SFEM_ERR_ARG;
  }
  SellPlan P;
  P.nrows = nrows; P.nslices = nslices; P.padded = padded;
  P.parts = parts; P.nparts = nparts;
  P.slice_ptr = slice_ptr; P.perm = perm; P.scols = scols; P.src = src; P.svals = svals; P.csr_vals = csr_vals;
  P.dirty = true;
  P.dirty_by = std::this_thread::get_id();
  std::lock_guard<std::mutex> lk(g_mu);
  auto old = g_plans.find(rowptr);
  if (old != g_plans.end()) {
    if (old->second.dirty) g_ndirty.fetch_sub(1);
    g_by_vals.erase(old->second.csr_vals);
  }
  g_plans[rowptr] = P;
  g_by_vals[csr_vals] = rowptr;
  g_ndirty.fetch_add(1);
  graph_epoch_bump();                       // captured launch sequences baked in the engine choice for this address
  return SFEM_OK;
}

void sfem_sell_unregister(const int* rowptr) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_plans.find(rowptr);
  if (it == g_plans.end()) return;
  if (it->second.dirty) g_ndirty.fetch_sub(1);
  g_by_vals.erase(it->second.csr_vals);
  g_plans.erase(it);
  graph_epoch_bump();
}

int sfem_sell_mark_dirty(const int* rowptr) {
  const double* v = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_plans.find(rowptr);
    if (it == g_plans.end()) return SFEM_OK;
    v = it->second.csr_vals;
  }
  sell_mark_dirty(v);
  return SFEM_OK;
}

int sfem_sell_sync(void* stream) { return sell_ensure_all((cudaStream_t)stream); }

int sfem_sell_set_min_rows(int n) {
  const int old = g_min_rows.exchange(n < 0 ? 0 : n);
  return old == -2 ? 0 : old;
}

int sfem_spmv_csr_f64_sell(int nrows, int ncols, int nnz, const int* rowptr, const int* cols, const double* vals,
                           const double* x, const double* b, double* y, int mode, int nb, void* stream) {
  if (nrows < 0 || nnz < 0 || mode < 0 || mode > 2 || (nb != 1 && nb != 2)) { set_error("sell spmv: bad arguments"); return SFEM_ERR_ARG; }
  if (mode == 1 && b == nullptr) { set_error("sell spmv: mode 1 needs b"); return SFEM_ERR_ARG; }
  if (nb == 2 && (reinterpret_cast<uintptr_t>(x) & 15u)) { set_error("sell spmv: nb = 2 needs a 16-byte aligned x"); return SFEM_ERR_ARG; }
  Csr A;
  A.nrows = nrows; A.ncols = ncols; A.nnz = nnz;
  A.rowptr = rowptr; A.cols = cols; A.vals = vals;
  const int r = sell_spmv(A, x, b, y, mode, nb, (cudaStream_t)stream);
  if (r == 0) { set_error("sell spmv: no sliced-ELL mirror registered for this matrix (or fewer rows than the threshold)"); return SFEM_ERR_ARG; }
  return r < 0 ? r : SFEM_OK;
}

}  // extern "C"
