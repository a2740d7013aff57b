// Multi-GPU plumbing (internal): peer-memory mailboxes over NVLink / NVSwitch.
//
// One process per GPU.  Every rank owns one device "mailbox" (cudaMalloc), exports it with CUDA IPC
// and maps the mailboxes of all peers; after that the data path never calls a library collective:
//   * halo exchange  = ONE kernel per exchange: block b packs the boundary values neighbour b needs and
//     stores them straight into that neighbour's mailbox (remote 16-byte stores over NVLink) as FLAGGED
//     words -- every double carries the sequence number of the exchange (see ll_store) -- then polls its
//     own mailbox until the neighbour's words of this exchange have arrived and copies them into the ghost
//     region of the vector: no fence, no separate flag, latency = one NVLink flight;
//   * all-reduce of the few Krylov scalars = inside the one-block scalar kernels: every rank stores its
//     partial sums (flagged words) into every peer's mailbox, polls for the nranks contributions and adds
//     them in rank order -> bit-identical on all ranks and run to run;
//   * vector all-reduce (restriction onto the replicated coarse hierarchy) = same scheme, chunked.
// Slots are double buffered by sequence parity; exchanges are symmetric (every pair sends both ways
// each time), so a slot is only rewritten after its reader has moved on.  Spins carry a time-out that
// raises the handle's error flag instead of hanging the GPU.
#pragma once
#include "sfem_common.cuh"

namespace sfem {

constexpr int kMaxRanks = 16;
constexpr int kAllreduceMaxK = 16;   // >= kMaxRanks: the Gershgorin all-gather of sfem_mg.cu publishes one slot per rank

// device-visible description of the communicator (passed by value to kernels)
struct DistDev {
  int rank = 0, nranks = 1;
  double* mailbox[kMaxRanks] = {};       // base of every rank's mailbox (peer-mapped; [rank] = local)
  unsigned long long* seq = nullptr;     // local sequence counters [0]=scalar all-reduce, [1]=vector all-reduce
  int* err = nullptr;                    // local error flag (time-out)
  // fixed layout at the start of every mailbox (in 8-byte words)
  long long sc_flag_off = 0;             // [2][nranks] flags of the scalar all-reduce
  long long sc_data_off = 0;             // [2][nranks][kAllreduceMaxK] flagged pairs = 2 words each
  long long vec_flag_off = 0;            // [2][nranks]
  long long vec_data_off = 0;            // [2][nranks][vec_cap]
  long long vec_cap = 0;
};

// one halo pattern (one multigrid level): device arrays of length nneigh
struct HaloDev {
  int nneigh = 0;
  int n_own = 0, n_loc = 0;
  const int* peer = nullptr;             // neighbour rank
  const int* send_cnt = nullptr;         // dofs sent to the neighbour
  const int* send_ptr = nullptr;         // offsets into send_idx, [nneigh+1]
  const int* send_idx = nullptr;         // local owned indices, in the order the receiver's ghost list expects
  const int* recv_cnt = nullptr;
  const int* recv_off = nullptr;         // first ghost slot (local index) filled by this neighbour
  const long long* peer_data_off = nullptr;   // where I write in the neighbour's mailbox: [2 parities] x cap words
  const long long* peer_flag_off = nullptr;   // [2] flags there
  const long long* my_data_off = nullptr;     // where the neighbour writes in my mailbox
  const long long* my_flag_off = nullptr;
  const long long* cap = nullptr;             // words per parity of each channel (>= 2 words per double sent: flagged pairs)
  unsigned long long* seq = nullptr;          // local, one counter per neighbour channel
};

struct Halo {
  HaloDev dev;
  int max_cnt = 0;
};

struct Dist {
  DistDev dev;
  size_t mailbox_words = 0;
  void* peer_base[kMaxRanks] = {};
  bool peer_open[kMaxRanks] = {};
};

// registry: halo pattern attached to a matrix (keyed by the device address of its rowptr)
const Halo* find_halo(const int* rowptr);
Dist* active_dist();                          // communicator of this process (nullptr: single GPU)
DistDev dist_dev();                           // its device view (nranks = 1 when none is active)

// fills the ghost region of x (n_loc entries x nb) from the owners; no-op when h == nullptr
int halo_exchange(const Halo* h, double* x, int nb, cudaStream_t st, int phase = 0);
// both patterns in ONE launch (one rendezvous with the neighbours): the Taylor-Hood apply's velocity + pressure ghosts
int halo_exchange_pair(const Halo* h1, double* x1, int nb1, const Halo* h2, double* x2, int nb2, cudaStream_t st);
// x[0..n) <- sum over ranks (in rank order); n <= vec_cap
int dist_allreduce_vec(Dist* d, double* x, int n, cudaStream_t st, int phase = 0);

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// spin until *flag >= want; returns false on time-out (10 s)
__device__ __forceinline__ bool spin_wait(const unsigned long long* flag, unsigned long long want, int* err) {
  const unsigned long long t0 = global_timer_ns();
  while (ld_acquire_sys(flag) < want) {
    if (global_timer_ns() - t0 > 10000000000ull) {
      if (err) *err = 1;
      return false;
    }
  }
  return true;
}

// ---- flagged ("LL") words: a double travels as two 8-byte words {flag : 32 | half of the double : 32}, written with ONE
// 16-byte store into the peer's mailbox.  An 8-byte word arrives atomically, so the receiver needs no fence and no separate
// flag: it polls the slot until both words carry the sequence number of this exchange.  This removes the two system-scope
// fence round trips of a data + fence + flag protocol from every exchange (the payloads here are a few kB: latency is all
// that matters), at twice the bytes on the wire.
__device__ __forceinline__ void ll_store(unsigned long long* dst2, double v, unsigned int flag) {
  const unsigned long long lo = ((unsigned long long)flag << 32) | (unsigned int)__double2loint(v);
  const unsigned long long hi = ((unsigned long long)flag << 32) | (unsigned int)__double2hiint(v);
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst2), "l"(lo), "l"(hi) : "memory");
}
__device__ __forceinline__ ulonglong2 ll_load(const unsigned long long* src2) {
  ulonglong2 w;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w.x), "=l"(w.y) : "l"(src2) : "memory");
  return w;
}
__device__ __forceinline__ bool ll_ready(const ulonglong2& w, unsigned int flag) {
  return (unsigned int)(w.x >> 32) == flag && (unsigned int)(w.y >> 32) == flag;
}
__device__ __forceinline__ double ll_value(const ulonglong2& w) {
  return __hiloint2double((int)(unsigned int)w.y, (int)(unsigned int)w.x);
}
// polls one slot until it carries `flag`; false on time-out (10 s)
__device__ __forceinline__ bool ll_wait(const unsigned long long* src2, unsigned int flag, double* out, int* err) {
  ulonglong2 w = ll_load(src2);
  if (!ll_ready(w, flag)) {
    const unsigned long long t0 = global_timer_ns();
    do {
      w = ll_load(src2);
      if (global_timer_ns() - t0 > 10000000000ull) {
        if (err) *err = 1;
        *out = 0.0;
        return false;
      }
    } while (!ll_ready(w, flag));
  }
  *out = ll_value(w);
  return true;
}

// All-reduce (sum) of K <= kAllreduceMaxK scalars, called by ONE thread of a one-block kernel.
// vals in/out.  With nranks == 1 it is a no-op.  Flagged words: every rank stores its K values into every peer's mailbox
// (posted 16-byte stores, no fence), then polls the nranks x K slots of its own mailbox -- all loads of a polling pass are
// issued before the first one is tested -- and adds the contributions in rank order (bit-identical on all ranks).
__device__ __forceinline__ void dist_allreduce_scalars(const DistDev& D, double* vals, int K) {
  if (D.nranks <= 1) return;
  const unsigned long long s = D.seq[0];
  const int par = (int)(s & 1ull);
  const unsigned int flag = (unsigned int)(s + 1);
  for (int q = 0; q < D.nranks; ++q) {
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(D.mailbox[q] + D.sc_data_off) +
                              2 * ((long long)par * D.nranks + D.rank) * kAllreduceMaxK;
    for (int k = 0; k < K; ++k) ll_store(dst + 2 * k, vals[k], flag);
  }
  const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(D.mailbox[D.rank] + D.sc_data_off) +
                                   2 * (long long)par * D.nranks * kAllreduceMaxK;
  if (K == 1) {
    const unsigned long long t0 = global_timer_ns();
    for (;;) {
      ulonglong2 w[kMaxRanks];
#pragma unroll
      for (int q = 0; q < kMaxRanks; ++q)
        if (q < D.nranks) w[q] = ll_load(mine + 2 * (long long)q * kAllreduceMaxK);
      bool ok = true;
      double acc = 0.0;
#pragma unroll
      for (int q = 0; q < kMaxRanks; ++q)
        if (q < D.nranks) { ok = ok && ll_ready(w[q], flag); acc += ll_value(w[q]); }
      if (ok) { vals[0] = acc; break; }
      if (global_timer_ns() - t0 > 10000000000ull) { if (D.err) *D.err = 1; break; }
    }
  } else {
    double acc[kAllreduceMaxK];
    for (int k = 0; k < K; ++k) acc[k] = 0.0;
    for (int q = 0; q < D.nranks; ++q)
      for (int k = 0; k < K; ++k) {
        double v;
        ll_wait(mine + 2 * ((long long)q * kAllreduceMaxK + k), flag, &v, D.err);
        acc[k] += v;
      }
    for (int k = 0; k < K; ++k) vals[k] = acc[k];
  }
  D.seq[0] = s + 1;
}
#endif

}  // namespace sfem
