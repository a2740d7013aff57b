// Multi-GPU plumbing (internal): peer-memory mailboxes over NVLink / NVSwitch.
//
// One process per GPU.  Every rank owns one device "mailbox" (cudaMalloc), exports it with CUDA IPC
// and maps the mailboxes of all peers; after that the data path never calls a library collective:
//   * halo exchange  = ONE kernel per exchange: block b packs the boundary values neighbour b needs and
//     stores them straight into that neighbour's mailbox (remote st.global over NVLink), publishes a
//     sequence flag with release semantics, then spins (acquire) on its own flag from the same neighbour
//     and copies the received values into the ghost region of the vector;
//   * all-reduce of the few Krylov scalars = inside the one-block scalar kernels: every rank stores its
//     partial sums into every peer's mailbox, waits for all flags and adds the nranks contributions in
//     rank order -> bit-identical on all ranks and run to run;
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
  long long sc_data_off = 0;             // [2][nranks][kAllreduceMaxK]
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
  const long long* cap = nullptr;             // words per parity of each channel (>= cnt * 2)
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

// All-reduce (sum) of K <= kAllreduceMaxK scalars, called by ONE thread of a one-block kernel.
// vals in/out.  With nranks == 1 it is a no-op.
__device__ __forceinline__ void dist_allreduce_scalars(const DistDev& D, double* vals, int K) {
  if (D.nranks <= 1) return;
  const unsigned long long s = D.seq[0];
  const int par = (int)(s & 1ull);
  for (int q = 0; q < D.nranks; ++q) {
    double* dst = D.mailbox[q] + D.sc_data_off + ((long long)par * D.nranks + D.rank) * kAllreduceMaxK;
    for (int k = 0; k < K; ++k) dst[k] = vals[k];
  }
  __threadfence_system();
  for (int q = 0; q < D.nranks; ++q) {
    unsigned long long* f = reinterpret_cast<unsigned long long*>(D.mailbox[q] + D.sc_flag_off) + par * D.nranks + D.rank;
    st_release_sys(f, s + 1);
  }
  double acc[kAllreduceMaxK];
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  const double* mine = D.mailbox[D.rank];
  for (int q = 0; q < D.nranks; ++q) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine + D.sc_flag_off) + par * D.nranks + q;
    spin_wait(f, s + 1, D.err);
    const volatile double* src = mine + D.sc_data_off + ((long long)par * D.nranks + q) * kAllreduceMaxK;
    for (int k = 0; k < K; ++k) acc[k] += src[k];
  }
  for (int k = 0; k < K; ++k) vals[k] = acc[k];
  D.seq[0] = s + 1;
}
#endif

}  // namespace sfem
