// Multi-GPU plumbing: peer-memory mailboxes, halo exchange and all-reduce kernels (see sfem_dist.h).
#include "sfem_dist.h"
#include "sfem_internal.h"

#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace sfem {

namespace {

std::mutex g_mu;
std::unordered_map<const int*, const Halo*> g_halos;
std::atomic<int> g_nhalos{0};
Dist* g_dist = nullptr;

// ---- halo exchange: block b <-> neighbour b, flagged words (sfem_dist.h)
// phase 0: send + wait + unpack (production); 1: send only; 2: wait + unpack only (single-GPU emulation in tests)
__device__ __forceinline__ void halo_block(const DistDev& D, const HaloDev& H, double* __restrict__ x, int b, int nb,
                                           int phase) {
  const int q = H.peer[b];
  const unsigned long long s = H.seq[b];
  const int par = (int)(s & 1ull);
  const unsigned int flag = (unsigned int)(s + 1);
  __syncthreads();                                            // everyone has read seq before thread 0 bumps it
  if (phase != 2) {                                           // pack + remote store
    const int cnt = H.send_cnt[b] * nb;
    const int* idx = H.send_idx + H.send_ptr[b];
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(D.mailbox[q] + H.peer_data_off[b] + (long long)par * H.cap[b]);
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const int li = idx[nb == 1 ? i : (i >> 1)];
      ll_store(dst + 2 * (long long)i, x[nb == 1 ? li : 2 * (long long)li + (i & 1)], flag);
    }
  }
  if (phase == 1) return;
  {                                                           // poll + unpack: mailbox -> ghost region
    const int cnt = H.recv_cnt[b] * nb;
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(D.mailbox[D.rank] + H.my_data_off[b] + (long long)par * H.cap[b]);
    double* dst = x + (size_t)H.recv_off[b] * nb;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      double v;
      ll_wait(src + 2 * (long long)i, flag, &v, D.err);
      dst[i] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) H.seq[b] = s + 1;
}

__global__ void __launch_bounds__(256) k_halo_exchange(DistDev D, HaloDev H, double* __restrict__ x, int nb, int phase) {
  halo_block(D, H, x, blockIdx.x, nb, phase);
}

// two halo patterns (Taylor-Hood apply: interleaved velocity ghosts of K + pressure ghosts of B^T, both of ONE Stokes
// vector) served by one launch
__global__ void __launch_bounds__(256) k_halo_exchange_pair(DistDev D, HaloDev H1, double* __restrict__ x1, int nb1,
                                                            HaloDev H2, double* __restrict__ x2, int nb2) {
  if ((int)blockIdx.x < H1.nneigh) halo_block(D, H1, x1, blockIdx.x, nb1, 0);
  else halo_block(D, H2, x2, blockIdx.x - H1.nneigh, nb2, 0);
}

// ---- vector all-reduce, phase 1: every block stores its chunk into all peers; the last block to
// finish publishes the flags
__global__ void __launch_bounds__(256) k_allreduce_vec_send(DistDev D, const double* __restrict__ x, int n,
                                                            unsigned int* __restrict__ ticket) {
  const unsigned long long s = D.seq[1];
  const int par = (int)(s & 1ull);
  for (int q = 0; q < D.nranks; ++q) {
    double* dst = D.mailbox[q] + D.vec_data_off + ((long long)par * D.nranks + D.rank) * D.vec_cap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = x[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) {
      *ticket = 0;
      __threadfence_system();
      for (int q = 0; q < D.nranks; ++q) {
        unsigned long long* f = reinterpret_cast<unsigned long long*>(D.mailbox[q] + D.vec_flag_off) + par * D.nranks + D.rank;
        st_release_sys(f, s + 1);
      }
    }
  }
}

// phase 2: wait for every rank's flag, add the contributions in rank order
__global__ void __launch_bounds__(256) k_allreduce_vec_recv(DistDev D, double* __restrict__ x, int n,
                                                            unsigned int* __restrict__ ticket) {
  const unsigned long long s = D.seq[1];
  const int par = (int)(s & 1ull);
  const double* mine = D.mailbox[D.rank];
  if (threadIdx.x == 0) {
    for (int q = 0; q < D.nranks; ++q) {
      const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine + D.vec_flag_off) + par * D.nranks + q;
      spin_wait(f, s + 1, D.err);
    }
  }
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (int q = 0; q < D.nranks; ++q) {
      const volatile double* src = mine + D.vec_data_off + ((long long)par * D.nranks + q) * D.vec_cap;
      acc += src[i];
    }
    x[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket + 1, 1u);
    if (t == gridDim.x - 1) {
      ticket[1] = 0;
      D.seq[1] = s + 1;
    }
  }
}

// ---- vector all-reduce in ONE launch: reduce-scatter + all-gather over flagged words.  Thread i owns element i of every
// chunk (chunk q = the part rank q reduces): it sends x[q][i] to rank q, polls the R - 1 contributions to its own chunk,
// adds them in rank order, sends the sum to every peer and polls the R - 1 sums of the other chunks.  (R - 1) / R of the
// vector travels twice -- against R - 1 full copies for the send-to-all scheme above -- and there is no fence, no ticket and
// no second kernel on the critical path; the reduced value of an element is computed by ONE rank, so all ranks hold
// bit-identical vectors.  Slots: [phase 2][parity 2][sender R][chunk] flagged pairs inside the mailbox's vector region.
__global__ void __launch_bounds__(256) k_allreduce_vec_ll(DistDev D, double* __restrict__ x, int n,
                                                          unsigned int* __restrict__ ticket) {
  const unsigned long long s = D.seq[1];
  const int par = (int)(s & 1ull);
  const unsigned int flag = (unsigned int)(s + 1);
  const int R = D.nranks, me = D.rank;
  const long long chunk = ((long long)n + R - 1) / R;
  auto slot = [&](int owner_rank, int phase, int sender, long long i) {
    return reinterpret_cast<unsigned long long*>(D.mailbox[owner_rank] + D.vec_data_off) +
           2 * ((((long long)phase * 2 + par) * R + sender) * chunk + i);
  };
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < chunk; i += (long long)gridDim.x * blockDim.x) {
    for (int q = 0; q < R; ++q) {                                   // reduce-scatter: send
      if (q == me) continue;
      const long long idx = q * chunk + i;
      ll_store(slot(q, 0, me, i), idx < n ? x[idx] : 0.0, flag);
    }
    double acc = 0.0;
    for (int q = 0; q < R; ++q) {                                   // my chunk: contributions in rank order
      double v;
      if (q == me) { const long long idx = me * chunk + i; v = idx < n ? x[idx] : 0.0; }
      else ll_wait(slot(me, 0, q, i), flag, &v, D.err);
      acc += v;
    }
    for (int q = 0; q < R; ++q)                                     // all-gather: send the reduced element
      if (q != me) ll_store(slot(q, 1, me, i), acc, flag);
    for (int q = 0; q < R; ++q) {
      double v = acc;
      if (q != me) ll_wait(slot(me, 1, q, i), flag, &v, D.err);
      const long long idx = q * chunk + i;
      if (idx < n) x[idx] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(ticket + 2, 1u);
    if (t == gridDim.x - 1) {
      ticket[2] = 0;
      D.seq[1] = s + 1;
    }
  }
}

__global__ void k_allreduce_scalars_test(DistDev D, double* vals, int K) {
  if (threadIdx.x == 0) dist_allreduce_scalars(D, vals, K);
}

}  // namespace

const Halo* find_halo(const int* rowptr) {
  if (g_nhalos.load(std::memory_order_relaxed) == 0) return nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_halos.find(rowptr);
  return it == g_halos.end() ? nullptr : it->second;
}

Dist* active_dist() { return g_dist; }

int halo_exchange(const Halo* h, double* x, int nb, cudaStream_t st, int phase) {
  if (h == nullptr || h->dev.nneigh == 0) return SFEM_OK;
  Dist* d = g_dist;
  if (!d) { set_error("halo exchange without an active communicator"); return SFEM_ERR_ARG; }
  Prof prof(PC_HALO, 16.0 * nb * h->max_cnt * h->dev.nneigh, st);
  k_halo_exchange<<<h->dev.nneigh, 256, 0, st>>>(d->dev, h->dev, x, nb, phase);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int halo_exchange_pair(const Halo* h1, double* x1, int nb1, const Halo* h2, double* x2, int nb2, cudaStream_t st) {
  if (h1 == nullptr || h1->dev.nneigh == 0) return halo_exchange(h2, x2, nb2, st);
  if (h2 == nullptr || h2->dev.nneigh == 0) return halo_exchange(h1, x1, nb1, st);
  Dist* d = g_dist;
  if (!d) { set_error("halo exchange without an active communicator"); return SFEM_ERR_ARG; }
  Prof prof(PC_HALO, 16.0 * (nb1 * h1->max_cnt * h1->dev.nneigh + nb2 * h2->max_cnt * h2->dev.nneigh), st);
  k_halo_exchange_pair<<<h1->dev.nneigh + h2->dev.nneigh, 256, 0, st>>>(d->dev, h1->dev, x1, nb1, h2->dev, x2, nb2);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int dist_allreduce_vec(Dist* d, double* x, int n, cudaStream_t st, int phase) {
  if (!d || d->dev.nranks <= 1 || n <= 0) return SFEM_OK;
  if (n > d->dev.vec_cap) { set_error("vector all-reduce larger than the mailbox capacity"); return SFEM_ERR_ARG; }
  unsigned int* ticket = reinterpret_cast<unsigned int*>(d->dev.seq + 8);
  Prof prof(PC_HALO, 16.0 * n * d->dev.nranks, st);
  if (phase == 0) {          // production: reduce-scatter + all-gather over flagged words, one launch
    const long long chunk = ((long long)n + d->dev.nranks - 1) / d->dev.nranks;
    k_allreduce_vec_ll<<<grid_for(chunk, 256, 2), 256, 0, st>>>(d->dev, x, n, ticket);
    SFEM_LAUNCH_CHECK();
    return SFEM_OK;
  }
  const int grid = grid_for(n, 256 * 2, 1);
  if (phase != 2) {
    k_allreduce_vec_send<<<grid, 256, 0, st>>>(d->dev, x, n, ticket);
    SFEM_LAUNCH_CHECK();
  }
  if (phase != 1) {
    k_allreduce_vec_recv<<<grid, 256, 0, st>>>(d->dev, x, n, ticket);
    SFEM_LAUNCH_CHECK();
  }
  return SFEM_OK;
}

}  // namespace sfem

using namespace sfem;

struct sfem_dist { sfem::Dist d; };
struct sfem_halo { sfem::Halo h; std::vector<void*> owned; };

extern "C" {

/* mailbox layout helper: words (8 bytes) reserved at the start of every mailbox for the all-reduces */
long long sfem_dist_header_words(int nranks, long long vec_cap) {
  // vector region: [2 phases][2 parities][nranks][chunk] flagged pairs (k_allreduce_vec_ll) or [2][nranks][vec_cap] plain
  // words (the two-kernel test path), whichever is larger
  const long long ll_words = 8LL * (vec_cap + nranks) + 16;
  const long long plain = 2LL * nranks * vec_cap;
  return 2LL * nranks + 4LL * nranks * kAllreduceMaxK + 2LL * nranks + (ll_words > plain ? ll_words : plain);
}

sfem_dist_t sfem_dist_create(int rank, int nranks, long long mailbox_words, long long vec_cap) {
  if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks ||
      mailbox_words < sfem_dist_header_words(nranks, vec_cap)) {
    set_error("sfem_dist_create: bad arguments");
    return nullptr;
  }
  sfem_dist* h = new sfem_dist();
  Dist& d = h->d;
  d.dev.rank = rank; d.dev.nranks = nranks;
  d.mailbox_words = (size_t)mailbox_words;
  double* mb = nullptr;
  if (cudaMalloc(&mb, d.mailbox_words * sizeof(double)) != cudaSuccess ||
      cudaMemset(mb, 0, d.mailbox_words * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&d.dev.seq, 16 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMemset(d.dev.seq, 0, 16 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaMalloc(&d.dev.err, sizeof(int)) != cudaSuccess || cudaMemset(d.dev.err, 0, sizeof(int)) != cudaSuccess) {
    set_error("sfem_dist_create: allocation failed");
    delete h;
    return nullptr;
  }
  d.dev.mailbox[rank] = mb;
  d.dev.sc_flag_off = 0;
  d.dev.sc_data_off = 2LL * nranks;
  d.dev.vec_flag_off = d.dev.sc_data_off + 4LL * nranks * kAllreduceMaxK;       // scalar slots are flagged pairs (2 words)
  d.dev.vec_data_off = d.dev.vec_flag_off + 2LL * nranks;
  d.dev.vec_cap = vec_cap;
  cudaDeviceSynchronize();
  return h;
}

/* out: 64 bytes (cudaIpcMemHandle_t) identifying the local mailbox */
int sfem_dist_ipc_handle(sfem_dist_t h, void* out64) {
  if (!h) { set_error("null dist handle"); return SFEM_ERR_ARG; }
  cudaIpcMemHandle_t mh;
  SFEM_CUDA(cudaIpcGetMemHandle(&mh, h->d.dev.mailbox[h->d.dev.rank]));
  static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(out64, &mh, 64);
  return SFEM_OK;
}

/* handles: nranks x 64 bytes, rank-ordered (all-gathered by the host); maps every peer mailbox */
int sfem_dist_open_peers(sfem_dist_t h, const void* handles) {
  if (!h) { set_error("null dist handle"); return SFEM_ERR_ARG; }
  Dist& d = h->d;
  for (int q = 0; q < d.dev.nranks; ++q) {
    if (q == d.dev.rank) continue;
    cudaIpcMemHandle_t mh;
    std::memcpy(&mh, static_cast<const char*>(handles) + 64 * q, 64);
    void* p = nullptr;
    SFEM_CUDA(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
    d.peer_base[q] = p;
    d.peer_open[q] = true;
    d.dev.mailbox[q] = static_cast<double*>(p);
  }
  return SFEM_OK;
}

/* single-process emulation for tests: rank q's mailbox is a plain device pointer of this process */
int sfem_dist_set_peer_pointer(sfem_dist_t h, int q, void* mailbox) {
  if (!h || q < 0 || q >= h->d.dev.nranks) { set_error("bad peer"); return SFEM_ERR_ARG; }
  h->d.dev.mailbox[q] = static_cast<double*>(mailbox);
  return SFEM_OK;
}

void* sfem_dist_mailbox(sfem_dist_t h) { return h ? h->d.dev.mailbox[h->d.dev.rank] : nullptr; }

/* makes h the communicator used by halo exchanges and the Krylov reductions of this process (NULL: none) */
int sfem_dist_activate(sfem_dist_t h) {
  g_dist = h ? &h->d : nullptr;
  graph_epoch_bump();
  return SFEM_OK;
}

/* 1 if a spin timed out since the last call (and clears it) */
int sfem_dist_error(sfem_dist_t h) {
  if (!h) return 0;
  int e = 0;
  cudaMemcpy(&e, h->d.dev.err, sizeof(int), cudaMemcpyDeviceToHost);
  if (e) cudaMemset(h->d.dev.err, 0, sizeof(int));
  return e;
}

void sfem_dist_destroy(sfem_dist_t h) {
  if (!h) return;
  if (g_dist == &h->d) g_dist = nullptr;
  graph_epoch_bump();
  Dist& d = h->d;
  for (int q = 0; q < d.dev.nranks; ++q)
    if (d.peer_open[q]) cudaIpcCloseMemHandle(d.peer_base[q]);
  cudaFree(d.dev.mailbox[d.dev.rank]);
  cudaFree(d.dev.seq);
  cudaFree(d.dev.err);
  delete h;
}

/* Halo pattern of one level.  All arrays are HOST arrays of length nneigh (send_idx: send_ptr[nneigh] ints);
 * offsets are in 8-byte words from the start of the respective mailbox; cap = words per parity slot. */
sfem_halo_t sfem_halo_create(int nneigh, int n_own, int n_loc, const int* peer, const int* send_ptr, const int* send_idx,
                             const int* recv_cnt, const int* recv_off, const long long* peer_data_off,
                             const long long* peer_flag_off, const long long* my_data_off, const long long* my_flag_off,
                             const long long* cap) {
  if (nneigh < 0 || n_own < 0 || n_loc < n_own) { set_error("sfem_halo_create: bad arguments"); return nullptr; }
  sfem_halo* hh = new sfem_halo();
  Halo& H = hh->h;
  H.dev.nneigh = nneigh; H.dev.n_own = n_own; H.dev.n_loc = n_loc;
  auto up = [&](const void* src, size_t bytes) -> void* {
    void* p = nullptr;
    if (bytes == 0) bytes = 8;
    if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    if (src) cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice); else cudaMemset(p, 0, bytes);
    hh->owned.push_back(p);
    return p;
  };
  std::vector<int> scnt(nneigh > 0 ? nneigh : 1, 0);
  for (int b = 0; b < nneigh; ++b) {
    scnt[b] = send_ptr[b + 1] - send_ptr[b];
    if (scnt[b] > H.max_cnt) H.max_cnt = scnt[b];
    if (recv_cnt[b] > H.max_cnt) H.max_cnt = recv_cnt[b];
  }
  const size_t ni = (size_t)nneigh * sizeof(int), nl = (size_t)nneigh * sizeof(long long);
  H.dev.peer = (const int*)up(peer, ni);
  H.dev.send_cnt = (const int*)up(scnt.data(), ni);
  H.dev.send_ptr = (const int*)up(send_ptr, ni + sizeof(int));
  H.dev.send_idx = (const int*)up(send_idx, (size_t)(nneigh > 0 ? send_ptr[nneigh] : 0) * sizeof(int));
  H.dev.recv_cnt = (const int*)up(recv_cnt, ni);
  H.dev.recv_off = (const int*)up(recv_off, ni);
  H.dev.peer_data_off = (const long long*)up(peer_data_off, nl);
  H.dev.peer_flag_off = (const long long*)up(peer_flag_off, nl);
  H.dev.my_data_off = (const long long*)up(my_data_off, nl);
  H.dev.my_flag_off = (const long long*)up(my_flag_off, nl);
  H.dev.cap = (const long long*)up(cap, nl);
  H.dev.seq = (unsigned long long*)up(nullptr, (size_t)(nneigh > 0 ? nneigh : 1) * sizeof(unsigned long long));
  for (void* p : hh->owned)
    if (!p) { set_error("sfem_halo_create: allocation failed"); sfem_halo_destroy(hh); return nullptr; }
  cudaDeviceSynchronize();
  return hh;
}

void sfem_halo_destroy(sfem_halo_t hh) {
  if (!hh) return;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto it = g_halos.begin(); it != g_halos.end();) {
      if (it->second == &hh->h) { it = g_halos.erase(it); g_nhalos.fetch_sub(1); } else ++it;
    }
  }
  graph_epoch_bump();
  for (void* p : hh->owned) cudaFree(p);
  delete hh;
}

/* every SpMV-family launch on the matrix whose rowptr lives at `rowptr` first fills the ghost entries of
 * its input vector through this halo pattern (vectors must hold n_loc entries per right-hand side) */
int sfem_halo_attach(const int* rowptr, sfem_halo_t hh) {
  if (!rowptr) { set_error("sfem_halo_attach: null matrix"); return SFEM_ERR_ARG; }
  std::lock_guard<std::mutex> lk(g_mu);
  graph_epoch_bump();                       // exchanges are baked into captured launch sequences
  auto it = g_halos.find(rowptr);
  if (hh == nullptr) {
    if (it != g_halos.end()) { g_halos.erase(it); g_nhalos.fetch_sub(1); }
    return SFEM_OK;
  }
  if (it == g_halos.end()) g_nhalos.fetch_add(1);
  g_halos[rowptr] = &hh->h;
  return SFEM_OK;
}

/* phase 0: whole exchange; 1 / 2: send-only / wait+unpack-only halves (single-GPU emulation of several ranks) */
int sfem_halo_exchange(sfem_halo_t hh, double* x, int nb, int phase, void* stream) {
  if (!hh || phase < 0 || phase > 2) { set_error("halo exchange: bad arguments"); return SFEM_ERR_ARG; }
  return halo_exchange(&hh->h, x, nb, (cudaStream_t)stream, phase);
}

int sfem_dist_allreduce_vec(sfem_dist_t h, double* x, int n, int phase, void* stream) {
  if (!h || phase < 0 || phase > 2) { set_error("allreduce vec: bad arguments"); return SFEM_ERR_ARG; }
  return dist_allreduce_vec(&h->d, x, n, (cudaStream_t)stream, phase);
}

/* vals[0..K) (device) <- sum over ranks, K <= 8 (the in-kernel scalar all-reduce, exposed for tests) */
int sfem_dist_allreduce_scalars(sfem_dist_t h, double* vals, int K, void* stream) {
  if (!h || K < 1 || K > kAllreduceMaxK) { set_error("allreduce scalars: bad arguments"); return SFEM_ERR_ARG; }
  k_allreduce_scalars_test<<<1, 32, 0, (cudaStream_t)stream>>>(h->d.dev, vals, K);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // extern "C"
