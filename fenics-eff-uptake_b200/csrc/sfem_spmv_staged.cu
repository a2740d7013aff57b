// Shared-memory-staged CSR SpMV for sm_100a: warp-specialised TMA (1-D bulk async copy) pipeline.
//
// A tile = `tile_rows` consecutive rows.  Because CSR rows are stored back to back, the tile's
// values / column indices are ONE contiguous span of HBM; the producer warp streams that span (and
// the tile's slice of rowptr) into a shared-memory stage with cp.async.bulk (UBLKCP in SASS),
// completion signalled on an mbarrier.  Eight consumer warps reduce rows out of shared memory
// (LANES lanes per row, shuffle reduction), gather x through the read-only path (x is L2 resident
// for the meshes of interest) and write y coalesced.  STAGES stages keep several tiles of HBM
// traffic in flight per CTA independent of consumer occupancy.
//
// Requirements (checked by the host wrapper): rowptr/cols/vals 16-byte aligned and allocated with
// at least 4 elements of padding past their logical end (span ends are rounded up to 16 B).
#include "sfem_common.cuh"
#include "sfem_internal.h"

#include <cstdlib>

namespace sfem {

namespace {

constexpr int kConsumerThreads = 256;
constexpr int kProducerThreads = 32;
constexpr int kMaxStages = 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct StageLayout {
  int cap;         // nnz capacity per stage (multiple of 4)
  int rp_cap;      // rowptr ints per stage (multiple of 4)
  size_t stage_bytes() const { return (size_t)cap * 12 + (size_t)rp_cap * 4; }
};

// MODE 0: y = A x; MODE 1: y = b - A x
template <int LANES, int MODE>
__global__ void __launch_bounds__(kConsumerThreads + kProducerThreads)
    k_spmv_staged(int nrows, int tile_rows, int ntiles, int stages, int cap, int rp_cap,
                  const int* __restrict__ rowptr, const int* __restrict__ cols, const double* __restrict__ vals,
                  const double* __restrict__ x, const double* __restrict__ b, double* __restrict__ y) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[kMaxStages];
  __shared__ uint64_t empty_bar[kMaxStages];

  const size_t stage_bytes = (size_t)cap * 12 + (size_t)rp_cap * 4;
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
    // ------------------------------------------------------------------ producer
    if (lane32 == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % stages;
        const uint32_t ph = (uint32_t)((it / stages) & 1);
        mbar_wait(&empty_bar[s], ph ^ 1u);
        const int r0 = tile * tile_rows;
        const int r1 = min(r0 + tile_rows, nrows);
        const int k0 = rowptr[r0] & ~3;
        const int k1 = (rowptr[r1] + 3) & ~3;
        const int cnt = k1 - k0;
        const int rcnt = ((r1 - r0 + 1) + 3) & ~3;
        unsigned char* base = smem + (size_t)s * stage_bytes;
        double* sv = reinterpret_cast<double*>(base);
        int* sc = reinterpret_cast<int*>(base + (size_t)cap * 8);
        int* sr = reinterpret_cast<int*>(base + (size_t)cap * 12);
        mbar_expect_tx(&full_bar[s], (uint32_t)(cnt * 12 + rcnt * 4));
        if (cnt > 0) {
          bulk_g2s(sv, vals + k0, (uint32_t)(cnt * 8), &full_bar[s]);
          bulk_g2s(sc, cols + k0, (uint32_t)(cnt * 4), &full_bar[s]);
        }
        bulk_g2s(sr, rowptr + r0, (uint32_t)(rcnt * 4), &full_bar[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ consumers
    const int ct = threadIdx.x - kProducerThreads;      // 0..255
    constexpr int ROWS = kConsumerThreads / LANES;       // rows per pass
    const int lane = ct % LANES;
    const int sub = ct / LANES;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int s = it % stages;
      const uint32_t ph = (uint32_t)((it / stages) & 1);
      mbar_wait(&full_bar[s], ph);
      const unsigned char* base = smem + (size_t)s * stage_bytes;
      const double* sv = reinterpret_cast<const double*>(base);
      const int* sc = reinterpret_cast<const int*>(base + (size_t)cap * 8);
      const int* sr = reinterpret_cast<const int*>(base + (size_t)cap * 12);
      const int r0 = tile * tile_rows;
      const int nr = min(tile_rows, nrows - r0);
      const int k0 = sr[0] & ~3;
      for (int lr = sub; lr - sub < nr; lr += ROWS) {     // uniform trip count across the warp
        const bool valid = lr < nr;
        double acc = 0.0;
        if (valid) {
          const int ks = sr[lr] - k0, ke = sr[lr + 1] - k0;
          for (int k = ks + lane; k < ke; k += LANES) acc = fma(sv[k], __ldg(x + sc[k]), acc);
        }
#pragma unroll
        for (int o = LANES >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (valid && lane == 0) {
          const int row = r0 + lr;
          y[row] = (MODE == 0) ? acc : (b[row] - acc);
        }
      }
      __syncwarp();
      if (lane32 == 0) mbar_arrive(&empty_bar[s]);
    }
  }
}

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

}  // namespace

// Host wrapper.  tile_cap must be the exact maximum nnz of a tile (computed from a host copy of
// rowptr); a smaller value would overflow the stage buffers.
int spmv_staged_plan(const Csr& A, int tile_rows, int tile_cap, int stages, const double* x, const double* b,
                     double* y, int mode, cudaStream_t st) {
  if (A.nrows <= 0) return SFEM_OK;
  if ((reinterpret_cast<uintptr_t>(A.rowptr) | reinterpret_cast<uintptr_t>(A.cols) |
       reinterpret_cast<uintptr_t>(A.vals)) & 15u) {
    set_error("staged spmv: CSR arrays must be 16-byte aligned");
    return SFEM_ERR_ARG;
  }
  if (tile_rows <= 0 || (tile_rows & 3) || tile_cap <= 0 || stages < 2 || stages > kMaxStages) {
    set_error("staged spmv: bad tile plan");
    return SFEM_ERR_ARG;
  }
  const int cap = (tile_cap + 7) & ~3;                 // + slack for the 16-byte rounding at both ends
  const int rp_cap = ((tile_rows + 1) + 3) & ~3;
  const size_t smem = (size_t)stages * ((size_t)cap * 12 + (size_t)rp_cap * 4);
  if (smem > 220 * 1024) {
    set_error("staged spmv: tile does not fit shared memory");
    return SFEM_ERR_ARG;
  }
  const int ntiles = (A.nrows + tile_rows - 1) / tile_rows;
  const int lanes = pick_lanes(A.nnz, A.nrows);
  const int threads = kConsumerThreads + kProducerThreads;
  int per_sm = (int)((220 * 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 6) per_sm = 6;
  per_sm = env_int("SFEM_STAGED_CTAS_PER_SM", per_sm);
  int grid = num_sms() * per_sm;
  if (grid > ntiles) grid = ntiles;

  Prof prof(PC_SPMV_STAGED, 12.0 * A.nnz + 4.0 * A.nrows + 8.0 * A.ncols + 8.0 * A.nrows * (mode == 0 ? 1 : 2), st);
#define SFEM_STAGED_LAUNCH(LN, MD)                                                                      \
  do {                                                                                                   \
    auto kern = k_spmv_staged<LN, MD>;                                                                   \
    SFEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    kern<<<grid, threads, smem, st>>>(A.nrows, tile_rows, ntiles, stages, cap, rp_cap, A.rowptr, A.cols, \
                                      A.vals, x, b, y);                                                  \
  } while (0)

  switch (lanes) {
    case 1:
    case 2:
      if (mode == 0) SFEM_STAGED_LAUNCH(2, 0); else SFEM_STAGED_LAUNCH(2, 1);
      break;
    case 4:
      if (mode == 0) SFEM_STAGED_LAUNCH(4, 0); else SFEM_STAGED_LAUNCH(4, 1);
      break;
    case 8:
      if (mode == 0) SFEM_STAGED_LAUNCH(8, 0); else SFEM_STAGED_LAUNCH(8, 1);
      break;
    case 16:
      if (mode == 0) SFEM_STAGED_LAUNCH(16, 0); else SFEM_STAGED_LAUNCH(16, 1);
      break;
    default:
      if (mode == 0) SFEM_STAGED_LAUNCH(32, 0); else SFEM_STAGED_LAUNCH(32, 1);
      break;
  }
#undef SFEM_STAGED_LAUNCH
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // namespace sfem
