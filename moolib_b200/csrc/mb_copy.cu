// HP-B: batched pitched 2-D byte copies -- the one kernel family behind Batcher.stack / Batcher.cat /
// EnvPool slab gathers / stackFields.  See include/moolib_b200.h for the reference call sites each entry replaces.
//
// Two implementations of the same contract (bit-exact byte movement):
//   * copy2d_ldg_kernel : 256-thread CTAs, persistent grid-stride over 16 KiB tiles, 4x16 B loads in flight per
//                         thread (ld.global.nc.L1::no_allocate.v4), coalesced 128 B-per-4-lanes stores.  Handles
//                         every alignment (head / 16 B body / tail, or 8/4/1 B lanes when src and dst are skewed).
//   * copy2d_tma_kernel : one elected lane per warp drives a ring of cp.async.bulk (UBLKCP) global->shared loads
//                         completing on mbarriers and cp.async.bulk shared->global stores; no register staging.
//                         Needs 16 B aligned src/dst/pitch/row_bytes; used when the whole table qualifies.
// Both are HBM-bound (2 x payload bytes); neither touches tensor cores.
#include "mb_common.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace mb {
namespace {

constexpr int kCopyThreads = 256;
constexpr uint32_t kTileBytes = 16384;  // = kCopyThreads * 4 * 16 B: one unrolled-by-4 pass per full tile
constexpr int kMaxJobs = MB_COPY_MAX_INLINE_JOBS;

enum : uint8_t {
  kModeBigRows = 0,     // row_bytes >= kTileBytes: a tile is a contiguous span inside one row
  kModeSmallVec16 = 1,  // small rows, everything 16 B aligned: a tile is `rpt` whole rows, flat 16 B vector loop
  kModeSmallGeneric = 2 // small rows, arbitrary alignment: a tile is `rpt` rows, one warp per row
};

struct CopyParams {
  mb_copy_job jobs[kMaxJobs];
  uint32_t tile_start[kMaxJobs + 1];  // exclusive prefix sum of tiles per job
  uint32_t aux[kMaxJobs];             // big rows: tiles per row; small rows: rows per tile
  uint8_t mode[kMaxJobs];
  uint32_t njobs;
  uint32_t pad_;
};
static_assert(sizeof(CopyParams) <= 4000, "CopyParams must fit the 4 KiB kernel parameter block");

// ---- span copies -------------------------------------------------------------------------------------------------

template <int W>
struct VecT;
template <>
struct VecT<8> {
  using type = uint2;
};
template <>
struct VecT<4> {
  using type = uint32_t;
};
template <>
struct VecT<2> {
  using type = uint16_t;
};
template <>
struct VecT<1> {
  using type = uint8_t;
};

// 16 B-lane body: all loads of an unrolled group are issued before the first store (memory-level parallelism).
__device__ __forceinline__ void copy_vec16(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t nvec,
                                           uint32_t tid, uint32_t nthr) {
  uint64_t base = 0;
  const uint64_t step = 4ull * nthr;
  for (; base + step <= nvec; base += step) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = ld_stream_v4(src + (base + tid + (uint64_t)k * nthr) * 16);
#pragma unroll
    for (int k = 0; k < 4; ++k) st_stream_v4(dst + (base + tid + (uint64_t)k * nthr) * 16, v[k]);
  }
  // remainder (< 4*nthr vectors): still issue the loads before the stores.  Out-of-range lanes re-load the last
  // vector (clamped index) instead of being predicated off, which keeps v[] in registers.
  if (base < nvec) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t i = min(base + tid + (uint64_t)k * nthr, nvec - 1);
      v[k] = ld_stream_v4(src + i * 16);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t i = base + tid + (uint64_t)k * nthr;
      if (i < nvec) st_stream_v4(dst + i * 16, v[k]);
    }
  }
}

template <int W>
__device__ __forceinline__ void copy_lanes(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t n,
                                           uint32_t tid, uint32_t nthr) {
  using T = typename VecT<W>::type;
  const T* s = reinterpret_cast<const T*>(src);
  T* d = reinterpret_cast<T*>(dst);
  for (uint64_t i = tid; i < n; i += nthr) d[i] = s[i];
}

// Skewed (src and dst differently aligned mod 16) spans: rare, kept out of line so the hot path stays lean.
__device__ __noinline__ void copy_span_skewed(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                              uint64_t len, uint32_t tid, uint32_t nthr) {
  const uint32_t ms = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
  const uint32_t md = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);
  const uint32_t skew = ms ^ md;
  if ((skew & 7u) == 0) {
    uint64_t head = (8u - (ms & 7u)) & 7u;
    if (head > len) head = len;
    if (tid < head) dst[tid] = src[tid];
    const uint64_t n = (len - head) >> 3;
    copy_lanes<8>(src + head, dst + head, n, tid, nthr);
    const uint64_t done = head + (n << 3);
    if (tid < len - done) dst[done + tid] = src[done + tid];
  } else if ((skew & 3u) == 0) {
    uint64_t head = (4u - (ms & 3u)) & 3u;
    if (head > len) head = len;
    if (tid < head) dst[tid] = src[tid];
    const uint64_t n = (len - head) >> 2;
    copy_lanes<4>(src + head, dst + head, n, tid, nthr);
    const uint64_t done = head + (n << 2);
    if (tid < len - done) dst[done + tid] = src[done + tid];
  } else if ((skew & 1u) == 0) {
    uint64_t head = ms & 1u;
    if (head > len) head = len;
    if (tid < head) dst[tid] = src[tid];
    const uint64_t n = (len - head) >> 1;
    copy_lanes<2>(src + head, dst + head, n, tid, nthr);
    const uint64_t done = head + (n << 1);
    if (tid < len - done) dst[done + tid] = src[done + tid];
  } else {
    copy_lanes<1>(src, dst, len, tid, nthr);
  }
}

// Copy `len` bytes src -> dst with `nthr` cooperating threads (a CTA or a warp), any alignment.
__device__ __forceinline__ void copy_span(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint64_t len,
                                          uint32_t tid, uint32_t nthr) {
  const uint32_t ms = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 15u);
  const uint32_t md = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u);
  if (ms == md) {
    uint64_t head = (16u - ms) & 15u;
    if (head > len) head = len;
    if (tid < head) dst[tid] = src[tid];
    const uint64_t nvec = (len - head) >> 4;
    copy_vec16(src + head, dst + head, nvec, tid, nthr);
    const uint64_t done = head + (nvec << 4);
    const uint64_t tail = len - done;
    if (tid < tail) dst[done + tid] = src[done + tid];
  } else {
    copy_span_skewed(src, dst, len, tid, nthr);
  }
}

// One tile of one job, executed by the whole CTA.
__device__ __forceinline__ void run_tile(const mb_copy_job& j, uint32_t mode, uint32_t aux, uint32_t t) {
  const uint8_t* src = static_cast<const uint8_t*>(j.src);
  uint8_t* dst = static_cast<uint8_t*>(j.dst);
  if (mode == kModeBigRows) {
    const uint32_t row = t / aux;
    const uint64_t col = (uint64_t)(t - row * aux) * kTileBytes;
    const uint64_t len = min((uint64_t)kTileBytes, j.row_bytes - col);
    copy_span(src + (int64_t)row * j.src_pitch + col, dst + (int64_t)row * j.dst_pitch + col, len, threadIdx.x,
              kCopyThreads);
  } else {
    const uint64_t row0 = (uint64_t)t * aux;
    const uint32_t nrows = (uint32_t)min((uint64_t)aux, j.rows - row0);
    if (mode == kModeSmallVec16) {
      const uint32_t vpr = (uint32_t)(j.row_bytes >> 4);
      const uint32_t total = nrows * vpr;  // <= kTileBytes/16
      const uint8_t* s0 = src + (int64_t)row0 * j.src_pitch;
      uint8_t* d0 = dst + (int64_t)row0 * j.dst_pitch;
      uint4 v[4];
      uint32_t r[4], c[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t i = min(threadIdx.x + k * kCopyThreads, total - 1);  // clamped: see copy_vec16
        r[k] = i / vpr;
        c[k] = i - r[k] * vpr;
        v[k] = ld_stream_v4(s0 + (int64_t)r[k] * j.src_pitch + (uint64_t)c[k] * 16);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (threadIdx.x + k * kCopyThreads < total)
          st_stream_v4(d0 + (int64_t)r[k] * j.dst_pitch + (uint64_t)c[k] * 16, v[k]);
    } else {
      const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      for (uint32_t r = warp; r < nrows; r += kCopyThreads / 32) {
        copy_span(src + (int64_t)(row0 + r) * j.src_pitch, dst + (int64_t)(row0 + r) * j.dst_pitch, j.row_bytes, lane,
                  32);
      }
    }
  }
}

__global__ void __launch_bounds__(kCopyThreads, 4) copy2d_ldg_kernel(const __grid_constant__ CopyParams p) {
  const uint32_t total = p.tile_start[p.njobs];
  for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
    // binary search: last job whose tile_start <= t (njobs <= 64 -> <= 6 steps, warp-uniform)
    uint32_t lo = 0, hi = p.njobs;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (p.tile_start[mid] <= t) lo = mid; else hi = mid;
    }
    run_tile(p.jobs[lo], p.mode[lo], p.aux[lo], t - p.tile_start[lo]);
  }
}

// Pointer-array gather (uniform rows, DEVICE-resident row pointers): K-B1 / K-B4.
struct GatherParams {
  uint8_t* dst;
  uint64_t dst_pitch;
  const void* const* src_rows;
  uint64_t row_bytes;
  uint64_t nrows;
  uint32_t tiles_per_row;  // big rows
  uint32_t rows_per_tile;  // small rows
  uint32_t total_tiles;
  uint32_t big;
};

__global__ void __launch_bounds__(kCopyThreads, 4) gather_rows_kernel(const __grid_constant__ GatherParams p) {
  for (uint32_t t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
    if (p.big) {
      const uint32_t row = t / p.tiles_per_row;
      const uint64_t col = (uint64_t)(t - row * p.tiles_per_row) * kTileBytes;
      const uint64_t len = min((uint64_t)kTileBytes, p.row_bytes - col);
      const uint8_t* src = static_cast<const uint8_t*>(p.src_rows[row]);
      copy_span(src + col, p.dst + row * p.dst_pitch + col, len, threadIdx.x, kCopyThreads);
    } else {
      const uint64_t row0 = (uint64_t)t * p.rows_per_tile;
      const uint32_t nrows = (uint32_t)min((uint64_t)p.rows_per_tile, p.nrows - row0);
      const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      for (uint32_t r = warp; r < nrows; r += kCopyThreads / 32) {
        const uint8_t* src = static_cast<const uint8_t*>(p.src_rows[row0 + r]);
        copy_span(src, p.dst + (row0 + r) * p.dst_pitch, p.row_bytes, lane, 32);
      }
    }
  }
}

// ---- TMA (bulk async copy) implementation -----------------------------------------------------------------------

constexpr int kTmaWarps = 4;
constexpr int kTmaStages = 6;
constexpr uint32_t kTmaTile = 8192;  // bytes per bulk copy
constexpr int kTmaStoresInFlight = 3;
constexpr uint32_t kTmaSmemBytes = kTmaWarps * kTmaStages * kTmaTile;  // 192 KiB
constexpr uint32_t kTmaMinRow = 2048;  // rows shorter than this go to the LDG kernel

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "MB_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra MB_DONE_%=;\n\t"
      "bra MB_WAIT_%=;\n\t"
      "MB_DONE_%=:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct TmaTile {
  const uint8_t* src;
  uint8_t* dst;
  uint32_t bytes;
};

// tile_start / aux here are in units of kTmaTile (aux = tiles per row; every row is tiled on its own).
__device__ __forceinline__ TmaTile tma_decode(const CopyParams& p, uint32_t t) {
  uint32_t lo = 0, hi = p.njobs;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (p.tile_start[mid] <= t) lo = mid; else hi = mid;
  }
  const mb_copy_job& j = p.jobs[lo];
  const uint32_t lt = t - p.tile_start[lo];
  const uint32_t tpr = p.aux[lo];
  const uint32_t row = lt / tpr;
  const uint64_t col = (uint64_t)(lt - row * tpr) * kTmaTile;
  TmaTile r;
  r.src = static_cast<const uint8_t*>(j.src) + (int64_t)row * j.src_pitch + col;
  r.dst = static_cast<uint8_t*>(j.dst) + (int64_t)row * j.dst_pitch + col;
  r.bytes = (uint32_t)min((uint64_t)kTmaTile, j.row_bytes - col);
  return r;
}

__global__ void __launch_bounds__(kTmaWarps * 32, 1) copy2d_tma_kernel(const __grid_constant__ CopyParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[kTmaWarps][kTmaStages];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane != 0) return;  // one elected lane per warp drives its own independent ring
  uint8_t* ring = smem + (size_t)warp * kTmaStages * kTmaTile;
  for (int s = 0; s < kTmaStages; ++s) mbar_init(&full[warp][s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  const uint32_t total = p.tile_start[p.njobs];
  const uint32_t nworkers = gridDim.x * kTmaWarps;
  const uint32_t w = blockIdx.x * kTmaWarps + warp;
  if (w >= total) return;
  const uint32_t mine = (total - w + nworkers - 1) / nworkers;  // tiles w, w+nworkers, ...

  // Loads run kTmaStages - kTmaStoresInFlight tiles ahead of the stores.
  constexpr int kAhead = kTmaStages - kTmaStoresInFlight;
  uint32_t issued = 0;
  auto issue_load = [&](uint32_t k) {
    const TmaTile tl = tma_decode(p, w + k * nworkers);
    const uint32_t s = k % kTmaStages;
    mbar_expect_tx(&full[warp][s], tl.bytes);
    bulk_g2s(ring + (size_t)s * kTmaTile, tl.src, tl.bytes, &full[warp][s]);
  };
  for (; issued < mine && issued < (uint32_t)kAhead; ++issued) issue_load(issued);
  for (uint32_t k = 0; k < mine; ++k) {
    const uint32_t s = k % kTmaStages;
    const TmaTile tl = tma_decode(p, w + k * nworkers);
    mbar_wait(&full[warp][s], (k / kTmaStages) & 1u);
    bulk_s2g(tl.dst, ring + (size_t)s * kTmaTile, tl.bytes);
    if (issued < mine) {
      // stage (issued % kTmaStages) was last used by tile issued - kTmaStages = k - kTmaStoresInFlight; its store
      // group must have finished reading smem, the kTmaStoresInFlight newer groups (k-2, k-1, k) may still be.
      bulk_wait_read<kTmaStoresInFlight>();
      issue_load(issued);
      ++issued;
    }
  }
  bulk_wait_all();
}

// ---- host side ---------------------------------------------------------------------------------------------------

enum CopyImpl { kImplAuto = 0, kImplLdg = 1, kImplTma = 2 };

CopyImpl copy_impl() {
  static CopyImpl impl = [] {
    const char* e = std::getenv("MB_COPY_IMPL");
    if (!e) return kImplAuto;
    if (!std::strcmp(e, "ldg")) return kImplLdg;
    if (!std::strcmp(e, "tma")) return kImplTma;
    return kImplAuto;
  }();
  return impl;
}

int grid_multiplier() {
  static int m = [] {
    const char* e = std::getenv("MB_COPY_CTAS_PER_SM");
    int v = e ? std::atoi(e) : 8;
    return v > 0 && v <= 32 ? v : 8;
  }();
  return m;
}

inline bool aligned16(uint64_t v) { return (v & 15u) == 0; }

bool job_tma_ok(const mb_copy_job& j) {
  return j.row_bytes >= kTmaMinRow && aligned16(reinterpret_cast<uintptr_t>(j.src)) &&
         aligned16(reinterpret_cast<uintptr_t>(j.dst)) && aligned16(j.row_bytes) &&
         (j.rows <= 1 || (aligned16((uint64_t)j.src_pitch) && aligned16((uint64_t)j.dst_pitch)));
}

int validate_job(const mb_copy_job& j, int i) {
  if (j.rows == 0 || j.row_bytes == 0) return MB_OK;
  MB_CHECK_ARG(j.src != nullptr && j.dst != nullptr, "mb_copy2d_batch: job %d has a null pointer", i);
  MB_CHECK_ARG(j.rows < (1ull << 31) && j.row_bytes < (1ull << 40), "mb_copy2d_batch: job %d too large", i);
  return MB_OK;
}

std::once_flag g_tma_attr_once;
cudaError_t g_tma_attr_err = cudaSuccess;

int launch_chunk(const mb_copy_job* jobs, int n, bool allow_tma, cudaStream_t stream) {
  CopyParams p;
  std::memset(&p, 0, sizeof(p));
  bool tma = allow_tma;
  uint32_t nj = 0;
  for (int i = 0; i < n; ++i) {
    if (jobs[i].rows == 0 || jobs[i].row_bytes == 0) continue;
    mb_copy_job j = jobs[i];
    if (j.rows > 1 && j.src_pitch == (int64_t)j.row_bytes && j.dst_pitch == (int64_t)j.row_bytes) {
      j.row_bytes *= j.rows;  // contiguous on both sides: one long row
      j.rows = 1;
    }
    p.jobs[nj++] = j;
    tma = tma && job_tma_ok(j);
  }
  if (nj == 0) return 0;
  p.njobs = nj;
  uint64_t tiles = 0;
  for (uint32_t i = 0; i < nj; ++i) {
    const mb_copy_job& j = p.jobs[i];
    p.tile_start[i] = (uint32_t)tiles;
    uint64_t t;
    if (tma) {
      const uint64_t tpr = (j.row_bytes + kTmaTile - 1) / kTmaTile;
      p.aux[i] = (uint32_t)tpr;
      t = tpr * j.rows;
    } else if (j.row_bytes >= kTileBytes) {
      const uint64_t tpr = (j.row_bytes + kTileBytes - 1) / kTileBytes;
      p.mode[i] = kModeBigRows;
      p.aux[i] = (uint32_t)tpr;
      t = tpr * j.rows;
    } else {
      const uint64_t rpt = std::max<uint64_t>(1, kTileBytes / j.row_bytes);
      const bool v16 = aligned16(reinterpret_cast<uintptr_t>(j.src)) && aligned16(reinterpret_cast<uintptr_t>(j.dst)) &&
                       aligned16(j.row_bytes) && aligned16((uint64_t)j.src_pitch) && aligned16((uint64_t)j.dst_pitch);
      p.mode[i] = v16 ? kModeSmallVec16 : kModeSmallGeneric;
      p.aux[i] = (uint32_t)rpt;
      t = (j.rows + rpt - 1) / rpt;
    }
    tiles += t;
    if (tiles >= (1ull << 31)) {
      set_error("mb_copy2d_batch: too many tiles in one launch");
      return MB_EINVAL;
    }
  }
  p.tile_start[nj] = (uint32_t)tiles;
  const int sms = sm_count(current_device());
  if (sms <= 0) return MB_ECUDA;
  if (tma) {
    std::call_once(g_tma_attr_once, [] {
      g_tma_attr_err = cudaFuncSetAttribute(copy2d_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)kTmaSmemBytes);
    });
    MB_CUDA(g_tma_attr_err);
    const uint32_t want = (uint32_t)((tiles + kTmaWarps - 1) / kTmaWarps);
    const uint32_t grid = std::min<uint32_t>(want, (uint32_t)sms);
    copy2d_tma_kernel<<<grid, kTmaWarps * 32, kTmaSmemBytes, stream>>>(p);
  } else {
    const uint32_t grid = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)sms * grid_multiplier());
    copy2d_ldg_kernel<<<grid, kCopyThreads, 0, stream>>>(p);
  }
  MB_CUDA(cudaGetLastError());
  return 1;
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" {

int mb_copy2d_batch(const mb_copy_job* jobs, int njobs, mb_stream_t stream_) {
  MB_CHECK_ARG(njobs >= 0 && (jobs != nullptr || njobs == 0), "mb_copy2d_batch: bad job table");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  for (int i = 0; i < njobs; ++i) {
    int rc = validate_job(jobs[i], i);
    if (rc) return rc;
  }
  const CopyImpl impl = copy_impl();
  // auto: TMA only pays when there is enough to stream; tiny tables are launch-latency bound either way.
  uint64_t total = 0;
  for (int i = 0; i < njobs; ++i) total += jobs[i].rows * jobs[i].row_bytes;
  const bool allow_tma = impl == kImplTma || (impl == kImplAuto && total >= (8ull << 20));
  int launches = 0;
  for (int i = 0; i < njobs; i += kMaxJobs) {
    int rc = launch_chunk(jobs + i, std::min(kMaxJobs, njobs - i), allow_tma, stream);
    if (rc < 0) return rc;
    launches += rc;
  }
  return launches;
}

int mb_gather_rows(void* dst, uint64_t dst_pitch, const void* const* src_rows_dev, uint64_t row_bytes, uint64_t nrows,
                   mb_stream_t stream_) {
  if (nrows == 0 || row_bytes == 0) return 0;
  MB_CHECK_ARG(dst && src_rows_dev, "mb_gather_rows: null pointer");
  MB_CHECK_ARG(dst_pitch >= row_bytes, "mb_gather_rows: dst_pitch < row_bytes");
  GatherParams p;
  p.dst = static_cast<uint8_t*>(dst);
  p.dst_pitch = dst_pitch;
  p.src_rows = src_rows_dev;
  p.row_bytes = row_bytes;
  p.nrows = nrows;
  p.big = row_bytes >= kTileBytes / 2;
  uint64_t tiles;
  if (p.big) {
    p.tiles_per_row = (uint32_t)((row_bytes + kTileBytes - 1) / kTileBytes);
    p.rows_per_tile = 1;
    tiles = (uint64_t)p.tiles_per_row * nrows;
  } else {
    p.tiles_per_row = 1;
    p.rows_per_tile = (uint32_t)std::max<uint64_t>(1, kTileBytes / row_bytes);
    tiles = (nrows + p.rows_per_tile - 1) / p.rows_per_tile;
  }
  MB_CHECK_ARG(tiles < (1ull << 31), "mb_gather_rows: too many tiles");
  p.total_tiles = (uint32_t)tiles;
  const int sms = sm_count(current_device());
  if (sms <= 0) return MB_ECUDA;
  const uint32_t grid = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)sms * grid_multiplier());
  gather_rows_kernel<<<grid, kCopyThreads, 0, static_cast<cudaStream_t>(stream_)>>>(p);
  MB_CUDA(cudaGetLastError());
  return 1;
}

int mb_stack_slot(void* dst_base, uint64_t outer, uint64_t size, uint64_t slot, uint64_t inner_bytes, const void* src,
                  mb_stream_t stream) {
  MB_CHECK_ARG(slot < size, "mb_stack_slot: slot %llu out of range (size %llu)", (unsigned long long)slot,
               (unsigned long long)size);
  mb_copy_job j;
  j.src = src;
  j.dst = static_cast<uint8_t*>(dst_base) + slot * inner_bytes;
  if (outer == 1) {
    j.rows = 1;
    j.row_bytes = inner_bytes;
  } else {
    j.rows = outer;
    j.row_bytes = inner_bytes;
  }
  j.src_pitch = (int64_t)inner_bytes;
  j.dst_pitch = (int64_t)(size * inner_bytes);
  return mb_copy2d_batch(&j, 1, stream);
}

int mb_cat_narrow(void* dst, const void* src, uint64_t outer, uint64_t dst_dim, uint64_t dst_off, uint64_t src_dim,
                  uint64_t src_off, uint64_t n, uint64_t inner_bytes, mb_stream_t stream) {
  MB_CHECK_ARG(dst_off + n <= dst_dim && src_off + n <= src_dim, "mb_cat_narrow: narrow out of range");
  mb_copy_job j;
  j.src = static_cast<const uint8_t*>(src) + src_off * inner_bytes;
  j.dst = static_cast<uint8_t*>(dst) + dst_off * inner_bytes;
  j.rows = outer;
  j.row_bytes = n * inner_bytes;
  j.src_pitch = (int64_t)(src_dim * inner_bytes);
  j.dst_pitch = (int64_t)(dst_dim * inner_bytes);
  // contiguous on both sides -> one long row (lets the big-row / TMA path take it)
  if (n == dst_dim && n == src_dim) {
    j.row_bytes *= outer;
    j.rows = 1;
  }
  return mb_copy2d_batch(&j, 1, stream);
}

}  // extern "C"

// ---- B3: action scatter into the host-mapped per-env mailboxes ---------------------------------------------------
namespace mb {
namespace {
__global__ void scatter_actions_kernel(uint32_t* __restrict__ counters, uint64_t stride, const int64_t* __restrict__ a,
                                       uint64_t n) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    volatile uint32_t* c = counters + i * stride;
    // single writer per mailbox (the learner); env workers only read (src/env.h:279-292)
    *c = *c + 1u + (uint32_t)a[i];
  }
  __threadfence_system();
}
}  // namespace
}  // namespace mb

extern "C" int mb_scatter_actions(uint32_t* counters, uint64_t stride, const int64_t* actions, uint64_t n,
                                  mb_stream_t stream) {
  if (n == 0) return 0;
  MB_CHECK_ARG(counters && actions && stride >= 1, "mb_scatter_actions: bad arguments");
  const uint32_t threads = 128;
  const uint32_t grid = (uint32_t)((n + threads - 1) / threads);
  mb::scatter_actions_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(counters, stride, actions, n);
  MB_CUDA(cudaGetLastError());
  return 1;
}
