// HP-B: batched pitched 2-D byte copies -- the one kernel family behind Batcher.stack / Batcher.cat /
// EnvPool slab gathers / stackFields.  See include/moolib_b200.h for the reference call sites each entry replaces.
//
// Two implementations of the same contract (bit-exact byte movement):
//   * copy2d_ldg_kernel : 256-thread CTAs, persistent grid-stride over 16 KiB tiles, 4x16 B loads in flight per
//                         thread (ld.global.nc.L1::no_allocate.v4), coalesced 128 B-per-4-lanes stores.  Handles
//                         every alignment (head / 16 B body / tail, or 8/4/1 B lanes when src and dst are skewed).
//   * copy2d_hybrid_kernel : for tables that carry bulk data.  Warps 0..W-1: one elected lane per warp drives a ring
//                         of cp.async.bulk (UBLKCP) global->shared loads completing on mbarriers and cp.async.bulk
//                         shared->global stores -- no register staging -- over the jobs that are 16 B aligned in
//                         src/dst/pitch/row_bytes.  The remaining warps run the LDG path over the table's other
//                         jobs (tiny leaves such as reward/done, skewed rows) in the SAME launch.
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
  kModeBigRows = 0,      // row_bytes >= kTileBytes: a tile is a contiguous span inside one row
  kModeSmallVec16 = 1,   // small rows, everything 16 B aligned: a tile is `rpt` whole rows, flat 16 B vector loop
  kModeSmallGeneric = 2, // small rows, arbitrary alignment: a tile is `rpt` rows, one warp per row
  kModeTma = 3           // bulk-async class: a tile is <= tma_tile bytes inside one row (jobs [0, n_tma))
};

struct CopyParams {
  mb_copy_job jobs[kMaxJobs];
  uint32_t tile_start[kMaxJobs + 1];  // exclusive prefix sum of tiles per job
  uint32_t aux[kMaxJobs];             // big rows: tiles per row; small rows: rows per tile
  uint8_t mode[kMaxJobs];
  uint32_t njobs;
  uint32_t n_tma;      // jobs [0, n_tma) are the bulk-async class (hybrid kernel only)
  uint32_t tma_tile;   // bytes per bulk copy
  uint16_t tma_warps;  // warps driving rings
  uint8_t tma_stages;
  uint8_t tma_stores;  // store groups allowed to be still reading shared memory
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

// One tile of one job, executed by `nthr` cooperating threads (a whole CTA, or the LDG warps of a hybrid CTA).
__device__ __forceinline__ void run_tile(const mb_copy_job& j, uint32_t mode, uint32_t aux, uint32_t t, uint32_t tid,
                                         uint32_t nthr) {
  const uint8_t* src = static_cast<const uint8_t*>(j.src);
  uint8_t* dst = static_cast<uint8_t*>(j.dst);
  if (mode == kModeBigRows) {
    const uint32_t row = t / aux;
    const uint64_t col = (uint64_t)(t - row * aux) * kTileBytes;
    const uint64_t len = min((uint64_t)kTileBytes, j.row_bytes - col);
    copy_span(src + (int64_t)row * j.src_pitch + col, dst + (int64_t)row * j.dst_pitch + col, len, tid, nthr);
  } else {
    const uint64_t row0 = (uint64_t)t * aux;
    const uint32_t nrows = (uint32_t)min((uint64_t)aux, j.rows - row0);
    if (mode == kModeSmallVec16) {
      const uint32_t vpr = (uint32_t)(j.row_bytes >> 4);
      const uint32_t total = nrows * vpr;  // <= kTileBytes/16
      const uint8_t* s0 = src + (int64_t)row0 * j.src_pitch;
      uint8_t* d0 = dst + (int64_t)row0 * j.dst_pitch;
      for (uint32_t base = 0; base < total; base += 4 * nthr) {
        uint4 v[4];
        uint32_t r[4], c[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t i = min(base + tid + k * nthr, total - 1);  // clamped: see copy_vec16
          r[k] = i / vpr;
          c[k] = i - r[k] * vpr;
          v[k] = ld_stream_v4(s0 + (int64_t)r[k] * j.src_pitch + (uint64_t)c[k] * 16);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (base + tid + k * nthr < total)
            st_stream_v4(d0 + (int64_t)r[k] * j.dst_pitch + (uint64_t)c[k] * 16, v[k]);
      }
    } else {
      const uint32_t warp = tid >> 5, lane = tid & 31;
      for (uint32_t r = warp; r < nrows; r += nthr / 32) {
        copy_span(src + (int64_t)(row0 + r) * j.src_pitch, dst + (int64_t)(row0 + r) * j.dst_pitch, j.row_bytes, lane,
                  32);
      }
    }
  }
}

__device__ __forceinline__ uint32_t find_job(const CopyParams& p, uint32_t t) {
  // binary search: last job whose tile_start <= t (njobs <= 64 -> <= 6 steps, warp-uniform)
  uint32_t lo = 0, hi = p.njobs;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (p.tile_start[mid] <= t) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kCopyThreads, 4) copy2d_ldg_kernel(const __grid_constant__ CopyParams p) {
  const uint32_t total = p.tile_start[p.njobs];
  for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
    const uint32_t j = find_job(p, t);
    run_tile(p.jobs[j], p.mode[j], p.aux[j], t - p.tile_start[j], threadIdx.x, kCopyThreads);
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

// ---- hybrid: bulk-async (TMA) rings + LDG warps in one launch ----------------------------------------------------

constexpr int kHybridWarps = 8;
constexpr int kMaxTmaStages = 16;
constexpr uint32_t kTmaMinRow = 2048;  // rows shorter than this stay on the LDG path

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
// wait until at most n of this thread's committed bulk groups are still READING shared memory
__device__ __forceinline__ void bulk_wait_read(uint32_t n) {
  switch (n) {
    case 0: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.bulk.wait_group.read 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.bulk.wait_group.read 6;" ::: "memory"); break;
    default: asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory"); break;
  }
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct TmaTile {
  const uint8_t* src;
  uint8_t* dst;
  uint32_t bytes;
};

// For the bulk class, tile_start / aux are in units of p.tma_tile (aux = tiles per row; rows are tiled one by one).
__device__ __forceinline__ TmaTile tma_decode(const CopyParams& p, uint32_t t) {
  uint32_t lo = 0, hi = p.n_tma;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (p.tile_start[mid] <= t) lo = mid; else hi = mid;
  }
  const mb_copy_job& j = p.jobs[lo];
  const uint32_t lt = t - p.tile_start[lo];
  const uint32_t tpr = p.aux[lo];
  const uint32_t row = lt / tpr;
  const uint64_t col = (uint64_t)(lt - row * tpr) * p.tma_tile;
  TmaTile r;
  r.src = static_cast<const uint8_t*>(j.src) + (int64_t)row * j.src_pitch + col;
  r.dst = static_cast<uint8_t*>(j.dst) + (int64_t)row * j.dst_pitch + col;
  r.bytes = (uint32_t)min((uint64_t)p.tma_tile, j.row_bytes - col);
  return r;
}

__global__ void __launch_bounds__(kHybridWarps * 32, 1) copy2d_hybrid_kernel(const __grid_constant__ CopyParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[kHybridWarps][kMaxTmaStages];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tma_total = p.tile_start[p.n_tma];
  if (warp >= p.tma_warps) {
    // ---- LDG warps: the table's non-bulk jobs (tiles [tma_total, total)) ----
    const uint32_t total = p.tile_start[p.njobs];
    const uint32_t nthr = (kHybridWarps - p.tma_warps) * 32;
    const uint32_t tid = threadIdx.x - p.tma_warps * 32;
    for (uint32_t t = tma_total + blockIdx.x; t < total; t += gridDim.x) {
      const uint32_t j = find_job(p, t);
      run_tile(p.jobs[j], p.mode[j], p.aux[j], t - p.tile_start[j], tid, nthr);
    }
    return;
  }
  if (lane != 0) return;  // one elected lane per warp drives its own independent ring
  const uint32_t stages = p.tma_stages, tile = p.tma_tile, stores = p.tma_stores;
  uint8_t* ring = smem + (size_t)warp * stages * tile;
  for (uint32_t s = 0; s < stages; ++s) mbar_init(&full[warp][s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

  const uint32_t nworkers = gridDim.x * p.tma_warps;
  const uint32_t w = blockIdx.x * p.tma_warps + warp;
  if (w >= tma_total) return;
  const uint32_t mine = (tma_total - w + nworkers - 1) / nworkers;  // tiles w, w+nworkers, ...

  // Loads run (stages - stores) tiles ahead of the stores.
  const uint32_t ahead = stages - stores;
  uint32_t issued = 0, ld_stage = 0;
  auto issue_load = [&]() {
    const TmaTile tl = tma_decode(p, w + issued * nworkers);
    mbar_expect_tx(&full[warp][ld_stage], tl.bytes);
    bulk_g2s(ring + (size_t)ld_stage * tile, tl.src, tl.bytes, &full[warp][ld_stage]);
    ++issued;
    if (++ld_stage == stages) ld_stage = 0;
  };
  while (issued < mine && issued < ahead) issue_load();
  uint32_t st_stage = 0, parity = 0;
  for (uint32_t k = 0; k < mine; ++k) {
    const TmaTile tl = tma_decode(p, w + k * nworkers);
    mbar_wait(&full[warp][st_stage], parity);
    bulk_s2g(tl.dst, ring + (size_t)st_stage * tile, tl.bytes);
    if (++st_stage == stages) {
      st_stage = 0;
      parity ^= 1u;
    }
    if (issued < mine) {
      // the stage about to be refilled was last used by tile k - stores; its store group must have finished
      // reading shared memory, the `stores` newer groups (.., k-1, k) may still be in flight
      bulk_wait_read(stores);
      issue_load();
    }
  }
  bulk_wait_all();
}

// ---- host side ---------------------------------------------------------------------------------------------------

enum CopyImpl { kImplAuto = 0, kImplLdg = 1, kImplTma = 2 };

long env_long(const char* name, long dflt, long lo, long hi) {
  const char* e = std::getenv(name);
  if (!e || !*e) return dflt;
  long v = std::strtol(e, nullptr, 0);
  return v < lo || v > hi ? dflt : v;
}

struct CopyTuning {
  CopyImpl impl;
  int ctas_per_sm;        // LDG kernel grid = min(tiles, SMs * ctas_per_sm)
  uint32_t tma_tile;      // bytes per bulk copy (multiple of 16)
  uint32_t tma_stages;    // ring depth per warp
  uint32_t tma_stores;    // store groups in flight per warp
  uint32_t tma_warps;     // ring-driving warps per CTA (the other 8 - tma_warps warps run the LDG path)
  uint64_t tma_min_bytes; // auto: use the hybrid kernel when the bulk class carries at least this much
};

const CopyTuning& tuning() {
  static const CopyTuning t = [] {
    CopyTuning c;
    const char* e = std::getenv("MB_COPY_IMPL");
    c.impl = !e ? kImplAuto : !std::strcmp(e, "ldg") ? kImplLdg : !std::strcmp(e, "tma") ? kImplTma : kImplAuto;
    c.ctas_per_sm = (int)env_long("MB_COPY_CTAS_PER_SM", 16, 1, 32);
    c.tma_tile = (uint32_t)env_long("MB_TMA_TILE", 16384, 512, 65536) & ~15u;
    c.tma_warps = (uint32_t)env_long("MB_TMA_WARPS", 3, 1, kHybridWarps - 1);
    c.tma_stages = (uint32_t)env_long("MB_TMA_STAGES", 4, 2, kMaxTmaStages);
    while ((uint64_t)c.tma_warps * c.tma_stages * c.tma_tile > 200u * 1024u && c.tma_stages > 2) --c.tma_stages;
    c.tma_stores = (uint32_t)env_long("MB_TMA_STORES", c.tma_stages / 2, 1, 7);
    if (c.tma_stores >= c.tma_stages) c.tma_stores = c.tma_stages - 1;
    c.tma_min_bytes = (uint64_t)env_long("MB_TMA_MIN_BYTES", 1l << 20, 0, 1l << 40);
    return c;
  }();
  return t;
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

std::mutex g_attr_mu;
uint32_t g_hybrid_smem_set = 0;

int launch_chunk(const mb_copy_job* jobs, int n, cudaStream_t stream) {
  const CopyTuning& tn = tuning();
  CopyParams p;
  std::memset(&p, 0, sizeof(p));
  // normalise, then order the bulk-async class first
  mb_copy_job norm[kMaxJobs];
  bool bulk[kMaxJobs];
  uint32_t nj = 0;
  uint64_t bulk_bytes = 0;
  for (int i = 0; i < n; ++i) {
    if (jobs[i].rows == 0 || jobs[i].row_bytes == 0) continue;
    mb_copy_job j = jobs[i];
    if (j.rows > 1 && j.src_pitch == (int64_t)j.row_bytes && j.dst_pitch == (int64_t)j.row_bytes) {
      j.row_bytes *= j.rows;  // contiguous on both sides: one long row
      j.rows = 1;
    }
    bulk[nj] = tn.impl != kImplLdg && job_tma_ok(j);
    if (bulk[nj]) {
      // Host-mapped sources (pinned EnvPool slabs) stay on the LDG path, which is the one validated for zero-copy
      // reads over PCIe; the bulk-async path is for device-resident sources.
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, j.src) != cudaSuccess) {
        cudaGetLastError();
        bulk[nj] = false;
      } else if (attr.type != cudaMemoryTypeDevice) {
        bulk[nj] = false;
      }
    }
    if (bulk[nj]) bulk_bytes += j.rows * j.row_bytes;
    norm[nj++] = j;
  }
  if (nj == 0) return 0;
  const bool hybrid = bulk_bytes > 0 && (tn.impl == kImplTma || bulk_bytes >= tn.tma_min_bytes);
  uint32_t k = 0;
  if (hybrid) {
    for (uint32_t i = 0; i < nj; ++i)
      if (bulk[i]) { p.jobs[k] = norm[i]; p.mode[k++] = kModeTma; }
    p.n_tma = k;
    for (uint32_t i = 0; i < nj; ++i)
      if (!bulk[i]) p.jobs[k++] = norm[i];
  } else {
    for (uint32_t i = 0; i < nj; ++i) p.jobs[k++] = norm[i];
  }
  p.njobs = nj;
  p.tma_tile = tn.tma_tile;
  p.tma_warps = (uint16_t)tn.tma_warps;
  p.tma_stages = (uint8_t)tn.tma_stages;
  p.tma_stores = (uint8_t)tn.tma_stores;
  uint64_t tiles = 0;
  for (uint32_t i = 0; i < nj; ++i) {
    const mb_copy_job& j = p.jobs[i];
    p.tile_start[i] = (uint32_t)tiles;
    uint64_t t;
    if (i < p.n_tma) {
      const uint64_t tpr = (j.row_bytes + tn.tma_tile - 1) / tn.tma_tile;
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
  if (hybrid) {
    const uint32_t smem = tn.tma_warps * tn.tma_stages * tn.tma_tile;
    {
      std::lock_guard<std::mutex> l(g_attr_mu);
      if (g_hybrid_smem_set < smem) {
        MB_CUDA(cudaFuncSetAttribute(copy2d_hybrid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g_hybrid_smem_set = smem;
      }
    }
    const uint64_t tma_tiles = p.tile_start[p.n_tma];
    const uint64_t want = std::max<uint64_t>((tma_tiles + tn.tma_warps - 1) / tn.tma_warps, tiles - tma_tiles);
    const uint32_t grid = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)sms);
    copy2d_hybrid_kernel<<<grid, kHybridWarps * 32, smem, stream>>>(p);
  } else {
    const uint32_t grid = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)sms * tn.ctas_per_sm);
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
  int launches = 0;
  for (int i = 0; i < njobs; i += kMaxJobs) {
    int rc = launch_chunk(jobs + i, std::min(kMaxJobs, njobs - i), stream);
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
  const uint32_t grid = (uint32_t)std::min<uint64_t>(tiles, (uint64_t)sms * tuning().ctas_per_sm);
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
