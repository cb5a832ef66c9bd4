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
#include <new>
#include <vector>

namespace mb {
namespace {

constexpr int kCopyThreads = 256;
constexpr uint32_t kTileBytes = 16384;  // = kCopyThreads * 4 * 16 B: one unrolled-by-4 pass per full tile
constexpr int kMaxJobs = MB_COPY_MAX_INLINE_JOBS;   // tables up to this length travel in the kernel parameters
constexpr int kSmallJobs = 64;                        // ... in the classic 4 KiB parameter block when they are short

enum : uint8_t {
  kModeBigRows = 0,      // row_bytes >= kTileBytes: a tile is a contiguous span inside one row
  kModeSmallVec16 = 1,   // small rows, everything 16 B aligned: a tile is `rpt` whole rows, flat 16 B vector loop
  kModeSmallGeneric = 2, // small rows, arbitrary alignment: a tile is `rpt` rows, one warp per row
  kModeTma = 3           // bulk-async class: a tile is <= tma_tile bytes inside one row (jobs [0, n_tma))
};

// CAP = 64: fits the 4 KiB parameter block (cheapest launch: the per-item stack/cat launches).  CAP = 512: the large
// (32764-byte) parameter space of CUDA 12.1+ -- a whole aligned unroll gather (T x leaves = 147 jobs) needs no upload.
template <int CAP>
struct CopyParamsT {
  mb_copy_job jobs[CAP];
  uint32_t tile_start[CAP + 1];  // exclusive prefix sum of tiles per job
  uint32_t aux[CAP];             // big rows: tiles per row; small rows: rows per tile
  uint8_t mode[CAP];
  uint32_t njobs;
  uint32_t n_tma;      // jobs [0, n_tma) are the bulk-async class (hybrid kernel only)
  uint32_t tma_tile;   // bytes per bulk copy
  uint16_t tma_warps;  // warps driving rings
  uint8_t tma_stages;
  uint8_t tma_stores;  // store groups allowed to be still reading shared memory
  uint8_t tma_contig;  // tile assignment of the ring workers: 0 strided, 1 contiguous spans
};
using CopyParams = CopyParamsT<kSmallJobs>;
using CopyParamsL = CopyParamsT<kMaxJobs>;
static_assert(sizeof(CopyParams) <= 4000, "CopyParams must fit the 4 KiB kernel parameter block");
static_assert(sizeof(CopyParamsL) <= 32764, "CopyParamsL must fit the large kernel parameter space");

// The same table in DEVICE memory (mb_copy2d_table): any number of jobs in one launch.  Uploaded once per launch by
// one async copy from the context's pinned staging; the kernels read it through L1/L2.
struct CopyParamsG {
  const mb_copy_job* jobs;
  const uint32_t* tile_start;  // njobs + 1
  const uint32_t* aux;
  const uint8_t* mode;
  uint32_t njobs;
  uint32_t n_tma;
  uint32_t tma_tile;
  uint16_t tma_warps;
  uint8_t tma_stages;
  uint8_t tma_stores;
  uint8_t tma_contig;
};

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

// Job of tile t when the caller's tile index only grows: walk a cursor forward (a worker's consecutive tiles are a few
// jobs apart), falling back to a binary search over the rest for long jumps.  With a device-resident table this
// replaces ~11 dependent global loads per tile by one or two.
template <class P>
__device__ __forceinline__ uint32_t seek_job(const P& p, uint32_t t, uint32_t cur, uint32_t end) {
#pragma unroll 1
  for (int step = 0; step < 12; ++step) {
    if (cur + 1 >= end || p.tile_start[cur + 1] > t) return cur;
    ++cur;
  }
  uint32_t lo = cur, hi = end;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (p.tile_start[mid] <= t) lo = mid; else hi = mid;
  }
  return lo;
}

// Tiles [lo, hi) of the LDG class, executed by `nthr` threads of CTA blockIdx.x.  Strided assignment for inline tables
// (few jobs, parameter space); contiguous spans for device-resident tables, so that a CTA's consecutive tiles are
// consecutive jobs and the table cursor moves one entry at a time (a unroll gather has ~1000 one-tile jobs: with a
// strided assignment every tile paid a ~20-load search, which made the tiny leaves the critical path).
template <class P>
__device__ __forceinline__ void ldg_tiles(const P& p, uint32_t lo, uint32_t hi, uint32_t first_job, uint32_t tid,
                                          uint32_t nthr) {
  uint32_t t, step, end;
  if (p.tma_contig) {
    const uint32_t per = (hi - lo + gridDim.x - 1) / gridDim.x;
    t = lo + blockIdx.x * per;
    end = min(hi, t + per);
    step = 1;
  } else {
    t = lo + blockIdx.x;
    end = hi;
    step = gridDim.x;
  }
  uint32_t j = first_job;
  for (; t < end; t += step) {
    j = seek_job(p, t, j, p.njobs);
    const mb_copy_job job = p.jobs[j];
    run_tile(job, p.mode[j], p.aux[j], t - p.tile_start[j], tid, nthr);
  }
}

template <class P>
__device__ __forceinline__ void ldg_body(const P& p) {
  ldg_tiles(p, 0, p.tile_start[p.njobs], 0, threadIdx.x, kCopyThreads);
}
__global__ void __launch_bounds__(kCopyThreads, 4) copy2d_ldg_kernel(const __grid_constant__ CopyParams p) { ldg_body(p); }
__global__ void __launch_bounds__(kCopyThreads, 4) copy2d_ldg_kernel_l(const __grid_constant__ CopyParamsL p) { ldg_body(p); }
__global__ void __launch_bounds__(kCopyThreads, 4) copy2d_ldg_table_kernel(const CopyParamsG p) { ldg_body(p); }

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
// A stream of tiles (the load stream or the store stream of one ring) keeps the job it is in: the table is consulted
// again only when a tile leaves that job's range.
struct TmaCursor {
  uint32_t job = 0, t0 = 1, t1 = 0;  // cached job and its tile range [t0, t1); empty at start
  uint32_t tpr = 1;
  const uint8_t* src = nullptr;
  uint8_t* dst = nullptr;
  uint64_t row_bytes = 0;
  int64_t src_pitch = 0, dst_pitch = 0;
};

template <class P>
__device__ __forceinline__ TmaTile tma_decode(const P& p, uint32_t t, TmaCursor& c) {
  if (t < c.t0 || t >= c.t1) {
    c.job = seek_job(p, t, t >= c.t1 && c.t1 != 0 ? c.job : 0u, p.n_tma);
    const mb_copy_job j = p.jobs[c.job];
    c.t0 = p.tile_start[c.job];
    c.t1 = p.tile_start[c.job + 1];
    c.tpr = p.aux[c.job];
    c.src = static_cast<const uint8_t*>(j.src);
    c.dst = static_cast<uint8_t*>(j.dst);
    c.row_bytes = j.row_bytes;
    c.src_pitch = j.src_pitch;
    c.dst_pitch = j.dst_pitch;
  }
  const uint32_t lt = t - c.t0;
  const uint32_t row = lt / c.tpr;
  const uint64_t col = (uint64_t)(lt - row * c.tpr) * p.tma_tile;
  TmaTile r;
  r.src = c.src + (int64_t)row * c.src_pitch + col;
  r.dst = c.dst + (int64_t)row * c.dst_pitch + col;
  r.bytes = (uint32_t)min((uint64_t)p.tma_tile, c.row_bytes - col);
  return r;
}

template <class P>
__device__ __forceinline__ void hybrid_body(const P& p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full[kHybridWarps][kMaxTmaStages];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tma_total = p.tile_start[p.n_tma];
  if (warp >= p.tma_warps) {
    // ---- LDG warps: the table's non-bulk jobs (tiles [tma_total, total)) ----
    ldg_tiles(p, tma_total, p.tile_start[p.njobs], p.n_tma, threadIdx.x - p.tma_warps * 32,
              (kHybridWarps - p.tma_warps) * 32);
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
  // Tile i of worker w.  Strided (w, w + nworkers, ...): neighbouring workers stream neighbouring tiles -- the tuned
  // layout for the few big jobs of an inline table.  Contiguous (a private span per worker): a worker stays inside one
  // job for many tiles, so a device-resident table of hundreds of jobs is consulted once per job, not once per tile.
  uint32_t first, step, mine;
  if (p.tma_contig) {
    const uint32_t per = (tma_total + nworkers - 1) / nworkers;
    first = w * per;
    step = 1;
    mine = first < tma_total ? min(per, tma_total - first) : 0;
  } else {
    first = w;
    step = nworkers;
    mine = w < tma_total ? (tma_total - w + nworkers - 1) / nworkers : 0;
  }
  if (mine == 0) return;

  // Loads run (stages - stores) tiles ahead of the stores.
  const uint32_t ahead = stages - stores;
  uint32_t issued = 0, ld_stage = 0;
  TmaCursor ld_cursor, st_cursor;  // the load stream and the store stream each remember the job they are in
  auto issue_load = [&]() {
    const TmaTile tl = tma_decode(p, first + issued * step, ld_cursor);
    mbar_expect_tx(&full[warp][ld_stage], tl.bytes);
    bulk_g2s(ring + (size_t)ld_stage * tile, tl.src, tl.bytes, &full[warp][ld_stage]);
    ++issued;
    if (++ld_stage == stages) ld_stage = 0;
  };
  while (issued < mine && issued < ahead) issue_load();
  uint32_t st_stage = 0, parity = 0;
  for (uint32_t k = 0; k < mine; ++k) {
    const TmaTile tl = tma_decode(p, first + k * step, st_cursor);
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
__global__ void __launch_bounds__(kHybridWarps * 32, 1) copy2d_hybrid_kernel(const __grid_constant__ CopyParams p) {
  hybrid_body(p);
}
__global__ void __launch_bounds__(kHybridWarps * 32, 1) copy2d_hybrid_kernel_l(const __grid_constant__ CopyParamsL p) {
  hybrid_body(p);
}
__global__ void __launch_bounds__(kHybridWarps * 32, 1) copy2d_hybrid_table_kernel(const CopyParamsG p) { hybrid_body(p); }

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
  int host_src_ctas;      // LDG grid cap when a table reads host-mapped memory (PCIe-bound)
  int inline_contig, table_contig;  // ring workers: contiguous tile spans (1) or strided tiles (0)
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
    c.host_src_ctas = (int)env_long("MB_COPY_HOST_SRC_CTAS", 64, 1, 4096);
    c.inline_contig = (int)env_long("MB_TMA_INLINE_CONTIG", 0, 0, 1);
    c.table_contig = (int)env_long("MB_TMA_TABLE_CONTIG", 1, 0, 1);
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
uint32_t g_hybrid_smem_set[3] = {0, 0, 0};  // [inline kernel, table kernel, large inline kernel]

// Where the caller says the sources live (MB_SRC_*); UNKNOWN asks the driver per job.
bool source_is_device(const mb_copy_job& j, int src_kind) {
  if (src_kind == MB_SRC_DEVICE) return true;
  if (src_kind == MB_SRC_HOST_MAPPED) return false;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, j.src) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeDevice;
}

// What a launch needs besides the table itself.
struct TablePlan {
  uint32_t njobs = 0, n_tma = 0;
  uint64_t tiles = 0, tma_tiles = 0;
  bool hybrid = false;
  bool any_host_src = false;
};

// Normalise `n` jobs, order the bulk-async class first and fill the table arrays (capacity >= n, tile_start n + 1).
// Shared by the inline (kernel-parameter) and the device-table launches.
int build_table(const mb_copy_job* jobs, int n, int src_kind, mb_copy_job* out_jobs, uint32_t* tile_start, uint32_t* aux,
                uint8_t* mode, std::vector<uint8_t>& bulk_scratch, TablePlan* plan) {
  const CopyTuning& tn = tuning();
  bulk_scratch.resize((size_t)n);
  uint8_t* bulk = bulk_scratch.data();
  uint32_t nj = 0;
  uint64_t bulk_bytes = 0;
  // pass 1: normalise into the tail of out_jobs' capacity order (stable), remember the class
  for (int i = 0; i < n; ++i) {
    if (jobs[i].rows == 0 || jobs[i].row_bytes == 0) continue;
    mb_copy_job j = jobs[i];
    if (j.rows > 1 && j.src_pitch == (int64_t)j.row_bytes && j.dst_pitch == (int64_t)j.row_bytes) {
      j.row_bytes *= j.rows;  // contiguous on both sides: one long row
      j.rows = 1;
    }
    bool is_bulk = tn.impl != kImplLdg && job_tma_ok(j);
    // Host-mapped sources (pinned EnvPool slabs) stay on the LDG path, which is the one validated for zero-copy
    // reads over PCIe; the bulk-async path is for device-resident sources.
    const bool dev_src = (is_bulk || src_kind != MB_SRC_UNKNOWN) ? source_is_device(j, src_kind) : true;
    if (!dev_src) {
      is_bulk = false;
      plan->any_host_src = true;
    }
    bulk[nj] = is_bulk;
    if (is_bulk) bulk_bytes += j.rows * j.row_bytes;
    out_jobs[nj++] = j;
  }
  plan->njobs = nj;
  if (nj == 0) return MB_OK;
  plan->hybrid = bulk_bytes > 0 && (tn.impl == kImplTma || bulk_bytes >= tn.tma_min_bytes);
  if (plan->hybrid) {
    // stable partition: the bulk-async class first, the rest behind it in their original order
    static thread_local std::vector<mb_copy_job> rest;
    rest.clear();
    uint32_t k = 0;
    for (uint32_t i = 0; i < nj; ++i) {
      if (bulk[i]) out_jobs[k++] = out_jobs[i];  // k <= i: never overwrites an unread entry
      else rest.push_back(out_jobs[i]);
    }
    plan->n_tma = k;
    for (const mb_copy_job& j : rest) out_jobs[k++] = j;
  }
  uint64_t tiles = 0;
  for (uint32_t i = 0; i < nj; ++i) {
    const mb_copy_job& j = out_jobs[i];
    tile_start[i] = (uint32_t)tiles;
    uint64_t t;
    if (i < plan->n_tma) {
      const uint64_t tpr = (j.row_bytes + tn.tma_tile - 1) / tn.tma_tile;
      mode[i] = kModeTma;
      aux[i] = (uint32_t)tpr;
      t = tpr * j.rows;
    } else if (j.row_bytes >= kTileBytes) {
      const uint64_t tpr = (j.row_bytes + kTileBytes - 1) / kTileBytes;
      mode[i] = kModeBigRows;
      aux[i] = (uint32_t)tpr;
      t = tpr * j.rows;
    } else {
      const uint64_t rpt = std::max<uint64_t>(1, kTileBytes / j.row_bytes);
      const bool v16 = aligned16(reinterpret_cast<uintptr_t>(j.src)) && aligned16(reinterpret_cast<uintptr_t>(j.dst)) &&
                       aligned16(j.row_bytes) && aligned16((uint64_t)j.src_pitch) && aligned16((uint64_t)j.dst_pitch);
      mode[i] = v16 ? kModeSmallVec16 : kModeSmallGeneric;
      aux[i] = (uint32_t)rpt;
      t = (j.rows + rpt - 1) / rpt;
    }
    tiles += t;
    if (i + 1 == plan->n_tma) plan->tma_tiles = tiles;
    if (tiles >= (1ull << 31)) {
      set_error("mb_copy2d: too many tiles in one launch");
      return MB_EINVAL;
    }
  }
  tile_start[nj] = (uint32_t)tiles;
  plan->tiles = tiles;
  return MB_OK;
}

struct LaunchShape {
  uint32_t grid = 0, smem = 0;
};

int launch_shape(const TablePlan& plan, int which_kernel, LaunchShape* out) {
  const CopyTuning& tn = tuning();
  const int sms = sm_count(current_device());
  if (sms <= 0) return MB_ECUDA;
  if (plan.hybrid) {
    const uint32_t smem = tn.tma_warps * tn.tma_stages * tn.tma_tile;
    {
      std::lock_guard<std::mutex> l(g_attr_mu);
      if (g_hybrid_smem_set[which_kernel] < smem) {
        if (which_kernel == 0)
          MB_CUDA(cudaFuncSetAttribute(copy2d_hybrid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else if (which_kernel == 1)
          MB_CUDA(cudaFuncSetAttribute(copy2d_hybrid_table_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else
          MB_CUDA(cudaFuncSetAttribute(copy2d_hybrid_kernel_l, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        g_hybrid_smem_set[which_kernel] = smem;
      }
    }
    const uint64_t want = std::max<uint64_t>((plan.tma_tiles + tn.tma_warps - 1) / tn.tma_warps, plan.tiles - plan.tma_tiles);
    out->grid = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)sms);
    out->smem = smem;
  } else {
    uint64_t cap = (uint64_t)sms * tn.ctas_per_sm;
    // sources in host-mapped memory are PCIe-bound: a few dozen CTAs keep the link full, the SMs stay free for the
    // kernels of other streams (actor inference) instead of hosting warps that wait on PCIe round trips
    if (plan.any_host_src) cap = std::min<uint64_t>(cap, (uint64_t)tn.host_src_ctas);
    out->grid = (uint32_t)std::min<uint64_t>(plan.tiles, cap);
    out->smem = 0;
  }
  return MB_OK;
}

template <class PT>
int launch_inline(const mb_copy_job* jobs, int n, int src_kind, int which_kernel, void (*ldg)(const PT), void (*hyb)(const PT),
                  cudaStream_t stream) {
  const CopyTuning& tn = tuning();
  static thread_local PT p;  // 29 KiB for the large variant: not on the stack of a Python thread
  static thread_local std::vector<uint8_t> scratch;
  TablePlan plan;
  int rc = build_table(jobs, n, src_kind, p.jobs, p.tile_start, p.aux, p.mode, scratch, &plan);
  if (rc) return rc;
  if (plan.njobs == 0) return 0;
  p.njobs = plan.njobs;
  p.n_tma = plan.n_tma;
  p.tma_tile = tn.tma_tile;
  p.tma_warps = (uint16_t)tn.tma_warps;
  p.tma_stages = (uint8_t)tn.tma_stages;
  p.tma_stores = (uint8_t)tn.tma_stores;
  p.tma_contig = (uint8_t)(which_kernel == 2 ? tn.table_contig : tn.inline_contig);
  LaunchShape ls;
  rc = launch_shape(plan, which_kernel, &ls);
  if (rc) return rc;
  if (plan.hybrid) hyb<<<ls.grid, kHybridWarps * 32, ls.smem, stream>>>(p);
  else ldg<<<ls.grid, kCopyThreads, 0, stream>>>(p);
  MB_CUDA(cudaGetLastError());
  return 1;
}

int launch_chunk(const mb_copy_job* jobs, int n, int src_kind, cudaStream_t stream) {
  if (n <= kSmallJobs) return launch_inline<CopyParams>(jobs, n, src_kind, 0, copy2d_ldg_kernel, copy2d_hybrid_kernel, stream);
  return launch_inline<CopyParamsL>(jobs, n, src_kind, 2, copy2d_ldg_kernel_l, copy2d_hybrid_kernel_l, stream);
}

}  // namespace
}  // namespace mb

// A context owns the staging for device-resident job tables: `depth` slots of pinned host memory + device memory.
struct mb_copy_ctx {
  int device = 0;
  uint32_t max_jobs = 0;
  static constexpr int kDepth = 4;
  struct Slot {
    uint8_t* host = nullptr;
    uint8_t* dev = nullptr;
    cudaEvent_t done = nullptr;
    bool used = false;
  } slot[kDepth];
  size_t off_tile = 0, off_aux = 0, off_mode = 0, bytes = 0;
  int next = 0;
  std::mutex mu;
  std::vector<uint8_t> scratch;
};

using namespace mb;

extern "C" {

int mb_copy2d_batch_ex(const mb_copy_job* jobs, int njobs, int src_kind, mb_stream_t stream_) {
  MB_CHECK_ARG(njobs >= 0 && (jobs != nullptr || njobs == 0), "mb_copy2d_batch: bad job table");
  MB_CHECK_ARG(src_kind >= MB_SRC_UNKNOWN && src_kind <= MB_SRC_HOST_MAPPED, "mb_copy2d_batch: bad src_kind %d", src_kind);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  for (int i = 0; i < njobs; ++i) {
    int rc = validate_job(jobs[i], i);
    if (rc) return rc;
  }
  int launches = 0;
  for (int i = 0; i < njobs; i += kMaxJobs) {
    int rc = launch_chunk(jobs + i, std::min(kMaxJobs, njobs - i), src_kind, stream);
    if (rc < 0) return rc;
    launches += rc;
  }
  return launches;
}

int mb_copy2d_batch(const mb_copy_job* jobs, int njobs, mb_stream_t stream) {
  return mb_copy2d_batch_ex(jobs, njobs, MB_SRC_UNKNOWN, stream);
}

int mb_copy_ctx_create(int device, uint32_t max_jobs, mb_copy_ctx** out) {
  MB_CHECK_ARG(out != nullptr, "mb_copy_ctx_create: out is null");
  *out = nullptr;
  MB_CHECK_ARG(max_jobs >= 1 && max_jobs <= (1u << 20), "mb_copy_ctx_create: max_jobs %u not in [1, 2^20]", max_jobs);
  int prev = -1;
  MB_CUDA(cudaGetDevice(&prev));
  MB_CUDA(cudaSetDevice(device));
  mb_copy_ctx* c = new (std::nothrow) mb_copy_ctx();
  if (!c) return MB_ENOMEM;
  c->device = device;
  c->max_jobs = max_jobs;
  auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
  c->off_tile = up((size_t)max_jobs * sizeof(mb_copy_job));
  c->off_aux = c->off_tile + up(((size_t)max_jobs + 1) * 4);
  c->off_mode = c->off_aux + up((size_t)max_jobs * 4);
  c->bytes = c->off_mode + up((size_t)max_jobs);
  int rc = MB_OK;
  for (auto& sl : c->slot) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&sl.host), c->bytes, cudaHostAllocDefault) != cudaSuccess ||
        cudaMalloc(reinterpret_cast<void**>(&sl.dev), c->bytes) != cudaSuccess ||
        cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming) != cudaSuccess) {
      rc = cuda_fail(cudaGetLastError(), "mb_copy_ctx_create allocation", __FILE__, __LINE__);
      break;
    }
  }
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
  if (rc != MB_OK) {
    mb_copy_ctx_destroy(c);
    return rc;
  }
  *out = c;
  return MB_OK;
}

int mb_copy_ctx_destroy(mb_copy_ctx* c) {
  if (!c) return MB_OK;
  for (auto& sl : c->slot) {
    if (sl.done) {
      if (sl.used) cudaEventSynchronize(sl.done);
      cudaEventDestroy(sl.done);
    }
    if (sl.host) cudaFreeHost(sl.host);
    if (sl.dev) cudaFree(sl.dev);
  }
  delete c;
  return MB_OK;
}

int mb_copy2d_table(mb_copy_ctx* c, const mb_copy_job* jobs, int njobs, int src_kind, mb_stream_t stream_) {
  MB_CHECK_ARG(c != nullptr, "mb_copy2d_table: null context");
  MB_CHECK_ARG(njobs >= 0 && (jobs != nullptr || njobs == 0), "mb_copy2d_table: bad job table");
  MB_CHECK_ARG(src_kind >= MB_SRC_UNKNOWN && src_kind <= MB_SRC_HOST_MAPPED, "mb_copy2d_table: bad src_kind %d", src_kind);
  if (njobs <= kMaxJobs) return mb_copy2d_batch_ex(jobs, njobs, src_kind, stream_);  // fits the kernel parameters
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  for (int i = 0; i < njobs; ++i) {
    int rc = validate_job(jobs[i], i);
    if (rc) return rc;
  }
  const CopyTuning& tn = tuning();
  std::lock_guard<std::mutex> l(c->mu);
  int launches = 0;
  for (int first = 0; first < njobs; first += (int)c->max_jobs) {
    const int n = std::min<int>((int)c->max_jobs, njobs - first);
    mb_copy_ctx::Slot& sl = c->slot[c->next];
    c->next = (c->next + 1) % mb_copy_ctx::kDepth;
    if (sl.used) MB_CUDA(cudaEventSynchronize(sl.done));  // the upload that last used this pinned slot has finished
    // the four arrays are packed for THIS table's length, so the upload is ~57 bytes per job and nothing else
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    const size_t off_tile = up((size_t)n * sizeof(mb_copy_job));
    const size_t off_aux = off_tile + up(((size_t)n + 1) * 4);
    const size_t off_mode = off_aux + up((size_t)n * 4);
    TablePlan plan;
    int rc = build_table(jobs + first, n, src_kind, reinterpret_cast<mb_copy_job*>(sl.host),
                         reinterpret_cast<uint32_t*>(sl.host + off_tile), reinterpret_cast<uint32_t*>(sl.host + off_aux),
                         sl.host + off_mode, c->scratch, &plan);
    if (rc) return rc;
    if (plan.njobs == 0) continue;
    const size_t upload = off_mode + plan.njobs;
    MB_CUDA(cudaMemcpyAsync(sl.dev, sl.host, upload, cudaMemcpyHostToDevice, stream));
    CopyParamsG p;
    p.jobs = reinterpret_cast<const mb_copy_job*>(sl.dev);
    p.tile_start = reinterpret_cast<const uint32_t*>(sl.dev + off_tile);
    p.aux = reinterpret_cast<const uint32_t*>(sl.dev + off_aux);
    p.mode = sl.dev + off_mode;
    p.njobs = plan.njobs;
    p.n_tma = plan.n_tma;
    p.tma_tile = tn.tma_tile;
    p.tma_warps = (uint16_t)tn.tma_warps;
    p.tma_stages = (uint8_t)tn.tma_stages;
    p.tma_stores = (uint8_t)tn.tma_stores;
    p.tma_contig = (uint8_t)tn.table_contig;
    LaunchShape ls;
    rc = launch_shape(plan, 1, &ls);
    if (rc) return rc;
    if (plan.hybrid) copy2d_hybrid_table_kernel<<<ls.grid, kHybridWarps * 32, ls.smem, stream>>>(p);
    else copy2d_ldg_table_kernel<<<ls.grid, kCopyThreads, 0, stream>>>(p);
    MB_CUDA(cudaGetLastError());
    MB_CUDA(cudaEventRecord(sl.done, stream));
    sl.used = true;
    ++launches;
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
