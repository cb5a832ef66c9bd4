// Learner-side steps adjacent to the two hot paths: launch-bound chains of tiny PyTorch ops
// in the reference's example, one launch each here.
//
//   K-L1  vtrace_kernel     : V-trace targets from log importance weights, the reverse scan over T and the policy-
//                             gradient advantages in ONE launch (reference: examples/common/vtrace.py:207-242 -- exp, two
//                             clamps, cat, three elementwise ops, a Python loop of T x 3 tiny kernels, stack, add, cat,
//                             clamp, three more elementwise ops: ~100 launches for T = 20).
//   K-L2  u8_to_f32_kernel  : x.float() * scale for uint8 observations in one pass (reference: examples/atari/models.py:94
//                             `x.float() / 255.0`, two elementwise passes; ATen computes a division by a scalar as a
//                             multiplication by its fp32 reciprocal, and so does this kernel).
// fp32 arithmetic in the reference's operation order with every rounding kept (no FMA contraction): results are
// bit-identical to the PyTorch restatement on the same device.
#include "mb_common.cuh"

#include <algorithm>

namespace mb {
namespace {

// torch.clamp(x, max=c): NaN propagates
__device__ __forceinline__ float clamp_max(float x, float c, bool on) {
  if (!on || x != x) return x;
  return fminf(x, c);
}

struct VtraceParams {
  const float* log_rhos;
  const float* discounts;
  const float* rewards;
  const float* values;
  const float* bootstrap;
  float* vs;
  float* pg;
  uint64_t T, B;
  float clip_rho, clip_pg_rho;
  int has_clip_rho, has_clip_pg_rho;
};

constexpr int kVtCols = 32;      // batch columns per block
constexpr int kVtThreads = 128;

// A block owns kVtCols columns.  All threads first pull the four [T, cols] input panels into shared memory with
// coalesced, independent loads (the scan itself is a chain of T dependent steps: fed from global memory it cost
// T x one DRAM latency, ~14 us for T = 20), one warp scans, all threads write the two output panels back.
__global__ void __launch_bounds__(kVtThreads) vtrace_kernel(const VtraceParams p) {
  extern __shared__ float vt_smem[];
  const uint64_t col0 = (uint64_t)blockIdx.x * kVtCols;
  const uint32_t ncol = (uint32_t)min((uint64_t)kVtCols, p.B - col0);
  const uint32_t T = (uint32_t)p.T;
  const uint32_t panel = T * kVtCols;
  float* s_lr = vt_smem;            // log_rhos, later vs
  float* s_d = s_lr + panel;        // discounts, later pg_advantages
  float* s_r = s_d + panel;
  float* s_v = s_r + panel;
  for (uint32_t i = threadIdx.x; i < panel; i += kVtThreads) {
    const uint32_t t = i / kVtCols, c = i - t * kVtCols;
    if (c < ncol) {
      const uint64_t g = (uint64_t)t * p.B + col0 + c;
      s_lr[i] = p.log_rhos[g];
      s_d[i] = p.discounts[g];
      s_r[i] = p.rewards[g];
      s_v[i] = p.values[g];
    }
  }
  __syncthreads();
  if (threadIdx.x < ncol) {
    const uint32_t c = threadIdx.x;
    const float boot = p.bootstrap[col0 + c];
    float acc = 0.f;        // vs_t - V(x_t), scanned backwards (vtrace.py:221-227)
    float v_next = boot;    // V(x_{t+1})
    float vs_next = boot;   // vs_{t+1}
    for (uint32_t t = T; t-- > 0;) {
      const uint32_t i = t * kVtCols + c;
      const float rho = expf(s_lr[i]);
      const float d = s_d[i], r = s_r[i], v = s_v[i];
      const float crho = clamp_max(rho, p.clip_rho, p.has_clip_rho != 0);
      const float cc = clamp_max(rho, 1.0f, true);
      // deltas = clipped_rhos * (rewards + discounts * values_t_plus_1 - values)
      const float delta = __fmul_rn(crho, __fsub_rn(__fadd_rn(r, __fmul_rn(d, v_next)), v));
      // acc = deltas[t] + discounts[t] * cs[t] * acc
      acc = __fadd_rn(delta, __fmul_rn(__fmul_rn(d, cc), acc));
      const float vs = __fadd_rn(acc, v);
      // pg_advantages = clipped_pg_rhos * (rewards + discounts * vs_t_plus_1 - values)
      const float cpg = clamp_max(rho, p.clip_pg_rho, p.has_clip_pg_rho != 0);
      s_d[i] = __fmul_rn(cpg, __fsub_rn(__fadd_rn(r, __fmul_rn(d, vs_next)), v));
      s_lr[i] = vs;
      v_next = v;
      vs_next = vs;
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < panel; i += kVtThreads) {
    const uint32_t t = i / kVtCols, c = i - t * kVtCols;
    if (c < ncol) {
      const uint64_t g = (uint64_t)t * p.B + col0 + c;
      p.vs[g] = s_lr[i];
      p.pg[g] = s_d[i];
    }
  }
}

// T too long for the shared-memory panels: one thread per column straight from global memory.
__global__ void __launch_bounds__(128) vtrace_long_kernel(const VtraceParams p) {
  const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p.B) return;
  const float boot = p.bootstrap[j];
  float acc = 0.f, v_next = boot, vs_next = boot;
  for (uint64_t t = p.T; t-- > 0;) {
    const uint64_t i = t * p.B + j;
    const float rho = expf(p.log_rhos[i]);
    const float d = p.discounts[i], r = p.rewards[i], v = p.values[i];
    const float crho = clamp_max(rho, p.clip_rho, p.has_clip_rho != 0);
    const float c = clamp_max(rho, 1.0f, true);
    const float delta = __fmul_rn(crho, __fsub_rn(__fadd_rn(r, __fmul_rn(d, v_next)), v));
    acc = __fadd_rn(delta, __fmul_rn(__fmul_rn(d, c), acc));
    const float vs = __fadd_rn(acc, v);
    const float cpg = clamp_max(rho, p.clip_pg_rho, p.has_clip_pg_rho != 0);
    p.pg[i] = __fmul_rn(cpg, __fsub_rn(__fadd_rn(r, __fmul_rn(d, vs_next)), v));
    p.vs[i] = vs;
    v_next = v;
    vs_next = vs;
  }
}

__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, uint64_t n,
                                                        float scale, int vec_ok) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec_ok) {
    // 16 observations per thread-iteration: one 16 B load, four 16 B stores
    const uint64_t nvec = n >> 4;
    for (; i < nvec; i += stride) {
      const uint4 q = ld_stream_v4(src + i * 16);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
      float4* out = reinterpret_cast<float4*>(dst + i * 16);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 f;
        f.x = __fmul_rn((float)(w[k] & 0xffu), scale);
        f.y = __fmul_rn((float)((w[k] >> 8) & 0xffu), scale);
        f.z = __fmul_rn((float)((w[k] >> 16) & 0xffu), scale);
        f.w = __fmul_rn((float)(w[k] >> 24), scale);
        out[k] = f;
      }
    }
    i = (nvec << 4) + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x);
  }
  for (; i < n; i += stride) dst[i] = __fmul_rn((float)src[i], scale);
}

}  // namespace
}  // namespace mb

using namespace mb;

extern "C" {

int mb_vtrace_f32(const float* log_rhos, const float* discounts, const float* rewards, const float* values,
                  const float* bootstrap_value, int has_clip_rho, float clip_rho, int has_clip_pg_rho, float clip_pg_rho,
                  uint64_t T, uint64_t B, float* vs_out, float* pg_advantages_out, mb_stream_t stream) {
  if (T == 0 || B == 0) return 0;
  MB_CHECK_ARG(log_rhos && discounts && rewards && values && bootstrap_value && vs_out && pg_advantages_out,
               "mb_vtrace_f32: null pointer");
  VtraceParams p;
  p.log_rhos = log_rhos;
  p.discounts = discounts;
  p.rewards = rewards;
  p.values = values;
  p.bootstrap = bootstrap_value;
  p.vs = vs_out;
  p.pg = pg_advantages_out;
  p.T = T;
  p.B = B;
  p.clip_rho = clip_rho;
  p.clip_pg_rho = clip_pg_rho;
  p.has_clip_rho = has_clip_rho;
  p.has_clip_pg_rho = has_clip_pg_rho;
  const size_t smem = (size_t)T * kVtCols * 4 * sizeof(float);
  if (smem <= 40 * 1024) {
    vtrace_kernel<<<(uint32_t)((B + kVtCols - 1) / kVtCols), kVtThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
  } else {
    vtrace_long_kernel<<<(uint32_t)((B + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  }
  MB_CUDA(cudaGetLastError());
  return 1;
}

int mb_u8_to_f32(const uint8_t* src, float* dst, uint64_t n, float scale, mb_stream_t stream) {
  if (n == 0) return 0;
  MB_CHECK_ARG(src && dst, "mb_u8_to_f32: null pointer");
  const int sms = sm_count(current_device());
  if (sms <= 0) return MB_ECUDA;
  const int vec_ok = (reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
  const uint64_t work = vec_ok ? std::max<uint64_t>(n >> 4, 1) : n;
  const uint32_t grid = (uint32_t)std::min<uint64_t>((work + 255) / 256, (uint64_t)sms * 8);
  u8_to_f32_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, n, scale, vec_ok);
  MB_CUDA(cudaGetLastError());
  return 1;
}

}  // extern "C"
