// Shared helpers for libpc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/phoneme_contrast.h"

namespace pc {

void set_error(const char* fmt, ...);
void count_launch();

#define PC_REQUIRE(cond, status, ...)        \
  do {                                       \
    if (!(cond)) {                           \
      pc::set_error(__VA_ARGS__);            \
      return (status);                       \
    }                                        \
  } while (0)

// Call right after a kernel launch.
#define PC_LAUNCH_CHECK(name)                                                   \
  do {                                                                          \
    pc::count_launch();                                                         \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      pc::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return PC_ECUDA;                                                          \
    }                                                                           \
  } while (0)

#define PC_CUDA(call)                                                           \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess) {                                                   \
      pc::set_error("%s: %s", #call, cudaGetErrorString(e__));                  \
      return PC_ECUDA;                                                          \
    }                                                                           \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// Programmatic dependent launch (PDL). A kernel launched through launch_pdl() may be scheduled while its predecessor on the
// stream is still draining: its CTAs run their prologue (barrier init, TMEM allocation, index tables built from kernel
// parameters) and then block in pdl_wait() until the predecessor grid has completed and its writes are visible.
// Contract for every kernel launched this way: NO global-memory access before pdl_wait(). pdl_trigger() lets the NEXT
// kernel on the stream start being scheduled; kernels that allocate TMEM call it only after their allocation so that a
// dependent CTA can never take TMEM columns ahead of a CTA it (transitively) waits for.
// Used for eager launches; inside a stream capture plain launches measured faster (see pdl_enabled in abi.cu; PC_PDL=0/1
// overrides). griddepcontrol.* are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled(cudaStream_t stream);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled(stream) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__host__ __device__ static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Unsigned division by a run-time constant without the ~20-instruction integer divide (Granlund & Montgomery): built once on
// the host, q = n / d for every 32-bit n.
struct FastDiv {
  uint32_t mul, s1, s2, d;
  __host__ static FastDiv make(uint32_t d) {
    FastDiv f;
    f.d = d;
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;                                   // ceil(log2 d)
    f.mul = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d) + 1u;
    f.s1 = l < 1 ? l : 1;
    f.s2 = l < 1 ? 0 : l - 1;
    return f;
  }
  __device__ __forceinline__ uint32_t div(uint32_t n) const {
    const uint32_t t = __umulhi(mul, n);
    return (t + ((n - t) >> s1)) >> s2;
  }
  __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
    q = div(n);
    r = n - q * d;
  }
};

// Train-mode BatchNorm coefficients of one channel from the fp64 sums (sum y, sum y^2): identical arithmetic wherever it is evaluated
// (pc_bn_finalize, the in-kernel finalisation of csrc/bn_act.cu, the stem statistics kernel). sum_y / sum_y2 are passed by value so that a
// caller may hand in numbers it has just computed.
__device__ __forceinline__ void bn_train_coeffs_v(double sum_y, double sum_y2, double count, float eps, float& mean, float& invstd, double& unbiased) {
  const double mu = sum_y / count;
  double var = sum_y2 / count - mu * mu;
  var = var < 0.0 ? 0.0 : var;
  mean = (float)mu;
  invstd = (float)(1.0 / sqrt(var + (double)eps));
  unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
}
__device__ __forceinline__ void bn_train_coeffs(const double* __restrict__ stats, int C, int c, double count, float eps, float& mean, float& invstd,
                                                double& unbiased) {
  bn_train_coeffs_v(stats[c], stats[C + c], count, eps, mean, invstd, unbiased);
}
// one channel of a PcBnFinalize: running statistics, published coefficients
__device__ __forceinline__ void bn_publish_channel(const PcBnFinalize& f, int c, float mean, float invstd, double unbiased) {
  const float g = f.gamma != nullptr ? f.gamma[c] : 1.f, b = f.beta != nullptr ? f.beta[c] : 0.f;
  const float sc = g * invstd;
  if (f.running_mean != nullptr) {
    f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * mean;
    f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)unbiased;
  }
  f.scale[c] = sc;
  f.shift[c] = b - mean * sc;
  if (f.mean != nullptr) f.mean[c] = mean;
  if (f.invstd != nullptr) f.invstd[c] = invstd;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Philox4x32-10 (counter-based RNG; Salmon et al. 2011), used for device noise and dropout masks.
struct Philox {
  __device__ static inline uint4 round10(uint4 ctr, uint2 key) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
      const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
      ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
      key.x += 0x9E3779B9u;
      key.y += 0xBB67AE85u;
    }
    return ctr;
  }
  __device__ static inline float u01(uint32_t x) { return ((x >> 8) + 0.5f) * (1.0f / 16777216.0f); }
  // two N(0,1) from two uniforms (Box-Muller)
  __device__ static inline float2 normal2(uint32_t a, uint32_t b) {
    const float u1 = u01(a), u2 = u01(b);
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincosf(6.283185307179586f * u2, &s, &c);
    return make_float2(r * c, r * s);
  }
};

}  // namespace pc
