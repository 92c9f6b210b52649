// Fused gradient-norm clip + Adam step on flat fp32 buffers (HBM-bound: 4 streams read, 3 written).
// Reference: torch.nn.utils.clip_grad_norm_ at src/training/trainer.py:147-150 followed by
// torch.optim.Adam(lr, weight_decay) from scripts/train.py:129-133 (L2 decay folded into the gradient,
// betas (0.9, 0.999), eps 1e-8; update arithmetic of torch/optim/adam.py:_single_tensor_adam).
#include "common.cuh"

namespace pc {

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, long long n, double* __restrict__ out) {
  __shared__ double sh[8];
  double s = 0.0;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(g + i * 4);
    s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float v = g[n4 * 4 + threadIdx.x];
    s += (double)v * v;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                 float lr, float beta1, float beta2, float eps, float wd, float max_norm, const double* __restrict__ norm_sq,
                 float prescale, float step_size, float inv_sqrt_bc2) {
  float coef = prescale;
  if (max_norm > 0.f && norm_sq != nullptr) {
    const float total = (float)(sqrt(norm_sq[0]) * (double)prescale);
    const float c = max_norm / (total + 1e-6f);
    coef *= c < 1.f ? c : 1.f;
  }
  const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    float gv = g[i] * coef;
    if (wd != 0.f) gv = fmaf(wd, pv, gv);
    float mv = m[i], vv = v[i];
    mv = mv + (gv - mv) * omb1;               // exp_avg.lerp_(grad, 1 - beta1)
    vv = vv * beta2 + omb2 * gv * gv;         // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
    const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
    p[i] = pv - step_size * (mv / denom);
    m[i] = mv;
    v[i] = vv;
  }
}

// Graph-capturable variant: the 1-based step count and the learning rate are read from device memory, so a captured
// launch stays valid across replays (bias corrections are recomputed on the device from *step_dev).
__global__ void __launch_bounds__(256)
clip_adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                     const float* __restrict__ lr_dev, float beta1, float beta2, float eps, float wd, float max_norm,
                     const double* __restrict__ norm_sq, float prescale, const long long* __restrict__ step_dev) {
  const double step = (double)step_dev[0];
  const float step_size = (float)((double)lr_dev[0] / (1.0 - pow((double)beta1, step)));
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)beta2, step)));
  float coef = prescale;
  if (max_norm > 0.f && norm_sq != nullptr) {
    const float total = (float)(sqrt(norm_sq[0]) * (double)prescale);
    const float c = max_norm / (total + 1e-6f);
    coef *= c < 1.f ? c : 1.f;
  }
  const float omb1 = 1.f - beta1, omb2 = 1.f - beta2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    float gv = g[i] * coef;
    if (wd != 0.f) gv = fmaf(wd, pv, gv);
    float mv = m[i], vv = v[i];
    mv = mv + (gv - mv) * omb1;
    vv = vv * beta2 + omb2 * gv * gv;
    const float denom = sqrtf(vv) * inv_sqrt_bc2 + eps;
    p[i] = pv - step_size * (mv / denom);
    m[i] = mv;
    v[i] = vv;
  }
}

__global__ void counter_add_kernel(long long* c, long long inc) { c[0] += inc; }

}  // namespace pc

using namespace pc;

extern "C" int pc_counter_add(int64_t* counter, int64_t inc, pc_stream_t stream) {
  PC_REQUIRE(counter != nullptr, PC_EINVAL, "pc_counter_add: null pointer");
  counter_add_kernel<<<1, 1, 0, stream>>>(reinterpret_cast<long long*>(counter), (long long)inc);
  PC_LAUNCH_CHECK("counter_add_kernel");
  return PC_OK;
}

extern "C" int pc_clip_adam_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev, float beta1, float beta2,
                                float eps, float weight_decay, float max_norm, const double* norm_sq, float grad_prescale,
                                const int64_t* step_dev, pc_stream_t stream) {
  PC_REQUIRE(p && g && m && v && lr_dev && step_dev && n > 0, PC_EINVAL, "pc_clip_adam_dev: bad arguments");
  PC_REQUIRE(max_norm <= 0.f || norm_sq != nullptr, PC_EINVAL, "pc_clip_adam_dev: clipping needs norm_sq");
  int grid = ceil_div(n, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  clip_adam_dev_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, n, lr_dev, beta1, beta2, eps, weight_decay, max_norm, norm_sq, grad_prescale,
                                                 reinterpret_cast<const long long*>(step_dev));
  PC_LAUNCH_CHECK("clip_adam_dev_kernel");
  return PC_OK;
}

extern "C" int pc_grad_sumsq(const float* g, int64_t n, double* norm_sq, pc_stream_t stream) {
  PC_REQUIRE(g && norm_sq && n > 0, PC_EINVAL, "pc_grad_sumsq: bad arguments");
  PC_REQUIRE((reinterpret_cast<uintptr_t>(g) & 15) == 0, PC_EINVAL, "pc_grad_sumsq: buffer must be 16-byte aligned");
  int grid = ceil_div(n / 4 + 1, 256);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  grad_sumsq_kernel<<<grid, 256, 0, stream>>>(g, n, norm_sq);
  PC_LAUNCH_CHECK("grad_sumsq_kernel");
  return PC_OK;
}

extern "C" int pc_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                            float eps, float weight_decay, float max_norm, const double* norm_sq, float grad_prescale,
                            int64_t step, pc_stream_t stream) {
  PC_REQUIRE(p && g && m && v && n > 0 && step >= 1, PC_EINVAL, "pc_clip_adam: bad arguments");
  PC_REQUIRE(max_norm <= 0.f || norm_sq != nullptr, PC_EINVAL, "pc_clip_adam: clipping needs norm_sq");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  int grid = ceil_div(n, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  clip_adam_kernel<<<grid, 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, max_norm, norm_sq,
                                             grad_prescale, step_size, inv_sqrt_bc2);
  PC_LAUNCH_CHECK("clip_adam_kernel");
  return PC_OK;
}
