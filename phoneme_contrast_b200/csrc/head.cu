// Network head: SpatialAttention gate + global average pool, then Linear + BatchNorm1d + L2 normalise,
// forward and backward. Reference: src/models/phoneme_cnn.py:129-143 (SpatialAttention), :117-124 / :295-302
// (pool, projection, F.normalize). Tensors are tiny here ([B,HW,C] with HW*C <= 32 K floats per sample,
// [B,128] embeddings); the kernels are latency-bound, so they are kept few and simple.
#include <stdlib.h>

#include "common.cuh"

namespace pc {

// one CTA per sample
__global__ void __launch_bounds__(256)
attn_pool_fwd_kernel(const float* __restrict__ a, int HW, int C, const float* __restrict__ w, const float* __restrict__ b0,
                     float* __restrict__ gate, float* __restrict__ pooled) {
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* ab = a + (size_t)b * HW * C;
  float* gb = gate + (size_t)b * HW;
  if (w != nullptr) {
    const float bias = b0 != nullptr ? b0[0] : 0.f;
    for (int p = warp; p < HW; p += 8) {
      float s = 0.f;
      for (int c = lane; c < C; c += 32) s = fmaf(ab[(size_t)p * C + c], w[c], s);
      s = warp_sum(s);
      if (lane == 0) gb[p] = 1.0f / (1.0f + expf(-(s + bias)));
    }
  } else {
    for (int p = tid; p < HW; p += 256) gb[p] = 1.0f;
  }
  __syncthreads();
  const float inv = 1.0f / (float)HW;
  for (int c = tid; c < C; c += 256) {
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s = fmaf(ab[(size_t)p * C + c], gb[p], s);
    pooled[(size_t)b * C + c] = s * inv;
  }
}

__global__ void __launch_bounds__(256)
attn_pool_bwd_kernel(const float* __restrict__ a, const float* __restrict__ gate, const float* __restrict__ dpooled, int HW,
                     int C, const float* __restrict__ w, float* __restrict__ da, float* __restrict__ dw,
                     float* __restrict__ db0) {
  extern __shared__ float du[];   // [HW]
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* ab = a + (size_t)b * HW * C;
  const float* gb = gate + (size_t)b * HW;
  const float* gp = dpooled + (size_t)b * C;
  float* dab = da + (size_t)b * HW * C;
  const float inv = 1.0f / (float)HW;
  if (w != nullptr) {
    float dbacc = 0.f;
    for (int p = warp; p < HW; p += 8) {
      float t = 0.f;
      for (int c = lane; c < C; c += 32) t = fmaf(ab[(size_t)p * C + c], gp[c], t);
      t = warp_sum(t) * inv;
      const float s = gb[p];
      const float d = s * (1.0f - s) * t;
      if (lane == 0) { du[p] = d; dbacc += d; }
    }
    __shared__ float dbs[8];
    if (lane == 0) dbs[warp] = dbacc;
    __syncthreads();
    if (tid == 0 && db0 != nullptr) {
      float s = 0.f;
      for (int k = 0; k < 8; ++k) s += dbs[k];
      atomicAdd(db0, s);
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += 256) {
    const float g = gp[c] * inv;
    if (w != nullptr) {
      const float wc = w[c];
      float dwacc = 0.f;
      for (int p = 0; p < HW; ++p) {
        const float av = ab[(size_t)p * C + c];
        dab[(size_t)p * C + c] = fmaf(g, gb[p], du[p] * wc);
        dwacc = fmaf(du[p], av, dwacc);
      }
      atomicAdd(dw + c, dwacc);
    } else {
      for (int p = 0; p < HW; ++p) dab[(size_t)p * C + c] = g;
    }
  }
}

// Single-pass forms for C = 128 * NV (NV float4 per lane): a warp owns pixels p = warp, warp + 8, ...; it reads row p ONCE (coalesced
// 512-byte segments), forms the gate from the row's dot product with w and accumulates gate * row into per-lane partial sums; the eight
// warps' partials are added through shared memory in a fixed order. The per-channel loops of the kernels above walk all HW pixels
// serially with C of 256 threads active (35 / 45 us at 64 x 250 x 128, the cnn_small head); these take one pass with all warps busy.
template <int NV>
__global__ void __launch_bounds__(256)
attn_pool_fwd1_kernel(const float* __restrict__ a, int HW, const float* __restrict__ w, const float* __restrict__ b0, float* __restrict__ gate,
                      float* __restrict__ pooled, float* __restrict__ scratch, unsigned int* __restrict__ counters) {
  constexpr int C = 128 * NV;
  __shared__ float4 part[8][32 * NV];
  __shared__ unsigned int s_ticket;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = gridDim.y, sy = blockIdx.y;          // S blocks share a sample's pixels (few samples: cnn_small's 64 would leave 84 SMs idle)
  const float4* ab = reinterpret_cast<const float4*>(a + (size_t)b * HW * C);
  float* gb = gate + (size_t)b * HW;
  float4 wv[NV], acc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    wv[j] = w != nullptr ? reinterpret_cast<const float4*>(w)[lane + 32 * j] : make_float4(0.f, 0.f, 0.f, 0.f);
    acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float bias = (w != nullptr && b0 != nullptr) ? b0[0] : 0.f;
  for (int p = warp + 8 * sy; p < HW; p += 8 * S) {
    float4 v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = ab[(size_t)p * (C / 4) + lane + 32 * j];
    float g = 1.0f;
    if (w != nullptr) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) s = fmaf(v[j].x, wv[j].x, fmaf(v[j].y, wv[j].y, fmaf(v[j].z, wv[j].z, fmaf(v[j].w, wv[j].w, s))));
      s = warp_sum(s);
      g = 1.0f / (1.0f + expf(-(s + bias)));
    }
    if (lane == 0) gb[p] = g;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      acc[j].x = fmaf(v[j].x, g, acc[j].x); acc[j].y = fmaf(v[j].y, g, acc[j].y);
      acc[j].z = fmaf(v[j].z, g, acc[j].z); acc[j].w = fmaf(v[j].w, g, acc[j].w);
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) part[warp][lane + 32 * j] = acc[j];
  __syncthreads();
  const float inv = 1.0f / (float)HW;
  float4* mine = reinterpret_cast<float4*>(scratch) + ((size_t)b * S + sy) * (C / 4);
  for (int i = tid; i < 32 * NV; i += 256) {
    float4 s = part[0][i];
#pragma unroll
    for (int k = 1; k < 8; ++k) { const float4 q = part[k][i]; s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w; }
    if (S == 1) reinterpret_cast<float4*>(pooled + (size_t)b * C)[i] = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
    else mine[i] = s;
  }
  if (S == 1) return;
  // the block that arrives last for this sample adds the S partial sums in a FIXED order (the result does not depend on scheduling)
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(&counters[b], 1u);
  __syncthreads();
  if (s_ticket != (unsigned)(S - 1)) return;
  __threadfence();
  const float4* all = reinterpret_cast<const float4*>(scratch) + (size_t)b * S * (C / 4);
  for (int i = tid; i < 32 * NV; i += 256) {
    float4 s = __ldcg(all + i);
    for (int k = 1; k < S; ++k) { const float4 q = __ldcg(all + (size_t)k * (C / 4) + i); s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w; }
    reinterpret_cast<float4*>(pooled + (size_t)b * C)[i] = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
  }
  if (tid == 0) counters[b] = 0u;                    // ready for the next launch (the counters start zeroed and stay so between launches)
}

template <int NV>
__global__ void __launch_bounds__(256)
attn_pool_bwd1_kernel(const float* __restrict__ a, const float* __restrict__ gate, const float* __restrict__ dpooled, int HW, const float* __restrict__ w,
                      float* __restrict__ da, float* __restrict__ dw, float* __restrict__ db0) {
  constexpr int C = 128 * NV;
  __shared__ float4 part[8][32 * NV];
  __shared__ float dbs[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float4* ab = reinterpret_cast<const float4*>(a + (size_t)b * HW * C);
  float4* dab = reinterpret_cast<float4*>(da + (size_t)b * HW * C);
  const float* gb = gate + (size_t)b * HW;
  const float inv = 1.0f / (float)HW;
  float4 gp[NV], wv[NV], dwacc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    gp[j] = reinterpret_cast<const float4*>(dpooled + (size_t)b * C)[lane + 32 * j];
    gp[j].x *= inv; gp[j].y *= inv; gp[j].z *= inv; gp[j].w *= inv;
    wv[j] = w != nullptr ? reinterpret_cast<const float4*>(w)[lane + 32 * j] : make_float4(0.f, 0.f, 0.f, 0.f);
    dwacc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float dbacc = 0.f;
  for (int p = warp + 8 * (int)blockIdx.y; p < HW; p += 8 * (int)gridDim.y) {
    if (w == nullptr) {
#pragma unroll
      for (int j = 0; j < NV; ++j) dab[(size_t)p * (C / 4) + lane + 32 * j] = gp[j];
      continue;
    }
    float4 v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = ab[(size_t)p * (C / 4) + lane + 32 * j];
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) t = fmaf(v[j].x, gp[j].x, fmaf(v[j].y, gp[j].y, fmaf(v[j].z, gp[j].z, fmaf(v[j].w, gp[j].w, t))));
    t = warp_sum(t);                                  // = (row . dpooled) / HW
    const float sg = gb[p];
    const float d = sg * (1.0f - sg) * t;
    dbacc += d;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float4 o;
      o.x = fmaf(gp[j].x, sg, d * wv[j].x); o.y = fmaf(gp[j].y, sg, d * wv[j].y);
      o.z = fmaf(gp[j].z, sg, d * wv[j].z); o.w = fmaf(gp[j].w, sg, d * wv[j].w);
      dab[(size_t)p * (C / 4) + lane + 32 * j] = o;
      dwacc[j].x = fmaf(d, v[j].x, dwacc[j].x); dwacc[j].y = fmaf(d, v[j].y, dwacc[j].y);
      dwacc[j].z = fmaf(d, v[j].z, dwacc[j].z); dwacc[j].w = fmaf(d, v[j].w, dwacc[j].w);
    }
  }
  if (w == nullptr) return;
#pragma unroll
  for (int j = 0; j < NV; ++j) part[warp][lane + 32 * j] = dwacc[j];
  if (lane == 0) dbs[warp] = dbacc;
  __syncthreads();
  for (int i = tid; i < 32 * NV; i += 256) {
    float4 s = part[0][i];
#pragma unroll
    for (int k = 1; k < 8; ++k) { const float4 q = part[k][i]; s.x += q.x; s.y += q.y; s.z += q.z; s.w += q.w; }
    atomicAdd(dw + 4 * i + 0, s.x); atomicAdd(dw + 4 * i + 1, s.y); atomicAdd(dw + 4 * i + 2, s.z); atomicAdd(dw + 4 * i + 3, s.w);
  }
  if (tid == 0 && db0 != nullptr) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += dbs[k];
    atomicAdd(db0, s);
  }
}

// ------------------------------------------------------------------------------------------------ projection head
// ws layout (floats): z [B*N] | mean [N] | invstd [N] | scratch [B*N]
// z[b][n] = x[b][:] . W[n][:] + bias[n]: one warp per (b, 4 consecutive n); lanes stride over K with float4 loads, so x and W
// rows are read as full 512-byte lines and the x row is shared by the four outputs.
__global__ void __launch_bounds__(256)
head_linear_kernel(const float* __restrict__ x, int B, int K, int N, const float* __restrict__ W,
                   const float* __restrict__ bias, float* __restrict__ z, int vec4) {
  const int lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n4 = (N + 3) >> 2;
  if (gw >= B * n4) return;
  const int b = gw / n4, n0 = (gw - b * n4) * 4;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  const float* xr = x + (size_t)b * K;
  if (vec4) {     // K % 4 == 0 and both base pointers 16-byte aligned (parameters may be views into a flat buffer)
    for (int k = lane * 4; k < K; k += 128) {
      const float4 xv = *reinterpret_cast<const float4*>(xr + k);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (n0 + q < N) {
          const float4 wv = *reinterpret_cast<const float4*>(W + (size_t)(n0 + q) * K + k);
          s[q] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, s[q]))));
        }
      }
    }
  } else {
    for (int k = lane; k < K; k += 32)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (n0 + q < N) s[q] = fmaf(xr[k], W[(size_t)(n0 + q) * K + k], s[q]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) s[q] = warp_sum(s[q]);
  if (lane < 4 && n0 + lane < N) {
    const float v = lane == 0 ? s[0] : (lane == 1 ? s[1] : (lane == 2 ? s[2] : s[3]));
    z[(size_t)b * N + n0 + lane] = v + (bias != nullptr ? bias[n0 + lane] : 0.f);
  }
}

// one CTA (128 threads) per feature n: batch statistics (two-pass) + running-stat update
__global__ void __launch_bounds__(128)
head_bn_stats_kernel(const float* __restrict__ z, int B, int N, float* __restrict__ rmean, float* __restrict__ rvar,
                     int64_t* __restrict__ nbt, float momentum, float eps, int training, float* __restrict__ mean_o,
                     float* __restrict__ invstd_o) {
  __shared__ double sh[4];
  __shared__ double mu_s;
  const int n = blockIdx.x, tid = threadIdx.x;
  if (!training) {
    if (tid == 0) { mean_o[n] = rmean[n]; invstd_o[n] = 1.0f / sqrtf(rvar[n] + eps); }
    return;
  }
  double s = 0.0;
  for (int b = tid; b < B; b += 128) s += (double)z[(size_t)b * N + n];
  s = warp_sum(s);
  if ((tid & 31) == 0) sh[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) mu_s = (sh[0] + sh[1] + sh[2] + sh[3]) / (double)B;
  __syncthreads();
  const double mu = mu_s;
  double q = 0.0;
  for (int b = tid; b < B; b += 128) { const double d = (double)z[(size_t)b * N + n] - mu; q += d * d; }
  q = warp_sum(q);
  __syncthreads();
  if ((tid & 31) == 0) sh[tid >> 5] = q;
  __syncthreads();
  if (tid == 0) {
    const double var = (sh[0] + sh[1] + sh[2] + sh[3]) / (double)B;
    mean_o[n] = (float)mu;
    invstd_o[n] = (float)(1.0 / sqrt(var + (double)eps));
    if (rmean != nullptr) {
      const double unbiased = B > 1 ? var * (double)B / (double)(B - 1) : var;
      rmean[n] = (1.f - momentum) * rmean[n] + momentum * (float)mu;
      rvar[n] = (1.f - momentum) * rvar[n] + momentum * (float)unbiased;
    }
    if (n == 0 && nbt != nullptr) nbt[0] += 1;
  }
}

// warp per row: zn = (z-mu)*invstd*gamma+beta ; e = zn / max(||zn||, 1e-12)
__global__ void __launch_bounds__(256)
head_normalize_kernel(const float* __restrict__ z, int B, int N, const float* __restrict__ mean, const float* __restrict__ invstd,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ emb) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  float ss = 0.f;
  for (int n = lane; n < N; n += 32) {
    const float v = fmaf((z[(size_t)row * N + n] - mean[n]) * invstd[n], gamma[n], beta[n]);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
  for (int n = lane; n < N; n += 32) {
    const float v = fmaf((z[(size_t)row * N + n] - mean[n]) * invstd[n], gamma[n], beta[n]);
    emb[(size_t)row * N + n] = v * inv;
  }
}

// warp per row: dzn = (demb - e (e.demb)) / ||zn||   -> scratch
__global__ void __launch_bounds__(256)
head_normalize_bwd_kernel(const float* __restrict__ demb, const float* __restrict__ z, int B, int N, const float* __restrict__ mean,
                          const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                          float* __restrict__ dzn) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  float ss = 0.f, dot = 0.f;
  for (int n = lane; n < N; n += 32) {
    const float v = fmaf((z[(size_t)row * N + n] - mean[n]) * invstd[n], gamma[n], beta[n]);
    ss = fmaf(v, v, ss);
    dot = fmaf(v, demb[(size_t)row * N + n], dot);
  }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float nrm = sqrtf(ss);
  const bool clamped = nrm < 1e-12f;           // F.normalize: x / clamp_min(norm, eps) -> constant denominator
  const float inv = 1.0f / fmaxf(nrm, 1e-12f);
  for (int n = lane; n < N; n += 32) {
    const float v = fmaf((z[(size_t)row * N + n] - mean[n]) * invstd[n], gamma[n], beta[n]);
    const float g = demb[(size_t)row * N + n];
    dzn[(size_t)row * N + n] = clamped ? g * inv : (g - v * inv * (dot * inv)) * inv;
  }
}

// one CTA per feature: BatchNorm1d backward in place on dzn -> dz
__global__ void __launch_bounds__(128)
head_bn_bwd_kernel(float* __restrict__ dzn, const float* __restrict__ z, int B, int N, const float* __restrict__ mean,
                   const float* __restrict__ invstd, const float* __restrict__ gamma, int training, float* __restrict__ dgamma,
                   float* __restrict__ dbeta, float* __restrict__ dbias) {
  __shared__ double sh[2][4];
  __shared__ float bc[2];
  const int n = blockIdx.x, tid = threadIdx.x;
  const float mu = mean[n], is = invstd[n], g = gamma[n];
  double s1 = 0.0, s2 = 0.0;
  for (int b = tid; b < B; b += 128) {
    const float d = dzn[(size_t)b * N + n];
    s1 += (double)d;
    s2 += (double)d * (double)((z[(size_t)b * N + n] - mu) * is);
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((tid & 31) == 0) { sh[0][tid >> 5] = s1; sh[1][tid >> 5] = s2; }
  __syncthreads();
  if (tid == 0) {
    const double a = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3], c = sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3];
    bc[0] = (float)a; bc[1] = (float)c;
    if (dbeta != nullptr) dbeta[n] = (float)a;
    if (dgamma != nullptr) dgamma[n] = (float)c;
  }
  __syncthreads();
  const float sdz = bc[0], sdzx = bc[1], invB = 1.0f / (float)B;
  double sb = 0.0;
  for (int b = tid; b < B; b += 128) {
    const float d = dzn[(size_t)b * N + n];
    float r;
    if (training) {
      const float xhat = (z[(size_t)b * N + n] - mu) * is;
      r = g * is * (d - sdz * invB - xhat * sdzx * invB);
    } else {
      r = g * is * d;
    }
    dzn[(size_t)b * N + n] = r;
    sb += (double)r;
  }
  sb = warp_sum(sb);
  __syncthreads();
  if ((tid & 31) == 0) sh[0][tid >> 5] = sb;
  __syncthreads();
  if (tid == 0 && dbias != nullptr) dbias[n] = (float)(sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3]);
}

// ---- BatchNorm1d with statistics synchronised across data-parallel ranks: reduce and apply as separate kernels, the caller exchanges
// the fp64 sums in between (pc_head_fwd_sync / pc_head_bwd_sync)
__global__ void __launch_bounds__(128) head_bn_sums_kernel(const float* __restrict__ z, int B, int N, double* __restrict__ sums) {
  __shared__ double sh[2][4];
  const int n = blockIdx.x, tid = threadIdx.x;
  double s = 0.0, q = 0.0;
  for (int b = tid; b < B; b += 128) { const double v = (double)z[(size_t)b * N + n]; s += v; q += v * v; }
  s = warp_sum(s); q = warp_sum(q);
  if ((tid & 31) == 0) { sh[0][tid >> 5] = s; sh[1][tid >> 5] = q; }
  __syncthreads();
  if (tid == 0) { sums[n] = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3]; sums[N + n] = sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3]; }
}

__global__ void __launch_bounds__(128) head_bn_from_sums_kernel(const double* __restrict__ sums, double count, int N, float* __restrict__ rmean,
                                                                float* __restrict__ rvar, int64_t* __restrict__ nbt, float momentum, float eps,
                                                                float* __restrict__ mean_o, float* __restrict__ invstd_o) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float mean, invstd;
  double unbiased;
  bn_train_coeffs_v(sums[n], sums[N + n], count, eps, mean, invstd, unbiased);
  mean_o[n] = mean;
  invstd_o[n] = invstd;
  if (rmean != nullptr) {
    rmean[n] = (1.f - momentum) * rmean[n] + momentum * mean;
    rvar[n] = (1.f - momentum) * rvar[n] + momentum * (float)unbiased;
  }
  if (n == 0 && nbt != nullptr) nbt[0] += 1;
}

__global__ void __launch_bounds__(128) head_bn_bwd_sums_kernel(const float* __restrict__ dzn, const float* __restrict__ z, int B, int N,
                                                               const float* __restrict__ mean, const float* __restrict__ invstd, double* __restrict__ sums) {
  __shared__ double sh[2][4];
  const int n = blockIdx.x, tid = threadIdx.x;
  const float mu = mean[n], is = invstd[n];
  double s1 = 0.0, s2 = 0.0;
  for (int b = tid; b < B; b += 128) {
    const float d = dzn[(size_t)b * N + n];
    s1 += (double)d;
    s2 += (double)d * (double)((z[(size_t)b * N + n] - mu) * is);
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if ((tid & 31) == 0) { sh[0][tid >> 5] = s1; sh[1][tid >> 5] = s2; }
  __syncthreads();
  if (tid == 0) { sums[n] = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3]; sums[N + n] = sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3]; }
}

// sums = the per-rank average of the global (sum d, sum d xhat): with the LOCAL row count B the projection terms are those of the global batch
__global__ void __launch_bounds__(128) head_bn_bwd_apply_kernel(float* __restrict__ dzn, const float* __restrict__ z, int B, int N,
                                                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                const float* __restrict__ gamma, const double* __restrict__ sums,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dbias) {
  __shared__ double sh[4];
  const int n = blockIdx.x, tid = threadIdx.x;
  const float mu = mean[n], is = invstd[n], g = gamma[n];
  const float sdz = (float)sums[n], sdzx = (float)sums[N + n], invB = 1.0f / (float)B;
  if (tid == 0) {
    if (dbeta != nullptr) dbeta[n] = sdz;
    if (dgamma != nullptr) dgamma[n] = sdzx;
  }
  double sb = 0.0;
  for (int b = tid; b < B; b += 128) {
    const float d = dzn[(size_t)b * N + n];
    const float xhat = (z[(size_t)b * N + n] - mu) * is;
    const float r = g * is * (d - sdz * invB - xhat * sdzx * invB);
    dzn[(size_t)b * N + n] = r;
    sb += (double)r;
  }
  sb = warp_sum(sb);
  if ((tid & 31) == 0) sh[tid >> 5] = sb;
  __syncthreads();
  if (tid == 0 && dbias != nullptr) dbias[n] = (float)(sh[0] + sh[1] + sh[2] + sh[3]);
}

// dW[n,k] = sum_b dz[b,n] x[b,k]
// dW[n][k] = sum_b dz[b][n] x[b][k]: block = 128 consecutive k of one n x 4 batch quarters (fixed-order smem reduce: deterministic)
__global__ void __launch_bounds__(512)
head_dw_kernel(const float* __restrict__ dz, const float* __restrict__ x, int B, int K, int N, float* __restrict__ dW) {
  __shared__ float part[4][128];
  const int kx = threadIdx.x & 127, by = threadIdx.x >> 7;
  const int kblocks = (K + 127) / 128;
  const int n = blockIdx.x / kblocks, k = (blockIdx.x - n * kblocks) * 128 + kx;
  const int bq = (B + 3) / 4, b0 = by * bq, b1 = min(B, b0 + bq);
  float s0 = 0.f, s1 = 0.f;
  if (k < K) {
    int b = b0;
    for (; b + 1 < b1; b += 2) {
      s0 = fmaf(dz[(size_t)b * N + n], x[(size_t)b * K + k], s0);
      s1 = fmaf(dz[(size_t)(b + 1) * N + n], x[(size_t)(b + 1) * K + k], s1);
    }
    if (b < b1) s0 = fmaf(dz[(size_t)b * N + n], x[(size_t)b * K + k], s0);
  }
  part[by][kx] = s0 + s1;
  __syncthreads();
  if (by == 0 && k < K) dW[(size_t)n * K + k] = (part[0][kx] + part[1][kx]) + (part[2][kx] + part[3][kx]);
}
// dx[b,k] = sum_n dz[b,n] W[n,k]
__global__ void head_dx_kernel(const float* __restrict__ dz, const float* __restrict__ W, int B, int K, int N, float* __restrict__ dx) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * K) return;
  const int k = idx % K, b = idx / K;
  float s = 0.f;
  for (int n = 0; n < N; ++n) s = fmaf(dz[(size_t)b * N + n], W[(size_t)n * K + k], s);
  dx[idx] = s;
}

}  // namespace pc

using namespace pc;

static bool attn_single_pass() {
  static const bool on = [] { const char* e = getenv("PC_ATTN_POOL1"); return e == nullptr || e[0] != '0'; }();
  return on;
}

// scratch: >= B * S * C floats, counters: >= B zeroed unsigned ints (reset by the kernel), S = pc_attn_pool_splits(B): how many blocks share
// one sample's pixels. Both may be NULL (one block per sample).
extern "C" int pc_attn_pool_splits(int B) {
  if (!attn_single_pass()) return 1;
  int s = (2 * kNumSMs) / (B > 0 ? B : 1);
  return s < 1 ? 1 : (s > 4 ? 4 : s);
}

static int attn_pool_fwd_impl(const float* a, int B, int HW, int C, const float* w, const float* b0, float* gate, float* pooled, float* scratch,
                              unsigned int* counters, pc_stream_t stream) {
  PC_REQUIRE(a && gate && pooled && B > 0 && HW > 0 && C > 0, PC_EINVAL, "pc_attn_pool_fwd: bad arguments");
  const bool al = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(pooled) | reinterpret_cast<uintptr_t>(scratch)) & 15) == 0;
  if (al && attn_single_pass() && (C == 128 || C == 256 || C == 512)) {
    const int S = (scratch != nullptr && counters != nullptr) ? pc_attn_pool_splits(B) : 1;
    const dim3 grid((unsigned)B, (unsigned)S);
    if (C == 128) attn_pool_fwd1_kernel<1><<<grid, 256, 0, stream>>>(a, HW, w, b0, gate, pooled, scratch, counters);
    else if (C == 256) attn_pool_fwd1_kernel<2><<<grid, 256, 0, stream>>>(a, HW, w, b0, gate, pooled, scratch, counters);
    else attn_pool_fwd1_kernel<4><<<grid, 256, 0, stream>>>(a, HW, w, b0, gate, pooled, scratch, counters);
    PC_LAUNCH_CHECK("attn_pool_fwd1_kernel");
    return PC_OK;
  }
  attn_pool_fwd_kernel<<<B, 256, 0, stream>>>(a, HW, C, w, b0, gate, pooled);
  PC_LAUNCH_CHECK("attn_pool_fwd_kernel");
  return PC_OK;
}

extern "C" int pc_attn_pool_fwd(const float* a, int B, int HW, int C, const float* w, const float* b0, float* gate,
                                float* pooled, pc_stream_t stream) {
  return attn_pool_fwd_impl(a, B, HW, C, w, b0, gate, pooled, nullptr, nullptr, stream);
}

extern "C" int pc_attn_pool_fwd_ws(const float* a, int B, int HW, int C, const float* w, const float* b0, float* gate, float* pooled,
                                   float* scratch, unsigned int* counters, pc_stream_t stream) {
  return attn_pool_fwd_impl(a, B, HW, C, w, b0, gate, pooled, scratch, counters, stream);
}

extern "C" int pc_attn_pool_bwd(const float* a, const float* gate, const float* dpooled, int B, int HW, int C, const float* w,
                                float* da, float* dw, float* db0, pc_stream_t stream) {
  PC_REQUIRE(a && gate && dpooled && da && B > 0 && HW > 0 && C > 0, PC_EINVAL, "pc_attn_pool_bwd: bad arguments");
  PC_REQUIRE(w == nullptr || (dw != nullptr && db0 != nullptr), PC_EINVAL, "pc_attn_pool_bwd: dw/db0 required with attention");
  PC_REQUIRE(HW <= 10240, PC_EUNSUPPORTED, "pc_attn_pool_bwd: HW=%d too large", HW);
  if (w != nullptr) {
    PC_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * C, stream));
    PC_CUDA(cudaMemsetAsync(db0, 0, sizeof(float), stream));
  }
  const bool al = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(dpooled) | reinterpret_cast<uintptr_t>(da)) & 15) == 0;
  if (al && attn_single_pass() && (C == 128 || C == 256 || C == 512)) {
    const dim3 grid((unsigned)B, (unsigned)pc_attn_pool_splits(B));      // blocks of one sample take disjoint pixels; dw / db0 are atomics already
    if (C == 128) attn_pool_bwd1_kernel<1><<<grid, 256, 0, stream>>>(a, gate, dpooled, HW, w, da, dw, db0);
    else if (C == 256) attn_pool_bwd1_kernel<2><<<grid, 256, 0, stream>>>(a, gate, dpooled, HW, w, da, dw, db0);
    else attn_pool_bwd1_kernel<4><<<grid, 256, 0, stream>>>(a, gate, dpooled, HW, w, da, dw, db0);
    PC_LAUNCH_CHECK("attn_pool_bwd1_kernel");
    return PC_OK;
  }
  attn_pool_bwd_kernel<<<B, 256, sizeof(float) * HW, stream>>>(a, gate, dpooled, HW, C, w, da, dw, db0);
  PC_LAUNCH_CHECK("attn_pool_bwd_kernel");
  return PC_OK;
}

extern "C" size_t pc_head_workspace(int B, int K, int N) {
  (void)K;
  return sizeof(float) * ((size_t)2 * B * N + 2 * (size_t)N);
}

extern "C" int pc_head_fwd(const float* x, int B, int K, int N, const float* W, const float* bias, const float* gamma,
                           const float* beta, float* running_mean, float* running_var, int64_t* num_batches_tracked,
                           float momentum, float eps, int training, float* emb, void* ws, pc_stream_t stream) {
  PC_REQUIRE(x && W && gamma && beta && emb && ws && B > 0 && K > 0 && N > 0, PC_EINVAL, "pc_head_fwd: bad arguments");
  PC_REQUIRE(!training || B > 1, PC_EINVAL, "Expected more than 1 value per channel when training");
  PC_REQUIRE(training || (running_mean && running_var), PC_EINVAL, "pc_head_fwd: eval mode needs running statistics");
  float* z = static_cast<float*>(ws);
  float* mean = z + (size_t)B * N;
  float* invstd = mean + N;
  const int vec4 = ((K & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(W)) & 15) == 0) ? 1 : 0;
  head_linear_kernel<<<ceil_div((long long)B * ((N + 3) / 4), 8), 256, 0, stream>>>(x, B, K, N, W, bias, z, vec4);
  PC_LAUNCH_CHECK("head_linear_kernel");
  head_bn_stats_kernel<<<N, 128, 0, stream>>>(z, B, N, running_mean, running_var, num_batches_tracked, momentum, eps, training, mean, invstd);
  PC_LAUNCH_CHECK("head_bn_stats_kernel");
  head_normalize_kernel<<<ceil_div(B, 8), 256, 0, stream>>>(z, B, N, mean, invstd, gamma, beta, emb);
  PC_LAUNCH_CHECK("head_normalize_kernel");
  return PC_OK;
}

extern "C" int pc_head_bwd(const float* demb, const float* x, int B, int K, int N, const float* W, const float* gamma,
                           const float* beta, int training, void* ws, float* dx, float* dW, float* dbias, float* dgamma,
                           float* dbeta, pc_stream_t stream) {
  PC_REQUIRE(demb && x && W && gamma && beta && ws && dx && dW && B > 0 && K > 0 && N > 0, PC_EINVAL, "pc_head_bwd: bad arguments");
  float* z = static_cast<float*>(ws);
  float* mean = z + (size_t)B * N;
  float* invstd = mean + N;
  float* dz = invstd + N;
  head_normalize_bwd_kernel<<<ceil_div(B, 8), 256, 0, stream>>>(demb, z, B, N, mean, invstd, gamma, beta, dz);
  PC_LAUNCH_CHECK("head_normalize_bwd_kernel");
  head_bn_bwd_kernel<<<N, 128, 0, stream>>>(dz, z, B, N, mean, invstd, gamma, training, dgamma, dbeta, dbias);
  PC_LAUNCH_CHECK("head_bn_bwd_kernel");
  head_dw_kernel<<<N * ceil_div(K, 128), 512, 0, stream>>>(dz, x, B, K, N, dW);
  PC_LAUNCH_CHECK("head_dw_kernel");
  head_dx_kernel<<<ceil_div((long long)B * K, 128), 128, 0, stream>>>(dz, W, B, K, N, dx);
  PC_LAUNCH_CHECK("head_dx_kernel");
  return PC_OK;
}

extern "C" int pc_head_fwd_sync(const float* x, int B, int K, int N, const float* W, const float* bias, const float* gamma, const float* beta,
                                float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum, float eps, float* emb, void* ws,
                                double* bn_sums, double count, int phase, pc_stream_t stream) {
  PC_REQUIRE(x && W && gamma && beta && emb && ws && bn_sums && B > 0 && K > 0 && N > 0 && (phase == 1 || phase == 2), PC_EINVAL, "pc_head_fwd_sync: bad arguments");
  float* z = static_cast<float*>(ws);
  float* mean = z + (size_t)B * N;
  float* invstd = mean + N;
  if (phase == 1) {
    const int vec4 = ((K & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(W)) & 15) == 0) ? 1 : 0;
    head_linear_kernel<<<ceil_div((long long)B * ((N + 3) / 4), 8), 256, 0, stream>>>(x, B, K, N, W, bias, z, vec4);
    PC_LAUNCH_CHECK("head_linear_kernel");
    head_bn_sums_kernel<<<N, 128, 0, stream>>>(z, B, N, bn_sums);
    PC_LAUNCH_CHECK("head_bn_sums_kernel");
    return PC_OK;
  }
  PC_REQUIRE(count > 1.0, PC_EINVAL, "Expected more than 1 value per channel when training");
  head_bn_from_sums_kernel<<<ceil_div(N, 128), 128, 0, stream>>>(bn_sums, count, N, running_mean, running_var, num_batches_tracked, momentum, eps, mean, invstd);
  PC_LAUNCH_CHECK("head_bn_from_sums_kernel");
  head_normalize_kernel<<<ceil_div(B, 8), 256, 0, stream>>>(z, B, N, mean, invstd, gamma, beta, emb);
  PC_LAUNCH_CHECK("head_normalize_kernel");
  return PC_OK;
}

extern "C" int pc_head_bwd_sync(const float* demb, const float* x, int B, int K, int N, const float* W, const float* gamma, const float* beta, void* ws,
                                float* dx, float* dW, float* dbias, float* dgamma, float* dbeta, double* bn_sums, int phase, pc_stream_t stream) {
  PC_REQUIRE(demb && x && W && gamma && beta && ws && dx && dW && bn_sums && B > 0 && K > 0 && N > 0 && (phase == 1 || phase == 2), PC_EINVAL,
             "pc_head_bwd_sync: bad arguments");
  float* z = static_cast<float*>(ws);
  float* mean = z + (size_t)B * N;
  float* invstd = mean + N;
  float* dz = invstd + N;
  if (phase == 1) {
    head_normalize_bwd_kernel<<<ceil_div(B, 8), 256, 0, stream>>>(demb, z, B, N, mean, invstd, gamma, beta, dz);
    PC_LAUNCH_CHECK("head_normalize_bwd_kernel");
    head_bn_bwd_sums_kernel<<<N, 128, 0, stream>>>(dz, z, B, N, mean, invstd, bn_sums);
    PC_LAUNCH_CHECK("head_bn_bwd_sums_kernel");
    return PC_OK;
  }
  head_bn_bwd_apply_kernel<<<N, 128, 0, stream>>>(dz, z, B, N, mean, invstd, gamma, bn_sums, dgamma, dbeta, dbias);
  PC_LAUNCH_CHECK("head_bn_bwd_apply_kernel");
  head_dw_kernel<<<N * ceil_div(K, 128), 512, 0, stream>>>(dz, x, B, K, N, dW);
  PC_LAUNCH_CHECK("head_dw_kernel");
  head_dx_kernel<<<ceil_div((long long)B * K, 128), 128, 0, stream>>>(dz, W, B, K, N, dx);
  PC_LAUNCH_CHECK("head_dx_kernel");
  return PC_OK;
}
