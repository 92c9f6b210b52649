// BatchNorm coefficient finalisation and the fused BatchNorm-apply / ReLU / max-pool / Dropout2d /
// residual-add elementwise passes, forward and backward. All tensors NHWC fp32; these kernels are
// HBM-bound: float4 channel vectors, grid-stride over pixels, per-channel reductions accumulated in
// registers -> shared memory -> one fp64 atomic per channel per CTA.
//
// Reference modules: nn.BatchNorm2d / ReLU / MaxPool2d / Dropout2d as composed in
// src/models/phoneme_cnn.py:36-63 (PhonemeNet blocks), :211-216 (init_conv), :173-184 (ResidualBlock).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pc {

// ------------------------------------------------------------------------------------------------ finalize
__global__ void bn_finalize_kernel(const double* __restrict__ stats, int C, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                                   int64_t* __restrict__ nbt, float momentum, float eps, int training,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_o,
                                   float* __restrict__ invstd_o) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && training && nbt != nullptr) nbt[0] += 1;
  if (c >= C) return;
  float mean, invstd;
  if (training) {
    double unbiased;
    bn_train_coeffs(stats, C, c, count, eps, mean, invstd, unbiased);
    if (rmean != nullptr) {
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)unbiased;
    }
  } else {
    mean = rmean[c];
    invstd = 1.0f / sqrtf(rvar[c] + eps);
  }
  const float g = gamma != nullptr ? gamma[c] : 1.f, b = beta != nullptr ? beta[c] : 0.f;
  const float s = g * invstd;
  scale[c] = s;
  shift[c] = b - mean * s;
  if (mean_o != nullptr) mean_o[c] = mean;
  if (invstd_o != nullptr) invstd_o[c] = invstd;
}

// The same finalisation INSIDE the kernel that first consumes the coefficients (pc_bn_act_split_fin, pc_bn_add_relu_fwd_fin): every
// block derives scale / shift of all C channels into shared memory (a few hundred fp64 operations), block 0 also publishes them
// (the backward reads them) and updates the running statistics. Saves one dependent launch per BatchNorm on the critical path.
__device__ __forceinline__ void bn_finalize_in_block(const PcBnFinalize& f, int C, float* __restrict__ s_scale, float* __restrict__ s_shift) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mean, invstd;
    double unbiased;
    bn_train_coeffs(f.stats, C, c, f.count, f.eps, mean, invstd, unbiased);
    const float g = f.gamma != nullptr ? f.gamma[c] : 1.f, b = f.beta != nullptr ? f.beta[c] : 0.f;
    const float sc = g * invstd, sh = b - mean * sc;
    s_scale[c] = sc;
    s_shift[c] = sh;
    if (blockIdx.x == 0) {
      if (f.running_mean != nullptr) {
        f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * mean;
        f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)unbiased;
      }
      f.scale[c] = sc;
      f.shift[c] = sh;
      if (f.mean != nullptr) f.mean[c] = mean;
      if (f.invstd != nullptr) f.invstd[c] = invstd;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && f.num_batches_tracked != nullptr) f.num_batches_tracked[0] += 1;
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 bn_relu4(float4 y, float4 s, float4 t) {
  return make_float4(fmaxf(fmaf(y.x, s.x, t.x), 0.f), fmaxf(fmaf(y.y, s.y, t.y), 0.f), fmaxf(fmaf(y.z, s.z, t.z), 0.f),
                     fmaxf(fmaf(y.w, s.w, t.w), 0.f));
}
#define PC_F4_ARR(v) {(v).x, (v).y, (v).z, (v).w}

// max |x| of a freshly written gradient tensor, accumulated with an integer atomicMax on the float's bit pattern (valid for
// non-negative floats). Consumed by the FP16X2 convolutions as their operand scale (include/phoneme_contrast.h).
__device__ __forceinline__ void amax_commit(float* amax, float local_max) {
  if (amax == nullptr) return;
  local_max = warp_max(local_max);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(amax), __float_as_uint(local_max));
}
__device__ __forceinline__ float absmax4(const float (&r)[4], float m) {
  return fmaxf(fmaxf(m, fmaxf(fabsf(r[0]), fabsf(r[1]))), fmaxf(fabsf(r[2]), fabsf(r[3])));
}

// FP16X2 activation planes are written unscaled (they follow a BatchNorm, so |a| is O(10)); a value beyond fp16's range would
// become inf in the hi plane and surface later as a NaN loss. Every plane writer therefore raises this sticky device flag when
// it meets |a| > 65504 (or a NaN); the host reads it with pc_f16_overflow_query (the trainer does so once per epoch, next to
// its loss read-back, and switches the model to the range-free tf32x3 engine).
__device__ unsigned int g_f16_overflow = 0u;
__device__ __forceinline__ void f16_range_check(float m) {
  if (!(m <= 65504.f)) atomicOr(&g_f16_overflow, 1u);
}

// optional second output of the forward kernels: the same 4 values in tensor-core operand form (fp16 hi | lo*2^11 planes of
// `plane_elems` elements each, element index o), so that the next convolution's gather can copy bytes (see pc_bn_act_split)
__device__ __forceinline__ void emit_planes4(unsigned char* __restrict__ planes, size_t plane_elems, size_t o, float4 r) {
  uint2 h, l;
  f16_range_check(fmaxf(fmaxf(fabsf(r.x), fabsf(r.y)), fmaxf(fabsf(r.z), fabsf(r.w))));
  tc::split_f16x2(r.x, r.y, h.x, l.x);
  tc::split_f16x2(r.z, r.w, h.y, l.y);
  *reinterpret_cast<uint2*>(planes + o * 2) = h;
  *reinterpret_cast<uint2*>(planes + (plane_elems + o) * 2) = l;
}

// ------------------------------------------------------------------------------------------------ forward
template <int POOL>
__global__ void __launch_bounds__(256)
bn_act_fwd_kernel(const float* __restrict__ y, int B, int H, int W, int C, int Ho, int Wo, const float* __restrict__ scale,
                  const float* __restrict__ shift, const float* __restrict__ drop, float* __restrict__ out,
                  uint8_t* __restrict__ argmax, unsigned char* __restrict__ planes, const PcBnFinalize fin) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float s_fin_fwd[];    // [2][C] when the coefficients are finalised here (pc_bn_act_fwd_fin)
  if (fin.stats != nullptr) {
    bn_finalize_in_block(fin, C, s_fin_fwd, s_fin_fwd + C);
    scale = s_fin_fwd; shift = s_fin_fwd + C;
  }
  const int C4 = C >> 2;
  const long long total = (long long)B * Ho * Wo * C4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % C4);
    long long p = idx / C4;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho);
    const int b = (int)(p / Ho);
    const int c = c4 * 4;
    const float4 s = ld4(scale + c), t = ld4(shift + c);
    float4 r;
    if (POOL == 0) {
      r = bn_relu4(ld4(y + (((size_t)b * H + ho) * W + wo) * C + c), s, t);
    } else if (POOL == 2) {
      r = make_float4(0.f, 0.f, 0.f, 0.f);  // relu output >= 0
#pragma unroll
      for (int kh = 0; kh < 2; ++kh)
#pragma unroll
        for (int kw = 0; kw < 2; ++kw) {
          const float4 a = bn_relu4(ld4(y + (((size_t)b * H + 2 * ho + kh) * W + 2 * wo + kw) * C + c), s, t);
          r.x = fmaxf(r.x, a.x); r.y = fmaxf(r.y, a.y); r.z = fmaxf(r.z, a.z); r.w = fmaxf(r.w, a.w);
        }
    } else {  // MaxPool2d(3, 2, 1): padding acts as -inf; first maximum in scan order wins (strict >)
      float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      int arg[4] = {0, 0, 0, 0};
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int h = 2 * ho - 1 + kh;
        if (h < 0 || h >= H) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int w = 2 * wo - 1 + kw;
          if (w < 0 || w >= W) continue;
          const float4 a4 = bn_relu4(ld4(y + (((size_t)b * H + h) * W + w) * C + c), s, t);
          const float a[4] = PC_F4_ARR(a4);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (a[q] > best[q]) { best[q] = a[q]; arg[q] = kh * 3 + kw; }
        }
      }
      r = make_float4(best[0], best[1], best[2], best[3]);
      if (argmax != nullptr) {
        uchar4 am = make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2], (unsigned char)arg[3]);
        *reinterpret_cast<uchar4*>(argmax + (((size_t)b * Ho + ho) * Wo + wo) * C + c) = am;
      }
    }
    if (drop != nullptr) {
      const float4 d = ld4(drop + (size_t)b * C + c);
      r.x *= d.x; r.y *= d.y; r.z *= d.z; r.w *= d.w;
    }
    const size_t oo = (((size_t)b * Ho + ho) * Wo + wo) * C + c;
    st4(out + oo, r);
    if (planes != nullptr) emit_planes4(planes, (size_t)B * Ho * Wo * C, oo, r);
  }
}

// dz (gradient w.r.t. the BatchNorm output) of input pixel p = (b,h,w), channels c..c+3, and xhat.
template <int POOL>
__device__ __forceinline__ void bn_act_dz(const float* __restrict__ dout, const float* __restrict__ y, int p, int b, int h, int w,
                                          int c, int H, int W, int C, int Ho, int Wo, float4 s, float4 t,
                                          const float* __restrict__ drop, const uint8_t* __restrict__ argmax,
                                          float dz[4], float4& yv) {
  yv = ld4(y + (size_t)p * C + c);
  const float4 a4 = bn_relu4(yv, s, t);
  const float a[4] = PC_F4_ARR(a4);
  float g[4] = {0.f, 0.f, 0.f, 0.f};
  if (POOL == 0) {
    const float4 d4 = ld4(dout + (size_t)p * C + c);
    g[0] = d4.x; g[1] = d4.y; g[2] = d4.z; g[3] = d4.w;
  } else if (POOL == 2) {
    const int ho = h >> 1, wo = w >> 1;
    if (ho < Ho && wo < Wo) {
      // recompute the window; the first maximum in scan order receives the gradient
      float best[4] = {-1.f, -1.f, -1.f, -1.f};
      int arg[4] = {0, 0, 0, 0};
#pragma unroll
      for (int kh = 0; kh < 2; ++kh)
#pragma unroll
        for (int kw = 0; kw < 2; ++kw) {
          const float4 o4 = bn_relu4(ld4(y + (((size_t)b * H + 2 * ho + kh) * W + 2 * wo + kw) * C + c), s, t);
          const float o[4] = PC_F4_ARR(o4);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (o[q] > best[q]) { best[q] = o[q]; arg[q] = kh * 2 + kw; }
        }
      const int me = (h & 1) * 2 + (w & 1);
      const float4 d4 = ld4(dout + (((size_t)b * Ho + ho) * Wo + wo) * C + c);
      const float d[4] = PC_F4_ARR(d4);
#pragma unroll
      for (int q = 0; q < 4; ++q) g[q] = (arg[q] == me) ? d[q] : 0.f;
    }
  } else {
    // MaxPool2d(3, 2, 1): input row h belongs to window row ho with tap kh = h + 1 - 2*ho in {0,1,2}:
    //   h even -> (ho = h/2, kh = 1);   h odd -> (ho = (h+1)/2, kh = 0) and (ho = (h-1)/2, kh = 2); same along w
    const int ho0 = (h + 1) >> 1, nh = 1 + (h & 1), wo0 = (w + 1) >> 1, nw = 1 + (w & 1);
#pragma unroll
    for (int ih = 0; ih < 2; ++ih) {
      const int ho = ho0 - ih;
      if (ih >= nh || ho >= Ho) continue;
      const int kh = h + 1 - 2 * ho;
#pragma unroll
      for (int iw = 0; iw < 2; ++iw) {
        const int wo = wo0 - iw;
        if (iw >= nw || wo >= Wo) continue;
        const int me = kh * 3 + (w + 1 - 2 * wo);
        const size_t o = (((size_t)b * Ho + ho) * Wo + wo) * C + c;
        const uchar4 am = *reinterpret_cast<const uchar4*>(argmax + o);
        const float4 d4 = ld4(dout + o);
        if (am.x == me) g[0] += d4.x;
        if (am.y == me) g[1] += d4.y;
        if (am.z == me) g[2] += d4.z;
        if (am.w == me) g[3] += d4.w;
      }
    }
  }
  float dr[4] = {1.f, 1.f, 1.f, 1.f};
  if (drop != nullptr) {
    const float4 d = ld4(drop + (size_t)b * C + c);
    dr[0] = d.x; dr[1] = d.y; dr[2] = d.z; dr[3] = d.w;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) dz[q] = (a[q] > 0.f) ? g[q] * dr[q] : 0.f;
}

// Block-wide maximum of a non-negative value, identical in every thread afterwards (order-independent, so every block of a
// grid derives the same number from the same global data).
__device__ __forceinline__ float block_max_all(float v, float* sh /*[9]*/) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, sh[w]);
    sh[8] = m;
  }
  __syncthreads();
  return sh[8];
}
__device__ __forceinline__ void max_commit(float* slot, float local_max) {   // slot may be null
  if (slot == nullptr) return;
  local_max = warp_max(local_max);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(slot), __float_as_uint(local_max));
}

// The same with ONE atomic per block and slot (up to three slots): per-warp atomics put 8 x 592 same-address operations per slot on
// the L2 atomic unit at the end of a reduce pass. sh: [3][8] floats. All threads of the block must call it.
template <int N>
__device__ __forceinline__ void max_commit_block(float* slots, const float (&local)[N], float* sh) {
  if (slots == nullptr) return;      // uniform across the block
  float w[N];
#pragma unroll
  for (int i = 0; i < N; ++i) w[i] = warp_max(local[i]);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) sh[i * 8 + (threadIdx.x >> 5)] = w[i];
  }
  __syncthreads();
  if (threadIdx.x < N) {
    float m = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) m = fmaxf(m, sh[threadIdx.x * 8 + k]);
    atomicMax(reinterpret_cast<unsigned int*>(slots + threadIdx.x), __float_as_uint(m));
  }
  __syncthreads();                   // sh is reused by the channel reduction that follows
}

// Visits the pixels of this block: f(p, b, h, w) with p the linear NHWC pixel index. Without pooling only p is needed (b
// only for the per-sample dropout multiplier), so the loop is flat and division-free; with pooling a block walks whole
// image rows, which keeps the (b, h) decomposition out of the inner loop.
template <int POOL, int UNROLL, typename F>
__device__ __forceinline__ void for_each_pixel(int B, int H, int W, int C4, bool need_b, F&& f) {
  const int ppb = blockDim.x / C4, lp = threadIdx.x / C4;
  if (POOL == 0) {
    const int npix = B * H * W, HW = H * W;      // < 2^31 (checked by the launcher)
#pragma unroll UNROLL
    for (int p = blockIdx.x * ppb + lp; p < npix; p += gridDim.x * ppb) f(p, need_b ? p / HW : 0, 0, 0);
  } else {
    const int rows = B * H;
    for (int row = blockIdx.x; row < rows; row += gridDim.x) {
      const int b = row / H, h = row - b * H;
#pragma unroll UNROLL
      for (int w = lp; w < W; w += ppb) f(row * W + w, b, h, w);
    }
  }
}

// MaxPool2d(2, 2) backward by WINDOW: a thread visits a 2x2 window of pre-pool pixels, loads its four y vectors once, finds each
// channel's arg max (first maximum in scan order, as bn_act_dz<2>) and hands every pixel of the window to f(p, dz, yv) with dz non-zero
// only at the arg max. The per-pixel walk (for_each_pixel<2> + bn_act_dz<2>) re-loads the whole window for each of its four pixels:
// 20 vector loads per window against 5 here. Windows along a ragged last row / column (odd H or W: no pooled output) pass dz = 0.
template <typename F>
__device__ __forceinline__ void for_each_window2(const float* __restrict__ dout, const float* __restrict__ y, int B, int H, int W, int C, int C4,
                                                 int c, int Ho, int Wo, float4 s, float4 t, const float* __restrict__ drop, F&& f) {
  const int ppb = blockDim.x / C4, lp = threadIdx.x / C4;
  const int Hw = (H + 1) >> 1, Ww = (W + 1) >> 1;
  const int per_img = Hw * Ww;
  const long long nwin = (long long)B * per_img;
#pragma unroll 1
  for (long long wi = (long long)blockIdx.x * ppb + lp; wi < nwin; wi += (long long)gridDim.x * ppb) {
    const int b = (int)(wi / per_img), r = (int)(wi - (long long)b * per_img);
    const int hw = r / Ww, ww = r - hw * Ww;
    const int h0 = 2 * hw, w0 = 2 * ww;
    float4 yv[4];
    float a[4][4];
    bool in[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int h = h0 + (k >> 1), w = w0 + (k & 1);
      in[k] = h < H && w < W;
      yv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in[k]) yv[k] = ld4(y + (((size_t)b * H + h) * W + w) * C + c);
      const float4 a4 = bn_relu4(yv[k], s, t);
      a[k][0] = a4.x; a[k][1] = a4.y; a[k][2] = a4.z; a[k][3] = a4.w;
    }
    float g[4] = {0.f, 0.f, 0.f, 0.f};     // gradient reaching each channel's arg-max pixel
    int arg[4] = {0, 0, 0, 0};
    if (hw < Ho && ww < Wo) {          // a full window (all four pixels exist) with a pooled output
      const float4 d4 = ld4(dout + (((size_t)b * Ho + hw) * Wo + ww) * C + c);
      const float d[4] = PC_F4_ARR(d4);
      float dr[4] = {1.f, 1.f, 1.f, 1.f};
      if (drop != nullptr) {
        const float4 dd = ld4(drop + (size_t)b * C + c);
        dr[0] = dd.x; dr[1] = dd.y; dr[2] = dd.z; dr[3] = dd.w;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float best = -1.f;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (a[k][q] > best) { best = a[k][q]; arg[q] = k; }
        g[q] = best > 0.f ? d[q] * dr[q] : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!in[k]) continue;
      const float dz[4] = {arg[0] == k ? g[0] : 0.f, arg[1] == k ? g[1] : 0.f, arg[2] == k ? g[2] : 0.f, arg[3] == k ? g[3] : 0.f};
      f((int)((((size_t)b * H + h0 + (k >> 1)) * W + w0 + (k & 1))), dz, yv[k]);
    }
  }
}

// Block-level per-channel reduction of NV values per thread (thread owns channels c4*4..+3), then fp64 atomics.
template <int NV>
__device__ __forceinline__ void block_channel_reduce(float (&acc)[NV][4], int C4, int c4_off, double* const* dst, float* sh) {
  // sh: [256][4] floats; C4 = channel quads handled by this block, starting at quad c4_off
  const int tid = threadIdx.x;
  const int c4 = tid % C4;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) sh[tid * 4 + q] = acc[v][q];
    __syncthreads();
    if (tid < C4) {
      double s[4] = {0.0, 0.0, 0.0, 0.0};
      for (int t = tid; t < (int)blockDim.x; t += C4)
#pragma unroll
        for (int q = 0; q < 4; ++q) s[q] += (double)sh[t * 4 + q];
#pragma unroll
      for (int q = 0; q < 4; ++q) atomicAdd(dst[v] + (c4_off + c4) * 4 + q, s[q]);
    }
  }
}

template <int POOL, int MINB>
__global__ void __launch_bounds__(256, MINB)
bn_act_bwd_reduce_kernel(const float* __restrict__ dout, const float* __restrict__ y, int B, int H, int W, int C, int Ho,
                         int Wo, const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ drop,
                         const uint8_t* __restrict__ argmax, double* __restrict__ sums, float* __restrict__ maxes) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[256 * 4];
  // a block covers a tile of C4 channel quads (gridDim.y tiles) so that the closing fp64 atomics per block stay few
  const int C4 = (C >> 2) / gridDim.y, c4_off = blockIdx.y * C4;
  const int c4 = c4_off + threadIdx.x % C4, c = c4 * 4;
  const float4 s = ld4(scale + c), t = ld4(shift + c), mu = ld4(mean + c), is = ld4(invstd + c);
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float gmax = 0.f, xmax = 0.f;     // max |dz| and max |xhat| (bound of |dy| for the pre-split gradient planes)
  auto accumulate = [&](const float (&dz)[4], const float4& yv) {
    const float xh[4] = {(yv.x - mu.x) * is.x, (yv.y - mu.y) * is.y, (yv.z - mu.z) * is.z, (yv.w - mu.w) * is.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      acc[0][q] += dz[q];
      acc[1][q] += dz[q] * xh[q];
    }
    gmax = absmax4(dz, gmax);
    xmax = absmax4(xh, xmax);
  };
  if (POOL == 2) {
    for_each_window2(dout, y, B, H, W, C, C4, c, Ho, Wo, s, t, drop, [&](int, const float (&dz)[4], const float4& yv) { accumulate(dz, yv); });
  } else {
    for_each_pixel<POOL, 4>(B, H, W, C4, drop != nullptr, [&](int p, int b, int h, int w) {
      float dz[4];
      float4 yv;
      bn_act_dz<POOL>(dout, y, p, b, h, w, c, H, W, C, Ho, Wo, s, t, drop, argmax, dz, yv);
      accumulate(dz, yv);
    });
  }
  {
    const float mx[2] = {gmax, xmax};
    max_commit_block<2>(maxes, mx, sh);
  }
  double* dst[2] = {sums, sums + C};
  block_channel_reduce<2>(acc, C4, c4_off, dst, sh);
}

// Gradient of the bias of the convolution that produced y (= sum over pixels of dy) WITHOUT a pass over dy. With the per-channel
// constants the apply kernels use -- a = fl(sum_dz / M), b = fl(sum_dz_xhat / M), mean and invstd in fp32 --
//     sum_p dy = scale * ( (sum_dz - M a) - b * sum_p xhat ),     sum_p xhat = invstd * (sum_p y - M mean),
// i.e. the two residuals left by rounding the batch mean and sum_dz / M to fp32: the bias is absorbed by the batch mean, its
// gradient is analytically zero and the reference reports fp32 round-off of this size there. sum_p y is the fp64 statistic the
// forward pass accumulated (`y_stats`). Replaces a column-sum kernel over every gradient tensor (125 us per cnn_deep step).
__device__ __forceinline__ float bias_grad_closed_form(float scale, float mean, float invstd, double sum_dz, float a, float b, double sum_y,
                                                       double M) {
  const double sum_xhat = (double)invstd * (sum_y - M * (double)mean);
  return (float)((double)scale * ((sum_dz - M * (double)a) - (double)b * sum_xhat));
}

template <int POOL>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ y, int B, int H, int W, int C, int Ho,
                        int Wo, const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ drop,
                        const uint8_t* __restrict__ argmax, const double* __restrict__ sums, float* __restrict__ dy,
                        float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dy_amax,
                        const float* __restrict__ maxes, unsigned char* __restrict__ dy_planes,
                        const double* __restrict__ y_stats, float* __restrict__ db_conv) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh_max[9];
  const int C4 = C >> 2;
  const int c4 = threadIdx.x % C4, c = c4 * 4;
  const int ppb = blockDim.x / C4;
  const int npix = B * H * W;
  const float invM = 1.0f / (float)npix;
  const float4 s = ld4(scale + c), t = ld4(shift + c), mu = ld4(mean + c), is = ld4(invstd + c);
  float sdz[4], sdzx[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    sdz[q] = (float)sums[c + q];
    sdzx[q] = (float)sums[C + c + q];
  }
  if (blockIdx.x == 0 && threadIdx.x < C4) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (dgamma != nullptr) dgamma[c + q] = sdzx[q];
      if (dbeta != nullptr) dbeta[c + q] = sdz[q];
    }
  }
  const float sv[4] = PC_F4_ARR(s), muv[4] = PC_F4_ARR(mu), isv[4] = PC_F4_ARR(is);
  if (blockIdx.x == 0 && threadIdx.x < C4 && db_conv != nullptr) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      db_conv[c + q] = bias_grad_closed_form(sv[q], muv[q], isv[q], sums[c + q], sdz[q] * invM, sdzx[q] * invM, y_stats[c + q], (double)npix);
  }
  // Pre-split output: dy is (also) written as fp16 hi | lo planes scaled by a power of two. The scale must be known before
  // the first element is written, so it comes from a BOUND of |dy| built from the reduce pass (max |dz|, max |xhat|, the
  // per-channel sums): |dy| <= |scale_c| (max|dz| + |sum dz_c| / M + max|xhat| |sum dz xhat_c| / M). Every block derives
  // the same bound; block 0 publishes it in dy_amax, which the consuming convolutions use to undo the scale.
  float pscale = 1.f;
  if (dy_planes != nullptr) {
    const float gmax = maxes[0], xmax = maxes[1];
    float lb = 0.f;
    for (int cc = threadIdx.x; cc < C; cc += blockDim.x)
      lb = fmaxf(lb, fabsf(scale[cc]) * (gmax + fabsf((float)sums[cc]) * invM + xmax * fabsf((float)sums[C + cc]) * invM));
    const float bound = block_max_all(lb, sh_max);
    pscale = tc::f16_operand_scale(bound);
    if (blockIdx.x == 0 && threadIdx.x == 0 && dy_amax != nullptr) dy_amax[0] = bound;
  }
  const size_t plane_elems = (size_t)npix * C;
  float lmax = 0.f;
  auto emit = [&](int p, const float (&dz)[4], const float4& yv) {
    const float yy[4] = PC_F4_ARR(yv);
    float r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float xhat = (yy[q] - muv[q]) * isv[q];
      r[q] = sv[q] * (dz[q] - sdz[q] * invM - xhat * sdzx[q] * invM);
    }
    lmax = absmax4(r, lmax);
    if (dy != nullptr) st4(dy + (size_t)p * C + c, make_float4(r[0], r[1], r[2], r[3]));
    if (dy_planes != nullptr)
      emit_planes4(dy_planes, plane_elems, (size_t)p * C + c, make_float4(r[0] * pscale, r[1] * pscale, r[2] * pscale, r[3] * pscale));
  };
  if (POOL == 2) {
    for_each_window2(dout, y, B, H, W, C, C4, c, Ho, Wo, s, t, drop, emit);
  } else {
    for_each_pixel<POOL, 2>(B, H, W, C4, drop != nullptr, [&](int p, int b, int h, int w) {
      float dz[4];
      float4 yv;
      bn_act_dz<POOL>(dout, y, p, b, h, w, c, H, W, C, Ho, Wo, s, t, drop, argmax, dz, yv);
      emit(p, dz, yv);
    });
  }
  if (dy_planes == nullptr) amax_commit(dy_amax, lmax);
}

// ------------------------------------------------------------------------------------------------ residual tail
__global__ void __launch_bounds__(256)
bn_add_relu_fwd_kernel(const float* __restrict__ y2, const float* __restrict__ scale2, const float* __restrict__ shift2,
                       const float* __restrict__ ysc, const float* __restrict__ sc_scale, const float* __restrict__ sc_shift,
                       long long n_pix, int C, float* __restrict__ out, unsigned char* __restrict__ planes, const PcBnFinalize fin2,
                       const PcBnFinalize fin_s) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float s_fin[];       // [4][C] when the coefficients are finalised here
  if (fin2.stats != nullptr) {
    bn_finalize_in_block(fin2, C, s_fin, s_fin + C);
    scale2 = s_fin; shift2 = s_fin + C;
    if (fin_s.stats != nullptr) {
      bn_finalize_in_block(fin_s, C, s_fin + 2 * C, s_fin + 3 * C);
      sc_scale = s_fin + 2 * C; sc_shift = s_fin + 3 * C;
    }
  }
  const int C4 = C >> 2;
  const long long total = n_pix * C4;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C4) * 4;
    const size_t o = (size_t)(idx / C4) * C + c;
    const float4 a = ld4(y2 + o), s = ld4(scale2 + c), t = ld4(shift2 + c);
    float4 r = ld4(ysc + o);
    if (sc_scale != nullptr) {
      const float4 ss = ld4(sc_scale + c), ts = ld4(sc_shift + c);
      r = make_float4(fmaf(r.x, ss.x, ts.x), fmaf(r.y, ss.y, ts.y), fmaf(r.z, ss.z, ts.z), fmaf(r.w, ss.w, ts.w));
    }
    r.x = fmaxf(fmaf(a.x, s.x, t.x) + r.x, 0.f);
    r.y = fmaxf(fmaf(a.y, s.y, t.y) + r.y, 0.f);
    r.z = fmaxf(fmaf(a.z, s.z, t.z) + r.z, 0.f);
    r.w = fmaxf(fmaf(a.w, s.w, t.w) + r.w, 0.f);
    st4(out + o, r);
    if (planes != nullptr) emit_planes4(planes, (size_t)n_pix * C, o, r);
  }
}

template <int MINB>
__global__ void __launch_bounds__(256, MINB)   // MINB blocks per SM = the grid's cap (reduce_grid2): the grid must be whole waves of what fits
bn_add_relu_bwd_reduce_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ y2,
                              const float* __restrict__ mean2, const float* __restrict__ invstd2,
                              const float* __restrict__ ysc, const float* __restrict__ mean_s,
                              const float* __restrict__ invstd_s, long long n_pix, int C, double* __restrict__ sums2,
                              double* __restrict__ sums_s, float* __restrict__ maxes) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[256 * 4];
  const int C4 = (C >> 2) / gridDim.y, c4_off = blockIdx.y * C4;     // channel tile of this block (see bn_act_bwd_reduce_kernel)
  const int c4 = c4_off + threadIdx.x % C4, c = c4 * 4;
  const int ppb = blockDim.x / C4;
  const float4 mu2 = ld4(mean2 + c), is2 = ld4(invstd2 + c);
  const bool proj = sums_s != nullptr;
  float4 mus = make_float4(0.f, 0.f, 0.f, 0.f), iss = mus;
  if (proj) { mus = ld4(mean_s + c); iss = ld4(invstd_s + c); }
  float acc[3][4] = {};
  float gmax = 0.f, x2max = 0.f, xsmax = 0.f;      // max |g|, |xhat2|, |xhat_s|: bound of |dy| for pre-split gradient planes
#pragma unroll 2
  for (long long p = (long long)blockIdx.x * ppb + threadIdx.x / C4; p < n_pix; p += (long long)gridDim.x * ppb) {
    const size_t o = (size_t)p * C + c;
    const float4 d = ld4(dout + o), ov = ld4(out + o), yv = ld4(y2 + o);
    const float g[4] = {ov.x > 0.f ? d.x : 0.f, ov.y > 0.f ? d.y : 0.f, ov.z > 0.f ? d.z : 0.f, ov.w > 0.f ? d.w : 0.f};
    const float x2[4] = {(yv.x - mu2.x) * is2.x, (yv.y - mu2.y) * is2.y, (yv.z - mu2.z) * is2.z, (yv.w - mu2.w) * is2.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      acc[0][q] += g[q];
      acc[1][q] += g[q] * x2[q];
    }
    gmax = absmax4(g, gmax);
    x2max = absmax4(x2, x2max);
    if (proj) {
      const float4 sv = ld4(ysc + o);
      const float xs[4] = {(sv.x - mus.x) * iss.x, (sv.y - mus.y) * iss.y, (sv.z - mus.z) * iss.z, (sv.w - mus.w) * iss.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[2][q] += g[q] * xs[q];
      xsmax = absmax4(xs, xsmax);
    }
  }
  {
    const float mx[3] = {gmax, x2max, xsmax};
    max_commit_block<3>(maxes, mx, sh);
  }
  // sums2 = [sum g, sum g*xhat2]; sums_s = [sum g, sum g*xhat_s]
  if (proj) {
    double* dst[3] = {sums2, sums2 + C, sums_s + C};
    block_channel_reduce<3>(acc, C4, c4_off, dst, sh);
  } else {
    float acc2[2][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { acc2[0][q] = acc[0][q]; acc2[1][q] = acc[1][q]; }
    double* dst[2] = {sums2, sums2 + C};
    block_channel_reduce<2>(acc2, C4, c4_off, dst, sh);
  }
}

__global__ void __launch_bounds__(256, 4)      // 4 resident blocks per SM: the 16-per-SM grid cap (ew_grid) is then exactly 4 waves
bn_add_relu_bwd_apply_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ y2,
                             const float* __restrict__ scale2, const float* __restrict__ mean2,
                             const float* __restrict__ invstd2, const double* __restrict__ sums2,
                             const float* __restrict__ ysc, const float* __restrict__ sc_scale,
                             const float* __restrict__ mean_s, const float* __restrict__ invstd_s,
                             const double* __restrict__ sums_s, long long n_pix, int C, float* __restrict__ dy2,
                             float* __restrict__ dysc, float* __restrict__ dgamma2, float* __restrict__ dbeta2,
                             float* __restrict__ dgamma_s, float* __restrict__ dbeta_s, float* __restrict__ dy2_amax,
                             float* __restrict__ dysc_amax, const float* __restrict__ maxes,
                             unsigned char* __restrict__ dy2_planes, unsigned char* __restrict__ dysc_planes,
                             const double* __restrict__ y2_stats, float* __restrict__ db2, const double* __restrict__ ysc_stats,
                             float* __restrict__ db_s) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh_max[9];
  const int C4 = C >> 2;
  const int c4 = threadIdx.x % C4, c = c4 * 4;
  const int ppb = blockDim.x / C4;
  const float invM = 1.0f / (float)n_pix;
  const bool proj = sc_scale != nullptr;
  const float4 s2 = ld4(scale2 + c), mu2 = ld4(mean2 + c), is2 = ld4(invstd2 + c);
  float4 ss = make_float4(1.f, 1.f, 1.f, 1.f), mus = make_float4(0.f, 0.f, 0.f, 0.f), iss = mus;
  if (proj) { ss = ld4(sc_scale + c); mus = ld4(mean_s + c); iss = ld4(invstd_s + c); }
  float sg[4], sgx2[4], sgxs[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    sg[q] = (float)sums2[c + q];
    sgx2[q] = (float)sums2[C + c + q];
    sgxs[q] = proj ? (float)sums_s[C + c + q] : 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.x < C4) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (dgamma2 != nullptr) dgamma2[c + q] = sgx2[q];
      if (dbeta2 != nullptr) dbeta2[c + q] = sg[q];
      if (proj && dgamma_s != nullptr) dgamma_s[c + q] = sgxs[q];
      if (proj && dbeta_s != nullptr) dbeta_s[c + q] = sg[q];
    }
  }
  const float s2v[4] = PC_F4_ARR(s2), mu2v[4] = PC_F4_ARR(mu2), is2v[4] = PC_F4_ARR(is2);
  const float ssv[4] = PC_F4_ARR(ss), musv[4] = PC_F4_ARR(mus), issv[4] = PC_F4_ARR(iss);
  if (blockIdx.x == 0 && threadIdx.x < C4) {        // conv bias gradients in closed form (see bias_grad_closed_form)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (db2 != nullptr)
        db2[c + q] = bias_grad_closed_form(s2v[q], mu2v[q], is2v[q], sums2[c + q], sg[q] * invM, sgx2[q] * invM, y2_stats[c + q], (double)n_pix);
      if (proj && db_s != nullptr)
        db_s[c + q] = bias_grad_closed_form(ssv[q], musv[q], issv[q], sums2[c + q], sg[q] * invM, sgxs[q] * invM, ysc_stats[c + q], (double)n_pix);
    }
  }
  // pre-split outputs: same scheme as bn_act_bwd_apply_kernel (power-of-two scale from a bound of |dy|, published in *_amax)
  float pscale2 = 1.f, pscales = 1.f;
  if (dy2_planes != nullptr || dysc_planes != nullptr) {
    const float gmax = maxes[0], x2max = maxes[1], xsmax = maxes[2];
    float lb2 = 0.f, lbs = 0.f;
    for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
      const float sgc = fabsf((float)sums2[cc]) * invM;
      lb2 = fmaxf(lb2, fabsf(scale2[cc]) * (gmax + sgc + x2max * fabsf((float)sums2[C + cc]) * invM));
      if (proj) lbs = fmaxf(lbs, fabsf(sc_scale[cc]) * (gmax + sgc + xsmax * fabsf((float)sums_s[C + cc]) * invM));
    }
    const float b2 = block_max_all(lb2, sh_max);
    const float bs = proj ? block_max_all(lbs, sh_max) : gmax;
    pscale2 = tc::f16_operand_scale(b2);
    pscales = tc::f16_operand_scale(bs);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      if (dy2_planes != nullptr && dy2_amax != nullptr) dy2_amax[0] = b2;
      if (dysc_planes != nullptr && dysc_amax != nullptr) dysc_amax[0] = bs;
    }
  }
  const size_t plane_elems = (size_t)n_pix * C;
  float lmax2 = 0.f, lmaxs = 0.f;
  for (long long p = (long long)blockIdx.x * ppb + threadIdx.x / C4; p < n_pix; p += (long long)gridDim.x * ppb) {
    const size_t o = (size_t)p * C + c;
    const float4 d = ld4(dout + o), ov = ld4(out + o), yv = ld4(y2 + o);
    const float g[4] = {ov.x > 0.f ? d.x : 0.f, ov.y > 0.f ? d.y : 0.f, ov.z > 0.f ? d.z : 0.f, ov.w > 0.f ? d.w : 0.f};
    const float yy[4] = PC_F4_ARR(yv);
    float r2[4], rs[4];
    float scv[4] = {0.f, 0.f, 0.f, 0.f};
    if (proj) {
      const float4 sv = ld4(ysc + o);
      scv[0] = sv.x; scv[1] = sv.y; scv[2] = sv.z; scv[3] = sv.w;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float x2 = (yy[q] - mu2v[q]) * is2v[q];
      r2[q] = s2v[q] * (g[q] - sg[q] * invM - x2 * sgx2[q] * invM);
      if (proj) {
        const float xs = (scv[q] - musv[q]) * issv[q];
        rs[q] = ssv[q] * (g[q] - sg[q] * invM - xs * sgxs[q] * invM);
      } else {
        rs[q] = g[q];
      }
    }
    lmax2 = absmax4(r2, lmax2);
    lmaxs = absmax4(rs, lmaxs);
    if (dy2 != nullptr) st4(dy2 + o, make_float4(r2[0], r2[1], r2[2], r2[3]));
    if (dysc != nullptr) st4(dysc + o, make_float4(rs[0], rs[1], rs[2], rs[3]));
    if (dy2_planes != nullptr) emit_planes4(dy2_planes, plane_elems, o, make_float4(r2[0] * pscale2, r2[1] * pscale2, r2[2] * pscale2, r2[3] * pscale2));
    if (dysc_planes != nullptr) emit_planes4(dysc_planes, plane_elems, o, make_float4(rs[0] * pscales, rs[1] * pscales, rs[2] * pscales, rs[3] * pscales));
  }
  if (dy2_planes == nullptr) amax_commit(dy2_amax, lmax2);
  if (dysc_planes == nullptr) amax_commit(dysc_amax, lmaxs);
}

// ------------------------------------------------------------------------------------------------ pre-split activation
// a = drop * relu(scale*y + shift) -> fp16 hi plane | fp16 lo*2^11 plane (tensor-core operand form, see tc_common.cuh).
// thread = 8 consecutive channels of one pixel: two float4 loads, two 16-byte stores.
__global__ void __launch_bounds__(256)
bn_act_split_kernel(const float* __restrict__ y, long long n_pix, int C, int hw, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ drop, int relu, unsigned char* __restrict__ planes,
                    const PcBnFinalize fin) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float s_fin[];       // [2][C] when the coefficients are finalised here
  if (fin.stats != nullptr) {
    bn_finalize_in_block(fin, C, s_fin, s_fin + C);
    scale = s_fin; shift = s_fin + C;
  }
  const int C8 = C >> 3;
  const long long total = n_pix * C8;
  const size_t plane_bytes = (size_t)n_pix * C * 2;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C8) * 8;
    const long long pix = idx / C8;
    const size_t o = (size_t)pix * C + c;
    float4 a = ld4(y + o), b = ld4(y + o + 4);
    if (scale != nullptr) {
      const float4 s0 = ld4(scale + c), s1 = ld4(scale + c + 4), t0 = ld4(shift + c), t1 = ld4(shift + c + 4);
      a = make_float4(fmaf(a.x, s0.x, t0.x), fmaf(a.y, s0.y, t0.y), fmaf(a.z, s0.z, t0.z), fmaf(a.w, s0.w, t0.w));
      b = make_float4(fmaf(b.x, s1.x, t1.x), fmaf(b.y, s1.y, t1.y), fmaf(b.z, s1.z, t1.z), fmaf(b.w, s1.w, t1.w));
    }
    if (relu) {
      a = make_float4(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f), fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
      b = make_float4(fmaxf(b.x, 0.f), fmaxf(b.y, 0.f), fmaxf(b.z, 0.f), fmaxf(b.w, 0.f));
    }
    if (drop != nullptr) {
      const size_t od = (size_t)(pix / hw) * C + c;
      const float4 d0 = ld4(drop + od), d1 = ld4(drop + od + 4);
      a = make_float4(a.x * d0.x, a.y * d0.y, a.z * d0.z, a.w * d0.w);
      b = make_float4(b.x * d1.x, b.y * d1.y, b.z * d1.z, b.w * d1.w);
    }
    uint4 h, l;
    f16_range_check(fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                          fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)))));
    tc::split_f16x2(a.x, a.y, h.x, l.x);
    tc::split_f16x2(a.z, a.w, h.y, l.y);
    tc::split_f16x2(b.x, b.y, h.z, l.z);
    tc::split_f16x2(b.z, b.w, h.w, l.w);
    *reinterpret_cast<uint4*>(planes + o * 2) = h;
    *reinterpret_cast<uint4*>(planes + plane_bytes + o * 2) = l;
  }
}

// ------------------------------------------------------------------------------------------------ dropout mask
__global__ void dropout2d_mask_kernel(float* __restrict__ drop, int n, float p, float keep_scale, uint64_t seed, uint64_t offset,
                                      const long long* __restrict__ step_dev) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i * 4 >= n) return;
  if (step_dev != nullptr) offset += (uint64_t)step_dev[0] << 24;      // graph-capturable: the call counter lives on the device
  const uint64_t ctr = offset + (uint64_t)i;
  const uint4 r = Philox::round10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0x44524f50u, 0u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (i * 4 + q < n) drop[i * 4 + q] = (Philox::u01(rv[q]) >= p) ? keep_scale : 0.f;
}

// reduction passes end with one fp64 atomic per channel per block onto the SAME few addresses (they serialise in L2), so
// they run on at most 4 blocks per SM
static inline int reduce_grid(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)kNumSMs * 4;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}
// 2-D grid of a reduction pass: y = channel tiles of 64 channels (each block then ends with 64 atomics per sum instead of C),
// x * y capped at 4 blocks per SM
static inline int reduce_bps() {
  static const int bps = [] { const char* e = getenv("PC_BN_REDUCE_BPS"); const int v = e ? atoi(e) : 4; return v >= 1 && v <= 4 ? v : 4; }();
  return bps;
}
static inline dim3 reduce_grid2(long long n_pix, int C) {
  const int c4 = C / 4;
  const int tiles = (c4 % 16 == 0) ? c4 / 16 : 1;
  const int c4t = c4 / tiles;
  const int ppb = 256 / c4t;
  long long gx = (n_pix + (long long)ppb * 4 - 1) / ((long long)ppb * 4);
  const long long cap = (long long)kNumSMs * reduce_bps() / tiles;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)tiles);
}

static inline int ew_grid(long long work_items, int per_block) {
  long long g = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)kNumSMs * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// Grid of a grid-stride elementwise kernel in WHOLE waves of its resident blocks (occupancy queried once per kernel and shared-memory
// size): ew_grid's fixed 16-blocks-per-SM cap left the last wave a third full for kernels that fit 3, 5 or 6 blocks per SM.
template <typename K>
static int ew_grid_waves(K kernel, size_t smem, long long work_items, int per_block) {
  struct Entry { const void* fn; size_t smem; int occ; };
  static Entry cache[32];
  static int n_cache = 0;
  int occ = 0;
  for (int i = 0; i < n_cache; ++i)
    if (cache[i].fn == (const void*)kernel && cache[i].smem == smem) occ = cache[i].occ;
  if (occ == 0) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 256, smem) != cudaSuccess || occ < 1) occ = 4;
    if (n_cache < 32) cache[n_cache++] = Entry{(const void*)kernel, smem, occ};
  }
  const long long wave = (long long)kNumSMs * occ;
  long long g = (work_items + per_block - 1) / per_block;
  if (g > wave) {
    long long waves = g / wave;
    if (waves > 4) waves = 4;
    g = waves * wave;
  }
  return (int)(g < 1 ? 1 : g);
}

static inline void pool_out_dims(int H, int W, int pool, int* Ho, int* Wo) {
  if (pool == 0) { *Ho = H; *Wo = W; }
  else if (pool == 2) { *Ho = H / 2; *Wo = W / 2; }
  else { *Ho = (H + 2 - 3) / 2 + 1; *Wo = (W + 2 - 3) / 2 + 1; }
}

}  // namespace pc

using namespace pc;

static int check_fin(const char* fn, const PcBnFinalize* f) {
  PC_REQUIRE(f->stats && f->count > 0.0 && f->scale && f->shift, PC_EINVAL, "%s: PcBnFinalize needs stats, a positive count and scale / shift outputs", fn);
  PC_REQUIRE((f->running_mean == nullptr) == (f->running_var == nullptr), PC_EINVAL, "%s: running_mean / running_var mismatch", fn);
  return PC_OK;
}

extern "C" int pc_bn_finalize(const double* stats, int C, double count, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                              float eps, int training, float* scale, float* shift, float* mean, float* invstd,
                              pc_stream_t stream) {
  PC_REQUIRE(C > 0 && scale && shift, PC_EINVAL, "pc_bn_finalize: bad arguments");
  PC_REQUIRE(training ? (stats != nullptr && count >= 1.0) : (running_mean && running_var), PC_EINVAL,
             "pc_bn_finalize: missing statistics");
  // nn.BatchNorm raises for a single value per channel in training mode (tests/test_models.py:52,97-103 avoid it)
  PC_REQUIRE(!training || count > 1.0, PC_EINVAL, "Expected more than 1 value per channel when training");
  launch_pdl(bn_finalize_kernel, dim3(ceil_div(C, 128)), dim3(128), 0, stream, stats, C, count, gamma, beta, running_mean, running_var,
                                                            num_batches_tracked, momentum, eps, training, scale, shift,
                                                            mean, invstd);
  PC_LAUNCH_CHECK("bn_finalize_kernel");
  return PC_OK;
}

#define PC_CHECK_C4(fn, C) \
  PC_REQUIRE((C) > 0 && (C) % 4 == 0 && (C) <= 1024 && 256 % ((C) / 4) == 0, PC_EUNSUPPORTED, fn ": channels=%d must be 4*2^k <= 1024", (C))

static int bn_act_fwd_impl(const float* y, int B, int H, int W, int C, const float* scale, const float* shift, const PcBnFinalize* fin,
                           const float* drop, int pool, float* out, uint8_t* argmax, void* planes, pc_stream_t stream) {
  PC_CHECK_C4("pc_bn_act_fwd", C);
  PC_REQUIRE(pool == 0 || pool == 2 || pool == 3, PC_EINVAL, "pc_bn_act_fwd: pool must be 0, 2 or 3");
  int Ho, Wo;
  pool_out_dims(H, W, pool, &Ho, &Wo);
  PC_REQUIRE(Ho > 0 && Wo > 0, PC_EINVAL, "pc_bn_act_fwd: input %dx%d too small for pooling", H, W);
  const long long total = (long long)B * Ho * Wo * (C / 4);
  PcBnFinalize f{};
  size_t smem = 0;
  if (fin != nullptr) { f = *fin; smem = sizeof(float) * 2 * (size_t)C; }
  const int grid = pool == 0 ? ew_grid_waves(bn_act_fwd_kernel<0>, smem, total, 256)
                             : (pool == 2 ? ew_grid_waves(bn_act_fwd_kernel<2>, smem, total, 256) : ew_grid_waves(bn_act_fwd_kernel<3>, smem, total, 256));
  if (pool == 0) launch_pdl((bn_act_fwd_kernel<0>), dim3(grid), dim3(256), smem, stream, y, B, H, W, C, Ho, Wo, scale, shift, drop, out, argmax, static_cast<unsigned char*>(planes), f);
  else if (pool == 2) launch_pdl((bn_act_fwd_kernel<2>), dim3(grid), dim3(256), smem, stream, y, B, H, W, C, Ho, Wo, scale, shift, drop, out, argmax, static_cast<unsigned char*>(planes), f);
  else launch_pdl((bn_act_fwd_kernel<3>), dim3(grid), dim3(256), smem, stream, y, B, H, W, C, Ho, Wo, scale, shift, drop, out, argmax, static_cast<unsigned char*>(planes), f);
  PC_LAUNCH_CHECK("bn_act_fwd_kernel");
  return PC_OK;
}

extern "C" int pc_bn_act_fwd(const float* y, int B, int H, int W, int C, const float* scale, const float* shift,
                             const float* drop, int pool, float* out, uint8_t* argmax, void* planes, pc_stream_t stream) {
  PC_REQUIRE(y && scale && shift && out && B > 0 && H > 0 && W > 0, PC_EINVAL, "pc_bn_act_fwd: bad arguments");
  return bn_act_fwd_impl(y, B, H, W, C, scale, shift, nullptr, drop, pool, out, argmax, planes, stream);
}

// pc_bn_act_fwd with the train-mode BatchNorm coefficients finalised inside the kernel (PcBnFinalize): one dependent launch less
extern "C" int pc_bn_act_fwd_fin(const float* y, int B, int H, int W, int C, const PcBnFinalize* fin, const float* drop, int pool, float* out,
                                 uint8_t* argmax, void* planes, pc_stream_t stream) {
  PC_REQUIRE(y && fin && fin->stats && fin->scale && fin->shift && out && B > 0 && H > 0 && W > 0 && fin->count > 0, PC_EINVAL, "pc_bn_act_fwd_fin: bad arguments");
  PC_REQUIRE(C <= 4096, PC_EUNSUPPORTED, "pc_bn_act_fwd_fin: too many channels for the in-block finalisation");
  return bn_act_fwd_impl(y, B, H, W, C, nullptr, nullptr, fin, drop, pool, out, argmax, planes, stream);
}

extern "C" int pc_bn_act_bwd_reduce(const float* dout, const float* y, int B, int H, int W, int C, const float* scale,
                                    const float* shift, const float* mean, const float* invstd, const float* drop,
                                    int pool, const uint8_t* argmax, double* sums, float* maxes, pc_stream_t stream) {
  PC_REQUIRE(dout && y && scale && shift && mean && invstd && sums, PC_EINVAL, "pc_bn_act_bwd_reduce: null pointer");
  PC_CHECK_C4("pc_bn_act_bwd_reduce", C);
  PC_REQUIRE(pool == 0 || pool == 2 || (pool == 3 && argmax), PC_EINVAL, "pc_bn_act_bwd_reduce: bad pool / argmax");
  int Ho, Wo;
  pool_out_dims(H, W, pool, &Ho, &Wo);
  const long long items = (long long)B * H * W * (C / 4);
  PC_REQUIRE((long long)B * H * W < (1LL << 31), PC_EUNSUPPORTED, "pc_bn_act_bwd_reduce: too many pixels");
  (void)items;
  const dim3 grid = reduce_grid2((long long)B * H * W, C);
  if (pool == 0) { if (reduce_bps() == 3) launch_pdl((bn_act_bwd_reduce_kernel<0, 3>), dim3(grid), dim3(256), 0, stream, dout, y, B, H, W, C, Ho, Wo, scale, shift, mean, invstd, drop, argmax, sums, maxes); else launch_pdl((bn_act_bwd_reduce_kernel<0, 4>), dim3(grid), dim3(256), 0, stream, dout, y, B, H, W, C, Ho, Wo, scale, shift, mean, invstd, drop, argmax, sums, maxes); }
  else if (pool == 2) { if (reduce_bps() == 3) launch_pdl((bn_act_bwd_reduce_kernel<2, 3>), dim3(grid), dim3(256), 0, stream, dout, y, B, H, W, C, Ho, Wo, scale, shift, mean, invstd, drop, argmax, sums, maxes); else launch_pdl((bn_act_bwd_reduce_kernel<2, 4>), dim3(grid), dim3(256), 0, stream, dout, y, B, H, W, C, Ho, Wo, scale, shift, mean, invstd, drop, argmax, sums, maxes); }
  else { if (reduce_bps() == 3) launch_pdl((bn_act_bwd_reduce_kernel<3, 3>), dim3(grid), dim3(256), 0, stream, dout, y, B, H, W, C, Ho, Wo, scale, shift, mean, invstd, drop, argmax, sums, maxes); else launch_pdl((bn_act_bwd_reduce_kernel<3, 4>), dim3(grid), dim3(256), 0, stream, dout, y, B, H, W, C, Ho, Wo, scale, shift, mean, invstd, drop, argmax, sums, maxes); }
  PC_LAUNCH_CHECK("bn_act_bwd_reduce_kernel");
  return PC_OK;
}

extern "C" int pc_bn_act_bwd_apply(const float* dout, const float* y, int B, int H, int W, int C, const float* scale,
                                   const float* shift, const float* mean, const float* invstd, const float* drop, int pool,
                                   const uint8_t* argmax, const double* sums, float* dy, float* dgamma, float* dbeta,
                                   float* dy_amax, const float* maxes, void* dy_planes, const double* y_stats, float* db_conv,
                                   pc_stream_t stream) {
  PC_REQUIRE(dout && y && scale && shift && mean && invstd && sums && (dy || dy_planes), PC_EINVAL, "pc_bn_act_bwd_apply: null pointer");
  PC_REQUIRE(db_conv == nullptr || y_stats != nullptr, PC_EINVAL, "pc_bn_act_bwd_apply: db_conv needs the forward statistics y_stats");
  PC_REQUIRE(dy_planes == nullptr || (maxes != nullptr && dy_amax != nullptr), PC_EINVAL,
             "pc_bn_act_bwd_apply: dy_planes needs the maxes of the reduce pass and a dy_amax slot");
  PC_CHECK_C4("pc_bn_act_bwd_apply", C);
  PC_REQUIRE(pool == 0 || pool == 2 || (pool == 3 && argmax), PC_EINVAL, "pc_bn_act_bwd_apply: bad pool / argmax");
  int Ho, Wo;
  pool_out_dims(H, W, pool, &Ho, &Wo);
  const long long items = (long long)B * H * W * (C / 4);
  const int grid = pool == 0 ? ew_grid_waves(bn_act_bwd_apply_kernel<0>, 0, items, 256 * 2)
                             : (pool == 2 ? ew_grid_waves(bn_act_bwd_apply_kernel<2>, 0, items, 256 * 2) : ew_grid_waves(bn_act_bwd_apply_kernel<3>, 0, items, 256 * 2));
  if (pool == 0) launch_pdl((bn_act_bwd_apply_kernel<0>), dim3(grid), dim3(256), 0, stream, dout, y, B, H, W, C, Ho, Wo, scale, shift, mean, invstd, drop, argmax, sums, dy, dgamma, dbeta, dy_amax, maxes, static_cast<unsigned char*>(dy_planes), y_stats, db_conv);
  else if (pool == 2) launch_pdl((bn_act_bwd_apply_kernel<2>), dim3(grid), dim3(256), 0, stream, dout, y, B, H, W, C, Ho, Wo, scale, shift, mean, invstd, drop, argmax, sums, dy, dgamma, dbeta, dy_amax, maxes, static_cast<unsigned char*>(dy_planes), y_stats, db_conv);
  else launch_pdl((bn_act_bwd_apply_kernel<3>), dim3(grid), dim3(256), 0, stream, dout, y, B, H, W, C, Ho, Wo, scale, shift, mean, invstd, drop, argmax, sums, dy, dgamma, dbeta, dy_amax, maxes, static_cast<unsigned char*>(dy_planes), y_stats, db_conv);
  PC_LAUNCH_CHECK("bn_act_bwd_apply_kernel");
  return PC_OK;
}

extern "C" int pc_bn_add_relu_fwd(const float* y2, const float* scale2, const float* shift2, const float* ysc,
                                  const float* sc_scale, const float* sc_shift, int64_t n_pix, int C, float* out,
                                  void* planes, pc_stream_t stream) {
  PC_REQUIRE(y2 && scale2 && shift2 && ysc && out && n_pix > 0, PC_EINVAL, "pc_bn_add_relu_fwd: bad arguments");
  PC_REQUIRE((sc_scale == nullptr) == (sc_shift == nullptr), PC_EINVAL, "pc_bn_add_relu_fwd: shortcut scale/shift mismatch");
  PC_CHECK_C4("pc_bn_add_relu_fwd", C);
  launch_pdl(bn_add_relu_fwd_kernel, dim3(ew_grid_waves(bn_add_relu_fwd_kernel, 0, n_pix * (C / 4), 256)), dim3(256), 0, stream, y2, scale2, shift2, ysc, sc_scale, sc_shift, n_pix, C, out, static_cast<unsigned char*>(planes),
             PcBnFinalize{}, PcBnFinalize{});
  PC_LAUNCH_CHECK("bn_add_relu_fwd_kernel");
  return PC_OK;
}

extern "C" int pc_bn_add_relu_fwd_fin(const float* y2, const PcBnFinalize* fin2, const float* ysc, const PcBnFinalize* fin_s, int64_t n_pix, int C,
                                      float* out, void* planes, pc_stream_t stream) {
  PC_REQUIRE(y2 && fin2 && ysc && out && n_pix > 0, PC_EINVAL, "pc_bn_add_relu_fwd_fin: bad arguments");
  PC_CHECK_C4("pc_bn_add_relu_fwd_fin", C);
  int rc = check_fin("pc_bn_add_relu_fwd_fin", fin2);
  if (rc != PC_OK) return rc;
  if (fin_s != nullptr) {
    rc = check_fin("pc_bn_add_relu_fwd_fin", fin_s);
    if (rc != PC_OK) return rc;
  }
  launch_pdl(bn_add_relu_fwd_kernel, dim3(ew_grid_waves(bn_add_relu_fwd_kernel, sizeof(float) * 4 * (size_t)C, n_pix * (C / 4), 256)), dim3(256), sizeof(float) * 4 * (size_t)C, stream, y2, (const float*)nullptr,
             (const float*)nullptr, ysc, (const float*)nullptr, (const float*)nullptr, (long long)n_pix, C, out, static_cast<unsigned char*>(planes), *fin2,
             fin_s != nullptr ? *fin_s : PcBnFinalize{});
  PC_LAUNCH_CHECK("bn_add_relu_fwd_kernel<fin>");
  return PC_OK;
}

extern "C" int pc_bn_add_relu_bwd_reduce(const float* dout, const float* out, const float* y2, const float* mean2,
                                         const float* invstd2, const float* ysc, const float* mean_s,
                                         const float* invstd_s, int64_t n_pix, int C, double* sums2, double* sums_s,
                                         float* maxes, pc_stream_t stream) {
  PC_REQUIRE(dout && out && y2 && mean2 && invstd2 && sums2 && n_pix > 0, PC_EINVAL, "pc_bn_add_relu_bwd_reduce: bad arguments");
  PC_REQUIRE(sums_s == nullptr || (ysc && mean_s && invstd_s), PC_EINVAL, "pc_bn_add_relu_bwd_reduce: shortcut pointers");
  PC_CHECK_C4("pc_bn_add_relu_bwd_reduce", C);
  if (reduce_bps() == 3) launch_pdl(bn_add_relu_bwd_reduce_kernel<3>, reduce_grid2(n_pix, C), dim3(256), 0, stream, dout, out, y2, mean2, invstd2, ysc, mean_s, invstd_s, n_pix, C, sums2, sums_s, maxes);
  else launch_pdl(bn_add_relu_bwd_reduce_kernel<4>, reduce_grid2(n_pix, C), dim3(256), 0, stream, dout, out, y2, mean2, invstd2, ysc, mean_s, invstd_s, n_pix, C, sums2, sums_s, maxes);
  PC_LAUNCH_CHECK("bn_add_relu_bwd_reduce_kernel");
  return PC_OK;
}

extern "C" int pc_bn_add_relu_bwd_apply(const float* dout, const float* out, const float* y2, const float* scale2,
                                        const float* mean2, const float* invstd2, const double* sums2, const float* ysc,
                                        const float* sc_scale, const float* mean_s, const float* invstd_s,
                                        const double* sums_s, int64_t n_pix, int C, float* dy2, float* dysc_or_dx,
                                        float* dgamma2, float* dbeta2, float* dgamma_s, float* dbeta_s, float* dy2_amax,
                                        float* dysc_amax, const float* maxes, void* dy2_planes, void* dysc_planes,
                                        const double* y2_stats, float* db2, const double* ysc_stats, float* db_s, pc_stream_t stream) {
  PC_REQUIRE(dout && out && y2 && scale2 && mean2 && invstd2 && sums2 && (dy2 || dy2_planes) && (dysc_or_dx || dysc_planes) && n_pix > 0,
             PC_EINVAL, "pc_bn_add_relu_bwd_apply: bad arguments");
  PC_REQUIRE((db2 == nullptr || y2_stats != nullptr) && (db_s == nullptr || ysc_stats != nullptr), PC_EINVAL,
             "pc_bn_add_relu_bwd_apply: db2 / db_s need the forward statistics of y2 / ysc");
  PC_REQUIRE((dy2_planes == nullptr || (maxes && dy2_amax)) && (dysc_planes == nullptr || (maxes && dysc_amax)), PC_EINVAL,
             "pc_bn_add_relu_bwd_apply: *_planes need the maxes of the reduce pass and the matching *_amax slot");
  PC_REQUIRE(sc_scale == nullptr || (ysc && mean_s && invstd_s && sums_s), PC_EINVAL, "pc_bn_add_relu_bwd_apply: shortcut pointers");
  PC_CHECK_C4("pc_bn_add_relu_bwd_apply", C);
  launch_pdl(bn_add_relu_bwd_apply_kernel, dim3(ew_grid_waves(bn_add_relu_bwd_apply_kernel, 0, n_pix * (C / 4), 256 * 2)), dim3(256), 0, stream, dout, out, y2, scale2, mean2, invstd2, sums2, ysc, sc_scale, mean_s, invstd_s, sums_s, n_pix, C, dy2, dysc_or_dx,
      dgamma2, dbeta2, dgamma_s, dbeta_s, dy2_amax, dysc_amax, maxes, static_cast<unsigned char*>(dy2_planes),
      static_cast<unsigned char*>(dysc_planes), y2_stats, db2, ysc_stats, db_s);
  PC_LAUNCH_CHECK("bn_add_relu_bwd_apply_kernel");
  return PC_OK;
}

extern "C" int pc_dropout2d_mask(float* drop, int B, int C, float p, uint64_t seed, uint64_t offset, const int64_t* step_dev,
                                 pc_stream_t stream) {
  PC_REQUIRE(drop && B > 0 && C > 0 && p >= 0.f && p < 1.f, PC_EINVAL, "pc_dropout2d_mask: bad arguments (p=%f)", p);
  const int n = B * C;
  launch_pdl(dropout2d_mask_kernel, dim3(ceil_div(ceil_div(n, 4), 128)), dim3(128), 0, stream, drop, n, p, 1.0f / (1.0f - p), seed, offset,
                                                                            reinterpret_cast<const long long*>(step_dev));
  PC_LAUNCH_CHECK("dropout2d_mask_kernel");
  return PC_OK;
}

extern "C" int pc_bn_act_split(const float* y, int64_t n_pix, int C, int hw, const float* scale, const float* shift, const float* drop,
                               int relu, void* planes, pc_stream_t stream) {
  PC_REQUIRE(y && planes && n_pix > 0 && hw > 0, PC_EINVAL, "pc_bn_act_split: bad arguments");
  PC_REQUIRE(C > 0 && C % 8 == 0, PC_EUNSUPPORTED, "pc_bn_act_split: channels=%d must be a multiple of 8", C);
  PC_REQUIRE((scale == nullptr) == (shift == nullptr), PC_EINVAL, "pc_bn_act_split: scale/shift mismatch");
  PC_REQUIRE((reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(planes) & 15) == 0, PC_EINVAL,
             "pc_bn_act_split: buffers must be 16-byte aligned");
  launch_pdl(bn_act_split_kernel, dim3(ew_grid_waves(bn_act_split_kernel, 0, n_pix * (C / 8), 256)), dim3(256), 0, stream, y, (long long)n_pix, C, hw, scale, shift, drop,
             relu, static_cast<unsigned char*>(planes), PcBnFinalize{});
  PC_LAUNCH_CHECK("bn_act_split_kernel");
  return PC_OK;
}

extern "C" int pc_bn_act_split_fin(const float* y, int64_t n_pix, int C, int hw, const PcBnFinalize* fin, const float* drop, int relu,
                                   void* planes, pc_stream_t stream) {
  PC_REQUIRE(y && planes && fin && n_pix > 0 && hw > 0, PC_EINVAL, "pc_bn_act_split_fin: bad arguments");
  PC_REQUIRE(C > 0 && C % 8 == 0 && C <= 1024, PC_EUNSUPPORTED, "pc_bn_act_split_fin: channels=%d must be a multiple of 8, at most 1024", C);
  const int rc = check_fin("pc_bn_act_split_fin", fin);
  if (rc != PC_OK) return rc;
  launch_pdl(bn_act_split_kernel, dim3(ew_grid_waves(bn_act_split_kernel, sizeof(float) * 2 * (size_t)C, n_pix * (C / 8), 256)), dim3(256), sizeof(float) * 2 * (size_t)C, stream, y, (long long)n_pix, C, hw,
             (const float*)nullptr, (const float*)nullptr, drop, relu, static_cast<unsigned char*>(planes), *fin);
  PC_LAUNCH_CHECK("bn_act_split_kernel<fin>");
  return PC_OK;
}

extern "C" int pc_f16_overflow_query(int reset, int* host_flag, pc_stream_t stream) {
  PC_REQUIRE(host_flag != nullptr, PC_EINVAL, "pc_f16_overflow_query: null pointer");
  unsigned int v = 0u;
  PC_CUDA(cudaMemcpyFromSymbolAsync(&v, pc::g_f16_overflow, sizeof(v), 0, cudaMemcpyDeviceToHost, stream));
  PC_CUDA(cudaStreamSynchronize(stream));
  if (reset && v != 0u) {
    const unsigned int zero = 0u;
    PC_CUDA(cudaMemcpyToSymbolAsync(pc::g_f16_overflow, &zero, sizeof(zero), 0, cudaMemcpyHostToDevice, stream));
    PC_CUDA(cudaStreamSynchronize(stream));
  }
  *host_flag = (int)v;
  return PC_OK;
}
