// Backward of the cnn_deep stem -- Conv2d(1, 64, 7, pad 3) -> BatchNorm2d -> ReLU -> MaxPool2d(3, 2, 1), reference
// src/models/phoneme_cnn.py:211-216 -- WITHOUT touching the full-resolution tensors.
//
// Round 1 ran three passes over the 265 MB pre-BatchNorm tensor y0 (BatchNorm-backward reduce, apply -> dy0 written, stem weight
// gradient reading dy0): 0.58 ms of a 4.0 ms step for 1.5 % of its FLOPs. None of it is needed:
//   * The gradient arriving from the pool, dz, is non-zero only at each window's argmax pixel, and there the BatchNorm output
//     equals the POOLED output p0 itself (p0 = max relu(bn(y)) = relu(bn(y[p*]))): the ReLU gate is p0 > 0 and
//     xhat[p*] = (p0 - beta) / gamma. So sum dz, sum dz * xhat (-> dbeta, dgamma and the BatchNorm projection terms) and
//     T1[o][t] = sum_p dz[p,o] x_t[p]   (x_t[p] = input sample under tap t of pixel p) follow from the pooled-resolution tensors
//     (dpool, p0, argmax: 148 MB instead of 1.1 GB of traffic).
//   * dy0 = s (dz - m1 - xhat m2), s = gamma * invstd, m1 = sum dz / M, m2 = sum dz xhat / M, hence
//         dW[o][t] = sum_p dy0[p,o] x_t[p] = s_o ( T1[o][t] - m1_o X1[t] - m2_o T2[o][t] ),
//         X1[t] = sum_p x_t[p],   T2[o][t] = sum_p xhat[p,o] x_t[p] = invstd_o ( sum_t' W[o][t'] G[t'][t] + b_o X1[t] - mu_o X1[t] ),
//         G[t'][t] = sum_p x_t'[p] x_t[p]   (49 x 49 Gram matrix of the input patches: data only, no channel index),
//     because y0 = W x + b is linear in the patches. The dense (all-pixel) parts of the gradient are thus closed forms in G and
//     X1; only the sparse part T1 needs a pass over data, and that pass is at pooled resolution.
//   * d(bias) = sum_p dy0 = 0 exactly (the batch mean absorbs the bias); the reference reports fp32 round-off there.
//
// Kernels: stem_gram_kernel (SIMT; G and X1 from lag correlations of each image, exact border handling; independent of the
// backward chain, launched on the weight-gradient stream during the forward), stem_bwd_pool_kernel (tcgen05: T1 as a GEMM whose
// reduction index is (pooled pixel, window position): operand rows built in shared memory from dpool / p0 / argmax and from the
// input patches, FP16X2 split, TMEM accumulators; per-channel sums in registers), stem_bwd_finish_kernel (the closed form, fp64).
#include "common.cuh"
#include "tc_common.cuh"

namespace pc {
namespace stemb {

using namespace pc::tc;

constexpr int KS = 7, NT = 49, LAG = 13, NLAG = 169;      // kernel size, taps, lags per axis
constexpr int CO = 64;                                     // output channels (the stem of the reference's cnn_deep)

// ------------------------------------------------------------------------------------------------ Gram matrix of the patches
// For tap t' = (r', s') and t = (r, s):  G[t'][t] = sum_{i in I(r')} sum_{j in J(s')} x[i][j] * xz[i + r - r'][j + s - s'],
// where xz is x zero-extended and I(r') = [max(0, r'-3), min(H-1, H-4+r')] (the rows some output pixel reads through tap row
// r'), J likewise. With f(i, j; d) = x[i][j] xz[i + dr][j + ds] summed over the WHOLE image (C0), over the first / last three
// rows (R), columns (C) and their 6 x 6 corner blocks (X), inclusion-exclusion gives every G[t'][t] from 49 numbers per lag.
// Thread roles: 13 row lags x 4 groups of 4 column lags x 4 row partitions (208 of 256 threads) for C0 / R -- each thread slides a
// 4-wide register window along the lagged row, i.e. 2 shared-memory loads per 4 multiply-adds; then one thread per (lag, boundary
// column) for C / X. The per-lag table [49] lives in shared memory and accumulates over the CTA's samples.
__global__ void __launch_bounds__(256) stem_gram_kernel(const float* __restrict__ x, int B, int H, int W, double* __restrict__ G,
                                                        double* __restrict__ X1) {
  extern __shared__ float sm[];
  const int Wz = W + 16, Hz = H + 12;              // 6 zero rows above / below, 6 zero columns left, 10 right (the 4-wide window over-reads)
  float* xz = sm;                                  // [(H+12)][(W+16)] zero-extended image
  float* tab = xz + Hz * Wz;                       // [NLAG][49]: c0 | rows[6] | cols[6] | corner[6][6]
  float* rowval = tab + NLAG * 49;                 // [H][7]: sum of row i over the columns tap column s reaches
  const int tid = threadIdx.x;
  for (int i = tid; i < NLAG * 49; i += blockDim.x) tab[i] = 0.f;
  float x1 = 0.f;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < Hz * Wz; i += blockDim.x) {
      const int hh = i / Wz - 6, ww = i % Wz - 6;
      xz[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[((size_t)b * H + hh) * W + ww] : 0.f;
    }
    __syncthreads();
    if (tid < 208) {
      const int part = tid & 3, g = (tid >> 2) & 3, dri = tid >> 4;       // row partition, column-lag group, row lag index
      const int dr = dri - 6, ds0 = 4 * g - 6;
      float c0[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = part; i < H; i += 4) {
        const float* a = xz + (i + 6) * Wz + 6;
        const float* bb = xz + (i + 6 + dr) * Wz + 6 + ds0;
        float w0 = bb[0], w1 = bb[1], w2 = bb[2];
        float rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int j = 0; j < W; ++j) {
          const float av = a[j], w3 = bb[j + 3];
          rs[0] = fmaf(av, w0, rs[0]); rs[1] = fmaf(av, w1, rs[1]); rs[2] = fmaf(av, w2, rs[2]); rs[3] = fmaf(av, w3, rs[3]);
          w0 = w1; w1 = w2; w2 = w3;
        }
        const int slot = i < 3 ? i : (i >= H - 3 ? i - (H - 6) : -1);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          c0[k] += rs[k];
          if (slot >= 0 && ds0 + k <= 6) atomicAdd(&tab[(dri * LAG + ds0 + k + 6) * 49 + 1 + slot], rs[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ds0 + k <= 6) atomicAdd(&tab[(dri * LAG + ds0 + k + 6) * 49], c0[k]);
    }
    // boundary columns and corners: one thread per (lag, boundary-column slot)
    for (int task = tid; task < NLAG * 6; task += blockDim.x) {
      const int lag = task / 6, cs = task - lag * 6;
      const int dr = lag / LAG - 6, ds = lag % LAG - 6;
      const int jc = cs < 3 ? cs : W - 6 + cs;
      float col = 0.f;
      for (int i = 0; i < H; ++i) col = fmaf(xz[(i + 6) * Wz + 6 + jc], xz[(i + 6 + dr) * Wz + 6 + jc + ds], col);
      float* t = tab + lag * 49;
      t[7 + cs] += col;
#pragma unroll
      for (int rsl = 0; rsl < 6; ++rsl) {
        const int i = rsl < 3 ? rsl : H - 6 + rsl;
        t[13 + rsl * 6 + cs] += xz[(i + 6) * Wz + 6 + jc] * xz[(i + 6 + dr) * Wz + 6 + jc + ds];
      }
    }
    // X1[(r, s)] = sum of x over rows [max(0, r-3), min(H-1, H-4+r)] x columns [max(0, s-3), min(W-1, W-4+s)]: one warp per row
    // forms the row total and removes the (at most three) border columns tap column s does not reach, then 49 threads add rows
    for (int i = tid >> 5; i < H; i += (int)(blockDim.x >> 5)) {
      const float* a = xz + (i + 6) * Wz + 6;
      float t = 0.f;
      for (int j = tid & 31; j < W; j += 32) t += a[j];
      t = warp_sum(t);
      if ((tid & 31) == 0) {
        float* rv = rowval + i * 7;
        rv[3] = t;
        rv[2] = t - a[W - 1]; rv[1] = rv[2] - a[W - 2]; rv[0] = rv[1] - a[W - 3];
        rv[4] = t - a[0]; rv[5] = rv[4] - a[1]; rv[6] = rv[5] - a[2];
      }
    }
    __syncthreads();
    if (tid < NT) {
      const int r = tid / KS, s2 = tid % KS;
      const int i0 = max(0, r - 3), i1 = min(H - 1, H - 4 + r);
      float t = 0.f;
      for (int i = i0; i <= i1; ++i) t += rowval[i * 7 + s2];
      x1 += t;
    }
  }
  __syncthreads();
  // assemble G[t'][t] for this CTA's samples and add it to the global fp64 matrix
  for (int pair = tid; pair < NT * NT; pair += blockDim.x) {
    const int tp = pair / NT, t = pair % NT;
    const int rp = tp / KS, sp = tp % KS, r = t / KS, s2 = t % KS;
    const float* e = tab + ((r - rp + 6) * LAG + (s2 - sp + 6)) * 49;
    // excluded rows: r' < 3 -> the last 3 - r' rows (slots 3 + r' .. 5); r' > 3 -> the first r' - 3 rows (slots 0 .. r' - 4)
    const int ra = rp < 3 ? 3 + rp : 0, rb = rp < 3 ? 6 : (rp > 3 ? rp - 3 : 0);
    const int ca = sp < 3 ? 3 + sp : 0, cb = sp < 3 ? 6 : (sp > 3 ? sp - 3 : 0);
    double v = (double)e[0];
    for (int k = ra; k < rb; ++k) v -= (double)e[1 + k];
    for (int k = ca; k < cb; ++k) v -= (double)e[7 + k];
    for (int k = ra; k < rb; ++k)
      for (int q = ca; q < cb; ++q) v += (double)e[13 + k * 6 + q];
    atomicAdd(G + pair, v);
  }
  if (tid < NT) atomicAdd(X1 + tid, (double)x1);
}

// ------------------------------------------------------------------------------------------------ max |x| of a tensor
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n4, float* __restrict__ out) {
  float m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(m));
}

// ------------------------------------------------------------------------------------------------ T1 on tensor cores
constexpr int QPB = 14;                    // pooled pixels per 128-row K block: row = ql * 9 + window position a (126 rows + 2 zero);
                                           // a block's pixels lie in ONE pooled row (pw = 14 seg + ql), so all its input patches fall
                                           // inside a 9 x 36 window of the image that is staged in shared memory once per block
constexpr int RG_H = 9, RG_W = 36, RG_LDH = 38;        // staged window: 9 x 36 samples; fp16 rows of 38 halves (even: 4-byte aligned chunks)
constexpr int RG_ASZ = RG_H * RG_LDH;                  // one of the four window arrays (hi / lo, two column parities)
constexpr int ROWS = 128;
constexpr uint32_t PART = ROWS * 128;      // one 64-wide group of one operand part: [128 rows][128 B], SWIZZLE_128B (MN-major)
constexpr uint32_t STAGE = 4 * PART;       // A_hi | A_lo | B_hi | B_lo
constexpr int NSTAGE = 3;
constexpr int NGROUPS = 3, NPROD = 128, PROD_WARPS = 4 * NGROUPS;
constexpr int THREADS = 32 * (PROD_WARPS + 1);

struct PoolParams {
  const float* dpool;        // [B][Hp][Wp][64] gradient w.r.t. the pooled output
  const float* p0;           // [B][Hp][Wp][64] pooled output
  const uint8_t* argmax;     // [B][Hp][Wp][64] window position kh * 3 + kw of the maximum
  const float* x;            // [B][H][W]
  const float* scale; const float* shift; const float* mean; const float* invstd;
  const float* amax;         // device scalar: max |dpool|
  float* t1_partial;         // [grid][64 taps][64 channels]
  double* sums;              // [2][64]: sum dz, sum dz * xhat
  int B, H, W, Hp, Wp;
  long long n_pooled;        // B * Hp * Wp
  int n_seg;                 // K blocks per pooled row: ceil(Wp / QPB)
  int n_blocks;              // B * Hp * n_seg
  FastDiv d_seg, d_hp;
};

__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // next 64-wide group
  d |= (uint64_t)(1024 >> 4) << 32;                   // next 8-row atom
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                             // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ uint32_t mn_off(int row, int j) { return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((j ^ (row & 7)) << 4)); }

__global__ void __launch_bounds__(THREADS, 1) stem_bwd_pool_kernel(const PoolParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* tiles = smem;                                           // [NSTAGE][STAGE]
  unsigned char* zeros = smem + (size_t)NSTAGE * STAGE;                  // [PART] second (all-zero) group of the A operand
  uint64_t* full = reinterpret_cast<uint64_t*>(zeros + PART);
  uint64_t* empty = full + NSTAGE;
  uint64_t* acc_full = empty + NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 2);                // [2][64]
  __half* s_reg = reinterpret_cast<__half*>(s_red + 128);                // [NGROUPS][hi0 | hi1 | lo0 | lo1][RG_H][RG_LDH] staged image windows, pre-split

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (int)(PART / 16); i += THREADS) reinterpret_cast<uint4*>(zeros)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 128; i += THREADS) s_red[i] = 0.f;
  if (warp == PROD_WARPS) {
    if (lane == 0) {
      for (int s = 0; s < NSTAGE; ++s) {
        mbar_init(&full[s], NPROD);
        mbar_init(&empty[s], 1);
      }
      mbar_init(acc_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // this CTA's K blocks: a contiguous range (the pooled pixels of a CTA stay within one or two images -> input patches hit L1)
  const int per = (p.n_blocks + gridDim.x - 1) / gridDim.x;
  const int blk0 = blockIdx.x * per, blk1 = min(p.n_blocks, blk0 + per);
  const int n_my = max(0, blk1 - blk0);

  if (warp < PROD_WARPS) {
    const int group = warp >> 2, gt = tid & (NPROD - 1);
    const float g_scale = f16_operand_scale(p.amax[0]);
    // ---- roles inside a group: every thread builds 8 patch chunks (rows pr + 16 i, taps 8 j ..); threads 0..111 additionally
    // one S unit (pooled pixel ql, channels 8 cj ..), threads 112..127 zero the two padding rows of S
    const int j = gt & 7, pr = gt >> 3;
    const int ql = gt >> 3, cj = gt & 7;
    float4 sc0 = make_float4(0, 0, 0, 0), sc1 = sc0, sh0 = sc0, sh1 = sc0, mu0 = sc0, mu1 = sc0, is0 = sc0, is1 = sc0;
    if (gt < 112) {
      sc0 = *reinterpret_cast<const float4*>(p.scale + 8 * cj); sc1 = *reinterpret_cast<const float4*>(p.scale + 8 * cj + 4);
      sh0 = *reinterpret_cast<const float4*>(p.shift + 8 * cj); sh1 = *reinterpret_cast<const float4*>(p.shift + 8 * cj + 4);
      mu0 = *reinterpret_cast<const float4*>(p.mean + 8 * cj); mu1 = *reinterpret_cast<const float4*>(p.mean + 8 * cj + 4);
      is0 = *reinterpret_cast<const float4*>(p.invstd + 8 * cj); is1 = *reinterpret_cast<const float4*>(p.invstd + 8 * cj + 4);
    }
    // xhat at the argmax pixel from the pooled value: bn(y) = p0  =>  xhat = ((p0 - shift) / scale - mean) * invstd = xk1 * p0 + xk0
    float xk1[8], xk0[8];
    {
      const float scv[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
      const float shv[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
      const float muv[8] = {mu0.x, mu0.y, mu0.z, mu0.w, mu1.x, mu1.y, mu1.z, mu1.w};
      const float isv[8] = {is0.x, is0.y, is0.z, is0.w, is1.x, is1.y, is1.z, is1.w};
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float inv = scv[q] != 0.f ? 1.f / scv[q] : 0.f;
        xk1[q] = inv * isv[q];
        xk0[q] = scv[q] != 0.f ? (-shv[q] * inv - muv[q]) * isv[q] : 0.f;
      }
    }
    float rs[8] = {0, 0, 0, 0, 0, 0, 0, 0}, rq[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // The reduction index of T1's M side is m = 8 * tr + ts (tap row tr = 16-byte chunk j of an operand row, tap column ts < 7 inside
    // it; m = 8 tr + 7 and chunk 7 carry junk that the finish kernel never reads): a chunk is then 8 CONSECUTIVE window samples,
    // copied as four 4-byte words from a window that was split into fp16 hi / lo ONCE per sample when it was staged. Each plane is
    // kept twice, sample cc at half index cc (copy 0) and cc + 1 (copy 1), so that chunks starting at an odd column are aligned too.
    __half* reg = s_reg + (size_t)group * 4 * RG_ASZ;
    int row_base[8];                   // half offset (inside the hi arrays) of the window sample under tap (0, 0) of my 8 rows (row = ql' * 9 + a')
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = pr + 16 * i;
      const int rq_ = row / 9, ra = row - rq_ * 9;
      const int cbase = 2 * min(rq_, QPB - 1) + ra % 3, sel = cbase & 1;
      row_base[i] = sel * RG_ASZ + (ra / 3) * RG_LDH + cbase + sel;
    }
    // Global loads of a K block -- its image window and its S unit (dpool, p0, argmax) -- are issued ONE block ahead into registers:
    // measured, a group otherwise sits out two full DRAM latencies per block (window -> barrier -> patch build; dpool / p0 -> S).
    struct Pre {
      float rv[3];
      float4 d0, d1, a0, a1;
      uint2 am;
      int n_valid;
    };
    auto issue = [&](int it, Pre& r) {
      uint32_t bph, seg, b, php;
      p.d_seg.divmod((uint32_t)(blk0 + it), bph, seg);
      p.d_hp.divmod(bph, b, php);
      const int pw0 = (int)seg * QPB;
      r.n_valid = min(QPB, p.Wp - pw0);
      // image window rows 2 ph - 4 .. 2 ph + 4, columns 2 pw0 - 4 .. 2 pw0 + 31 (zero outside the image)
      const float* img = p.x + (size_t)b * p.H * p.W;
      const int h0 = 2 * (int)php - 4, w0 = 2 * pw0 - 4;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int e = gt + 128 * k;
        const int rr = e / RG_W, cc = e - rr * RG_W;
        const int hi = h0 + rr, wi = w0 + cc;
        r.rv[k] = (e < RG_H * RG_W && (unsigned)hi < (unsigned)p.H && (unsigned)wi < (unsigned)p.W) ? img[hi * p.W + wi] : 0.f;
      }
      r.d0 = r.d1 = r.a0 = r.a1 = make_float4(0.f, 0.f, 0.f, 0.f);
      r.am = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
      if (gt < 112 && ql < r.n_valid) {
        const size_t o = (((size_t)b * p.Hp + php) * p.Wp + pw0 + ql) * CO + 8 * cj;
        r.d0 = *reinterpret_cast<const float4*>(p.dpool + o); r.d1 = *reinterpret_cast<const float4*>(p.dpool + o + 4);
        r.a0 = *reinterpret_cast<const float4*>(p.p0 + o); r.a1 = *reinterpret_cast<const float4*>(p.p0 + o + 4);
        r.am = *reinterpret_cast<const uint2*>(p.argmax + o);
      }
    };
    Pre cur;
    if (group < n_my) issue(group, cur);
    for (int it = group; it < n_my; it += NGROUPS) {
      const int s = it % NSTAGE;
      const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
      Pre nxt = cur;
      if (it + NGROUPS < n_my) issue(it + NGROUPS, nxt);
      const int n_valid = cur.n_valid;
      // ---------------- stage the image window
      asm volatile("bar.sync %0, 128;" ::"r"(2 + group) : "memory");      // the previous block's patch reads are done
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int e = gt + 128 * k;
        const int rr = e / RG_W, cc = e - rr * RG_W;
        if (e < RG_H * RG_W) {
          const float v = cur.rv[k];
          const __half vh = __float2half_rn(v);
          const __half vl = __float2half_rn((v - __half2float(vh)) * kF16LoScale);
          __half* w0 = reg + rr * RG_LDH + cc;
          w0[0] = vh; w0[RG_ASZ + 1] = vh; w0[2 * RG_ASZ] = vl; w0[3 * RG_ASZ + 1] = vl;
        }
      }
      // ---------------- S unit
      uint4 hh = make_uint4(0u, 0u, 0u, 0u), ll = hh;
      const uint2 am = cur.am;
      if (gt < 112 && ql < n_valid) {
        const float d[8] = {cur.d0.x, cur.d0.y, cur.d0.z, cur.d0.w, cur.d1.x, cur.d1.y, cur.d1.z, cur.d1.w};
        const float a[8] = {cur.a0.x, cur.a0.y, cur.a0.z, cur.a0.w, cur.a1.x, cur.a1.y, cur.a1.z, cur.a1.w};
        float dz[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          dz[q] = a[q] > 0.f ? d[q] : 0.f;                                   // ReLU gate: pooled output > 0
          const float xh = fmaf(a[q], xk1[q], xk0[q]);
          rs[q] += dz[q];
          rq[q] = fmaf(dz[q], xh, rq[q]);
          dz[q] *= g_scale;
        }
        split_f16x2(dz[0], dz[1], hh.x, ll.x); split_f16x2(dz[2], dz[3], hh.y, ll.y);
        split_f16x2(dz[4], dz[5], hh.z, ll.z); split_f16x2(dz[6], dz[7], hh.w, ll.w);
      }
      asm volatile("bar.sync %0, 128;" ::"r"(2 + group) : "memory");        // staged window visible to the group
      mbar_wait(&empty[s], ph ^ 1u);
      unsigned char* a_hi = tiles + (size_t)s * STAGE;
      unsigned char* a_lo = a_hi + PART;
      unsigned char* b_hi = a_hi + 2 * PART;
      unsigned char* b_lo = a_hi + 3 * PART;
      // ---------------- patch chunks: row (ql', a'), chunk j = tap row j of the conv pixel (2 ph - 1 + a'/3, 2 (pw0 + ql') - 1 + a'%3),
      // i.e. window samples [a'/3 + j][2 ql' + a'%3 .. + 7]
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = pr + 16 * i;
        uint4 h = make_uint4(0u, 0u, 0u, 0u), l = h;
        if (j < KS) {
          const uint32_t* ph_ = reinterpret_cast<const uint32_t*>(reg + row_base[i] + j * RG_LDH);
          const uint32_t* pl_ = reinterpret_cast<const uint32_t*>(reg + row_base[i] + j * RG_LDH + 2 * RG_ASZ);
          h = make_uint4(ph_[0], ph_[1], ph_[2], ph_[3]);
          l = make_uint4(pl_[0], pl_[1], pl_[2], pl_[3]);
        }
        const uint32_t off = mn_off(row, j);
        *reinterpret_cast<uint4*>(a_hi + off) = h;
        *reinterpret_cast<uint4*>(a_lo + off) = l;
      }
      if (gt < 112) {
        // row ql * 9 + a holds this pixel's (scaled) gradient in the channels whose argmax is window position a, zero elsewhere:
        // clear my nine 16-byte chunks of both planes, then drop each channel's two halves into the row its argmax names
#pragma unroll
        for (int a = 0; a < 9; ++a) {
          const uint32_t off = mn_off(ql * 9 + a, cj);
          *reinterpret_cast<uint4*>(b_hi + off) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(b_lo + off) = make_uint4(0u, 0u, 0u, 0u);
        }
        const uint32_t amv[2] = {am.x, am.y}, hw[4] = {hh.x, hh.y, hh.z, hh.w}, lw[4] = {ll.x, ll.y, ll.z, ll.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t a = (amv[q >> 2] >> (8 * (q & 3))) & 0xFFu;
          if (a < 9u) {
            const uint32_t off = mn_off(ql * 9 + (int)a, cj) + 2u * (uint32_t)q;
            *reinterpret_cast<unsigned short*>(b_hi + off) = (unsigned short)(hw[q >> 1] >> (16 * (q & 1)));
            *reinterpret_cast<unsigned short*>(b_lo + off) = (unsigned short)(lw[q >> 1] >> (16 * (q & 1)));
          }
        }
      } else {
        const int u = gt - 112;                       // rows 126, 127 of S: zero
        const uint32_t off = mn_off(126 + (u >> 3), u & 7);
        *reinterpret_cast<uint4*>(b_hi + off) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(b_lo + off) = make_uint4(0u, 0u, 0u, 0u);
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
      cur = nxt;        // (the register moves wait for the prefetched loads: keep them behind this block's tile build)
    }
    // ---------------- per-channel sums of this CTA
    if (gt < 112) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        atomicAdd(&s_red[8 * cj + q], rs[q]);
        atomicAdd(&s_red[64 + 8 * cj + q], rq[q]);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * PROD_WARPS) : "memory");
    if (tid < 128) atomicAdd(p.sums + tid, (double)s_red[tid]);
    // ---------------- epilogue: T1 partial of this CTA (rows = taps 0..63 live in TMEM lanes 0..63: lane quarters 0 and 1)
    if ((warp & 3) < 2 && (warp >> 2) < 2) {
      float* dst = p.t1_partial + (size_t)blockIdx.x * 64 * CO;
      const int quarter = warp & 3, c0 = 32 * (warp >> 2);
      const int row = quarter * 32 + lane;
      float v[32];
      if (n_my == 0) {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0.f;
      } else {
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(taddr, r0);
        tmem_ld_32x32(taddr + 128, r1);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r0[k]) + __uint_as_float(r1[k]);
        tmem_ld_32x32(taddr + 64, r0);
        tmem_ld_32x32(taddr + 192, r1);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = fmaf(__uint_as_float(r0[k]) + __uint_as_float(r1[k]), kF16LoInv, v[k]);
      }
#pragma unroll
      for (int k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(dst + (size_t)row * CO + c0 + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    }
  } else {
    // ---------------- MMA issuer: D[128 (taps | zero half) x 64 channels] += patch^T S over the 128 rows of a stage
    if (lane == 0 && n_my > 0) {
      const uint32_t idesc = instr_desc(0u, 128, 64) | (1u << 15) | (1u << 16);       // both operands MN-major
      const uint32_t idesc2 = instr_desc(0u, 128, 128) | (1u << 15) | (1u << 16);
      const uint32_t zero_addr = smem_u32(zeros);
      for (int it = 0; it < n_my; ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t base = smem_u32(tiles + (size_t)s * STAGE);
        // A: group 0 = the patch tile, group 1 = the shared zero tile (taps 64..127 do not exist); B: [S_hi ; S_lo] back to back
        const uint64_t a_hi = desc_mn(base, zero_addr - base);
        const uint64_t a_lo = desc_mn(base + PART, zero_addr - (base + PART));
        const uint64_t b_hi = desc_mn(base + 2 * PART, PART);
#pragma unroll
        for (int kk = 0; kk < ROWS / 16; ++kk) {
          const uint64_t adv = (uint64_t)(kk * (2048 >> 4));
          const int ks = it * (ROWS / 16) + kk;
          const uint32_t d_set = tmem_base + (uint32_t)((ks & 1) * 128);
          mma_bf16(d_set, a_hi + adv, b_hi + adv, idesc2, ks < 2 ? 0u : 1u);
          mma_bf16(d_set + 64, a_lo + adv, b_hi + adv, idesc, 1u);
        }
        mma_commit(&empty[s]);
      }
      mma_commit(acc_full);
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == PROD_WARPS) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------ closed form
__global__ void __launch_bounds__(256) stem_bwd_finish_kernel(const float* __restrict__ t1_partial, int n_partial, const double* __restrict__ sums,
                                                              const double* __restrict__ G, const double* __restrict__ X1,
                                                              const float* __restrict__ w, const float* __restrict__ bias,
                                                              const float* __restrict__ gamma, const float* __restrict__ mean,
                                                              const float* __restrict__ invstd, const float* __restrict__ amax, double M,
                                                              float* __restrict__ dw, float* __restrict__ db, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta) {
  const double unscale = 1.0 / (double)f16_operand_scale(amax[0]);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < CO * NT; idx += gridDim.x * blockDim.x) {
    const int o = idx / NT, t = idx % NT;
    double t1 = 0.0;
    const int trow = 8 * (t / KS) + t % KS;             // T1's M index of tap t (see stem_bwd_pool_kernel)
    for (int k = 0; k < n_partial; ++k) t1 += (double)t1_partial[((size_t)k * 64 + trow) * CO + o];
    t1 *= unscale;
    double yx = 0.0;
    for (int tp = 0; tp < NT; ++tp) yx += (double)w[o * NT + tp] * G[tp * NT + t];
    const double b = bias != nullptr ? (double)bias[o] : 0.0;
    yx += b * X1[t];
    const double is = (double)invstd[o], mu = (double)mean[o], g = (double)(gamma != nullptr ? gamma[o] : 1.f);
    const double t2 = is * (yx - mu * X1[t]);
    const double m1 = sums[o] / M, m2 = sums[CO + o] / M;
    dw[idx] = (float)(g * is * (t1 - m1 * X1[t] - m2 * t2));
    if (t == 0) {
      if (db != nullptr) db[o] = 0.f;
      if (dgamma != nullptr) dgamma[o] = (float)sums[CO + o];
      if (dbeta != nullptr) dbeta[o] = (float)sums[o];
    }
  }
}

}  // namespace stemb
}  // namespace pc

using namespace pc;
using namespace pc::stemb;

extern "C" int pc_stem_bwd_supported(int k, int Cout, int H, int W) {
  return (k == 7 && Cout == 64 && H >= 7 && W >= 7 && (size_t)(H + 12) * (W + 16) * 4 + NLAG * 49 * 4 + (size_t)H * 7 * 4 <= 200 * 1024) ? 1 : 0;
}

// G [49*49] and X1 [49] (fp64, ACCUMULATED into: the caller zeroes them) of the stem's input patches, x [B][H][W].
extern "C" int pc_stem_gram(const float* x, int B, int H, int W, double* G, double* X1, pc_stream_t stream) {
  PC_REQUIRE(x && G && X1 && B > 0, PC_EINVAL, "pc_stem_gram: bad arguments");
  PC_REQUIRE(pc_stem_bwd_supported(7, 64, H, W), PC_EUNSUPPORTED, "pc_stem_gram: image %dx%d not covered", H, W);
  const size_t smem = ((size_t)(H + 12) * (W + 16) + NLAG * 49 + (size_t)H * 7) * sizeof(float);
  static size_t conf = 0;
  if (smem > conf) {
    PC_CUDA(cudaFuncSetAttribute(stem_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  const int grid = B < 2 * kNumSMs ? B : 2 * kNumSMs;
  stem_gram_kernel<<<grid, 256, smem, stream>>>(x, B, H, W, G, X1);
  PC_LAUNCH_CHECK("stem_gram_kernel");
  return PC_OK;
}

extern "C" size_t pc_stem_bwd_workspace(void) { return (size_t)kNumSMs * 64 * CO * sizeof(float) + 64; }

// Everything the stem's backward produces (there is no gradient w.r.t. the input): dw [64][1][7][7], db, dgamma, dbeta.
// sums [2][64] fp64 and amax_slot (float) must be zero on entry; G / X1 from pc_stem_gram on the same x.
extern "C" int pc_stem_bwd(const float* dpool, const float* p0, const uint8_t* argmax, const float* x, int B, int H, int W, const float* w_oihw,
                           const float* bias, const float* gamma, const float* scale, const float* shift, const float* mean,
                           const float* invstd, const double* G, const double* X1, double* sums, float* amax_slot, void* workspace,
                           size_t workspace_bytes, float* dw, float* db, float* dgamma, float* dbeta, pc_stream_t stream) {
  PC_REQUIRE(dpool && p0 && argmax && x && w_oihw && scale && shift && mean && invstd && G && X1 && sums && amax_slot && workspace && dw,
             PC_EINVAL, "pc_stem_bwd: null pointer");
  PC_REQUIRE(pc_stem_bwd_supported(7, 64, H, W), PC_EUNSUPPORTED, "pc_stem_bwd: shape not covered");
  PC_REQUIRE(workspace_bytes >= pc_stem_bwd_workspace(), PC_EINVAL, "pc_stem_bwd: workspace too small");
  const int Hp = (H + 2 - 3) / 2 + 1, Wp = (W + 2 - 3) / 2 + 1;
  const long long n_pooled = (long long)B * Hp * Wp;
  PC_REQUIRE(n_pooled * CO < (1LL << 31), PC_EUNSUPPORTED, "pc_stem_bwd: batch too large for 32-bit pooled indices");
  {
    const long long n4 = n_pooled * CO / 4;
    int grid = ceil_div(n4, 256 * 8);
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    absmax_kernel<<<grid, 256, 0, stream>>>(dpool, n4, amax_slot);
    PC_LAUNCH_CHECK("absmax_kernel");
  }
  PoolParams p{};
  p.dpool = dpool; p.p0 = p0; p.argmax = argmax; p.x = x; p.scale = scale; p.shift = shift; p.mean = mean; p.invstd = invstd;
  p.amax = amax_slot; p.t1_partial = static_cast<float*>(workspace); p.sums = sums;
  p.B = B; p.H = H; p.W = W; p.Hp = Hp; p.Wp = Wp; p.n_pooled = n_pooled;
  p.n_seg = ceil_div(Wp, QPB);
  p.n_blocks = B * Hp * p.n_seg;
  p.d_seg = FastDiv::make((uint32_t)p.n_seg); p.d_hp = FastDiv::make((uint32_t)Hp);
  const size_t smem = (size_t)NSTAGE * STAGE + PART + sizeof(uint64_t) * (2 * NSTAGE + 1) + 16 + 128 * sizeof(float) + (size_t)NGROUPS * 4 * RG_ASZ * sizeof(__half) + 1024;
  static bool conf = false;
  if (!conf) {
    PC_CUDA(cudaFuncSetAttribute(stem_bwd_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = true;
  }
  const int grid = p.n_blocks < kNumSMs ? p.n_blocks : kNumSMs;
  stem_bwd_pool_kernel<<<grid, THREADS, smem, stream>>>(p);
  PC_LAUNCH_CHECK("stem_bwd_pool_kernel");
  stem_bwd_finish_kernel<<<ceil_div(CO * NT, 256), 256, 0, stream>>>(p.t1_partial, grid, sums, G, X1, w_oihw, bias, gamma, mean, invstd, amax_slot,
                                                                   (double)B * H * W, dw, db, dgamma, dbeta);
  PC_LAUNCH_CHECK("stem_bwd_finish_kernel");
  return PC_OK;
}
