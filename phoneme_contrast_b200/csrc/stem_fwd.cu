// Forward of the cnn_deep stem as ONE pass: Conv2d(1, 64, 7, pad 3) -> BatchNorm2d -> ReLU -> MaxPool2d(3, 2, 1)
// (reference src/models/phoneme_cnn.py:211-216) on tcgen05, writing only the pooled output.
//
// Round 1 wrote the 265 MB pre-BatchNorm tensor y0 with a SIMT convolution (0.26 ms), then read it again to normalise and pool
// (0.14 ms) -- a two-pass structure forced by BatchNorm needing the batch statistics of y0 before it can be applied. For THIS
// layer the statistics do not need y0: y0 = W x_patch + b is linear in the 49-tap input patches, so
//     sum_p y0[p,o]   = sum_t W[o,t] X1[t] + M b_o
//     sum_p y0[p,o]^2 = sum_{t,t'} W[o,t] W[o,t'] G[t,t'] + 2 b_o sum_t W[o,t] X1[t] + M b_o^2
// with G / X1 the Gram matrix / tap sums of the patches (csrc/stem_bwd.cu computes them from lag correlations of the input; the
// backward needs them anyway). pc_stem_stats_from_gram evaluates these closed forms in fp64; BatchNorm's scale / shift are then
// known BEFORE the convolution runs and the whole stem becomes a single kernel whose epilogue normalises, rectifies and pools:
//   work item  = one pooled row (b, ph): conv rows 2ph-1, 2ph, 2ph+1 (those inside the image), one 128-pixel MMA tile each;
//   producers  = 2 groups x 4 warps: stage the 9 x (W+6) input window of the item in shared memory, build the K-major
//                [128 pixels][64 taps] fp16 hi / lo operand tiles of its rows (taps 49..63 zero);
//   MMA        = a_hi x [w_hi ; w_lo] (N = 128: main | corr) + a_lo x w_hi per 16-tap k-step, weights resident in smem,
//                accumulators in a ring of four 128-column TMEM slots;
//   epilogue   = 4 warps: TMEM -> + bias -> scale / shift -> ReLU -> shared act[row][channel][w]; then every pooled pixel takes
//                the first maximum of its 3x3 window in scan order (PyTorch's tie rule) and the pooled value, its window
//                position (argmax, for the backward) and, optionally, the fp16 hi | lo planes for block 0 are stored.
// y0 is never materialised: 265 MB less to write, read twice and keep.
#include "common.cuh"
#include "tc_common.cuh"

namespace pc {
namespace stemf {

using namespace pc::tc;

constexpr int KS = 7, NT = 49, CO = 64;
constexpr int NSTAGE = 3, NSLOT = 4;
constexpr int NGROUPS = 2, NPROD = 128, PROD_WARPS = 4 * NGROUPS, EPI_WARPS = 8;
constexpr int THREADS = 32 * (PROD_WARPS + 1 + EPI_WARPS);
constexpr uint32_t A_PART = 128 * 128;          // [128 pixels][128 B]
constexpr uint32_t STAGE = 2 * A_PART;          // hi | lo
constexpr uint32_t B_BYTES = 2 * CO * 128;      // [w_hi rows ; w_lo rows]

struct Params {
  const float* x; const float* w; const float* bias; const float* scale; const float* shift;
  float* p0; uint8_t* argmax; unsigned char* planes;
  int B, H, W, Hp, Wp, WLD, LDH;     // LDH: halves per row of the pre-split window arrays (even)
  int n_items;                 // B * Hp
  FastDiv d_hp;
  size_t plane_elems;
};

// Conv rows 2ph-1 .. 2ph+1 inside [0, H) feed pooled row ph. Row 2ph-1 is also row 2(ph-1)+1 of the previous pooled row: a CTA walks
// consecutive items, so unless the item is the CTA's first or starts an image (`fresh`) that row is still in the activation ring
// (slot (h + 1) % 3) and only two new rows are convolved.
__device__ __forceinline__ int first_kh(bool fresh) { return fresh ? 0 : 1; }
__device__ __forceinline__ int tiles_of(int ph, int H, bool fresh) {
  int n = 0;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh) n += (kh >= first_kh(fresh) && (unsigned)(2 * ph - 1 + kh) < (unsigned)H) ? 1 : 0;
  return n;
}
__device__ __forceinline__ int act_slot(int h) { return (h + 1) % 3; }

__global__ void __launch_bounds__(THREADS, 1) stem_fwd_kernel(const Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* stages = smem;                                              // [NSTAGE][STAGE]
  unsigned char* b_tile = smem + (size_t)NSTAGE * STAGE;                     // [2 * 64 rows][128 B]
  float* act = reinterpret_cast<float*>(b_tile + B_BYTES);                   // [3][64][WLD]
  __half* win = reinterpret_cast<__half*>(act + 3 * CO * p.WLD);             // [NGROUPS][4 arrays: hi0 | hi1 | lo0 | lo1][9][LDH] fp16
  uint64_t* full = reinterpret_cast<uint64_t*>(win + (size_t)NGROUPS * 4 * 9 * p.LDH + 8);
  uint64_t* empty = full + NSTAGE;
  uint64_t* acc_full = empty + NSTAGE;        // [NSLOT]
  uint64_t* acc_empty = acc_full + NSLOT;     // [NSLOT]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + NSLOT);
  float* s_co = reinterpret_cast<float*>(tmem_slot + 4);                     // [3][64] bias | scale | shift (16-byte aligned: read as float4)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // weights -> K-major SWIZZLE_128B fp16 hi / lo rows. The reduction index is k = 8 * tr + ts (tap row tr = 16-byte chunk, tap
  // column ts < 7 inside it; k = 8 tr + 7 and chunk 7 are zero): a chunk of the A operand is then 8 CONSECUTIVE input samples
  for (int i = tid; i < CO * 8; i += THREADS) {
    const int o = i >> 3, j = i & 7;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = (j < KS && q < KS) ? p.w[o * NT + j * KS + q] : 0.f;
    uint4 h, l;
    split_f16x2(v[0], v[1], h.x, l.x); split_f16x2(v[2], v[3], h.y, l.y);
    split_f16x2(v[4], v[5], h.z, l.z); split_f16x2(v[6], v[7], h.w, l.w);
    *reinterpret_cast<uint4*>(b_tile + sw128_offset((uint32_t)o, (uint32_t)j)) = h;
    *reinterpret_cast<uint4*>(b_tile + CO * 128 + sw128_offset((uint32_t)o, (uint32_t)j)) = l;
  }
  for (int i = tid; i < CO; i += THREADS) {       // bn(y + b) = scale * y + (scale * b + shift)
    const float bsv = p.bias != nullptr ? p.bias[i] : 0.f;
    s_co[i] = bsv;
    s_co[CO + i] = p.scale[i];
    s_co[2 * CO + i] = fmaf(p.scale[i], bsv, p.shift[i]);
  }
  for (int i = tid; i < NGROUPS * 4 * 9 * p.LDH / 2; i += THREADS) reinterpret_cast<uint32_t*>(win)[i] = 0u;
  // activation rows: column 0 is w = -1, columns 1 .. W the pixels, the rest padding; padding holds -1 (< any ReLU output), so the
  // pooling windows need no bounds checks
  for (int i = tid; i < 3 * CO * p.WLD; i += THREADS) act[i] = -1.f;
  if (warp == PROD_WARPS) {
    if (lane == 0) {
      for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], NPROD); mbar_init(&empty[s], 1); }
      for (int s = 0; s < NSLOT; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EPI_WARPS); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per = (p.n_items + gridDim.x - 1) / gridDim.x;
  const int it0 = blockIdx.x * per, it1 = min(p.n_items, it0 + per);

  if (warp < PROD_WARPS) {
    // ================================================================================= producers
    const int group = warp >> 2, gt = tid & (NPROD - 1);
    const int j = gt & 7, pr = gt >> 3;
    // pre-split window of the current item: fp16 hi and lo*2^11 planes, each stored twice -- copy 0 with sample cc at half index
    // cc, copy 1 at cc + 1 -- so that the 8 consecutive samples of ANY operand chunk start on a 4-byte boundary in one of them
    const int asz = 9 * p.LDH;
    __half* whi0 = win + (size_t)group * 4 * asz;
    __half* whi1 = whi0 + asz;
    __half* wlo0 = whi1 + asz;
    __half* wlo1 = wlo0 + asz;
    const int odd = pr & 1;                                   // parity of my rows w = pr + 16 i
    const __half* my_hi = odd ? whi1 : whi0;
    const __half* my_lo = odd ? wlo1 : wlo0;
    int tile_idx = 0;       // global tile counter of this CTA at the start of the current item
    for (int it = it0; it < it1; ++it) {
      uint32_t b, ph;
      p.d_hp.divmod((uint32_t)it, b, ph);
      const bool fresh = it == it0 || ph == 0u;
      const int nt = tiles_of((int)ph, p.H, fresh);
      if (((it - it0) & (NGROUPS - 1)) == group) {
        // ---- window: input rows 2ph-4 .. 2ph+4, columns -3 .. W+2 (zero outside), split once per sample
        const float* img = p.x + (size_t)b * p.H * p.W;
        const int h0 = 2 * (int)ph - 4;
        asm volatile("bar.sync %0, 128;" ::"r"(2 + group) : "memory");
        const int rr0 = first_kh(fresh);                      // window row 0 only feeds conv row 2ph-1
        for (int e = gt + rr0 * (p.W + 6); e < 9 * (p.W + 6); e += NPROD) {
          const int rr = e / (p.W + 6), cc = e - rr * (p.W + 6);
          const int hi = h0 + rr, wi = cc - 3;
          const float v = ((unsigned)hi < (unsigned)p.H && (unsigned)wi < (unsigned)p.W) ? img[hi * p.W + wi] : 0.f;
          const __half vh = __float2half_rn(v);
          const __half vl = __float2half_rn((v - __half2float(vh)) * kF16LoScale);
          const int o0 = rr * p.LDH + cc;
          whi0[o0] = vh; whi1[o0 + 1] = vh; wlo0[o0] = vl; wlo1[o0 + 1] = vl;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(2 + group) : "memory");
        int k = 0;
#pragma unroll 1
        for (int kh = rr0; kh < 3; ++kh) {
          const int h = 2 * (int)ph - 1 + kh;
          if ((unsigned)h >= (unsigned)p.H) continue;
          const int t_glob = tile_idx + k;
          ++k;
          const int s = t_glob % NSTAGE;
          const uint32_t phs = (uint32_t)(t_glob / NSTAGE) & 1u;
          // conv pixel (h, w), chunk j = tap row tr: input samples (h + tr - 3, w - 3 .. w + 4) = window row kh + tr, columns w .. w + 7
          uint4 hh[8], ll[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int w = pr + 16 * i;
            hh[i] = make_uint4(0u, 0u, 0u, 0u);
            ll[i] = hh[i];
            if (j < KS && w < p.W) {
              const int o0 = (kh + j) * p.LDH + w + odd;          // even half index
              const uint32_t* ph_ = reinterpret_cast<const uint32_t*>(my_hi + o0);
              const uint32_t* pl_ = reinterpret_cast<const uint32_t*>(my_lo + o0);
              hh[i] = make_uint4(ph_[0], ph_[1], ph_[2], ph_[3]);
              ll[i] = make_uint4(pl_[0], pl_[1], pl_[2], pl_[3]);
            }
          }
          // Two producer groups share one stage ring, so a group may reach use n of a stage while use n-1 (the other group's) has
          // not even been filled; a parity wait on `empty` alone would then alias to an older phase. Waiting first for use n-1 to
          // be FILLED pins the phase: by then `empty` has completed exactly the phases 0 .. n-2.
          if (t_glob >= NSTAGE) mbar_wait(&full[s], (uint32_t)(t_glob / NSTAGE - 1) & 1u);
          mbar_wait(&empty[s], phs ^ 1u);
          unsigned char* a_hi = stages + (size_t)s * STAGE;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t off = sw128_offset((uint32_t)(pr + 16 * i), (uint32_t)j);
            *reinterpret_cast<uint4*>(a_hi + off) = hh[i];
            *reinterpret_cast<uint4*>(a_hi + A_PART + off) = ll[i];
          }
          fence_proxy_async();
          mbar_arrive(&full[s]);
        }
      }
      tile_idx += nt;
    }
  } else if (warp == PROD_WARPS) {
    // ================================================================================= MMA issuer
    if (lane == 0) {
      const uint32_t idesc = instr_desc(0u, 128, CO), idesc2 = instr_desc(0u, 128, 2 * CO);
      const uint64_t b_hi = smem_desc_sw128(smem_u32(b_tile));
      int t_glob = 0;
      for (int it = it0; it < it1; ++it) {
        uint32_t b, ph;
        p.d_hp.divmod((uint32_t)it, b, ph);
        const int nt = tiles_of((int)ph, p.H, it == it0 || ph == 0u);
        for (int k = 0; k < nt; ++k, ++t_glob) {
          const int s = t_glob % NSTAGE, slot = t_glob % NSLOT;
          mbar_wait(&acc_empty[slot], ((uint32_t)(t_glob / NSLOT) & 1u) ^ 1u);
          mbar_wait(&full[s], (uint32_t)(t_glob / NSTAGE) & 1u);
          tc_fence_after();
          const uint32_t base = smem_u32(stages + (size_t)s * STAGE);
          const uint64_t a_hi = smem_desc_sw128(base), a_lo = smem_desc_sw128(base + A_PART);
          const uint32_t d = tmem_base + (uint32_t)slot * 128u;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            mma_bf16(d, a_hi + (uint64_t)(kk * 2), b_hi + (uint64_t)(kk * 2), idesc2, kk > 0 ? 1u : 0u);
            mma_bf16(d + CO, a_lo + (uint64_t)(kk * 2), b_hi + (uint64_t)(kk * 2), idesc, 1u);
          }
          mma_commit(&empty[s]);
          mma_commit(&acc_full[slot]);
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================================= epilogue: normalise, rectify, pool
    // 8 warps: TMEM lane quarter = warp % 4 (hardware rule), two warps per quarter split the 64 channels
    const int quarter = warp & 3;
    const int chalf = (warp - PROD_WARPS - 1) >> 2;
    const int et = (warp - PROD_WARPS - 1) * 32 + lane;   // 0..255
    const int wpix = quarter * 32 + lane;                 // conv pixel (column) of my TMEM lane
    const int c0 = 32 * chalf;
    // pooling role: channel c, quarter rq of the pooled row
    const int c = et & 63, rq = et >> 6;
    const int per = (p.Wp + 3) >> 2;
    const int pw0 = rq * per, pw1 = min(p.Wp, pw0 + per);
    int t_glob = 0;
    for (int it = it0; it < it1; ++it) {
      uint32_t b, ph;
      p.d_hp.divmod((uint32_t)it, b, ph);
      // ---- part 1: the item's new conv rows from TMEM into act[slot][c][1 + w] (BatchNorm + ReLU applied)
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");      // previous item's pooling no longer reads act
#pragma unroll 1
      for (int kh = first_kh(it == it0 || ph == 0u); kh < 3; ++kh) {
        const int h = 2 * (int)ph - 1 + kh;
        if ((unsigned)h >= (unsigned)p.H) continue;
        const int slot = t_glob % NSLOT;
        mbar_wait(&acc_full[slot], (uint32_t)(t_glob / NSLOT) & 1u);
        tc_fence_after();
        ++t_glob;
        const uint32_t t_row = tmem_base + (uint32_t)slot * 128u + ((uint32_t)(quarter * 32) << 16);
        float* arow = act + (size_t)act_slot(h) * CO * p.WLD + wpix + 1;
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(t_row + (uint32_t)c0, r0);
        tmem_ld_32x32(t_row + (uint32_t)(c0 + CO), r1);
        tmem_ld_wait();
        if (wpix < p.W) {
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            const float4 sc = *reinterpret_cast<const float4*>(s_co + CO + c0 + k), tt = *reinterpret_cast<const float4*>(s_co + 2 * CO + c0 + k);
            const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, ttv[4] = {tt.x, tt.y, tt.z, tt.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float y = fmaf(__uint_as_float(r1[k + q]), kF16LoInv, __uint_as_float(r0[k + q]));
              arow[(size_t)(c0 + k + q) * p.WLD] = fmaxf(fmaf(y, scv[q], ttv[q]), 0.f);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[slot]);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");      // act complete
      // ---- part 2: pooled row ph. Window of pooled pixel pw = act columns 2pw, 2pw+1, 2pw+2 (column 0 is w = -1): one 8-byte
      // load per row and pixel, the third column is the first of the next pair. Rows scanned in order with a strict compare =
      // first maximum in (kh, kw) scan order (PyTorch's tie rule).
      const float* rowp[3];
      bool rv[3];
      float2 cur[3];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int h = 2 * (int)ph - 1 + kh;
        rv[kh] = (unsigned)h < (unsigned)p.H;
        rowp[kh] = act + ((size_t)act_slot(rv[kh] ? h : 0) * CO + c) * p.WLD;
        cur[kh] = (rv[kh] && pw0 < pw1) ? *reinterpret_cast<const float2*>(rowp[kh] + 2 * pw0) : make_float2(-1.f, -1.f);
      }
      size_t o = (((size_t)b * p.Hp + ph) * p.Wp + pw0) * CO + c;
      for (int pw = pw0; pw < pw1; ++pw, o += CO) {
        float best = -INFINITY;
        int arg = 0;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          if (!rv[kh]) continue;
          const float2 nx = *reinterpret_cast<const float2*>(rowp[kh] + 2 * pw + 2);
          if (cur[kh].x > best) { best = cur[kh].x; arg = kh * 3; }
          if (cur[kh].y > best) { best = cur[kh].y; arg = kh * 3 + 1; }
          if (nx.x > best) { best = nx.x; arg = kh * 3 + 2; }
          cur[kh] = nx;
        }
        p.p0[o] = best;
        if (p.argmax != nullptr) p.argmax[o] = (uint8_t)arg;
        if (p.planes != nullptr) {
          const __half hi = __float2half_rn(best);
          const __half lo = __float2half_rn((best - __half2float(hi)) * kF16LoScale);
          reinterpret_cast<__half*>(p.planes)[o] = hi;
          reinterpret_cast<__half*>(p.planes)[p.plane_elems + o] = lo;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == PROD_WARPS) tmem_dealloc(tmem_base, 512);
}

// stats[0][o] = sum_p y0, stats[1][o] = sum_p y0^2 from the Gram matrix / tap sums of the input patches (fp64); block = channel
__global__ void __launch_bounds__(64) stem_stats_kernel(const double* __restrict__ G, const double* __restrict__ X1, const float* __restrict__ w,
                                                        const float* __restrict__ bias, double M, double* __restrict__ stats, const PcBnFinalize fin) {
  __shared__ double s_w[NT], s_a[64], s_b[64];
  const int o = blockIdx.x, t = threadIdx.x;
  if (t < NT) s_w[t] = (double)w[o * NT + t];
  __syncthreads();
  double a = 0.0, q = 0.0;
  if (t < NT) {
    double r = 0.0;
    for (int u = 0; u < NT; ++u) r = fma(G[t * NT + u], s_w[u], r);
    q = s_w[t] * r;
    a = s_w[t] * X1[t];
  }
  s_a[t] = a; s_b[t] = q;
  __syncthreads();
  if (t == 0) {
    double wx = 0.0, wgw = 0.0;
    for (int u = 0; u < NT; ++u) { wx += s_a[u]; wgw += s_b[u]; }
    const double b = bias != nullptr ? (double)bias[o] : 0.0;
    const double sy = wx + M * b, sy2 = wgw + 2.0 * b * wx + M * b * b;
    stats[o] = sy;
    stats[CO + o] = sy2;
    if (fin.scale != nullptr) {        // BatchNorm finalisation of this channel right here (one launch less before the stem forward)
      float mean, invstd;
      double unbiased;
      bn_train_coeffs_v(sy, sy2, M, fin.eps, mean, invstd, unbiased);
      bn_publish_channel(fin, o, mean, invstd, unbiased);
      if (o == 0 && fin.num_batches_tracked != nullptr) fin.num_batches_tracked[0] += 1;
    }
  }
}

}  // namespace stemf
}  // namespace pc

using namespace pc;
using namespace pc::stemf;

extern "C" int pc_stem_fwd_supported(int k, int Cout, int H, int W) { return (k == 7 && Cout == 64 && H >= 3 && W >= 7 && W <= 128) ? 1 : 0; }

extern "C" int pc_stem_stats_from_gram(const double* G, const double* X1, const float* w_oihw, const float* bias, int B, int H, int W,
                                       double* stats, const PcBnFinalize* fin, pc_stream_t stream) {
  PC_REQUIRE(G && X1 && w_oihw && stats, PC_EINVAL, "pc_stem_stats_from_gram: null pointer");
  PC_REQUIRE(fin == nullptr || (fin->scale && fin->shift), PC_EINVAL, "pc_stem_stats_from_gram: PcBnFinalize needs scale / shift outputs");
  stem_stats_kernel<<<CO, 64, 0, stream>>>(G, X1, w_oihw, bias, (double)B * H * W, stats, fin != nullptr ? *fin : PcBnFinalize{});
  PC_LAUNCH_CHECK("stem_stats_kernel");
  return PC_OK;
}

// x [B][H][W] -> p0 [B][Hp][Wp][64] = maxpool3x3s2p1(relu(scale * (conv7x7(x) + bias) + shift)); argmax (may be NULL) receives the
// window position kh*3+kw of each maximum; planes (may be NULL) the fp16 hi | lo planes of p0 (pc_bn_act_split layout).
extern "C" int pc_stem_fwd(const float* x, const float* w_oihw, const float* bias, const float* scale, const float* shift, int B, int H,
                           int W, float* p0, uint8_t* argmax, void* planes, pc_stream_t stream) {
  PC_REQUIRE(x && w_oihw && scale && shift && p0 && B > 0, PC_EINVAL, "pc_stem_fwd: bad arguments");
  PC_REQUIRE(pc_stem_fwd_supported(7, 64, H, W), PC_EUNSUPPORTED, "pc_stem_fwd: image %dx%d not covered (one 128-pixel tile per row)", H, W);
  Params p{};
  p.x = x; p.w = w_oihw; p.bias = bias; p.scale = scale; p.shift = shift; p.p0 = p0; p.argmax = argmax;
  p.planes = static_cast<unsigned char*>(planes);
  p.B = B; p.H = H; p.W = W; p.Hp = (H + 2 - 3) / 2 + 1; p.Wp = (W + 2 - 3) / 2 + 1;
  p.WLD = 2 * ((p.Wp + 1) | 1);          // 2 * odd (conflict-free 8-byte loads across channels), >= 2 Wp + 2 columns
  p.LDH = (W + 6 + 8 + 2 + 1) & ~1;
  p.n_items = B * p.Hp;
  p.d_hp = FastDiv::make((uint32_t)p.Hp);
  p.plane_elems = (size_t)B * p.Hp * p.Wp * CO;
  const size_t smem = (size_t)NSTAGE * STAGE + B_BYTES + sizeof(float) * (3 * (size_t)CO * p.WLD) + sizeof(__half) * ((size_t)NGROUPS * 4 * 9 * p.LDH + 8) +
                      sizeof(uint64_t) * (2 * NSTAGE + 2 * NSLOT + 1) + 16 + sizeof(float) * 3 * CO + 1024;
  PC_REQUIRE(smem <= 227 * 1024, PC_EUNSUPPORTED, "pc_stem_fwd: row width %d needs %zu B of shared memory", W, smem);
  static size_t conf = 0;
  if (smem > conf) {
    PC_CUDA(cudaFuncSetAttribute(stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
  stem_fwd_kernel<<<grid, THREADS, smem, stream>>>(p);
  PC_LAUNCH_CHECK("stem_fwd_kernel");
  return PC_OK;
}
