// Halo-resident tcgen05 convolution engine for sm_100a: stride-1 3x3 "same" convolutions (forward and data gradient) on the
// FP16X2 operand planes, persistent CTAs, TMA-fed, double-buffered TMEM, weight tiles multicast across a thread-block cluster.
//
// Why. ncu / timing of the round-1 kernels (one 128-pixel tile per CTA, the A operand re-gathered for each of the 9 taps) put
// every layer at 6-8 TB/s of L2 -> shared-memory traffic with the tensor pipe 22-54 % busy: the kernels are bound by operand
// traffic out of L2, not by the tensor cores. Per 128 x BN x 64 k-chunk they move 32 KB of A (hi + lo) and 2 * BN * 128 B of B.
// This engine removes most of both:
//   A  The activation planes [B][H][W][C] are addressed in a PADDED position space q = (b * (H+1) + hp) * (W+1) + wp, where
//      hp = 0 is a zero row above each image and wp = 0 a zero column left of each row (the right neighbour of a row's last
//      pixel is the next row's zero column; the row below an image is the next image's zero row). Tap (dr, ds) of position q
//      is then position q + dr * (W+1) + ds -- a pure ROW SHIFT. A tile's 128 positions plus a halo of W+2 positions on either
//      side are loaded ONCE per 64-channel chunk by a few tiled TMA boxes {64 ch, W+1, RB rows} starting at w = -1 / h = -1
//      (out-of-bounds elements are zero-filled by the TMA unit, which materialises the padding without storing it in HBM), and
//      all 9 taps issue their MMAs from the same shared-memory region through UMMA descriptors whose start address is advanced
//      by the tap's row shift (SWIZZLE_128B is a function of absolute shared-memory address bits for both the TMA write and the
//      UMMA read, so a start address that is only 128-byte aligned is legal; checked on hardware by profiles/tools/
//      probe_umma_tma.cu). A traffic drops from 9 x 128 rows to (128 + 2 (W+2) + box rounding) rows per chunk.
//      Positions with hp = 0 or wp = 0 are computed and discarded (7 % of the rows at 20x51 ... 34 % at 3x7).
//   B  The CTAs of a cluster work on different position tiles of the same output-channel tile in lockstep; each loads 1/CL of
//      every weight stage and multicasts it into all CL shared memories (cp.async.bulk ... .multicast::cluster), the stage is
//      released by tcgen05.commit arriving on the `empty` barrier of every CTA of the cluster.
// The accumulators of tile i+1 are produced while the epilogue warps drain tile i from the other TMEM buffer.
//
// Warp roles (10 warps): 0 = TMA / bulk-copy producer (one lane), 1 = TMEM allocation + MMA issue (one lane), 2-9 = epilogue
// (TMEM lane quarter = warp % 4, two warps per quarter splitting the columns): TMEM -> registers -> (main + 2^-11 corr) * scale + bias -> NHWC fp32 store for real pixels,
// BatchNorm sum / sum-of-squares accumulated per CTA in shared memory over all its tiles and flushed once (fp64 atomics).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pc {
namespace halo {

using namespace pc::tc;

constexpr int BM = 128;
constexpr int MAX_BOX = 8;
constexpr int MAX_BST = 8;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 64 + 32 * EPI_WARPS;

struct Params {
  const unsigned char* Bp;   // packed weights [kc = tap * cpt + chunk][hi | lo][Npad][128 B] (pc_pack_conv_weight_tc, FP16X2)
  const float* bias;
  const float* a_amax;       // dgrad: device scalar behind the power-of-two scale of the dy planes (null: 1)
  float* C;                  // [B][H][W][Nn] fp32
  double* stats;             // [2][Nn] or null
  int B, H, W, Ca, Nn, Npad;
  int Wp, Pimg, RB, box_pos, nbox;
  long long Q;
  int n_mtiles, n_ntiles, n_items, cpt, accumulate, cluster;
  int tap_shift[9];          // row shift of weight tap r * 3 + s
  int ntaps;                 // 9 (3x3) or 1 (1x1 shortcut convolutions)
  int halo;                  // positions staged on either side of a tile: W + 2 (3x3) or 0 (1x1)
  int a_stride;              // 1x1 forward with stride 2: the activation boxes sample every a_stride-th pixel (TMA elementStrides)
  int Hout, Wout, o_stride;  // output image and pixel stride: position (hp, wp) -> pixel (o_stride (hp-1) + o_off_h, o_stride (wp-1) + o_off_w)
  int o_off_h, o_off_w;      // (stride-2 data gradients: o_stride = 2, offsets = the output parity class; pixels past the image are dropped)
  int tap_id[9];             // weight tap (r * 3 + s) behind issued tap t (a stride-2 data gradient issues 1 / 2 / 2 / 4 of the nine per class)
  int b_stages;
  uint32_t region_bytes;     // one part (hi or lo) of an A region: nbox * box_pos * 128, rounded up to 1024
  uint32_t a_tx_bytes;       // bytes the TMA boxes of one region deliver (both parts, unrounded)
  FastDiv d_pimg, d_wp, d_box, d_hp1, d_nt;
  // Fused BatchNorm-backward REDUCE pass (data gradient of conv2 -> backward of bn1, pool-free): with red.y set the epilogue also forms
  // dz = [bn(y) > 0] * dA * drop and xhat = (y - mean) * invstd for every pixel it has just produced and accumulates sum dz, sum dz xhat
  // per channel plus max |dz|, max |xhat| -- what pc_bn_act_bwd_reduce would re-read dA and y from HBM for.
  PcBnBwdReduce red;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t n_clusters_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 5-D tiled TMA load (coordinates innermost first: channel, w, h, image, plane), completion in bytes on `bar`
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
// same box brought into L2 only (no shared memory, no barrier). Tried one tile / two chunks ahead of the real loads: every layer got
// SLOWER (64ch 58 -> 68 us, 128ch 77 -> 89, 512ch 107 -> 124), so the kernels do not use it; kept for experiments.
__device__ __forceinline__ void tma_prefetch_5d(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global [%0, {%1, %2, %3, %4, %5}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
// 1-D bulk copy replicated into the same shared-memory offset of every CTA in `mask`; each destination's mbarrier (same offset)
// receives the byte count
__device__ __forceinline__ void bulk_g2s_mcast(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mma_commit_mcast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

__device__ __forceinline__ float warp_reduce_scatter32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// Epilogue of both kernels: EPI_WARPS = 8 warps (2 .. 9); warp w reads TMEM lanes 32 * (w % 4) .. + 31 (its hardware lane quarter =
// 32 tile rows) and the 32-column chunks (w - 2) / 4, + 2, ...: two warps share a lane quarter and split the columns, so each
// scheduler hosts two epilogue warps whose TMEM loads / global stores overlap. Per item: wait for the accumulator buffer, combine
// (main + 2^-11 corr) * scale + bias, store the rows that are real pixels, accumulate the BatchNorm sums, release the buffer.
// BatchNorm sums: BN = 64 layers have one output-channel tile and up to 15 tiles per CTA with a 3.5 k-cycle main loop, so a
// 62-shuffle cross-row reduction per chunk and tile would make the epilogue the bottleneck; each thread instead keeps its row's
// running sum / sum of squares of its 32 columns in registers over ALL its tiles and the cross-row reduction happens once per CTA.
template <int BN, int NACC, int NBUF, bool RED>
__device__ __forceinline__ void epilogue_warps(const Params& p, uint32_t tmem_base, float a_scale, uint64_t* acc_full, uint64_t* acc_empty,
                                               float* s_sum, float* s_sq, const float* s_bias, uint32_t cl_id, uint32_t n_cl, int CL, uint32_t rank) {
  constexpr uint32_t TM_BUF = NACC * BN;
  constexpr bool REGSTATS = (BN <= 64);
  constexpr int NCH = BN >= 64 ? BN / 64 : 1;        // 32-column chunks per warp and tile (BN = 32: one chunk, read by the `half` 0 warps only)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3;
  const int half = (warp - 2) >> 2;
  const int row = quarter * 32 + lane;
  const float out_scale = 1.f / a_scale;
  const bool want_red = RED;                                // fused BatchNorm-backward reduce (see Params::red); never accumulates into C
  const bool want_stats = p.stats != nullptr || want_red;   // the same accumulators serve (sum v, sum v^2) or (sum dz, sum dz xhat)
  const float* s_rscale = s_bias + p.Npad;                  // [Npad] each, filled by the kernel prologue when want_red
  const float* s_rshift = s_rscale + p.Npad;
  const float* s_rmean = s_rshift + p.Npad;
  const float* s_rinv = s_rmean + p.Npad;
  float mx_dz = 0.f, mx_xh = 0.f;
  float rs[REGSTATS ? 32 : 1], rq[REGSTATS ? 32 : 1];
#pragma unroll
  for (int k = 0; k < (REGSTATS ? 32 : 1); ++k) rs[k] = rq[k] = 0.f;
  uint32_t ti = 0;
  for (uint32_t item = cl_id; item < (uint32_t)p.n_items; item += n_cl, ++ti) {
    uint32_t mg, nt;
    p.d_nt.divmod(item, mg, nt);
    const int m_tile = (int)(mg * (uint32_t)CL + rank);
    const int n0 = (int)nt * BN;
    const long long q = (long long)m_tile * BM + row;
    bool valid = q < p.Q;
    long long pix = 0;
    uint32_t b = 0;
    if (valid) {
      uint32_t rem, hp, wp;
      p.d_pimg.divmod((uint32_t)q, b, rem);
      p.d_wp.divmod(rem, hp, wp);
      const int oh = p.o_stride * ((int)hp - 1) + p.o_off_h, ow = p.o_stride * ((int)wp - 1) + p.o_off_w;
      valid = hp >= 1u && wp >= 1u && oh < p.Hout && ow < p.Wout;
      pix = ((long long)b * p.Hout + oh) * p.Wout + ow;
    }
    float* dst_row = p.C + (size_t)(valid ? pix : 0) * p.Nn + n0;
    const uint32_t tb = ti % NBUF;
    // accumulate: the values already in the output row are fetched BEFORE waiting for the accumulator (their DRAM latency used to sit
    // in the epilogue, 57 -> 96 us for the 64-channel data gradient that adds into the shortcut gradient)
    float4 oldv[8];
    auto load_old = [&](int c0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        oldv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!RED && p.accumulate && valid && n0 + c0 + 4 * k < p.Nn) oldv[k] = *reinterpret_cast<const float4*>(dst_row + c0 + 4 * k);
      }
    };
    const bool idle = BN == 32 && half == 1;                 // BN = 32: the second warp of each lane quarter has no columns
    if (!idle) load_old(32 * half);
    // fused reduce: this row's y chunk (and dropout multipliers) are fetched before waiting for the accumulator as well
    const float* y_row = want_red ? p.red.y + (size_t)(valid ? pix : 0) * p.Nn + n0 : nullptr;
    const float* d_row = (want_red && p.red.drop != nullptr) ? p.red.drop + (size_t)b * p.Nn + n0 : nullptr;
    float4 yv[8];
    auto load_y = [&](int c0) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        yv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (RED && valid && n0 + c0 + 4 * k < p.Nn) yv[k] = *reinterpret_cast<const float4*>(y_row + c0 + 4 * k);
      }
    };
    if (!idle) load_y(32 * half);
    mbar_wait(&acc_full[tb], (ti / NBUF) & 1u);
    tc_fence_after();
    const uint32_t t_row = tmem_base + tb * TM_BUF + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
    for (int ci = 0; ci < (idle ? 0 : NCH); ++ci) {
      const int c0 = 32 * (half + 2 * ci);
      if (ci > 0) { load_old(c0); load_y(c0); }
      float v[32], u[32];
      {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(t_row + (uint32_t)c0, r0);
        tmem_ld_32x32(t_row + (uint32_t)(c0 + BN), r1);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) { v[k] = __uint_as_float(r0[k]); u[k] = __uint_as_float(r1[k]); }
      }
      if (NACC == 4) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(t_row + (uint32_t)(c0 + 2 * BN), r0);
        tmem_ld_32x32(t_row + (uint32_t)(c0 + 3 * BN), r1);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) { v[k] += __uint_as_float(r0[k]); u[k] += __uint_as_float(r1[k]); }
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        const float acc = fmaf(u[k], kF16LoInv, v[k]) * out_scale;
        v[k] = valid ? acc + s_bias[n0 + c0 + k] : 0.f;
      }
      if (valid) {
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
          if (n0 + c0 + k < p.Nn) {
            float4 o = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
            float* d = dst_row + c0 + k;
            if (!RED && p.accumulate) {
              const float4 old = oldv[k >> 2];
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *reinterpret_cast<float4*>(d) = o;
          }
        }
      }
      if (want_red) {
        // v[k] = dA of a real pixel (0 otherwise) -> v[k] = dz, u[k] = dz * xhat
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int c = n0 + c0 + k;
          const float4 y4 = yv[k >> 2];
          const float yy = (k & 3) == 0 ? y4.x : ((k & 3) == 1 ? y4.y : ((k & 3) == 2 ? y4.z : y4.w));
          const bool on = valid && c < p.Nn && fmaf(yy, s_rscale[c], s_rshift[c]) > 0.f;
          const float dr = d_row != nullptr && valid && c < p.Nn ? __ldg(d_row + c0 + k) : 1.f;
          const float dzv = on ? v[k] * dr : 0.f;
          const float xh = (valid && c < p.Nn) ? (yy - s_rmean[c]) * s_rinv[c] : 0.f;
          v[k] = dzv;
          u[k] = dzv * xh;
          mx_dz = fmaxf(mx_dz, fabsf(dzv));
          mx_xh = fmaxf(mx_xh, fabsf(xh));
        }
      }
      if (want_stats) {
        if (REGSTATS) {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            rs[k & (REGSTATS ? 31 : 0)] += v[k];
            rq[k & (REGSTATS ? 31 : 0)] = want_red ? rq[k & (REGSTATS ? 31 : 0)] + u[k] : fmaf(v[k], v[k], rq[k & (REGSTATS ? 31 : 0)]);
          }
        } else {
          float sq[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) sq[k] = want_red ? u[k] : v[k] * v[k];
          const float cs = warp_reduce_scatter32(v, lane);
          const float cq = warp_reduce_scatter32(sq, lane);
          atomicAdd(&s_sum[n0 + c0 + lane], cs);
          atomicAdd(&s_sq[n0 + c0 + lane], cq);
        }
      }
    }
    // this warp has finished reading the buffer: hand it back to the MMA issuer
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&acc_empty[tb]);
  }
  if (want_stats) {
    if (REGSTATS && !(BN == 32 && half == 1)) {        // one cross-row reduction per CTA (n_ntiles == 1: channel = 32 * half + lane)
      float a[32], b[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) { a[k] = rs[k & (REGSTATS ? 31 : 0)]; b[k] = rq[k & (REGSTATS ? 31 : 0)]; }
      const float cs = warp_reduce_scatter32(a, lane);
      const float cq = warp_reduce_scatter32(b, lane);
      atomicAdd(&s_sum[32 * half + lane], cs);
      atomicAdd(&s_sq[32 * half + lane], cq);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
    double* dst = want_red ? p.red.sums : p.stats;
    for (int i = tid - 64; i < p.Nn; i += 32 * EPI_WARPS) {
      const float a = s_sum[i], b = s_sq[i];
      if (a != 0.f || b != 0.f) {
        atomicAdd(dst + i, (double)a);
        atomicAdd(dst + p.Nn + i, (double)b);
      }
    }
    if (want_red && p.red.maxes != nullptr) {        // non-negative floats order like their bit patterns
      mx_dz = warp_max(mx_dz);
      mx_xh = warp_max(mx_xh);
      if (lane == 0) {
        atomicMax(reinterpret_cast<unsigned int*>(p.red.maxes), __float_as_uint(mx_dz));
        atomicMax(reinterpret_cast<unsigned int*>(p.red.maxes + 1), __float_as_uint(mx_xh));
      }
    }
  }
}

// NACC accumulators of BN columns per TMEM buffer: [main | corr] (2) or [main0 | corr0 | main1 | corr1] alternating per k-step (4).
// The tensor core adds into its fp32 accumulator with truncation; over a K = 4608 reduction a single [main | corr] pair measured
// 5.6e-6 of max|y| against fp64 where the alternating sets give ~2e-6, so the deep-K layers (>= 256 gathered channels) use NACC = 4
// with ONE TMEM buffer (their tiles run 36-72 k-chunks, the un-overlapped epilogue is < 5 % of a tile) and the shallow ones two
// buffers (the epilogue of tile i hides under the main loop of tile i + 1).
template <int BN, int NACC, int NBUF, bool RED>
__global__ void __launch_bounds__(THREADS, 1) conv_halo_kernel(const __grid_constant__ CUtensorMap amap, const Params p) {
  constexpr uint32_t B_STAGE = 2u * BN * 128u;          // hi rows then lo rows
  constexpr uint32_t TM_BUF = NACC * BN;                 // TMEM columns per accumulator buffer
  static_assert(NBUF * TM_BUF <= 512, "the accumulator buffers must fit the 512 TMEM columns");

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t a_buf_bytes = 2u * p.region_bytes;
  unsigned char* a_buf = smem;                                                // [2][hi region | lo region]
  unsigned char* b_buf = smem + 2 * (size_t)a_buf_bytes;                      // [b_stages][B_STAGE]   (a_buf_bytes is a multiple of 1024)
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + (size_t)p.b_stages * B_STAGE);
  uint64_t* a_full = bars;              // [2]
  uint64_t* a_empty = bars + 2;         // [2]
  uint64_t* b_full = bars + 4;          // [MAX_BST]
  uint64_t* b_empty = b_full + MAX_BST; // [MAX_BST]
  uint64_t* acc_full = b_empty + MAX_BST;   // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_sum = reinterpret_cast<float*>(tmem_slot + 4);   // [Npad]
  float* s_sq = s_sum + p.Npad;                             // [Npad]
  float* s_bias = s_sq + p.Npad;                            // [Npad]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int CL = p.cluster;
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  const uint32_t cl_id = CL > 1 ? cluster_id_x() : blockIdx.x;
  const uint32_t n_cl = CL > 1 ? n_clusters_x() : gridDim.x;
  const uint16_t cl_mask = (uint16_t)((1u << CL) - 1u);

  for (int i = tid; i < p.Npad; i += THREADS) {
    s_sum[i] = 0.f;
    s_sq[i] = 0.f;
    s_bias[i] = (p.bias != nullptr && i < p.Nn) ? p.bias[i] : 0.f;
    if (RED) {
      const bool in = i < p.Nn;
      s_bias[p.Npad + i] = in ? p.red.scale[i] : 0.f;
      s_bias[2 * p.Npad + i] = in ? p.red.shift[i] : 0.f;
      s_bias[3 * p.Npad + i] = in ? p.red.mean[i] : 0.f;
      s_bias[4 * p.Npad + i] = in ? p.red.invstd[i] : 0.f;
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&a_full[i], 1);
        mbar_init(&a_empty[i], 1);
        mbar_init(&acc_full[i], 1);
        mbar_init(&acc_empty[i], EPI_WARPS);
      }
      for (int i = 0; i < p.b_stages; ++i) {
        mbar_init(&b_full[i], 1);
        mbar_init(&b_empty[i], (uint32_t)CL);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, NBUF * TM_BUF);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (CL > 1) cluster_sync_all();     // every CTA's barriers exist before a peer multicasts into them / arrives on them
  const uint32_t tmem_base = *tmem_slot;
  const float a_scale = p.a_amax != nullptr ? f16_operand_scale(p.a_amax[0]) : 1.f;

  if (warp == 0) {
    // ================================================================================= producer
    if (lane == 0) {
      const size_t kc_stride = (size_t)2 * p.Npad * 128;
      // my share of a weight stage: CL == 1: both parts (2 copies of BN rows); CL == 2: one part; CL == 4: half a part
      const int n_copy = CL == 1 ? 2 : 1;
      const uint32_t copy_rows = CL <= 2 ? (uint32_t)BN : (uint32_t)BN / 2u;
      uint32_t ai = 0, bi = 0;     // running A-region / B-stage counters (phase = (count / depth) & 1)
      for (uint32_t item = cl_id; item < (uint32_t)p.n_items; item += n_cl) {
        uint32_t mg, nt;
        p.d_nt.divmod(item, mg, nt);
        const int m_tile = (int)(mg * (uint32_t)CL + rank);
        const int n0 = (int)nt * BN;
        const int q_lo = m_tile * BM - p.halo;
        const int bx0 = floor_div(q_lo, p.box_pos);
        for (int cc = 0; cc < p.cpt; ++cc) {
          // ---- A region of this chunk: nbox boxes of RB padded rows, hi and lo planes
          const uint32_t ab = ai & 1u;
          mbar_wait(&a_empty[ab], ((ai >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&a_full[ab], p.a_tx_bytes);
          unsigned char* dst = a_buf + (size_t)ab * a_buf_bytes;
          for (int j = 0; j < p.nbox; ++j) {
            const int rg = (bx0 + j) * p.RB;                    // global padded-row index of the box's first row
            const int b = floor_div(rg, p.H + 1);
            const int hp0 = rg - b * (p.H + 1);
            tma_load_5d(dst + (size_t)j * p.box_pos * 128, &amap, &a_full[ab], cc * 64, -p.a_stride, p.a_stride * (hp0 - 1), b, 0);
            tma_load_5d(dst + p.region_bytes + (size_t)j * p.box_pos * 128, &amap, &a_full[ab], cc * 64, -p.a_stride, p.a_stride * (hp0 - 1), b, 1);
          }
          ++ai;
          // ---- weight stages of this chunk, one per tap
          for (int t = 0; t < p.ntaps; ++t) {
            const uint32_t s = bi % (uint32_t)p.b_stages;
            mbar_wait(&b_empty[s], ((bi / (uint32_t)p.b_stages) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&b_full[s], B_STAGE);
            const unsigned char* src = p.Bp + (size_t)(p.tap_id[t] * p.cpt + cc) * kc_stride + (size_t)n0 * 128;
            unsigned char* bd = b_buf + (size_t)s * B_STAGE;
            for (int k = 0; k < n_copy; ++k) {
              const uint32_t row0 = (CL == 1 ? (uint32_t)k * BN : rank * copy_rows);      // row inside the [hi ; lo] stage
              const uint32_t part = row0 / BN, prow = row0 - part * BN;
              const unsigned char* sp = src + (size_t)part * p.Npad * 128 + (size_t)prow * 128;
              if (CL == 1) bulk_g2s(bd + (size_t)row0 * 128, sp, copy_rows * 128u, &b_full[s]);
              else bulk_g2s_mcast(bd + (size_t)row0 * 128, sp, copy_rows * 128u, &b_full[s], cl_mask);
            }
            ++bi;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================================= MMA issuer
    if (lane == 0) {
      const uint32_t idesc = instr_desc(0u, BM, BN);
      const uint32_t idesc2 = instr_desc(0u, BM, 2 * BN);
      uint32_t ai = 0, bi = 0, ti = 0;
      for (uint32_t item = cl_id; item < (uint32_t)p.n_items; item += n_cl) {
        uint32_t mg, nt;
        p.d_nt.divmod(item, mg, nt);
        const int m_tile = (int)(mg * (uint32_t)CL + rank);
        const int q_lo = m_tile * BM - p.halo;
        const int bx0 = floor_div(q_lo, p.box_pos);
        const int delta = m_tile * BM - bx0 * p.box_pos;          // row of the tile's first position inside the region
        const uint32_t tb = ti % NBUF;
        mbar_wait(&acc_empty[tb], ((ti / NBUF) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_buf = tmem_base + tb * TM_BUF;
        int ks = 0;
        for (int cc = 0; cc < p.cpt; ++cc) {
          const uint32_t ab = ai & 1u;
          mbar_wait(&a_full[ab], (ai >> 1) & 1u);
          tc_fence_after();
          const uint32_t a_base = smem_u32(a_buf + (size_t)ab * a_buf_bytes);
          for (int t = 0; t < p.ntaps; ++t) {
            const uint32_t s = bi % (uint32_t)p.b_stages;
            mbar_wait(&b_full[s], (bi / (uint32_t)p.b_stages) & 1u);
            tc_fence_after();
            const uint32_t a_row = a_base + (uint32_t)(delta + p.tap_shift[t]) * 128u;
            const uint64_t a_hi = smem_desc_sw128(a_row);
            const uint64_t a_lo = smem_desc_sw128(a_row + p.region_bytes);
            const uint32_t b_addr = smem_u32(b_buf + (size_t)s * B_STAGE);
            const uint64_t b_hi = smem_desc_sw128(b_addr);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk, ++ks) {
              const uint64_t adv = (uint64_t)(kk * 2);
              const uint32_t d_set = d_buf + (NACC == 4 ? (uint32_t)((ks & 1) * 2 * BN) : 0u);
              // a_hi * [b_hi ; b_lo] -> [main | corr] in one N = 2 BN instruction, then a_lo * b_hi into corr
              mma_bf16(d_set, a_hi + adv, b_hi + adv, idesc2, ks < NACC / 2 ? 0u : 1u);
              mma_bf16(d_set + BN, a_lo + adv, b_hi + adv, idesc, 1u);
            }
            if (CL == 1) mma_commit(&b_empty[s]);
            else mma_commit_mcast(&b_empty[s], cl_mask);
            ++bi;
          }
          mma_commit(&a_empty[ab]);
          ++ai;
        }
        mma_commit(&acc_full[tb]);
        ++ti;
      }
    }
    __syncwarp();
  } else {
    // ================================================================================= epilogue (warps 2..5)
    epilogue_warps<BN, NACC, NBUF, RED>(p, tmem_base, a_scale, acc_full, acc_empty, s_sum, s_sq, s_bias, cl_id, n_cl, CL, rank);
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (CL > 1) cluster_sync_all();     // no CTA leaves while a peer may still multicast into it or arrive on its barriers
  if (warp == 1) tmem_dealloc(tmem_base, NBUF * TM_BUF);
}

// ------------------------------------------------------------------------------------------------ weight-resident variant
// 64 -> 64 channel layers (cnn_deep block 0: the largest position count, 2184 tiles at 256 views). Measured with the streaming
// kernel above, their time follows bytes delivered into the SM -- 80 KB of activation region + 147 KB of weights per tile at
// ~23 B/clk/SM (the 6.5 TB/s aggregate L2 -> SM rate every kernel of this library tops out at) -- not the 3.5 k cycles of
// tensor work. All 9 taps' weights of such a layer are only 147 KB, so this variant keeps them RESIDENT in shared memory for the
// CTA's whole life and streams activations only. That leaves room for one activation region; the overlap a second buffer would
// give comes from phasing the two operand planes instead: the MMAs that read the hi plane (36 per tile, a_hi x [b_hi ; b_lo])
// are issued first, then the 36 that read the lo plane (a_lo x b_hi), and each plane has its own full / empty barrier pair, so
// the lo plane of tile i + 1 streams in under the hi MMAs of tile i + 1 and the hi plane of tile i + 1 under the lo MMAs of tile i.
template <int BN, int NACC, int NBUF, bool RED>
__global__ void __launch_bounds__(THREADS, 1) conv_halo_res_kernel(const __grid_constant__ CUtensorMap amap, const Params p) {
  constexpr uint32_t B_STAGE = 2u * BN * 128u;
  constexpr uint32_t TM_BUF = NACC * BN;
  static_assert(NBUF * TM_BUF <= 512, "the accumulator buffers must fit the 512 TMEM columns");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* a_buf = smem;                                                // [hi region | lo region]
  unsigned char* b_buf = smem + 2 * (size_t)p.region_bytes;                   // [9 taps][B_STAGE] resident
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_buf + (size_t)9 * B_STAGE);
  uint64_t* a_full = bars;              // [2] hi, lo
  uint64_t* a_empty = bars + 2;         // [2]
  uint64_t* b_full = bars + 4;          // [1]
  uint64_t* acc_full = bars + 6;        // [2]
  uint64_t* acc_empty = bars + 8;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* s_sum = reinterpret_cast<float*>(tmem_slot + 4);
  float* s_sq = s_sum + p.Npad;
  float* s_bias = s_sq + p.Npad;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < p.Npad; i += THREADS) {
    s_sum[i] = 0.f;
    s_sq[i] = 0.f;
    s_bias[i] = (p.bias != nullptr && i < p.Nn) ? p.bias[i] : 0.f;
    if (RED) {
      const bool in = i < p.Nn;
      s_bias[p.Npad + i] = in ? p.red.scale[i] : 0.f;
      s_bias[2 * p.Npad + i] = in ? p.red.shift[i] : 0.f;
      s_bias[3 * p.Npad + i] = in ? p.red.mean[i] : 0.f;
      s_bias[4 * p.Npad + i] = in ? p.red.invstd[i] : 0.f;
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&a_full[i], 1);
        mbar_init(&a_empty[i], 1);
        mbar_init(&acc_full[i], 1);
        mbar_init(&acc_empty[i], EPI_WARPS);
      }
      mbar_init(&b_full[0], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, NBUF * TM_BUF);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const float a_scale = p.a_amax != nullptr ? f16_operand_scale(p.a_amax[0]) : 1.f;
  const uint32_t part_tx = p.a_tx_bytes / 2u;

  if (warp == 0) {
    if (lane == 0) {
      // weights: once
      mbar_arrive_expect_tx(&b_full[0], 9u * B_STAGE);
      for (int t = 0; t < 9; ++t) {
        const unsigned char* src = p.Bp + (size_t)t * 2 * p.Npad * 128;
        bulk_g2s(b_buf + (size_t)t * B_STAGE, src, BN * 128u, &b_full[0]);
        bulk_g2s(b_buf + (size_t)t * B_STAGE + BN * 128u, src + (size_t)p.Npad * 128, BN * 128u, &b_full[0]);
      }
      uint32_t ti = 0;
      for (uint32_t item = blockIdx.x; item < (uint32_t)p.n_items; item += gridDim.x, ++ti) {
        const int m_tile = (int)item;
        const int bx0 = floor_div(m_tile * BM - p.Wp - 1, p.box_pos);
        for (int part = 0; part < 2; ++part) {
          mbar_wait(&a_empty[part], (ti & 1u) ^ 1u);
          mbar_arrive_expect_tx(&a_full[part], part_tx);
          unsigned char* dst = a_buf + (size_t)part * p.region_bytes;
          for (int j = 0; j < p.nbox; ++j) {
            const int rg = (bx0 + j) * p.RB;
            const int b = floor_div(rg, p.H + 1);
            const int hp0 = rg - b * (p.H + 1);
            tma_load_5d(dst + (size_t)j * p.box_pos * 128, &amap, &a_full[part], 0, -1, hp0 - 1, b, part);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = instr_desc(0u, BM, BN);
      const uint32_t idesc2 = instr_desc(0u, BM, 2 * BN);
      mbar_wait(&b_full[0], 0);
      tc_fence_after();
      const uint32_t b_base = smem_u32(b_buf);
      const uint32_t a_base = smem_u32(a_buf);
      uint32_t ti = 0;
      for (uint32_t item = blockIdx.x; item < (uint32_t)p.n_items; item += gridDim.x, ++ti) {
        const int m_tile = (int)item;
        const int bx0 = floor_div(m_tile * BM - p.Wp - 1, p.box_pos);
        const int delta = m_tile * BM - bx0 * p.box_pos;
        const uint32_t tb = ti % NBUF;
        mbar_wait(&acc_empty[tb], ((ti / NBUF) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_buf = tmem_base + tb * TM_BUF;
        // ---- phase 1: the hi plane against [b_hi ; b_lo] -> [main | corr]
        mbar_wait(&a_full[0], ti & 1u);
        tc_fence_after();
        int ks = 0;
        for (int t = 0; t < 9; ++t) {
          const uint64_t a_hi = smem_desc_sw128(a_base + (uint32_t)(delta + p.tap_shift[t]) * 128u);
          const uint64_t b_hi = smem_desc_sw128(b_base + (uint32_t)t * B_STAGE);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk, ++ks) {
            const uint32_t d_set = d_buf + (NACC == 4 ? (uint32_t)((ks & 1) * 2 * BN) : 0u);
            mma_bf16(d_set, a_hi + (uint64_t)(kk * 2), b_hi + (uint64_t)(kk * 2), idesc2, ks < NACC / 2 ? 0u : 1u);
          }
        }
        mma_commit(&a_empty[0]);
        // ---- phase 2: the lo plane against b_hi -> corr
        mbar_wait(&a_full[1], ti & 1u);
        tc_fence_after();
        ks = 0;
        for (int t = 0; t < 9; ++t) {
          const uint64_t a_lo = smem_desc_sw128(a_base + p.region_bytes + (uint32_t)(delta + p.tap_shift[t]) * 128u);
          const uint64_t b_hi = smem_desc_sw128(b_base + (uint32_t)t * B_STAGE);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk, ++ks) {
            const uint32_t d_set = d_buf + (NACC == 4 ? (uint32_t)((ks & 1) * 2 * BN) : 0u);
            mma_bf16(d_set + BN, a_lo + (uint64_t)(kk * 2), b_hi + (uint64_t)(kk * 2), idesc, 1u);
          }
        }
        mma_commit(&a_empty[1]);
        mma_commit(&acc_full[tb]);
      }
    }
    __syncwarp();
  } else {
    epilogue_warps<BN, NACC, NBUF, RED>(p, tmem_base, a_scale, acc_full, acc_empty, s_sum, s_sq, s_bias, blockIdx.x, gridDim.x, 1, 0u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, NBUF * TM_BUF);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}


static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e == nullptr ? dflt : atoi(e);
}

struct Plan {
  int RB, box_pos, nbox, b_stages;
  uint32_t region_bytes;
  size_t smem;            // streaming kernel (0: it does not fit)
  size_t res_smem;        // weight-resident kernel (0: not applicable / does not fit)
};
// Geometry of the activation regions plus the shared-memory footprints of the two kernels. cpt / n_ntiles / ntaps decide whether the
// weight-resident variant applies (one 64-channel chunk, one output tile, nine taps). False when neither kernel fits.
static bool make_plan(int H, int W, int BN, int Npad, Plan& pl, int halo = -1, int cpt = 0, int n_ntiles = 0, int ntaps = 9, bool explicit_taps = false) {
  const int Wp = W + 1;
  if (halo < 0) halo = Wp + 1;
  // RB: rows per TMA box; must divide H + 1 so that a box never straddles two images. Largest divisor with <= 64 positions.
  int RB = 1;
  for (int d = 1; d <= H + 1; ++d)
    if ((H + 1) % d == 0 && d * Wp <= 64) RB = d;
  if (RB * Wp > 256 || Wp > 256) return false;
  pl.RB = RB;
  pl.box_pos = RB * Wp;
  // worst case: the region starts up to box_pos - 1 positions before q_lo
  pl.nbox = (halo + pl.box_pos - 1 + BM + halo + pl.box_pos - 1) / pl.box_pos;
  if (pl.nbox > MAX_BOX) return false;
  pl.region_bytes = (uint32_t)pl.nbox * pl.box_pos * 128u;
  pl.region_bytes = (pl.region_bytes + 1023u) & ~1023u;
  const size_t budget = 227 * 1024;
  const size_t tail = sizeof(float) * 7 * (size_t)Npad + 1024;
  // streaming kernel: two activation buffers (hi | lo each) + >= 2 weight stages; needs 64-column output tiles
  pl.smem = 0; pl.b_stages = 0;
  const size_t fixed = 2 * (size_t)2 * pl.region_bytes + sizeof(uint64_t) * (8 + 2 * MAX_BST) + 16 + tail;
  if (BN >= 64 && fixed + 2 * (size_t)2 * BN * 128 <= budget) {
    int st = (int)((budget - fixed) / ((size_t)2 * BN * 128));
    if (st > MAX_BST) st = MAX_BST;
    pl.b_stages = st;
    pl.smem = fixed + (size_t)st * 2 * BN * 128;
  }
  // weight-resident kernel: one activation region + all nine taps' weights
  pl.res_smem = 0;
  if (env_int("PC_HALO_RESIDENT", 1) != 0 && ntaps == 9 && !explicit_taps && BN <= 64 && cpt == 1 && n_ntiles == 1) {
    const size_t rs = 2 * (size_t)pl.region_bytes + 9 * (size_t)2 * BN * 128 + sizeof(uint64_t) * 10 + 16 + tail;
    if (rs <= budget) pl.res_smem = rs;
  }
  return pl.smem != 0 || pl.res_smem != 0;
}

// (32: layers with 32 produced channels, weight-resident kernel only -- their operand is packed with 32 rows per part, csrc/conv_tc.cu npad_of)
static inline int pick_bn(int Nn) { return Nn <= 32 ? 32 : (Nn <= 64 ? 64 : 128); }

// H, W: the image the POSITION space is built on (the output image of a forward pass, the dy image of a data gradient).
// ntaps = 9: 3x3 stride-1 "same" convolution. ntaps = 1: 1x1 convolution (the residual shortcuts); a_stride = 2 samples every second
// pixel of the (Ha x Wa) activation tensor (stride-2 forward), o_stride = 2 scatters the rows to every second pixel of the
// (Hout x Wout) output (stride-2 data gradient, accumulate only: the other pixels receive nothing from this convolution).
// taps (optional): an explicit tap list -- position shift and weight tap id per issued tap -- with the output parity offsets of a
// stride-2 3x3 data gradient: dx[2i + p, 2j + q] = sum over the taps (r, s) with r = p + 1 (mod 2), s = q + 1 (mod 2) of
// dy[i + a, j + b] w[r][s]^T, a = 1 for r = 0 (else 0), b = 1 for s = 0 (else 0): a stride-1 "convolution" of dy per output class.
struct TapSpec { int n; int shift[9]; int id[9]; int off_h, off_w; };
static int run(const void* planes, const void* wp, const float* bias, const float* a_amax, float* out, double* stats, int B, int H, int W, int Ca,
               int Nn, int dgrad, int accumulate, pc_stream_t stream, int ntaps = 9, int a_stride = 1, int Ha = 0, int Wa = 0, int Hout = 0,
               int Wout = 0, int o_stride = 1, const TapSpec* taps = nullptr, const PcBnBwdReduce* red = nullptr) {
  if (taps != nullptr) ntaps = taps->n;
  const int BN = pick_bn(Nn);
  const int Npad = ceil_div(Nn, BN) * BN;
  if (Ha == 0) { Ha = H; Wa = W; }
  if (Hout == 0) { Hout = H; Wout = W; }
  const int halo = (ntaps == 1 && taps == nullptr) ? 0 : W + 2;
  const int cpt_ = ceil_div(Ca, 64);          // a 32-channel tensor is one chunk whose channels 32..63 the TMA unit zero-fills (box > tensor)
  Plan pl;
  PC_REQUIRE(make_plan(H, W, BN, Npad, pl, halo, cpt_, Npad / BN, ntaps, taps != nullptr), PC_EUNSUPPORTED, "conv_halo: image %dx%d does not fit the halo plan", H, W);
  PC_REQUIRE(a_stride * (W + 1) <= 256 && a_stride * pl.RB <= 256, PC_EUNSUPPORTED, "conv_halo: strided box exceeds the TMA box limit");
  EncodeTiledFn enc = encode_tiled();
  PC_REQUIRE(enc != nullptr, PC_ECUDA, "conv_halo: cuTensorMapEncodeTiled is not available from this driver");
  Params p{};
  p.Bp = static_cast<const unsigned char*>(wp); p.bias = bias; p.a_amax = a_amax; p.C = out; p.stats = stats;
  if (red != nullptr) p.red = *red;
  p.B = B; p.H = H; p.W = W; p.Ca = Ca; p.Nn = Nn; p.Npad = Npad;
  p.Wp = W + 1; p.Pimg = (H + 1) * (W + 1); p.RB = pl.RB; p.box_pos = pl.box_pos; p.nbox = pl.nbox;
  p.Q = (long long)B * p.Pimg;
  p.n_mtiles = ceil_div(p.Q, BM);
  p.n_ntiles = Npad / BN;
  p.cpt = cpt_;
  p.accumulate = accumulate;
  int CL = env_int("PC_HALO_CLUSTER", 1);
  if (CL != 1 && CL != 2 && CL != 4) CL = 1;
  p.cluster = CL;
  p.n_items = ceil_div(p.n_mtiles, CL) * p.n_ntiles;
  // forward: tap (r, s) reads position q + (r-1) (W+1) + (s-1); data gradient: dx[q] = sum dy[q - (r-1)(W+1) - (s-1)] w[r][s]
  for (int r = 0; r < 3; ++r)
    for (int s = 0; s < 3; ++s) p.tap_shift[r * 3 + s] = ntaps == 1 ? 0 : (dgrad ? -1 : 1) * ((r - 1) * p.Wp + (s - 1));
  for (int t = 0; t < 9; ++t) p.tap_id[t] = ntaps == 1 ? 0 : t;
  p.o_off_h = p.o_off_w = 0;
  if (taps != nullptr) {
    for (int t = 0; t < taps->n; ++t) { p.tap_shift[t] = taps->shift[t]; p.tap_id[t] = taps->id[t]; }
    p.o_off_h = taps->off_h; p.o_off_w = taps->off_w;
  }
  p.ntaps = ntaps; p.halo = halo; p.a_stride = a_stride; p.Hout = Hout; p.Wout = Wout; p.o_stride = o_stride;
  p.b_stages = pl.b_stages;
  p.region_bytes = pl.region_bytes;
  p.a_tx_bytes = 2u * (uint32_t)pl.nbox * (uint32_t)pl.box_pos * 128u;
  p.d_pimg = FastDiv::make((uint32_t)p.Pimg); p.d_wp = FastDiv::make((uint32_t)p.Wp); p.d_nt = FastDiv::make((uint32_t)p.n_ntiles);
  p.d_box = FastDiv::make((uint32_t)p.box_pos); p.d_hp1 = FastDiv::make((uint32_t)(H + 1));

  CUtensorMap amap;
  const cuuint64_t dims[5] = {(cuuint64_t)Ca, (cuuint64_t)Wa, (cuuint64_t)Ha, (cuuint64_t)B, 2};
  const cuuint64_t strides[4] = {(cuuint64_t)Ca * 2, (cuuint64_t)Wa * Ca * 2, (cuuint64_t)Ha * Wa * Ca * 2, (cuuint64_t)B * Ha * Wa * Ca * 2};
  // with an element stride s the box spans s * n tensor elements and delivers n of them (every s-th)
  const cuuint32_t box[5] = {64, (cuuint32_t)(a_stride * p.Wp), (cuuint32_t)(a_stride * p.RB), 1, 1};
  const cuuint32_t es[5] = {1, (cuuint32_t)a_stride, (cuuint32_t)a_stride, 1, 1};
  const CUresult cr = enc(&amap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(planes), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PC_REQUIRE(cr == CUDA_SUCCESS, PC_ECUDA, "conv_halo: cuTensorMapEncodeTiled failed (CUresult %d)", (int)cr);

  // weight-resident variant: one 64-channel chunk, one output tile of <= 64 channels, and all nine taps' weights + one activation
  // region fit the 227 KB of shared memory
  const size_t res_smem = pl.res_smem;
  const bool resident = res_smem != 0;
  PC_REQUIRE(resident || pl.smem != 0, PC_EUNSUPPORTED, "conv_halo: no kernel variant fits this layer");
  PC_REQUIRE(BN != 32 || (resident && red == nullptr), PC_EUNSUPPORTED, "conv_halo: 32 produced channels run the weight-resident kernel only");
  if (resident) {
    p.cluster = CL = 1;
    p.n_items = p.n_mtiles;
    p.d_nt = FastDiv::make(1u);
    const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
#define PC_HALO_RES(BN_, RED_)                                                                                                        \
  do {                                                                                                                                \
    static size_t conf = 0;                                                                                                           \
    if (res_smem > conf) {                                                                                                            \
      PC_CUDA(cudaFuncSetAttribute((conv_halo_res_kernel<BN_, 2, 2, RED_>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)res_smem)); \
      conf = res_smem;                                                                                                                \
    }                                                                                                                                 \
    conv_halo_res_kernel<BN_, 2, 2, RED_><<<grid, THREADS, res_smem, stream>>>(amap, p);                                              \
  } while (0)
    if (BN == 32) PC_HALO_RES(32, false);
    else if (red != nullptr) PC_HALO_RES(64, true);
    else PC_HALO_RES(64, false);
#undef PC_HALO_RES
    PC_LAUNCH_CHECK("conv_halo_res_kernel");
    return PC_OK;
  }
  const int max_clusters = kNumSMs / CL;
  const int n_clusters = p.n_items < max_clusters ? p.n_items : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_clusters * CL));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
#define PC_HALO_LAUNCH1(BN_, NACC_, NBUF_, RED_)                                                                                       \
  do {                                                                                                                          \
    static size_t conf = 0;                                                                                                     \
    if (pl.smem > conf) {                                                                                                       \
      PC_CUDA(cudaFuncSetAttribute((conv_halo_kernel<BN_, NACC_, NBUF_, RED_>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem)); \
      conf = pl.smem;                                                                                                           \
    }                                                                                                                           \
    PC_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<BN_, NACC_, NBUF_, RED_>, amap, p));                                             \
  } while (0)
#define PC_HALO_LAUNCH(BN_, NACC_, NBUF_)                       \
  do {                                                          \
    if (red != nullptr) PC_HALO_LAUNCH1(BN_, NACC_, NBUF_, true); \
    else PC_HALO_LAUNCH1(BN_, NACC_, NBUF_, false);               \
  } while (0)
  if (BN == 64 && p.cpt <= 2) PC_HALO_LAUNCH(64, 2, 2);
  else if (BN == 64) PC_HALO_LAUNCH(64, 4, 2);
  else if (p.cpt >= 4) PC_HALO_LAUNCH(128, 4, 1);
  else PC_HALO_LAUNCH(128, 2, 2);
#undef PC_HALO_LAUNCH
#undef PC_HALO_LAUNCH1
  PC_LAUNCH_CHECK("conv_halo_kernel");
  return PC_OK;
}

}  // namespace halo
}  // namespace pc

using namespace pc;

// 1 when the halo engine covers this convolution: stride 1, 3x3, pad 1, gathered and produced channels multiples of 64.
extern "C" int pc_conv_halo_supported(const PcConvGeom* g, int dgrad) {
  if (g == nullptr) return 0;
  if (!pc::halo::env_int("PC_CONV_HALO", 1)) return 0;
  const int ca = dgrad ? g->Cout : g->Cin, nn = dgrad ? g->Cin : g->Cout;
  // channel counts: multiples of 64 -- or exactly 32 for the stride-1 3x3 layers (cnn_small): 32 gathered channels are one 64-channel
  // chunk whose upper half the TMA boxes zero-fill (box wider than the tensor), 32 produced channels run the weight-resident kernel with
  // a 32-column tile
  const bool strict = ca % 64 == 0 && nn % 64 == 0 && nn >= 64;
  const bool relaxed = (ca % 64 == 0 || ca == 32) && (nn % 64 == 0 || nn == 32);
  if (!relaxed) return 0;
  if (!strict && !(g->R == 3 && g->S == 3 && g->stride == 1 && g->pad == 1 && pc::halo::env_int("PC_HALO_C32", 1))) return 0;
  if (g->R == 1 && g->S == 1 && g->pad == 0 && (g->stride == 1 || g->stride == 2)) {
    // 1x1 (shortcut) convolutions: position space = the (Ho x Wo) image; PC_HALO_1X1=0 keeps them on the per-tap-gather kernel
    if (!pc::halo::env_int("PC_HALO_1X1", 1)) return 0;
    if ((long long)g->B * (g->Ho + 1) * (g->Wo + 1) + 4096 >= (1LL << 31) || (long long)g->B * g->H * g->W * (ca > nn ? ca : nn) >= (1LL << 31)) return 0;
    pc::halo::Plan pl1;
    const int bn1 = pc::halo::pick_bn(nn);
    if (g->stride * (g->Wo + 1) > 256) return 0;
    return pc::halo::make_plan(g->Ho, g->Wo, bn1, ceil_div(nn, bn1) * bn1, pl1, 0) && g->stride * pl1.RB <= 256 ? 1 : 0;
  }
  if (dgrad && g->R == 3 && g->S == 3 && g->stride == 2 && g->pad == 1) {
    // stride-2 3x3 data gradient as four stride-1 problems over dy, one per output parity class (see TapSpec). Correct
    // (tests/test_gpu_halo.py::test_halo_stride2_dgrad) but measured SLOWER than the per-tap-gather kernel with its parity-class tap
    // skipping at 256 views -- pc_conv_dgrad 0.716 ms -> 0.800 (128- and 256-channel dy) -> 0.882 (all three layers): every class
    // re-loads the dy tile + halo and runs 1 - 4 taps per tile. OFF by default (PC_HALO_S2=1 enables, PC_HALO_ALL=1 adds the 512-channel
    // layer); the version worth building keeps the four classes' accumulators in TMEM and loads dy once.
    if (!pc::halo::env_int("PC_HALO_S2", 0)) return 0;
    if (ca > 256 && !pc::halo::env_int("PC_HALO_ALL", 0)) return 0;
    if ((long long)g->B * (g->Ho + 1) * (g->Wo + 1) + 4096 >= (1LL << 31) || (long long)g->B * g->H * g->W * (ca > nn ? ca : nn) >= (1LL << 31)) return 0;
    pc::halo::Plan pl2;
    const int bn2 = pc::halo::pick_bn(nn);
    return pc::halo::make_plan(g->Ho, g->Wo, bn2, ceil_div(nn, bn2) * bn2, pl2) ? 1 : 0;
  }
  if (g->R != 3 || g->S != 3 || g->stride != 1 || g->pad != 1 || g->Ho != g->H || g->Wo != g->W) return 0;
  // Measured per layer at 256 views (profiles/r2_halo_bench.md): 64ch 131 -> 58 us, 128ch 98 -> 77, 512ch 119 -> 107, but 256ch
  // (5x13 images) 76 -> 88: with 4 chunks x 288 KB of streamed weights per tile and only 168 position tiles the per-tap-gather
  // kernel's two resident... one-tile-per-CTA grid balances better. PC_HALO_ALL=1 forces the halo engine everywhere.
  if (ca == 256 && !pc::halo::env_int("PC_HALO_ALL", 0)) return 0;
  const long long Q = (long long)g->B * (g->H + 1) * (g->W + 1);
  if (Q + 4096 >= (1LL << 31) || (long long)g->B * g->H * g->W * (ca > nn ? ca : nn) >= (1LL << 31)) return 0;
  pc::halo::Plan pl;
  const int bn = pc::halo::pick_bn(nn);
  const int npad = ceil_div(nn, bn) * bn;
  if (!pc::halo::make_plan(g->H, g->W, bn, npad, pl, -1, ceil_div(ca, 64), npad / bn)) return 0;
  return (bn != 32 || pl.res_smem != 0) ? 1 : 0;
}

extern "C" int pc_conv_fwd_halo(const void* x_planes, const void* wp, const float* bias, const PcConvGeom* g, float* y, double* stats,
                                pc_stream_t stream) {
  PC_REQUIRE(x_planes && wp && g && y, PC_EINVAL, "pc_conv_fwd_halo: null pointer");
  if (!pc_conv_halo_supported(g, 0)) return PC_EUNSUPPORTED;
  if (g->R == 1)
    return pc::halo::run(x_planes, wp, bias, nullptr, y, stats, g->B, g->Ho, g->Wo, g->Cin, g->Cout, 0, 0, stream, 1, g->stride, g->H, g->W, g->Ho,
                         g->Wo, 1);
  return pc::halo::run(x_planes, wp, bias, nullptr, y, stats, g->B, g->H, g->W, g->Cin, g->Cout, 0, 0, stream);
}

extern "C" int pc_conv_dgrad_halo(const void* dy_planes, const void* wp, const PcConvGeom* g, float* dx, int accumulate, const float* dy_amax,
                                  pc_stream_t stream) {
  PC_REQUIRE(dy_planes && wp && g && dx && dy_amax, PC_EINVAL, "pc_conv_dgrad_halo: null pointer");
  if (!pc_conv_halo_supported(g, 1)) return PC_EUNSUPPORTED;
  if (g->R == 3 && g->stride == 2) {
    const int Wp = g->Wo + 1;
    for (int pc_ = 0; pc_ < 2; ++pc_)
      for (int qc = 0; qc < 2; ++qc) {
        pc::halo::TapSpec ts{};
        ts.off_h = pc_; ts.off_w = qc;
        for (int r = 0; r < 3; ++r) {
          if (((r + 1) & 1) != pc_) continue;             // h = 2 ho + r - 1: the parity of h is that of r + 1
          for (int sx = 0; sx < 3; ++sx) {
            if (((sx + 1) & 1) != qc) continue;
            // ho = (h + 1 - r) / 2 = i + (p + 1 - r) / 2 with h = 2 i + p: r = 0 -> i + 1 (p = 1), r = 1 -> i (p = 0), r = 2 -> i (p = 1)
            const int a = r == 0 ? 1 : 0, b = sx == 0 ? 1 : 0;
            ts.shift[ts.n] = a * Wp + b;
            ts.id[ts.n] = r * 3 + sx;
            ++ts.n;
          }
        }
        const int rc = pc::halo::run(dy_planes, wp, nullptr, dy_amax, dx, nullptr, g->B, g->Ho, g->Wo, g->Cout, g->Cin, 1, accumulate, stream, ts.n, 1,
                                     g->Ho, g->Wo, g->H, g->W, 2, &ts);
        if (rc != PC_OK) return rc;
      }
    return PC_OK;
  }
  if (g->R == 1) {
    // dx[b, s ho, s wo, :] (+)= dy[b, ho, wo, :] W^T; with stride 2 the remaining pixels get no contribution, so only accumulation is defined here
    if (g->stride != 1 && !accumulate) return PC_EUNSUPPORTED;
    return pc::halo::run(dy_planes, wp, nullptr, dy_amax, dx, nullptr, g->B, g->Ho, g->Wo, g->Cout, g->Cin, 1, accumulate, stream, 1, 1, g->Ho, g->Wo,
                         g->H, g->W, g->stride);
  }
  return pc::halo::run(dy_planes, wp, nullptr, dy_amax, dx, nullptr, g->B, g->H, g->W, g->Cout, g->Cin, 1, accumulate, stream);
}

// Stride-1 3x3 data gradient with the REDUCE pass of the BatchNorm backward that consumes dx fused into the epilogue (Params::red):
// dx = dgrad(dy) is written as usual; red->sums / red->maxes (zeroed by the caller) receive what pc_bn_act_bwd_reduce(dx, red->y, ...,
// pool = 0) would have produced, without re-reading dx and y from HBM.
extern "C" int pc_conv_dgrad_halo_bnred(const void* dy_planes, const void* wp, const PcConvGeom* g, float* dx, const float* dy_amax,
                                        const PcBnBwdReduce* red, pc_stream_t stream) {
  PC_REQUIRE(dy_planes && wp && g && dx && dy_amax && red, PC_EINVAL, "pc_conv_dgrad_halo_bnred: null pointer");
  PC_REQUIRE(red->y && red->scale && red->shift && red->mean && red->invstd && red->sums, PC_EINVAL, "pc_conv_dgrad_halo_bnred: incomplete reduce record");
  if (!pc_conv_halo_supported(g, 1) || g->R != 3 || g->stride != 1) return PC_EUNSUPPORTED;
  return pc::halo::run(dy_planes, wp, nullptr, dy_amax, dx, nullptr, g->B, g->H, g->W, g->Cout, g->Cin, 1, 0, stream, 9, 1, 0, 0, 0, 0, 1, nullptr, red);
}
