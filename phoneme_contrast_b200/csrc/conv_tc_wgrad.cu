// tcgen05 tensor-core weight gradient (TF32x3):
//
//   dW[(tap,c)][n] = sum over output pixels m of  xform(x)[pix(m,tap)][c] * dy[m][n]
//
// GEMM view: D[128 (tap,c) rows x BN out-channels] += A'[128 x 8 pixels] * B'[BN x 8 pixels]^T per k-step, where the
// REDUCTION dimension is the pixel index. In NHWC both operands have the pixel index as their slow dimension and the
// channel index contiguous, i.e. they are "MN-major" in UMMA terms. MN-major tf32 operands have exactly one legal
// shared-memory layout, SWIZZLE_128B_BASE32B: atoms of 4 pixel rows x 128 B (32 channels) whose 32-byte chunks are
// XOR-ed with the row index. Tiles are stored as [channel group of 32][pixel group of 4][4 rows][128 B] and described
// with LBO = 4096 B (next channel group), SBO = 512 B (next pixel group), a_major = b_major = MN.
// A 128-row M tile is a run of 128 consecutive (tap, channel) indices, so for Cin = 64 it covers two taps and the
// producers gather two different input pixels per output pixel.
//
// FP16X2 variant (F16 = true): operands split into fp16 hi / lo*2^11 (tc_common.cuh: split_f16x2), kind::f16 MMAs with
// K = 16 pixels per instruction. 16-bit MN-major operands use the ordinary SWIZZLE_128B layout: atoms of 8 pixel rows x
// 128 B (64 channels), 16-byte chunk index XOR (row & 7); tiles are [channel group of 64][32 pixels][128 B] with
// LBO = 4096 B (next channel group) and SBO = 1024 B (next 8 pixels). A stage still holds 32 pixels (now 2 k-steps) in
// half the bytes, dy is pre-scaled by a power of two from its max magnitude (dy_amax) and the scale is undone in the
// epilogue.
//
// grid = (ceil(K/128), ceil(Cout/BN), splits over pixels); each CTA writes its partial tile to
// partial[split][K+1][Cout] and conv_wgrad_reduce_kernel (conv_simt.cu) sums the splits in a fixed order into OIHW.
#include "common.cuh"
#include "tc_common.cuh"

namespace pc {
namespace tcwg {

using namespace pc::tc;

constexpr int NPROD = 128;             // threads per producer group; NGROUPS groups: 3 with one CTA per SM (BN = 128),
                                       // 2 with two co-resident CTAs per SM (BN = 64), as in conv_tc.cu
constexpr int PIX = 32;                 // pixels per stage (4 k-steps of 8)
constexpr int MAX_STAGES = 6;
constexpr uint32_t SMEM_BUDGET = 200 * 1024;
constexpr int NACC = 4;

struct XformDev {
  const float* scale;
  const float* shift;
  const float* drop;
  int relu;
};

struct Params {
  const float* x;
  const float* dy;
  float* partial;
  XformDev xf;
  PcConvGeom g;
  int K, M, rows_per_split, stages;
  const float* dy_amax;   // FP16X2: device scalar max|dy| (null: scale 1)
  size_t plane_bytes;     // PRESPLIT: byte distance between the hi and lo planes of x
  FastDiv div_wo, div_ho; // pixel index -> (b, ho, wo) without integer divides (twice per pixel row and stage otherwise)
  int dy_presplit;        // dy is a pair of fp16 planes already scaled by f16_operand_scale(*dy_amax) (F16 only)
  size_t dy_plane_bytes;
};

__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t smem_addr_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr_bytes >> 4) & 0x3FFF);
  d |= (uint64_t)(4096 >> 4) << 16;      // LBO: next 32-channel group
  d |= (uint64_t)(512 >> 4) << 32;       // SBO: next 4-pixel group (one swizzle atom = 4 pixel rows x 128 B)
  d |= (uint64_t)1 << 46;                // version
  d |= (uint64_t)1 << 61;                // SWIZZLE_128B_BASE32B: the only layout for MN-major tf32 operands
  return d;
}

// 16-bit MN-major operand, SWIZZLE_128B: 8-row x 128-byte atoms
__device__ __forceinline__ uint64_t smem_desc_mn_sw128_h(uint32_t smem_addr_bytes, uint32_t group_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr_bytes >> 4) & 0x3FFF);
  d |= (uint64_t)(group_bytes >> 4) << 16;   // LBO: next 64-channel group (pixels per stage x 128 B)
  d |= (uint64_t)(1024 >> 4) << 32;      // SBO: next 8-pixel group
  d |= (uint64_t)1 << 46;                // version
  d |= (uint64_t)2 << 61;                // SWIZZLE_128B
  return d;
}

// channels per 128-byte row: 32 (tf32) or 64 (fp16); a 128-row M tile is 128 / CPG channel groups
// STEM stages hold 64 pixels instead of 32: that kernel streams dy once from HBM with nothing else to do, so the bytes
// each producer group has in flight set its speed
template <int BN, bool F16, bool STEM = false>
__host__ __device__ constexpr uint32_t stage_bytes() {
  return 2u * ((F16 ? 2u : 4u) * (STEM ? 8192u : 4096u) + (BN / (F16 ? 64 : 32)) * (STEM ? 8192u : 4096u));
}

// STEM (with F16): single input channel. The M tile's first 64 rows are the R*S taps (zero padded), the second channel
// group stays zero; thread (pixel row, j) gathers taps 8j..8j+7 of its pixel's window with scalar loads.
// PRESPLIT (with F16): x is the pair of fp16 planes written by pc_bn_act_split; the A' gather copies 16-byte chunks.
template <int BN, int NGROUPS, int MINB, bool F16, bool STEM = false, bool PRESPLIT = false>
__global__ void __launch_bounds__(32 * (4 * NGROUPS + 1), MINB) wgrad_tc_kernel(const Params p) {
  static_assert(!STEM || F16, "the stem gather is built for the FP16X2 tiles only");
  static_assert(!PRESPLIT || (F16 && !STEM), "pre-split planes feed the FP16X2 tiles only");
  constexpr int PROD_WARPS = 4 * NGROUPS;
  constexpr int CPG = F16 ? 64 : 32;          // channels per 128-byte row (one channel group)
  constexpr int GA = 128 / CPG, GB = BN / CPG; // channel groups of the A' (M) and B' (N) tiles
  constexpr int NV = F16 ? 2 : 1;             // float4 loads per thread, row and group (8 or 4 channels -> one 16-byte chunk)
  constexpr int PIXS = STEM ? 64 : PIX;       // pixels per stage
  constexpr int HR = PIXS / 16;               // pixel rows per thread and stage (rows pr + 16*h)
  constexpr uint32_t GRP = PIXS * 128;        // bytes of one channel group of a stage
  constexpr uint32_t A_PART = GA * GRP, B_PART = GB * GRP;
  constexpr uint32_t STAGE = stage_bytes<BN, F16, STEM>();
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int S = p.stages;
  unsigned char* tiles = smem;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * STAGE);
  uint64_t* empty = full + MAX_STAGES;
  uint64_t* acc_full = empty + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const PcConvGeom g = p.g;
  const int kt = blockIdx.x, n0 = blockIdx.y * BN, split = blockIdx.z;
  const int m_begin = split * p.rows_per_split;
  const int m_end = min(p.M, m_begin + p.rows_per_split);
  const int n_stages = (m_end - m_begin + PIXS - 1) / PIXS;

  if (warp == PROD_WARPS) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(&full[s], NPROD);
        mbar_init(&empty[s], 1);
      }
      mbar_init(acc_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, NACC * BN);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();   // see common.cuh: only after the TMEM allocation
  pdl_wait();      // nothing above touched global memory

  if (warp < PROD_WARPS) {
    const int group = warp >> 2;
    const int gt = tid & (NPROD - 1);
    const int j = gt & 7, pr = gt >> 3;       // 16-byte chunk, pixel row pr (+16)
    // the channel groups of this M tile: (tap, first channel of my 16-byte chunk)
    int q_tr[GA], q_ts[GA], q_c0[GA];
    bool q_ok[GA];
#pragma unroll
    for (int q = 0; q < GA; ++q) {
      const int kidx0 = 128 * kt + CPG * q;
      q_ok[q] = kidx0 < p.K;
      const int tap = q_ok[q] ? kidx0 / g.Cin : 0;
      q_c0[q] = (q_ok[q] ? kidx0 - tap * g.Cin : 0) + 4 * NV * j;
      q_tr[q] = tap / g.S;
      q_ts[q] = tap - q_tr[q] * g.S;
    }
    const bool has_aff = p.xf.scale != nullptr, has_relu = p.xf.relu != 0, has_drop = p.xf.drop != nullptr;
    float4 q_sc[GA][NV], q_sh[GA][NV];
#pragma unroll
    for (int q = 0; q < GA; ++q)
#pragma unroll
      for (int vv = 0; vv < NV; ++vv) {
        q_sc[q][vv] = make_float4(1.f, 1.f, 1.f, 1.f);
        q_sh[q][vv] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_aff && q_ok[q]) {
          q_sc[q][vv] = *reinterpret_cast<const float4*>(p.xf.scale + q_c0[q] + 4 * vv);
          q_sh[q][vv] = *reinterpret_cast<const float4*>(p.xf.shift + q_c0[q] + 4 * vv);
        }
      }
    const bool use_scale = F16 && p.dy_amax != nullptr;
    const float b_scale = use_scale ? f16_operand_scale(p.dy_amax[0]) : 1.f;
    int stem_dh[STEM ? 8 : 1], stem_dw[STEM ? 8 : 1];
    float bsum[STEM ? 8 : 1] = {};     // STEM: bias gradient (column sums of dy) of my 8 channels, folded into this pass over dy
    if (STEM) {
      // rows 64..127 of the M tile (second channel group of a_hi / a_lo) are never written again: zero them once
      constexpr int V = GRP / 16;     // uint4 per group
      for (int i = tid; i < S * 2 * V; i += 32 * PROD_WARPS) {
        const int sidx = i / (2 * V), rem = i - sidx * 2 * V;
        unsigned char* part = tiles + (size_t)sidx * STAGE + (rem >= V ? A_PART : 0) + GRP;
        *reinterpret_cast<uint4*>(part + (rem % V) * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, %0;" ::"n"(32 * PROD_WARPS) : "memory");
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int t = 8 * j + q, tr = t / g.S;
        stem_dh[q] = t < p.K ? tr - g.pad : -100000;      // out-of-range sentinel -> bounds test fails
        stem_dw[q] = t - tr * g.S - g.pad;
      }
    }
    for (int st = group; st < n_stages; st += NGROUPS) {
      const int s = st % S;
      const uint32_t ph = (uint32_t)(st / S) & 1u;
      float4 av[HR][STEM ? 1 : GA][NV], bv[HR][GB][NV];
#pragma unroll
      for (int h = 0; h < HR; ++h) {
        const int pl = pr + 16 * h;
        const int m = m_begin + st * PIXS + pl;
        const bool mv = m < m_end;
        int b = 0, ho = 0, wo = 0;
        if (mv) {
          uint32_t t_, wo_, b_, ho_;
          p.div_wo.divmod((uint32_t)m, t_, wo_);
          p.div_ho.divmod(t_, b_, ho_);
          wo = (int)wo_; ho = (int)ho_; b = (int)b_;
        }
#pragma unroll
        for (int q = 0; q < GB; ++q)
#pragma unroll
          for (int vv = 0; vv < NV; ++vv) {
            const int n = n0 + CPG * q + 4 * NV * j + 4 * vv;
            if (F16 && p.dy_presplit) {      // vv = 0: hi bits, vv = 1: lo bits of the 8 channels n0 + CPG*q + 8*j ..
              uint4 bits = make_uint4(0u, 0u, 0u, 0u);
              const int n8 = n0 + CPG * q + 8 * j;
              if (mv && n8 < g.Cout)
                bits = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.dy) + (vv ? p.dy_plane_bytes : 0) +
                                                       ((size_t)m * g.Cout + n8) * 2);
              bv[h][q][vv] = make_float4(__uint_as_float(bits.x), __uint_as_float(bits.y), __uint_as_float(bits.z), __uint_as_float(bits.w));
              continue;
            }
            bv[h][q][vv] = (mv && n < g.Cout) ? *reinterpret_cast<const float4*>(p.dy + (size_t)m * g.Cout + n) : make_float4(0.f, 0.f, 0.f, 0.f);
            if (STEM) {
              bsum[(4 * vv) % (STEM ? 8 : 1)] += bv[h][q][vv].x; bsum[(4 * vv + 1) % (STEM ? 8 : 1)] += bv[h][q][vv].y;
              bsum[(4 * vv + 2) % (STEM ? 8 : 1)] += bv[h][q][vv].z; bsum[(4 * vv + 3) % (STEM ? 8 : 1)] += bv[h][q][vv].w;
            }
          }
        if (STEM) {
          float tv[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int hi = ho + stem_dh[q], wi = wo + stem_dw[q];
            const bool ok = mv && (unsigned)hi < (unsigned)g.H && (unsigned)wi < (unsigned)g.W;
            tv[q] = ok ? p.x[((size_t)b * g.H + hi) * g.W + wi] : 0.f;
          }
          av[h][0][0] = make_float4(tv[0], tv[1], tv[2], tv[3]);
          av[h][0][NV - 1] = make_float4(tv[4], tv[5], tv[6], tv[7]);
        }
#pragma unroll
        for (int q = 0; q < (STEM ? 0 : GA); ++q) {
          const int hi = ho * g.stride - g.pad + q_tr[q], wi = wo * g.stride - g.pad + q_ts[q];
          const bool ok = mv && q_ok[q] && (unsigned)hi < (unsigned)g.H && (unsigned)wi < (unsigned)g.W;
          if (PRESPLIT) {       // hi bits -> av[..][0], lo bits -> av[..][NV-1]
            uint4 hb = make_uint4(0u, 0u, 0u, 0u), lb = hb;
            if (ok) {
              const size_t e = ((((size_t)b * g.H + hi) * g.W + wi) * g.Cin + q_c0[q]) * 2;
              hb = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.x) + e);
              lb = *reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(p.x) + p.plane_bytes + e);
            }
            av[h][q][0] = make_float4(__uint_as_float(hb.x), __uint_as_float(hb.y), __uint_as_float(hb.z), __uint_as_float(hb.w));
            av[h][q][NV - 1] = make_float4(__uint_as_float(lb.x), __uint_as_float(lb.y), __uint_as_float(lb.z), __uint_as_float(lb.w));
            continue;
          }
#pragma unroll
          for (int vv = 0; vv < NV; ++vv) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) {
              v = *reinterpret_cast<const float4*>(p.x + (((size_t)b * g.H + hi) * g.W + wi) * g.Cin + q_c0[q] + 4 * vv);
              if (has_aff) v = make_float4(fmaf(v.x, q_sc[q][vv].x, q_sh[q][vv].x), fmaf(v.y, q_sc[q][vv].y, q_sh[q][vv].y), fmaf(v.z, q_sc[q][vv].z, q_sh[q][vv].z), fmaf(v.w, q_sc[q][vv].w, q_sh[q][vv].w));
              if (has_relu) v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
              if (has_drop) {
                const float4 d = *reinterpret_cast<const float4*>(p.xf.drop + (size_t)b * g.Cin + q_c0[q] + 4 * vv);
                v = make_float4(v.x * d.x, v.y * d.y, v.z * d.z, v.w * d.w);
              }
            }
            av[h][q][vv] = v;
          }
        }
      }
      mbar_wait(&empty[s], ph ^ 1u);
      unsigned char* a_hi = tiles + (size_t)s * STAGE;
      unsigned char* a_lo = a_hi + A_PART;
      unsigned char* b_hi = a_hi + 2 * A_PART;
      unsigned char* b_lo = b_hi + B_PART;
#pragma unroll
      for (int h = 0; h < HR; ++h) {
        const int pl = pr + 16 * h;
        if (F16) {
          // atom = 8 pixel rows x 128 B; 16-byte chunk index XOR (row & 7)
          const uint32_t off = (uint32_t)((pl >> 3) * 1024 + (pl & 7) * 128 + ((j ^ (pl & 7)) << 4));
#pragma unroll
          for (int q = 0; q < (STEM ? 1 : GA); ++q) {
            uint4 hh, ll;
            if (PRESPLIT) {
              hh = make_uint4(__float_as_uint(av[h][q][0].x), __float_as_uint(av[h][q][0].y), __float_as_uint(av[h][q][0].z), __float_as_uint(av[h][q][0].w));
              ll = make_uint4(__float_as_uint(av[h][q][NV - 1].x), __float_as_uint(av[h][q][NV - 1].y), __float_as_uint(av[h][q][NV - 1].z), __float_as_uint(av[h][q][NV - 1].w));
              *reinterpret_cast<uint4*>(a_hi + q * GRP + off) = hh;
              *reinterpret_cast<uint4*>(a_lo + q * GRP + off) = ll;
              continue;
            }
            split_f16x2(av[h][q][0].x, av[h][q][0].y, hh.x, ll.x); split_f16x2(av[h][q][0].z, av[h][q][0].w, hh.y, ll.y);
            split_f16x2(av[h][q][NV - 1].x, av[h][q][NV - 1].y, hh.z, ll.z); split_f16x2(av[h][q][NV - 1].z, av[h][q][NV - 1].w, hh.w, ll.w);
            *reinterpret_cast<uint4*>(a_hi + q * GRP + off) = hh;
            *reinterpret_cast<uint4*>(a_lo + q * GRP + off) = ll;
          }
#pragma unroll
          for (int q = 0; q < GB; ++q) {
            float4 b0 = bv[h][q][0], b1 = bv[h][q][NV - 1];
            if (p.dy_presplit) {
              *reinterpret_cast<uint4*>(b_hi + q * GRP + off) = make_uint4(__float_as_uint(b0.x), __float_as_uint(b0.y), __float_as_uint(b0.z), __float_as_uint(b0.w));
              *reinterpret_cast<uint4*>(b_lo + q * GRP + off) = make_uint4(__float_as_uint(b1.x), __float_as_uint(b1.y), __float_as_uint(b1.z), __float_as_uint(b1.w));
              continue;
            }
            if (use_scale) {
              b0 = make_float4(b0.x * b_scale, b0.y * b_scale, b0.z * b_scale, b0.w * b_scale);
              b1 = make_float4(b1.x * b_scale, b1.y * b_scale, b1.z * b_scale, b1.w * b_scale);
            }
            uint4 hh, ll;
            split_f16x2(b0.x, b0.y, hh.x, ll.x); split_f16x2(b0.z, b0.w, hh.y, ll.y);
            split_f16x2(b1.x, b1.y, hh.z, ll.z); split_f16x2(b1.z, b1.w, hh.w, ll.w);
            *reinterpret_cast<uint4*>(b_hi + q * GRP + off) = hh;
            *reinterpret_cast<uint4*>(b_lo + q * GRP + off) = ll;
          }
        } else {
          // atom = 4 pixel rows x 128 B; 32-byte chunk index XOR (row & 3)  (Swizzle<2,5,2> on the byte address)
          const uint32_t off = (uint32_t)((pl >> 2) * 512 + (pl & 3) * 128 + ((((j >> 1) ^ (pl & 3)) << 5) | ((j & 1) << 4)));
#pragma unroll
          for (int q = 0; q < GA; ++q) {
            float hh[4], ll[4];
            split_tf32(av[h][q][0].x, hh[0], ll[0]); split_tf32(av[h][q][0].y, hh[1], ll[1]);
            split_tf32(av[h][q][0].z, hh[2], ll[2]); split_tf32(av[h][q][0].w, hh[3], ll[3]);
            *reinterpret_cast<float4*>(a_hi + q * GRP + off) = make_float4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<float4*>(a_lo + q * GRP + off) = make_float4(ll[0], ll[1], ll[2], ll[3]);
          }
#pragma unroll
          for (int q = 0; q < GB; ++q) {
            float hh[4], ll[4];
            split_tf32(bv[h][q][0].x, hh[0], ll[0]); split_tf32(bv[h][q][0].y, hh[1], ll[1]);
            split_tf32(bv[h][q][0].z, hh[2], ll[2]); split_tf32(bv[h][q][0].w, hh[3], ll[3]);
            *reinterpret_cast<float4*>(b_hi + q * GRP + off) = make_float4(hh[0], hh[1], hh[2], hh[3]);
            *reinterpret_cast<float4*>(b_lo + q * GRP + off) = make_float4(ll[0], ll[1], ll[2], ll[3]);
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&full[s]);
    }

    // ---- epilogue: TMEM lane = (tap,c) row of the tile; 32-column chunks spread over the producer groups
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (STEM) {
      // all MMAs have completed, so the stage buffers are free: sum the per-thread bias partials per channel there and
      // write them as the bias row of this split's partial
      float* s_b = reinterpret_cast<float*>(tiles);
      if (tid < BN) s_b[tid] = 0.f;
      asm volatile("bar.sync 1, %0;" ::"n"(32 * PROD_WARPS) : "memory");
#pragma unroll
      for (int e = 0; e < (STEM ? 8 : 1); ++e) atomicAdd(s_b + 8 * j + e, bsum[e]);
      asm volatile("bar.sync 1, %0;" ::"n"(32 * PROD_WARPS) : "memory");
      if (tid < BN && n0 + tid < g.Cout) p.partial[((size_t)split * (p.K + 1) + p.K) * g.Cout + n0 + tid] = s_b[tid];
    }
    const int row = (warp & 3) * 32 + lane;
    const int kidx = 128 * kt + row;
    float* dst = p.partial + ((size_t)split * (p.K + 1) + (kidx < p.K ? kidx : 0)) * g.Cout + n0;
#pragma unroll 1
    for (int c0 = 32 * group; c0 < BN; c0 += 32 * NGROUPS) {
      float v[32];
      uint32_t raw[32];
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
      tmem_ld_32x32(taddr, raw);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(raw[q]);
      // accumulator sets [main0 | corr0 | main1 | corr1]
      tmem_ld_32x32(taddr + 2 * BN, raw);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] += __uint_as_float(raw[q]);
      float u[32];
      tmem_ld_32x32(taddr + BN, raw);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) u[q] = __uint_as_float(raw[q]);
      tmem_ld_32x32(taddr + 3 * BN, raw);
      tmem_ld_wait();
      constexpr float corr_scale = F16 ? kF16LoInv : 1.f;
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] = fmaf(u[q] + __uint_as_float(raw[q]), corr_scale, v[q]);
      if (use_scale) {
        const float inv = 1.f / b_scale;
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] *= inv;
      }
      if (kidx < p.K) {
#pragma unroll
        for (int q = 0; q < 32; q += 4)
          if (n0 + c0 + q < g.Cout) *reinterpret_cast<float4*>(dst + c0 + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
      }
    }
  } else {
    // ---- MMA issuer
    if (lane == 0) {
      // kind::tf32, D fp32, A and B MN-major (bits 15, 16), N = BN, M = 128
      constexpr uint32_t FMT = F16 ? 0u : 2u;
      const uint32_t idesc = instr_desc(FMT, 128, BN) | (1u << 15) | (1u << 16);
      const uint32_t idesc2 = instr_desc(FMT, 128, 2 * BN) | (1u << 15) | (1u << 16);
      constexpr int KSTEPS = F16 ? PIXS / 16 : 4;     // per stage: 16 (fp16) or 8 (tf32) pixels per MMA
      for (int st = 0; st < n_stages; ++st) {
        const int s = st % S;
        const uint32_t ph = (uint32_t)(st / S) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint32_t base = smem_u32(tiles + (size_t)s * STAGE);
        const uint64_t a_hi = F16 ? smem_desc_mn_sw128_h(base, GRP) : smem_desc_mn_sw128(base);
        const uint64_t a_lo = F16 ? smem_desc_mn_sw128_h(base + A_PART, GRP) : smem_desc_mn_sw128(base + A_PART);
        const uint64_t b_hi = F16 ? smem_desc_mn_sw128_h(base + 2 * A_PART, GRP) : smem_desc_mn_sw128(base + 2 * A_PART);
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk) {
          const uint64_t adv = (uint64_t)(kk * ((F16 ? 2048 : 1024) >> 4));   // next pixel group of one MMA
          const int ks = st * KSTEPS + kk;
          // one N = 2*BN MMA forms a_hi*[b_hi; b_lo] (main | correction), a second adds a_lo*b_hi to the correction half;
          // two accumulator sets alternate per k-step (same scheme as conv_tc.cu)
          const uint32_t d_set = tmem_base + (uint32_t)((ks & 1) * 2 * BN);
          if (F16) {
            mma_bf16(d_set, a_hi + adv, b_hi + adv, idesc2, ks < 2 ? 0u : 1u);
            mma_bf16(d_set + BN, a_lo + adv, b_hi + adv, idesc, 1u);
          } else {
            mma_tf32(d_set, a_hi + adv, b_hi + adv, idesc2, ks < 2 ? 0u : 1u);
            mma_tf32(d_set + BN, a_lo + adv, b_hi + adv, idesc, 1u);
          }
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
  if (warp == PROD_WARPS) tmem_dealloc(tmem_base, NACC * BN);
}

// bias gradient: db[n] = sum_m dy[m][n] -> written into partial[0][K][n]; other splits' bias rows are zeroed
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dy, int M, int C, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[256 * 4];
  const int C4 = C >> 2, c4 = threadIdx.x % C4, ppb = 256 / C4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int m = blockIdx.x * ppb + threadIdx.x / C4; m < M; m += gridDim.x * ppb) {
    const float4 v = *reinterpret_cast<const float4*>(dy + (size_t)m * C + c4 * 4);
    acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) sh[threadIdx.x * 4 + q] = acc[q];
  __syncthreads();
  if (threadIdx.x < C4) {
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int t = threadIdx.x; t < 256; t += C4)
#pragma unroll
      for (int q = 0; q < 4; ++q) s[q] += sh[t * 4 + q];
#pragma unroll
    for (int q = 0; q < 4; ++q) out[(size_t)blockIdx.x * C + c4 * 4 + q] = s[q];
  }
}

// the same from pre-split gradient planes: value = (hi + lo * 2^-11) / f16_operand_scale(*amax); thread = 8 channels of a pixel
__global__ void __launch_bounds__(256) colsum_planes_kernel(const unsigned char* __restrict__ planes, size_t plane_bytes, int M, int C,
                                                            const float* __restrict__ amax, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sh[256 * 8];
  const int C8 = C >> 3, c8 = threadIdx.x % C8, ppb = 256 / C8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int m = blockIdx.x * ppb + threadIdx.x / C8; m < M; m += gridDim.x * ppb) {
    const size_t o = ((size_t)m * C + c8 * 8) * 2;
    const uint4 h = *reinterpret_cast<const uint4*>(planes + o), l = *reinterpret_cast<const uint4*>(planes + plane_bytes + o);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hw[q])), lf = __half22float2(*reinterpret_cast<const __half2*>(&lw[q]));
      acc[2 * q] += fmaf(lf.x, kF16LoInv, hf.x);
      acc[2 * q + 1] += fmaf(lf.y, kF16LoInv, hf.y);
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) sh[threadIdx.x * 8 + q] = acc[q];
  __syncthreads();
  if (threadIdx.x < C8) {
    const float inv = 1.f / f16_operand_scale(amax[0]);
    float sres[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = threadIdx.x; t < 256; t += C8)
#pragma unroll
      for (int q = 0; q < 8; ++q) sres[q] += sh[t * 8 + q];
#pragma unroll
    for (int q = 0; q < 8; ++q) out[(size_t)blockIdx.x * C + c8 * 8 + q] = sres[q] * inv;
  }
}

static inline int pick_bn(int n) { return n <= 64 ? 64 : 128; }

static int plan(const PcConvGeom* g, int* splits, int* rps) {
  const int K = g->R * g->S * g->Cin;
  const long long M = (long long)g->B * g->Ho * g->Wo;
  const int bn = pick_bn(g->Cout);
  const int tiles = ceil_div(K, 128) * ceil_div(g->Cout, bn);
  // fill (at most) two full waves of the 148 SMs: one CTA per SM is resident, so tiles*splits just above a multiple of
  // 148 would cost a whole extra wave
  int sp = (2 * kNumSMs) / tiles;
  const int max_sp = (int)(M / (PIX * 8) > 0 ? M / (PIX * 8) : 1);
  if (sp > max_sp) sp = max_sp;
  if (sp < 1) sp = 1;
  const int pixs = g->Cin == 1 ? 64 : PIX;
  int r = ceil_div(M, sp);
  r = ceil_div(r, pixs) * pixs;
  sp = ceil_div(M, r);
  *splits = sp;
  *rps = r;
  return bn;
}

}  // namespace tcwg
}  // namespace pc

using namespace pc;
using namespace pc::tcwg;

namespace pc { void launch_wgrad_reduce(const float* partial, int n_splits, int R, int S, int Cin, int Cout, float* dw, float* db, pc_stream_t stream,
                                        const float* bias_partial = nullptr, int n_bias = 0); }

// single-channel stem on the FP16X2 tiles: taps as the (padded) M rows
extern "C" int pc_conv_wgrad_tc_stem_supported(const PcConvGeom* g) {
  if (g == nullptr || g->Cin != 1) return 0;
  const long long M = (long long)g->B * g->Ho * g->Wo;
  return (g->R == g->S && g->R * g->S <= 64 && g->stride == 1 && g->pad == g->R / 2 && (g->Cout == 32 || g->Cout == 64) &&
          M * g->Cout < (1LL << 31)) ? 1 : 0;
}

extern "C" int pc_conv_wgrad_tc_supported(const PcConvGeom* g) {
  if (g == nullptr) return 0;
  const bool cout_pow2 = g->Cout >= 32 && g->Cout <= 1024 && (g->Cout & (g->Cout - 1)) == 0;
  const long long M = (long long)g->B * g->Ho * g->Wo;
  const int cmax = g->Cout > g->Cin ? g->Cout : g->Cin;
  return (g->Cin % 32 == 0 && cout_pow2 && g->R * g->S <= 32 && M * cmax < (1LL << 31)) ? 1 : 0;
}

extern "C" size_t pc_conv_wgrad_tc_workspace(const PcConvGeom* g) {
  if (g == nullptr) return 0;
  int sp, rps;
  plan(g, &sp, &rps);
  const size_t part = (size_t)sp * (size_t)(g->R * g->S * g->Cin + 1) * g->Cout * sizeof(float);
  const size_t cs = (size_t)kNumSMs * 2 * g->Cout * sizeof(float);
  return part + cs;
}

extern "C" int pc_conv_wgrad_tc(const float* x, const float* dy, const PcConvGeom* g, const PcInXform* xf, float* dw_oihw, float* db,
                                void* workspace, size_t workspace_bytes, int prec, const float* dy_amax, int dy_presplit,
                                pc_stream_t stream) {
  PC_REQUIRE(x && dy && g && dw_oihw && workspace, PC_EINVAL, "pc_conv_wgrad_tc: null pointer");
  const bool stem = g->Cin == 1;
  PC_REQUIRE(stem ? (prec == PC_PREC_FP16X2 && pc_conv_wgrad_tc_stem_supported(g)) : pc_conv_wgrad_tc_supported(g), PC_EUNSUPPORTED,
             "pc_conv_wgrad_tc: shape not covered (Cin %% 32, Cout %% 4; single-channel stem: FP16X2, k*k <= 64, Cout 32|64)");
  PC_REQUIRE(workspace_bytes >= pc_conv_wgrad_tc_workspace(g), PC_EINVAL, "pc_conv_wgrad_tc: workspace too small");
  int sp, rps;
  const int bn = plan(g, &sp, &rps);
  Params p{};
  p.x = x; p.dy = dy; p.partial = static_cast<float*>(workspace);
  if (xf != nullptr) { p.xf.scale = xf->scale; p.xf.shift = xf->shift; p.xf.drop = xf->drop; p.xf.relu = xf->relu; }
  const bool presplit = xf != nullptr && xf->presplit != 0;
  if (presplit) {
    PC_REQUIRE(prec == PC_PREC_FP16X2 && g->Cin % 64 == 0 && !xf->scale && !xf->shift && !xf->drop && !xf->relu, PC_EINVAL,
               "pc_conv_wgrad: pre-split input planes take no further transform and need PC_PREC_FP16X2 with Cin %% 64 == 0");
    p.plane_bytes = (size_t)g->B * g->H * g->W * g->Cin * 2;
  }
  p.g = *g;
  p.div_wo = FastDiv::make((uint32_t)g->Wo);
  p.div_ho = FastDiv::make((uint32_t)g->Ho);
  p.K = g->R * g->S * g->Cin;
  p.M = g->B * g->Ho * g->Wo;
  p.rows_per_split = rps;
  const bool f16 = prec == PC_PREC_FP16X2 && (g->Cin % 64 == 0 || stem);   // 64-channel groups; otherwise the TF32x3 tiles
  p.dy_amax = f16 ? dy_amax : nullptr;
  if (dy_presplit) {
    PC_REQUIRE(f16 && !stem && dy_amax != nullptr && g->Cout % 8 == 0 && 256 % (g->Cout / 8) == 0, PC_EINVAL,
               "pc_conv_wgrad: pre-split dy needs the FP16X2 tiles (Cin %% 64 == 0), dy_amax, and Cout = 8 * 2^k <= 2048");
    p.dy_presplit = 1;
    p.dy_plane_bytes = (size_t)p.M * g->Cout * 2;
  }
  const uint32_t st = stem ? stage_bytes<64, true, true>() : bn == 64 ? (f16 ? stage_bytes<64, true>() : stage_bytes<64, false>())
                               : (f16 ? stage_bytes<128, true>() : stage_bytes<128, false>());
  int stages = (int)(SMEM_BUDGET / st);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  p.stages = stages;
  const size_t smem = (size_t)stages * st + sizeof(uint64_t) * (2 * MAX_STAGES + 1) + 16 + 1024;
  dim3 grid(ceil_div(p.K, 128), ceil_div(g->Cout, bn), sp);
  // (two co-resident CTAs per SM were tried here as in conv_tc.cu: the register cap made the two-operand gather spill and
  //  the kernel got slower, so the weight gradient keeps one CTA per SM)
#define PC_WG_LAUNCH(BN_, F16_)                                                                                              \
  do {                                                                                                                     \
    static size_t conf = 0;                                                                                                \
    if (smem > conf) {                                                                                                     \
      PC_CUDA(cudaFuncSetAttribute((wgrad_tc_kernel<BN_, 3, 1, F16_>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      conf = smem;                                                                                                         \
    }                                                                                                                      \
    launch_pdl(wgrad_tc_kernel<BN_, 3, 1, F16_>, grid, dim3(32 * 13), smem, stream, p);                                    \
  } while (0)
  if (stem) {
    static size_t conf = 0;
    if (smem > conf) {
      PC_CUDA(cudaFuncSetAttribute((wgrad_tc_kernel<64, 3, 1, true, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      conf = smem;
    }
    launch_pdl(wgrad_tc_kernel<64, 3, 1, true, true>, grid, dim3(32 * 13), smem, stream, p);
  } else if (presplit) {
#define PC_WG_LAUNCH_PS(BN_)                                                                                                 \
  do {                                                                                                                     \
    static size_t conf = 0;                                                                                                \
    if (smem > conf) {                                                                                                     \
      PC_CUDA(cudaFuncSetAttribute((wgrad_tc_kernel<BN_, 3, 1, true, false, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      conf = smem;                                                                                                         \
    }                                                                                                                      \
    launch_pdl(wgrad_tc_kernel<BN_, 3, 1, true, false, true>, grid, dim3(32 * 13), smem, stream, p);                       \
  } while (0)
    if (bn == 64) PC_WG_LAUNCH_PS(64); else PC_WG_LAUNCH_PS(128);
#undef PC_WG_LAUNCH_PS
  } else if (bn == 64) {
    if (f16) PC_WG_LAUNCH(64, true); else PC_WG_LAUNCH(64, false);
  } else {
    if (f16) PC_WG_LAUNCH(128, true); else PC_WG_LAUNCH(128, false);
  }
#undef PC_WG_LAUNCH
  PC_LAUNCH_CHECK("wgrad_tc_kernel");
  // bias gradient: per-CTA column sums of dy -> bias rows of the partial buffer (one row per colsum CTA, appended after the
  // split partials), then the common reduce
  if (stem) {     // the stem kernel wrote its own bias rows
    launch_wgrad_reduce(p.partial, sp, g->R, g->S, g->Cin, g->Cout, dw_oihw, db, stream);
    PC_LAUNCH_CHECK("conv_wgrad_reduce_kernel");
    return PC_OK;
  }
  if (db == nullptr) {   // the caller gets the bias gradient elsewhere (closed form of the BatchNorm backward, csrc/bn_act.cu): no column sums
    launch_wgrad_reduce(p.partial, sp, g->R, g->S, g->Cin, g->Cout, dw_oihw, nullptr, stream);
    PC_LAUNCH_CHECK("conv_wgrad_reduce_kernel");
    return PC_OK;
  }
  float* cs = p.partial + (size_t)sp * (size_t)(p.K + 1) * g->Cout;
  const int cs_ctas = kNumSMs * 2;
  if (dy_presplit)
    launch_pdl(colsum_planes_kernel, dim3(cs_ctas), dim3(256), 0, stream, reinterpret_cast<const unsigned char*>(dy), p.dy_plane_bytes, p.M,
               g->Cout, dy_amax, cs);
  else
    launch_pdl(colsum_kernel, dim3(cs_ctas), dim3(256), 0, stream, dy, p.M, g->Cout, cs);
  PC_LAUNCH_CHECK("colsum_kernel");
  launch_wgrad_reduce(p.partial, sp, g->R, g->S, g->Cin, g->Cout, dw_oihw, db, stream, cs, cs_ctas);   // one launch: dw and db
  PC_LAUNCH_CHECK("conv_wgrad_reduce_kernel");
  return PC_OK;
}
