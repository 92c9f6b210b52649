// Supervised contrastive loss on the tcgen05 tensor cores (label form, D = 64 | 128): the similarity tiles S = F_i F_j^T are
// formed in TMEM with the FP16x2 operand split of conv_tc.cu (fp32-level accuracy) and consumed straight from TMEM by the
// online masked log-sum-exp / positive-sum epilogue, so the N x N logits never reach shared memory, let alone HBM.
// Reference: SupervisedContrastiveLoss.forward, src/training/losses.py:49-84 (same quantities as supcon.cu, whose SIMT
// kernels remain the path for a user mask, small N and other D).
//
//   pack      F [N,D] fp32 -> K-major fp16 hi | lo image [D/64][2][Npad rows][128 B], SWIZZLE_128B: every operand tile of
//             the main kernel is then a contiguous 16 KB (128 rows) or 8 KB (64 rows) block fetched with ONE bulk async copy
//             -- there are no producer warps and no conversion work in the main loop.
//   forward   grid (row tiles of 128, column splits). Per CTA: the A tile (its 128 rows, all of D) is loaded once and stays
//             in shared memory; 64-column B tiles stream through a 3-stage bulk-copy ring; two 256-column TMEM buffers
//             ([main | corr] x 2 alternating sets, see conv_tc.cu) let the MMAs of tile t+1 run under the epilogue of tile t.
//             warps 0-7 epilogue (TMEM lane quarter = warp & 3 -> row, column half = warp >> 2), warp 8 MMA issuer + TMEM
//             owner, warp 9 bulk-copy loader. Each CTA writes partial row statistics (max, sum exp, #pos, sum pos logits) of
//             its column range; supcon_merge_kernel combines the splits into the (m, den, n_pos, s_pos) rows and the row loss.
#include "common.cuh"
#include "tc_common.cuh"

namespace pc {
namespace sctc {

using namespace pc::tc;

constexpr int BM = 128, BN = 64, NST = 3;
constexpr int EPI_WARPS = 8, THREADS = 32 * (EPI_WARPS + 2);

struct Params {
  const unsigned char* Fp;     // packed image (see pack_rows_kernel)
  const long long* labels;
  float* partial;              // [splits][nrows][4]
  int N, Npad, KC, row0, nrows, col_tiles, tiles_per_split;
  float invT;
};

// one thread per (row, 16-byte chunk): 8 consecutive d of one row -> fp16 hi and lo*2^11
__global__ void __launch_bounds__(256) pack_rows_kernel(const float* __restrict__ F, int N, int Npad, int D, unsigned char* __restrict__ out) {
  const int cpr = D >> 3;      // chunks per row
  const long long total = (long long)Npad * cpr;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx / cpr), ch = (int)(idx - (long long)r * cpr);
    const int kc = ch >> 3, j = ch & 7;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (r < N) {
      a = *reinterpret_cast<const float4*>(F + (size_t)r * D + ch * 8);
      b = *reinterpret_cast<const float4*>(F + (size_t)r * D + ch * 8 + 4);
    }
    uint4 h, l;
    split_f16x2(a.x, a.y, h.x, l.x);
    split_f16x2(a.z, a.w, h.y, l.y);
    split_f16x2(b.x, b.y, h.z, l.z);
    split_f16x2(b.z, b.w, h.w, l.w);
    unsigned char* base = out + ((size_t)(kc * 2) * Npad + r) * 128 + (size_t)((j ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(base) = h;
    *reinterpret_cast<uint4*>(base + (size_t)Npad * 128) = l;
  }
}

__global__ void __launch_bounds__(THREADS, 1) supcon_fwd_tc_kernel(const Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int KC = p.KC;
  const uint32_t A_BYTES = (uint32_t)KC * 2u * 16384u, B_STAGE = (uint32_t)KC * 2u * 8192u;
  unsigned char* a_tile = smem;                                   // [KC][hi|lo][128 rows][128 B]
  unsigned char* b_ring = smem + A_BYTES;                         // [NST][KC][hi|lo][64 rows][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)NST * B_STAGE);
  uint64_t* a_full = bars;                 // [1]
  uint64_t* b_full = bars + 1;             // [NST]
  uint64_t* b_empty = b_full + NST;        // [NST]
  uint64_t* acc_full = b_empty + NST;      // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_merge = reinterpret_cast<float*>(tmem_slot + 4);      // [128][4]: column half 1 -> half 0
  long long* s_lab = reinterpret_cast<long long*>(s_merge + 128 * 4);   // [tiles_per_split * 64] labels of my columns

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i0 = p.row0 + blockIdx.x * BM;                       // first global row of this tile
  const int t_begin = blockIdx.y * p.tiles_per_split;
  const int t_end = min(p.col_tiles, t_begin + p.tiles_per_split);
  const int n_tiles = t_end - t_begin;

  if (warp == EPI_WARPS) {
    if (lane == 0) {
      mbar_init(a_full, 1);
      for (int s = 0; s < NST; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 32 * EPI_WARPS); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < EPI_WARPS) {
    // ============================================================ epilogue: thread = (row, column half)
    const int r = (warp & 3) * 32 + lane, half = warp >> 2;
    const int i = i0 + r;
    const long long yi = (i < p.N) ? p.labels[i] : 0;
    // labels of this CTA's whole column range -> shared memory once (broadcast LDS in the loop instead of global loads)
    for (int q = tid; q < n_tiles * BN; q += 32 * EPI_WARPS) {
      const int j = t_begin * BN + q;
      s_lab[q] = j < p.N ? p.labels[j] : 0;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
    // base-2 domain: z2 = z * log2(e), so that exp(z - m) = exp2(z2 - m2) is one ex2.approx
    const float c2 = p.invT * 1.4426950408889634f;
    float m2 = -INFINITY, den = 0.f, npos = 0.f, sraw2 = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int buf = t & 1;
      const int j0 = (t_begin + t) * BN + half * 32;
      const long long* lab = s_lab + t * BN + half * 32;
      mbar_wait(&acc_full[buf], (uint32_t)(t >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(buf * 256 + half * 32);
      uint32_t raw[32];
      float v[32], u[32];
      // accumulator sets [main0 | corr0 | main1 | corr1] of 64 columns each
      tmem_ld_32x32(taddr, raw);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(raw[q]);
      tmem_ld_32x32(taddr + 128, raw);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] += __uint_as_float(raw[q]);
      tmem_ld_32x32(taddr + 64, raw);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) u[q] = __uint_as_float(raw[q]);
      tmem_ld_32x32(taddr + 192, raw);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&acc_empty[buf]);            // TMEM buffer may be overwritten by the MMAs of tile t + 2
      // interior tiles (no diagonal element of this CTA's rows, no column past N) take the check-free path
      const int jt = (t_begin + t) * BN;
      const bool edge = (jt + BN > p.N) || (jt < i0 + BM && jt + BN > i0);
      float tmax = -INFINITY;
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        v[q] = fmaf(u[q] + __uint_as_float(raw[q]), kF16LoInv, v[q]) * c2;     // z2_ij
        if (!edge || j0 + q < p.N) tmax = fmaxf(tmax, v[q]);
      }
      const float m_new = fmaxf(m2, tmax);
      if (i < p.N && m_new > -INFINITY) {
        float dsum = 0.f;
        if (!edge) {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            dsum += exp2f(v[q] - m_new);
            if (yi == lab[q]) { npos += 1.f; sraw2 += v[q]; }
          }
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) {
            const int j = j0 + q;
            if (j < p.N && j != i) {
              dsum += exp2f(v[q] - m_new);
              if (yi == lab[q]) { npos += 1.f; sraw2 += v[q]; }
            }
          }
        }
        den = den * exp2f(m2 - m_new) + dsum;      // exp2(-inf) = 0 on the first tile
        m2 = m_new;
      }
    }
    const float m = m2 * 0.6931471805599453f, sraw = sraw2 * 0.6931471805599453f;   // back to natural units
    // merge the two column halves of a row, then write this split's partial statistics
    if (half == 1) {
      s_merge[r * 4 + 0] = m; s_merge[r * 4 + 1] = den; s_merge[r * 4 + 2] = npos; s_merge[r * 4 + 3] = sraw;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
    if (half == 0) {
      const float m1 = s_merge[r * 4 + 0], d1 = s_merge[r * 4 + 1];
      const float mm = fmaxf(m, m1);
      float d = 0.f;
      if (mm > -INFINITY) d = den * expf(m - mm) + d1 * expf(m1 - mm);
      const int lr = blockIdx.x * BM + r;      // row inside the block [row0, row0 + nrows)
      if (lr < p.nrows && i < p.N) {
        float* dst = p.partial + ((size_t)blockIdx.y * p.nrows + lr) * 4;
        dst[0] = mm; dst[1] = d; dst[2] = npos + s_merge[r * 4 + 2]; dst[3] = sraw + s_merge[r * 4 + 3];
      }
    }
  } else if (warp == EPI_WARPS) {
    // ============================================================ MMA issuer
    if (lane == 0 && n_tiles > 0) {
      const uint32_t idesc = instr_desc(0u, BM, BN), idesc2 = instr_desc(0u, BM, 2 * BN);
      mbar_wait(a_full, 0);
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % NST, buf = t & 1;
        mbar_wait(&b_full[s], (uint32_t)(t / NST) & 1u);
        mbar_wait(&acc_empty[buf], ((uint32_t)(t >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t a_base = smem_u32(a_tile), b_base = smem_u32(b_ring + (size_t)s * B_STAGE);
        for (int kc = 0; kc < KC; ++kc) {
          const uint64_t a_hi = smem_desc_sw128(a_base + (uint32_t)(kc * 2) * 16384u);
          const uint64_t a_lo = smem_desc_sw128(a_base + (uint32_t)(kc * 2 + 1) * 16384u);
          const uint64_t b_hi = smem_desc_sw128(b_base + (uint32_t)(kc * 2) * 8192u);     // b_lo follows: N = 128 covers both
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 2);
            const int ks = kc * 4 + kk;
            const uint32_t d_set = tmem_base + (uint32_t)(buf * 256 + (ks & 1) * 128);
            mma_bf16(d_set, a_hi + adv, b_hi + adv, idesc2, ks < 2 ? 0u : 1u);
            mma_bf16(d_set + 64, a_lo + adv, b_hi + adv, idesc, 1u);
          }
        }
        mma_commit(&b_empty[s]);
        mma_commit(&acc_full[buf]);
      }
    }
    __syncwarp();
  } else {
    // ============================================================ bulk-copy loader
    if (lane == 0 && n_tiles > 0) {
      const size_t part_stride = (size_t)p.Npad * 128;
      mbar_arrive_expect_tx(a_full, A_BYTES);
      for (int q = 0; q < KC * 2; ++q) bulk_g2s(a_tile + (size_t)q * 16384, p.Fp + (size_t)q * part_stride + (size_t)i0 * 128, 16384, a_full);
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % NST;
        mbar_wait(&b_empty[s], ((uint32_t)(t / NST) & 1u) ^ 1u);
        unsigned char* dst = b_ring + (size_t)s * B_STAGE;
        const size_t j0 = (size_t)(t_begin + t) * BN;
        mbar_arrive_expect_tx(&b_full[s], B_STAGE);
        for (int q = 0; q < KC * 2; ++q) bulk_g2s(dst + (size_t)q * 8192, p.Fp + (size_t)q * part_stride + j0 * 128, 8192, &b_full[s]);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == EPI_WARPS) tmem_dealloc(tmem_base, 512);
}

// combine the column splits of each row: stats (m, den + 1e-6, n_pos, s_pos) and the row loss (losses.py:66-80)
__global__ void __launch_bounds__(256) supcon_merge_kernel(const float* __restrict__ partial, int splits, int nrows, float t_ratio,
                                                           float* __restrict__ stats, float* __restrict__ row_loss) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  float m = -INFINITY;
  for (int s = 0; s < splits; ++s) m = fmaxf(m, partial[((size_t)s * nrows + r) * 4]);
  float den = 0.f, npos = 0.f, sraw = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float* q = partial + ((size_t)s * nrows + r) * 4;
    if (q[0] > -INFINITY) den += q[1] * expf(q[0] - m);
    npos += q[2];
    sraw += q[3];
  }
  const float d = den + 1e-6f;
  const float spos = sraw - npos * m;
  const float nn = (npos == 0.f) ? 1.f : npos;
  float* st = stats + (size_t)r * 4;
  st[0] = m; st[1] = d; st[2] = npos; st[3] = spos;
  row_loss[r] = -t_ratio * (spos - npos * logf(d)) / nn;
}

// ---------------------------------------------------------------------------------------------------------------- backward
// dF_i = (c / T) * sum_{j != i} w'_ij F_j,   w'_ij = e_ij - [y_i == y_j] (1/n_i + 1/n_j),
// e_ij = exp(z_ij - m_i) / den_i + exp(z_ij - m_j) / den_j          (same algebra as supcon_bwd_kernel in supcon.cu)
//
// Per column tile: MMA1 forms S = F_i F_j^T in TMEM columns [0,256); the epilogue warps turn it into W' (fp32), split it into
// fp16 hi | lo and store it as the K-major A operand of MMA2; MMA2 accumulates dF[128 x D] += W'[128 x 64] F_j[64 x D] into
// TMEM columns [256,512). The B operand of MMA2 is the SAME shared-memory tile MMA1 used: a [64 rows j][128 B = 64 d] block
// with the 128-byte swizzle is a K-major operand when its rows are the N index (MMA1) and an MN-major operand when its rows
// are the K index (MMA2), so F_j is fetched once. With LBO = 8 KB one N = 2D MMA walks [hi g0 | lo g0 | hi g1 | lo g1]
// (g = 64-wide d group), i.e. main and correction columns interleave per group; the w_lo * b_hi term is one N = 64 MMA per group.
struct BwdParams {
  const unsigned char* Fp;
  const long long* labels;
  const float* stats_all;      // [N][4] (m, den, n_pos, s_pos) of ALL rows
  const float* grad_scale;     // device scalar or null
  float* partial;              // [splits][nrows][D]
  int N, Npad, KC, D, row0, nrows, col_tiles, tiles_per_split;
  float invT, coef;
};

__device__ __forceinline__ uint64_t smem_desc_mn_h(uint32_t smem_addr_bytes, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr_bytes >> 4) & 0x3FFF);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;     // next 64-wide d group
  d |= (uint64_t)(1024 >> 4) << 32;          // next 8 rows (K direction)
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}

constexpr int NSTB = 2;

__global__ void __launch_bounds__(THREADS, 1) supcon_bwd_tc_kernel(const BwdParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int KC = p.KC;
  const uint32_t A_BYTES = (uint32_t)KC * 2u * 16384u, B_STAGE = (uint32_t)KC * 2u * 8192u;
  unsigned char* a_tile = smem;
  unsigned char* b_ring = a_tile + A_BYTES;                         // [NSTB][KC][hi|lo][64 rows][128 B]
  unsigned char* w_buf = b_ring + (size_t)NSTB * B_STAGE;           // [2][hi|lo][128 rows][128 B]
  float4* cs_f = reinterpret_cast<float4*>(w_buf + 2 * 32768);      // [2][64] (m2_j, h_j/den_j, 1/n_j, -)
  long long* cs_y = reinterpret_cast<long long*>(cs_f + 2 * BN);    // [2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(cs_y + 2 * BN);
  uint64_t* a_full = bars;
  uint64_t* b_full = bars + 1;            // [NSTB]
  uint64_t* b_empty = b_full + NSTB;      // [NSTB]
  uint64_t* s_full = b_empty + NSTB;      // [1]
  uint64_t* s_empty = s_full + 1;         // [1]
  uint64_t* w_full = s_empty + 1;         // [2]
  uint64_t* w_empty = w_full + 2;         // [2]
  uint64_t* d_full = w_empty + 2;         // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int i0 = p.row0 + blockIdx.x * BM;
  const int t_begin = blockIdx.y * p.tiles_per_split;
  const int n_tiles = min(p.col_tiles, t_begin + p.tiles_per_split) - t_begin;

  if (warp == EPI_WARPS) {
    if (lane == 0) {
      mbar_init(a_full, 1);
      for (int s = 0; s < NSTB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
      mbar_init(s_full, 1);
      mbar_init(s_empty, 32 * EPI_WARPS);
      for (int b = 0; b < 2; ++b) { mbar_init(&w_full[b], 32 * EPI_WARPS); mbar_init(&w_empty[b], 1); }
      mbar_init(d_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr float kLog2e = 1.4426950408889634f;

  if (warp < EPI_WARPS) {
    // ============================================================ S -> W' (thread = row, column half), then the dF read-out
    const int r = (warp & 3) * 32 + lane, half = warp >> 2;
    const int i = i0 + r;
    float mi2 = 0.f, idi = 0.f, ini = 0.f;
    long long yi = 0;
    if (i < p.N) {
      const float4 st = *reinterpret_cast<const float4*>(p.stats_all + (size_t)i * 4);
      mi2 = st.x * kLog2e;
      idi = (st.z != 0.f ? 1.f : 0.f) / st.y;
      ini = st.z != 0.f ? 1.f / st.z : 0.f;
      yi = p.labels[i];
    }
    const float c2 = p.invT * kLog2e;
    for (int t = 0; t < n_tiles; ++t) {
      const int jt = (t_begin + t) * BN;
      const int cb = t & 1;
      if (tid < BN) {                       // column statistics of this tile -> shared memory
        const int j = jt + tid;
        float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
        long long y = 0;
        if (j < p.N) {
          const float4 st = *reinterpret_cast<const float4*>(p.stats_all + (size_t)j * 4);
          c = make_float4(st.x * kLog2e, (st.z != 0.f ? 1.f : 0.f) / st.y, st.z != 0.f ? 1.f / st.z : 0.f, 0.f);
          y = p.labels[j];
        }
        cs_f[cb * BN + tid] = c;
        cs_y[cb * BN + tid] = y;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
      mbar_wait(s_full, (uint32_t)t & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * 32);
      uint32_t raw[32];
      float v[32], u[32];
      tmem_ld_32x32(taddr, raw);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(raw[q]);
      tmem_ld_32x32(taddr + 128, raw);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] += __uint_as_float(raw[q]);
      tmem_ld_32x32(taddr + 64, raw);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 32; ++q) u[q] = __uint_as_float(raw[q]);
      tmem_ld_32x32(taddr + 192, raw);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_empty);                 // S may be overwritten by MMA1 of the next tile
      const float4* cf = cs_f + cb * BN + half * 32;
      const long long* cy = cs_y + cb * BN + half * 32;
#pragma unroll
      for (int q = 0; q < 32; ++q) {
        const int j = jt + half * 32 + q;
        const float z2 = fmaf(u[q] + __uint_as_float(raw[q]), kF16LoInv, v[q]) * c2;
        const float4 c = cf[q];
        float w = exp2f(z2 - mi2) * idi + exp2f(z2 - c.x) * c.y;
        if (yi == cy[q]) w -= ini + c.z;
        v[q] = (i < p.N && j < p.N && j != i) ? w : 0.f;
      }
      mbar_wait(&w_empty[cb], (((uint32_t)t >> 1) & 1u) ^ 1u);
      unsigned char* w_hi = w_buf + (size_t)cb * 32768 + (size_t)r * 128;
      unsigned char* w_lo = w_hi + 16384;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint4 h, l;
        split_f16x2(v[8 * k + 0], v[8 * k + 1], h.x, l.x);
        split_f16x2(v[8 * k + 2], v[8 * k + 3], h.y, l.y);
        split_f16x2(v[8 * k + 4], v[8 * k + 5], h.z, l.z);
        split_f16x2(v[8 * k + 6], v[8 * k + 7], h.w, l.w);
        const uint32_t off = (uint32_t)(((half * 4 + k) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(w_hi + off) = h;
        *reinterpret_cast<uint4*>(w_lo + off) = l;
      }
      fence_proxy_async();
      mbar_arrive(&w_full[cb]);
    }
    // ---- read-out: thread (row, half) owns d group `half` (64 columns) of its row
    if (n_tiles > 0) {
      mbar_wait(d_full, 0);
      tc_fence_after();
    }
    const float scale = p.coef * (p.grad_scale != nullptr ? p.grad_scale[0] : 1.f) * p.invT;
    const int lr = blockIdx.x * BM + r;
    if (half < KC) {
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        float v[32];
        if (n_tiles > 0) {
          uint32_t raw[32], rc[32];
          const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(256 + half * 128 + cc * 32);
          tmem_ld_32x32(taddr, raw);
          tmem_ld_32x32(taddr + 64, rc);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = fmaf(__uint_as_float(rc[q]), kF16LoInv, __uint_as_float(raw[q])) * scale;
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = 0.f;
        }
        if (lr < p.nrows && i < p.N) {
          float* dst = p.partial + ((size_t)blockIdx.y * p.nrows + lr) * p.D + half * 64 + cc * 32;
#pragma unroll
          for (int q = 0; q < 32; q += 4) *reinterpret_cast<float4*>(dst + q) = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
        }
      }
    }
  } else if (warp == EPI_WARPS) {
    // ============================================================ MMA issuer
    if (lane == 0 && n_tiles > 0) {
      const uint32_t idesc_s = instr_desc(0u, BM, BN), idesc_s2 = instr_desc(0u, BM, 2 * BN);
      const uint32_t idesc_d2 = instr_desc(0u, BM, 2 * 64 * KC) | (1u << 16);     // B MN-major, N = 2D
      const uint32_t idesc_d1 = instr_desc(0u, BM, 64) | (1u << 16);
      const uint32_t a_base = smem_u32(a_tile);
      auto mma1 = [&](int t) {
        const uint32_t b_base = smem_u32(b_ring + (size_t)(t % NSTB) * B_STAGE);
        for (int kc = 0; kc < KC; ++kc) {
          const uint64_t a_hi = smem_desc_sw128(a_base + (uint32_t)(kc * 2) * 16384u);
          const uint64_t a_lo = smem_desc_sw128(a_base + (uint32_t)(kc * 2 + 1) * 16384u);
          const uint64_t b_hi = smem_desc_sw128(b_base + (uint32_t)(kc * 2) * 8192u);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 2);
            const int ks = kc * 4 + kk;
            const uint32_t d_set = tmem_base + (uint32_t)((ks & 1) * 128);
            mma_bf16(d_set, a_hi + adv, b_hi + adv, idesc_s2, ks < 2 ? 0u : 1u);
            mma_bf16(d_set + 64, a_lo + adv, b_hi + adv, idesc_s, 1u);
          }
        }
      };
      mbar_wait(a_full, 0);
      mbar_wait(&b_full[0], 0);
      tc_fence_after();
      mma1(0);
      mma_commit(s_full);
      for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) {
          mbar_wait(&b_full[(t + 1) % NSTB], (uint32_t)((t + 1) / NSTB) & 1u);
          mbar_wait(s_empty, (uint32_t)t & 1u);
          tc_fence_after();
          mma1(t + 1);
          mma_commit(s_full);
        }
        const int cb = t & 1;
        mbar_wait(&w_full[cb], ((uint32_t)t >> 1) & 1u);
        tc_fence_after();
        const uint32_t w_hi = smem_u32(w_buf + (size_t)cb * 32768), w_lo = w_hi + 16384u;
        const uint32_t b_base = smem_u32(b_ring + (size_t)(t % NSTB) * B_STAGE);
        const uint64_t wa_hi = smem_desc_sw128(w_hi), wa_lo = smem_desc_sw128(w_lo);
        const uint64_t b_all = smem_desc_mn_h(b_base, 8192u);           // [hi g0 | lo g0 | hi g1 | lo g1]
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t adv_a = (uint64_t)(kk * 2), adv_b = (uint64_t)(kk * (2048 >> 4));
          const uint32_t acc = (t == 0 && kk == 0) ? 0u : 1u;
          mma_bf16(tmem_base + 256, wa_hi + adv_a, b_all + adv_b, idesc_d2, acc);
          for (int g = 0; g < KC; ++g) {
            const uint64_t b_g = smem_desc_mn_h(b_base + (uint32_t)g * 16384u, 8192u);
            mma_bf16(tmem_base + 256 + (uint32_t)(g * 128 + 64), wa_lo + adv_a, b_g + adv_b, idesc_d1, 1u);
          }
        }
        mma_commit(&b_empty[t % NSTB]);
        mma_commit(&w_empty[cb]);
      }
      mma_commit(d_full);
    }
    __syncwarp();
  } else {
    // ============================================================ bulk-copy loader
    if (lane == 0 && n_tiles > 0) {
      const size_t part_stride = (size_t)p.Npad * 128;
      mbar_arrive_expect_tx(a_full, A_BYTES);
      for (int q = 0; q < KC * 2; ++q) bulk_g2s(a_tile + (size_t)q * 16384, p.Fp + (size_t)q * part_stride + (size_t)i0 * 128, 16384, a_full);
      for (int t = 0; t < n_tiles; ++t) {
        const int s = t % NSTB;
        mbar_wait(&b_empty[s], ((uint32_t)(t / NSTB) & 1u) ^ 1u);
        unsigned char* dst = b_ring + (size_t)s * B_STAGE;
        const size_t j0 = (size_t)(t_begin + t) * BN;
        mbar_arrive_expect_tx(&b_full[s], B_STAGE);
        for (int q = 0; q < KC * 2; ++q) bulk_g2s(dst + (size_t)q * 8192, p.Fp + (size_t)q * part_stride + j0 * 128, 8192, &b_full[s]);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == EPI_WARPS) tmem_dealloc(tmem_base, 512);
}

// dF[r][:] = sum over column splits, in split order (deterministic)
__global__ void __launch_bounds__(256) supcon_dF_reduce_kernel(const float* __restrict__ partial, int splits, long long n4, float* __restrict__ dF) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n4; k += (long long)gridDim.x * blockDim.x) {
    float4 a = reinterpret_cast<const float4*>(partial)[k];
    for (int s = 1; s < splits; ++s) {
      const float4 b = reinterpret_cast<const float4*>(partial)[(size_t)s * n4 + k];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    reinterpret_cast<float4*>(dF)[k] = a;
  }
}

// one extra all-zero tile of rows: a 128-row A tile starting at any row0 + 128k <= N stays inside the image
static inline int npad_of(int N) { return (ceil_div(N, 128) + 1) * 128; }
static inline int splits_of(int N, int nrows) {
  const int row_tiles = ceil_div(nrows, BM), col_tiles = ceil_div(N, BN);
  int sp = (2 * kNumSMs) / row_tiles;                 // about two CTAs' worth of work per SM
  const int max_sp = col_tiles / 8 > 0 ? col_tiles / 8 : 1;      // at least 8 column tiles per CTA (the A tile load is amortised)
  if (sp > max_sp) sp = max_sp;
  if (sp < 1) sp = 1;
  int tps = ceil_div(col_tiles, sp);
  if (tps > 64) tps = 64;                              // the CTA keeps its columns' labels in shared memory (<= 32 KB)
  return ceil_div(col_tiles, tps);
}

}  // namespace sctc
}  // namespace pc

using namespace pc;
using namespace pc::sctc;

extern "C" int pc_supcon_tc_supported(int N, int D, int row0, int nrows) {
  return (N >= 128 && (D == 64 || D == 128) && row0 % 8 == 0 && nrows > 0 && row0 >= 0 && row0 + nrows <= N) ? 1 : 0;
}

extern "C" size_t pc_supcon_tc_workspace(int N, int D, int nrows) {
  if (N <= 0 || D <= 0 || nrows <= 0) return 0;
  const size_t packed = (size_t)(D / 64) * 2 * npad_of(N) * 128;
  const size_t part = (size_t)splits_of(N, nrows) * nrows * 4 * sizeof(float);
  return packed + part + 256;
}

extern "C" int pc_supcon_fwd_tc(const float* F, const int64_t* labels, int N, int D, int row0, int nrows, float temperature,
                                float base_temperature, void* workspace, size_t workspace_bytes, float* stats, float* row_loss,
                                pc_stream_t stream) {
  PC_REQUIRE(N > 1, PC_EINVAL, "Batch size must be greater than 1 for contrastive loss");  // losses.py:44-45
  PC_REQUIRE(F && labels && stats && row_loss && workspace, PC_EINVAL, "pc_supcon_fwd_tc: null pointer");
  PC_REQUIRE(pc_supcon_tc_supported(N, D, row0, nrows), PC_EUNSUPPORTED,
             "pc_supcon_fwd_tc: needs N >= 128, D = 64 | 128, row0 %% 8 == 0 (got N=%d D=%d row0=%d nrows=%d)", N, D, row0, nrows);
  PC_REQUIRE(temperature > 0.f && base_temperature > 0.f, PC_EINVAL, "pc_supcon_fwd_tc: temperatures must be positive");
  PC_REQUIRE(workspace_bytes >= pc_supcon_tc_workspace(N, D, nrows), PC_EINVAL, "pc_supcon_fwd_tc: workspace too small");
  PC_REQUIRE((reinterpret_cast<uintptr_t>(F) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 127) == 0, PC_EINVAL,
             "pc_supcon_fwd_tc: F must be 16-byte and the workspace 128-byte aligned");
  const int Npad = npad_of(N), KC = D / 64;
  unsigned char* Fp = static_cast<unsigned char*>(workspace);
  const size_t packed = (size_t)KC * 2 * Npad * 128;
  float* partial = reinterpret_cast<float*>(Fp + ((packed + 255) & ~(size_t)255));
  const long long chunks = (long long)Npad * (D / 8);
  int pgrid = ceil_div(chunks, 256);
  if (pgrid > kNumSMs * 8) pgrid = kNumSMs * 8;
  pack_rows_kernel<<<pgrid, 256, 0, stream>>>(F, N, Npad, D, Fp);
  PC_LAUNCH_CHECK("supcon pack_rows_kernel");

  Params p{};
  p.Fp = Fp; p.labels = reinterpret_cast<const long long*>(labels); p.partial = partial;
  p.N = N; p.Npad = Npad; p.KC = KC; p.row0 = row0; p.nrows = nrows;
  p.col_tiles = ceil_div(N, BN);
  const int sp = splits_of(N, nrows);
  p.tiles_per_split = ceil_div(p.col_tiles, sp);
  p.invT = 1.0f / temperature;
  const size_t smem = (size_t)KC * 2 * 16384 + (size_t)NST * KC * 2 * 8192 + sizeof(uint64_t) * (1 + 2 * NST + 4) + 16 + sizeof(float) * 128 * 4 +
                      sizeof(long long) * (size_t)p.tiles_per_split * BN + 1024;
  static size_t conf = 0;
  if (smem > conf) {
    PC_CUDA(cudaFuncSetAttribute(supcon_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  dim3 grid(ceil_div(nrows, BM), sp);
  supcon_fwd_tc_kernel<<<grid, THREADS, smem, stream>>>(p);
  PC_LAUNCH_CHECK("supcon_fwd_tc_kernel");
  supcon_merge_kernel<<<ceil_div(nrows, 256), 256, 0, stream>>>(partial, sp, nrows, temperature / base_temperature, stats, row_loss);
  PC_LAUNCH_CHECK("supcon_merge_kernel");
  return PC_OK;
}

extern "C" size_t pc_supcon_bwd_tc_workspace(int N, int D, int nrows) {
  if (N <= 0 || D <= 0 || nrows <= 0) return 0;
  const size_t packed = (size_t)(D / 64) * 2 * npad_of(N) * 128;
  const size_t part = (size_t)splits_of(N, nrows) * nrows * D * sizeof(float);
  return packed + part + 256;
}

extern "C" int pc_supcon_bwd_tc(const float* F, const int64_t* labels, int N, int D, int row0, int nrows, float temperature, float coef,
                                const float* grad_scale, const float* stats_all, void* workspace, size_t workspace_bytes, float* dF,
                                pc_stream_t stream) {
  PC_REQUIRE(N > 1, PC_EINVAL, "Batch size must be greater than 1 for contrastive loss");
  PC_REQUIRE(F && labels && stats_all && dF && workspace, PC_EINVAL, "pc_supcon_bwd_tc: null pointer");
  PC_REQUIRE(pc_supcon_tc_supported(N, D, row0, nrows), PC_EUNSUPPORTED,
             "pc_supcon_bwd_tc: needs N >= 128, D = 64 | 128, row0 %% 8 == 0 (got N=%d D=%d row0=%d nrows=%d)", N, D, row0, nrows);
  PC_REQUIRE(workspace_bytes >= pc_supcon_bwd_tc_workspace(N, D, nrows), PC_EINVAL, "pc_supcon_bwd_tc: workspace too small");
  PC_REQUIRE((reinterpret_cast<uintptr_t>(F) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 127) == 0 &&
                 (reinterpret_cast<uintptr_t>(stats_all) & 15) == 0 && (reinterpret_cast<uintptr_t>(dF) & 15) == 0,
             PC_EINVAL, "pc_supcon_bwd_tc: F, stats_all, dF must be 16-byte and the workspace 128-byte aligned");
  const int Npad = npad_of(N), KC = D / 64;
  unsigned char* Fp = static_cast<unsigned char*>(workspace);
  const size_t packed = (size_t)KC * 2 * Npad * 128;
  float* partial = reinterpret_cast<float*>(Fp + ((packed + 255) & ~(size_t)255));
  const long long chunks = (long long)Npad * (D / 8);
  int pgrid = ceil_div(chunks, 256);
  if (pgrid > kNumSMs * 8) pgrid = kNumSMs * 8;
  pack_rows_kernel<<<pgrid, 256, 0, stream>>>(F, N, Npad, D, Fp);
  PC_LAUNCH_CHECK("supcon pack_rows_kernel");

  BwdParams p{};
  p.Fp = Fp; p.labels = reinterpret_cast<const long long*>(labels); p.stats_all = stats_all; p.grad_scale = grad_scale;
  p.partial = partial;
  p.N = N; p.Npad = Npad; p.KC = KC; p.D = D; p.row0 = row0; p.nrows = nrows;
  p.col_tiles = ceil_div(N, BN);
  const int sp = splits_of(N, nrows);
  p.tiles_per_split = ceil_div(p.col_tiles, sp);
  p.invT = 1.0f / temperature;
  p.coef = coef;
  const size_t smem = (size_t)KC * 2 * 16384 + (size_t)NSTB * KC * 2 * 8192 + 2 * 32768 + 2 * BN * (sizeof(float4) + sizeof(long long)) +
                      sizeof(uint64_t) * 16 + 16 + 1024;
  static size_t conf = 0;
  if (smem > conf) {
    PC_CUDA(cudaFuncSetAttribute(supcon_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conf = smem;
  }
  dim3 grid(ceil_div(nrows, BM), sp);
  supcon_bwd_tc_kernel<<<grid, THREADS, smem, stream>>>(p);
  PC_LAUNCH_CHECK("supcon_bwd_tc_kernel");
  const long long n4 = (long long)nrows * D / 4;
  int rgrid = ceil_div(n4, 256);
  if (rgrid > kNumSMs * 8) rgrid = kNumSMs * 8;
  supcon_dF_reduce_kernel<<<rgrid, 256, 0, stream>>>(partial, sp, n4, dF);
  PC_LAUNCH_CHECK("supcon_dF_reduce_kernel");
  return PC_OK;
}
