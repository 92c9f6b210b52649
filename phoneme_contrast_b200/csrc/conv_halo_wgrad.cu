// Halo-resident tcgen05 weight gradient for the stride-1 3x3 "same" convolutions on the FP16X2 operand planes.
//
//   dW[o][c][r][s] = sum over positions q of  dy[q][o] * x[q + (r-1) (W+1) + (s-1)][c]
//
// in the PADDED position space of csrc/conv_halo.cu (q = (b (H+1) + hp) (W+1) + wp; hp = 0 / wp = 0 are a zero row / column that the
// TMA unit materialises by out-of-bounds fill), where a filter tap is a pure row shift. The round-1 kernel (conv_tc_wgrad.cu)
// re-gathers x for every tap and dy for every 128-row (tap, channel) tile: measured, its time IS its L2 -> shared-memory traffic
// (64 -> 64 channels at 20x51: 9 x 67 MB + 5 x 67 MB = 0.94 GB at the ~7 TB/s every kernel here tops out at = 135 us for 42 us of
// tensor work). Here both operands are loaded ONCE per position tile by tiled TMA boxes and the taps come from descriptors:
//   * reduction index k = position. Both operands are MN-major (a 128-byte row = 64 channels of one position, SWIZZLE_128B as the
//     TMA writes it), so advancing the reduction index by 16 positions is + 2048 bytes on a descriptor start address.
//   * B = x: the three column taps ds = -1, 0, +1 are three N groups of ONE region, group stride LBO = 128 bytes (one position):
//     N = 3 x 64 channels per instruction. The row tap dr is a different unit of work (the region is loaded with its rows shifted
//     by dr, again by TMA coordinates).
//   * A = dy: M = 128 output channels (two 64-channel regions, LBO = region stride). Layers with only 64 output channels fill
//     M = 128 with TWO row taps instead: group 1 = the same region one padded row (LBO = (W+1) * 128 bytes) further, which pairs
//     dy[k + (W+1)] with x[k + ds], i.e. tap dr - 1; nine taps then take two passes (dr in {0, -1}, then {+1, unused}).
//   * three products per k-step (dy_hi x_hi -> main; dy_hi x_lo, dy_lo x_hi -> corr), accumulators [128 x 192] main | corr in
//     TMEM (384 of 512 columns), read out once per CTA; partials [unit][split][128][192] reduced by wgrad_halo_reduce_kernel.
//   * positions past the end of a tile's data multiply ZERO rows of x (each x buffer keeps 8 zero guard rows in front of and a
//     zero tail behind the rows TMA writes), so the k-step count is simply ceil(tile positions / 16).
// Descriptor start addresses that are only 128-byte aligned and group strides below the 1024-byte swizzle atom are legal because
// the swizzle is a function of absolute shared-memory address bits for the TMA write and the UMMA read alike (the forward engine
// relies on the former, csrc/stem_bwd.cu on arbitrary LBOs; tests/test_gpu_halo.py checks this kernel against fp64).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pc {
namespace halowg {

using namespace pc::tc;

constexpr int NSTAGE = 2;
constexpr int EPI_WARPS = 4;
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int NN = 192;              // N of every MMA: 3 column taps x 64 input channels
constexpr int GUARD = 8;             // zero rows in front of the x rows
constexpr int MAX_BOX = 8;

struct Params {
  float* partial;            // [n_units][splits][128][192]
  const float* dy_amax;      // device scalar behind the power-of-two scale of the dy planes
  int B, H, W, Cin, Cout;
  int Wp, RB, NB, n_box;     // box = NB images x RB padded rows x (W+1) positions; a tile = n_box boxes
  int box_pos, tile_pos;     // positions per box / tile
  int n_tiles;               // position tiles over the whole tensor
  int pair;                  // 1: 64 output channels, M groups = two row taps (see above)
  int n_units, splits;
  int n_cc, n_ot;            // 64-channel input chunks, 128-channel output tiles
  int nks;                   // k-steps per tile = ceil(tile_pos / 16)
  uint32_t dy_chunk_bytes;   // one 64-channel dy region (one plane), rounded up to 1024
  uint32_t x_plane_bytes;    // one x plane buffer (guard + rows + tail), rounded up to 1024
  uint32_t stage_bytes;
  uint32_t tx_bytes;         // bytes the TMA boxes of one stage deliver
};

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}
// MN-major 16-bit operand, SWIZZLE_128B: LBO = byte distance between 64-element groups of the M / N dimension, SBO = 1024 (8 rows of k)
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// unit -> (output tile, input chunk, row tap of M group 0 relative to the x rows loaded, row shift of the x rows)
struct Unit { int ot, cc, dr_x; };
__device__ __forceinline__ Unit decode_unit(const Params& p, int u) {
  Unit r;
  if (p.pair) {              // unit = (cc, pass): pass 0 loads x unshifted (M groups: dr = 0, -1), pass 1 shifted by +1 (dr = +1, unused)
    r.ot = 0; r.cc = u >> 1; r.dr_x = (u & 1);
  } else {                   // unit = ((ot, cc), dr)
    const int dr = u % 3, rest = u / 3;
    r.dr_x = dr - 1; r.cc = rest % p.n_cc; r.ot = rest / p.n_cc;
  }
  return r;
}

__global__ void __launch_bounds__(THREADS, 1) wgrad_halo_kernel(const __grid_constant__ CUtensorMap dymap, const __grid_constant__ CUtensorMap xmap,
                                                                const Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // stage: [dy chunk 0: hi | lo] [dy chunk 1: hi | lo] [x: hi | lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NSTAGE * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + NSTAGE;
  uint64_t* acc_full = bars + 2 * NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int unit = blockIdx.x / p.splits, split = blockIdx.x - unit * p.splits;
  const Unit un = decode_unit(p, unit);
  const int per = (p.n_tiles + p.splits - 1) / p.splits;
  const int t_begin = split * per, t_end = min(p.n_tiles, t_begin + per);
  const int n_my = max(0, t_end - t_begin);
  const int n_dy_chunks = p.pair ? 1 : 2;

  // everything TMA does not write stays zero for the CTA's life (x guard rows / tails, dy tails)
  for (uint32_t i = tid; i < (uint32_t)NSTAGE * p.stage_bytes / 16u; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(acc_full, 1);
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
  const uint32_t dy_bytes = 2u * p.dy_chunk_bytes;                  // hi | lo of one chunk

  if (warp == 0) {
    // ================================================================================= TMA producer
    if (lane == 0) {
      const int dy_rb = p.pair ? p.RB + 1 : p.RB;     // the pair mode stages one more padded row per box (M group 1 reads one row ahead)
      (void)dy_rb;
      for (int i = 0; i < n_my; ++i) {
        const int s = i % NSTAGE;
        mbar_wait(&empty[s], (((uint32_t)(i / NSTAGE)) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&full[s], p.tx_bytes);
        unsigned char* st = smem + (size_t)s * p.stage_bytes;
        const int tile = t_begin + i;
        for (int j = 0; j < p.n_box; ++j) {
          const int bx = tile * p.n_box + j;                         // global box index
          int b, hp0;
          if (p.NB > 1) { b = bx * p.NB; hp0 = 0; }
          else {
            const int rg = bx * p.RB;                                // global padded row of the box's first row
            b = rg / (p.H + 1);
            hp0 = rg - b * (p.H + 1);
          }
          const uint32_t boff = (uint32_t)j * (uint32_t)p.box_pos * 128u;
          for (int ch = 0; ch < n_dy_chunks; ++ch) {
            unsigned char* d = st + (size_t)ch * dy_bytes + boff;
            const int c0 = un.ot * 128 + ch * 64;
            tma_load_5d(d, &dymap, &full[s], c0, -1, hp0 - 1, b, 0);
            tma_load_5d(d + p.dy_chunk_bytes, &dymap, &full[s], c0, -1, hp0 - 1, b, 1);
          }
          unsigned char* xd = st + (size_t)n_dy_chunks * dy_bytes + (uint32_t)GUARD * 128u + boff;
          tma_load_5d(xd, &xmap, &full[s], un.cc * 64, -1, hp0 - 1 + un.dr_x, b, 0);
          tma_load_5d(xd + p.x_plane_bytes, &xmap, &full[s], un.cc * 64, -1, hp0 - 1 + un.dr_x, b, 1);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================================= MMA issuer
    if (lane == 0 && n_my > 0) {
      const uint32_t idesc = instr_desc(0u, 128, NN) | (1u << 15) | (1u << 16);       // both operands MN-major
      const uint32_t a_lbo = p.pair ? (uint32_t)p.Wp * 128u : dy_bytes;
      const uint32_t d_main = tmem_base, d_corr = tmem_base + (uint32_t)NN;
      for (int i = 0; i < n_my; ++i) {
        const int s = i % NSTAGE;
        mbar_wait(&full[s], ((uint32_t)(i / NSTAGE)) & 1u);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + (size_t)s * p.stage_bytes);
        const uint32_t x_hi_addr = st + (uint32_t)n_dy_chunks * dy_bytes + (uint32_t)(GUARD - 1) * 128u;   // column tap ds = -1 starts one row early
        const uint64_t a_hi = desc_mn(st, a_lbo), a_lo = desc_mn(st + p.dy_chunk_bytes, a_lbo);
        const uint64_t b_hi = desc_mn(x_hi_addr, 128u), b_lo = desc_mn(x_hi_addr + p.x_plane_bytes, 128u);
        for (int ks = 0; ks < p.nks; ++ks) {
          const uint64_t adv = (uint64_t)ks * (2048u >> 4);
          const uint32_t first = (i == 0 && ks == 0) ? 0u : 1u;
          mma_bf16(d_main, a_hi + adv, b_hi + adv, idesc, first);
          mma_bf16(d_corr, a_hi + adv, b_lo + adv, idesc, first);
          mma_bf16(d_corr, a_lo + adv, b_hi + adv, idesc, 1u);
        }
        mma_commit(&empty[s]);
      }
      mma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // ================================================================================= epilogue: TMEM lane = M row, 192 columns
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    float* dst = p.partial + ((size_t)blockIdx.x * 128 + row) * NN;
    const float inv = 1.f / f16_operand_scale(p.dy_amax[0]);
    if (n_my > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < NN; c0 += 32) {
      float v[32];
      if (n_my > 0) {
        uint32_t r0[32], r1[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0;
        tmem_ld_32x32(taddr, r0);
        tmem_ld_32x32(taddr + (uint32_t)NN, r1);
        tmem_ld_wait();
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = fmaf(__uint_as_float(r1[k]), kF16LoInv, __uint_as_float(r0[k])) * inv;
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < 32; k += 4) *reinterpret_cast<float4*>(dst + c0 + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// partial [unit][split][128 rows][192 = (column tap s, channel)] -> dw OIHW. Block = one output channel x one 64-channel input chunk:
// the 576 consecutive floats dw[o][cc*64 .. +63][r][s]. threadIdx.y = split lane (splits y, y + 4, ... summed in fp64), the four lane
// sums are combined in a fixed order, so the result does not depend on scheduling.
__global__ void __launch_bounds__(192 * 4) wgrad_halo_reduce_kernel(const float* __restrict__ partial, int splits, int n_cc, int Cin, int pair,
                                                                    float* __restrict__ dw) {
  __shared__ double sh[4][3][NN];
  const int o = blockIdx.x, cc = blockIdx.y, n = threadIdx.x, ty = threadIdx.y;
  // reads follow the partials' own layout (192 consecutive floats per row tap), the (c, r, s) order of OIHW is restored through shared memory
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    int unit, m;
    if (pair) {              // pass 0: M group 0 = row tap dr = 0 (r = 1), group 1 = dr = -1 (r = 0); pass 1: group 0 = dr = +1 (r = 2)
      unit = cc * 2 + (r == 2 ? 1 : 0);
      m = o + (r == 0 ? 64 : 0);
    } else {
      unit = ((o >> 7) * n_cc + cc) * 3 + r;
      m = o & 127;
    }
    const float* src = partial + (((size_t)unit * splits) * 128 + m) * NN + n;
    double acc = 0.0;
#pragma unroll 4
    for (int sp = ty; sp < splits; sp += 4) acc += (double)src[(size_t)sp * 128 * NN];
    sh[ty][r][n] = acc;
  }
  __syncthreads();
  const int j = ty * 192 + n;
  const int c_here = min(64, Cin - cc * 64);       // channels of this chunk that exist (32 for the 32-channel layers)
  if (j < 9 * c_here) {
    const int c = j / 9, tap = j - 9 * c, r = tap / 3, sc = tap - 3 * r;
    const int nn = sc * 64 + c;
    dw[((size_t)o * Cin + cc * 64) * 9 + j] = (float)(((sh[0][r][nn] + sh[1][r][nn]) + sh[2][r][nn]) + sh[3][r][nn]);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}
static inline int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e == nullptr ? dflt : atoi(e);
}

struct Plan {
  int RB, NB, n_box, box_pos, tile_pos, n_tiles, nks, pair, n_units, splits, n_cc, n_ot;
  uint32_t dy_chunk_bytes, x_plane_bytes, stage_bytes, tx_bytes;
  size_t smem, partial_bytes;
};

static void size_plan(const PcConvGeom* g, Plan& pl) {
  const int Wp = g->W + 1;
  pl.box_pos = pl.NB * pl.RB * Wp;
  pl.tile_pos = pl.n_box * pl.box_pos;
  pl.nks = (pl.tile_pos + 15) / 16;
  // dy region rows: what TMA writes (+ one padded row in pair mode) or the rows the k-steps read (+ the row M group 1 reads ahead)
  const int dy_rows_tma = pl.pair ? pl.tile_pos + Wp : pl.tile_pos;
  const int dy_rows = 16 * pl.nks + (pl.pair ? Wp : 0);
  pl.dy_chunk_bytes = ((uint32_t)(dy_rows > dy_rows_tma ? dy_rows : dy_rows_tma) * 128u + 1023u) & ~1023u;
  const int x_rows = GUARD + 16 * pl.nks + 2;        // guard | rows the k-steps read through the three column taps
  pl.x_plane_bytes = ((uint32_t)x_rows * 128u + 1023u) & ~1023u;
  const int n_dy_chunks = pl.pair ? 1 : 2;
  pl.stage_bytes = (uint32_t)n_dy_chunks * 2u * pl.dy_chunk_bytes + 2u * pl.x_plane_bytes;
  pl.tx_bytes = (uint32_t)pl.n_box * ((uint32_t)n_dy_chunks * 2u * (uint32_t)(pl.pair ? pl.box_pos + Wp : pl.box_pos) * 128u + 2u * (uint32_t)pl.box_pos * 128u);
  pl.smem = (size_t)NSTAGE * pl.stage_bytes + sizeof(uint64_t) * (2 * NSTAGE + 1) + 16 + 1024;
}

static bool make_plan(const PcConvGeom* g, Plan& pl) {
  const int H = g->H, Wp = g->W + 1;
  if (Wp > 256) return false;
  pl.pair = g->Cout <= 64 ? 1 : 0;          // (32 output channels: the dy boxes zero-fill channels 32..63, the reduce only reads rows o < Cout)
  // A tile = n_box TMA boxes of NB images x RB padded rows; RB divides H + 1 (a box never straddles two images), several images per
  // box only for whole-image boxes (and not in pair mode, whose extra row per box assumes one image). Among the shapes that fit the
  // shared memory pick the one that wastes the fewest k-step rows (tile positions are rounded up to 16), then the largest tile.
  bool found = false;
  double best_waste = 0.0;
  Plan cand = pl;
  for (int RB = 1; RB <= H + 1; ++RB) {
    if ((H + 1) % RB != 0 || RB * Wp > 240) continue;
    for (int NB = 1; NB <= (RB == H + 1 && !pl.pair ? 16 : 1); ++NB) {
      if (g->B % NB != 0 || NB * RB * Wp > 240 || NB > 256) continue;
      for (int n_box = 1; n_box <= (pl.pair ? 1 : MAX_BOX); ++n_box) {
        cand.RB = RB; cand.NB = NB; cand.n_box = n_box;
        size_plan(g, cand);
        if (cand.tile_pos > 240 || cand.tile_pos < 64 || cand.smem > 227 * 1024) continue;
        const double waste = (double)(16 * cand.nks) / cand.tile_pos + 0.002 * (NB > 1 ? 0 : n_box);   // fewer TMA instructions on ties
        if (!found || waste < best_waste - 1e-9 || (waste < best_waste + 1e-9 && cand.tile_pos > pl.tile_pos)) {
          found = true; best_waste = waste;
          pl.RB = RB; pl.NB = NB; pl.n_box = n_box;
          size_plan(g, pl);
        }
      }
    }
  }
  if (!found) return false;
  const long long boxes_total = pl.NB > 1 ? (long long)(g->B / pl.NB) : (long long)g->B * ((H + 1) / pl.RB);
  pl.n_tiles = (int)((boxes_total + pl.n_box - 1) / pl.n_box);
  pl.n_cc = (g->Cin + 63) / 64;            // (32 input channels: one chunk whose upper half the x boxes zero-fill)
  pl.n_ot = pl.pair ? 1 : g->Cout / 128;
  pl.n_units = pl.pair ? 2 * pl.n_cc : pl.n_ot * pl.n_cc * 3;
  // splits: fill the 148 SMs in (nearly) whole waves with at least ~4 tiles per CTA
  int best = 1;
  double best_eff = 0.0;
  for (int sp = 1; sp <= 148; ++sp) {
    if (sp > 1 && pl.n_tiles / sp < 4) break;
    const int ctas = pl.n_units * sp;
    const int waves = (ctas + kNumSMs - 1) / kNumSMs;
    if (waves > 3) break;
    const double eff = (double)ctas / ((double)waves * kNumSMs);
    if (eff > best_eff + 0.03) { best_eff = eff; best = sp; }
  }
  pl.splits = best;
  {   // experiment knob: more, shorter CTAs (PC_WGRAD_HALO_SPLIT_MUL = 2, 3, ...) so that a high-priority main stream gets SMs back sooner.
      // Measured on the cnn_deep step: x1 3.00 ms, x3 3.23 ms, x6 3.49 ms (with or without the priority): every CTA zero-fills its
      // stages, allocates TMEM and writes a 98 KB partial, so shorter CTAs cost more than the scheduling freedom returns.
    const int mul = env_int("PC_WGRAD_HALO_SPLIT_MUL", 1);
    if (mul > 1) {
      int sp = best * mul;
      if (sp > pl.n_tiles) sp = pl.n_tiles;
      if (sp >= 1) pl.splits = sp;
    }
  }
  pl.partial_bytes = (size_t)pl.n_units * pl.splits * 128 * NN * sizeof(float);
  return true;
}

}  // namespace halowg
}  // namespace pc

using namespace pc;

// 1 when the halo weight-gradient engine covers this convolution: stride-1 3x3 pad-1, FP16X2 planes on both operands, input channels
// a multiple of 64 (or 32), output channels 64 (or 32) or a multiple of 128. PC_WGRAD_HALO=0 disables it.
extern "C" int pc_conv_wgrad_halo_supported(const PcConvGeom* g) {
  if (g == nullptr || !pc::halowg::env_int("PC_WGRAD_HALO", 1)) return 0;
  if (g->R != 3 || g->S != 3 || g->stride != 1 || g->pad != 1 || g->Ho != g->H || g->Wo != g->W) return 0;
  if (!(g->Cin % 64 == 0 || g->Cin == 32) || !(g->Cout == 64 || g->Cout == 32 || g->Cout % 128 == 0)) return 0;
  const long long Q = (long long)g->B * (g->H + 1) * (g->W + 1);
  const int cmax = g->Cin > g->Cout ? g->Cin : g->Cout;
  if (Q + 4096 >= (1LL << 31) || (long long)g->B * g->H * g->W * cmax >= (1LL << 31)) return 0;
  pc::halowg::Plan pl;
  return pc::halowg::make_plan(g, pl) ? 1 : 0;
}

extern "C" size_t pc_conv_wgrad_halo_workspace(const PcConvGeom* g) {
  pc::halowg::Plan pl;
  if (g == nullptr || !pc::halowg::make_plan(g, pl)) return 0;
  return pl.partial_bytes;
}

// x_planes / dy_planes: fp16 hi | lo planes ([2][B][H][W][C], pc_bn_act_split layout), dy scaled by f16_operand_scale(*dy_amax).
extern "C" int pc_conv_wgrad_halo(const void* x_planes, const void* dy_planes, const PcConvGeom* g, float* dw_oihw, void* workspace,
                                  size_t workspace_bytes, const float* dy_amax, pc_stream_t stream) {
  using namespace pc::halowg;
  PC_REQUIRE(x_planes && dy_planes && g && dw_oihw && workspace && dy_amax, PC_EINVAL, "pc_conv_wgrad_halo: null pointer");
  if (!pc_conv_wgrad_halo_supported(g)) return PC_EUNSUPPORTED;
  Plan pl;
  PC_REQUIRE(make_plan(g, pl), PC_EUNSUPPORTED, "pc_conv_wgrad_halo: no plan for this shape");
  PC_REQUIRE(workspace_bytes >= pl.partial_bytes, PC_EINVAL, "pc_conv_wgrad_halo: workspace too small (%zu < %zu)", workspace_bytes, pl.partial_bytes);
  EncodeTiledFn enc = encode_tiled();
  PC_REQUIRE(enc != nullptr, PC_ECUDA, "pc_conv_wgrad_halo: cuTensorMapEncodeTiled is not available from this driver");
  Params p{};
  p.partial = static_cast<float*>(workspace); p.dy_amax = dy_amax;
  p.B = g->B; p.H = g->H; p.W = g->W; p.Cin = g->Cin; p.Cout = g->Cout;
  p.Wp = g->W + 1; p.RB = pl.RB; p.NB = pl.NB; p.n_box = pl.n_box; p.box_pos = pl.box_pos; p.tile_pos = pl.tile_pos; p.n_tiles = pl.n_tiles;
  p.pair = pl.pair; p.n_units = pl.n_units; p.splits = pl.splits; p.n_cc = pl.n_cc; p.n_ot = pl.n_ot; p.nks = pl.nks;
  p.dy_chunk_bytes = pl.dy_chunk_bytes; p.x_plane_bytes = pl.x_plane_bytes; p.stage_bytes = pl.stage_bytes; p.tx_bytes = pl.tx_bytes;

  auto make_map = [&](CUtensorMap* m, const void* base, int C, int rows) -> CUresult {
    const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)g->W, (cuuint64_t)g->H, (cuuint64_t)g->B, 2};
    const cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)g->W * C * 2, (cuuint64_t)g->H * g->W * C * 2, (cuuint64_t)g->B * g->H * g->W * C * 2};
    const cuuint32_t box[5] = {64, (cuuint32_t)p.Wp, (cuuint32_t)rows, (cuuint32_t)p.NB, 1};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  };
  CUtensorMap dymap, xmap;
  CUresult cr = make_map(&dymap, dy_planes, g->Cout, pl.pair ? pl.RB + 1 : pl.RB);
  PC_REQUIRE(cr == CUDA_SUCCESS, PC_ECUDA, "pc_conv_wgrad_halo: cuTensorMapEncodeTiled(dy) failed (CUresult %d)", (int)cr);
  cr = make_map(&xmap, x_planes, g->Cin, pl.RB);
  PC_REQUIRE(cr == CUDA_SUCCESS, PC_ECUDA, "pc_conv_wgrad_halo: cuTensorMapEncodeTiled(x) failed (CUresult %d)", (int)cr);
  static size_t conf = 0;
  if (pl.smem > conf) {
    PC_CUDA(cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    conf = pl.smem;
  }
  wgrad_halo_kernel<<<pl.n_units * pl.splits, THREADS, pl.smem, stream>>>(dymap, xmap, p);
  PC_LAUNCH_CHECK("wgrad_halo_kernel");
  wgrad_halo_reduce_kernel<<<dim3(g->Cout, pl.n_cc), dim3(192, 4), 0, stream>>>(p.partial, pl.splits, pl.n_cc, g->Cin, pl.pair, dw_oihw);
  PC_LAUNCH_CHECK("wgrad_halo_reduce_kernel");
  return PC_OK;
}
