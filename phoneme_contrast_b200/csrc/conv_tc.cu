// tcgen05 tensor-core implicit-GEMM convolution for sm_100a (forward and data-gradient), plus the plain
// C = A * B^T GEMM built from the same tile engine (used by tests as a unit check of the tensor-core path).
//
//   D[128 x BN] (fp32, TMEM)  +=  A[128 x K] * B[BN x K]^T        K = taps x channels, walked in 128-byte k-chunks
//
// Warp roles (14 warps, one CTA per 128-row M tile x BN-column N tile):
//   warps 0-11 three producer groups of 4 warps (group g fills k-chunks g, g+3, ...), then all join the epilogue. The A operand is an im2col GATHER with the previous layer's
//              BatchNorm-apply + ReLU + Dropout2d folded in, so it cannot come from TMA: 8 lanes fetch one
//              pixel's 128 contiguous bytes (coalesced), transform, convert, and store them into the
//              SWIZZLE_128B K-major tile the UMMA descriptor expects; fence.proxy.async; mbarrier arrive.
//   warp 12    allocates TMEM, then one lane issues tcgen05.mma (D in TMEM) and tcgen05.commit per stage.
//   warp 13    one lane streams the pre-swizzled weight tiles with 1-D bulk async copies (cp.async.bulk, TMA
//              engine) that complete on the same "full" mbarrier as the producers' arrivals.
//   epilogue   tcgen05.ld (32 lanes x 32 columns per instruction) -> + bias -> BatchNorm sum / sum-of-squares via a
//              31-shuffle warp reduce-scatter -> NHWC fp32 store.
//
// Precision modes (include/phoneme_contrast.h PC_PREC_*):
//   TF32X3  fp32 operands split as x = hi + lo (hi = top 19 bits, lo = x - hi exactly); three kind::tf32 MMAs
//           per k-step (lo*hi, hi*lo, hi*hi) accumulate ~fp32-accurate products in the fp32 TMEM accumulator.
//           This is the mode that keeps the 1e-4 parity bar ("fp32 parity on tensor cores").
//   FP16X2  fp32 operands split as x = hi + lo * 2^-11 with hi, lo in fp16 (tc_common.cuh: split_f16x2): the same three
//           products and the same ~22-bit operand precision as TF32X3, but kind::f16 MMAs run at twice the tf32 rate
//           and every operand byte in shared memory carries twice the K extent (64 channels per 128-byte row). fp16's
//           narrow exponent is handled by a power-of-two operand scale derived from the tensor's max magnitude
//           (`a_amax`, used for the gradient operand of dgrad) and undone in the epilogue.
//   BF16    operands rounded to bf16, one kind::f16 MMA per k-step (the 1e-2 tolerance mode).
#include <stdlib.h>
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace pc {
namespace tcconv {

using namespace pc::tc;

constexpr int BM = 128;
constexpr int NPROD = 128;            // threads that fill one stage (one producer group)
// Producer groups: group g owns k-chunks g, g+NGROUPS, ... so that the global-load latency of several stages is in flight
// at once (one group alone is latency-bound). Two shapes are built:
//   NGROUPS = 3, one CTA per SM  (BN = 128: the 4 x 128 TMEM accumulator columns fill the SM's tensor memory)
//   NGROUPS = 2, two CTAs per SM (BN <= 64: 256 TMEM columns and ~100 KB smem each) -- while one CTA is in its prologue,
//   MMA drain or epilogue (~38 % of a K = 576 tile), the other's main loop keeps the shared-memory pipes busy.
constexpr int MAX_STAGES = 6;
constexpr uint32_t SMEM_BUDGET = 200 * 1024;

struct XformDev {
  const float* scale;
  const float* shift;
  const float* drop;
  int relu;
};

struct Params {
  const float* A;
  const unsigned char* Bp;
  const float* bias;
  const float* a_amax;   // FP16X2: device scalar max|A| -> power-of-two operand scale (null: scale 1)
  float* C;
  double* stats;
  XformDev xf;
  PcConvGeom g;
  int mode;          // 0 conv fwd gather, 1 conv dgrad gather, 2 plain row-major A [M][lda]
  long long M;
  int Nn, Npad, Ca, n_kc, lda, accumulate, stages;
  int aff_floats;    // 2 * Ca when the gather applies a per-channel affine (kept in shared memory), else 0
  int presplit;      // A is a pair of fp16 planes (hi, lo) instead of fp32 (FP16X2, mode 0 only)
  size_t plane_bytes;
  long long* dbg;    // optional [gridDim.x*gridDim.y][16] clock64 stamps (diagnostics, see pc_tc_set_debug)
};

#define PC_STAMP(slot)                                                                                 \
  do {                                                                                                 \
    if (p.dbg != nullptr) p.dbg[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] = clock64(); \
  } while (0)

template <int PREC> struct Prec;
// NACC: TMEM accumulators (of BN columns) per tile. The tensor core adds into its fp32 accumulator with truncation, so a
// long dependent chain of accumulations picks up a bias of ~2^-24 per step. TF32X3 therefore keeps the small correction
// terms apart from the hi*hi terms and alternates between two accumulator sets [main | corr] per k-step; the epilogue adds
// the four partial sums in registers with round-to-nearest (measured: ~10x lower error than a single accumulator).
template <> struct Prec<PC_PREC_TF32X3> { static constexpr int BKC = 32, PARTS = 2, NACC = 4; };
template <> struct Prec<PC_PREC_BF16> { static constexpr int BKC = 64, PARTS = 1, NACC = 1; };
template <> struct Prec<PC_PREC_FP16X2> { static constexpr int BKC = 64, PARTS = 2, NACC = 4; };

template <int BN, int PREC>
__host__ __device__ constexpr uint32_t stage_bytes() { return (uint32_t)Prec<PREC>::PARTS * (BM * 128 + BN * 128); }

// v[32] per lane (lane = row) -> returns sum over the 32 lanes of column `lane`
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

// PRESPLIT: the gathered tensor is already stored as fp16 hi | lo planes (pc_bn_act_split): the producers only copy
// 16-byte chunks into the swizzled tile (no BatchNorm / ReLU / dropout / split arithmetic, which the plain gather repeats
// for every tap and every N tile that touches an input element).
template <int BN, int PREC, int NGROUPS, int MINB, bool PRESPLIT>
__global__ void __launch_bounds__(32 * (4 * NGROUPS + 2), MINB) igemm_tc_kernel(const Params p) {
  constexpr int PROD_WARPS = 4 * NGROUPS;
  using P = Prec<PREC>;
  constexpr int BKC = P::BKC, PARTS = P::PARTS, NACC = P::NACC;
  constexpr uint32_t A_PART = BM * 128, B_PART = BN * 128;
  constexpr uint32_t STAGE = stage_bytes<BN, PREC>();

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // SWIZZLE_128B tiles need a 1024-byte aligned base: align by hand (the launch reserves 1 KB of slack)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int S = p.stages;
  unsigned char* tiles = smem;                                                    // [S][A parts | B parts]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * STAGE);        // [S]
  uint64_t* empty = full + MAX_STAGES;                                            // [S]
  uint64_t* acc_full = empty + MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);                        // [BN]
  float* s_sum = s_bias + BN;                                                     // [BN]
  float* s_sq = s_sum + BN;                                                       // [BN]
  int* s_off0 = reinterpret_cast<int*>(s_sq + BN);                                // [BM] per-row window-origin offset
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_off0 + BM);                    // [BM] per-row valid-tap bitmask
  int* s_pix = reinterpret_cast<int*>(s_mask + BM);                               // [BM] output row (pixel) index, -1 = none
  int* s_smp = s_pix + BM;                                                        // [BM] sample (batch) index of the row
  int* s_nact = s_smp + BM;                                                       // [1] number of active k-chunks
  float* s_aff = reinterpret_cast<float*>(s_nact + 4);                            // [2][Ca] fused BatchNorm scale | shift (aff_floats)
  int4* s_kc = reinterpret_cast<int4*>(s_aff + p.aff_floats);                     // [n_kc] active k-chunks: {kc, tap, c_base, tap_off}

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) PC_STAMP(0);
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  if (tid < BN) {
    const int n = n0 + tid;
    (void)n;
    s_sum[tid] = 0.f;
    s_sq[tid] = 0.f;
  }
  // ---- row_info: one thread per tile row computes the element offset of the row's window origin in the gathered
  // tensor and the bitmask of taps that fall inside it; per k-chunk a row then costs one bit test and one add.
  //   fwd   : pixel (b,ho,wo) reads x[b, ho*stride - pad + r, wo*stride - pad + s, :]          -> off0 + (r*Wa + s)*Ca
  //   dgrad : pixel (b,h,w) reads dy[b, (h+pad-r)/stride, (w+pad-s)/stride, :] when divisible  -> off0 - ((r/stride)*Wa + s/stride)*Ca
  //           (for a valid tap (h+pad-r)/stride == floor((h+pad)/stride) - floor(r/stride))
  if (tid < BM) {
    const PcConvGeom g = p.g;
    const int Hr = p.mode == 0 ? g.Ho : g.H, Wr = p.mode == 0 ? g.Wo : g.W;
    const int Ha = p.mode == 0 ? g.H : g.Ho, Wa = p.mode == 0 ? g.W : g.Wo;
    const int st = g.stride;
    const int m = (int)m0 + tid;
    int off = 0, pix = -1;
    uint32_t mask = 0u;
    if (m < (int)p.M) {
      if (p.mode == 2) {
        off = m * p.lda;
        mask = 1u;
        pix = m;
      } else {
        int wq, hq, bq;
        if (p.mode == 1 && st == 2) {
          // stride-2 data gradient: enumerate input pixels parity-class-major ((h&1, w&1) = (0,0),(0,1),(1,0),(1,1)) so that the
          // rows of a tile share their valid taps (a pixel of one class receives from only 1, 2, 2 or 4 of the 9 taps) and the
          // k-chunks of taps nobody in the tile needs are skipped altogether.
          int loc = m, c = 0, nh = 0, nw = 0;
          for (; c < 4; ++c) {
            nh = (Hr - (c >> 1) + 1) >> 1;
            nw = (Wr - (c & 1) + 1) >> 1;
            const int cnt = g.B * nh * nw;
            if (loc < cnt) break;
            loc -= cnt;
          }
          const int w2 = loc % nw, t = loc / nw;
          wq = 2 * w2 + (c & 1);
          hq = 2 * (t % nh) + (c >> 1);
          bq = t / nh;
        } else {
          wq = m % Wr;
          const int t = m / Wr;
          hq = t % Hr;
          bq = t / Hr;
        }
        pix = (bq * Hr + hq) * Wr + wq;
        int hb, wb;
        if (p.mode == 0) {
          hb = hq * st - g.pad;
          wb = wq * st - g.pad;
          uint32_t wmask = 0u;
          for (int ts = 0; ts < g.S; ++ts)
            if ((unsigned)(wb + ts) < (unsigned)Wa) wmask |= 1u << ts;
          for (int tr = 0; tr < g.R; ++tr)
            if ((unsigned)(hb + tr) < (unsigned)Ha) mask |= wmask << (tr * g.S);
        } else {
          const int hp = hq + g.pad, wp = wq + g.pad;
          hb = hp / st;
          wb = wp / st;
          const int hpar = hp - hb * st, wpar = wp - wb * st;
          uint32_t wmask = 0u;
          for (int ts = 0; ts < g.S; ++ts) {
            const int q = ts / st;
            if (ts - q * st == wpar && (unsigned)(wb - q) < (unsigned)Wa) wmask |= 1u << ts;
          }
          for (int tr = 0; tr < g.R; ++tr) {
            const int q = tr / st;
            if (tr - q * st == hpar && (unsigned)(hb - q) < (unsigned)Ha) mask |= wmask << (tr * g.S);
          }
        }
        off = ((bq * Ha + hb) * Wa + wb) * p.Ca;
      }
    }
    s_pix[tid] = pix;
    s_smp[tid] = pix >= 0 ? pix / (Hr * Wr) : 0;
    s_off0[tid] = off;
    s_mask[tid] = mask;
    asm volatile("bar.sync 2, 128;" ::: "memory");
    if (warp == 0) {
      uint32_t tm = s_mask[lane] | s_mask[lane + 32] | s_mask[lane + 64] | s_mask[lane + 96];
      tm = __reduce_or_sync(0xffffffffu, tm);                      // taps that at least one row of this tile needs
      const int cpt_ = p.Ca / BKC;
      int n_act = 0;
      for (int base = 0; base < p.n_kc; base += 32) {
        const int kc = base + lane;
        const bool act = kc < p.n_kc && ((tm >> (kc / cpt_)) & 1u);
        const uint32_t bal = __ballot_sync(0xffffffffu, act);
        if (act) {
          // per-stage gather constants, computed once here instead of once per producer thread and stage
          const int tap = kc / cpt_, tr = tap / p.g.S, ts = tap - tr * p.g.S;
          const int Wa_ = p.mode == 0 ? p.g.W : p.g.Wo;
          int tap_off = 0;
          if (p.mode == 0) tap_off = (tr * Wa_ + ts) * p.Ca;
          else if (p.mode == 1) tap_off = -((tr / p.g.stride) * Wa_ + ts / p.g.stride) * p.Ca;
          s_kc[n_act + __popc(bal & ((1u << lane) - 1u))] = make_int4(kc, tap, (kc - tap * cpt_) * BKC, tap_off);
        }
        n_act += __popc(bal);
      }
      if (lane == 0) s_nact[0] = n_act;
    }
  }
  if (warp == PROD_WARPS) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(&full[s], NPROD + 1);
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
  pdl_trigger();   // TMEM is allocated: the next kernel on the stream may start its prologue behind us
  pdl_wait();      // everything above used kernel parameters only; from here on we read the predecessor's output
  if (tid < BN) {  // read by the epilogue warps after their "bar.sync 3" below
    const int n = n0 + tid;
    s_bias[tid] = (p.bias != nullptr && n < p.Nn) ? p.bias[n] : 0.f;
  }
  if (p.aff_floats != 0 && warp < PROD_WARPS) {
    for (int c = tid; c < p.Ca; c += 32 * PROD_WARPS) {
      s_aff[c] = p.xf.scale[c];
      s_aff[p.Ca + c] = p.xf.shift[c];
    }
    asm volatile("bar.sync 4, %0;" ::"n"(32 * PROD_WARPS) : "memory");
  }
  if (tid == 0) PC_STAMP(1);

  if (warp < PROD_WARPS) {
    // ============================================================ producers
    const int group = warp >> 2;            // producer group
    const int gt = tid & (NPROD - 1);       // thread index inside the group
    const int j = gt & 7;           // 16-byte chunk of the 128-byte k-chunk row
    const int rg = gt >> 3;         // rows rg + 16*i
    // per-row window-origin offsets and valid-tap masks were computed once per CTA (see row_info above)
    const PcConvGeom g = p.g;
    const int Ha = p.mode == 0 ? g.H : g.Ho, Wa = p.mode == 0 ? g.W : g.Wo;
    int off0[8];
    uint32_t tapmask[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      off0[i] = s_off0[rg + 16 * i];
      tapmask[i] = s_mask[rg + 16 * i];
    }
    const int cpt = p.Ca / BKC;     // k-chunks per tap
    constexpr int EPC = (PREC == PC_PREC_TF32X3) ? 4 : 8; // source elements per 16-byte destination chunk
    const bool use_scale = (PREC == PC_PREC_FP16X2 && p.a_amax != nullptr);
    const float a_scale = use_scale ? f16_operand_scale(p.a_amax[0]) : 1.f;
    const bool has_aff = (p.mode == 0 && p.xf.scale != nullptr);
    const bool has_relu = (p.mode == 0 && p.xf.relu);
    const bool has_drop = (p.mode == 0 && p.xf.drop != nullptr);
    const uint32_t soff0 = sw128_offset((uint32_t)rg, (uint32_t)j);     // rows rg + 16*i share (row & 7): + 2048*i
    if (tid == 0) PC_STAMP(2);
    const int n_act = s_nact[0];
    // (A software pipeline inside a group -- prefetching the next stage's rows into registers freed by the conversion -- was
    //  measured: no gain on the BN = 128 kernels and slower on the BN = 64 ones; the plain order below is kept.)
    float v[8][EPC];
    int tap = 0, c0 = 0, tap_off = 0;
    const float* src_base = p.A;
    auto setup = [&](int it_) {
      const int4 st = s_kc[it_];
      tap = st.y;
      c0 = st.z + j * EPC;                        // first channel of my chunk
      tap_off = st.w;
      src_base = p.A + tap_off + c0;
    };
    const unsigned char* planes = reinterpret_cast<const unsigned char*>(p.A);     // PRESPLIT: hi plane, lo plane at + plane_bytes
    auto load_half = [&](auto half) {
      constexpr int H = decltype(half)::value;
      if (PRESPLIT) {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int i = 4 * H + ii;
          const bool ok = (tapmask[i] >> tap) & 1u;
          uint4 h = make_uint4(0u, 0u, 0u, 0u), l = h;
          if (ok) {
            const size_t e = (size_t)(off0[i] + tap_off + c0) * 2;       // element index -> byte offset in an fp16 plane
            h = *reinterpret_cast<const uint4*>(planes + e);
            l = *reinterpret_cast<const uint4*>(planes + p.plane_bytes + e);
          }
          v[i][0] = __uint_as_float(h.x); v[i][1] = __uint_as_float(h.y); v[i][2] = __uint_as_float(h.z); v[i][3] = __uint_as_float(h.w);
          v[i][4 % EPC] = __uint_as_float(l.x); v[i][5 % EPC] = __uint_as_float(l.y); v[i][6 % EPC] = __uint_as_float(l.z); v[i][7 % EPC] = __uint_as_float(l.w);
        }
        return;
      }
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        const int i = 4 * H + ii;
        const bool ok = (tapmask[i] >> tap) & 1u;
#pragma unroll
        for (int q = 0; q < EPC; q += 4) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) t = *reinterpret_cast<const float4*>(src_base + off0[i] + q);
          v[i][q] = t.x; v[i][q + 1] = t.y; v[i][q + 2] = t.z; v[i][q + 3] = t.w;
        }
      }
    };
    using Half0 = std::integral_constant<int, 0>;
    using Half1 = std::integral_constant<int, 1>;
    for (int it = group; it < n_act; it += NGROUPS) {
      const int s = it % S;
      const uint32_t ph = (uint32_t)(it / S) & 1u;
      // this stage's loads are issued here, ahead of the wait for the smem slot
      setup(it);
      load_half(Half0{});
      load_half(Half1{});
      const int cur_tap = tap, cur_c0 = c0;
      float sc[EPC], sh[EPC];
      if (has_aff) {
#pragma unroll
        for (int q = 0; q < EPC; q += 4) {
          const float4 a = *reinterpret_cast<const float4*>(s_aff + cur_c0 + q);
          const float4 b = *reinterpret_cast<const float4*>(s_aff + p.Ca + cur_c0 + q);
          sc[q] = a.x; sc[q + 1] = a.y; sc[q + 2] = a.z; sc[q + 3] = a.w;
          sh[q] = b.x; sh[q + 1] = b.y; sh[q + 2] = b.z; sh[q + 3] = b.w;
        }
      }
      mbar_wait(&empty[s], ph ^ 1u);
      unsigned char* a_hi = tiles + (size_t)s * STAGE + soff0;
      unsigned char* a_lo = a_hi + A_PART;
      auto convert_half = [&](auto half) {
        constexpr int H = decltype(half)::value;
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int i = 4 * H + ii;
          if (PRESPLIT) {       // bits already in operand form: copy
            *reinterpret_cast<uint4*>(a_hi + 2048 * i) = make_uint4(__float_as_uint(v[i][0]), __float_as_uint(v[i][1]), __float_as_uint(v[i][2]), __float_as_uint(v[i][3]));
            *reinterpret_cast<uint4*>(a_lo + 2048 * i) = make_uint4(__float_as_uint(v[i][4 % EPC]), __float_as_uint(v[i][5 % EPC]), __float_as_uint(v[i][6 % EPC]), __float_as_uint(v[i][7 % EPC]));
            continue;
          }
          const bool ok = (tapmask[i] >> cur_tap) & 1u;
          if (ok) {
            if (has_aff) {
#pragma unroll
              for (int q = 0; q < EPC; ++q) v[i][q] = fmaf(v[i][q], sc[q], sh[q]);
            }
            if (has_relu) {
#pragma unroll
              for (int q = 0; q < EPC; ++q) v[i][q] = fmaxf(v[i][q], 0.f);
            }
            if (has_drop) {
              const int b = s_smp[rg + 16 * i];                          // a valid tap reads the output pixel's own sample
#pragma unroll
              for (int q = 0; q < EPC; q += 4) {
                const float4 d = *reinterpret_cast<const float4*>(p.xf.drop + (size_t)b * p.Ca + cur_c0 + q);
                v[i][q] *= d.x; v[i][q + 1] *= d.y; v[i][q + 2] *= d.z; v[i][q + 3] *= d.w;
              }
            }
          }
          if (PREC == PC_PREC_TF32X3) {
            float h[4], l[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) split_tf32(v[i][q], h[q], l[q]);
            *reinterpret_cast<float4*>(a_hi + 2048 * i) = make_float4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<float4*>(a_lo + 2048 * i) = make_float4(l[0], l[1], l[2], l[3]);
          } else if (PREC == PC_PREC_FP16X2) {
            uint4 h, l;
            if (use_scale) {
#pragma unroll
              for (int q = 0; q < EPC; ++q) v[i][q] *= a_scale;
            }
            split_f16x2(v[i][0], v[i][1], h.x, l.x);
            split_f16x2(v[i][2], v[i][3], h.y, l.y);
            split_f16x2(v[i][4 % EPC], v[i][5 % EPC], h.z, l.z);
            split_f16x2(v[i][6 % EPC], v[i][7 % EPC], h.w, l.w);
            *reinterpret_cast<uint4*>(a_hi + 2048 * i) = h;
            *reinterpret_cast<uint4*>(a_lo + 2048 * i) = l;
          } else {
            uint4 w;
            w.x = pack_bf16(v[i][0], v[i][1]); w.y = pack_bf16(v[i][2], v[i][3]);
            w.z = pack_bf16(v[i][4 % EPC], v[i][5 % EPC]); w.w = pack_bf16(v[i][6 % EPC], v[i][7 % EPC]);
            *reinterpret_cast<uint4*>(a_hi + 2048 * i) = w;
          }
        }
      };
      convert_half(Half0{});
      convert_half(Half1{});
      fence_proxy_async();
      mbar_arrive(&full[s]);
      if (tid == 0 && it == 0) PC_STAMP(3);
      if (tid == 0 && it == 3) PC_STAMP(4);
    }
    if (tid == 0) PC_STAMP(5);

    // ============================================================ epilogue
    // warp w reads TMEM lanes 32*(w%4).. (its hardware lane quarter) = tile rows, and the 32-column chunks w/4, w/4+NGROUPS, ..
    asm volatile("bar.sync 3, %0;" ::"n"(32 * PROD_WARPS) : "memory");     // s_bias visible to all epilogue warps
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (tid == 0) PC_STAMP(6);
    const float out_scale = 1.f / a_scale;      // exact: a_scale is a power of two
    const int r = (warp & 3) * 32 + lane;
    const int pix = s_pix[r];
    const bool valid = pix >= 0;
    float* dst_row = p.C + (size_t)(valid ? pix : 0) * p.Nn + n0;
#pragma unroll 1
    for (int c0 = 32 * group; c0 < BN; c0 += 32 * NGROUPS) {
      float v[32];
      if (n_act == 0) {                      // no tap reaches this tile (e.g. odd pixels of a 1x1 stride-2 conv): zeros
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = 0.f;
      } else {
        uint32_t raw[32];
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0;
        tmem_ld_32x32(taddr, raw);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = __uint_as_float(raw[q]);
        if (NACC == 4) {
          // accumulator sets [main0 | corr0 | main1 | corr1]: v = (main0 + main1) + corr_scale * (corr0 + corr1)
          float u[32];
          tmem_ld_32x32(taddr + 2 * BN, raw);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] += __uint_as_float(raw[q]);
          tmem_ld_32x32(taddr + BN, raw);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) u[q] = __uint_as_float(raw[q]);
          tmem_ld_32x32(taddr + 3 * BN, raw);
          tmem_ld_wait();
          constexpr float corr_scale = (PREC == PC_PREC_FP16X2) ? kF16LoInv : 1.f;
#pragma unroll
          for (int q = 0; q < 32; ++q) v[q] = fmaf(u[q] + __uint_as_float(raw[q]), corr_scale, v[q]);
          if (PREC == PC_PREC_FP16X2 && p.a_amax != nullptr) {
#pragma unroll
            for (int q = 0; q < 32; ++q) v[q] *= out_scale;
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] = valid ? v[q] + s_bias[c0 + q] : 0.f;
      if (valid) {
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          if (n0 + c0 + q < p.Nn) {
            float4 o = make_float4(v[q], v[q + 1], v[q + 2], v[q + 3]);
            float* d = dst_row + c0 + q;
            if (p.accumulate) {
              const float4 old = *reinterpret_cast<const float4*>(d);
              o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            *reinterpret_cast<float4*>(d) = o;
          }
        }
      }
      if (p.stats != nullptr) {
        float sq[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) sq[q] = v[q] * v[q];
        const float cs = warp_reduce_scatter32(v, lane);
        const float cq = warp_reduce_scatter32(sq, lane);
        atomicAdd(&s_sum[c0 + lane], cs);
        atomicAdd(&s_sq[c0 + lane], cq);
      }
    }
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, %0;" ::"n"(32 * PROD_WARPS) : "memory");     // producer/epilogue warps only
      if (tid < BN) {
        const int n = n0 + tid;
        if (n < p.Nn) {
          atomicAdd(p.stats + n, (double)s_sum[tid]);
          atomicAdd(p.stats + p.Nn + n, (double)s_sq[tid]);
        }
      }
    }
  } else if (warp == PROD_WARPS) {
    // ============================================================ MMA issuer
    if (lane == 0) {
      constexpr uint32_t FMT = PREC == PC_PREC_TF32X3 ? 2u : (PREC == PC_PREC_BF16 ? 1u : 0u);
      const uint32_t idesc = instr_desc(FMT, BM, BN);
      const uint32_t idesc2 = instr_desc(FMT, BM, 2 * BN);
      const int n_act = s_nact[0];
      for (int it = 0; it < n_act; ++it) {
        const int s = it % S;
        const uint32_t ph = (uint32_t)(it / S) & 1u;
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (it == 0) PC_STAMP(8);
        if (it == 1) PC_STAMP(9);
        if (it == 4) PC_STAMP(10);
        const uint32_t base = smem_u32(tiles + (size_t)s * STAGE);
        const uint64_t a_hi = smem_desc_sw128(base);
        const uint64_t a_lo = smem_desc_sw128(base + A_PART);
        const uint64_t b_hi = smem_desc_sw128(base + PARTS * A_PART);
        const uint64_t b_lo = smem_desc_sw128(base + PARTS * A_PART + B_PART);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t adv = (uint64_t)(kk * 2);       // 32 bytes per k-step, in 16-byte units
          const int ks = it * 4 + kk;                    // global k-step
          if (PREC == PC_PREC_TF32X3) {
            // b_lo is stored right behind b_hi, so ONE MMA with N = 2*BN forms a_hi*b_hi (columns [0,BN): main) and
            // a_hi*b_lo (columns [BN,2BN): correction) while reading a_hi from shared memory once; the second MMA adds
            // a_lo*b_hi into the correction half. Two accumulator sets alternate per k-step (see NACC note above).
            const uint32_t d_set = tmem_base + (uint32_t)((ks & 1) * 2 * BN);
            mma_tf32(d_set, a_hi + adv, b_hi + adv, idesc2, ks < 2 ? 0u : 1u);
            mma_tf32(d_set + BN, a_lo + adv, b_hi + adv, idesc, 1u);
          } else if (PREC == PC_PREC_FP16X2) {
            // same scheme with kind::f16 (K = 16 per instruction, still 32 bytes per k-step); the correction half holds
            // 2^11 x its value (split_f16x2) and is rescaled in the epilogue
            const uint32_t d_set = tmem_base + (uint32_t)((ks & 1) * 2 * BN);
            mma_bf16(d_set, a_hi + adv, b_hi + adv, idesc2, ks < 2 ? 0u : 1u);
            mma_bf16(d_set + BN, a_lo + adv, b_hi + adv, idesc, 1u);
          } else {
            mma_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc, ks == 0 ? 0u : 1u);
          }
        }
        mma_commit(&empty[s]);
      }
      mma_commit(acc_full);
      PC_STAMP(11);
    }
    __syncwarp();
  } else {
    // ============================================================ weight loader
    if (lane == 0) {
      const size_t kc_stride = (size_t)PARTS * p.Npad * 128;
      const int n_act = s_nact[0];
      for (int it = 0; it < n_act; ++it) {
        const int kc = s_kc[it].x;
        const int s = it % S;
        const uint32_t ph = (uint32_t)(it / S) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        unsigned char* b_dst = tiles + (size_t)s * STAGE + PARTS * A_PART;
        const unsigned char* src = p.Bp + (size_t)kc * kc_stride + (size_t)n0 * 128;
        mbar_arrive_expect_tx(&full[s], PARTS * B_PART);
#pragma unroll
        for (int part = 0; part < PARTS; ++part)
          bulk_g2s(b_dst + part * B_PART, src + (size_t)part * p.Npad * 128, B_PART, &full[s]);
      }
    }
    __syncwarp();
  }

  if (tid == 0) PC_STAMP(7);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == PROD_WARPS) tmem_dealloc(tmem_base, NACC * BN);
  if (tid == 0) PC_STAMP(12);
}

// ------------------------------------------------------------------------------------------------ weight / B packing
// out layout: [n_kc][PARTS][Npad rows][128 bytes, 16-byte chunk index XOR (row & 7)]
// src_mode 0: conv fwd   B[n = o][k = (r,s,c)] = w[o][c][r][s]      (Ca = I)
// src_mode 1: conv dgrad B[n = c][k = (r,s,o)] = w[o][c][r][s]      (Ca = O)
// src_mode 2: plain      B[n][k] = src[n * ld + k]                  (taps = 1, Ca = K)
template <int PREC>
__device__ __forceinline__ void pack_b_item(const float* __restrict__ src, int O, int I, int R, int S, int src_mode, int ld, int Nn,
                                            int Npad, int Ca, unsigned char* __restrict__ out, long long idx) {
  using P = Prec<PREC>;
  constexpr int BKC = P::BKC, PARTS = P::PARTS;
  constexpr int EPC = (PREC == PC_PREC_TF32X3) ? 4 : 8;
  (void)O;
  const int cpt = (Ca + BKC - 1) / BKC;      // the last chunk of a 32-channel layer on the FP16X2 halo engine is zero-padded to 64
  const int j = (int)(idx & 7);
  const int n = (int)((idx >> 3) % Npad);
  const int kc = (int)((idx >> 3) / Npad);
  const int tap = kc / cpt, c0 = (kc % cpt) * BKC + j * EPC;
  float v[EPC];
#pragma unroll
  for (int q = 0; q < EPC; ++q) {
    float x = 0.f;
    const int c = c0 + q;
    if (n < Nn && c < Ca) {
      if (src_mode == 0) x = src[(((size_t)n * I + c) * R + tap / S) * S + tap % S];
      else if (src_mode == 1) x = src[(((size_t)c * I + n) * R + tap / S) * S + tap % S];
      else x = src[(size_t)n * ld + c];
    }
    v[q] = x;
  }
  unsigned char* base = out + ((size_t)kc * PARTS * Npad + n) * 128 + (size_t)((j ^ (n & 7)) << 4);
  if (PREC == PC_PREC_TF32X3) {
    float h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split_tf32(v[q], h[q], l[q]);
    *reinterpret_cast<float4*>(base) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(base + (size_t)Npad * 128) = make_float4(l[0], l[1], l[2], l[3]);
  } else if (PREC == PC_PREC_FP16X2) {
    uint4 h, l;
    split_f16x2(v[0], v[1], h.x, l.x);
    split_f16x2(v[2], v[3], h.y, l.y);
    split_f16x2(v[4 % EPC], v[5 % EPC], h.z, l.z);
    split_f16x2(v[6 % EPC], v[7 % EPC], h.w, l.w);
    *reinterpret_cast<uint4*>(base) = h;
    *reinterpret_cast<uint4*>(base + (size_t)Npad * 128) = l;
  } else {
    uint4 w;
    w.x = pack_bf16(v[0], v[1]); w.y = pack_bf16(v[2], v[3]);
    w.z = pack_bf16(v[4 % EPC], v[5 % EPC]); w.w = pack_bf16(v[6 % EPC], v[7 % EPC]);
    *reinterpret_cast<uint4*>(base) = w;
  }
}

template <int PREC>
__global__ void pack_b_kernel(const float* __restrict__ src, int O, int I, int R, int S, int src_mode, int ld, int Nn, int Npad,
                              int Ca, unsigned char* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int taps = src_mode == 2 ? 1 : R * S;
  const long long total = (long long)taps * ((Ca + Prec<PREC>::BKC - 1) / Prec<PREC>::BKC) * Npad * 8;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
    pack_b_item<PREC>(src, O, I, R, S, src_mode, ld, Nn, Npad, Ca, out, idx);
}

// every layer's weight operand in one launch: job j owns items [item_begin_j, item_begin_{j+1})
__global__ void __launch_bounds__(256) pack_b_batch_kernel(const PcPackJob* __restrict__ jobs, int n_jobs, long long total) {
  pdl_trigger();
  pdl_wait();
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_jobs - 1;               // last job whose item_begin <= idx
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].item_begin <= idx) lo = mid; else hi = mid - 1;
    }
    const PcPackJob jb = jobs[lo];
    const int ca = jb.dgrad ? jb.O : jb.I, nn = jb.dgrad ? jb.I : jb.O;
    const int bn = nn <= 32 ? 32 : (nn <= 64 ? 64 : 128);
    const int npad = (nn + bn - 1) / bn * bn;
    const long long li = idx - jb.item_begin;
    unsigned char* out = static_cast<unsigned char*>(jb.out);
    if (jb.prec == PC_PREC_TF32X3) pack_b_item<PC_PREC_TF32X3>(jb.w_oihw, jb.O, jb.I, jb.R, jb.S, jb.dgrad, 0, nn, npad, ca, out, li);
    else if (jb.prec == PC_PREC_FP16X2) pack_b_item<PC_PREC_FP16X2>(jb.w_oihw, jb.O, jb.I, jb.R, jb.S, jb.dgrad, 0, nn, npad, ca, out, li);
    else pack_b_item<PC_PREC_BF16>(jb.w_oihw, jb.O, jb.I, jb.R, jb.S, jb.dgrad, 0, nn, npad, ca, out, li);
  }
}

static inline int pick_bn(int Nn) { return Nn <= 32 ? 32 : (Nn <= 64 ? 64 : 128); }
static inline int npad_of(int Nn) { const int bn = pick_bn(Nn); return ceil_div(Nn, bn) * bn; }
static inline int bkc_of(int prec) { return prec == PC_PREC_TF32X3 ? 32 : 64; }
static inline int parts_of(int prec) { return prec == PC_PREC_BF16 ? 1 : 2; }
static inline bool tc_prec(int prec) { return prec == PC_PREC_TF32X3 || prec == PC_PREC_BF16 || prec == PC_PREC_FP16X2; }

template <int PREC>
static void launch_pack(int grid, pc_stream_t stream, const float* src, int O, int I, int R, int S, int src_mode, int ld, int nn,
                        int npad, int ca, void* out) {
  launch_pdl((pack_b_kernel<PREC>), dim3(grid), dim3(256), 0, stream, src, O, I, R, S, src_mode, ld, nn, npad, ca, (unsigned char*)out);
}
static void launch_pack(int prec, int grid, pc_stream_t stream, const float* src, int O, int I, int R, int S, int src_mode, int ld,
                        int nn, int npad, int ca, void* out) {
  if (prec == PC_PREC_TF32X3) launch_pack<PC_PREC_TF32X3>(grid, stream, src, O, I, R, S, src_mode, ld, nn, npad, ca, out);
  else if (prec == PC_PREC_FP16X2) launch_pack<PC_PREC_FP16X2>(grid, stream, src, O, I, R, S, src_mode, ld, nn, npad, ca, out);
  else launch_pack<PC_PREC_BF16>(grid, stream, src, O, I, R, S, src_mode, ld, nn, npad, ca, out);
}

template <int BN, int PREC>
static int launch(const Params& p0, pc_stream_t stream) {
  Params p = p0;
  constexpr bool kTwoPerSm = (BN <= 64) && (Prec<PREC>::NACC * BN <= 256);
  constexpr int NG = kTwoPerSm ? 2 : 3;
  constexpr int MINB = kTwoPerSm ? 2 : 1;
  constexpr int THREADS = 32 * (4 * NG + 2);
  const uint32_t st = stage_bytes<BN, PREC>();
  int stages = (int)((kTwoPerSm ? 100u * 1024u : SMEM_BUDGET) / st);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages > p.n_kc) stages = p.n_kc < 2 ? 2 : p.n_kc;
  p.stages = stages;
  p.aff_floats = (p.mode == 0 && p.xf.scale != nullptr) ? 2 * p.Ca : 0;
  const size_t smem = (size_t)stages * st + sizeof(uint64_t) * (2 * MAX_STAGES + 1) + 16 + sizeof(float) * 3 * BN + sizeof(int) * (4 * BM + 4) + sizeof(float) * (size_t)p.aff_floats + sizeof(int4) * (size_t)(p.n_kc + 1) + 1024;
  dim3 grid(ceil_div(p.M, BM), p.Npad / BN);
  if constexpr (PREC == PC_PREC_FP16X2) {
    if (p.presplit) {
      static size_t configured_ps = 0;
      if (smem > configured_ps) {
        PC_CUDA(cudaFuncSetAttribute((igemm_tc_kernel<BN, PREC, NG, MINB, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_ps = smem;
      }
      launch_pdl(igemm_tc_kernel<BN, PREC, NG, MINB, true>, dim3(grid), dim3(THREADS), smem, stream, p);
      PC_LAUNCH_CHECK("igemm_tc_kernel<presplit>");
      return PC_OK;
    }
  }
  PC_REQUIRE(!p.presplit, PC_EUNSUPPORTED, "pre-split fp16 input planes need the FP16X2 precision");
  static size_t configured = 0;
  if (smem > configured) {
    PC_CUDA(cudaFuncSetAttribute((igemm_tc_kernel<BN, PREC, NG, MINB, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  launch_pdl(igemm_tc_kernel<BN, PREC, NG, MINB, false>, dim3(grid), dim3(THREADS), smem, stream, p);
  PC_LAUNCH_CHECK("igemm_tc_kernel");
  return PC_OK;
}

static long long* g_dbg = nullptr;

static int dispatch(const Params& p_in, int prec, pc_stream_t stream) {
  Params p = p_in;
  p.dbg = g_dbg;
  const int bn = pick_bn(p.Nn);
  if (prec == PC_PREC_TF32X3) {
    if (bn == 32) return launch<32, PC_PREC_TF32X3>(p, stream);
    if (bn == 64) return launch<64, PC_PREC_TF32X3>(p, stream);
    return launch<128, PC_PREC_TF32X3>(p, stream);
  }
  if (prec == PC_PREC_FP16X2) {
    if (bn == 32) return launch<32, PC_PREC_FP16X2>(p, stream);
    if (bn == 64) return launch<64, PC_PREC_FP16X2>(p, stream);
    return launch<128, PC_PREC_FP16X2>(p, stream);
  }
  if (bn == 32) return launch<32, PC_PREC_BF16>(p, stream);
  if (bn == 64) return launch<64, PC_PREC_BF16>(p, stream);
  return launch<128, PC_PREC_BF16>(p, stream);
}

}  // namespace tcconv
}  // namespace pc

using namespace pc;
using namespace pc::tcconv;

// Diagnostics: when set, every tensor-core tile writes 16 clock64() stamps (see PC_STAMP) to buf[tile][16].
extern "C" void pc_tc_set_debug(long long* buf) { pc::tcconv::g_dbg = buf; }

extern "C" int pc_conv_halo_supported(const PcConvGeom* g, int dgrad);

extern "C" int pc_conv_tc_supported(const PcConvGeom* g, int dgrad, int prec) {
  if (g == nullptr || !tc_prec(prec)) return 0;
  const int ca = dgrad ? g->Cout : g->Cin, nn = dgrad ? g->Cin : g->Cout;
  // 32-bit element offsets inside the kernel: the gathered tensor and the output must have fewer than 2^31 elements;
  // the tap bitmask holds at most 32 taps
  const long long in_elems = (long long)g->B * g->H * g->W * g->Cin, out_elems = (long long)g->B * g->Ho * g->Wo * g->Cout;
  if (in_elems >= (1LL << 31) || out_elems >= (1LL << 31) || g->R * g->S > 32) return 0;
  if (nn % 4 != 0 || nn < 16) return 0;
  if (ca % bkc_of(prec) == 0) return 1;
  // 32 gathered channels on the FP16X2 engine: only through the halo engine, whose TMA boxes zero-fill channels 32..63 (the weight
  // operand is packed with a zero-padded 64-channel chunk); the per-tap-gather kernel needs whole 64-channel chunks
  return (prec == PC_PREC_FP16X2 && ca == 32 && pc_conv_halo_supported(g, dgrad)) ? 1 : 0;
}

extern "C" size_t pc_conv_tc_packed_bytes(int O, int I, int R, int S, int dgrad, int prec) {
  const int ca = dgrad ? O : I, nn = dgrad ? I : O;
  return (size_t)R * S * ceil_div(ca, bkc_of(prec)) * parts_of(prec) * npad_of(nn) * 128;
}

extern "C" int pc_pack_conv_weight_tc(const float* w_oihw, int O, int I, int R, int S, int dgrad, int prec, void* out,
                                      pc_stream_t stream) {
  PC_REQUIRE(w_oihw && out && O > 0 && I > 0 && R > 0 && S > 0, PC_EINVAL, "pc_pack_conv_weight_tc: bad arguments");
  PC_REQUIRE(tc_prec(prec), PC_EINVAL, "pc_pack_conv_weight_tc: precision must be TF32X3, FP16X2 or BF16");
  const int ca = dgrad ? O : I, nn = dgrad ? I : O;
  PC_REQUIRE(ca % bkc_of(prec) == 0 || (prec == PC_PREC_FP16X2 && ca == 32), PC_EUNSUPPORTED, "pc_pack_conv_weight_tc: %d channels not a multiple of %d", ca,
             bkc_of(prec));
  const int npad = npad_of(nn);
  const long long total = (long long)R * S * ceil_div(ca, bkc_of(prec)) * npad * 8;
  int grid = ceil_div(total, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  launch_pack(prec, grid, stream, w_oihw, O, I, R, S, dgrad ? 1 : 0, 0, nn, npad, ca, out);
  PC_LAUNCH_CHECK("pack_b_kernel");
  return PC_OK;
}

extern "C" int64_t pc_pack_conv_weight_tc_items(int O, int I, int R, int S, int dgrad, int prec) {
  if (!tc_prec(prec)) return 0;
  const int ca = dgrad ? O : I, nn = dgrad ? I : O;
  return (int64_t)R * S * ceil_div(ca, bkc_of(prec)) * npad_of(nn) * 8;
}

extern "C" int pc_pack_conv_weights_tc_batch(const PcPackJob* jobs, int n_jobs, int64_t total_items, pc_stream_t stream) {
  PC_REQUIRE(jobs && n_jobs > 0 && total_items > 0, PC_EINVAL, "pc_pack_conv_weights_tc_batch: bad arguments");
  int grid = ceil_div(total_items, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  launch_pdl(pack_b_batch_kernel, dim3(grid), dim3(256), 0, stream, jobs, n_jobs, (long long)total_items);
  PC_LAUNCH_CHECK("pack_b_batch_kernel");
  return PC_OK;
}

extern "C" int pc_conv_halo_supported(const PcConvGeom* g, int dgrad);
extern "C" int pc_conv_fwd_halo(const void* x_planes, const void* wp, const float* bias, const PcConvGeom* g, float* y, double* stats, pc_stream_t stream);
extern "C" int pc_conv_dgrad_halo(const void* dy_planes, const void* wp, const PcConvGeom* g, float* dx, int accumulate, const float* dy_amax,
                                  pc_stream_t stream);

extern "C" int pc_conv_fwd_tc(const float* x, const void* wp, const float* bias, const PcConvGeom* g, const PcInXform* xf, float* y,
                              double* stats, int prec, pc_stream_t stream) {
  if (!pc_conv_tc_supported(g, 0, prec)) return PC_EUNSUPPORTED;
  // stride-1 3x3 layers on pre-split planes: the halo-resident persistent engine (conv_halo.cu); same weight operand, same outputs
  if (prec == PC_PREC_FP16X2 && xf != nullptr && xf->presplit && !xf->scale && !xf->shift && !xf->drop && !xf->relu && g_dbg == nullptr &&
      pc_conv_halo_supported(g, 0))
    return pc_conv_fwd_halo(x, wp, bias, g, y, stats, stream);
  PC_REQUIRE(g->Cin % bkc_of(prec) == 0, PC_EUNSUPPORTED, "pc_conv_fwd: %d input channels reach the FP16X2 engine only as pre-split planes on the halo path", g->Cin);
  Params p{};
  p.A = x; p.Bp = (const unsigned char*)wp; p.bias = bias; p.C = y; p.stats = stats;
  if (xf != nullptr) { p.xf.scale = xf->scale; p.xf.shift = xf->shift; p.xf.drop = xf->drop; p.xf.relu = xf->relu; }
  if (xf != nullptr && xf->presplit) {
    PC_REQUIRE(prec == PC_PREC_FP16X2 && !xf->scale && !xf->shift && !xf->drop && !xf->relu, PC_EINVAL,
               "pc_conv_fwd: pre-split input planes take no further transform and need PC_PREC_FP16X2");
    p.presplit = 1;
    p.plane_bytes = (size_t)g->B * g->H * g->W * g->Cin * 2;
  }
  p.g = *g; p.mode = 0;
  p.M = (long long)g->B * g->Ho * g->Wo;
  p.Nn = g->Cout; p.Npad = npad_of(g->Cout); p.Ca = g->Cin;
  p.n_kc = g->R * g->S * (g->Cin / bkc_of(prec));
  p.accumulate = 0;
  return dispatch(p, prec, stream);
}

extern "C" int pc_conv_dgrad_tc(const float* dy, const void* wp, const PcConvGeom* g, float* dx, int accumulate, int prec,
                                const float* dy_amax, int dy_presplit, pc_stream_t stream) {
  if (!pc_conv_tc_supported(g, 1, prec)) return PC_EUNSUPPORTED;
  if (prec == PC_PREC_FP16X2 && dy_presplit && dy_amax != nullptr && g_dbg == nullptr && pc_conv_halo_supported(g, 1) &&
      (g->R != 1 || g->stride == 1 || accumulate))      // a strided 1x1 data gradient only touches every second pixel: accumulate-only
    return pc_conv_dgrad_halo(dy, wp, g, dx, accumulate, dy_amax, stream);
  PC_REQUIRE(g->Cout % bkc_of(prec) == 0, PC_EUNSUPPORTED, "pc_conv_dgrad: %d gradient channels reach the FP16X2 engine only as pre-split planes on the halo path", g->Cout);
  Params p{};
  p.A = dy; p.Bp = (const unsigned char*)wp; p.C = dx; p.a_amax = dy_amax;
  if (dy_presplit) {      // dy = fp16 hi | lo planes already scaled by f16_operand_scale(*dy_amax) (pc_bn_*_bwd_apply)
    PC_REQUIRE(prec == PC_PREC_FP16X2 && dy_amax != nullptr, PC_EINVAL, "pc_conv_dgrad: pre-split dy needs PC_PREC_FP16X2 and dy_amax");
    p.presplit = 1;
    p.plane_bytes = (size_t)g->B * g->Ho * g->Wo * g->Cout * 2;
  }
  p.g = *g; p.mode = 1;
  p.M = (long long)g->B * g->H * g->W;
  p.Nn = g->Cin; p.Npad = npad_of(g->Cin); p.Ca = g->Cout;
  p.n_kc = g->R * g->S * (g->Cout / bkc_of(prec));
  p.accumulate = accumulate;
  return dispatch(p, prec, stream);
}

// C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) on the tensor cores; ws >= pc_tc_gemm_workspace(N, K, prec) bytes.
extern "C" size_t pc_tc_gemm_workspace(int N, int K, int prec) {
  return (size_t)(K / bkc_of(prec)) * parts_of(prec) * npad_of(N) * 128;
}

extern "C" int pc_tc_gemm(const float* A, const float* B, const float* bias, float* C, int M, int N, int K, int prec, void* ws,
                          size_t ws_bytes, pc_stream_t stream) {
  PC_REQUIRE(A && B && C && ws && M > 0 && N > 0 && K > 0, PC_EINVAL, "pc_tc_gemm: bad arguments");
  PC_REQUIRE(tc_prec(prec), PC_EINVAL, "pc_tc_gemm: precision must be TF32X3, FP16X2 or BF16");
  PC_REQUIRE(K % bkc_of(prec) == 0 && N % 4 == 0 && N >= 16, PC_EUNSUPPORTED, "pc_tc_gemm: K=%d must be a multiple of %d, N=%d of 4", K,
             bkc_of(prec), N);
  PC_REQUIRE(ws_bytes >= pc_tc_gemm_workspace(N, K, prec), PC_EINVAL, "pc_tc_gemm: workspace too small");
  const int npad = npad_of(N);
  const long long total = (long long)(K / bkc_of(prec)) * npad * 8;
  int grid = ceil_div(total, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  launch_pack(prec, grid, stream, B, 0, 0, 1, 1, 2, K, N, npad, K, ws);
  PC_LAUNCH_CHECK("pack_b_kernel");
  Params p{};
  p.A = A; p.Bp = (const unsigned char*)ws; p.bias = bias; p.C = C;
  p.mode = 2; p.M = M; p.Nn = N; p.Npad = npad; p.Ca = K; p.n_kc = K / bkc_of(prec); p.lda = K;
  p.g.S = 1; p.g.R = 1; p.g.stride = 1;
  return dispatch(p, prec, stream);
}
