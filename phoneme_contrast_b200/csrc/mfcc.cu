// MFCC / log-mel front end with the view augmentations fused into the epilogue.
//
// One CTA per clip: frames are loaded (coalesced, reflect-padded, Hann-windowed), transformed with an
// in-shared-memory mixed-radix real FFT (400 real points = one 200-point complex FFT, 200 = 8 x 5 x 5,
// + split post-processing), reduced through the sparse (banded) mel filterbank, and kept as a
// log-mel tile [n_mels][T] in shared memory. Each view of the clip (gain, time/freq mask, noise --
// src/datasets/dataset.py:79-98) is then a cheap epilogue over that tile: gain is a dB offset, the
// top_db clamp needs the clip-wide maximum (hence the resident tile), the DCT-II is a small
// shared-memory contraction, masks / noise are applied as the result is stored. The spectrogram is
// computed ONCE per clip no matter how many views are emitted.
//
// Reference arithmetic (torchaudio 2.x, see oracle/mfcc_oracle.py for the line-by-line restatement):
//   stft(center, reflect, periodic Hann 400, hop 160) -> |.|^2 -> fb[201,80] -> 10 log10(max(.,1e-10))
//   -> max(., amax - 80) -> dct[80,40]           (functional.py:123-144, 390-405; _transforms.py:701-718)
#include "common.cuh"

namespace pc {

constexpr int FE_NFFT = 400, FE_NC = 200, FE_FC = 16;   // complex points, frames per chunk
constexpr int FE_ZLD = FE_NC + 1;                       // float2 row stride (odd -> fewer bank conflicts)
constexpr int FE_PLD = 204;
constexpr int FE_THREADS = 256;
constexpr float FE_LOG_FLOOR = 1e-37f;

struct FeParams {
  const float* wave; int n_clips, S, wave_ld;
  const float* window; const int* fb_start; const int* fb_len; const float* fb_w; const float* dct; const float* tw;
  int hop, n_mels, n_mfcc, T, TLD, Lsz, dct_sz, xs_len, fb_ld, dct_alias;
  int dct_mma;      // DCT on mma.sync (n_mels, n_mfcc multiples of 8, n_mfcc <= 64, T <= 128): dct staged as tf32 hi | lo arrays
  float preemph;    // optional pre-emphasis y[n] = x[n] - a x[n-1], y[0] = x[0] (0 = off: the reference's code path)
  const PcViewDesc* views; int n_views, views_per_clip;
  const float* noise; int kind, clamp_mode; float top_db; const float* clamp_ref; float* clip_max_out; float* out;
  long long* dbg;   // optional [grid][8] per-phase cycle totals (diagnostics: pc_fe_set_debug)
};
#define FE_T(slot)                                              \
  do {                                                          \
    if (p.dbg != nullptr && tid == 0) {                         \
      const long long now = clock64();                          \
      p.dbg[(size_t)blockIdx.x * 8 + (slot)] += now - t_last;   \
      t_last = now;                                             \
    }                                                           \
  } while (0)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

__device__ __forceinline__ void dft4(const float2 y[4], float2 Y[4]) {
  const float2 c0 = cadd(y[0], y[2]), c1 = csub(y[0], y[2]), c2 = cadd(y[1], y[3]), c3 = cmul_mi(csub(y[1], y[3]));
  Y[0] = cadd(c0, c2); Y[1] = cadd(c1, c3); Y[2] = csub(c0, c2); Y[3] = csub(c1, c3);
}
__device__ __forceinline__ void dft8(float2 x[8]) {
  const float r = 0.70710678118654752f;
  float2 a[4], b[4], E[4], O[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { a[j] = cadd(x[j], x[j + 4]); b[j] = csub(x[j], x[j + 4]); }
  b[1] = make_float2(r * (b[1].x + b[1].y), r * (b[1].y - b[1].x));       // * (1 - i)/sqrt2
  b[2] = cmul_mi(b[2]);                                                   // * (-i)
  b[3] = make_float2(r * (b[3].y - b[3].x), r * (-b[3].x - b[3].y));      // * (-1 - i)/sqrt2
  dft4(a, E);
  dft4(b, O);
#pragma unroll
  for (int q = 0; q < 4; ++q) { x[2 * q] = E[q]; x[2 * q + 1] = O[q]; }
}
__device__ __forceinline__ void dft5(float2 x[5]) {
  const float c1 = 0.30901699437494742f, c2 = -0.80901699437494742f, s1 = 0.95105651629515357f, s2 = 0.58778525229247313f;
  const float2 t1 = cadd(x[1], x[4]), t2 = cadd(x[2], x[3]), t3 = csub(x[1], x[4]), t4 = csub(x[2], x[3]);
  const float2 m1 = make_float2(x[0].x + c1 * t1.x + c2 * t2.x, x[0].y + c1 * t1.y + c2 * t2.y);
  const float2 m2 = make_float2(x[0].x + c2 * t1.x + c1 * t2.x, x[0].y + c2 * t1.y + c1 * t2.y);
  const float2 n1 = make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y);
  const float2 n2 = make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y);
  x[0] = make_float2(x[0].x + t1.x + t2.x, x[0].y + t1.y + t2.y);
  // m - i n = (m.x + n.y, m.y - n.x) ; m + i n = (m.x - n.y, m.y + n.x)
  x[1] = make_float2(m1.x + n1.y, m1.y - n1.x);
  x[4] = make_float2(m1.x - n1.y, m1.y + n1.x);
  x[2] = make_float2(m2.x + n2.y, m2.y - n2.x);
  x[3] = make_float2(m2.x - n2.y, m2.y + n2.x);
}

// ---- DCT on the tensor cores: out[t][c] = sum_m val[m][t] dct[m][c] is a [T x n_mels] x [n_mels x n_mfcc] product per clip. It ran
// as scalar FMAs (28 % of the kernel's instructions for two views); mma.sync m16n8k8 with the 3xTF32 split (hi * hi + hi * lo +
// lo * hi, fp32 accumulate: ~2^-21 relative) does a 16 x 8 x 8 block per instruction. tcgen05 is not the tool for a 101 x 40 x 80
// product inside an FFT kernel: no TMEM / descriptor set-up, operands straight from the log-mel tile in shared memory.
__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// Device noise: element (f, t) of a view takes one of the two Box-Muller normals of the Philox block of its row PAIR (f >> 1, t), so
// a thread that owns two adjacent rows pays one Philox call for both (it used to take one of four normals per call). Fast-math
// log / sincos: these are noise samples, checked statistically.
__device__ __forceinline__ float2 fe_noise_pair(uint32_t view, int f_pair, int t, int T, uint32_t seed) {
  const uint4 rr = Philox::round10(make_uint4((uint32_t)(f_pair * T + t), view, 0x4e4f4953u, 0u), make_uint2(seed, 0x70635f66u));
  const float u1 = Philox::u01(rr.x), u2 = Philox::u01(rr.y);
  const float r = sqrtf(-2.0f * __logf(u1));
  float sn, cs;
  __sincosf(6.283185307179586f * u2, &sn, &cs);
  return make_float2(r * cs, r * sn);
}
__device__ __forceinline__ float fe_noise(uint32_t view, int f, int t, int T, uint32_t seed) {
  const float2 n2 = fe_noise_pair(view, f >> 1, t, T, seed);
  return (f & 1) ? n2.y : n2.x;
}

// position of natural-order bin k (0..199) after the in-place 8 x 5 x 5 passes
__device__ __forceinline__ int fft_pos(int k) {
  const int k1 = k & 7, k2 = k >> 3;
  const int d = k2 / 5, c = k2 - 5 * d;
  return 25 * k1 + 5 * c + d;
}

__global__ void __launch_bounds__(FE_THREADS, 2) frontend_kernel(FeParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: Z [FC][ZLD] float2 | P [FC][PLD] | L [n_mels][TLD] | win [400] | W200 [200] float2 | W400 [201] float2 | dct [n_mels*n_mfcc]
  float2* Z = reinterpret_cast<float2*>(smem_raw);
  float* P = reinterpret_cast<float*>(Z + FE_FC * FE_ZLD);
  float* L = P + FE_FC * FE_PLD;
  float* win = L + p.Lsz;   // Lsz = n_mels*TLD rounded up to 4 floats (keeps float2/float4 tables aligned)
  float2* W200 = reinterpret_cast<float2*>(win + FE_NFFT);
  float2* W400 = W200 + 200;
  // the DCT matrix is only needed after the last frame chunk: when it fits it aliases the FFT work buffers (Z, P), which
  // keeps the CTA at ~104 KB so that two CTAs stay resident per SM
  float* dcts = p.dct_alias ? reinterpret_cast<float*>(Z) : reinterpret_cast<float*>(W400 + 202);     // [dct_sz] (+ [dct_sz] tf32 lo parts with dct_mma)
  float* xs = reinterpret_cast<float*>(W400 + 202) + (p.dct_alias ? 0 : 2 * p.dct_sz);   // 2 x [(FC-1)*hop + n_fft] samples of the current / next frame chunk (reflect-padded)
  float* fbw = xs + 2 * p.xs_len;                                  // [n_mels][fb_ld] filter weights
  int* fbs = reinterpret_cast<int*>(fbw + p.n_mels * p.fb_ld);     // [n_mels] first bin, [n_mels] length
  int* posk = fbs + 2 * p.n_mels;                                  // [101 + 101]: position of bin k, then of bin 200 - k, after the FFT passes
  float* csum = reinterpret_cast<float*>(posk + 202);              // [64] column sums of the DCT matrix
  float2* twa = reinterpret_cast<float2*>(csum + 64);              // [7][25] twiddles of the radix-8 pass in access order: W200^(n2 * k1), k1 = 1..7
  __shared__ float red[FE_THREADS / 32];
  __shared__ float lmax_s;

  const int tid = threadIdx.x;
  long long t_last = clock64();
  const int V = p.views_per_clip;
  const int view0 = blockIdx.x * V;
  const int clip = p.views != nullptr ? p.views[view0].clip : blockIdx.x;
  const float* x = p.wave + (size_t)clip * p.wave_ld;
  const int S = p.S, T = p.T;

  for (int i = tid; i < FE_NFFT; i += FE_THREADS) win[i] = p.window[i];
  for (int i = tid; i < 200; i += FE_THREADS) W200[i] = make_float2(p.tw[2 * i], p.tw[2 * i + 1]);
  for (int i = tid; i < 201; i += FE_THREADS) W400[i] = make_float2(p.tw[400 + 2 * i], p.tw[400 + 2 * i + 1]);
  auto load_dct = [&]() {     // plain fp32, or tf32 hi | lo parts for the tensor-core DCT
    for (int i = tid; i < p.n_mels * p.n_mfcc; i += FE_THREADS) {
      const float v = p.dct[i];
      if (p.dct_mma) {
        const float hi = __uint_as_float(f2tf32(v));
        dcts[i] = hi;
        dcts[p.dct_sz + i] = __uint_as_float(f2tf32(v - hi));
      } else {
        dcts[i] = v;
      }
    }
  };
  if (p.kind == PC_FE_MFCC && !p.dct_alias) load_dct();
  for (int i = tid; i < p.n_mels * p.fb_ld; i += FE_THREADS) {
    const int m = i / p.fb_ld, q = i - m * p.fb_ld;
    fbw[i] = p.fb_w[m * PC_FB_MAXW + q];
  }
  for (int i = tid; i < p.n_mels; i += FE_THREADS) { fbs[i] = p.fb_start[i]; fbs[p.n_mels + i] = p.fb_len[i]; }
  for (int k = tid; k < 101; k += FE_THREADS) { posk[k] = fft_pos(k % 200); posk[101 + k] = fft_pos((200 - k) % 200); }
  for (int i = tid; i < 7 * 25; i += FE_THREADS) {
    const int k1 = i / 25 + 1, n2 = i % 25;
    twa[i] = make_float2(p.tw[2 * (n2 * k1)], p.tw[2 * (n2 * k1) + 1]);
  }
  __syncthreads();

  // Stage the sample span of the frame chunk starting at frame f into dst: 16-byte cp.async for interior vectors (no
  // register staging, completion tracked per commit group), plain loads with reflect padding at the clip edges.
  const float pre = p.preemph;
  const bool vec_ok = pre == 0.f && ((p.hop & 3) == 0) && ((p.wave_ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  auto stage_chunk = [&](int f, float* dst) {
    const int nfc = min(FE_FC, T - f);
    const int base = f * p.hop - FE_NFFT / 2;
    const int nv = ((nfc - 1) * p.hop + FE_NFFT + 3) >> 2;
    for (int v = tid; v < nv; v += FE_THREADS) {
      const int i = base + 4 * v;
      if (vec_ok && i >= 0 && i + 3 < S) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + 4 * v)), "l"(x + i) : "memory");
      } else {
        float t[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          int ii = i + e;
          ii = ii < 0 ? -ii : (ii >= S ? 2 * (S - 1) - ii : ii);
          ii = ii < 0 ? 0 : (ii >= S ? S - 1 : ii);      // only reachable in the unused tail of the last vector
          t[e] = x[ii];
          if (pre != 0.f && ii > 0) t[e] = fmaf(-pre, x[ii - 1], t[e]);     // the pre-emphasised signal is what gets reflect-padded
        }
        *reinterpret_cast<float4*>(dst + 4 * v) = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
  };
  stage_chunk(0, xs);
  asm volatile("cp.async.commit_group;" ::: "memory");
  float lm = -INFINITY;
  const bool hop_even = (p.hop & 1) == 0;
  static_assert(FE_FC == 16, "the mel loop's index split assumes 16-frame chunks");

  for (int f0 = 0; f0 < T; f0 += FE_FC) {
    const int nf = min(FE_FC, T - f0);
    // ---- P0: the chunk's sample span was staged into xs[buf] by the previous iteration's prefetch (cp.async, overlapped
    // with that chunk's FFT); stage the NEXT chunk now, then wait for the current one.
    const int buf = (f0 / FE_FC) & 1;
    if (f0 + FE_FC < T) stage_chunk(f0 + FE_FC, xs + (buf ^ 1) * p.xs_len);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();
    const float* xc = xs + buf * p.xs_len;
    FE_T(0);
    // ---- A: 25 radix-8 butterflies per frame over stride-25 points, twiddle W200^(n2*k1). The points come straight from the
    // staged samples: complex point n = (x[2n] w[2n], x[2n+1] w[2n+1]) of the Hann-windowed frame (hop may be odd: scalar loads)
    for (int i = tid; i < nf * 25; i += FE_THREADS) {
      const int fl = i / 25, n2 = i - fl * 25;
      float2* z = Z + fl * FE_ZLD;
      const float* sp = xc + fl * p.hop + 2 * n2;
      float2 v[8];
#pragma unroll
      for (int n1 = 0; n1 < 8; ++n1) {
        const float2 wv = *reinterpret_cast<const float2*>(win + 50 * n1 + 2 * n2);
        float2 xv;
        if (hop_even) xv = *reinterpret_cast<const float2*>(sp + 50 * n1);      // 8-byte aligned: xs, fl * hop and 2 n are even
        else xv = make_float2(sp[50 * n1], sp[50 * n1 + 1]);
        v[n1] = make_float2(xv.x * wv.x, xv.y * wv.y);
      }
      dft8(v);
#pragma unroll
      for (int k1 = 0; k1 < 8; ++k1) z[25 * k1 + n2] = k1 == 0 ? v[0] : cmul(v[k1], twa[(k1 - 1) * 25 + n2]);
    }
    __syncthreads();
    FE_T(1);
    // ---- B1: radix-5 over a (stride 5), twiddle W25^(b*c) = W200^(8*b*c)
#pragma unroll 2
    for (int i = tid; i < nf * 40; i += FE_THREADS) {
      const int fl = i / 40, r = i - fl * 40;
      const int k1 = r / 5, b = r - k1 * 5;
      float2* z = Z + fl * FE_ZLD + 25 * k1 + b;
      float2 v[5];
#pragma unroll
      for (int a = 0; a < 5; ++a) v[a] = z[5 * a];
      dft5(v);
#pragma unroll
      for (int c = 0; c < 5; ++c) z[5 * c] = (c == 0 || b == 0) ? v[c] : cmul(v[c], W200[8 * b * c]);
    }
    __syncthreads();
    // ---- B2: radix-5 over b (stride 1); output bin k = k1 + 8*(c + 5*d) stays at position 25*k1 + 5*c + d
#pragma unroll 2
    for (int i = tid; i < nf * 40; i += FE_THREADS) {
      const int fl = i / 40, r = i - fl * 40;
      float2* z = Z + fl * FE_ZLD + 5 * r;   // r = 5*k1 + c
      float2 v[5];
#pragma unroll
      for (int b = 0; b < 5; ++b) v[b] = z[b];
      dft5(v);
#pragma unroll
      for (int d = 0; d < 5; ++d) z[d] = v[d];
    }
    __syncthreads();
    FE_T(2);
    // ---- post: real-FFT split, power spectrum bins k and 200-k together. All 16 frame slots are processed (thread = (bin, frame),
    // frame fastest; slots beyond nf hold stale finite data and are never stored) and the power spectrum is written TRANSPOSED,
    // PT[bin][16 frames], so that the filterbank below reads four frames per 16-byte load.
    float* PT = P;
#pragma unroll 2
    for (int i = tid; i < 101 * FE_FC; i += FE_THREADS) {
      const int k = i >> 4, fl = i & 15;
      const float2* z = Z + fl * FE_ZLD;
      const float2 zk = z[posk[k]];
      const float2 zn = z[posk[101 + k]];
      // Xe = (zk + conj(zn))/2 ; Xo = -i (zk - conj(zn))/2 ; X[k] = Xe + W400^k Xo ; X[200-k] = conj(Xe - W400^k Xo)
      const float2 xe = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
      const float2 dd = make_float2(0.5f * (zk.x - zn.x), 0.5f * (zk.y + zn.y));
      const float2 xo = cmul_mi(dd);
      const float2 wx = cmul(W400[k], xo);
      const float2 a = cadd(xe, wx), b = csub(xe, wx);
      PT[k * FE_FC + fl] = a.x * a.x + a.y * a.y;
      PT[(200 - k) * FE_FC + fl] = b.x * b.x + b.y * b.y;
    }
    __syncthreads();
    FE_T(3);
    // ---- mel: banded filterbank, stored as 10 log10(mel); thread = (mel bin, group of four frames)
    for (int i = tid; i < 4 * p.n_mels; i += FE_THREADS) {
      const int m = i >> 2, fq = i & 3;
      const int s = fbs[m], len = fbs[p.n_mels + m];
      const float* wq = fbw + m * p.fb_ld;
      const float* pq = PT + s * FE_FC + 4 * fq;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int q = 0; q < len; ++q) {
        const float w = wq[q];
        const float4 pv = *reinterpret_cast<const float4*>(pq + q * FE_FC);
        acc.x = fmaf(pv.x, w, acc.x); acc.y = fmaf(pv.y, w, acc.y); acc.z = fmaf(pv.z, w, acc.z); acc.w = fmaf(pv.w, w, acc.w);
      }
      // 10 log10(x) = 10 log10(2) log2(x)
      const float lv[4] = {3.0102999566398120f * __log2f(fmaxf(acc.x, FE_LOG_FLOOR)), 3.0102999566398120f * __log2f(fmaxf(acc.y, FE_LOG_FLOOR)),
                           3.0102999566398120f * __log2f(fmaxf(acc.z, FE_LOG_FLOOR)), 3.0102999566398120f * __log2f(fmaxf(acc.w, FE_LOG_FLOOR))};
      float* lp = L + m * p.TLD + f0 + 4 * fq;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (4 * fq + e < nf) {
          lp[e] = lv[e];
          lm = fmaxf(lm, lv[e]);            // clip-wide maximum, tracked as the tile is written
        }
      }
    }
    __syncthreads();
    FE_T(4);
  }

  if (p.kind == PC_FE_MFCC && p.dct_alias) load_dct();     // FFT buffers are free now (the frame loop ended with a barrier)
  // ---- clip-wide maximum of the log-mel tile (the barriers below also publish the DCT matrix)
  lm = warp_max(lm);
  if ((tid & 31) == 0) red[tid >> 5] = lm;
  __syncthreads();
  if (tid == 0) {
    float v = red[0];
    for (int w = 1; w < FE_THREADS / 32; ++w) v = fmaxf(v, red[w]);
    lmax_s = v;
  }
  if (p.kind == PC_FE_MFCC && p.dct_mma && tid < p.n_mfcc) {       // column sums of the DCT matrix (the gain enters as G * csum[c])
    float a = 0.f;
    for (int m = 0; m < p.n_mels; ++m) a += dcts[m * p.n_mfcc + tid] + dcts[p.dct_sz + m * p.n_mfcc + tid];
    csum[tid] = a;
  }
  __syncthreads();
  const float lmax = lmax_s;
  const float amin_db = -100.0f;   // 10 log10(1e-10)

  FE_T(5);
  const int n_out = p.kind == PC_FE_MFCC ? p.n_mfcc : p.n_mels;
  float dacc[8][4];          // tensor-core DCT accumulators of this warp's 16 frames (up to 8 coefficient tiles)
  float dct_clamp = 0.f;     // clamp level they were computed with
  for (int v = 0; v < V; ++v) {
    const int view = view0 + v;
    if (view >= p.n_views) break;
    PcViewDesc d;
    if (p.views != nullptr) d = p.views[view];
    else { d.gain = 1.f; d.t0 = d.t1 = d.f0 = d.f1 = 0; d.noise_level = 0.f; d.noise_seed = 0u; d.clip = clip; }
    // gain g scales power by g^2: +20 log10|g| dB (dataset.py:165-167 multiplies the waveform)
    const float G = d.gain == 1.0f ? 0.f : 20.0f * log10f(fabsf(d.gain));
    const float vmax = fmaxf(lmax + G, amin_db);
    float floor_db = -INFINITY;
    if (p.clamp_mode == PC_CLAMP_PER_CLIP) floor_db = vmax - p.top_db;
    else if (p.clamp_mode == PC_CLAMP_GIVEN) floor_db = p.clamp_ref[0] - p.top_db;
    if (tid == 0 && p.clip_max_out != nullptr) p.clip_max_out[view] = vmax;
    const float lo = fmaxf(amin_db, floor_db);
    float* o = p.out + (size_t)view * n_out * T;
    const float* nz = p.noise != nullptr ? p.noise + (size_t)view * n_out * T : nullptr;

    if (p.kind == PC_FE_MFCC && p.dct_mma) {
      // val_v[m][t] = max(L + G_v, lo_v) = G_v + max(L, lo_v - G_v): the DCT of the clamped tile is shared by every view of the clip
      // with the same clamp level (all of them unless a gain pushes the floor below -100 dB), the gain enters as G_v * csum[c].
      // Warp w owns frames 16 w .. 16 w + 15 and all coefficient tiles; accumulators stay in registers across the views.
      const int warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tig = lane & 3;
      const int t0 = 16 * warp;
      const int n_nt = p.n_mfcc >> 3;
      const float lclamp = lo - G;
      if (t0 < T) {
        if (v == 0 || lclamp != dct_clamp) {
          dct_clamp = lclamp;
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) dacc[nt][0] = dacc[nt][1] = dacc[nt][2] = dacc[nt][3] = 0.f;
          const int ta = min(t0 + gq, T - 1), tb = min(t0 + gq + 8, T - 1);
          for (int ks = 0; ks < (p.n_mels >> 3); ++ks) {
            const float* Lk = L + (8 * ks + tig) * p.TLD;
            const float av[4] = {fmaxf(Lk[ta], lclamp), fmaxf(Lk[tb], lclamp), fmaxf(Lk[4 * p.TLD + ta], lclamp), fmaxf(Lk[4 * p.TLD + tb], lclamp)};
            uint32_t ah[4], al[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              ah[e] = f2tf32(av[e]);
              al[e] = f2tf32(av[e] - __uint_as_float(ah[e]));
            }
            const float* dk = dcts + (8 * ks + tig) * p.n_mfcc + gq;
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
              if (nt < n_nt) {
                const uint32_t bh0 = __float_as_uint(dk[8 * nt]), bh1 = __float_as_uint(dk[4 * p.n_mfcc + 8 * nt]);
                const uint32_t bl0 = __float_as_uint(dk[p.dct_sz + 8 * nt]), bl1 = __float_as_uint(dk[p.dct_sz + 4 * p.n_mfcc + 8 * nt]);
                mma_tf32_16x8x8(dacc[nt], al, bh0, bh1);      // small terms first
                mma_tf32_16x8x8(dacc[nt], ah, bl0, bl1);
                mma_tf32_16x8x8(dacc[nt], ah, bh0, bh1);
              }
            }
          }
        }
        // accumulator element e of tile nt: frame t0 + gq + 8 (e >> 1), coefficient 8 nt + 2 tig + (e & 1)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          if (nt < n_nt) {
            const int c0 = 8 * nt + 2 * tig;
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
              const int t = t0 + gq + 8 * hrow;
              if (t < T) {
                float r0 = fmaf(G, csum[c0], dacc[nt][2 * hrow]), r1 = fmaf(G, csum[c0 + 1], dacc[nt][2 * hrow + 1]);
                const bool tm = t >= d.t0 && t < d.t1;
                if (tm || (c0 >= d.f0 && c0 < d.f1)) r0 = 0.f;
                if (tm || (c0 + 1 >= d.f0 && c0 + 1 < d.f1)) r1 = 0.f;
                if (d.noise_level != 0.f) {
                  float2 nv;
                  if (nz != nullptr) nv = make_float2(nz[c0 * T + t], nz[(c0 + 1) * T + t]);
                  else nv = fe_noise_pair((uint32_t)view, c0 >> 1, t, T, d.noise_seed);
                  r0 = fmaf(nv.x, d.noise_level, r0);
                  r1 = fmaf(nv.y, d.noise_level, r1);
                }
                o[c0 * T + t] = r0;
                o[(c0 + 1) * T + t] = r1;
              }
            }
          }
        }
      }
    } else if (p.kind == PC_FE_MFCC) {
      const int n_cg = (p.n_mfcc + 7) >> 3;
      for (int i = tid; i < n_cg * T; i += FE_THREADS) {
        const int cg = i / T, t = i - cg * T;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int c0 = cg * 8;
        if (c0 + 8 <= p.n_mfcc && (p.n_mfcc & 3) == 0) {
#pragma unroll 4
          for (int m = 0; m < p.n_mels; ++m) {
            const float val = fmaxf(L[m * p.TLD + t] + G, lo);
            const float4 d0 = *reinterpret_cast<const float4*>(dcts + m * p.n_mfcc + c0);
            const float4 d1 = *reinterpret_cast<const float4*>(dcts + m * p.n_mfcc + c0 + 4);
            acc[0] = fmaf(val, d0.x, acc[0]); acc[1] = fmaf(val, d0.y, acc[1]); acc[2] = fmaf(val, d0.z, acc[2]); acc[3] = fmaf(val, d0.w, acc[3]);
            acc[4] = fmaf(val, d1.x, acc[4]); acc[5] = fmaf(val, d1.y, acc[5]); acc[6] = fmaf(val, d1.z, acc[6]); acc[7] = fmaf(val, d1.w, acc[7]);
          }
        } else {
          for (int m = 0; m < p.n_mels; ++m) {
            const float val = fmaxf(L[m * p.TLD + t] + G, lo);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (c0 + q < p.n_mfcc) acc[q] = fmaf(val, dcts[m * p.n_mfcc + c0 + q], acc[q]);
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int c = c0 + q;
          if (c >= p.n_mfcc) break;
          float r = acc[q];
          if ((t >= d.t0 && t < d.t1) || (c >= d.f0 && c < d.f1)) r = 0.f;
          if (d.noise_level != 0.f) {
            float nv;
            if (nz != nullptr) nv = nz[c * T + t];
            else nv = fe_noise((uint32_t)view, c, t, T, d.noise_seed);
            r = fmaf(nv, d.noise_level, r);
          }
          o[c * T + t] = r;
        }
      }
    } else {
      for (int i = tid; i < p.n_mels * T; i += FE_THREADS) {
        const int m = i / T, t = i - m * T;
        float r = fmaxf(L[m * p.TLD + t] + G, lo);
        if ((t >= d.t0 && t < d.t1) || (m >= d.f0 && m < d.f1)) r = 0.f;
        if (d.noise_level != 0.f) {
          float nv;
          if (nz != nullptr) nv = nz[i];
          else nv = fe_noise((uint32_t)view, m, t, T, d.noise_seed);
          r = fmaf(nv, d.noise_level, r);
        }
        o[i] = r;
      }
    }
  }
  FE_T(6);
}

__global__ void __launch_bounds__(256) reduce_max_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
  __shared__ float red[8];
  float v = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 256) v = fmaxf(v, x[i]);
  v = warp_max(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) v = fmaxf(v, red[w]);
    out[0] = fmaxf(v, red[0]);
  }
}

__global__ void augment_apply_kernel(const float* __restrict__ x, const PcViewDesc* __restrict__ views, int n_views, int F, int T,
                                     const float* __restrict__ noise, float* __restrict__ out) {
  const long long total = (long long)n_views * F * T;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(idx % T);
    const int f = (int)((idx / T) % F);
    const int v = (int)(idx / ((long long)T * F));
    const PcViewDesc d = views[v];
    float r = x[idx];
    if ((t >= d.t0 && t < d.t1) || (f >= d.f0 && f < d.f1)) r = 0.f;
    if (d.noise_level != 0.f) {
      float nv;
      if (noise != nullptr) nv = noise[idx];
      else nv = fe_noise((uint32_t)v, f, t, T, d.noise_seed);
      r = fmaf(nv, d.noise_level, r);
    }
    out[idx] = r;
  }
}

// torchaudio compute_deltas, win_length 5, replicate padding: d_t = sum_{k=-2..2} k x_{t+k} / 10
__global__ void compute_deltas_kernel(const float* __restrict__ x, int rows, int T, float* __restrict__ out) {
  const long long total = (long long)rows * T;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(idx % T);
    const float* r = x + (idx - t);
    float s = 0.f;
#pragma unroll
    for (int k = -2; k <= 2; ++k) {
      int tt = t + k;
      tt = tt < 0 ? 0 : (tt >= T ? T - 1 : tt);
      s = fmaf((float)k, r[tt], s);
    }
    out[idx] = s / 10.0f;
  }
}

static long long* g_fe_dbg = nullptr;
}  // namespace pc

using namespace pc;

// Diagnostics: per-CTA cycle totals of the front-end phases (load, radix-8, radix-5 x2 [slots 1,2 = B1, B2+post...], mel, max, epilogue).
extern "C" void pc_fe_set_debug(long long* buf) { pc::g_fe_dbg = buf; }

extern "C" int pc_frontend_fwd(const float* wave, int n_clips, int S, int wave_ld, const PcMfccConsts* c, const PcViewDesc* views,
                               int n_views, const float* noise, int kind, int clamp_mode, float top_db, const float* clamp_ref,
                               float* clip_max_out, float* out, pc_stream_t stream) {
  PC_REQUIRE(wave && c && out && n_clips > 0 && S > 0 && wave_ld >= S, PC_EINVAL, "pc_frontend_fwd: bad arguments");
  PC_REQUIRE(c->n_fft == FE_NFFT, PC_EUNSUPPORTED, "pc_frontend_fwd: n_fft=%d not built (the in-smem FFT is specialised for 400)", c->n_fft);
  PC_REQUIRE(c->hop > 0 && c->n_mels > 0 && c->n_mels <= 128, PC_EUNSUPPORTED, "pc_frontend_fwd: hop=%d n_mels=%d unsupported", c->hop, c->n_mels);
  PC_REQUIRE(kind == PC_FE_MFCC || kind == PC_FE_LOGMEL, PC_EINVAL, "pc_frontend_fwd: bad kind");
  PC_REQUIRE(kind != PC_FE_MFCC || (c->n_mfcc > 0 && c->n_mfcc <= c->n_mels && c->dct), PC_EINVAL,
             "Cannot select more MFCC coefficients than # mel bins");
  PC_REQUIRE(S > FE_NFFT / 2, PC_EINVAL, "pc_frontend_fwd: reflect padding needs more than %d samples (got %d)", FE_NFFT / 2, S);
  PC_REQUIRE(clamp_mode != PC_CLAMP_GIVEN || clamp_ref, PC_EINVAL, "pc_frontend_fwd: clamp_ref required");
  PC_REQUIRE(c->window && c->fb_start && c->fb_len && c->fb_w && c->tw, PC_EINVAL, "pc_frontend_fwd: null constants");
  int V = 1, n_ctas = n_clips;
  if (views != nullptr) {
    PC_REQUIRE(n_views > 0 && n_views % n_clips == 0, PC_EINVAL, "pc_frontend_fwd: n_views must be a multiple of n_clips (views grouped by clip)");
    V = n_views / n_clips;
  } else {
    n_views = n_clips;
  }
  FeParams p;
  p.wave = wave; p.n_clips = n_clips; p.S = S; p.wave_ld = wave_ld;
  p.window = c->window; p.fb_start = c->fb_start; p.fb_len = c->fb_len; p.fb_w = c->fb_w; p.dct = c->dct; p.tw = c->tw;
  p.hop = c->hop; p.n_mels = c->n_mels; p.n_mfcc = kind == PC_FE_MFCC ? c->n_mfcc : 0;
  p.preemph = c->preemph;
  p.T = 1 + S / c->hop;
  p.TLD = p.T + ((8 - p.T % 32) + 32) % 32;      // TLD % 32 == 8: the tensor-core DCT's fragment loads (4 mel rows x 8 frames) hit 32 distinct banks
  p.Lsz = (p.n_mels * p.TLD + 3) & ~3;
  p.dct_sz = (p.n_mels * p.n_mfcc + 3) & ~3;
  p.xs_len = ((FE_FC - 1) * c->hop + FE_NFFT + 7) & ~3;
  p.fb_ld = c->fb_wmax > 0 ? ((c->fb_wmax + 3) & ~3) : PC_FB_MAXW;
  p.dct_mma = (kind == PC_FE_MFCC && p.n_mels % 8 == 0 && p.n_mfcc % 8 == 0 && p.n_mfcc <= 64 && p.T <= 16 * (FE_THREADS / 32)) ? 1 : 0;
  p.dct_alias = (sizeof(float) * 2 * (size_t)p.dct_sz <= sizeof(float2) * FE_FC * FE_ZLD + sizeof(float) * FE_FC * FE_PLD) ? 1 : 0;
  p.views = views; p.n_views = n_views; p.views_per_clip = V;
  p.noise = noise; p.kind = kind; p.clamp_mode = clamp_mode; p.top_db = top_db; p.clamp_ref = clamp_ref;
  p.clip_max_out = clip_max_out; p.out = out;
  p.dbg = g_fe_dbg;
  size_t smem = sizeof(float2) * FE_FC * FE_ZLD + sizeof(float) * (FE_FC * FE_PLD + (size_t)p.Lsz + FE_NFFT) +
                sizeof(float2) * (200 + 202) + sizeof(float) * ((p.dct_alias ? 0 : 2 * (size_t)p.dct_sz) + 2 * (size_t)p.xs_len + (size_t)p.n_mels * p.fb_ld + 2 * (size_t)p.n_mels + 202 + 64 + 2 * 7 * 25 + 4);
  PC_REQUIRE(smem <= 227 * 1024, PC_EUNSUPPORTED, "pc_frontend_fwd: clip of %d samples (%d frames) needs %zu B shared memory (> 227 KB)", S, p.T, smem);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    PC_CUDA(cudaFuncSetAttribute(frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  frontend_kernel<<<n_ctas, FE_THREADS, smem, stream>>>(p);
  PC_LAUNCH_CHECK("frontend_kernel");
  return PC_OK;
}

extern "C" int pc_reduce_max(const float* x, int n, float* out, pc_stream_t stream) {
  PC_REQUIRE(x && out && n > 0, PC_EINVAL, "pc_reduce_max: bad arguments");
  reduce_max_kernel<<<1, 256, 0, stream>>>(x, n, out);
  PC_LAUNCH_CHECK("reduce_max_kernel");
  return PC_OK;
}

extern "C" int pc_augment_apply(const float* x, const PcViewDesc* views, int n_views, int F, int T, const float* noise, float* out,
                                pc_stream_t stream) {
  PC_REQUIRE(x && views && out && n_views > 0 && F > 0 && T > 0, PC_EINVAL, "pc_augment_apply: bad arguments");
  const long long total = (long long)n_views * F * T;
  int grid = ceil_div(total, 256);
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  augment_apply_kernel<<<grid, 256, 0, stream>>>(x, views, n_views, F, T, noise, out);
  PC_LAUNCH_CHECK("augment_apply_kernel");
  return PC_OK;
}

extern "C" int pc_compute_deltas(const float* x, int rows, int T, float* out, pc_stream_t stream) {
  PC_REQUIRE(x && out && rows > 0 && T > 0, PC_EINVAL, "pc_compute_deltas: bad arguments");
  const long long total = (long long)rows * T;
  int grid = ceil_div(total, 256);
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  compute_deltas_kernel<<<grid, 256, 0, stream>>>(x, rows, T, out);
  PC_LAUNCH_CHECK("compute_deltas_kernel");
  return PC_OK;
}
