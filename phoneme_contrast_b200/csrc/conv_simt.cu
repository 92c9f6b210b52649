// fp32 SIMT convolution kernels (NHWC): implicit-GEMM forward / data-gradient / weight-gradient for
// Cin % 16 == 0 layers, and direct kernels for the Cin == 1 stem convolutions (3x3 -> 32, 7x7 -> 64).
// These are the exact-fp32 path (PC_PREC_FP32); the tcgen05 tensor-core path lives in conv_tc.cu.
//
// Reference layers: nn.Conv2d in src/models/phoneme_cnn.py:36-61 (PhonemeNet), :159-170 (ResidualBlock),
// :212 (init_conv). GEMM view: M = B*Ho*Wo output pixels, N = Cout, K = R*S*Cin.
//
// Fusions: (1) the previous BatchNorm-apply + ReLU + Dropout2d multiplier is applied while the input
// tile is gathered (PcInXform), so that activation never round-trips HBM; (2) the epilogue adds the
// bias and accumulates the per-channel sum / sum-of-squares the next train-mode BatchNorm needs.
#include <stdlib.h>

#include "common.cuh"

namespace pc {

constexpr int CG_BM = 128, CG_BK = 16, CG_THREADS = 256;

struct XformDev {
  const float* scale;
  const float* shift;
  const float* drop;
  int relu;
};

// ------------------------------------------------------------------------------------------------ weight packing
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int O, int I, int R, int S, float* __restrict__ wf,
                                        float* __restrict__ wd) {
  const long long n = (long long)O * I * R * S;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (long long)gridDim.x * blockDim.x) {
    // idx enumerates OIHW
    const int s = (int)(idx % S);
    const int r = (int)((idx / S) % R);
    const int i = (int)((idx / ((long long)S * R)) % I);
    const int o = (int)(idx / ((long long)S * R * I));
    const float v = w[idx];
    if (wf != nullptr) wf[((size_t)(r * S + s) * I + i) * O + o] = v;   // [(r,s,c)][o]
    if (wd != nullptr) wd[((size_t)(r * S + s) * O + o) * I + i] = v;   // [(r,s,o)][c]
  }
}

// ------------------------------------------------------------------------------------------------ implicit GEMM
// MODE 0: forward   A = xform(x)[B,H,W,Cin], rows = output pixels, taps shift the input window.
// MODE 1: dgrad     A = dy[B,Ho,Wo,Cout],    rows = input pixels; tap (r,s) contributes iff (h+pad-r) % stride == 0.
// C[M x N] = sum_taps A_tap[M x Ca] * Wp[tap][Ca x N]   (Ca = channels of A, N = output channels of this GEMM)
template <int BN, int MODE>
__global__ void __launch_bounds__(CG_THREADS)
conv_igemm_kernel(const float* __restrict__ A, const float* __restrict__ Wp, const float* __restrict__ bias, PcConvGeom g,
                  XformDev xf, float* __restrict__ Cout_ptr, double* __restrict__ stats, int accumulate) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[2][CG_BK][CG_BM + 4];
  __shared__ __align__(16) float Bs[2][CG_BK][BN + 4];
  // BatchNorm-statistics scratch [2][16][BN] aliases As after the main loop (2*16*BN <= 2*16*(BM+4) floats)
  float (*red)[16][BN] = reinterpret_cast<float (*)[16][BN]>(&As[0][0][0]);
  static_assert(BN <= CG_BM + 4, "stats scratch must fit in As");

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // GEMM dims for this mode
  const int Ca = MODE == 0 ? g.Cin : g.Cout;     // channels of the gathered operand
  const int Nn = MODE == 0 ? g.Cout : g.Cin;     // GEMM N
  const int Hr = MODE == 0 ? g.Ho : g.H, Wr = MODE == 0 ? g.Wo : g.W;   // row-pixel grid
  const int Ha = MODE == 0 ? g.H : g.Ho, Wa = MODE == 0 ? g.W : g.Wo;   // gathered tensor grid
  const long long M = (long long)g.B * Hr * Wr;
  const long long m0 = (long long)blockIdx.x * CG_BM;
  const int n0 = blockIdx.y * BN;

  // loader: row = tid & 127, channel half = tid >> 7 (8 floats each)
  const int lrow = tid & 127, lhalf = tid >> 7;
  const long long lm = m0 + lrow;
  const bool lvalid = lm < M;
  int lb = 0, lh = 0, lw = 0;
  if (lvalid) {
    lw = (int)(lm % Wr);
    lh = (int)((lm / Wr) % Hr);
    lb = (int)(lm / ((long long)Wr * Hr));
  }
  const int chunks_per_tap = Ca / CG_BK;
  const int n_chunks = g.R * g.S * chunks_per_tap;

  // B loader: BK x BN floats = 16*BN/4 float4; thread t loads float4 #t (+256 ..)
  constexpr int B_F4 = CG_BK * BN / 4;
  constexpr int B_PER_T = (B_F4 + CG_THREADS - 1) / CG_THREADS;

  float4 ra[2];
  float4 rb[B_PER_T];

  auto load_chunk = [&](int chunk) {
    const int tap = chunk / chunks_per_tap;
    const int c0 = (chunk - tap * chunks_per_tap) * CG_BK + lhalf * 8;
    const int r = tap / g.S, s = tap - r * g.S;
    bool ok = lvalid;
    int ha, wa;
    if (MODE == 0) {
      ha = lh * g.stride - g.pad + r;
      wa = lw * g.stride - g.pad + s;
    } else {
      const int hn = lh + g.pad - r, wn = lw + g.pad - s;
      ok = ok && hn >= 0 && wn >= 0 && (hn % g.stride == 0) && (wn % g.stride == 0);
      ha = hn / g.stride;
      wa = wn / g.stride;
    }
    ok = ok && ha >= 0 && ha < Ha && wa >= 0 && wa < Wa;
    if (ok) {
      const float* p = A + (((size_t)lb * Ha + ha) * Wa + wa) * Ca + c0;
      ra[0] = *reinterpret_cast<const float4*>(p);
      ra[1] = *reinterpret_cast<const float4*>(p + 4);
      if (MODE == 0 && xf.scale != nullptr) {
        const float4 s0 = *reinterpret_cast<const float4*>(xf.scale + c0), s1 = *reinterpret_cast<const float4*>(xf.scale + c0 + 4);
        const float4 t0 = *reinterpret_cast<const float4*>(xf.shift + c0), t1 = *reinterpret_cast<const float4*>(xf.shift + c0 + 4);
        ra[0] = make_float4(fmaf(ra[0].x, s0.x, t0.x), fmaf(ra[0].y, s0.y, t0.y), fmaf(ra[0].z, s0.z, t0.z), fmaf(ra[0].w, s0.w, t0.w));
        ra[1] = make_float4(fmaf(ra[1].x, s1.x, t1.x), fmaf(ra[1].y, s1.y, t1.y), fmaf(ra[1].z, s1.z, t1.z), fmaf(ra[1].w, s1.w, t1.w));
      }
      if (MODE == 0 && xf.relu) {
        ra[0] = make_float4(fmaxf(ra[0].x, 0.f), fmaxf(ra[0].y, 0.f), fmaxf(ra[0].z, 0.f), fmaxf(ra[0].w, 0.f));
        ra[1] = make_float4(fmaxf(ra[1].x, 0.f), fmaxf(ra[1].y, 0.f), fmaxf(ra[1].z, 0.f), fmaxf(ra[1].w, 0.f));
      }
      if (MODE == 0 && xf.drop != nullptr) {
        const float4 d0 = *reinterpret_cast<const float4*>(xf.drop + (size_t)lb * Ca + c0);
        const float4 d1 = *reinterpret_cast<const float4*>(xf.drop + (size_t)lb * Ca + c0 + 4);
        ra[0] = make_float4(ra[0].x * d0.x, ra[0].y * d0.y, ra[0].z * d0.z, ra[0].w * d0.w);
        ra[1] = make_float4(ra[1].x * d1.x, ra[1].y * d1.y, ra[1].z * d1.z, ra[1].w * d1.w);
      }
    } else {
      ra[0] = ra[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float* wrow = Wp + (size_t)chunk * CG_BK * Nn;   // chunk*BK == tap*Ca + c_base
#pragma unroll
    for (int q = 0; q < B_PER_T; ++q) {
      const int f = tid + q * CG_THREADS;
      rb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f < B_F4) {
        const int kk = f / (BN / 4), nq = f % (BN / 4);
        const int n = n0 + nq * 4;
        if (n < Nn) rb[q] = *reinterpret_cast<const float4*>(wrow + (size_t)kk * Nn + n);
      }
    }
  };
  auto store_chunk = [&](int buf) {
    const int kb = lhalf * 8;
    As[buf][kb + 0][lrow] = ra[0].x; As[buf][kb + 1][lrow] = ra[0].y; As[buf][kb + 2][lrow] = ra[0].z; As[buf][kb + 3][lrow] = ra[0].w;
    As[buf][kb + 4][lrow] = ra[1].x; As[buf][kb + 5][lrow] = ra[1].y; As[buf][kb + 6][lrow] = ra[1].z; As[buf][kb + 7][lrow] = ra[1].w;
#pragma unroll
    for (int q = 0; q < B_PER_T; ++q) {
      const int f = tid + q * CG_THREADS;
      if (f < B_F4) {
        const int kk = f / (BN / 4), nq = f % (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][kk][nq * 4]) = rb[q];
      }
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_chunk(0);
  store_chunk(0);
  __syncthreads();
  for (int chunk = 0; chunk < n_chunks; ++chunk) {
    const int buf = chunk & 1;
    if (chunk + 1 < n_chunks) load_chunk(chunk + 1);
#pragma unroll
    for (int k = 0; k < CG_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN];
      if (TN == 2) {
        const float2 b = *reinterpret_cast<const float2*>(&Bs[buf][k][tx * 2]);
        bv[0] = b.x; bv[1] = b.y;
      } else {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * TN + j]);
          bv[j] = b.x; bv[j + 1] = b.y; bv[j + 2] = b.z; bv[j + 3] = b.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (chunk + 1 < n_chunks) store_chunk(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue: bias, store, BatchNorm statistics
  float bsum[TN], bsq[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) { bsum[j] = 0.f; bsq[j] = 0.f; }
  float bj[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int n = n0 + tx * TN + j;
    bj[j] = (bias != nullptr && n < Nn) ? bias[n] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + ty * 8 + i;
    if (m < M) {
      float* dst = Cout_ptr + (size_t)m * Nn + n0 + tx * TN;
      float v[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) v[j] = acc[i][j] + bj[j];
      if (n0 + tx * TN < Nn) {   // Nn % 4 == 0 and TN in {2,4,8}: whole vector in or out except TN == 8 / Nn % 8 != 0
        if (TN == 2) {
          float2 o = make_float2(v[0], v[1]);
          if (accumulate) { const float2 p = *reinterpret_cast<const float2*>(dst); o.x += p.x; o.y += p.y; }
          *reinterpret_cast<float2*>(dst) = o;
        } else {
#pragma unroll
          for (int j = 0; j < TN; j += 4) {
            if (n0 + tx * TN + j < Nn) {
              float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              if (accumulate) { const float4 p = *reinterpret_cast<const float4*>(dst + j); o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
              *reinterpret_cast<float4*>(dst + j) = o;
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < TN; ++j) { bsum[j] += v[j]; bsq[j] = fmaf(v[j], v[j], bsq[j]); }
    }
  }
  if (stats != nullptr) {
#pragma unroll
    for (int j = 0; j < TN; ++j) { red[0][ty][tx * TN + j] = bsum[j]; red[1][ty][tx * TN + j] = bsq[j]; }
    __syncthreads();
    if (tid < BN) {
      const int n = n0 + tid;
      if (n < Nn) {
        double s = 0.0, q = 0.0;
#pragma unroll
        for (int t = 0; t < 16; ++t) { s += (double)red[0][t][tid]; q += (double)red[1][t][tid]; }
        atomicAdd(stats + n, s);
        atomicAdd(stats + Nn + n, q);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ wgrad
// dWp[(tap, c)][n] = sum_m xform(x)[pix(m,tap)][c] * dy[m][n]; CTA = (tap, 64-channel tile, 64-n tile, M split).
// partial layout: [split][K + 1][Cout] (row K = column sums of dy = bias gradient, written by tap 0 / c-tile 0 CTAs).
constexpr int WG_TC = 64, WG_TN = 64, WG_BM = 16;
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, PcConvGeom g, XformDev xf,
                  float* __restrict__ partial, int n_splits, int rows_per_split) {
  __shared__ __align__(16) float Xs[WG_BM][WG_TC + 4];
  __shared__ __align__(16) float Ds[WG_BM][WG_TN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;   // thread: c = ty*4.., n = tx*4..
  const int c_tiles = ceil_div(g.Cin, WG_TC);
  const int tap = blockIdx.x / c_tiles, ct = blockIdx.x % c_tiles;
  const int r = tap / g.S, s = tap % g.S;
  const int c0 = ct * WG_TC, n0 = blockIdx.y * WG_TN;
  const int split = blockIdx.z;
  const long long M = (long long)g.B * g.Ho * g.Wo;
  const long long m_begin = (long long)split * rows_per_split;
  long long m_end = m_begin + rows_per_split;
  if (m_end > M) m_end = M;
  const bool do_bias = (tap == 0 && ct == 0);

  // loader: row = tid >> 4 (0..15), float4 index = tid & 15 -> 64 floats per row for each operand
  const int lr = tid >> 4, lq = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bacc = 0.f;   // bias partial for column n0 + (tid & 63), rows handled by tid >> 6

  for (long long mb = m_begin; mb < m_end; mb += WG_BM) {
    const long long m = mb + lr;
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), dv = xv;
    if (m < m_end) {
      const int wo = (int)(m % g.Wo);
      const int ho = (int)((m / g.Wo) % g.Ho);
      const int b = (int)(m / ((long long)g.Wo * g.Ho));
      const int h = ho * g.stride - g.pad + r, w = wo * g.stride - g.pad + s;
      const int c = c0 + lq * 4;
      if (h >= 0 && h < g.H && w >= 0 && w < g.W && c < g.Cin) {
        xv = *reinterpret_cast<const float4*>(x + (((size_t)b * g.H + h) * g.W + w) * g.Cin + c);
        if (xf.scale != nullptr) {
          const float4 sc = *reinterpret_cast<const float4*>(xf.scale + c), sh = *reinterpret_cast<const float4*>(xf.shift + c);
          xv = make_float4(fmaf(xv.x, sc.x, sh.x), fmaf(xv.y, sc.y, sh.y), fmaf(xv.z, sc.z, sh.z), fmaf(xv.w, sc.w, sh.w));
        }
        if (xf.relu) xv = make_float4(fmaxf(xv.x, 0.f), fmaxf(xv.y, 0.f), fmaxf(xv.z, 0.f), fmaxf(xv.w, 0.f));
        if (xf.drop != nullptr) {
          const float4 d = *reinterpret_cast<const float4*>(xf.drop + (size_t)b * g.Cin + c);
          xv = make_float4(xv.x * d.x, xv.y * d.y, xv.z * d.z, xv.w * d.w);
        }
      }
      const int n = n0 + lq * 4;
      if (n < g.Cout) dv = *reinterpret_cast<const float4*>(dy + (size_t)m * g.Cout + n);
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&Xs[lr][lq * 4]) = xv;
    *reinterpret_cast<float4*>(&Ds[lr][lq * 4]) = dv;
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < WG_BM; ++mm) {
      const float4 a = *reinterpret_cast<const float4*>(&Xs[mm][ty * 4]);
      const float4 d = *reinterpret_cast<const float4*>(&Ds[mm][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, dvv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], dvv[j], acc[i][j]);
    }
    if (do_bias) {
#pragma unroll
      for (int mm = 0; mm < WG_BM / 4; ++mm) bacc += Ds[(tid >> 6) * (WG_BM / 4) + mm][tid & 63];
    }
  }
  const int K = g.R * g.S * g.Cin;
  float* pbase = partial + (size_t)split * (K + 1) * g.Cout;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty * 4 + i;
    const int n = n0 + tx * 4;
    if (c < g.Cin && n < g.Cout)
      *reinterpret_cast<float4*>(pbase + ((size_t)tap * g.Cin + c) * g.Cout + n) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
  if (do_bias) {
    __syncthreads();
    float* red = &Xs[0][0];   // reuse: 256 floats
    red[tid] = bacc;
    __syncthreads();
    if (tid < 64) {
      const int n = n0 + tid;
      if (n < g.Cout) pbase[(size_t)K * g.Cout + n] = red[tid] + red[tid + 64] + red[tid + 128] + red[tid + 192];
    }
  }
}

// partial [splits][K+1][Cout] -> dw OIHW + db, fixed summation order. Block = 32 float4 columns x 8 split lanes: lane y sums
// splits y, y+8, ... (independent 16-byte loads in flight), then the 8 partial sums are combined in a fixed order through
// shared memory, so the result does not depend on scheduling.
__global__ void __launch_bounds__(256)
conv_wgrad_reduce_kernel(const float* __restrict__ partial, int n_splits, int R, int S, int Cin, int Cout,
                         float* __restrict__ dw, float* __restrict__ db, const float* __restrict__ bias_partial, int n_bias) {
  pdl_trigger();
  pdl_wait();
  __shared__ double sh[8][32][4];
  const int K = R * S * Cin;
  const long long total = (long long)(K + 1) * Cout;      // floats per split; Cout % 4 == 0
  const long long total4 = total >> 2;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long i4 = (long long)blockIdx.x * 32 + tx;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (i4 < total4) {
    // the bias row (k == K) may come from a separate set of partials [n_bias][Cout] (column sums of dy) instead of the splits
    const bool bias_row = bias_partial != nullptr && (i4 << 2) >= (long long)K * Cout;
    const float4* src = bias_row ? reinterpret_cast<const float4*>(bias_partial) + (i4 - ((long long)K * Cout >> 2))
                                 : reinterpret_cast<const float4*>(partial) + i4;
    const size_t stride4 = bias_row ? (size_t)(Cout >> 2) : (size_t)total4;
    const int n_part = bias_row ? n_bias : n_splits;
    for (int sp = ty; sp < n_part; sp += 8) {
      const float4 a = src[(size_t)sp * stride4];
      s0 += (double)a.x; s1 += (double)a.y; s2 += (double)a.z; s3 += (double)a.w;
    }
  }
  sh[ty][tx][0] = s0; sh[ty][tx][1] = s1; sh[ty][tx][2] = s2; sh[ty][tx][3] = s3;
  __syncthreads();
  if (ty == 0 && i4 < total4) {
    double sv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double t = sh[0][tx][q];
#pragma unroll
      for (int y = 1; y < 8; ++y) t += sh[y][tx][q];
      sv[q] = t;
    }
    const long long idx = i4 << 2;
    const int n = (int)(idx % Cout);
    const int k = (int)(idx / Cout);
    if (k == K) {
      if (db != nullptr) {
#pragma unroll
        for (int q = 0; q < 4; ++q) db[n + q] = (float)sv[q];
      }
    } else if (dw != nullptr) {
      const int c = k % Cin, tap = k / Cin;
      const int r = tap / S, ss = tap % S;
#pragma unroll
      for (int q = 0; q < 4; ++q) dw[(((size_t)(n + q) * Cin + c) * R + r) * S + ss] = (float)sv[q];
    }
  }
}


// The same reduction for the layers with FEW splits and MANY outputs (256- / 512-channel layers: 2 .. 8 splits, up to 2.4 M weights),
// where the kernel above is bound by its scattered 4-byte OIHW stores (one 32-byte sector each). Block = 32 output channels x 8
// input channels x all taps: partial rows are read as 128-byte segments, the tile is transposed in shared memory and every output
// channel's run of 8 * taps consecutive OIHW floats leaves as one segment. No bias row (db comes from the BatchNorm backward).
__global__ void __launch_bounds__(256)
conv_wgrad_reduce_t_kernel(const float* __restrict__ partial, int n_splits, int taps, int Cin, int Cout, float* __restrict__ dw) {
  pdl_trigger();
  pdl_wait();
  __shared__ float tile[32][8 * 9 + 1];
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 8;
  const int K = taps * Cin, rows = taps * 8;
  const size_t stride = (size_t)(K + 1) * Cout;
  for (int item = threadIdx.x; item < rows * 8; item += 256) {
    const int row = item >> 3, f4 = item & 7;
    const int tap = row >> 3, c = row & 7;
    const float* src = partial + (size_t)(tap * Cin + c0 + c) * Cout + n0 + 4 * f4;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 4
    for (int sp = 0; sp < n_splits; ++sp) {
      const float4 a = *reinterpret_cast<const float4*>(src + (size_t)sp * stride);
      s0 += (double)a.x; s1 += (double)a.y; s2 += (double)a.z; s3 += (double)a.w;
    }
    const int j = c * taps + tap;
    tile[4 * f4 + 0][j] = (float)s0; tile[4 * f4 + 1][j] = (float)s1; tile[4 * f4 + 2][j] = (float)s2; tile[4 * f4 + 3][j] = (float)s3;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * rows; i += 256) {
    const int n = i / rows, j = i - n * rows;
    dw[((size_t)(n0 + n) * Cin + c0) * taps + j] = tile[n][j];
  }
}

// ------------------------------------------------------------------------------------------------ stem (Cin == 1)
// y[b,h,w,n] = bias[n] + sum_{r,s} x[b,h+r-p,w+s-p] * w[n][r][s]; lane = output channel (coalesced NHWC store),
// each warp walks a strip of pixels; weights for the lane's channel(s) live in registers.
// WSPLIT: one warp per (output row, 32-channel group) instead of one warp per row owning COUT/32 channels per lane (half the
// weight registers, 16 resident warps instead of 8). Measured SLOWER for 7x7 / 64 channels (289 vs 260 us: every shared-memory
// window load then feeds half as many FMAs), so it is not used; kept as a switch for other shapes.
template <int KS, int COUT, bool WSPLIT = false>   // kernel size, output channels (16 | 32 | 64); lane owns channels cb + lane + 32*q
__global__ void __launch_bounds__(WSPLIT ? 256 * (COUT / 32) : 256)
conv_stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w_oihw, const float* __restrict__ bias, int B,
                     int H, int W, float* __restrict__ y, double* __restrict__ stats) {
  pdl_trigger();
  pdl_wait();
  constexpr int P = KS / 2, CPL = WSPLIT ? 1 : (COUT + 31) / 32;
  constexpr int TH = 8, TW = 32;   // output tile per CTA
  __shared__ float xs[TH + KS - 1][TW + KS - 1 + 1];
  __shared__ float red[2][8][COUT];
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 7;     // warp = output row of the tile
  const int cb = WSPLIT ? (threadIdx.x >> 8) * 32 : 0;                  // first channel of this warp's group
  const int tiles_w = ceil_div(W, TW), tiles_h = ceil_div(H, TH);
  const int tile = blockIdx.x;
  const int b = tile / (tiles_w * tiles_h);
  const int th = (tile / tiles_w) % tiles_h, tw = tile % tiles_w;
  const int h0 = th * TH, w0 = tw * TW;
  for (int i = threadIdx.x; i < (TH + KS - 1) * (TW + KS - 1); i += blockDim.x) {
    const int rr = i / (TW + KS - 1), cc = i % (TW + KS - 1);
    const int h = h0 + rr - P, w = w0 + cc - P;
    xs[rr][cc] = (h >= 0 && h < H && w >= 0 && w < W) ? x[((size_t)b * H + h) * W + w] : 0.f;
  }
  // weights: coalesced global -> shared, then each lane picks up its channels' taps (row stride KS*KS is odd: conflict-free)
  __shared__ float ws[COUT * KS * KS];
  for (int i = threadIdx.x; i < COUT * KS * KS; i += blockDim.x) ws[i] = w_oihw[i];
  __syncthreads();
  float wr[CPL][KS * KS], bv[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) {
    const int n = cb + lane + 32 * q;
    bv[q] = (bias != nullptr && n < COUT) ? bias[n] : 0.f;
#pragma unroll
    for (int t = 0; t < KS * KS; ++t) wr[q][t] = n < COUT ? ws[n * KS * KS + t] : 0.f;
  }
  float ssum[CPL], ssq[CPL];
#pragma unroll
  for (int q = 0; q < CPL; ++q) { ssum[q] = 0.f; ssq[q] = 0.f; }
  // warp `warp` handles output row h0 + warp; 4 adjacent pixels per iteration share one (KS+3)-wide window row held in
  // registers, so each shared-memory (broadcast) load feeds up to 4*CPL FMAs instead of CPL
  const int h = h0 + warp;
  if (h < H) {
    for (int cw = 0; cw < TW && w0 + cw < W; cw += 4) {
      float acc[4][CPL];
#pragma unroll
      for (int px = 0; px < 4; ++px)
#pragma unroll
        for (int q = 0; q < CPL; ++q) acc[px][q] = bv[q];
#pragma unroll
      for (int r = 0; r < KS; ++r) {
        float xw[KS + 3];
#pragma unroll
        for (int t = 0; t < KS + 3; ++t) xw[t] = xs[warp + r][cw + t];
#pragma unroll
        for (int s = 0; s < KS; ++s)
#pragma unroll
          for (int px = 0; px < 4; ++px)
#pragma unroll
            for (int q = 0; q < CPL; ++q) acc[px][q] = fmaf(xw[px + s], wr[q][r * KS + s], acc[px][q]);
      }
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        const int w = w0 + cw + px;
        if (w < W) {
          float* dst = y + (((size_t)b * H + h) * W + w) * COUT;
#pragma unroll
          for (int q = 0; q < CPL; ++q) {
            if (cb + lane + 32 * q < COUT) dst[cb + lane + 32 * q] = acc[px][q];
            ssum[q] += acc[px][q];
            ssq[q] = fmaf(acc[px][q], acc[px][q], ssq[q]);
          }
        }
      }
    }
  }
  if (stats != nullptr) {
#pragma unroll
    for (int q = 0; q < CPL; ++q)
      if (cb + lane + 32 * q < COUT) { red[0][warp][cb + lane + 32 * q] = ssum[q]; red[1][warp][cb + lane + 32 * q] = ssq[q]; }
    __syncthreads();
    if (threadIdx.x < COUT) {
      double s = 0.0, q2 = 0.0;
#pragma unroll
      for (int t = 0; t < 8; ++t) { s += (double)red[0][t][threadIdx.x]; q2 += (double)red[1][t][threadIdx.x]; }
      atomicAdd(stats + threadIdx.x, s);
      atomicAdd(stats + COUT + threadIdx.x, q2);
    }
  }
}

// dw[n][r][s] partial over a chunk of pixels; partial layout [chunk][KS*KS + 1][COUT] (last row = bias gradient)
template <int KS, int COUT>
__global__ void __launch_bounds__(256)
conv_stem_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int B, int H, int W,
                       float* __restrict__ partial) {
  pdl_trigger();
  pdl_wait();
  constexpr int P = KS / 2, CPL = (COUT + 31) / 32, T = KS * KS;
  constexpr int TH = 8, TW = 32;
  __shared__ float xs[TH + KS - 1][TW + KS - 1 + 1];
  __shared__ float red[8][COUT];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tiles_w = ceil_div(W, TW), tiles_h = ceil_div(H, TH);
  const int n_tiles = B * tiles_w * tiles_h;
  float acc[CPL][T + 1];
#pragma unroll
  for (int q = 0; q < CPL; ++q)
#pragma unroll
    for (int t = 0; t <= T; ++t) acc[q][t] = 0.f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {   // persistent: one partial per CTA
    const int b = tile / (tiles_w * tiles_h);
    const int th = (tile / tiles_w) % tiles_h, tw = tile % tiles_w;
    const int h0 = th * TH, w0 = tw * TW;
    __syncthreads();
    for (int i = threadIdx.x; i < (TH + KS - 1) * (TW + KS - 1); i += 256) {
      const int rr = i / (TW + KS - 1), cc = i % (TW + KS - 1);
      const int h = h0 + rr - P, w = w0 + cc - P;
      xs[rr][cc] = (h >= 0 && h < H && w >= 0 && w < W) ? x[((size_t)b * H + h) * W + w] : 0.f;
    }
    __syncthreads();
    const int h = h0 + warp;
    if (h < H) {
      // dy for the next 4 pixels is fetched while the current 4 are being accumulated (the loop is otherwise exposed to
      // the full global-load latency once per iteration: only 8 warps are resident)
      auto load_d = [&](int cw, float (&d)[4][CPL]) {
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          const int w = w0 + cw + px;
          const bool ok = cw < TW && w < W;
          const float* src = dy + (((size_t)b * H + h) * W + (ok ? w : 0)) * COUT;
#pragma unroll
          for (int q = 0; q < CPL; ++q) d[px][q] = (ok && lane + 32 * q < COUT) ? src[lane + 32 * q] : 0.f;
        }
      };
      float dn[4][CPL];
      load_d(0, dn);
      for (int cw = 0; cw < TW && w0 + cw < W; cw += 4) {
        float d[4][CPL];
#pragma unroll
        for (int px = 0; px < 4; ++px)
#pragma unroll
          for (int q = 0; q < CPL; ++q) { d[px][q] = dn[px][q]; acc[q][T] += d[px][q]; }
        load_d(cw + 4, dn);
#pragma unroll
        for (int r = 0; r < KS; ++r) {
          float xw[KS + 3];
#pragma unroll
          for (int t = 0; t < KS + 3; ++t) xw[t] = xs[warp + r][cw + t];
#pragma unroll
          for (int s = 0; s < KS; ++s)
#pragma unroll
            for (int px = 0; px < 4; ++px)
#pragma unroll
              for (int q = 0; q < CPL; ++q) acc[q][r * KS + s] = fmaf(xw[px + s], d[px][q], acc[q][r * KS + s]);
        }
      }
    }
  }
  float* pbase = partial + (size_t)blockIdx.x * (T + 1) * COUT;
#pragma unroll
  for (int t = 0; t <= T; ++t) {   // unrolled so acc[][] stays in registers
    __syncthreads();
#pragma unroll
    for (int q = 0; q < CPL; ++q)
      if (lane + 32 * q < COUT) red[warp][lane + 32 * q] = acc[q][t];
    __syncthreads();
    if (threadIdx.x < COUT) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
      pbase[(size_t)t * COUT + threadIdx.x] = s;
    }
  }
}

void launch_wgrad_reduce(const float* partial, int n_splits, int R, int S, int Cin, int Cout, float* dw, float* db, pc_stream_t stream,
                         const float* bias_partial, int n_bias) {
  if (db == nullptr && bias_partial == nullptr && dw != nullptr && n_splits <= 8 && R * S <= 9 && Cin % 8 == 0 && Cout % 32 == 0 &&
      (long long)R * S * Cin * Cout >= 200000) {
    launch_pdl(conv_wgrad_reduce_t_kernel, dim3(Cout / 32, Cin / 8), dim3(256), 0, stream, partial, n_splits, R * S, Cin, Cout, dw);
    count_launch();
    return;
  }
  const long long total4 = ((long long)(R * S * Cin + 1) * Cout) >> 2;
  launch_pdl(conv_wgrad_reduce_kernel, dim3(ceil_div(total4, 32)), dim3(256), 0, stream, partial, n_splits, R, S, Cin, Cout, dw, db,
             bias_partial, n_bias);
  count_launch();
}

static inline XformDev to_dev(const PcInXform* xf) {
  XformDev d{nullptr, nullptr, nullptr, 0};
  if (xf != nullptr) { d.scale = xf->scale; d.shift = xf->shift; d.drop = xf->drop; d.relu = xf->relu; }
  return d;
}

static int check_geom(const char* fn, const PcConvGeom* g) {
  PC_REQUIRE(g != nullptr, PC_EINVAL, "%s: null geometry", fn);
  PC_REQUIRE(g->B > 0 && g->H > 0 && g->W > 0 && g->Cin > 0 && g->Cout > 0 && g->R > 0 && g->S > 0 && g->stride > 0 && g->pad >= 0,
             PC_EINVAL, "%s: bad geometry", fn);
  const int ho = (g->H + 2 * g->pad - g->R) / g->stride + 1, wo = (g->W + 2 * g->pad - g->S) / g->stride + 1;
  PC_REQUIRE(ho == g->Ho && wo == g->Wo && ho > 0 && wo > 0, PC_EINVAL, "%s: output %dx%d inconsistent with input %dx%d k%d s%d p%d", fn,
             g->Ho, g->Wo, g->H, g->W, g->R, g->stride, g->pad);
  return PC_OK;
}

static inline int stem_wgrad_ctas(const PcConvGeom* g) {
  const int tiles = g->B * ceil_div(g->H, 8) * ceil_div(g->W, 32);
  return tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs;
}

static int wgrad_splits(const PcConvGeom* g, int* rows_per_split) {
  const long long M = (long long)g->B * g->Ho * g->Wo;
  const int tiles = g->R * g->S * ceil_div(g->Cin, WG_TC) * ceil_div(g->Cout, WG_TN);
  int splits = ceil_div(4LL * kNumSMs, tiles);
  const int max_splits = ceil_div(M, 256);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int rps = ceil_div(M, splits);
  rps = ceil_div(rps, WG_BM) * WG_BM;
  splits = ceil_div(M, rps);
  *rows_per_split = rps;
  return splits;
}

}  // namespace pc

using namespace pc;

extern "C" int pc_pack_conv_weight(const float* w_oihw, int O, int I, int R, int S, float* wf, float* wd, pc_stream_t stream) {
  PC_REQUIRE(w_oihw && (wf || wd) && O > 0 && I > 0 && R > 0 && S > 0, PC_EINVAL, "pc_pack_conv_weight: bad arguments");
  const long long n = (long long)O * I * R * S;
  int grid = ceil_div(n, 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  pack_conv_weight_kernel<<<grid, 256, 0, stream>>>(w_oihw, O, I, R, S, wf, wd);
  PC_LAUNCH_CHECK("pack_conv_weight_kernel");
  return PC_OK;
}

// tensor-core path (conv_tc.cu); returns PC_EUNSUPPORTED when the shape is not covered.
extern "C" int pc_conv_fwd_tc(const float* x, const void* wp, const float* bias, const PcConvGeom* g, const PcInXform* xf,
                              float* y, double* stats, int prec, pc_stream_t stream);
extern "C" int pc_conv_dgrad_tc(const float* dy, const void* wp, const PcConvGeom* g, float* dx, int accumulate, int prec,
                                const float* dy_amax, int dy_presplit, pc_stream_t stream);

extern "C" int pc_conv_fwd(const float* x, const float* wf, const float* bias, const PcConvGeom* g, const PcInXform* xf,
                           float* y, double* stats, int prec, pc_stream_t stream) {
  int rc = check_geom("pc_conv_fwd", g);
  if (rc != PC_OK) return rc;
  PC_REQUIRE(x && wf && y, PC_EINVAL, "pc_conv_fwd: null pointer");
  if (g->Cin == 1) {
    // stem: wf is the OIHW weight itself ([Cout][R*S])
    PC_REQUIRE(g->stride == 1 && g->R == g->S && g->pad == g->R / 2, PC_EUNSUPPORTED, "pc_conv_fwd: stem conv must be stride 1, 'same' padding");
    PC_REQUIRE(xf == nullptr || (xf->scale == nullptr && xf->drop == nullptr && !xf->relu), PC_EUNSUPPORTED, "pc_conv_fwd: stem conv takes no input transform");
    const int tiles = g->B * ceil_div(g->H, 8) * ceil_div(g->W, 32);
    if (g->R == 3 && g->Cout == 16) launch_pdl((conv_stem_fwd_kernel<3, 16>), dim3(tiles), dim3(256), 0, stream, x, wf, bias, g->B, g->H, g->W, y, stats);
    else if (g->R == 3 && g->Cout == 32) launch_pdl((conv_stem_fwd_kernel<3, 32>), dim3(tiles), dim3(256), 0, stream, x, wf, bias, g->B, g->H, g->W, y, stats);
    else if (g->R == 3 && g->Cout == 64) launch_pdl((conv_stem_fwd_kernel<3, 64>), dim3(tiles), dim3(256), 0, stream, x, wf, bias, g->B, g->H, g->W, y, stats);
    else if (g->R == 7 && g->Cout == 16) launch_pdl((conv_stem_fwd_kernel<7, 16>), dim3(tiles), dim3(256), 0, stream, x, wf, bias, g->B, g->H, g->W, y, stats);
    else if (g->R == 7 && g->Cout == 32) launch_pdl((conv_stem_fwd_kernel<7, 32>), dim3(tiles), dim3(256), 0, stream, x, wf, bias, g->B, g->H, g->W, y, stats);
    else if (g->R == 7 && g->Cout == 64) launch_pdl((conv_stem_fwd_kernel<7, 64>), dim3(tiles), dim3(256), 0, stream, x, wf, bias, g->B, g->H, g->W, y, stats);
    else PC_REQUIRE(false, PC_EUNSUPPORTED, "pc_conv_fwd: stem conv k=%d Cout=%d not built (3|7 x 16|32|64)", g->R, g->Cout);
    PC_LAUNCH_CHECK("conv_stem_fwd_kernel");
    return PC_OK;
  }
  PC_REQUIRE(g->Cin % CG_BK == 0 && g->Cout % 4 == 0, PC_EUNSUPPORTED, "pc_conv_fwd: Cin=%d must be a multiple of 16 and Cout=%d of 4", g->Cin, g->Cout);
  if (prec != PC_PREC_FP32) {
    // tensor-core path: `wf` is the buffer produced by pc_pack_conv_weight_tc for this precision
    rc = pc_conv_fwd_tc(x, wf, bias, g, xf, y, stats, prec, stream);
    PC_REQUIRE(rc != PC_EUNSUPPORTED, PC_EUNSUPPORTED, "pc_conv_fwd: shape not covered by the tensor-core path (check pc_conv_tc_supported)");
    return rc;
  }
  PC_REQUIRE(xf == nullptr || !xf->presplit, PC_EUNSUPPORTED, "pc_conv_fwd: pre-split input planes need the FP16X2 tensor-core path");
  const long long M = (long long)g->B * g->Ho * g->Wo;
  const XformDev d = to_dev(xf);
  if (g->Cout <= 32) {
    dim3 grid(ceil_div(M, CG_BM), ceil_div(g->Cout, 32));
    conv_igemm_kernel<32, 0><<<grid, CG_THREADS, 0, stream>>>(x, wf, bias, *g, d, y, stats, 0);
  } else if (g->Cout <= 64) {
    dim3 grid(ceil_div(M, CG_BM), ceil_div(g->Cout, 64));
    conv_igemm_kernel<64, 0><<<grid, CG_THREADS, 0, stream>>>(x, wf, bias, *g, d, y, stats, 0);
  } else {
    dim3 grid(ceil_div(M, CG_BM), ceil_div(g->Cout, 128));
    conv_igemm_kernel<128, 0><<<grid, CG_THREADS, 0, stream>>>(x, wf, bias, *g, d, y, stats, 0);
  }
  PC_LAUNCH_CHECK("conv_igemm_kernel<fwd>");
  return PC_OK;
}

extern "C" int pc_conv_dgrad(const float* dy, const float* wd, const PcConvGeom* g, float* dx, int accumulate, int prec,
                             const float* dy_amax, int dy_presplit, pc_stream_t stream) {
  int rc = check_geom("pc_conv_dgrad", g);
  if (rc != PC_OK) return rc;
  PC_REQUIRE(dy && wd && dx, PC_EINVAL, "pc_conv_dgrad: null pointer");
  PC_REQUIRE(g->Cout % CG_BK == 0 && g->Cin % 4 == 0, PC_EUNSUPPORTED, "pc_conv_dgrad: Cout=%d must be a multiple of 16 and Cin=%d of 4", g->Cout, g->Cin);
  if (prec != PC_PREC_FP32) {
    rc = pc_conv_dgrad_tc(dy, wd, g, dx, accumulate, prec, dy_amax, dy_presplit, stream);
    PC_REQUIRE(rc != PC_EUNSUPPORTED, PC_EUNSUPPORTED, "pc_conv_dgrad: shape not covered by the tensor-core path (check pc_conv_tc_supported)");
    return rc;
  }
  PC_REQUIRE(!dy_presplit, PC_EUNSUPPORTED, "pc_conv_dgrad: pre-split dy needs the FP16X2 tensor-core path");
  const long long M = (long long)g->B * g->H * g->W;
  const XformDev d{nullptr, nullptr, nullptr, 0};
  if (g->Cin <= 32) {
    dim3 grid(ceil_div(M, CG_BM), ceil_div(g->Cin, 32));
    conv_igemm_kernel<32, 1><<<grid, CG_THREADS, 0, stream>>>(dy, wd, nullptr, *g, d, dx, nullptr, accumulate);
  } else if (g->Cin <= 64) {
    dim3 grid(ceil_div(M, CG_BM), ceil_div(g->Cin, 64));
    conv_igemm_kernel<64, 1><<<grid, CG_THREADS, 0, stream>>>(dy, wd, nullptr, *g, d, dx, nullptr, accumulate);
  } else {
    dim3 grid(ceil_div(M, CG_BM), ceil_div(g->Cin, 128));
    conv_igemm_kernel<128, 1><<<grid, CG_THREADS, 0, stream>>>(dy, wd, nullptr, *g, d, dx, nullptr, accumulate);
  }
  PC_LAUNCH_CHECK("conv_igemm_kernel<dgrad>");
  return PC_OK;
}

extern "C" int pc_conv_wgrad_tc_supported(const PcConvGeom* g);
extern "C" int pc_conv_wgrad_tc_stem_supported(const PcConvGeom* g);
extern "C" size_t pc_conv_wgrad_tc_workspace(const PcConvGeom* g);
extern "C" int pc_conv_wgrad_tc(const float* x, const float* dy, const PcConvGeom* g, const PcInXform* xf, float* dw_oihw, float* db,
                                void* workspace, size_t workspace_bytes, int prec, const float* dy_amax, int dy_presplit, pc_stream_t stream);

static size_t wgrad_workspace_simt(const PcConvGeom* g);

extern "C" int pc_conv_wgrad_halo_supported(const PcConvGeom* g);
extern "C" size_t pc_conv_wgrad_halo_workspace(const PcConvGeom* g);
extern "C" int pc_conv_wgrad_halo(const void* x_planes, const void* dy_planes, const PcConvGeom* g, float* dw_oihw, void* workspace,
                                  size_t workspace_bytes, const float* dy_amax, pc_stream_t stream);

extern "C" size_t pc_conv_wgrad_workspace(const PcConvGeom* g) {
  if (g == nullptr) return 0;
  size_t a = wgrad_workspace_simt(g);
  if (pc_conv_wgrad_halo_supported(g)) {
    const size_t h = pc_conv_wgrad_halo_workspace(g);
    if (h > a) a = h;
  }
  if ((g->Cin != 1 && pc_conv_wgrad_tc_supported(g)) || pc_conv_wgrad_tc_stem_supported(g)) {
    const size_t b = pc_conv_wgrad_tc_workspace(g);
    if (b > a) a = b;
  }
  return a;
}

static size_t wgrad_workspace_simt(const PcConvGeom* g) {
  if (g->Cin == 1) return (size_t)stem_wgrad_ctas(g) * (size_t)(g->R * g->S + 1) * g->Cout * sizeof(float);
  int rps;
  const int splits = wgrad_splits(g, &rps);
  return (size_t)splits * (size_t)(g->R * g->S * g->Cin + 1) * g->Cout * sizeof(float);
}

extern "C" int pc_conv_wgrad(const float* x, const float* dy, const PcConvGeom* g, const PcInXform* xf, float* dw_oihw,
                             float* db, void* workspace, size_t workspace_bytes, int prec, const float* dy_amax, int dy_presplit,
                             pc_stream_t stream) {
  int rc = check_geom("pc_conv_wgrad", g);
  if (rc != PC_OK) return rc;
  PC_REQUIRE(x && dy && dw_oihw && workspace, PC_EINVAL, "pc_conv_wgrad: null pointer");
  PC_REQUIRE(workspace_bytes >= pc_conv_wgrad_workspace(g), PC_EINVAL, "pc_conv_wgrad: workspace too small (%zu < %zu)", workspace_bytes,
             pc_conv_wgrad_workspace(g));
  // stride-1 3x3 layers with both operands in plane form and the bias gradient coming from the BatchNorm backward: the halo engine
  // (csrc/conv_halo_wgrad.cu: operands loaded once per position tile by TMA, taps from descriptors)
  if (prec == PC_PREC_FP16X2 && dy_presplit && xf != nullptr && xf->presplit && db == nullptr && dy_amax != nullptr && pc_conv_wgrad_halo_supported(g))
    return pc_conv_wgrad_halo(x, dy, g, dw_oihw, workspace, workspace_bytes, dy_amax, stream);
  // tensor-core path (TF32x3) for eligible layers whenever a tensor-core precision is requested
  if (prec != PC_PREC_FP32 && g->Cin != 1 && pc_conv_wgrad_tc_supported(g))
    return pc_conv_wgrad_tc(x, dy, g, xf, dw_oihw, db, workspace, workspace_bytes, prec, dy_amax, dy_presplit, stream);
  // single-channel stem on the tensor cores (taps as M rows: built for the 49 taps of cnn_deep's 7x7 stem; with the 9 taps of cnn_small's
  // 3x3 stem most of the 128-row tile is padding -- 59 us against the direct kernel below, so 3x3 stems stay on that one unless
  // PC_WGRAD_TC_STEM3=1)
  static const bool stem3_tc = [] { const char* e = getenv("PC_WGRAD_TC_STEM3"); return e != nullptr && e[0] == '1'; }();
  if (prec == PC_PREC_FP16X2 && xf == nullptr && pc_conv_wgrad_tc_stem_supported(g) && (g->R > 3 || stem3_tc))
    return pc_conv_wgrad_tc(x, dy, g, xf, dw_oihw, db, workspace, workspace_bytes, prec, dy_amax, 0, stream);
  PC_REQUIRE(!dy_presplit, PC_EUNSUPPORTED, "pc_conv_wgrad: pre-split dy needs the FP16X2 tensor-core path");
  PC_REQUIRE(xf == nullptr || !xf->presplit, PC_EUNSUPPORTED, "pc_conv_wgrad: pre-split input planes need the FP16X2 tensor-core path");
  float* partial = static_cast<float*>(workspace);
  int n_partials;
  if (g->Cin == 1) {
    PC_REQUIRE(g->stride == 1 && g->R == g->S && g->pad == g->R / 2, PC_EUNSUPPORTED, "pc_conv_wgrad: stem conv must be stride 1, 'same' padding");
    const int tiles = stem_wgrad_ctas(g);
    if (g->R == 3 && g->Cout == 16) launch_pdl((conv_stem_wgrad_kernel<3, 16>), dim3(tiles), dim3(256), 0, stream, x, dy, g->B, g->H, g->W, partial);
    else if (g->R == 3 && g->Cout == 32) launch_pdl((conv_stem_wgrad_kernel<3, 32>), dim3(tiles), dim3(256), 0, stream, x, dy, g->B, g->H, g->W, partial);
    else if (g->R == 3 && g->Cout == 64) launch_pdl((conv_stem_wgrad_kernel<3, 64>), dim3(tiles), dim3(256), 0, stream, x, dy, g->B, g->H, g->W, partial);
    else if (g->R == 7 && g->Cout == 16) launch_pdl((conv_stem_wgrad_kernel<7, 16>), dim3(tiles), dim3(256), 0, stream, x, dy, g->B, g->H, g->W, partial);
    else if (g->R == 7 && g->Cout == 32) launch_pdl((conv_stem_wgrad_kernel<7, 32>), dim3(tiles), dim3(256), 0, stream, x, dy, g->B, g->H, g->W, partial);
    else if (g->R == 7 && g->Cout == 64) launch_pdl((conv_stem_wgrad_kernel<7, 64>), dim3(tiles), dim3(256), 0, stream, x, dy, g->B, g->H, g->W, partial);
    else PC_REQUIRE(false, PC_EUNSUPPORTED, "pc_conv_wgrad: stem conv k=%d Cout=%d not built (3|7 x 16|32|64)", g->R, g->Cout);
    PC_LAUNCH_CHECK("conv_stem_wgrad_kernel");
    n_partials = tiles;
  } else {
    PC_REQUIRE(g->Cin % 4 == 0 && g->Cout % 4 == 0, PC_EUNSUPPORTED, "pc_conv_wgrad: channels must be multiples of 4");
    int rps;
    const int splits = wgrad_splits(g, &rps);
    dim3 grid(g->R * g->S * ceil_div(g->Cin, WG_TC), ceil_div(g->Cout, WG_TN), splits);
    conv_wgrad_kernel<<<grid, 256, 0, stream>>>(x, dy, *g, to_dev(xf), partial, splits, rps);
    PC_LAUNCH_CHECK("conv_wgrad_kernel");
    n_partials = splits;
  }
  const long long total4 = ((long long)(g->R * g->S * g->Cin + 1) * g->Cout) >> 2;
  int grid = ceil_div(total4, 32);
  launch_pdl(conv_wgrad_reduce_kernel, dim3(grid), dim3(256), 0, stream, partial, n_partials, g->R, g->S, g->Cin, g->Cout, dw_oihw, db,
             (const float*)nullptr, 0);
  PC_LAUNCH_CHECK("conv_wgrad_reduce_kernel");
  return PC_OK;
}
