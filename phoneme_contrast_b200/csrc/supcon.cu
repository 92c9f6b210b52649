// Supervised contrastive loss, forward and backward, flash-style: the N x N logits live only in
// registers / shared memory. fp32 SIMT path (exact fp32 parity; used for every N, and the only path
// at the reference's batch sizes N = 64..512 where the op is launch-latency bound).
//
// Reference arithmetic: src/training/losses.py:49-84:
//   z_ij = f_i.f_j / T ; m_i = max_j z_ij (diagonal included) ; den_i = sum_{j!=i} exp(z_ij-m_i) + 1e-6
//   P_ij = [y_i==y_j, j!=i] (or user mask * (1-I)) ; n_i = sum_j P_ij
//   l_i = -(T/Tb) * (sum_j P_ij (z_ij-m_i) - n_i log den_i) / (n_i==0 ? 1 : n_i)
// Backward (m_i detached, losses.py:65):
//   G_ij = -c (P_ij/nn_i - h_i e^{z_ij-m_i} O_ij / den_i),  dF_i = (1/T) sum_j (G_ij + G_ji) f_j
#include "common.cuh"

namespace pc {

constexpr int SC_BM = 64, SC_BN = 64, SC_BK = 32, SC_THREADS = 256;
constexpr int SC_DC = 128;  // d-columns of dF produced per CTA pass in the backward

// S tile (64x64) += Fi[64 x D] * Fj[64 x D]^T, 4x4 per thread (rows ty*4.., cols tx*4..)
__device__ __forceinline__ void sc_tile_gemm(const float* __restrict__ F, int N, int D, int i0, int j0,
                                             float (*As)[SC_BM + 4], float (*Bs)[SC_BN + 4], float acc[4][4]) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  // loader mapping: row = tid & 63, k-quad = tid >> 6 (0..3) -> each thread loads 2 float4 per operand per chunk
  const int lr = tid & 63, lq = tid >> 6;
  for (int k0 = 0; k0 < D; k0 += SC_BK) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int kq = lq + 4 * h;  // 0..7 -> k offset kq*4
      const int k = k0 + kq * 4;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      if (k < D) {
        if (i0 + lr < N) va = *reinterpret_cast<const float4*>(F + (size_t)(i0 + lr) * D + k);
        if (j0 + lr < N) vb = *reinterpret_cast<const float4*>(F + (size_t)(j0 + lr) * D + k);
      }
      As[kq * 4 + 0][lr] = va.x; As[kq * 4 + 1][lr] = va.y; As[kq * 4 + 2][lr] = va.z; As[kq * 4 + 3][lr] = va.w;
      Bs[kq * 4 + 0][lr] = vb.x; Bs[kq * 4 + 1][lr] = vb.y; Bs[kq * 4 + 2][lr] = vb.z; Bs[kq * 4 + 3][lr] = vb.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SC_BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ float group16_max(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float group16_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(SC_THREADS)
supcon_fwd_kernel(const float* __restrict__ F, const int64_t* __restrict__ labels, const float* __restrict__ mask,
                  int N, int D, int row0, int nrows, float invT, float t_ratio, float* __restrict__ stats,
                  float* __restrict__ row_loss) {
  __shared__ __align__(16) float As[SC_BK][SC_BM + 4];
  __shared__ __align__(16) float Bs[SC_BK][SC_BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = row0 + blockIdx.x * SC_BM;
  const int row_end = row0 + nrows;

  float m[4], den[4], npos[4], sraw[4];
  int64_t yi[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    m[r] = -INFINITY; den[r] = 0.f; npos[r] = 0.f; sraw[r] = 0.f;
    const int i = i0 + ty * 4 + r;
    yi[r] = (labels != nullptr && i < N) ? labels[i] : 0;
  }
  for (int j0 = 0; j0 < N; j0 += SC_BN) {
    float acc[4][4];
    sc_tile_gemm(F, N, D, i0, j0, As, Bs, acc);
    int64_t yj[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c;
      yj[c] = (labels != nullptr && j < N) ? labels[j] : 0;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty * 4 + r;
      float z[4], tmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + tx * 4 + c;
        z[c] = acc[r][c] * invT;
        if (j < N) tmax = fmaxf(tmax, z[c]);
      }
      tmax = group16_max(tmax);
      const float m_new = fmaxf(m[r], tmax);
      float dsum = 0.f, psum = 0.f, zsum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + tx * 4 + c;
        if (j < N && j != i && i < N) {
          dsum += expf(z[c] - m_new);
          const float p = (mask != nullptr) ? mask[(size_t)i * N + j] : (yi[r] == yj[c] ? 1.f : 0.f);
          psum += p;
          zsum = fmaf(p, z[c], zsum);
        }
      }
      dsum = group16_sum(dsum);
      psum = group16_sum(psum);
      zsum = group16_sum(zsum);
      den[r] = den[r] * expf(m[r] - m_new) + dsum;   // exp(-inf) = 0 on the first tile
      m[r] = m_new;
      npos[r] += psum;
      sraw[r] += zsum;
    }
  }
  if (tx == 0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty * 4 + r;
      if (i < row_end && i < N) {
        const float d = den[r] + 1e-6f;
        const float spos = sraw[r] - npos[r] * m[r];
        const float nn = (npos[r] == 0.f) ? 1.f : npos[r];
        float* st = stats + (size_t)(i - row0) * 4;
        st[0] = m[r]; st[1] = d; st[2] = npos[r]; st[3] = spos;
        row_loss[i - row0] = -t_ratio * (spos - npos[r] * logf(d)) / nn;
      }
    }
  }
}

// dF rows [i0, i0+64) x d-columns [dc0, dc0+128)
__global__ void __launch_bounds__(SC_THREADS)
supcon_bwd_kernel(const float* __restrict__ F, const int64_t* __restrict__ labels, const float* __restrict__ mask,
                  int N, int D, int row0, int nrows, float invT, float coef, const float* __restrict__ grad_scale,
                  const float* __restrict__ stats_all, float* __restrict__ dF) {
  extern __shared__ __align__(16) float smem[];
  float (*As)[SC_BM + 4] = reinterpret_cast<float (*)[SC_BM + 4]>(smem);
  float (*Bs)[SC_BN + 4] = reinterpret_cast<float (*)[SC_BN + 4]>(smem + SC_BK * (SC_BM + 4));
  float (*Ws)[SC_BN + 1] = reinterpret_cast<float (*)[SC_BN + 1]>(smem + 2 * SC_BK * (SC_BM + 4));
  float (*Fj)[SC_DC] = reinterpret_cast<float (*)[SC_DC]>(smem + 2 * SC_BK * (SC_BM + 4) + SC_BM * (SC_BN + 1));

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = row0 + blockIdx.x * SC_BM;
  const int dc0 = blockIdx.y * SC_DC;
  const int row_end = row0 + nrows;
  const float c = coef * (grad_scale != nullptr ? grad_scale[0] : 1.f);

  // per-row constants of my 4 rows (for the W tile) -- rows ty*4+r
  float mi[4], inv_den_i[4], inv_n_i[4];
  int64_t yi[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i < N) {
      const float* st = stats_all + (size_t)i * 4;
      mi[r] = st[0];
      const float h = st[2] != 0.f ? 1.f : 0.f;
      inv_den_i[r] = h / st[1];
      inv_n_i[r] = st[2] != 0.f ? 1.f / st[2] : 0.f;   // P_ij = 0 whenever n_i = 0 (labels); with a float mask nn=1 but sum P = 0
      if (mask != nullptr && st[2] == 0.f) inv_n_i[r] = 1.f;
      yi[r] = labels != nullptr ? labels[i] : 0;
    } else {
      mi[r] = 0.f; inv_den_i[r] = 0.f; inv_n_i[r] = 0.f; yi[r] = 0;
    }
  }
  // second GEMM: thread computes rows ty*4..+4 x dcols tx*8..+8
  float out[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int d = 0; d < 8; ++d) out[r][d] = 0.f;

  for (int j0 = 0; j0 < N; j0 += SC_BN) {
    float acc[4][4];
    sc_tile_gemm(F, N, D, i0, j0, As, Bs, acc);   // ends with __syncthreads()
    // stage Fj [64 x 128] natural layout
    for (int q = tid; q < SC_BN * (SC_DC / 4); q += SC_THREADS) {
      const int jr = q / (SC_DC / 4), dq = q % (SC_DC / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + jr < N && dc0 + dq * 4 < D) v = *reinterpret_cast<const float4*>(F + (size_t)(j0 + jr) * D + dc0 + dq * 4);
      *reinterpret_cast<float4*>(&Fj[jr][dq * 4]) = v;
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const int j = j0 + tx * 4 + cc;
      float mj = 0.f, inv_den_j = 0.f, inv_n_j = 0.f;
      int64_t yj = 0;
      if (j < N) {
        const float* st = stats_all + (size_t)j * 4;
        mj = st[0];
        inv_den_j = (st[2] != 0.f ? 1.f : 0.f) / st[1];
        inv_n_j = st[2] != 0.f ? 1.f / st[2] : (mask != nullptr ? 1.f : 0.f);
        yj = labels != nullptr ? labels[j] : 0;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty * 4 + r;
        float w = 0.f;
        if (i < N && j < N && i != j) {
          const float z = acc[r][cc] * invT;
          float pij, pji;
          if (mask != nullptr) {
            pij = mask[(size_t)i * N + j];
            pji = mask[(size_t)j * N + i];
          } else {
            pij = pji = (yi[r] == yj) ? 1.f : 0.f;
          }
          const float e = expf(z - mi[r]) * inv_den_i[r] + expf(z - mj) * inv_den_j;
          w = -c * (pij * inv_n_i[r] + pji * inv_n_j - e);
        }
        Ws[ty * 4 + r][tx * 4 + cc] = w;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < SC_BN; ++j) {
      const float4 b0 = *reinterpret_cast<const float4*>(&Fj[j][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Fj[j][tx * 8 + 4]);
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float a = Ws[ty * 4 + r][j];
#pragma unroll
        for (int d = 0; d < 8; ++d) out[r][d] = fmaf(a, bv[d], out[r][d]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i < row_end && i < N) {
#pragma unroll
      for (int d = 0; d < 8; d += 4) {
        const int dc = dc0 + tx * 8 + d;
        if (dc < D) {
          float4 v = make_float4(out[r][d] * invT, out[r][d + 1] * invT, out[r][d + 2] * invT, out[r][d + 3] * invT);
          *reinterpret_cast<float4*>(dF + (size_t)(i - row0) * D + dc) = v;
        }
      }
    }
  }
}

// out[0] = scale * sum x[0..n) in a fixed order (single CTA, fp64 accumulation)
__global__ void __launch_bounds__(256) sum_scaled_kernel(const float* __restrict__ x, int n, float scale, float* __restrict__ out) {
  __shared__ double sh[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += (double)x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    out[0] = (float)(t * (double)scale);
  }
}

}  // namespace pc

extern "C" int pc_supcon_fwd(const float* F, const int64_t* labels, const float* mask, int N, int D, int row0, int nrows,
                             float temperature, float base_temperature, float* stats, float* row_loss, pc_stream_t stream) {
  using namespace pc;
  PC_REQUIRE(N > 1, PC_EINVAL, "Batch size must be greater than 1 for contrastive loss");  // losses.py:44-45
  PC_REQUIRE(D > 0 && D % 4 == 0, PC_EUNSUPPORTED, "pc_supcon_fwd: embedding dim %d must be a multiple of 4", D);
  PC_REQUIRE(F && stats && row_loss && (labels || mask), PC_EINVAL, "pc_supcon_fwd: null pointer");
  PC_REQUIRE(row0 >= 0 && nrows > 0 && row0 + nrows <= N, PC_EINVAL, "pc_supcon_fwd: bad row block [%d,+%d) of %d", row0, nrows, N);
  PC_REQUIRE(temperature > 0.f && base_temperature > 0.f, PC_EINVAL, "pc_supcon_fwd: temperatures must be positive");
  dim3 grid(ceil_div(nrows, SC_BM));
  supcon_fwd_kernel<<<grid, SC_THREADS, 0, stream>>>(F, labels, mask, N, D, row0, nrows, 1.0f / temperature,
                                                      temperature / base_temperature, stats, row_loss);
  PC_LAUNCH_CHECK("supcon_fwd_kernel");
  return PC_OK;
}

extern "C" int pc_sum_scaled(const float* x, int n, float scale, float* out, pc_stream_t stream) {
  using namespace pc;
  PC_REQUIRE(x && out && n > 0, PC_EINVAL, "pc_sum_scaled: bad arguments");
  sum_scaled_kernel<<<1, 256, 0, stream>>>(x, n, scale, out);
  PC_LAUNCH_CHECK("sum_scaled_kernel");
  return PC_OK;
}

extern "C" int pc_supcon_bwd(const float* F, const int64_t* labels, const float* mask, int N, int D, int row0, int nrows,
                             float temperature, float coef, const float* grad_scale, const float* stats_all, float* dF,
                             pc_stream_t stream) {
  using namespace pc;
  PC_REQUIRE(N > 1, PC_EINVAL, "Batch size must be greater than 1 for contrastive loss");
  PC_REQUIRE(D > 0 && D % 4 == 0, PC_EUNSUPPORTED, "pc_supcon_bwd: embedding dim %d must be a multiple of 4", D);
  PC_REQUIRE(F && stats_all && dF && (labels || mask), PC_EINVAL, "pc_supcon_bwd: null pointer");
  PC_REQUIRE(row0 >= 0 && nrows > 0 && row0 + nrows <= N, PC_EINVAL, "pc_supcon_bwd: bad row block");
  static const size_t smem = sizeof(float) * (2 * SC_BK * (SC_BM + 4) + SC_BM * (SC_BN + 1) + SC_BN * SC_DC);
  static bool attr_set = false;
  if (!attr_set) {
    PC_CUDA(cudaFuncSetAttribute(supcon_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid(ceil_div(nrows, SC_BM), ceil_div(D, SC_DC));
  supcon_bwd_kernel<<<grid, SC_THREADS, smem, stream>>>(F, labels, mask, N, D, row0, nrows, 1.0f / temperature, coef,
                                                         grad_scale, stats_all, dF);
  PC_LAUNCH_CHECK("supcon_bwd_kernel");
  return PC_OK;
}
