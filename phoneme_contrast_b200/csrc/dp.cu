// Data-parallel exchange helpers: everything a rank does around the NCCL calls that is not the row-block SupCon
// kernels themselves. One all_gather moves embeddings AND labels: each local row is packed as D fp32 values followed by the two
// 32-bit halves of its int64 label (bit-cast, never converted); after the gather one launch splits the [N][D+2] buffer into the
// contiguous F [N][D] / labels [N] the loss kernels take. The loss value needs no collective of its own: the gathered per-row
// statistics (max, denominator, positive count, positive sum -- already exchanged for the backward) determine every row's loss
// (reference losses.py:69-82), so each rank reduces them in the same fixed order and obtains the same global mean.
#include "common.cuh"

namespace pc {

__global__ void __launch_bounds__(256) dp_pack_kernel(const float* __restrict__ emb, const int64_t* __restrict__ labels, int n, int D,
                                                      float* __restrict__ packed) {
  const int ld = D + 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * ld; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld), c = (int)(i - (long long)r * ld);
    float v;
    if (c < D) v = emb[(size_t)r * D + c];
    else {
      const unsigned long long bits = (unsigned long long)labels[r];
      v = __uint_as_float(c == D ? (uint32_t)(bits & 0xFFFFFFFFull) : (uint32_t)(bits >> 32));
    }
    packed[i] = v;
  }
}

__global__ void __launch_bounds__(256) dp_unpack_kernel(const float* __restrict__ packed, int N, int D, float* __restrict__ F,
                                                        int64_t* __restrict__ labels) {
  const int ld = D + 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)N * ld; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld), c = (int)(i - (long long)r * ld);
    if (c < D) F[(size_t)r * D + c] = packed[i];
    else if (c == D) {
      const unsigned long long lo = __float_as_uint(packed[i]), hi = __float_as_uint(packed[i + 1]);
      labels[r] = (int64_t)(lo | (hi << 32));
    }
  }
}

// out[0] = scale * sum_i -(T/T_base) * (spos_i - npos_i * log den_i) / max(npos_i, 1), fixed summation order (fp64)
__global__ void __launch_bounds__(256) loss_from_stats_kernel(const float* __restrict__ stats, int N, float t_ratio, float scale,
                                                              float* __restrict__ out) {
  __shared__ double sh[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < N; i += 256) {
    const float4 st = *reinterpret_cast<const float4*>(stats + (size_t)i * 4);      // m, den (+1e-6), npos, spos
    const float nn = st.z == 0.f ? 1.f : st.z;
    s += (double)(-t_ratio * (st.w - st.z * logf(st.y)) / nn);
  }
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

using namespace pc;

extern "C" int pc_dp_pack(const float* emb, const int64_t* labels, int n, int D, float* packed, pc_stream_t stream) {
  PC_REQUIRE(emb && labels && packed && n > 0 && D > 0, PC_EINVAL, "pc_dp_pack: bad arguments");
  int grid = ceil_div((long long)n * (D + 2), 256);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  dp_pack_kernel<<<grid, 256, 0, stream>>>(emb, labels, n, D, packed);
  PC_LAUNCH_CHECK("dp_pack_kernel");
  return PC_OK;
}

extern "C" int pc_dp_unpack(const float* packed, int N, int D, float* F, int64_t* labels, pc_stream_t stream) {
  PC_REQUIRE(packed && F && labels && N > 0 && D > 0, PC_EINVAL, "pc_dp_unpack: bad arguments");
  int grid = ceil_div((long long)N * (D + 2), 256);
  if (grid > kNumSMs * 4) grid = kNumSMs * 4;
  dp_unpack_kernel<<<grid, 256, 0, stream>>>(packed, N, D, F, labels);
  PC_LAUNCH_CHECK("dp_unpack_kernel");
  return PC_OK;
}

extern "C" int pc_supcon_loss_from_stats(const float* stats, int N, float temperature, float base_temperature, float scale, float* out,
                                         pc_stream_t stream) {
  PC_REQUIRE(stats && out && N > 0 && base_temperature > 0.f, PC_EINVAL, "pc_supcon_loss_from_stats: bad arguments");
  loss_from_stats_kernel<<<1, 256, 0, stream>>>(stats, N, temperature / base_temperature, scale, out);
  PC_LAUNCH_CHECK("loss_from_stats_kernel");
  return PC_OK;
}
