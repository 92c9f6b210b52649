// Stand-alone hardware probe (round 2): settles three questions the halo-resident convolution engine depends on.
//   T1  tcgen05.mma, K-major SWIZZLE_128B A operand whose descriptor start address is advanced by an ARBITRARY number of
//       128-byte rows (not a multiple of the 1024-byte swizzle atom): is the swizzle taken from absolute shared-memory address
//       bits (base_offset field = 0 works), or does the descriptor's base_offset field have to carry (addr >> 7) & 7 ?
//   T2  cp.async.bulk.tensor (tiled, 4-D, SWIZZLE_128B): box wider than the tensor (W+1 columns starting at w = -1, rows starting
//       at h = -1), zero fill of out-of-bounds elements, destination only 128-byte aligned: layout and swizzle phase in smem.
//   T3  cp.async.bulk.tensor im2col mode (4-D): stride-2 3x3 pad-1 and 1x1 stride-2 gathers of 128 consecutive output pixels.
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/probe profiles/tools/probe_umma_tma.cu && /tmp/probe
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {     // never hangs the box: gives up after ~1 s
  for (long long i = 0; i < 200000000LL; ++i)
    if (mbar_try_wait(bar, parity)) return true;
  return false;
}

// ------------------------------------------------------------------------------------------------------------ T1
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7u) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
// A_rows [R][64] fp16 (row-major), Bm [64][64] fp16 (n-major rows of k). out[variant][shift_idx][128][64] fp32
__global__ void __launch_bounds__(128) t1_kernel(const __half* __restrict__ A_rows, int R, const __half* __restrict__ Bm,
                                                 const int* __restrict__ shifts, int n_shifts, float* __restrict__ out, int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  unsigned char* a_tile = smem;                       // R rows x 128 B, swizzled by absolute address
  unsigned char* b_tile = smem + (size_t)R * 128;     // 64 rows x 128 B (R is a multiple of 8 -> 1024-aligned)
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < R * 8; i += 128) {
    const int r = i >> 3, j = i & 7;
    *reinterpret_cast<uint4*>(a_tile + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A_rows + (size_t)r * 64 + j * 8);
  }
  for (int i = tid; i < 64 * 8; i += 128) {
    const int r = i >> 3, j = i & 7;
    *reinterpret_cast<uint4*>(b_tile + r * 128 + ((j ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(Bm + (size_t)r * 64 + j * 8);
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);   // f16 x f16 -> f32, M 128, N 64, K-major both
  uint32_t phase = 0;
  for (int variant = 0; variant < 2; ++variant) {
    for (int si = 0; si < n_shifts; ++si) {
      if (tid == 0) {
        const uint32_t a_addr = smem_u32(a_tile) + (uint32_t)shifts[si] * 128u;
        const uint32_t bo = variant == 0 ? 0u : ((a_addr >> 7) & 7u);
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t da = desc_sw128(a_addr, bo) + (uint64_t)(kk * 2);
          const uint64_t db = desc_sw128(smem_u32(b_tile), 0) + (uint64_t)(kk * 2);
          const uint32_t acc = kk > 0;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      if (!mbar_wait_bounded(&bar, phase)) { if (tid == 0) atomicExch(status, 100 + si); }
      phase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float* o = out + (((size_t)variant * n_shifts + si) * 128 + tid) * 64;
      for (int c0 = 0; c0 < 64; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 8; ++q) o[c0 + q] = __uint_as_float(v[q]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}

// ------------------------------------------------------------------------------------------------------------ T2 / T3
// One thread issues `n_loads` TMA loads (tiled 4-D, or im2col 4-D with offsets) into smem at dst_off[i]; the CTA then dumps `dump_bytes`
// of raw shared memory (from the 1024-aligned base) to global memory together with the low address bits of the base.
struct LoadCmd { int c, w, h, n; int off_w, off_h; int dst_off; int bytes; };
__global__ void __launch_bounds__(128) tma_kernel(const __grid_constant__ CUtensorMap map, const LoadCmd* __restrict__ cmds, int n_loads, int im2col,
                                                  unsigned char* __restrict__ dump, int dump_bytes, int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char raw[];
  unsigned char* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  for (int i = tid; i < dump_bytes / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0xDEADBEEFu;
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    uint32_t total = 0;
    for (int i = 0; i < n_loads; ++i) total += (uint32_t)cmds[i].bytes;
    mbar_expect_tx(&bar, total);
    for (int i = 0; i < n_loads; ++i) {
      const LoadCmd c = cmds[i];
      const uint32_t dst = smem_u32(smem) + (uint32_t)c.dst_off;
      if (!im2col) {
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(dst), "l"(&map), "r"(smem_u32(&bar)), "r"(c.c), "r"(c.w), "r"(c.h), "r"(c.n) : "memory");
      } else {
        const uint16_t ow = (uint16_t)c.off_w, oh = (uint16_t)c.off_h;
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                     ::"r"(dst), "l"(&map), "r"(smem_u32(&bar)), "r"(c.c), "r"(c.w), "r"(c.h), "r"(c.n), "h"(ow), "h"(oh) : "memory");
      }
    }
  }
  if (!mbar_wait_bounded(&bar, 0)) { if (tid == 0) atomicExch(status, 200); }
  __syncthreads();
  for (int i = tid; i < dump_bytes / 4; i += 128) reinterpret_cast<uint32_t*>(dump)[i] = reinterpret_cast<uint32_t*>(smem)[i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);
static void* entry(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !fn) { printf("driver entry point %s unavailable\n", name); exit(3); }
  return fn;
}

static float h2f(__half h) { return __half2float(h); }

// value of element (row, k) in a dumped K-major SW128 tile whose row 0 sits at byte offset `off` from the 1024-aligned base
static float tile_at(const std::vector<unsigned char>& d, int off, int row, int k) {
  const int rb = off + row * 128;                    // absolute-address swizzle: chunk ^= (addr >> 7) & 7
  const int chunk = (k >> 3) ^ ((rb >> 7) & 7);
  __half h;
  memcpy(&h, &d[rb + chunk * 16 + (k & 7) * 2], 2);
  return h2f(h);
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  int* status;
  CK(cudaMalloc(&status, 4));
  CK(cudaMemset(status, 0, 4));
  int fails = 0;
  // ------------------------------------------------------------------ T1
  {
    const int R = 256;
    std::vector<__half> A(R * 64), Bm(64 * 64);
    srand(1);
    for (auto& v : A) v = __float2half((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : Bm) v = __float2half((rand() % 2001 - 1000) / 1000.f);
    const int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 27, 52, 53, 100};
    const int ns = sizeof(shifts) / sizeof(int);
    __half *dA, *dB; int* dS; float* dO;
    CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, Bm.size() * 2)); CK(cudaMalloc(&dS, sizeof(shifts)));
    CK(cudaMalloc(&dO, (size_t)2 * ns * 128 * 64 * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bm.data(), Bm.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dS, shifts, sizeof(shifts), cudaMemcpyHostToDevice));
    const int smem = R * 128 + 64 * 128 + 1024;
    CK(cudaFuncSetAttribute(t1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    t1_kernel<<<1, 128, smem>>>(dA, R, dB, dS, ns, dO, status);
    CK(cudaDeviceSynchronize());
    std::vector<float> O((size_t)2 * ns * 128 * 64);
    CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
    for (int variant = 0; variant < 2; ++variant)
      for (int si = 0; si < ns; ++si) {
        double maxerr = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) ref += (double)h2f(A[(size_t)(shifts[si] + m) * 64 + k]) * h2f(Bm[n * 64 + k]);
            maxerr = fmax(maxerr, fabs(ref - O[(((size_t)variant * ns + si) * 128 + m) * 64 + n]));
          }
        printf("T1 base_offset=%s shift=%3d rows: max|err| = %.3e %s\n", variant ? "(addr>>7)&7" : "0          ", shifts[si], maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
      }
  }
  // ------------------------------------------------------------------ T2: tiled 4-D with halo zero fill
  EncodeTiledFn enc_tiled = (EncodeTiledFn)entry("cuTensorMapEncodeTiled");
  EncodeIm2colFn enc_im2col = (EncodeIm2colFn)entry("cuTensorMapEncodeIm2col");
  {
    const int B = 3, H = 5, W = 13, C = 128;
    std::vector<__half> X((size_t)B * H * W * C);
    for (size_t i = 0; i < X.size(); ++i) X[i] = __float2half((float)((i * 7919) % 2039) / 64.f + 1.f);   // never zero
    __half* dX; CK(cudaMalloc(&dX, X.size() * 2 + (1 << 17))); CK(cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap map;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(W + 1), 3, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc_tiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dX, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("T2 encode tiled (box W+1 = %d > W = %d): CUresult %d\n", W + 1, W, (int)r);
    if (r == CUDA_SUCCESS) {
      // two boxes: rows h = -1..1 of image 1 at dst 640 (128-aligned only), rows h = 2..4 at dst 640 + 3*14*128, third box: b = 3 (all OOB)
      LoadCmd cmds[3] = {{64, -1, -1, 1, 0, 0, 640, 3 * 14 * 128}, {64, -1, 2, 1, 0, 0, 640 + 3 * 14 * 128, 3 * 14 * 128}, {0, -1, -1, 3, 0, 0, 640 + 6 * 14 * 128, 3 * 14 * 128}};
      LoadCmd* dC; CK(cudaMalloc(&dC, sizeof(cmds))); CK(cudaMemcpy(dC, cmds, sizeof(cmds), cudaMemcpyHostToDevice));
      const int dump_bytes = 20 * 1024;
      unsigned char* dD; CK(cudaMalloc(&dD, dump_bytes));
      CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      tma_kernel<<<1, 128, 64 * 1024>>>(map, dC, 3, 0, dD, dump_bytes, status);
      CK(cudaDeviceSynchronize());
      std::vector<unsigned char> D(dump_bytes);
      CK(cudaMemcpy(D.data(), dD, dump_bytes, cudaMemcpyDeviceToHost));
      int bad = 0, checked = 0;
      for (int box_i = 0; box_i < 3; ++box_i)
        for (int hh = 0; hh < 3; ++hh)
          for (int ww = 0; ww < W + 1; ++ww)
            for (int k = 0; k < 64; ++k) {
              const int b = box_i == 2 ? 3 : 1, h = cmds[box_i].h + hh, w = -1 + ww, c = cmds[box_i].c + k;
              const float ref = (b < B && h >= 0 && h < H && w >= 0 && w < W) ? h2f(X[(((size_t)b * H + h) * W + w) * C + c]) : 0.f;
              const float got = tile_at(D, cmds[box_i].dst_off, hh * (W + 1) + ww, k);
              ++checked;
              if (got != ref) { if (bad < 5) printf("   T2 mismatch box %d hh %d ww %d k %d: got %g want %g\n", box_i, hh, ww, k, got, ref); ++bad; }
            }
      printf("T2 tiled halo boxes into a 128B-aligned destination, absolute-address swizzle: %d / %d mismatches %s\n", bad, checked, bad ? "MISMATCH" : "OK");
      fails += bad != 0;
    } else fails++;
  }
  // ------------------------------------------------------------------ T3: im2col mode
  for (int cfg = 0; cfg < 3; ++cfg) {
    const int B = 3, H = 10, W = 26, C = 128;
    const int R = cfg == 1 ? 1 : 3, pad = cfg == 1 ? 0 : 1, stride = cfg == 2 ? 1 : 2;
    const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - R) / stride + 1;
    std::vector<__half> X((size_t)B * H * W * C);
    for (size_t i = 0; i < X.size(); ++i) X[i] = __float2half((float)((i * 104729) % 2039) / 64.f + 1.f);
    __half* dX; CK(cudaMalloc(&dX, X.size() * 2 + (1 << 17))); CK(cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap map;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    int lower[2] = {-pad, -pad}, upper[2] = {pad - (R - 1), pad - (R - 1)};
    cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = enc_im2col(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dX, dims, strides, lower, upper, 64, 128, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("T3 cfg %d (R %d pad %d stride %d, Ho %d Wo %d) encode im2col: CUresult %d\n", cfg, R, pad, stride, Ho, Wo, (int)r);
    if (r != CUDA_SUCCESS) { fails++; continue; }
    const int m0s[3] = {0, 100, B * Ho * Wo - 60};        // tile starts: image start, mid-row, last partial tile (runs off the end)
    for (int t = 0; t < 3; ++t) {
      const int m0 = m0s[t];
      const int taps[3][2] = {{0, 0}, {R - 1, R - 1}, {R / 2, 0}};
      for (int ti = 0; ti < (R == 1 ? 1 : 3); ++ti) {
        const int tr = taps[ti][0], ts = taps[ti][1];
        const int q0 = m0 % Wo, p0 = (m0 / Wo) % Ho, n0 = m0 / (Wo * Ho);
        LoadCmd cmd = {64, q0 * stride - pad, p0 * stride - pad, n0, ts, tr, 0, 128 * 128};
        LoadCmd* dC; CK(cudaMalloc(&dC, sizeof(cmd))); CK(cudaMemcpy(dC, &cmd, sizeof(cmd), cudaMemcpyHostToDevice));
        const int dump_bytes = 16 * 1024;
        unsigned char* dD; CK(cudaMalloc(&dD, dump_bytes));
        tma_kernel<<<1, 128, 64 * 1024>>>(map, dC, 1, 1, dD, dump_bytes, status);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned char> D(dump_bytes);
        CK(cudaMemcpy(D.data(), dD, dump_bytes, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int i = 0; i < 128; ++i) {
          const int m = m0 + i, q = m % Wo, p = (m / Wo) % Ho, n = m / (Wo * Ho);
          const int h = p * stride - pad + tr, w = q * stride - pad + ts;
          for (int k = 0; k < 64; ++k) {
            const float ref = (n < B && h >= 0 && h < H && w >= 0 && w < W) ? h2f(X[(((size_t)n * H + h) * W + w) * C + 64 + k]) : 0.f;
            const float got = tile_at(D, 0, i, k);
            if (got != ref) { if (bad < 3) printf("   T3 mismatch row %d (n %d p %d q %d) k %d: got %g want %g\n", i, n, p, q, k, got, ref); ++bad; }
          }
        }
        printf("T3 cfg %d tile m0 = %4d tap (r %d, s %d): %d mismatches %s\n", cfg, m0, tr, ts, bad, bad ? "MISMATCH" : "OK");
        fails += bad != 0;
      }
    }
  }
  int st = 0;
  CK(cudaMemcpy(&st, status, 4, cudaMemcpyDeviceToHost));
  printf("status word %d (0 = no barrier time-outs); failing groups %d\n", st, fails);
  return 0;
}
