// Data-parallel exchanges over NVLink / NVSwitch PEER MEMORY (new work, the reference is single-process).
//
// Every rank owns one "peer region" (a cudaMalloc'ed block exported with cudaIpcGetMemHandle and mapped by all other ranks of the
// node), laid out identically on all ranks: [flags | gathered embeddings F | gathered labels | gathered row statistics | flat gradient bucket].
// The three exchanges of a training step are then plain kernels that load / store through the mapped pointers -- no NCCL call, no
// host involvement -- so that the WHOLE data-parallel step (forward, loss, backward, exchanges, clip + Adam) is ONE captured CUDA graph:
//
//   pc_dp_gather_peer    store the local [n][D] embeddings and [n] labels straight into rows [row0, row0 + n) of EVERY rank's gathered
//                        F / label buffers (the all_gather is the producer's store loop; nothing to pack or unpack)
//   pc_peer_bcast        same for the [n][4] SupCon row statistics (or any small block)
//   pc_peer_allreduce    two-shot sum of the flat gradient bucket: rank r loads slice r from every rank (fixed rank order, so the
//                        result is bit-identical everywhere and independent of timing), adds, and stores the sum back into
//                        slice r of EVERY rank
//   pc_peer_barrier      flag barrier between the ranks (one 32-bit slot per (channel, source rank) in every region): arrive =
//                        fence.sys + st.release.sys of the channel's epoch into every peer's slot, wait = ld.acquire.sys spin on
//                        the own slots. Two independent channels, so that the tail of the gradient bucket can be exchanged on a side
//                        stream while the main stream continues with the backward.
// A barrier that does not complete within `timeout_ms` (a peer died) raises a sticky error word instead of hanging the GPU; every
// later barrier then returns at once, and the host reads the word with pc_peer_error.
#include <cuda.h>
#include <string.h>

#include "common.cuh"

namespace pc {
namespace peer {

constexpr int MAX_RANKS = 16;

struct Bases {
  unsigned long long p[MAX_RANKS];
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// flags: [2 channels][MAX_RANKS] arrival slots, then [2] epochs, then the error word (all 32-bit, at flag_off of every region)
constexpr int FLAG_WORDS = 2 * MAX_RANKS + 2 + 1;

__global__ void __launch_bounds__(32) barrier_kernel(const Bases b, size_t flag_off, int rank, int R, int chan, unsigned long long timeout_ns) {
  unsigned int* mine = reinterpret_cast<unsigned int*>(b.p[rank] + flag_off);
  unsigned int* epoch = mine + 2 * MAX_RANKS + chan;
  unsigned int* err = mine + 2 * MAX_RANKS + 2;
  const unsigned int e = *epoch + 1u;
  const int t = threadIdx.x;
  const bool active = *reinterpret_cast<volatile unsigned int*>(err) == 0u && t < R;
  if (active) {
    // arrive: everything this GPU wrote before (this and earlier kernels of the stream) is ordered before the flag
    __threadfence_system();
    unsigned int* theirs = reinterpret_cast<unsigned int*>(b.p[t] + flag_off) + chan * MAX_RANKS + rank;
    st_release_sys(theirs, e);
  }
  __syncwarp();                  // every arrival is issued before any lane starts to spin
  if (active) {
    // wait for rank t's arrival
    const unsigned int* slot = mine + chan * MAX_RANKS + t;
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(ld_acquire_sys(slot) - e) < 0) {
      if (globaltimer_ns() - t0 > timeout_ns) {
        atomicExch(err, 1u + (unsigned)t);
        break;
      }
    }
  }
  __syncwarp();
  if (t == 0) *epoch = e;
}

// the all_gather of embeddings and labels as the producer's store loop: emb[i][:] -> row (row0 + i) of the [N][D] field at f_off and
// labels[i] -> element (row0 + i) of the [N] int64 field at y_off of EVERY rank's region (D % 4 == 0; no packing / unpacking pass)
__global__ void __launch_bounds__(256) gather_rows_kernel(const float4* __restrict__ emb, const int64_t* __restrict__ labels, int n, int D4, const Bases b,
                                                          size_t f_off, size_t y_off, int row0, int R) {
  const long long total = (long long)n * D4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total + n; i += (long long)gridDim.x * blockDim.x) {
    if (i < total) {
      const float4 v = emb[i];
      for (int k = 0; k < R; ++k) reinterpret_cast<float4*>(b.p[k] + f_off)[(size_t)row0 * D4 + i] = v;
    } else {
      const int r = (int)(i - total);
      const int64_t v = labels[r];
      for (int k = 0; k < R; ++k) reinterpret_cast<int64_t*>(b.p[k] + y_off)[row0 + r] = v;
    }
  }
}

// src[0 .. n4) (16-byte words) -> word (dst_word0 + i) of every rank's buffer
__global__ void __launch_bounds__(256) bcast_kernel(const uint4* __restrict__ src, long long n4, const Bases b, size_t dst_off, long long dst_word0, int R) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = src[i];
    for (int k = 0; k < R; ++k) reinterpret_cast<uint4*>(b.p[k] + dst_off)[dst_word0 + i] = v;
  }
}

// Two-shot all-reduce (sum) of n4 float4 words at byte offset `off` of every region. Rank r owns words [r * per, (r + 1) * per).
template <int RT>
__global__ void __launch_bounds__(512) allreduce_kernel(const Bases b, size_t off, long long n4, long long per, int rank, int R_rt) {
  const int R = RT > 0 ? RT : R_rt;
  const long long lo = (long long)rank * per;
  long long hi = lo + per;
  if (hi > n4) hi = n4;
  for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
    float4 v[RT > 0 ? RT : MAX_RANKS];
#pragma unroll
    for (int k = 0; k < (RT > 0 ? RT : MAX_RANKS); ++k)
      if (k < R) v[k] = __ldcg(reinterpret_cast<const float4*>(b.p[k] + off) + i);      // all loads in flight before the adds
    float4 s = v[0];
#pragma unroll
    for (int k = 1; k < (RT > 0 ? RT : MAX_RANKS); ++k)
      if (k < R) { s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w; }
#pragma unroll
    for (int k = 0; k < (RT > 0 ? RT : MAX_RANKS); ++k)
      if (k < R) reinterpret_cast<float4*>(b.p[k] + off)[i] = s;
  }
}

// dst[i] = scale * sum over r (rank order) of slots[r][i]: the local half of a one-shot all-reduce of a few hundred fp64 values (every
// rank stored its [n] block into slot `rank` of every region before the barrier). Same order everywhere -> bit-identical results.
__global__ void __launch_bounds__(256) sum_slots_kernel(const double* __restrict__ slots, int R, int n, double scale, double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int r = 0; r < R; ++r) s += __ldcg(slots + (size_t)r * n + i);
  dst[i] = s * scale;
}

static int load_bases(const unsigned long long* bases, int R, Bases& b) {
  PC_REQUIRE(bases != nullptr && R >= 1 && R <= MAX_RANKS, PC_EINVAL, "peer: need 1..%d region base addresses", MAX_RANKS);
  for (int k = 0; k < MAX_RANKS; ++k) b.p[k] = k < R ? bases[k] : 0ull;
  for (int k = 0; k < R; ++k) PC_REQUIRE(b.p[k] != 0ull && (b.p[k] & 15ull) == 0ull, PC_EINVAL, "peer: region base %d is null or not 16-byte aligned", k);
  return PC_OK;
}

typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
static GetRangeFn get_range_fn() {
  static GetRangeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (GetRangeFn)f;
  }();
  return fn;
}

}  // namespace peer
}  // namespace pc

using namespace pc;
using pc::peer::Bases;

extern "C" int pc_peer_flag_bytes(void) { return (int)(sizeof(unsigned int) * pc::peer::FLAG_WORDS + 255) / 256 * 256; }
extern "C" int pc_peer_max_ranks(void) { return pc::peer::MAX_RANKS; }

extern "C" int pc_peer_export(const void* ptr, unsigned char* handle64, size_t* offset) {
  PC_REQUIRE(ptr && handle64 && offset, PC_EINVAL, "pc_peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  pc::peer::GetRangeFn gr = pc::peer::get_range_fn();
  PC_REQUIRE(gr != nullptr, PC_ECUDA, "pc_peer_export: cuMemGetAddressRange is not available from this driver");
  CUdeviceptr base = 0;
  size_t size = 0;
  const CUresult cr = gr(&base, &size, (CUdeviceptr)ptr);
  PC_REQUIRE(cr == CUDA_SUCCESS, PC_ECUDA, "pc_peer_export: cuMemGetAddressRange failed (CUresult %d)", (int)cr);
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base));
  if (e != cudaSuccess) (void)cudaGetLastError();      // do not leave the error for the next launch check: the caller falls back to NCCL
  PC_REQUIRE(e == cudaSuccess, PC_ECUDA,
             "pc_peer_export: cudaIpcGetMemHandle failed (%s); the region must come from cudaMalloc (PyTorch's default caching allocator, not "
             "expandable_segments)", cudaGetErrorString(e));
  memcpy(handle64, &h, 64);
  *offset = (size_t)((CUdeviceptr)ptr - base);
  return PC_OK;
}

extern "C" int pc_peer_open(const unsigned char* handle64, void** base) {
  PC_REQUIRE(handle64 && base, PC_EINVAL, "pc_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  const cudaError_t e = cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) (void)cudaGetLastError();
  PC_REQUIRE(e == cudaSuccess, PC_ECUDA, "pc_peer_open: cudaIpcOpenMemHandle failed (%s)", cudaGetErrorString(e));
  return PC_OK;
}

extern "C" int pc_peer_close(void* base) {
  PC_REQUIRE(base, PC_EINVAL, "pc_peer_close: null pointer");
  PC_CUDA(cudaIpcCloseMemHandle(base));
  return PC_OK;
}

extern "C" int pc_peer_barrier(const unsigned long long* bases, int R, int rank, size_t flag_off, int channel, int timeout_ms, pc_stream_t stream) {
  Bases b;
  if (int rc = pc::peer::load_bases(bases, R, b)) return rc;
  PC_REQUIRE(rank >= 0 && rank < R && (channel == 0 || channel == 1) && timeout_ms > 0, PC_EINVAL, "pc_peer_barrier: bad arguments");
  pc::peer::barrier_kernel<<<1, 32, 0, stream>>>(b, flag_off, rank, R, channel, (unsigned long long)timeout_ms * 1000000ull);
  PC_LAUNCH_CHECK("peer::barrier_kernel");
  return PC_OK;
}

extern "C" int pc_peer_error(const unsigned long long* bases, int R, int rank, size_t flag_off, int reset, int* out, pc_stream_t stream) {
  PC_REQUIRE(bases && out && rank >= 0 && rank < R, PC_EINVAL, "pc_peer_error: bad arguments");
  unsigned int* err = reinterpret_cast<unsigned int*>(bases[rank] + flag_off) + 2 * pc::peer::MAX_RANKS + 2;
  unsigned int v = 0;
  PC_CUDA(cudaMemcpyAsync(&v, err, sizeof(v), cudaMemcpyDeviceToHost, stream));
  PC_CUDA(cudaStreamSynchronize(stream));
  if (reset && v != 0u) PC_CUDA(cudaMemsetAsync(err, 0, sizeof(v), stream));
  *out = (int)v;
  return PC_OK;
}

extern "C" int pc_dp_gather_peer(const float* emb, const int64_t* labels, int n, int D, const unsigned long long* bases, int R, size_t f_off, size_t y_off,
                                 int row0, pc_stream_t stream) {
  PC_REQUIRE(emb && labels && n > 0 && D > 0 && D % 4 == 0 && row0 >= 0 && f_off % 16 == 0 && y_off % 8 == 0 && (reinterpret_cast<uintptr_t>(emb) & 15u) == 0,
             PC_EINVAL, "pc_dp_gather_peer: bad arguments (D must be a multiple of 4, emb and the field offsets 16-byte aligned)");
  Bases b;
  if (int rc = pc::peer::load_bases(bases, R, b)) return rc;
  int grid = ceil_div((long long)n * (D / 4) + n, 256);
  if (grid > kNumSMs) grid = kNumSMs;
  pc::peer::gather_rows_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(emb), labels, n, D / 4, b, f_off, y_off, row0, R);
  PC_LAUNCH_CHECK("peer::gather_rows_kernel");
  return PC_OK;
}

extern "C" int pc_peer_bcast(const void* src, size_t bytes, const unsigned long long* bases, int R, size_t dst_off, pc_stream_t stream) {
  PC_REQUIRE(src && bytes > 0 && bytes % 16 == 0 && dst_off % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0, PC_EINVAL,
             "pc_peer_bcast: source, size and destination offset must be 16-byte aligned");
  Bases b;
  if (int rc = pc::peer::load_bases(bases, R, b)) return rc;
  const long long n4 = (long long)(bytes / 16);
  int grid = ceil_div(n4, 256);
  if (grid > kNumSMs) grid = kNumSMs;
  pc::peer::bcast_kernel<<<grid, 256, 0, stream>>>(static_cast<const uint4*>(src), n4, b, 0, (long long)(dst_off / 16), R);
  PC_LAUNCH_CHECK("peer::bcast_kernel");
  return PC_OK;
}

extern "C" int pc_peer_allreduce(const unsigned long long* bases, int R, int rank, size_t off, long long count, int blocks, pc_stream_t stream) {
  PC_REQUIRE(off % 16 == 0 && count > 0 && count % 4 == 0 && rank >= 0 && rank < R, PC_EINVAL,
             "pc_peer_allreduce: the offset must be 16-byte aligned and the element count a multiple of 4");
  Bases b;
  if (int rc = pc::peer::load_bases(bases, R, b)) return rc;
  const long long n4 = count / 4;
  const long long per = (n4 + R - 1) / R;
  if (blocks <= 0) blocks = 64;
  const long long need = (per + 511) / 512;
  if (blocks > need) blocks = (int)(need > 0 ? need : 1);
  switch (R) {
    case 2: pc::peer::allreduce_kernel<2><<<blocks, 512, 0, stream>>>(b, off, n4, per, rank, R); break;
    case 4: pc::peer::allreduce_kernel<4><<<blocks, 512, 0, stream>>>(b, off, n4, per, rank, R); break;
    case 8: pc::peer::allreduce_kernel<8><<<blocks, 512, 0, stream>>>(b, off, n4, per, rank, R); break;
    default: pc::peer::allreduce_kernel<0><<<blocks, 512, 0, stream>>>(b, off, n4, per, rank, R); break;
  }
  PC_LAUNCH_CHECK("peer::allreduce_kernel");
  return PC_OK;
}

extern "C" int pc_peer_sum_slots(const double* slots, int R, int n, double scale, double* dst, pc_stream_t stream) {
  PC_REQUIRE(slots && dst && R >= 1 && n > 0, PC_EINVAL, "pc_peer_sum_slots: bad arguments");
  pc::peer::sum_slots_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(slots, R, n, scale, dst);
  PC_LAUNCH_CHECK("peer::sum_slots_kernel");
  return PC_OK;
}

// A region of the library's own (cudaMalloc, zero-filled): used when the caller's allocator hands out memory that cannot be exported
// over CUDA IPC (PyTorch's expandable segments are virtual-memory mappings, not cudaMalloc blocks).
extern "C" int pc_peer_alloc(size_t bytes, void** ptr) {
  PC_REQUIRE(ptr && bytes > 0, PC_EINVAL, "pc_peer_alloc: bad arguments");
  PC_CUDA(cudaMalloc(ptr, bytes));
  PC_CUDA(cudaMemset(*ptr, 0, bytes));
  PC_CUDA(cudaDeviceSynchronize());
  return PC_OK;
}

extern "C" int pc_peer_free(void* ptr) {
  PC_REQUIRE(ptr, PC_EINVAL, "pc_peer_free: null pointer");
  PC_CUDA(cudaFree(ptr));
  return PC_OK;
}
