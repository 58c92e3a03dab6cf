// Peer-to-peer plumbing for the data-parallel gradient exchange over NVLink 5 / NVSwitch (row a11 of
// SURVEY.md section 8; replaces the NCCLAllreduce ops of detectron/modeling/optimizer_wsl.py:52-72).
//
// The bulk data moves with the COPY ENGINES (cudaMemcpyAsync between peer-mapped buffers), so the SMs
// stay on the tensor-core GEMMs the exchange overlaps with; only the cross-GPU ordering needs kernels:
//   signal: after a stream's copies into a peer have completed (stream order), publish a sequence number
//           into that peer's flag word with a system-scope release store;
//   wait:   spin (one thread per flag, nanosleep back-off) until every flag has reached the sequence
//           number, with a system-scope acquire load; a watchdog bounds the spin so that a lost peer
//           surfaces as an error code instead of a hung GPU.
#include <cuda.h>
#include <algorithm>
#include <cstring>
#include "common.cuh"

namespace nawsod {
namespace {

constexpr int kMaxSignals = 64;
struct SignalArgs { uint32_t* ptr[kMaxSignals]; int n; uint32_t value; };

__global__ void p2p_signal_kernel(const SignalArgs a) {
  const int i = threadIdx.x;
  if (i < a.n) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.ptr[i]), "r"(a.value) : "memory");
  }
}

__global__ void p2p_wait_kernel(const uint32_t* __restrict__ flags, int n, uint32_t value, unsigned long long timeout_ns,
                                uint32_t* __restrict__ status) {
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    unsigned ns = 32;
    while (true) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
      if (static_cast<int32_t>(v - value) >= 0) break;          // sequence numbers wrap
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > timeout_ns) { if (status) atomicMax(status, 1u); break; }
      __nanosleep(ns);
      if (ns < 1024) ns <<= 1;
    }
  }
  __threadfence_system();
}

// SM-driven scatter / broadcast over NVLink: ONE launch moves `bytes` from src[i] to dst[i] for every peer i
// (the reduce-scatter leg sends a different gradient slice to each owner, the all-gather leg the same operand
// slice to everyone) and then publishes the sequence number, so a bucket costs one kernel instead of W-1 copies
// + a signal.  The kernel needs no shared memory and few registers, so its CTAs run on the SMs the persistent
// tcgen05 GEMMs already occupy -- the transfer genuinely overlaps the math (an NCCL kernel has to wait for
// those SMs, and the copy engines reach only ~400 GB/s of egress when eight ranks scatter at once).  Work is
// cut into 32 KB pieces dealt round-robin over the peers, so all links carry traffic all the time; each thread
// keeps four independent 16-byte loads in flight; remote stores are posted.
constexpr int kMaxPeers = 15;
constexpr int kScatterThreads = 512;
constexpr int kScatterUnroll = 4;
struct ScatterArgs {
  const char* src[kMaxPeers]; char* dst[kMaxPeers];
  int npeers; long long bytes;
  uint32_t* flag[kMaxPeers + 1]; int nflags; uint32_t value;
  int slot;
};
__device__ unsigned int g_scatter_done[128];

__global__ void __launch_bounds__(kScatterThreads) p2p_scatter_kernel(const ScatterArgs a) {
  constexpr long long kPiece = (long long)kScatterThreads * 16 * kScatterUnroll;
  const long long pieces_pp = (a.bytes + kPiece - 1) / kPiece;
  const long long total = pieces_pp * a.npeers;
  for (long long c = blockIdx.x; c < total; c += gridDim.x) {
    const int peer = static_cast<int>(c % a.npeers);
    const long long base = (c / a.npeers) * kPiece + (long long)threadIdx.x * 16;
    const char* s = a.src[peer];
    char* d = a.dst[peer];
    uint4 v[kScatterUnroll];
#pragma unroll
    for (int u = 0; u < kScatterUnroll; ++u) {
      const long long o = base + (long long)u * kScatterThreads * 16;
      if (o < a.bytes) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                                    : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(s + o));
    }
#pragma unroll
    for (int u = 0; u < kScatterUnroll; ++u) {
      const long long o = base + (long long)u * kScatterThreads * 16;
      if (o < a.bytes) asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(d + o), "r"(v[u].x), "r"(v[u].y), "r"(v[u].z), "r"(v[u].w) : "memory");
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(&g_scatter_done[a.slot], 1u);
    if (prev == gridDim.x - 1) {                 // last CTA: every piece of every CTA is out -> publish
      g_scatter_done[a.slot] = 0;
      __threadfence_system();
      for (int i = 0; i < a.nflags; ++i)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.flag[i]), "r"(a.value) : "memory");
    }
  }
}


// Measured in round 2 and removed: a TMA-driven variant of this kernel (one warp per CTA issuing cp.async.bulk global -> shared ->
// peer global through a 3 x 8 KB ring).  It moves the bytes without SM loads / stores, but it needs ~25 KB of shared memory per
// CTA, and an SM whose shared-memory carve-out was sized for a persistent tcgen05 GEMM CTA (193 KB -> the 196 KB configuration)
// has 2 KB left: the scatter CTAs could only start in the gaps between GEMM launches (8 GPUs: a 180 MB panel took 0.56 ms
// instead of 0.3 ms, profiles/r2h_bench_n8_tma.json) and, with the next step's fc6 waiting inside the kernel for the operands
// they carry, not at all.  This kernel needs no shared memory and co-resides.
}  // namespace
}  // namespace nawsod

using namespace nawsod;

extern "C" int nawsod_p2p_scatter(const void* const* srcs, void* const* dsts, int npeers, int64_t bytes,
                                  void* const* flag_ptrs, int nflags, uint32_t value, int slot, void* stream) {
  NAWSOD_REQUIRE(npeers >= 0 && npeers <= kMaxPeers && nflags >= 0 && nflags <= kMaxPeers + 1 && slot >= 0 && slot < 128,
                 NAWSOD_ERR_ARG, "p2p_scatter: at most %d peers, %d flags, slot in [0, 128)", kMaxPeers, kMaxPeers + 1);
  NAWSOD_REQUIRE(bytes >= 0 && bytes % 16 == 0, NAWSOD_ERR_ALIGN, "p2p_scatter: byte count must be a multiple of 16");
  ScatterArgs a;
  a.npeers = (bytes == 0) ? 0 : npeers; a.bytes = bytes; a.nflags = nflags; a.value = value; a.slot = slot;
  for (int i = 0; i < npeers; ++i) {
    NAWSOD_REQUIRE(srcs && dsts && srcs[i] && dsts[i] && aligned16(srcs[i]) && aligned16(dsts[i]), NAWSOD_ERR_ALIGN,
                   "p2p_scatter: source / destination %d is null or not 16-byte aligned", i);
    a.src[i] = static_cast<const char*>(srcs[i]); a.dst[i] = static_cast<char*>(dsts[i]);
  }
  for (int i = 0; i < nflags; ++i) {
    NAWSOD_REQUIRE(flag_ptrs && flag_ptrs[i], NAWSOD_ERR_ARG, "p2p_scatter: null flag pointer %d", i);
    a.flag[i] = static_cast<uint32_t*>(flag_ptrs[i]);
  }
  if (a.npeers == 0 && nflags == 0) return NAWSOD_OK;
  const long long piece = (long long)kScatterThreads * 16 * kScatterUnroll;
  const long long total = std::max<long long>(1, ((bytes + piece - 1) / piece) * std::max(a.npeers, 0));
  // 32 CTAs keep the links busy and leave the co-resident GEMMs most of their issue slots (8 GPUs: 5.07 ms per step against
  // 5.15 ms with two CTAs per SM, profiles/r2c_bench_n8_sm32.json / _sm.json)
  const long long want = get_tuning("p2p_ctas", 0) > 0 ? get_tuning("p2p_ctas", 0) : 32;
  const int grid = (int)std::max<long long>(1, std::min<long long>(total, want));
  p2p_scatter_kernel<<<grid, kScatterThreads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_p2p_enable_peer_access(int peer_device) {
  int dev = 0, can = 0;
  NAWSOD_CUDA_OK(cudaGetDevice(&dev));
  if (peer_device == dev) return NAWSOD_OK;
  NAWSOD_CUDA_OK(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  NAWSOD_REQUIRE(can, NAWSOD_ERR_UNSUPPORTED, "p2p: device %d cannot access device %d (no NVLink / PCIe peer path)", dev, peer_device);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return NAWSOD_OK; }
  NAWSOD_CUDA_OK(e);
  return NAWSOD_OK;
}

extern "C" int nawsod_p2p_get_mem_handle(const void* ptr, void* handle_out, int64_t handle_bytes, int64_t* offset_out) {
  NAWSOD_REQUIRE(ptr && handle_out && offset_out && handle_bytes == (int64_t)sizeof(cudaIpcMemHandle_t), NAWSOD_ERR_ARG,
                 "p2p_get_mem_handle: need a device pointer and a %d-byte handle buffer", (int)sizeof(cudaIpcMemHandle_t));
  cudaIpcMemHandle_t h;
  NAWSOD_CUDA_OK(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));      // names the whole underlying allocation
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  NAWSOD_CUDA_OK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &q));
  NAWSOD_REQUIRE(fp && q == cudaDriverEntryPointSuccess, NAWSOD_ERR_CUDA, "cuMemGetAddressRange is not available");
  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = reinterpret_cast<RangeFn>(fp)(&base, &size, reinterpret_cast<CUdeviceptr>(ptr));
  NAWSOD_REQUIRE(r == CUDA_SUCCESS, NAWSOD_ERR_CUDA, "cuMemGetAddressRange failed with %d", (int)r);
  memcpy(handle_out, &h, sizeof(h));
  *offset_out = (int64_t)(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return NAWSOD_OK;
}

extern "C" int nawsod_p2p_open_mem_handle(const void* handle, int64_t handle_bytes, void** base) {
  NAWSOD_REQUIRE(handle && base && handle_bytes == (int64_t)sizeof(cudaIpcMemHandle_t), NAWSOD_ERR_ARG,
                 "p2p_open_mem_handle: need a %d-byte cudaIpcMemHandle_t", (int)sizeof(cudaIpcMemHandle_t));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  // opened in the CURRENT device's context, with peer access to the exporting device: kernels of this
  // device may then dereference the mapping (a mapping opened in the exporter's context may not be)
  NAWSOD_CUDA_OK(cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess));
  return NAWSOD_OK;
}

extern "C" int nawsod_p2p_close_mem_handle(void* base) {
  if (!base) return NAWSOD_OK;
  NAWSOD_CUDA_OK(cudaIpcCloseMemHandle(base));
  return NAWSOD_OK;
}

extern "C" int nawsod_p2p_copy(void* dst, const void* src, int64_t bytes, void* stream) {
  NAWSOD_REQUIRE(bytes >= 0, NAWSOD_ERR_SHAPE, "p2p_copy: negative size");
  if (bytes == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(dst && src, NAWSOD_ERR_ARG, "p2p_copy: null pointer");
  NAWSOD_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
  return NAWSOD_OK;
}

extern "C" int nawsod_p2p_signal(void* const* flag_ptrs, int n, uint32_t value, void* stream) {
  NAWSOD_REQUIRE(n >= 0 && n <= kMaxSignals, NAWSOD_ERR_ARG, "p2p_signal: 0..%d flags per call", kMaxSignals);
  if (n == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(flag_ptrs, NAWSOD_ERR_ARG, "p2p_signal: null pointer table");
  SignalArgs a;
  a.n = n; a.value = value;
  for (int i = 0; i < n; ++i) {
    NAWSOD_REQUIRE(flag_ptrs[i] && (reinterpret_cast<uintptr_t>(flag_ptrs[i]) & 3u) == 0, NAWSOD_ERR_ARG, "p2p_signal: bad flag pointer %d", i);
    a.ptr[i] = static_cast<uint32_t*>(flag_ptrs[i]);
  }
  p2p_signal_kernel<<<1, kMaxSignals, 0, static_cast<cudaStream_t>(stream)>>>(a);
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}

extern "C" int nawsod_p2p_wait(const void* flags, int n, uint32_t value, int64_t timeout_ms, void* status, void* stream) {
  NAWSOD_REQUIRE(n >= 0 && timeout_ms > 0, NAWSOD_ERR_ARG, "p2p_wait: bad arguments");
  if (n == 0) return NAWSOD_OK;
  NAWSOD_REQUIRE(flags, NAWSOD_ERR_ARG, "p2p_wait: null flags");
  p2p_wait_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint32_t*>(flags), n, value,
                                                                   (unsigned long long)timeout_ms * 1000000ull,
                                                                   static_cast<uint32_t*>(status));
  NAWSOD_LAUNCH_OK();
  return NAWSOD_OK;
}
