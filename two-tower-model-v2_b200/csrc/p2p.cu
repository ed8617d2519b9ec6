// All-gather of small per-rank records over NVLink peer memory, done by our own kernels (no NCCL on the
// data path): every rank PUSHES its record into its slot of every peer's receive buffer with plain
// 16-byte stores, then raises a per-(peer, rank) sequence flag; the consumer side runs a one-warp wait
// kernel on the flags and then reads its local receive buffer.  The peer pointers come from CUDA IPC
// handles exchanged once at set-up (two-tower-model-v2_b200/sharded.py); this file only sees pointers.
//
// Ordering: every block fences its peer stores at system scope before counting itself done; the block
// that completes the count fences again and publishes the flags with system-scope release stores.  The
// waiter polls with system-scope acquire loads.  Receive buffers are double-buffered by the caller
// (parity of the sequence number): a peer can only be one exchange ahead of the slowest rank.
#include <string.h>
#include "tt_common.cuh"

namespace tt {

constexpr int P2P_MAX_RANKS = 64;

struct PushPeers {
  unsigned char* dst[P2P_MAX_RANKS];   // where my record goes on peer g
  int* flag[P2P_MAX_RANKS];            // my flag cell on peer g
};

__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
p2p_push_kernel(const uint4* __restrict__ src, size_t n16, PushPeers peers, int G, int seq, unsigned int* done) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const uint4 v = src[i];
    for (int g = 0; g < G; ++g) reinterpret_cast<uint4*>(peers.dst[g])[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = (atomicAdd(done, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence_system();
    if (threadIdx.x < G) st_release_sys(peers.flag[threadIdx.x], seq);
    if (threadIdx.x == 0) *done = 0u;       // ready for the next push on this stream
  }
}

// One warp: lane g waits until rank g's record of exchange `seq` has landed.  A rank that never shows up
// must not hang the device: after `timeout_cycles` the kernel gives up and reports through *timed_out.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void p2p_wait_kernel(const int* __restrict__ flags, int G, int seq, unsigned long long timeout_ns, int* timed_out) {
  const unsigned long long t0 = globaltimer_ns();     // wall-clock nanoseconds: independent of the SM clock
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    while (ld_acquire_sys(flags + g) - seq < 0) {
      if (globaltimer_ns() - t0 > timeout_ns) { atomicExch(timed_out, seq); return; }
      __nanosleep(200);
    }
  }
}

}  // namespace tt

using namespace tt;

extern "C" __attribute__((visibility("default"))) int tt_p2p_push(const void* src, size_t nbytes, void* const* peer_dst, int32_t* const* peer_flag, int G,
                                                     int32_t seq, uint32_t* done_counter, void* stream) {
  TT_CHECK_ARG(src && peer_dst && peer_flag && done_counter, "null pointer");
  TT_CHECK_ARG(G >= 1 && G <= P2P_MAX_RANKS, "need 1 <= G <= 64");
  TT_CHECK_ARG(nbytes % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0, "record must be 16-byte aligned and sized");
  PushPeers pp{};
  for (int g = 0; g < G; ++g) {
    TT_CHECK_ARG(peer_dst[g] && peer_flag[g] && (reinterpret_cast<uintptr_t>(peer_dst[g]) & 15) == 0, "bad peer pointer");
    pp.dst[g] = reinterpret_cast<unsigned char*>(peer_dst[g]);
    pp.flag[g] = reinterpret_cast<int*>(peer_flag[g]);
  }
  const size_t n16 = nbytes / 16;
  size_t blocks = (n16 + 255) / 256;
  const size_t cap = (size_t)num_sms() * 2;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  p2p_push_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(src), n16, pp, G, seq,
                                                                    done_counter);
  TT_CHECK_LAUNCH();
  return TT_OK;
}

extern "C" __attribute__((visibility("default"))) int tt_p2p_wait(const int32_t* flags, int G, int32_t seq, double timeout_seconds, int32_t* timed_out,
                                                     void* stream) {
  TT_CHECK_ARG(flags && timed_out, "null pointer");
  TT_CHECK_ARG(G >= 1 && G <= P2P_MAX_RANKS, "need 1 <= G <= 64");
  TT_CHECK_ARG(timeout_seconds > 0.0 && timeout_seconds < 3600.0, "timeout_seconds out of range");
  const unsigned long long ns = (unsigned long long)(timeout_seconds * 1.0e9);
  p2p_wait_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(flags, G, seq, ns, timed_out);
  TT_CHECK_LAUNCH();
  return TT_OK;
}

extern "C" __attribute__((visibility("default"))) int tt_p2p_enable_peer(int peer_device) {
  int dev = 0, can = 0;
  TT_CHECK_CUDA(cudaGetDevice(&dev));
  if (peer_device == dev) return TT_OK;
  TT_CHECK_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
  if (!can) { set_error("tt_p2p_enable_peer: device " + std::to_string(dev) + " cannot access device " + std::to_string(peer_device)); return TT_ERR_UNSUPPORTED; }
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return TT_OK; }
  TT_CHECK_CUDA(e);
  return TT_OK;
}

// Receive buffers are allocated here (cudaMalloc, not a sub-block of a caching allocator) so that the IPC
// handle describes exactly this buffer, and peers are opened on the CALLING rank's device: the mapping then
// lives in this device's address space with peer access enabled (cudaIpcMemLazyEnablePeerAccess), which is
// what kernels launched on this device need.
extern "C" __attribute__((visibility("default"))) int tt_p2p_alloc(size_t bytes, void** dev_ptr, void* handle64) {
  TT_CHECK_ARG(dev_ptr && handle64 && bytes > 0, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  TT_CHECK_CUDA(cudaMalloc(&p, bytes));
  TT_CHECK_CUDA(cudaMemset(p, 0, bytes));
  TT_CHECK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); TT_CHECK_CUDA(e); }
  memcpy(handle64, &h, sizeof(h));
  *dev_ptr = p;
  return TT_OK;
}

extern "C" __attribute__((visibility("default"))) int tt_p2p_open(const void* handle64, void** dev_ptr) {
  TT_CHECK_ARG(dev_ptr && handle64, "bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  TT_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return TT_OK;
}

extern "C" __attribute__((visibility("default"))) int tt_p2p_close(void* dev_ptr) {
  if (dev_ptr) TT_CHECK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return TT_OK;
}

extern "C" __attribute__((visibility("default"))) int tt_p2p_free(void* dev_ptr) {
  if (dev_ptr) TT_CHECK_CUDA(cudaFree(dev_ptr));
  return TT_OK;
}
