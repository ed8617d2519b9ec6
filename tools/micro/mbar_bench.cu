// Microbenchmark: cost of mbarrier hand-offs in a warp-specialised producer/consumer chain (no data, no MMAs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I two-tower-model-v2_b200/csrc -I include -o gpurun_out/mbar_bench tools/micro/mbar_bench.cu
// Roles per CTA: P (warp 0 lane 0) -> S (NS warps) -> M (warp 1 lane 0) -> back to S; rings of RAW / BST stages.
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace tt::ptx;

__device__ __forceinline__ void wait_plain(uint32_t bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }

template <int NS, int VARIANT>   // VARIANT bit0: one-lane polling; bit1: no watchdog
__global__ void __launch_bounds__(640, 1) chain_kernel(int iters, long long* out) {
  constexpr int RAW = 6, BST = 3;
  __shared__ uint64_t full_raw[RAW], empty_raw[RAW], full_b[BST], empty_b[BST];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < RAW; ++i) { mbar_init(smem_u32(full_raw + i), 1); mbar_init(smem_u32(empty_raw + i), NS); }
    for (int i = 0; i < BST; ++i) { mbar_init(smem_u32(full_b + i), NS); mbar_init(smem_u32(empty_b + i), 1); }
    fence_barrier_init();
  }
  __syncthreads();
  auto W = [&](uint64_t* b, uint32_t par, bool warpwide) {
    const uint32_t a = smem_u32(b);
    if (warpwide && (VARIANT & 1)) { if (lane == 0) { if (VARIANT & 2) wait_plain(a, par); else mbar_wait(a, par, 1); } __syncwarp(); }
    else { if (VARIANT & 2) wait_plain(a, par); else mbar_wait(a, par, 1); }
  };
  const long long t0 = clock64();
  if (warp == 0) {
    if (lane == 0) { int s = 0; uint32_t ph = 0; for (int n = 0; n < iters; ++n) { W(empty_raw + s, ph ^ 1, false); mbar_arrive(smem_u32(full_raw + s)); if (++s == RAW) { s = 0; ph ^= 1; } } }
  } else if (warp == 1) {
    if (lane == 0) { int s = 0; uint32_t ph = 0; for (int n = 0; n < iters; ++n) { W(full_b + s, ph, false); mbar_arrive(smem_u32(empty_b + s)); if (++s == BST) { s = 0; ph ^= 1; } } }
  } else if (warp >= 2 && warp < 2 + NS) {
    int rs = 0, bs = 0; uint32_t rp = 0, bp = 0;
    for (int n = 0; n < iters; ++n) {
      W(full_raw + rs, rp, true);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(empty_raw + rs));
      if (++rs == RAW) { rs = 0; rp ^= 1; }
      W(empty_b + bs, bp ^ 1, true);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(full_b + bs));
      if (++bs == BST) { bs = 0; bp ^= 1; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

template <int NS, int VARIANT>
void run(const char* name, int threads) {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  const int iters = 2000;
  chain_kernel<NS, VARIANT><<<148, threads>>>(iters, d);
  chain_kernel<NS, VARIANT><<<148, threads>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-48s threads %4d: %7.1f cycles per hand-off round (%s)\n", name, threads, (double)mx / iters, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<4, 0>("4 consumer warps, all lanes poll, watchdog", 640);
  run<4, 1>("4 consumer warps, one lane polls, watchdog", 640);
  run<4, 2>("4 consumer warps, all lanes poll, plain loop", 640);
  run<4, 3>("4 consumer warps, one lane polls, plain loop", 640);
  run<1, 3>("1 consumer warp, one lane polls, plain loop", 640);
  run<8, 3>("8 consumer warps, one lane polls, plain loop", 640);
  run<4, 3>("4 consumer warps, one lane, plain, 192 threads", 192);
  run<4, 0>("4 consumer warps, all lanes, watchdog, 192 thr", 192);
  return 0;
}
