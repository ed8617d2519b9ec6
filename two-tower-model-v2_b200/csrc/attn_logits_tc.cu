// Attention-score MLP on the tensor cores at fp32 accuracy (reference: src/models/buyer_tower.py:32-36,85-86):
//     logit[r] = W2 . relu(W1 x_r + b1) + b2
// A single-pass bf16/tf32 MLP misses the 1e-5 tolerance of the pooled output (the event weight, up to 10,
// multiplies the logit inside exp), so the hidden layer uses the 3xTF32 split: every fp32 operand is written
// as big + small with big = rna_tf32(v) and small = rna_tf32(v - big) (both exactly representable in tf32),
// and  x.w ~= xb.wb + xb.ws + xs.wb  (the dropped xs.ws term is ~2^-22 relative), accumulated in fp32 in TMEM
// by tcgen05.mma kind::tf32.  Error vs fp32 FMA arithmetic: ~1e-7 relative, the same order as fp32 itself.
//
// Persistent CTAs (one per SM, 384 threads) loop over 128-row tiles of x:
//   warp 0 lane 0     : TMA producer - per 32-column K-block the raw fp32 tile x[128 x 32] and the big/small W1[Hp x 32]
//                       (128-byte rows, 128B swizzle) into a 3-stage ring
//   warps 2,3,8-11    : splitters - rewrite each landed x tile in place as `big` and write `small` next to it
//                       (same swizzled offsets), fence.proxy.async, arrive (W1 is pre-split once per call)
//   warp 1 lane 0     : MMA issuer - 4 K=8 steps x 3 MMAs per K-block, M = 128, N = Hp, accumulators
//                       double-buffered in TMEM (2 x Hp columns)
//   warps 4-7         : epilogue - tcgen05.ld the [128 x Hp] hidden pre-activations of the finished tile,
//                       relu(.+b1).W2 + b2, one row per thread; overlaps the next tile's MMAs
//   warp 2            : also the TMEM allocator
// Rooflines: tf32 tensor pipe (3 x 2*D*Hp flop per row at half the bf16 rate) and shared-memory bandwidth
// (every MMA streams both operands from shared memory; the splitters add one read and two writes per tile).
#include <cuda.h>
#include <mutex>
#include "tt_common.cuh"
#include "sm100_ptx.cuh"
#include "flat_internal.cuh"

namespace tt {

using namespace ptx;

constexpr int TC_BM = 128;          // rows per CTA
constexpr int TC_BK = 32;           // fp32 per K-block = one 128-byte swizzle row
constexpr int TC_THREADS = 384;
constexpr int TC_SPLIT_WARPS = 6;
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_X_TILE = TC_BM * TC_BK * 4;   // 16 KB

struct AttnTcParams {
  long long R;
  int H, Hp;          // hidden units, padded to a multiple of 32 (<= 256)
  int num_kb;         // ceil(D / 32)
  int num_stages;
  int stage_bytes;    // 2 x-tiles + 2 W-tiles
  int tmem_cols;
  int num_tiles;
  const float* b1;
  const float* W2;
  const float* b2;
  float* logits;
};

__device__ __forceinline__ uint32_t rna_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Instruction descriptor for kind::tf32: A/B = tf32 (K-major), D = fp32, dense.
__host__ __device__ constexpr uint32_t make_idesc_tf32_f32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// In place: tile -> big, small written at the same offsets of `small_tile`.  16-byte chunks strided over the
// TC_SPLIT_WARPS*32 splitter threads, four chunks in flight per thread.
__device__ __forceinline__ void split_chunk(const float4 v, uint4& b, uint4& s) {
  b.x = rna_tf32(v.x); b.y = rna_tf32(v.y); b.z = rna_tf32(v.z); b.w = rna_tf32(v.w);
  s.x = rna_tf32(v.x - __uint_as_float(b.x)); s.y = rna_tf32(v.y - __uint_as_float(b.y));
  s.z = rna_tf32(v.z - __uint_as_float(b.z)); s.w = rna_tf32(v.w - __uint_as_float(b.w));
}
__device__ __forceinline__ void split_tile(uint8_t* tile, uint8_t* small_tile, int bytes, int t) {
  constexpr int STEP = TC_SPLIT_WARPS * 32 * 16;
  int off = t * 16;
  for (; off + 3 * STEP < bytes; off += 4 * STEP) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(tile + off + u * STEP);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint4 b, s;
      split_chunk(v[u], b, s);
      *reinterpret_cast<uint4*>(tile + off + u * STEP) = b;
      *reinterpret_cast<uint4*>(small_tile + off + u * STEP) = s;
    }
  }
  for (; off < bytes; off += STEP) {
    uint4 b, s;
    split_chunk(*reinterpret_cast<const float4*>(tile + off), b, s);
    *reinterpret_cast<uint4*>(tile + off) = b;
    *reinterpret_cast<uint4*>(small_tile + off) = s;
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
attn_logits_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                      const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int w_tile = p.Hp * TC_BK * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.num_stages * p.stage_bytes);
  uint64_t* full_bar = bars;                          // [TC_MAX_STAGES]  TMA landed
  uint64_t* split_bar = bars + TC_MAX_STAGES;         // [TC_MAX_STAGES]  big/small written (one arrival per splitter warp)
  uint64_t* empty_bar = bars + 2 * TC_MAX_STAGES;     // [TC_MAX_STAGES]  MMAs of the stage retired
  uint64_t* tmem_full_bar = bars + 3 * TC_MAX_STAGES; // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;       // [2]  one arrival per epilogue warp
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int my_tiles = (p.num_tiles > (int)blockIdx.x) ? (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0 && lane == 0) { prefetch_tensormap(&tmap_x); prefetch_tensormap(&tmap_w); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.num_stages; ++i) {
      mbar_init(smem_u32(full_bar + i), 1);
      mbar_init(smem_u32(split_bar + i), TC_SPLIT_WARPS);
      mbar_init(smem_u32(empty_bar + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(tmem_full_bar + i), 1);
      mbar_init(smem_u32(tmem_empty_bar + i), 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(smem_u32(tmem_ptr_smem), (uint32_t)p.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const bool splitter = (warp == 2 || warp == 3 || warp >= 8);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const long long row0 = (long long)((int)blockIdx.x + i * (int)gridDim.x) * TC_BM;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(smem_u32(empty_bar + stage), phase ^ 1, 400 + stage);
          const uint32_t fb = smem_u32(full_bar + stage);
          uint8_t* st = smem + (size_t)stage * p.stage_bytes;
          mbar_arrive_expect_tx(fb, (uint32_t)(TC_X_TILE + 2 * w_tile));
          tma_load_2d(smem_u32(st), &tmap_x, fb, kb * TC_BK, (int)row0);                     // x raw (split in place)
          tma_load_2d(smem_u32(st + 2 * TC_X_TILE), &tmap_w, fb, kb * TC_BK, 0);             // W big  (pre-split)
          tma_load_2d(smem_u32(st + 2 * TC_X_TILE + w_tile), &tmap_w, fb, kb * TC_BK, p.Hp); // W small
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32_f32(TC_BM, p.Hp);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int buf = i & 1;
        mbar_wait(smem_u32(tmem_empty_bar + buf), (((uint32_t)i >> 1) & 1u) ^ 1u, 440 + buf);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.Hp);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(smem_u32(split_bar + stage), phase, 410 + stage);
          tc_fence_after();
          uint8_t* st = smem + (size_t)stage * p.stage_bytes;
          const uint64_t xb = make_smem_desc_sw128(smem_u32(st));
          const uint64_t xs = make_smem_desc_sw128(smem_u32(st + TC_X_TILE));
          const uint64_t wb = make_smem_desc_sw128(smem_u32(st + 2 * TC_X_TILE));
          const uint64_t ws = make_smem_desc_sw128(smem_u32(st + 2 * TC_X_TILE + w_tile));
#pragma unroll
          for (int k = 0; k < TC_BK / 8; ++k) {
            const uint64_t koff = (uint64_t)((k * 8 * 4) >> 4);       // 32 bytes per K = 8 step
            mma_tf32_ss(d_tmem, xb + koff, wb + koff, idesc, (uint32_t)((kb | k) != 0));
            mma_tf32_ss(d_tmem, xb + koff, ws + koff, idesc, 1u);
            mma_tf32_ss(d_tmem, xs + koff, wb + koff, idesc, 1u);
          }
          mma_commit(smem_u32(empty_bar + stage));
          if (kb == p.num_kb - 1) mma_commit(smem_u32(tmem_full_bar + buf));
          if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (splitter) {
    const int sw = (warp < 4) ? warp - 2 : warp - 6;          // 0..5
    const int t = sw * 32 + lane;
    int stage = 0;
    uint32_t phase = 0;
    for (int n = 0; n < my_tiles * p.num_kb; ++n) {
      mbar_wait(smem_u32(full_bar + stage), phase, 420 + stage);
      uint8_t* st = smem + (size_t)stage * p.stage_bytes;
      split_tile(st, st + TC_X_TILE, TC_X_TILE, t);
      fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(split_bar + stage));
      if (++stage == p.num_stages) { stage = 0; phase ^= 1; }
    }
  } else {
    // ---- epilogue (warps 4-7): one row per thread ---------------------------------------------------
    const int quad = warp & 3;
    for (int i = 0; i < my_tiles; ++i) {
      const int buf = i & 1;
      const long long row0 = (long long)((int)blockIdx.x + i * (int)gridDim.x) * TC_BM;
      mbar_wait(smem_u32(tmem_full_bar + buf), ((uint32_t)i >> 1) & 1u, 430 + buf);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * p.Hp);
      float logit = 0.f;
      for (int c = 0; c < p.Hp; c += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(taddr + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int h = c + j;
          if (h < p.H) logit = fmaf(fmaxf(__uint_as_float(v[j]) + __ldg(p.b1 + h), 0.f), __ldg(p.W2 + h), logit);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(tmem_empty_bar + buf));
      const long long r = row0 + quad * 32 + lane;
      if (r < p.R) p.logits[r] = logit + __ldg(p.b2);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols); }
}

// W1 f32 [H, D] -> Wsplit f32 [2][Hp][D]: big = rna_tf32(w), small = rna_tf32(w - big); rows >= H are zero.
// Done once per call (the same weights serve every row tile).
__global__ void attn_split_w_kernel(const float* __restrict__ W1, int H, int Hp, int D, float* __restrict__ out) {
  const long long n = (long long)Hp * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int h = (int)(i / D);
    const float w = (h < H) ? W1[i] : 0.f;
    const uint32_t b = rna_tf32(w);
    out[i] = __uint_as_float(b);
    out[n + i] = __uint_as_float(rna_tf32(w - __uint_as_float(b)));
  }
}

// Library-private stream-ordered memory pool (one per device) for the split weights: blocks are reused across
// calls instead of going back to the driver at every synchronisation (the default pool's behaviour).
cudaMemPool_t scratch_pool(int dev) {
  static std::mutex mu;
  static cudaMemPool_t pools[64] = {};
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!pools[dev]) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaMemPool_t pool = nullptr;
    if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    unsigned long long keep = ~0ull;
    (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    pools[dev] = pool;
  }
  return pools[dev];
}

// Returns TT_ERR_UNSUPPORTED (without setting an error) when the shape does not fit this kernel.
int launch_attn_logits_tc(const float* x, long long R, int D, const float* W1, const float* b1, const float* W2,
                          const float* b2, int H, float* logits, cudaStream_t st) {
  const int Hp = (H + 31) / 32 * 32;
  if (Hp > 256 || D % 4 != 0 || R >= (1LL << 31)) return TT_ERR_UNSUPPORTED;
  AttnTcParams p{};
  p.R = R; p.H = H; p.Hp = Hp;
  p.num_kb = (D + TC_BK - 1) / TC_BK;
  p.stage_bytes = 2 * TC_X_TILE + 2 * Hp * TC_BK * 4;
  const int budget = 227 * 1024 - 1024 - 256;
  p.num_stages = budget / p.stage_bytes;
  if (p.num_stages > TC_MAX_STAGES) p.num_stages = TC_MAX_STAGES;
  if (p.num_stages < 2) return TT_ERR_UNSUPPORTED;
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * Hp) p.tmem_cols <<= 1;      // two accumulator buffers
  p.b1 = b1; p.W2 = W2; p.b2 = b2; p.logits = logits;
  // split weights live in a stream-ordered scratch allocation (freed after the kernel on the same stream)
  float* wsplit = nullptr;
  int dev = 0;
  TT_CHECK_CUDA(cudaGetDevice(&dev));
  cudaMemPool_t pool = scratch_pool(dev);
  if (!pool) return TT_ERR_UNSUPPORTED;          // no stream-ordered allocator: the FFMA2 kernel serves the call
  TT_CHECK_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&wsplit), (size_t)2 * Hp * D * sizeof(float), pool, st));
  attn_split_w_kernel<<<64, 256, 0, st>>>(W1, H, Hp, D, wsplit);
  TT_CHECK_LAUNCH();
  CUtensorMap tx, tw;
  if (int e = make_tmap_f32(&tx, x, R, D, TC_BM, TC_BK)) { cudaFreeAsync(wsplit, st); return e; }
  if (int e = make_tmap_f32(&tw, wsplit, 2LL * Hp, D, Hp, TC_BK)) { cudaFreeAsync(wsplit, st); return e; }
  const size_t smem = (size_t)p.num_stages * p.stage_bytes + 1024 + 256;
  TT_CHECK_CUDA(cudaFuncSetAttribute(attn_logits_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  p.num_tiles = (int)((R + TC_BM - 1) / TC_BM);
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  attn_logits_tc_kernel<<<(unsigned)grid, TC_THREADS, smem, st>>>(tx, tw, p);
  count_launch();
  const cudaError_t le = cudaGetLastError();
  TT_CHECK_CUDA(cudaFreeAsync(wsplit, st));
  TT_CHECK_CUDA(le);
  return TT_OK;
}

}  // namespace tt
