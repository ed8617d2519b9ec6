// Fused attention pooling of the buyer tower (reference: src/models/buyer_tower.py:70-101):
//     logit_s = W2 . relu(W1 x_s + b1) + b2 ;  a = softmax_s(logit * w) ;  out = normalize(sum_s a_s x_s)
// ONE kernel, ONE pass over x: a 64-row tile of x lands in shared memory (TMA), is rewritten IN PLACE as the two fp16
// pieces the tensor cores consume, and the same shared-memory tile feeds the softmax-weighted row sum once its logits
// are known.  x crosses HBM -> L2 -> SM exactly once.
//
// Arithmetic of the hidden layer (fp32-accurate at the fp16 tensor rate): every fp32 operand is split into two fp16
// pieces, v*2^e = hi + lo with hi = rn_f16(v*2^e), lo = rn_f16(v*2^e - hi)  (22 significant bits; 2^e is a power of
// two chosen so that `lo` stays a normal fp16: 16 for x, from max|W1| for the weights), and
//     x.w ~= hi_x.hi_w + hi_x.lo_w + lo_x.hi_w          (dropped lo.lo term ~2^-22 relative)
// is accumulated in fp32 in TMEM by three tcgen05.mma kind::f16 per K step; the epilogue undoes 2^e.
// The tensor core TRUNCATES (round toward zero) when it adds into the accumulator: 72 updates of one accumulator
// bias the hidden pre-activations by ~6e-6 relative, which the event weight (up to 10) turns into 1.2e-5 element-wise
// on the pooled output (measured; reproduced by a numpy model of truncating accumulation).  So the large hi.hi
// products go to one accumulator (24 updates) and the two cross terms, 2^-11 smaller, to a second one; the epilogue
// adds the two in fp32.  The pooled sum uses x = (hi + lo) / 16 (22 bits).  Measured error of the pooled output:
// 4e-6 element-wise, 1e-6 norm-wise - the level of an fp32 sgemm.
// A value outside the fp16 range after scaling (|x| > 4094, or a non-finite input) raises a device flag and a
// predicated fp32 CUDA-core kernel recomputes the call: no host synchronisation, always the fp32 answer.
//
// Orientation: D[hidden(128) x rows(64)] = W1 . x^T.  The A operand W1 (hi and lo, 2 x 192 TMEM columns for
// D = 384) is written ONCE per CTA into tensor memory (tcgen05.st); the B operand (64 rows of x, hi and lo tiles) is the
// only MMA operand read from shared memory.
//
// Persistent CTAs (one per SM, 20 warps); CTA c owns a contiguous range of buyers and walks its rows in 64-row tiles,
// two tile buffers of nkb x 16 KB (192 KB for D = 384):
//   warp 4 lane 0   : TMA producer - as soon as a buffer is free, the whole next tile: per 64-column K-block two
//                                    128B-swizzled fp32 boxes [64 rows x 32 cols] (8 KB each), one mbarrier per K-block
//   warps 6-13      : splitters    - K-block by K-block as they land: 8 lanes read one row's 256 raw bytes, then write
//                                    `hi` over the first box and `lo` over the second (UMMA K-major 128B-swizzle layout),
//                                    fence.proxy.async, arrive
//   warp 5 lane 0   : MMA issuer   - per K-block 4 K-steps x 3 MMAs (M = 128 hidden, N = 64 rows, K = 16) into the main and
//                                    the cross-term accumulator (2 x 64 TMEM columns), one commit per tile
//   warps 0-3       : epilogue     - lane = hidden unit: relu(acc*2^-e + b1)*W2, butterfly transpose-reduce over the 128
//                                    hidden units -> one logit per row (shared memory, or global in logits-only mode)
//   warps 14-19     : pooling      - warp k owns the 64 columns of K-block k: online softmax per buyer (running max /
//                                    normaliser / 8 partial sums per lane, buyers may straddle tiles), rows read back
//                                    from the hi/lo tiles in shared memory; at a buyer's last row the six warps combine
//                                    their squared norms and write the normalised [D] row; then the buffer is free.
// Every wait is polled by ONE lane per warp (32 lanes polling one mbarrier are served one after the other; a no-op
// pipeline of the previous version of this kernel, which polled with every lane and handed 16 KB K-blocks through two
// rings, already took 72 us at C2).
// Rooflines: HBM (x once: B*S*D*4 bytes); tensor pipe 3 x 2*D*128 flop per row at the fp16 rate; shared memory
// ~5.5 KB per row (TMA write, splitter read + write, MMA B reads, pooling reads).
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>
#include "tt_common.cuh"
#include "sm100_ptx.cuh"
#include "flat_internal.cuh"

namespace tt {

using namespace ptx;

constexpr int AF_TILE = 64;                          // rows of x per MMA tile (UMMA N)
constexpr int AF_M = 128;                            // hidden units (UMMA M); H <= 128
constexpr int AF_KB_BYTES = 2 * AF_TILE * 128;       // one (tile, K-block): two fp32 boxes [64 x 32]  ->  hi | lo fp16 [64 x 64]
constexpr int AF_MAX_KB = 6;                         // D <= 384: W hi + lo = 2 * 6 * 32 = 384 TMEM columns
constexpr int AF_EPI_WARP0 = 0;                      // 4 warps: TMEM lane quadrant = warp index
constexpr int AF_TMA_WARP = 4;
constexpr int AF_MMA_WARP = 5;                       // also allocates / frees tensor memory
constexpr int AF_SPLIT_WARP0 = 6;
constexpr int AF_SPLIT_WARPS = 8;
constexpr int AF_POOL_WARP0 = AF_SPLIT_WARP0 + AF_SPLIT_WARPS;
constexpr int AF_POOL_WARPS = AF_MAX_KB;             // one per K-block
constexpr int AF_WARPS = AF_POOL_WARP0 + AF_POOL_WARPS;
constexpr int AF_THREADS = AF_WARPS * 32;
constexpr int AF_ACC_COLS = 128;                     // main + cross-term accumulator (64 columns each), then W hi, W lo
constexpr int AF_SMAX = 32768;                       // events per buyer (bounded by the fp32 fallback kernel's shared memory)
constexpr float AF_X_SCALE = 16.0f;
constexpr float AF_F16_MAX = 65504.0f;

struct AttnFusedParams {
  const float* x;          // [R, D]
  const float* w;          // [B, S]       (pool mode)
  float* out;              // [B, D]       (pool mode)
  float* logits_out;       // [R]          (logits-only mode)
  const uint4* Wp;         // [2][nkb][8][128] uint4: fp16 pieces of W1 * 2^kw, 8 K elements per uint4, by hidden unit
  const float* b1;
  const float* W2;
  const float* b2;
  const float* inv_scale;  // 2^-(kw + 4), written by the weight-preparation kernel
  int* flag;               // raised when a value leaves the fp16 range
  long long R;
  int B, S, D, H, nkb;
  int mode;                // TT_B200_ATTN_MODE, timing experiments only (wrong results): 4 no cross-term MMAs, 8 no MMAs
  long long* trace;        // TT_B200_ATTN_TRACE: clock64 of CTA 0's pipeline events, [AF_TRACE_TILES][AF_TRACE_SLOTS]
};
constexpr int AF_TRACE_TILES = 32, AF_TRACE_SLOTS = 32;
#define AF_TR(t, slot) do { if (p.trace && blockIdx.x == 0 && (t) < AF_TRACE_TILES) p.trace[(t) * AF_TRACE_SLOTS + (slot)] = clock64(); } while (0)

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_shared() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One lane polls, the warp joins at __syncwarp.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int tag) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity, tag);
  __syncwarp();
}

// 8 consecutive fp32 -> 8 fp16 `hi` + 8 fp16 `lo` of the scaled values
__device__ __forceinline__ void split8(const float4 a, const float4 b, uint4& hi, uint4& lo, float& mabs) {
  const float v[8] = {a.x * AF_X_SCALE, a.y * AF_X_SCALE, a.z * AF_X_SCALE, a.w * AF_X_SCALE,
                      b.x * AF_X_SCALE, b.y * AF_X_SCALE, b.z * AF_X_SCALE, b.w * AF_X_SCALE};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float2 back = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(v[2 * i] - back.x, v[2 * i + 1] - back.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    mabs = fmaxf(mabs, fmaxf(fabsf(v[2 * i]), fabsf(v[2 * i + 1])));      // (a NaN propagates through the MMA like in fp32)
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ __half2 as_half2(uint32_t u) {
  __half2 h;
  memcpy(&h, &u, sizeof(h));
  return h;
}
// one 16-byte chunk (8 columns) of one row:  acc += coef * hi (fp32),  accl += coef * lo (packed fp16)
__device__ __forceinline__ void fma_chunk(float (&acc)[8], __half2 (&accl)[4], const uint4 hi, const uint4 lo, const float coef) {
  const __half2 ch = __float2half2_rn(coef);
  const float2 a0 = __half22float2(as_half2(hi.x)), a1 = __half22float2(as_half2(hi.y));
  const float2 a2 = __half22float2(as_half2(hi.z)), a3 = __half22float2(as_half2(hi.w));
  acc[0] = fmaf(a0.x, coef, acc[0]); acc[1] = fmaf(a0.y, coef, acc[1]);
  acc[2] = fmaf(a1.x, coef, acc[2]); acc[3] = fmaf(a1.y, coef, acc[3]);
  acc[4] = fmaf(a2.x, coef, acc[4]); acc[5] = fmaf(a2.y, coef, acc[5]);
  acc[6] = fmaf(a3.x, coef, acc[6]); acc[7] = fmaf(a3.y, coef, acc[7]);
  accl[0] = __hfma2(as_half2(lo.x), ch, accl[0]);
  accl[1] = __hfma2(as_half2(lo.y), ch, accl[1]);
  accl[2] = __hfma2(as_half2(lo.z), ch, accl[2]);
  accl[3] = __hfma2(as_half2(lo.w), ch, accl[3]);
}

template <bool POOL>
__global__ void __launch_bounds__(AF_THREADS, 1)
attn_pool_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const AttnFusedParams p) {
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~(uintptr_t)1023);
  const int nkb = p.nkb;
  const int buf_bytes = nkb * AF_KB_BYTES;
  uint8_t* tiles = smem;                                                               // [2][nkb][16 KB]
  float* logits_s = reinterpret_cast<float*>(tiles + 2 * (size_t)buf_bytes);           // [2][64]
  float* partial = logits_s + 2 * AF_TILE;                                             // [2][4][64]
  float* ssq = partial + 2 * 4 * AF_TILE;                                              // [2][8]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ssq + 16);
  uint64_t* full_raw = bars;                              // [2][AF_MAX_KB]  TMA landed
  uint64_t* full_b = full_raw + 2 * AF_MAX_KB;            // [2][AF_MAX_KB]  one arrival per splitter warp
  uint64_t* buf_free = full_b + 2 * AF_MAX_KB;            // [2]  one arrival per pooling warp
  uint64_t* logits_full = buf_free + 2;                   // [2]  one arrival per logits-writing epilogue warp
  uint64_t* acc_full = logits_full + 2;                   // [1]  tcgen05.commit
  uint64_t* acc_empty = acc_full + 1;                     // [1]  one arrival per epilogue warp
  uint64_t* w_bar = acc_empty + 1;                        // [1]  W1 pieces are in TMEM (one arrival per epilogue warp)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- this CTA's buyers / rows -------------------------------------------------------------------
  const int b0 = (int)((long long)blockIdx.x * p.B / gridDim.x);
  const int b1 = (int)((long long)(blockIdx.x + 1) * p.B / gridDim.x);
  const long long r0 = (long long)b0 * p.S;
  const long long nrows = (long long)(b1 - b0) * p.S;
  const int ntiles = (int)((nrows + AF_TILE - 1) / AF_TILE);

  if (warp == AF_TMA_WARP && lane == 0) prefetch_tensormap(&tmap_x);
  if (warp == AF_SPLIT_WARP0 && lane == 0) {
    for (int i = 0; i < 2 * AF_MAX_KB; ++i) { mbar_init(smem_u32(full_raw + i), 1); mbar_init(smem_u32(full_b + i), AF_SPLIT_WARPS); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(buf_free + i), (uint32_t)nkb); mbar_init(smem_u32(logits_full + i), 2); }
    mbar_init(smem_u32(acc_full), 1);
    mbar_init(smem_u32(acc_empty), 4);
    mbar_init(smem_u32(w_bar), 4);
    fence_barrier_init();
  }
  if (warp == AF_MMA_WARP) { tmem_alloc(smem_u32(tmem_ptr_smem), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t w_hi_col = AF_ACC_COLS, w_lo_col = AF_ACC_COLS + (uint32_t)nkb * 32u;

  if (warp == AF_TMA_WARP) {
    // =========================== TMA producer =====================================================
    if (lane == 0) {
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const uint32_t k = (uint32_t)(t >> 1);
        mbar_wait(smem_u32(buf_free + buf), (k & 1u) ^ 1u, 500 + buf);
        AF_TR(t, 1);
        const long long row = r0 + (long long)t * AF_TILE;          // < 2^31 (checked on the host)
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t fb = smem_u32(full_raw + buf * AF_MAX_KB + kb);
          const uint32_t dst = smem_u32(tiles + (size_t)buf * buf_bytes + (size_t)kb * AF_KB_BYTES);
          mbar_arrive_expect_tx(fb, (uint32_t)AF_KB_BYTES);
          tma_load_2d(dst, &tmap_x, fb, kb * 64, (int)row);
          tma_load_2d(dst + AF_KB_BYTES / 2, &tmap_x, fb, kb * 64 + 32, (int)row);
        }
        AF_TR(t, 0);
      }
    }
    __syncwarp();
  } else if (warp == AF_MMA_WARP) {
    // =========================== MMA issuer =======================================================
    if (lane == 0) {
      // per K = 16 step TWO MMAs: the hi and lo tiles of a K-block are adjacent in shared memory (same 8-row-group
      // stride), so one N = 128 MMA with A = hi_w computes hi_w.[hi_x | lo_x] into the main (columns 0-63) and the
      // cross-term accumulator (columns 64-127) at once; a second N = 64 MMA adds lo_w.hi_x to the cross terms.
      // A tensor-memory A operand costs ~100 cycles per MMA whatever N is (64 B/cycle TMEM read), so 2 instead of 3
      // MMAs per step is a third less tensor time.
      constexpr uint32_t idesc128 = make_idesc_f16_f32(AF_M, 2 * AF_TILE);
      constexpr uint32_t idesc64 = make_idesc_f16_f32(AF_M, AF_TILE);
      mbar_wait(smem_u32(w_bar), 0, 510);
      tc_fence_after();
      const uint32_t d_main = tmem_base, d_cross = tmem_base + (uint32_t)AF_TILE;
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        const uint32_t k = (uint32_t)(t >> 1);
        mbar_wait(smem_u32(acc_empty), ((uint32_t)t & 1u) ^ 1u, 520);
        AF_TR(t, 25);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(full_b + buf * AF_MAX_KB + kb), k & 1u, 530 + kb);
          AF_TR(t, 14 + kb);
          tc_fence_after();
          const uint32_t base = smem_u32(tiles + (size_t)buf * buf_bytes + (size_t)kb * AF_KB_BYTES);
          const uint64_t xh = make_smem_desc_sw128(base);          // N = 128: rows 0-63 = hi tile, rows 64-127 = lo tile
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t koff = (uint64_t)((ks * 16 * 2) >> 4);               // 32 bytes per K = 16 step
            const uint32_t a_hi = tmem_base + w_hi_col + (uint32_t)((kb * 4 + ks) * 8);
            const uint32_t a_lo = tmem_base + w_lo_col + (uint32_t)((kb * 4 + ks) * 8);
            const uint32_t first = (uint32_t)((kb | ks) != 0);
            if (!(p.mode & 8)) mma_f16_ts(d_main, a_hi, xh + koff, idesc128, first);              // hi_w.[hi_x | lo_x]
            if (!(p.mode & 12)) mma_f16_ts(d_cross, a_lo, xh + koff, idesc64, 1u);                // + lo_w.hi_x
          }
        }
        mma_commit(smem_u32(acc_full));
        AF_TR(t, 20);
      }
    }
    __syncwarp();
  } else if (warp < 4) {
    // =========================== epilogue: lane = hidden unit ========================================
    const int e = warp;                          // TMEM lane quadrant
    const int h = e * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(e * 32) << 16);
    pdl_wait();      // the weight-preparation kernel (previous launch) has written Wp / inv_scale / flag
    // one-time: this CTA's copy of the W1 pieces into tensor memory (A operand of every MMA)
    for (int j = 0; j < 2 * nkb; ++j) {
      const uint4* src = p.Wp + (size_t)j * 8 * AF_M + h;
      uint32_t r[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 v = __ldg(src + q * AF_M);
        r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
      }
      __syncwarp();
      tmem_st_32x32(lane_base + w_hi_col + (uint32_t)(j * 32), r);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(w_bar));

    const float b1h = (h < p.H) ? __ldg(p.b1 + h) : 0.f;
    const float w2h = (h < p.H) ? __ldg(p.W2 + h) : 0.f;
    const float sc = __ldg(p.inv_scale);
    const float b2v = __ldg(p.b2);
    for (int t = 0; t < ntiles; ++t) {
      mbar_wait_warp(smem_u32(acc_full), (uint32_t)t & 1u, 540);
      if (warp == 0 && lane == 0) AF_TR(t, 21);
      tc_fence_after();
      float* part = partial + (size_t)(t & 1) * 4 * AF_TILE + e * AF_TILE;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32], c[32];
        __syncwarp();
        tmem_ld_32x32(lane_base + (uint32_t)(half * 32), v);                     // main accumulator
        tmem_ld_32x32(lane_base + (uint32_t)(AF_TILE + half * 32), c);           // cross terms
        tmem_ld_wait();
        if (half == 1) {                          // everything is in registers: hand the accumulators back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(acc_empty));
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(c[i]));
        float tv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) tv[i] = w2h * fmaxf(fmaf(__uint_as_float(v[i]), sc, b1h), 0.f);
        // butterfly transpose-reduce: afterwards lane l holds the sum over the warp's 32 hidden units of column l
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < off; ++i) {
            const float send = upper ? tv[i] : tv[i + off];
            const float keep = upper ? tv[i + off] : tv[i];
            tv[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        part[half * 32 + lane] = tv[0];
      }
      named_bar_sync(1, 128);                     // the four partial sums of every column are in shared memory
      if (e < 2) {
        const int n = e * 32 + lane;
        const float* pp = partial + (size_t)(t & 1) * 4 * AF_TILE + n;
        const float logit = ((pp[0] + pp[AF_TILE]) + pp[2 * AF_TILE]) + pp[3 * AF_TILE] + b2v;
        const long long rl = (long long)t * AF_TILE + n;
        if (POOL) logits_s[(t & 1) * AF_TILE + n] = logit;
        else if (rl < nrows) p.logits_out[r0 + rl] = logit;
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(logits_full + (t & 1)));
        if (warp == 0 && lane == 0) AF_TR(t, 22);
      }
    }
  } else if (warp >= AF_SPLIT_WARP0 && warp < AF_POOL_WARP0) {
    // =========================== splitters: raw fp32 K-block -> hi | lo fp16, in place =================
    const int tid = (warp - AF_SPLIT_WARP0) * 32 + lane;          // 0..255
    float mabs = 0.f;
    constexpr int UPT = 512 / (AF_SPLIT_WARPS * 32);              // 32-byte units per thread per K-block
    // the 8 lanes of a row sit in one warp and handle the same unit index i: a row is read completely (all lanes of the
    // load instruction) before any lane overwrites it
    int row[UPT], rd0[UPT], rd1[UPT], wr[UPT];
    bool swap[UPT];
#pragma unroll
    for (int i = 0; i < UPT; ++i) {
      const int u = tid + AF_SPLIT_WARPS * 32 * i;
      row[i] = u >> 3;
      const int sub = u & 7, box = sub >> 2, j = sub & 3;
      const int x7 = row[i] & 7;
      const int p0 = (2 * j) ^ x7, p1 = p0 ^ 1;                    // swizzled slots of raw chunks 2j and 2j+1
      // box-1 lanes read the odd chunk first: the 8 lanes of a quarter-warp then touch 8 distinct 16-byte slots
      swap[i] = box != 0;
      const int off = box * (AF_KB_BYTES / 2) + row[i] * 128;
      rd0[i] = off + (box ? p1 : p0) * 16;
      rd1[i] = off + (box ? p0 : p1) * 16;
      wr[i] = row[i] * 128 + ((4 * box + j) ^ x7) * 16;             // slot of fp16 chunk c = 4*box + j in its row
    }
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const uint32_t k = (uint32_t)(t >> 1);
      for (int kb = 0; kb < nkb; ++kb) {
        uint8_t* base = tiles + (size_t)buf * buf_bytes + (size_t)kb * AF_KB_BYTES;
        mbar_wait_warp(smem_u32(full_raw + buf * AF_MAX_KB + kb), k & 1u, 550 + kb);
        if (tid == 0) AF_TR(t, 2 + kb);
        float4 fa[UPT], fb[UPT];
#pragma unroll
        for (int i = 0; i < UPT; ++i) {
          const float4 first = *reinterpret_cast<const float4*>(base + rd0[i]);
          const float4 second = *reinterpret_cast<const float4*>(base + rd1[i]);
          fa[i] = swap[i] ? second : first;                          // raw chunk 2j   (columns 8j .. 8j+3 of the box)
          fb[i] = swap[i] ? first : second;                          // raw chunk 2j+1 (columns 8j+4 .. 8j+7)
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < UPT; ++i) {
          uint4 hi, lo;
          split8(fa[i], fb[i], hi, lo, mabs);
          *reinterpret_cast<uint4*>(base + wr[i]) = hi;
          *reinterpret_cast<uint4*>(base + AF_KB_BYTES / 2 + wr[i]) = lo;
        }
        fence_proxy_async_shared();        // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(full_b + buf * AF_MAX_KB + kb));
        if (tid == 0) AF_TR(t, 8 + kb);
      }
    }
    pdl_wait();      // the weight-preparation kernel cleared the flag
    // out of the fp16 range, or large enough for the fp16 sum of the lo pieces of S rows to overflow
    if (!(mabs <= AF_F16_MAX) || (POOL && !(mabs * (float)p.S <= 1.0e8f))) atomicOr(p.flag, 1);
  } else if (warp >= AF_POOL_WARP0 && warp - AF_POOL_WARP0 < nkb) {
    // =========================== pooling: warp pw owns columns [64 pw, 64 pw + 64) ======================
    const int pw = warp - AF_POOL_WARP0;
    const int c = lane & 7, rs = lane >> 3;        // 16-byte chunk (8 columns) / row subset: rows n = 4g + rs
    const int S = p.S;
    // sum_s a_s x_s with x = (hi + lo) / 16: the hi pieces accumulate in fp32, the lo pieces (2^-11 of hi) in packed
    // fp16 (their rounding is 2^-22 of the sum; the splitters flag inputs large enough to overflow an fp16 sum)
    float acc[8];
    __half2 accl[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) accl[i] = __float2half2_rn(0.f);
    float m_run = -INFINITY, l_run = 0.f;
    int nfin = 0;
    int next_end = S;                               // CTA-local row at which the current buyer ends
    int bl = 0;                                     // current buyer, CTA-local
    const float* wrow = POOL ? p.w + r0 : nullptr;
    const uint32_t tiles_u32 = smem_u32(tiles);
    const int nrows_i = (int)nrows;
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      const uint32_t k = (uint32_t)(t >> 1);
      const int n0 = t * AF_TILE;
      const int nvalid = (nrows_i - n0 < AF_TILE) ? (nrows_i - n0) : AF_TILE;
      float w0 = 0.f, w1 = 0.f;
      if (POOL) {                                   // event weights of the tile's rows: in flight while the logits are computed
        if (lane < nvalid) w0 = __ldg(wrow + n0 + lane);
        if (lane + 32 < nvalid) w1 = __ldg(wrow + n0 + lane + 32);
      }
      mbar_wait_warp(smem_u32(logits_full + buf), k & 1u, 570 + buf);
      if (pw == 0 && lane == 0) AF_TR(t, 23);
      if (POOL) {
        const uint32_t base = tiles_u32 + (uint32_t)buf * (uint32_t)buf_bytes + (uint32_t)pw * AF_KB_BYTES;
        const float z0 = (lane < nvalid) ? logits_s[buf * AF_TILE + lane] * w0 : -INFINITY;
        const float z1 = (lane + 32 < nvalid) ? logits_s[buf * AF_TILE + lane + 32] * w1 : -INFINITY;
        int n_lo = 0;
        while (n_lo < nvalid) {
          const int bend = next_end - n0;                                 // end of the current buyer, tile coordinates
          const int n_hi = bend < nvalid ? bend : nvalid;
          // ---- online softmax step (buyer_tower.py:89-92) -----------------------------------------
          const bool in0 = lane >= n_lo && lane < n_hi, in1 = lane + 32 >= n_lo && lane + 32 < n_hi;
          const float smax = warp_max(fmaxf(in0 ? z0 : -INFINITY, in1 ? z1 : -INFINITY));
          const float m_new = fmaxf(m_run, smax);
          const float scale = (m_run == -INFINITY) ? 0.f : expf(m_run - m_new);
          const float p0 = in0 ? expf(z0 - m_new) : 0.f, p1 = in1 ? expf(z1 - m_new) : 0.f;
          l_run = l_run * scale + warp_sum(p0 + p1);
          m_run = m_new;
          if (scale != 1.f) {
            const __half2 sh = __float2half2_rn(scale);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] *= scale;
#pragma unroll
            for (int i = 0; i < 4; ++i) accl[i] = __hmul2(accl[i], sh);
          }
          // ---- weighted row sum (buyer_tower.py:96) from the hi/lo tiles, two 4-row groups per step ----
          const int g_end = (n_hi - 1) >> 2;
          int g = n_lo >> 2;
          for (; g < g_end; g += 2) {
            const int nA = 4 * g + rs, nB = nA + 4;
            const float pA = __shfl_sync(0xffffffffu, (g < 8) ? p0 : p1, nA & 31);       // 0 outside [n_lo, n_hi)
            const float pB = __shfl_sync(0xffffffffu, (g + 1 < 8) ? p0 : p1, nB & 31);
            const uint32_t aA = base + (uint32_t)(nA * 128 + ((c ^ (nA & 7)) << 4));
            const uint32_t aB = base + (uint32_t)(nB * 128 + ((c ^ (nB & 7)) << 4));
            const uint4 hA = lds128(aA), lA = lds128(aA + AF_KB_BYTES / 2);
            const uint4 hB = lds128(aB), lB = lds128(aB + AF_KB_BYTES / 2);
            fma_chunk(acc, accl, hA, lA, pA);
            fma_chunk(acc, accl, hB, lB, pB);
          }
          if (g == g_end) {
            const int nA = 4 * g + rs;
            const float pA = __shfl_sync(0xffffffffu, (g < 8) ? p0 : p1, nA & 31);
            const uint32_t aA = base + (uint32_t)(nA * 128 + ((c ^ (nA & 7)) << 4));
            const uint4 hA = lds128(aA), lA = lds128(aA + AF_KB_BYTES / 2);
            fma_chunk(acc, accl, hA, lA, pA);
          }
          if (bend <= nvalid) {
            // ---- the buyer is complete: F.normalize(p=2, dim=1, eps=1e-12) (buyer_tower.py:99) -------
            const float inv = 1.0f / (AF_X_SCALE * l_run);
            float ss = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 lf = __half22float2(accl[i >> 1]);
              float a = acc[i] + ((i & 1) ? lf.y : lf.x);
              a += __shfl_xor_sync(0xffffffffu, a, 8);
              a += __shfl_xor_sync(0xffffffffu, a, 16);
              a *= inv;
              acc[i] = a;
              ss = fmaf(a, a, ss);
            }
            ss += __shfl_xor_sync(0xffffffffu, ss, 1);
            ss += __shfl_xor_sync(0xffffffffu, ss, 2);
            ss += __shfl_xor_sync(0xffffffffu, ss, 4);
            float* sq = ssq + (nfin & 1) * 8;
            if (lane == 0) sq[pw] = ss;
            named_bar_sync(2, nkb * 32);
            float tot = 0.f;
            for (int j = 0; j < nkb; ++j) tot += sq[j];
            const float rden = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
            if (rs == 0) {
              float4* op = reinterpret_cast<float4*>(p.out + ((long long)b0 + bl) * p.D + pw * 64 + c * 8);
              op[0] = make_float4(acc[0] * rden, acc[1] * rden, acc[2] * rden, acc[3] * rden);
              op[1] = make_float4(acc[4] * rden, acc[5] * rden, acc[6] * rden, acc[7] * rden);
            }
            ++nfin;
            ++bl;
            next_end += S;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) accl[i] = __float2half2_rn(0.f);
            m_run = -INFINITY;
            l_run = 0.f;
          }
          n_lo = n_hi;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(buf_free + buf));
      if (pw == 0 && lane == 0) AF_TR(t, 24);
    }
  }

  // ---- teardown -----------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == AF_MMA_WARP) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// W1 f32 [H, D] -> fp16 pieces of W1 * 2^kw laid out for the epilogue threads' coalesced loads:
//   Wp[part][kb][q][m] (uint4 = 8 consecutive K elements k = kb*64 + q*8 .. +7 of hidden unit m; part 0 = hi, 1 = lo)
// kw puts max|W1| * 2^kw into [8192, 16384), so that `lo` (2^-11 of `hi`) stays a normal fp16.  Every block reduces
// max|W1| over the whole (L2-resident, 196 KB) matrix itself - cheaper than a second launch - and then converts its
// share: one thread per (hidden unit, 8 consecutive k), i.e. 32-byte coalesced reads.
constexpr int AF_PREP_THREADS = 256;
__global__ void __launch_bounds__(AF_PREP_THREADS)
attn_fused_prep_w_kernel(const float* __restrict__ W1, int H, int D, int nkb, uint4* __restrict__ Wp,
                         float* __restrict__ inv_scale, int* __restrict__ flag) {
  __shared__ float red[AF_PREP_THREADS / 32];
  __shared__ float s_scale;
  float m = 0.f;
  bool bad = false;
  const int n = H * D;
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(W1) & 15) == 0) {
    const float4* w4 = reinterpret_cast<const float4*>(W1);
    for (int i = threadIdx.x; i < (n >> 2); i += blockDim.x) {
      const float4 v = __ldg(w4 + i);
      const float a = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
      if (!(fabsf(v.x) <= 3.0e38f) || !(fabsf(v.y) <= 3.0e38f) || !(fabsf(v.z) <= 3.0e38f) || !(fabsf(v.w) <= 3.0e38f)) bad = true;
      m = fmaxf(m, a);
    }
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float a = fabsf(__ldg(W1 + i));
      if (!(a <= 3.0e38f)) bad = true;
      m = fmaxf(m, a);
    }
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  const int anybad = __syncthreads_or(bad ? 1 : 0);
  if (threadIdx.x == 0) {
    float mm = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) mm = fmaxf(mm, red[i]);
    int e = 0;
    float scale = 1.f;
    if (mm > 0.f && !anybad) { (void)frexpf(mm, &e); scale = ldexpf(1.f, 14 - e); }     // mm = f * 2^e, f in [0.5, 1)
    s_scale = scale;
    if (blockIdx.x == 0) {
      *inv_scale = 1.0f / (scale * AF_X_SCALE);
      *flag = anybad ? 1 : 0;             // also clears the flag for this call
    }
  }
  __syncthreads();
  const float scale = s_scale;
  const int k8n = nkb * 8;                                  // groups of 8 consecutive k per hidden unit
  const int units = AF_M * k8n;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < units; u += gridDim.x * blockDim.x) {
    const int mrow = u / k8n, k8 = u % k8n;
    const int kb = k8 >> 3, q = k8 & 7;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k8 * 8 + j;
      v[j] = (mrow < H && k < D) ? __ldg(W1 + (size_t)mrow * D + k) * scale : 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2 hh = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
      const float2 back = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(v[2 * j] - back.x, v[2 * j + 1] - back.y);
      hi[j] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[j] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    Wp[((size_t)(0 * nkb + kb) * 8 + q) * AF_M + mrow] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    Wp[((size_t)(1 * nkb + kb) * 8 + q) * AF_M + mrow] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// Predicated fp32 recomputation (CUDA cores, one CTA per buyer): runs only when the fused kernel found a value
// outside the fp16 range.  Plain fp32 FMA arithmetic of buyer_tower.py:85-99; speed is irrelevant here.
__global__ void __launch_bounds__(256)
attn_pool_fallback_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ W1,
                          const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                          int H, float* __restrict__ out, int B, int S, int D, const int* __restrict__ flag) {
  pdl_wait();
  if (*flag == 0) return;
  extern __shared__ float fsm[];
  float* lg = fsm;                 // [S]
  float* hsum = fsm + S;           // [8] per-warp partial logits
  __shared__ float bc[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
  const float* xb = x + (long long)b * S * D;
  for (int s = 0; s < S; ++s) {
    const float* row = xb + (long long)s * D;
    float part = 0.f;
    for (int h = warp; h < H; h += 8) {
      float a = 0.f;
      for (int d = lane; d < D; d += 32) a = fmaf(__ldg(row + d), __ldg(W1 + (long long)h * D + d), a);
      a = warp_sum(a);
      part = fmaf(fmaxf(a + __ldg(b1 + h), 0.f), __ldg(W2 + h), part);
    }
    if (lane == 0) hsum[warp] = part;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += hsum[i];
      lg[s] = (t + __ldg(b2)) * __ldg(w + (long long)b * S + s);
    }
    __syncthreads();
  }
  if (tid == 0) {
    float m = -INFINITY, tot = 0.f;
    for (int s = 0; s < S; ++s) m = fmaxf(m, lg[s]);
    for (int s = 0; s < S; ++s) tot += expf(lg[s] - m);
    bc[0] = m; bc[1] = tot;
  }
  __syncthreads();
  const float m = bc[0], tot = bc[1];
  float ss = 0.f;
  for (int d = tid; d < D; d += blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) a = fmaf(__ldg(xb + (long long)s * D + d), expf(lg[s] - m) / tot, a);
    out[(long long)b * D + d] = a;
    ss += a * a;
  }
  ss = warp_sum(ss);
  __syncthreads();
  if (lane == 0) hsum[warp] = ss;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += hsum[i]; bc[0] = fmaxf(sqrtf(t), 1e-12f); }
  __syncthreads();
  const float denom = bc[0];
  for (int d = tid; d < D; d += blockDim.x) out[(long long)b * D + d] /= denom;
  __syncthreads();
  }
}

static int fused_mode() {
  static const int m = [] { const char* e = getenv("TT_B200_ATTN_MODE"); return e ? atoi(e) : 0; }();
  return m;
}

static size_t fused_smem_bytes(int nkb) {
  return 2 * (size_t)nkb * AF_KB_BYTES + (2 * AF_TILE + 2 * 4 * AF_TILE + 16) * sizeof(float) + (4 * AF_MAX_KB + 7) * sizeof(uint64_t) + 16 + 1024;
}

struct FusedWs { size_t wp, meta, total; };
static FusedWs fused_ws_layout(int nkb) {
  FusedWs w{};
  w.wp = 0;
  w.meta = align_up((size_t)2 * nkb * 8 * AF_M * sizeof(uint4), 256);
  w.total = w.meta + 256;
  return w;
}

static bool fused_shape_ok(const float* x, const float* out, long long B, long long S, int D, int H) {
  return D % 64 == 0 && D <= 64 * AF_MAX_KB && H >= 1 && H <= AF_M && S >= 1 && S <= AF_SMAX && B * S >= 64 &&
         B * S < (1LL << 31) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
}

// Logits-only mode (tt_attention_logits; e.g. once over the whole catalog for the gather path): the same kernel without
// the pooling warps, logits written to global memory; the weight pieces live in a stream-ordered scratch block.
int launch_attn_logits_fused(const float* x, long long R, int D, const float* W1, const float* b1, const float* W2,
                             const float* b2, int H, float* logits, cudaStream_t st) {
  if (!(D % 64 == 0 && D <= 64 * AF_MAX_KB && H >= 1 && H <= AF_M && R >= 64 && R < (1LL << 31) &&
        (reinterpret_cast<uintptr_t>(x) & 15) == 0))
    return TT_ERR_UNSUPPORTED;
  int dev = 0;
  TT_CHECK_CUDA(cudaGetDevice(&dev));
  cudaMemPool_t pool = scratch_pool(dev);
  if (!pool) return TT_ERR_UNSUPPORTED;
  const int nkb = D / 64;
  const FusedWs lay = fused_ws_layout(nkb);
  unsigned char* ws = nullptr;
  TT_CHECK_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&ws), lay.total, pool, st));
  uint4* Wp = reinterpret_cast<uint4*>(ws + lay.wp);
  float* inv_scale = reinterpret_cast<float*>(ws + lay.meta);
  int* flag = reinterpret_cast<int*>(ws + lay.meta + 16);
  attn_fused_prep_w_kernel<<<(AF_M * nkb * 8 + AF_PREP_THREADS - 1) / AF_PREP_THREADS, AF_PREP_THREADS, 0, st>>>(W1, H, D, nkb, Wp, inv_scale, flag);
  count_launch();
  int rc = TT_OK;
  CUtensorMap tx;
  if ((rc = make_tmap_f32(&tx, x, R, D, AF_TILE, 32)) == TT_OK) {
    AttnFusedParams p{};
    p.x = x; p.logits_out = logits; p.Wp = Wp; p.b1 = b1; p.W2 = W2; p.b2 = b2; p.inv_scale = inv_scale; p.flag = flag;
    p.R = R; p.B = (int)R; p.S = 1; p.D = D; p.H = H; p.nkb = nkb;
    p.mode = fused_mode();
    const size_t smem = fused_smem_bytes(nkb);
    const long long tiles = (R + AF_TILE - 1) / AF_TILE;
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    cudaError_t ce = cudaFuncSetAttribute(attn_pool_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce == cudaSuccess) { count_launch(); ce = launch_pdl(attn_pool_fused_kernel<false>, dim3(grid), dim3(AF_THREADS), smem, st, tx, p); }
    if (ce != cudaSuccess) { set_error(std::string("launch_attn_logits_fused: ") + cudaGetErrorString(ce)); rc = TT_ERR_CUDA; }
    if (rc == TT_OK) rc = launch_attn_logits_generic_if(x, R, D, W1, b1, W2, b2, H, logits, flag, st);
  }
  const cudaError_t fe = cudaFreeAsync(ws, st);
  if (rc == TT_OK && fe != cudaSuccess) { set_error("launch_attn_logits_fused: cudaFreeAsync failed"); rc = TT_ERR_CUDA; }
  return rc;
}

}  // namespace tt

using namespace tt;

extern "C" __attribute__((visibility("default"))) size_t tt_pool_attention_fused_workspace_bytes(int B, int S, int D, int H) {
  if (B < 1 || S < 1 || D < 1 || H < 1) return 0;
  const int nkb = (D + 63) / 64;
  // the fused path needs the fp16 weight pieces + a flag; other shapes run logits + pooling and need [B,S] logits
  return fused_ws_layout(nkb < 1 ? 1 : nkb).total + align_up((size_t)B * S * sizeof(float), 256);
}

extern "C" __attribute__((visibility("default"))) int tt_pool_attention_fused(const float* x, const float* w, const float* W1, const float* b1,
                                                                 const float* W2, const float* b2, int H, float* out, int B, int S,
                                                                 int D, void* workspace, size_t workspace_bytes, void* stream) {
  TT_CHECK_ARG(x && w && W1 && b1 && W2 && b2 && out && workspace, "null pointer");
  TT_CHECK_ARG(B >= 0 && S >= 1 && D >= 1 && H >= 1, "need B >= 0, S >= 1, D >= 1, H >= 1");
  TT_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  if (B == 0) return TT_OK;
  if (workspace_bytes < tt_pool_attention_fused_workspace_bytes(B, S, D, H)) {
    set_error("tt_pool_attention_fused: workspace too small");
    return TT_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int nkb = (D + 63) / 64;
  const FusedWs lay = fused_ws_layout(nkb);
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  if (!fused_shape_ok(x, out, B, S, D, H)) {
    // shapes outside the fused kernel: logits (tensor cores when aligned, else CUDA cores) + softmax pooling
    float* logits = reinterpret_cast<float*>(ws + lay.total);
    if (int e = tt_attention_logits(x, (int64_t)B * S, D, W1, b1, W2, b2, H, logits, stream)) return e;
    return tt_pool_attention(x, logits, w, out, B, S, D, stream);
  }
  uint4* Wp = reinterpret_cast<uint4*>(ws + lay.wp);
  float* inv_scale = reinterpret_cast<float*>(ws + lay.meta);
  int* flag = reinterpret_cast<int*>(ws + lay.meta + 16);
  attn_fused_prep_w_kernel<<<(AF_M * nkb * 8 + AF_PREP_THREADS - 1) / AF_PREP_THREADS, AF_PREP_THREADS, 0, st>>>(W1, H, D, nkb, Wp, inv_scale, flag);
  TT_CHECK_LAUNCH();
  const size_t smem = fused_smem_bytes(nkb);
  TT_CHECK_CUDA(cudaFuncSetAttribute(attn_pool_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int sms = num_sms();
  CUtensorMap tx;
  if (int e = make_tmap_f32(&tx, x, (long long)B * S, D, AF_TILE, 32)) return e;
  AttnFusedParams p{};
  p.x = x; p.w = w; p.out = out; p.Wp = Wp;
  p.b1 = b1; p.W2 = W2; p.b2 = b2; p.inv_scale = inv_scale; p.flag = flag;
  p.R = (long long)B * S; p.B = B; p.S = S; p.D = D; p.H = H; p.nkb = nkb;
  p.mode = fused_mode();
  const char* trace_path = getenv("TT_B200_ATTN_TRACE");       // debugging aid: synchronises and writes CTA 0's event clocks
  if (trace_path) {
    TT_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&p.trace), AF_TRACE_TILES * AF_TRACE_SLOTS * sizeof(long long)));
    TT_CHECK_CUDA(cudaMemsetAsync(p.trace, 0, AF_TRACE_TILES * AF_TRACE_SLOTS * sizeof(long long), st));
  }
  const int grid = B < sms ? B : sms;
  // programmatic dependent launch: the x stream (TMA, splitters) starts under the weight-preparation kernel's tail;
  // only the epilogue warps (W1 pieces, scale) and the flag wait for it.  The preparation kernel itself is a plain
  // launch, so everything that produced x has completed before either kernel starts.
  count_launch();
  TT_CHECK_CUDA(launch_pdl(attn_pool_fused_kernel<true>, dim3(grid), dim3(AF_THREADS), smem, st, tx, p));
  if (trace_path) {
    static long long host_trace[AF_TRACE_TILES * AF_TRACE_SLOTS];
    TT_CHECK_CUDA(cudaStreamSynchronize(st));
    TT_CHECK_CUDA(cudaMemcpy(host_trace, p.trace, sizeof(host_trace), cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    if (FILE* f = fopen(trace_path, "w")) {
      for (int t = 0; t < AF_TRACE_TILES; ++t) {
        for (int i = 0; i < AF_TRACE_SLOTS; ++i) fprintf(f, "%lld ", host_trace[t * AF_TRACE_SLOTS + i]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  const size_t fsm = (size_t)(S + 8) * sizeof(float);
  if (fsm > 48 * 1024)
    TT_CHECK_CUDA(cudaFuncSetAttribute(attn_pool_fallback_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
  count_launch();
  const int fgrid = B < num_sms() * 8 ? B : num_sms() * 8;      // grid-stride: an all-exit launch stays a few microseconds
  TT_CHECK_CUDA(launch_pdl(attn_pool_fallback_kernel, dim3(fgrid), dim3(256), fsm, st, x, w, W1, b1, W2, b2, H, out, B, S, D,
                           (const int*)flag));
  return TT_OK;
}
