// Fused attention pooling of the buyer tower (reference: src/models/buyer_tower.py:70-101), CTA-pair version:
//     logit_s = W2 . relu(W1 x_s + b1) + b2 ;  a = softmax_s(logit * w) ;  out = normalize(sum_s a_s x_s)
// ONE kernel, ONE pass over x in HBM: the score MLP runs on the tensor cores while the rows stream in, and the
// softmax-weighted sum re-reads the same rows a few microseconds later from L2 (the TMA producer is throttled so that
// the rows between "first touched" and "pooled" stay a fraction of L2).
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
// adds the two in fp32.
// A value outside the fp16 range after scaling (|x| > 4094, or a non-finite input) raises a device flag and a
// predicated fp32 CUDA-core kernel recomputes the call: no host synchronisation, always the fp32 answer.
//
// Why a CTA pair.  The previous version kept W1 (hi + lo = 196 KB for 128 x 384) in tensor memory as the A operand
// (tcgen05.mma with a TMEM A operand, N = 64 rows of x): measured 100 cycles per MMA, 1.6 cycles per accumulator
// column, a third of the tensor rate - 85 us of MMA time at C2 on top of a pipeline whose per-K-block hand-offs cost
// as much again.  Both operands from shared memory run at the full rate, but W1 does not fit one SM next to the x
// tiles.  A cluster of two CTAs issues cta_group::2 MMAs with M = 256 (128 rows of x from EACH CTA's shared memory),
// N = 128 hidden units whose W1 rows are SPLIT between the two CTAs (64 each: 96 KB of hi + lo per CTA, loaded once),
// K = 16: 6 KB of operand reads per SM per MMA = 48 cycles, 3 MMAs per K step, 3.5k cycles per 256 rows.
// The two CTAs of a pair own disjoint buyer ranges (each pools its own buyers); only the MMA is shared.
//
// Per CTA (24 warps), rows in tiles of 128, K-blocks of 64 columns through a 3-stage ring of 32 KB:
//   warp 0 lane 0   : TMA producer  - raw fp32 [128 rows x 64 cols] per K-block as two 128B-swizzled boxes [128 x 32]
//   warps 8-15      : splitters     - rewrite the landed K-block IN PLACE: 8 lanes read one row's 256 raw bytes, then
//                                     write `hi` over the first box and `lo` over the second (UMMA K-major 128B-swizzle
//                                     layout), fence.proxy.async, arrive on the LEADER's barrier
//   warp 1 lane 0   : MMA issuer (leader CTA) - per K-block 4 K-steps x 3 MMAs into the main and the cross-term
//                                     accumulator (2 x 128 TMEM columns, double-buffered); tcgen05.commit multicast frees
//                                     the stage in both CTAs
//   warps 4-7       : epilogue      - lane = row: logit = b2 + sum_h W2[h] relu(acc[h] 2^-e + b1[h]) straight from the
//                                     accumulator columns, no cross-lane reduction
//   warps 16-23     : pooling       - one warp per buyer, as soon as the tile holding the buyer's last row is done:
//                                     softmax of logit*weight, weighted row sum (128-bit loads, L2 hits), L2 normalise
//   warp 2          : TMEM allocator
// Rooflines: HBM (x once: B*S*D*4 bytes); L2 -> SM x twice; tensor pipe 3 x 2*D*128 flop per row.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include "tt_common.cuh"
#include "sm100_ptx.cuh"
#include "flat_internal.cuh"

namespace tt {

using namespace ptx;

constexpr int AP_ROWS = 128;                         // rows of x per CTA per tile (UMMA M = 256 over the pair)
constexpr int AP_H = 128;                            // hidden units (UMMA N); H <= 128, 64 rows of W1 per CTA
constexpr int AP_STAGE = AP_ROWS * 64 * 4;           // 32 KB: raw fp32 K-block -> hi 16 KB | lo 16 KB
constexpr int AP_STAGES = 3;
constexpr int AP_W_KB = 2 * (AP_H / 2) * 128;        // 16 KB per K-block: this CTA's 64 rows of W hi (8 KB) | W lo (8 KB)
constexpr int AP_MAX_KB = 6;                         // D <= 384
constexpr int AP_RMAX = 6144;                        // rows (logits) one CTA may own per launch
constexpr int AP_SPLIT_WARP0 = 8, AP_SPLIT_WARPS = 8;
constexpr int AP_POOL_WARP0 = 16, AP_POOL_WARPS = 8;
constexpr int AP_WARPS = 24;
constexpr int AP_THREADS = AP_WARPS * 32;
constexpr int AP_LAG = 3;                            // tiles the x stream may run ahead of the pooling re-read
constexpr float AP_X_SCALE = 16.0f;
constexpr float AP_F16_MAX = 65504.0f;

struct AttnPairParams {
  const float* x;          // [R, D]
  const float* w;          // [B, S]       (pool mode)
  float* out;              // [B, D]       (pool mode)
  float* logits_out;       // [R]          (logits-only mode)
  const float* b1;
  const float* W2;
  const float* b2;
  const float* inv_scale;  // 2^-(kw + 4), written by the weight-preparation kernel
  int* flag;               // raised when a value leaves the fp16 range
  long long R;
  int B, S, D, H, nkb;
  int mode;                // TT_B200_ATTN_MODE, timing experiments only (wrong results): 8 no MMAs, 16 no throttle
  long long* trace;        // TT_B200_ATTN_TRACE: clock64 of CTA 0's pipeline events, [AP_TRACE_TILES][AP_TRACE_SLOTS]
};
constexpr int AP_TRACE_TILES = 16, AP_TRACE_SLOTS = 40;
#define AP_TR(t, slot) do { if (p.trace && blockIdx.x == 0 && (t) < AP_TRACE_TILES) p.trace[(t) * AP_TRACE_SLOTS + (slot)] = clock64(); } while (0)

__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_cta_shared_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_cta_shared_add_u32(unsigned int* p, unsigned int v) {
  asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// arrive (release at cluster scope) on the barrier at the same offset in CTA `cta` of the pair
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 remAddr32;\n\t"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [remAddr32];\n\t}"
      ::"r"(bar), "r"(cta)
      : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity, int tag) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("tt_b200: mbarrier watchdog: block %d thread %d tag %d parity %u\n", (int)blockIdx.x, (int)threadIdx.x, tag, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mma_f16_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 8 consecutive fp32 -> 8 fp16 `hi` + 8 fp16 `lo` of the scaled values
__device__ __forceinline__ void split8_pair(const float4 a, const float4 b, uint4& hi, uint4& lo, float& mabs) {
  const float v[8] = {a.x * AP_X_SCALE, a.y * AP_X_SCALE, a.z * AP_X_SCALE, a.w * AP_X_SCALE,
                      b.x * AP_X_SCALE, b.y * AP_X_SCALE, b.z * AP_X_SCALE, b.w * AP_X_SCALE};
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

template <bool POOL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(AP_THREADS, 1)
attn_pool_pair_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const AttnPairParams p) {
  extern __shared__ uint8_t smem_raw_[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_) + 1023) & ~(uintptr_t)1023);
  const int nkb = p.nkb;
  uint8_t* w_tiles = smem;                                                        // [nkb][16 KB]  W hi | W lo (this CTA's 64 rows)
  uint8_t* ring = w_tiles + (size_t)AP_MAX_KB * AP_W_KB;                          // [AP_STAGES][32 KB]
  float* logits_s = reinterpret_cast<float*>(ring + (size_t)AP_STAGES * AP_STAGE); // [AP_RMAX]  (pool mode)
  float2* b1w2 = reinterpret_cast<float2*>(logits_s + AP_RMAX);                   // [128] {b1[h], W2[h]}
  uint64_t* bars = reinterpret_cast<uint64_t*>(b1w2 + AP_H);
  uint64_t* full_raw = bars;                          // [AP_STAGES]  local: TMA landed
  uint64_t* full_b = full_raw + AP_STAGES;            // [AP_STAGES]  leader: one arrival per splitter warp of both CTAs
  uint64_t* empty = full_b + AP_STAGES;               // [AP_STAGES]  both: tcgen05.commit multicast
  uint64_t* acc_full = empty + AP_STAGES;             // [2]  both: tcgen05.commit multicast
  uint64_t* acc_empty = acc_full + 2;                 // [2]  leader: one arrival per epilogue warp of both CTAs
  uint64_t* w_full = acc_empty + 2;                   // [1]  leader: W tiles of both CTAs landed
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_full + 1);
  unsigned int* done_cnt = tmem_ptr_smem + 1;         // += 1 per epilogue warp per finished tile
  unsigned int* pooled_cnt = done_cnt + 1;            // += 1 per pooled buyer

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;

  // ---- this pair's buyers, split between its two CTAs ------------------------------------------------
  const int npairs = (int)(gridDim.x >> 1), pair = (int)(blockIdx.x >> 1);
  const int bp0 = (int)((long long)pair * p.B / npairs);
  const int bp1 = (int)((long long)(pair + 1) * p.B / npairs);
  const int bmid = bp0 + (bp1 - bp0 + 1) / 2;
  const int b0 = leader ? bp0 : bmid, b1 = leader ? bmid : bp1;
  const long long r0 = (long long)b0 * p.S;
  const int nrows = (b1 - b0) * p.S;                                    // <= AP_RMAX in pool mode (host)
  const int nrows_a = (bmid - bp0) * p.S, nrows_b = (bp1 - bmid) * p.S;
  const int ntiles = (max(nrows_a, nrows_b) + AP_ROWS - 1) / AP_ROWS;   // the pair walks in lockstep

  if (warp == 0 && lane == 0) { prefetch_tensormap(&tmap_x); prefetch_tensormap(&tmap_w); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < AP_STAGES; ++i) {
      mbar_init(smem_u32(full_raw + i), 1);
      mbar_init(smem_u32(full_b + i), 2 * AP_SPLIT_WARPS);
      mbar_init(smem_u32(empty + i), 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(acc_full + i), 1); mbar_init(smem_u32(acc_empty + i), 8); }
    mbar_init(smem_u32(w_full), 2);
    *done_cnt = 0u;
    *pooled_cnt = 0u;
    fence_barrier_init();
  }
  cluster_sync();          // peer barriers must be initialised before any remote arrive / 2-SM alloc
  if (warp == 2) { tmem_alloc_2cta(smem_u32(tmem_ptr_smem), 512); tmem_relinquish_2cta(); }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // =========================== TMA producer =====================================================
    if (lane == 0) {
      pdl_wait();      // the weight-preparation kernel (previous launch) has written the W pieces
      // this CTA's 64 rows of W hi and W lo per K-block; the bytes of both CTAs are counted on the leader's barrier
      const uint32_t wb = smem_u32(w_full);
      if (leader) mbar_arrive_expect_tx(wb, (uint32_t)(2 * nkb * AP_W_KB));
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t dst = smem_u32(w_tiles + (size_t)kb * AP_W_KB);
        tma_load_2d_2cta(dst, &tmap_w, wb, kb * 64, (int)cta_rank * (AP_H / 2));
        tma_load_2d_2cta(dst + AP_W_KB / 2, &tmap_w, wb, kb * 64, AP_H + (int)cta_rank * (AP_H / 2));
      }
      if (!leader) mbar_arrive_remote(wb, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        if (POOL && !(p.mode & 16) && t > AP_LAG) {
          // throttle: the rows of tile t - AP_LAG and earlier should have been pooled (approximately: the pooling warps
          // finish buyers out of order) before more rows enter L2
          const int rows_before = (t - AP_LAG) * AP_ROWS;
          int need = rows_before / p.S - AP_POOL_WARPS;
          const int nb = b1 - b0;
          if (need > nb) need = nb;
          while (need > 0 && (int)ld_acquire_cta_shared_u32(pooled_cnt) < need) __nanosleep(64);
        }
        const long long row = r0 + (long long)t * AP_ROWS;          // rows past the tensor are zero-filled
        const int rowc = (int)(row < p.R ? row : p.R);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(empty + stage), phase ^ 1, 500 + stage);
          const uint32_t fb = smem_u32(full_raw + stage);
          const uint32_t dst = smem_u32(ring + (size_t)stage * AP_STAGE);
          mbar_arrive_expect_tx(fb, (uint32_t)AP_STAGE);
          tma_load_2d(dst, &tmap_x, fb, kb * 64, rowc);
          tma_load_2d(dst + AP_STAGE / 2, &tmap_x, fb, kb * 64 + 32, rowc);
          AP_TR(t, kb);
          if (++stage == AP_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA) ==========================================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc_f16_f32(2 * AP_ROWS, AP_H);
      mbar_wait(smem_u32(w_full), 0, 510);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        mbar_wait(smem_u32(acc_empty + buf), (((uint32_t)t >> 1) & 1u) ^ 1u, 520 + buf);
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)(buf * 2 * AP_H), d_cross = d_main + (uint32_t)AP_H;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(smem_u32(full_b + stage), phase, 530 + stage);
          AP_TR(t, 18 + kb);
          tc_fence_after();
          const uint32_t a_base = smem_u32(ring + (size_t)stage * AP_STAGE);
          const uint32_t b_base = smem_u32(w_tiles + (size_t)kb * AP_W_KB);
          const uint64_t xh = make_smem_desc_sw128(a_base), xl = make_smem_desc_sw128(a_base + AP_STAGE / 2);
          const uint64_t wh = make_smem_desc_sw128(b_base), wl = make_smem_desc_sw128(b_base + AP_W_KB / 2);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t koff = (uint64_t)((ks * 16 * 2) >> 4);               // 32 bytes per K = 16 step
            const uint32_t first = (uint32_t)((kb | ks) != 0);
            if (!(p.mode & 8)) {
              mma_f16_ss_2cta(d_main, xh + koff, wh + koff, idesc, first);        // hi_x.hi_w -> main accumulator
              mma_f16_ss_2cta(d_cross, xh + koff, wl + koff, idesc, first);       // hi_x.lo_w
              mma_f16_ss_2cta(d_cross, xl + koff, wh + koff, idesc, 1u);          // lo_x.hi_w -> cross accumulator
            }
          }
          mma_commit_2cta(smem_u32(empty + stage));
          AP_TR(t, 24 + kb);
          if (kb == nkb - 1) mma_commit_2cta(smem_u32(acc_full + buf));
          if (++stage == AP_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4 && warp < 8) {
    // =========================== epilogue: lane = row =============================================
    const int quad = warp & 3;
    const int rl_in_tile = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    pdl_wait();      // inv_scale comes from the weight-preparation kernel
    for (int h = (warp - 4) * 32 + lane; h < AP_H; h += 128)
      b1w2[h] = (h < p.H) ? make_float2(__ldg(p.b1 + h), __ldg(p.W2 + h)) : make_float2(0.f, 0.f);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const float sc = __ldg(p.inv_scale);
    const float b2v = __ldg(p.b2);
    for (int t = 0; t < ntiles; ++t) {
      const int buf = t & 1;
      mbar_wait(smem_u32(acc_full + buf), ((uint32_t)t >> 1) & 1u, 540 + buf);
      if (warp == 4 && lane == 0) AP_TR(t, 30);
      tc_fence_after();
      const uint32_t col0 = (uint32_t)(buf * 2 * AP_H);
      float logit = b2v;
#pragma unroll 1
      for (int cch = 0; cch < AP_H / 32; ++cch) {
        uint32_t v[32], c[32];
        __syncwarp();
        tmem_ld_32x32(lane_base + col0 + (uint32_t)(cch * 32), v);                     // main accumulator
        tmem_ld_32x32(lane_base + col0 + (uint32_t)(AP_H + cch * 32), c);              // cross terms
        tmem_ld_wait();
        if (cch == AP_H / 32 - 1) {                    // everything is in registers: hand the accumulators back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (leader) mbar_arrive(smem_u32(acc_empty + buf)); else mbar_arrive_remote(smem_u32(acc_empty + buf), 0); }
        }
        const float4* bw = reinterpret_cast<const float4*>(b1w2 + cch * 32);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float4 q = bw[i >> 1];                   // {b1[h], W2[h], b1[h+1], W2[h+1]} (broadcast)
          const float a0 = fmaf(__uint_as_float(v[i]) + __uint_as_float(c[i]), sc, q.x);
          const float a1 = fmaf(__uint_as_float(v[i + 1]) + __uint_as_float(c[i + 1]), sc, q.z);
          logit = fmaf(fmaxf(a0, 0.f), q.y, logit);
          logit = fmaf(fmaxf(a1, 0.f), q.w, logit);
        }
      }
      const int rl = t * AP_ROWS + rl_in_tile;
      if (rl < nrows) {
        if (POOL) logits_s[rl] = logit;
        else p.logits_out[r0 + rl] = logit;
      }
      if (POOL) {
        __syncwarp();
        if (lane == 0) red_release_cta_shared_add_u32(done_cnt, 1u);
      }
      if (warp == 4 && lane == 0) AP_TR(t, 31);
    }
  } else if (warp >= AP_SPLIT_WARP0 && warp < AP_POOL_WARP0) {
    // =========================== splitters: raw fp32 K-block -> hi | lo fp16, in place =================
    const int tid = (warp - AP_SPLIT_WARP0) * 32 + lane;          // 0..255
    float mabs = 0.f;
    constexpr int UPT = (AP_ROWS * 8) / (AP_SPLIT_WARPS * 32);    // 32-byte units per thread per K-block (4)
    // the 8 lanes of a row sit in one warp and handle the same unit index i: a row is read completely (all lanes of
    // the load instructions) before any lane overwrites it
    int rd0[UPT], rd1[UPT], wr[UPT];
    const int sub = tid & 7, box = sub >> 2, j = sub & 3;
#pragma unroll
    for (int i = 0; i < UPT; ++i) {
      const int row = (tid >> 3) + 32 * i;
      const int x7 = row & 7;
      const int p0 = (2 * j) ^ x7, p1 = p0 ^ 1;                    // swizzled slots of raw chunks 2j and 2j+1
      // box-1 lanes read the odd chunk first: the 8 lanes of a quarter-warp then touch 8 distinct 16-byte slots
      const int off = box * (AP_STAGE / 2) + row * 128;
      rd0[i] = off + (box ? p1 : p0) * 16;
      rd1[i] = off + (box ? p0 : p1) * 16;
      wr[i] = row * 128 + ((4 * box + j) ^ x7) * 16;               // slot of fp16 chunk c = 4*box + j in its row
    }
    int stage = 0;
    uint32_t phase = 0;
    const int nsteps = ntiles * nkb;
    for (int n = 0; n < nsteps; ++n) {
      uint8_t* base = ring + (size_t)stage * AP_STAGE;
      mbar_wait(smem_u32(full_raw + stage), phase, 550 + stage);
      if (tid == 0) AP_TR(n / nkb, 6 + n % nkb);
      float4 fa[UPT], fb[UPT];
#pragma unroll
      for (int i = 0; i < UPT; ++i) {
        const float4 first = *reinterpret_cast<const float4*>(base + rd0[i]);
        const float4 second = *reinterpret_cast<const float4*>(base + rd1[i]);
        fa[i] = box ? second : first;                              // raw chunk 2j   (columns 8j .. 8j+3 of the box)
        fb[i] = box ? first : second;                              // raw chunk 2j+1 (columns 8j+4 .. 8j+7)
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < UPT; ++i) {
        uint4 hi, lo;
        split8_pair(fa[i], fb[i], hi, lo, mabs);
        *reinterpret_cast<uint4*>(base + wr[i]) = hi;
        *reinterpret_cast<uint4*>(base + AP_STAGE / 2 + wr[i]) = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(smem_u32(full_b + stage));
        else mbar_arrive_remote(smem_u32(full_b + stage), 0);
      }
      if (tid == 0) AP_TR(n / nkb, 12 + n % nkb);
      if (++stage == AP_STAGES) { stage = 0; phase ^= 1; }
    }
    pdl_wait();      // the weight-preparation kernel cleared the flag
    if (!(mabs <= AP_F16_MAX)) atomicOr(p.flag, 1);
  } else if (POOL && warp >= AP_POOL_WARP0) {
    // =========================== pooling: one warp per buyer ==========================================
    const int pw = warp - AP_POOL_WARP0;
    const int S = p.S, D = p.D;
    const int nvalid4 = D >> 2;
    constexpr int NV = 3, U = 2;
    for (int b = b0 + pw; b < b1; b += AP_POOL_WARPS) {
      const int rl0 = (b - b0) * S;
      const unsigned int need = 4u * (unsigned int)((rl0 + S - 1) / AP_ROWS + 1);
      if (lane == 0) {
        while (ld_acquire_cta_shared_u32(done_cnt) < need) __nanosleep(64);
      }
      __syncwarp();
      const float* lg = logits_s + rl0;
      const float* wb = p.w + (long long)b * S;
      // ---- softmax normaliser (buyer_tower.py:89-92) -----------------------------------------------
      float m = -INFINITY, tot = 0.f;
      for (int s = lane; s < S; s += 32) m = fmaxf(m, lg[s] * __ldg(wb + s));
      m = warp_max(m);
      for (int s = lane; s < S; s += 32) tot += expf(lg[s] * __ldg(wb + s) - m);
      tot = warp_sum(tot);
      // ---- weighted row sum (buyer_tower.py:96) ----------------------------------------------------
      float4 acc[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* xb = reinterpret_cast<const float4*>(p.x + ((long long)b * S) * D);
      for (int s0 = 0; s0 < S; s0 += 32) {
        const int s = s0 + lane;
        const float coef = (s < S) ? expf(lg[s] * __ldg(wb + s) - m) / tot : 0.f;
        const int nrow = min(32, S - s0);
        for (int j0 = 0; j0 < nrow; j0 += U) {
          float4 buf[U][NV];
          float cf[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int jj = j0 + u;
            const bool ok = jj < nrow;
            cf[u] = ok ? __shfl_sync(0xffffffffu, coef, ok ? jj : 0) : 0.f;
            const float4* rp = xb + (long long)(s0 + (ok ? jj : 0)) * nvalid4;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              const int c4 = v * 32 + lane;
              buf[u][v] = (ok && c4 < nvalid4) ? ldg_stream(rp + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              acc[v].x = fmaf(buf[u][v].x, cf[u], acc[v].x);
              acc[v].y = fmaf(buf[u][v].y, cf[u], acc[v].y);
              acc[v].z = fmaf(buf[u][v].z, cf[u], acc[v].z);
              acc[v].w = fmaf(buf[u][v].w, cf[u], acc[v].w);
            }
          }
        }
      }
      // ---- F.normalize(p=2, dim=1, eps=1e-12) (buyer_tower.py:99) ------------------------------------
      float ss = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        ss += acc[v].x * acc[v].x + acc[v].y * acc[v].y + acc[v].z * acc[v].z + acc[v].w * acc[v].w;
      ss = warp_sum(ss);
      const float denom = fmaxf(sqrtf(ss), 1e-12f);
      float4* op = reinterpret_cast<float4*>(p.out + (long long)b * D);
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c4 = v * 32 + lane;
        if (c4 < nvalid4) op[c4] = make_float4(acc[v].x / denom, acc[v].y / denom, acc[v].z / denom, acc[v].w / denom);
      }
      __syncwarp();
      if (lane == 0) red_release_cta_shared_add_u32(pooled_cnt, 1u);
    }
  }

  // ---- teardown -----------------------------------------------------------------------------------
  tc_fence_before();
  cluster_sync();
  if (warp == 2) { tc_fence_after(); tmem_dealloc_2cta(tmem_base, 512); }
}

// W1 f32 [H, D] -> fp16 pieces of W1 * 2^kw as two row-major [128, D] matrices (hi rows 0-127, lo rows 128-255; rows
// h >= H are zero): the TMA boxes [64 rows x 64 cols] of the pair kernel land them in the UMMA K-major layout.
// kw puts max|W1| * 2^kw into [8192, 16384), so that `lo` (2^-11 of `hi`) stays a normal fp16.  Every block reduces
// max|W1| over the whole (L2-resident, 196 KB) matrix itself - cheaper than a second launch.
constexpr int AP_PREP_THREADS = 256;
__global__ void __launch_bounds__(AP_PREP_THREADS)
attn_pair_prep_w_kernel(const float* __restrict__ W1, int H, int D, __half* __restrict__ Wp, float* __restrict__ inv_scale,
                        int* __restrict__ flag) {
  __shared__ float red[AP_PREP_THREADS / 32];
  __shared__ float s_scale;
  float m = 0.f;
  bool bad = false;
  const int n = H * D;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float a = fabsf(__ldg(W1 + i));
    if (!(a <= 3.0e38f)) bad = true;
    m = fmaxf(m, a);
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
      *inv_scale = 1.0f / (scale * AP_X_SCALE);
      *flag = anybad ? 1 : 0;             // also clears the flag for this call
    }
  }
  __syncthreads();
  const float scale = s_scale;
  const int total = AP_H * D;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int h = i / D;
    const float v = (h < H) ? __ldg(W1 + i) * scale : 0.f;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    Wp[i] = hi;
    Wp[total + i] = lo;
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

static size_t pair_smem_bytes() {
  return (size_t)AP_MAX_KB * AP_W_KB + (size_t)AP_STAGES * AP_STAGE + AP_RMAX * sizeof(float) + AP_H * sizeof(float2) +
         (3 * AP_STAGES + 5) * sizeof(uint64_t) + 16 + 1024;
}

struct FusedWs { size_t wp, meta, total; };
static FusedWs fused_ws_layout(int D) {
  FusedWs w{};
  w.wp = 0;
  w.meta = align_up((size_t)2 * AP_H * D * sizeof(__half), 256);
  w.total = w.meta + 256;
  return w;
}

static bool fused_shape_ok(const float* x, const float* out, long long B, long long S, int D, int H) {
  return D % 64 == 0 && D <= 64 * AP_MAX_KB && H >= 1 && H <= AP_H && S >= 1 && S <= AP_RMAX && B * S >= 256 &&
         B * S < (1LL << 31) && num_sms() >= 2 && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
         ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
}

static int prep_weights(const float* W1, int H, int D, __half* Wp, float* inv_scale, int* flag, cudaStream_t st) {
  const int blocks = (AP_H * D + AP_PREP_THREADS * 4 - 1) / (AP_PREP_THREADS * 4);
  attn_pair_prep_w_kernel<<<blocks, AP_PREP_THREADS, 0, st>>>(W1, H, D, Wp, inv_scale, flag);
  TT_CHECK_LAUNCH();
  return TT_OK;
}

// Logits-only mode (tt_attention_logits; e.g. once over the whole catalog for the gather path): the same kernel without
// the pooling warps, logits written to global memory; the weight pieces live in a stream-ordered scratch block.
int launch_attn_logits_fused(const float* x, long long R, int D, const float* W1, const float* b1, const float* W2,
                             const float* b2, int H, float* logits, cudaStream_t st) {
  if (!(D % 64 == 0 && D <= 64 * AP_MAX_KB && H >= 1 && H <= AP_H && R >= 256 && R < (1LL << 31) && num_sms() >= 2 &&
        (reinterpret_cast<uintptr_t>(x) & 15) == 0))
    return TT_ERR_UNSUPPORTED;
  int dev = 0;
  TT_CHECK_CUDA(cudaGetDevice(&dev));
  cudaMemPool_t pool = scratch_pool(dev);
  if (!pool) return TT_ERR_UNSUPPORTED;
  const int nkb = D / 64;
  const FusedWs lay = fused_ws_layout(D);
  unsigned char* ws = nullptr;
  TT_CHECK_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void**>(&ws), lay.total, pool, st));
  __half* Wp = reinterpret_cast<__half*>(ws + lay.wp);
  float* inv_scale = reinterpret_cast<float*>(ws + lay.meta);
  int* flag = reinterpret_cast<int*>(ws + lay.meta + 16);
  int rc = prep_weights(W1, H, D, Wp, inv_scale, flag, st);
  CUtensorMap tx, tw;
  if (rc == TT_OK) rc = make_tmap_f32(&tx, x, R, D, AP_ROWS, 32);
  if (rc == TT_OK) rc = make_tmap_16bit(&tw, Wp, 2 * AP_H, D, AP_H / 2);
  if (rc == TT_OK) {
    AttnPairParams p{};
    p.x = x; p.logits_out = logits; p.b1 = b1; p.W2 = W2; p.b2 = b2; p.inv_scale = inv_scale; p.flag = flag;
    p.R = R; p.B = (int)R; p.S = 1; p.D = D; p.H = H; p.nkb = nkb;
    p.mode = fused_mode();
    const size_t smem = pair_smem_bytes();
    const long long tiles = (R + 2 * AP_ROWS - 1) / (2 * AP_ROWS);
    const int npairs = (int)(tiles < num_sms() / 2 ? tiles : num_sms() / 2);
    cudaError_t ce = cudaFuncSetAttribute(attn_pool_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ce == cudaSuccess) { count_launch(); ce = launch_pdl(attn_pool_pair_kernel<false>, dim3(2 * npairs), dim3(AP_THREADS), smem, st, tx, tw, p); }
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
  // the fused path needs the fp16 weight pieces + a flag; other shapes run logits + pooling and need [B,S] logits
  return fused_ws_layout((D + 63) / 64 * 64).total + align_up((size_t)B * S * sizeof(float), 256);
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
  const FusedWs lay = fused_ws_layout((D + 63) / 64 * 64);
  unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
  if (!fused_shape_ok(x, out, B, S, D, H)) {
    // shapes outside the fused kernel: logits (tensor cores when aligned, else CUDA cores) + softmax pooling
    float* logits = reinterpret_cast<float*>(ws + lay.total);
    if (int e = tt_attention_logits(x, (int64_t)B * S, D, W1, b1, W2, b2, H, logits, stream)) return e;
    return tt_pool_attention(x, logits, w, out, B, S, D, stream);
  }
  const int nkb = D / 64;
  __half* Wp = reinterpret_cast<__half*>(ws + lay.wp);
  float* inv_scale = reinterpret_cast<float*>(ws + lay.meta);
  int* flag = reinterpret_cast<int*>(ws + lay.meta + 16);
  if (int e = prep_weights(W1, H, D, Wp, inv_scale, flag, st)) return e;
  const size_t smem = pair_smem_bytes();
  TT_CHECK_CUDA(cudaFuncSetAttribute(attn_pool_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CUtensorMap tw;
  if (int e = make_tmap_16bit(&tw, Wp, 2 * AP_H, D, AP_H / 2)) return e;
  // every launch covers at most npairs * 2 * floor(AP_RMAX / S) buyers (a CTA keeps its rows' logits in shared memory)
  const int max_pairs = num_sms() / 2;
  const long long per_cta = AP_RMAX / S;
  for (long long bdone = 0; bdone < B;) {
    const long long nb = (B - bdone < 2 * per_cta * max_pairs) ? (B - bdone) : 2 * per_cta * max_pairs;
    CUtensorMap tx;
    if (int e = make_tmap_f32(&tx, x + bdone * S * D, nb * S, D, AP_ROWS, 32)) return e;
    AttnPairParams p{};
    p.x = x + bdone * S * D; p.w = w + bdone * S; p.out = out + bdone * D;
    p.b1 = b1; p.W2 = W2; p.b2 = b2; p.inv_scale = inv_scale; p.flag = flag;
    p.R = nb * S; p.B = (int)nb; p.S = S; p.D = D; p.H = H; p.nkb = nkb;
    p.mode = fused_mode();
    const char* trace_path = getenv("TT_B200_ATTN_TRACE");       // debugging aid: synchronises and writes CTA 0's event clocks
    if (trace_path) {
      TT_CHECK_CUDA(cudaMalloc(reinterpret_cast<void**>(&p.trace), AP_TRACE_TILES * AP_TRACE_SLOTS * sizeof(long long)));
      TT_CHECK_CUDA(cudaMemsetAsync(p.trace, 0, AP_TRACE_TILES * AP_TRACE_SLOTS * sizeof(long long), st));
    }
    // enough pairs that no CTA owns more than per_cta buyers, no more than the rows fill
    long long npairs = (nb + 2 * per_cta - 1) / (2 * per_cta);
    const long long by_rows = (nb * S + 2 * AP_ROWS - 1) / (2 * AP_ROWS);
    if (npairs < by_rows) npairs = by_rows;
    if (npairs > nb) npairs = nb;
    if (npairs > max_pairs) npairs = max_pairs;
    // programmatic dependent launch: the kernel's set-up runs under the weight-preparation kernel's tail; the producer,
    // the epilogue warps (scale) and the flag wait for it.  The preparation kernel itself is a plain launch, so
    // everything that produced x has completed before either kernel starts.
    count_launch();
    TT_CHECK_CUDA(launch_pdl(attn_pool_pair_kernel<true>, dim3((unsigned)(2 * npairs)), dim3(AP_THREADS), smem, st, tx, tw, p));
    if (trace_path) {
      static long long host_trace[AP_TRACE_TILES * AP_TRACE_SLOTS];
      TT_CHECK_CUDA(cudaStreamSynchronize(st));
      TT_CHECK_CUDA(cudaMemcpy(host_trace, p.trace, sizeof(host_trace), cudaMemcpyDeviceToHost));
      cudaFree(p.trace);
      if (FILE* f = fopen(trace_path, "w")) {
        for (int t = 0; t < AP_TRACE_TILES; ++t) {
          for (int i = 0; i < AP_TRACE_SLOTS; ++i) fprintf(f, "%lld ", host_trace[t * AP_TRACE_SLOTS + i]);
          fprintf(f, "\n");
        }
        fclose(f);
      }
    }
    bdone += nb;
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
