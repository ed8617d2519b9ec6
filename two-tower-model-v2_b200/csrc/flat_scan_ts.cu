// Tensor-core catalog scan for large query batches with the QUERIES IN TENSOR MEMORY (the hot loop of
// faiss.IndexFlatIP.search as called at src/inference/vector_db.py:197 with nq > 256).
//
// The shared-memory-operand kernel (flat_scan.cu) keeps one 128-query block per CTA resident in shared memory:
// every catalog tile is then re-read from L2 by each of the nq/256 query units (L2->SM traffic = 16 B/cycle/SM x 2,
// 5.6 TB/s at nq = 4096, the ceiling this kernel hit), and every MMA streams its A operand from shared memory.
// Here a CTA pair holds NQB query blocks of 256 queries (128 per CTA) in TENSOR MEMORY (tcgen05.st, once) and
// issues tcgen05.mma.cta_group::2 with A from TMEM:
//   * twice the queries per catalog byte: L2->SM traffic halves (NQB = 2), and so does the power that goes with it;
//   * shared memory only holds catalog tiles (a deep ring of whole 64-row tiles) and is read at 32 B/cycle;
//   * TMEM: 2 accumulator slots x 64 columns + NQB x Dp/2 query columns (D = 384: 128 + 2 x 192 = 512).
// Work item j = (tile t, block b), b fastest; accumulator slot j & 1: the epilogue of item j overlaps the MMAs of
// item j + 1, which for NQB = 2 read the SAME catalog tile again (it stays in its ring slot until block NQB-1 is done).
//
// CTA anatomy (256 threads, 2-CTA cluster): warp 0 lane 0 TMA producer (one mbarrier per 64-row tile: num_kb boxes
// of [32 rows x 64 bf16] per CTA), warp 1 lane 0 of the leader MMA issuer (M = 256, N = 64, K = 16), warp 2 TMEM
// allocator, warps 4-7 epilogue: first the query rows global -> registers -> TMEM, then per item tcgen05.ld of the
// 64 accumulator columns, buffer handed back at once, threshold filter + sparse cooperative candidate extraction
// exactly as in flat_scan.cu.
#include "flat_scan_common.cuh"

namespace tt {

constexpr int TS_N = 64;                       // catalog rows per tile
constexpr int TS_HALF = TS_N / 2;              // rows of a tile one CTA loads
constexpr int TS_KB_BYTES = TS_HALF * 128;     // one K-block of this CTA's half tile: 4 KB
constexpr int TS_MAX_SLOTS = 8;
constexpr int TS_QCOL0 = 128;                  // TMEM: accumulator slots at columns 0 and 64, query blocks from 128

struct ScanTsParams {
  long long N;
  int nq, num_kb, nqu, nslices, num_tiles, num_slots, slot_bytes, qcols, seg_cap, Dp;
  const float* thr;
  unsigned int* seg_cnt;
  uint2* cand;
  const uint4* qh;       // bf16 [nq_pad, Dp]
};

// The MAIN epilogue of flat_scan.cu for one 64-column item: v0 / v1 are accumulator columns 0-31 / 32-63 of this
// thread's query; hits (rare) are extracted cooperatively through a 256-byte per-warp scratch row.
__device__ __forceinline__ void ts_consume(const uint32_t (&v0)[32], const uint32_t (&v1)[32], float thr, int q, unsigned int& my_cnt,
                                           uint32_t row0, uint32_t* scratch, const ScanTsParams& p, int slice, int lane) {
  const float m0 = max_tree(v0);
  const float m1 = max_tree(v1);
  unsigned int hitmask = __ballot_sync(0xffffffffu, fmaxf(m0, m1) >= thr);
  while (hitmask) {
    const int src = __ffs(hitmask) - 1;
    hitmask &= hitmask - 1;
    const unsigned int halves = __shfl_sync(0xffffffffu, (m0 >= thr ? 1u : 0u) | (m1 >= thr ? 2u : 0u), src);
    if (lane == src) {
      if (halves & 1u) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(scratch + j) = make_uint4(v0[j], v0[j + 1], v0[j + 2], v0[j + 3]);
      }
      if (halves & 2u) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(scratch + 32 + j) = make_uint4(v1[j], v1[j + 1], v1[j + 2], v1[j + 3]);
      }
    }
    __syncwarp();
    const float thr_s = __shfl_sync(0xffffffffu, thr, src);
    const unsigned int cnt_s = __shfl_sync(0xffffffffu, my_cnt, src);
    const int q_s = __shfl_sync(0xffffffffu, q, src);
    const uint32_t x0 = (halves & 1u) ? scratch[lane] : 0xff800000u;
    const uint32_t x1 = (halves & 2u) ? scratch[32 + lane] : 0xff800000u;
    const bool h0 = __uint_as_float(x0) >= thr_s, h1 = __uint_as_float(x1) >= thr_s;
    const unsigned int b0 = __ballot_sync(0xffffffffu, h0), b1 = __ballot_sync(0xffffffffu, h1);
    const unsigned int lt = (1u << lane) - 1u;
    const unsigned int p0 = cnt_s + __popc(b0 & lt), p1 = cnt_s + __popc(b0) + __popc(b1 & lt);
    uint2* seg = p.cand + ((size_t)q_s * p.nslices + slice) * p.seg_cap;
    const uint32_t rbase = row0 + (uint32_t)lane;
    if (h0 && p0 < (unsigned int)p.seg_cap) seg[p0] = make_uint2(x0, rbase);
    if (h1 && p1 < (unsigned int)p.seg_cap) seg[p1] = make_uint2(x1, rbase + 32u);
    if (lane == src) my_cnt += __popc(b0) + __popc(b1);
    __syncwarp();
  }
}

template <int NQB>
__global__ void __launch_bounds__(SCAN_THREADS, 1)
flat_scan_ts_kernel(const __grid_constant__ CUtensorMap tmap_x, const ScanTsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)p.num_slots * p.slot_bytes);
  uint64_t* full_bar = bars;                          // [TS_MAX_SLOTS]  whole tile landed (both CTAs)
  uint64_t* empty_bar = bars + TS_MAX_SLOTS;          // [TS_MAX_SLOTS]  MMAs of every block over the tile retired
  uint64_t* q_bar = bars + 2 * TS_MAX_SLOTS;          // [1]  query blocks are in TMEM (4 warps x 2 CTAs)
  uint64_t* tmem_full_bar = q_bar + 1;                // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;       // [2]  one arrival per epilogue warp of both CTAs
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint32_t* scratch_base = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + BAR_BYTES);   // 4 x 256 B

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  pdl_trigger();

  const int unit = (int)(blockIdx.x >> 1);
  const int qu = unit % p.nqu;
  const int slice = unit / p.nqu;
  const int ntiles = (p.num_tiles > slice) ? (p.num_tiles - slice + p.nslices - 1) / p.nslices : 0;

  if (warp == 0 && lane == 0) prefetch_tensormap(&tmap_x);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.num_slots; ++i) {
      mbar_init(smem_u32(full_bar + i), 2);            // leader's arrive.expect_tx + the partner's remote arrive
      mbar_init(smem_u32(empty_bar + i), 1);
    }
    mbar_init(smem_u32(q_bar), 8);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(tmem_full_bar + i), 1);
      mbar_init(smem_u32(tmem_empty_bar + i), 8);
    }
    fence_barrier_init();
  }
  cluster_sync();
  if (warp == 2) { tmem_alloc_2cta(smem_u32(tmem_ptr_smem), TMEM_COLS); tmem_relinquish_2cta(); }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0) {
    // =========================== TMA producer ===============================================
    if (lane == 0) {
      int slot = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const long long tile = (long long)slice + (long long)t * p.nslices;
        const int row0 = (int)(tile * TS_N) + (int)cta_rank * TS_HALF;
        mbar_wait(smem_u32(empty_bar + slot), phase ^ 1, 600 + slot);
        const uint32_t fb = smem_u32(full_bar + slot);
        const uint32_t dst = smem_u32(ring + (size_t)slot * p.slot_bytes);
        if (leader) mbar_arrive_expect_tx(fb, (uint32_t)p.slot_bytes * 2u);
        for (int kb = 0; kb < p.num_kb; ++kb) tma_load_2d_2cta(dst + (uint32_t)(kb * TS_KB_BYTES), &tmap_x, fb, kb * BLOCK_K, row0);
        if (!leader) mbar_arrive_remote(fb, 0);
        if (++slot == p.num_slots) { slot = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA) ======================================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc_bf16_f32(256, TS_N);
      mbar_wait(smem_u32(q_bar), 0, 610);
      tc_fence_after();
      int slot = 0;
      uint32_t phase = 0;
      uint32_t j = 0;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(smem_u32(full_bar + slot), phase, 620 + slot);
        tc_fence_after();
        const uint64_t b_desc0 = make_smem_desc_sw128(smem_u32(ring + (size_t)slot * p.slot_bytes));
#pragma unroll
        for (int b = 0; b < NQB; ++b, ++j) {
          const uint32_t acc = j & 1u;
          mbar_wait(smem_u32(tmem_empty_bar + acc), ((j >> 1) & 1u) ^ 1u, 630 + (int)acc);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * (uint32_t)TS_N;
          const uint32_t a0 = tmem_base + (uint32_t)TS_QCOL0 + (uint32_t)(b * p.qcols);
          for (int kb = 0; kb < p.num_kb; ++kb) {
            const uint64_t bd = b_desc0 + (uint64_t)((kb * TS_KB_BYTES) >> 4);
            const uint32_t ad = a0 + (uint32_t)(kb * 32);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              mma_bf16_ts_2cta(d_tmem, ad + (uint32_t)(k * 8), bd + (uint64_t)((k * UMMA_K * 2) >> 4), idesc, (uint32_t)((kb | k) != 0));
          }
          mma_commit_2cta(smem_u32(tmem_full_bar + acc));
        }
        mma_commit_2cta(smem_u32(empty_bar + slot));      // every block has read this tile
        if (++slot == p.num_slots) { slot = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // =========================== epilogue ====================================================
    const int quad = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
    // queries of this CTA -> tensor memory (A operand): lane = query row, 32 columns (64 bf16) per K-block
    int q[NQB];
    float thr[NQB];
    unsigned int cnt[NQB];
#pragma unroll
    for (int b = 0; b < NQB; ++b) {
      q[b] = (qu * NQB + b) * 256 + (int)cta_rank * 128 + quad * 32 + lane;
      const uint4* src = p.qh + (size_t)q[b] * (size_t)(p.Dp >> 3);
      for (int kb = 0; kb < p.num_kb; ++kb) {
        uint32_t r[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = __ldg(src + kb * 8 + c);
          r[4 * c] = v.x; r[4 * c + 1] = v.y; r[4 * c + 2] = v.z; r[4 * c + 3] = v.w;
        }
        __syncwarp();
        tmem_st_32x32(lane_base + (uint32_t)TS_QCOL0 + (uint32_t)(b * p.qcols + kb * 32), r);
      }
      thr[b] = (q[b] < p.nq) ? __ldg(p.thr + q[b]) : INFINITY;
      cnt[b] = 0u;
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (leader) mbar_arrive(smem_u32(q_bar)); else mbar_arrive_remote(smem_u32(q_bar), 0);
    }
    uint32_t* scratch = scratch_base + (warp - 4) * 64;
    uint32_t j = 0;
    for (int t = 0; t < ntiles; ++t) {
      const long long tile = (long long)slice + (long long)t * p.nslices;
      const long long row0 = tile * TS_N;
      const int ncols = (int)min((long long)TS_N, p.N - row0);
#pragma unroll
      for (int b = 0; b < NQB; ++b, ++j) {
        const uint32_t acc = j & 1u;
        mbar_wait(smem_u32(tmem_full_bar + acc), (j >> 1) & 1u, 640 + (int)acc);
        tc_fence_after();
        const uint32_t taddr = lane_base + acc * (uint32_t)TS_N;
        const bool valid = q[b] < p.nq;
        if (ncols == TS_N) {
          uint32_t v0[32], v1[32];
          __syncwarp();
          tmem_ld_32x32(taddr, v0);
          tmem_ld_32x32(taddr + 32u, v1);
          tmem_ld_wait();
          tc_fence_before();                  // the values are in registers: hand the accumulator slot back at once
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(smem_u32(tmem_empty_bar + acc)); else mbar_arrive_remote(smem_u32(tmem_empty_bar + acc), 0);
          }
          ts_consume(v0, v1, thr[b], valid ? q[b] : 0, cnt[b], (uint32_t)row0, scratch, p, slice, lane);
        } else {
          uint2* seg = p.cand + ((size_t)(valid ? q[b] : 0) * p.nslices + slice) * p.seg_cap;
          append_columns(taddr, ncols, thr[b], (uint32_t)row0, seg, cnt[b], (unsigned int)p.seg_cap);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(smem_u32(tmem_empty_bar + acc)); else mbar_arrive_remote(smem_u32(tmem_empty_bar + acc), 0);
          }
        }
      }
    }
#pragma unroll
    for (int b = 0; b < NQB; ++b)
      if (q[b] < p.nq) p.seg_cnt[(size_t)q[b] * p.nslices + slice] = cnt[b];
  }

  tc_fence_before();
  cluster_sync();
  if (warp == 2) { tc_fence_after(); tmem_dealloc_2cta(tmem_base, TMEM_COLS); }
}

// bf16 [rows, pitch] row-major, box = [box_rows, 64 columns], 128-byte swizzle (flat_scan.cu)
int make_tmap_bf16(void* tensor_map, const void* base, long long rows, int pitch, int box_rows);

template <int NQB>
static int launch_ts(const CUtensorMap& tx, const ScanTsParams& sp, int units, size_t smem, cudaStream_t st) {
  auto kern = flat_scan_ts_kernel<NQB>;
  TT_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * units));
  cfg.blockDim = dim3(SCAN_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  count_launch();
  TT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tx, sp));
  return TT_OK;
}

int launch_main_scan_ts(const ScanPlan& pl, const void* qh, const void* Xh, long long N, int nq, float* thr,
                        unsigned int* seg_cnt, void* cand, cudaStream_t st) {
  CUtensorMap tx;
  if (int e = make_tmap_bf16(&tx, Xh, N, pl.Dp, TS_HALF)) return e;
  ScanTsParams sp{};
  sp.N = N; sp.nq = nq; sp.num_kb = pl.num_kb; sp.nqu = pl.ts_nqu; sp.nslices = pl.main_slices; sp.num_tiles = pl.ts_tiles;
  sp.num_slots = pl.ts_slots; sp.slot_bytes = pl.num_kb * TS_KB_BYTES; sp.qcols = pl.Dp / 2; sp.seg_cap = pl.seg_cap; sp.Dp = pl.Dp;
  sp.thr = thr; sp.seg_cnt = seg_cnt; sp.cand = reinterpret_cast<uint2*>(cand); sp.qh = reinterpret_cast<const uint4*>(qh);
  const size_t smem = (size_t)sp.num_slots * sp.slot_bytes + 1024 + BAR_BYTES + SCRATCH_BYTES;
  profile_scan_begin(st);
  const int e = (pl.ts_nqb == 2) ? launch_ts<2>(tx, sp, pl.main_slices * pl.ts_nqu, smem, st)
                                 : launch_ts<1>(tx, sp, pl.main_slices * pl.ts_nqu, smem, st);
  profile_scan_end(st);
  return e;
}

}  // namespace tt
