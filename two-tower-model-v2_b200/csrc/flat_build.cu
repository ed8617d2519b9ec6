// Index build and query preparation for the exact inner-product index
// (reference: src/inference/vector_db.py:44-54 build_index, :152-156 / :189-193 query renormalise).
//
//   Xn[r,:] = X[r,:] / (||X[r,:]||_2 + 1e-8)              fp32, what IndexFlatIP would store
//   Xh[r,:] = bf16_rn(Xn[r,:]), zero padded to pitch Dp     the copy the tcgen05 scan streams
//   stats   = { max_r ||Xh[r]||, max_r ||Xh[r] - Xn[r]|| }  bounds the bf16 scoring error
//
// One warp per row; the row is read once (kept in registers for D <= 1024), HBM-bound.
#include "tt_common.cuh"

namespace tt {

__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
  // non-negative floats order like their bit patterns
  atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// Normalises one row (warp-cooperative).  `src` f32[D]; writes dst_n f32[D] (may alias src) and
// dst_h bf16[Dp].  Returns (||h||^2, ||h - n||^2, ||n||^2) reduced over the warp.
template <int MAXV>   // row cached in registers when D <= MAXV*128 floats and D % 4 == 0
__device__ __forceinline__ void normalise_row(const float* __restrict__ src, int D, int Dp,
                                              float* dst_n, __nv_bfloat16* dst_h, int lane,
                                              float& hh, float& dd, float& nn, bool normalize = true) {
  float ss = 0.f;
  const bool vec = (D % 4 == 0) && (D <= MAXV * 128) &&
                   ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(dst_n) & 15) == 0);
  hh = dd = nn = 0.f;
  if (vec) {
    float4 v[MAXV];
    const int n4 = D >> 2;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = i * 32 + lane;
      v[i] = (c < n4) ? reinterpret_cast<const float4*>(src)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
    }
    ss = warp_sum(ss);
    const float den = normalize ? (sqrtf(ss) + 1e-8f) : 1.0f;   // vector_db.py:44-45
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = i * 32 + lane;
      if (c < n4) {
        float4 n = make_float4(v[i].x / den, v[i].y / den, v[i].z / den, v[i].w / den);
        reinterpret_cast<float4*>(dst_n)[c] = n;
        __nv_bfloat16 h0 = __float2bfloat16_rn(n.x), h1 = __float2bfloat16_rn(n.y);
        __nv_bfloat16 h2 = __float2bfloat16_rn(n.z), h3 = __float2bfloat16_rn(n.w);
        const float f0 = __bfloat162float(h0), f1 = __bfloat162float(h1);
        const float f2 = __bfloat162float(h2), f3 = __bfloat162float(h3);
        hh += f0 * f0 + f1 * f1 + f2 * f2 + f3 * f3;
        dd += (f0 - n.x) * (f0 - n.x) + (f1 - n.y) * (f1 - n.y) + (f2 - n.z) * (f2 - n.z) + (f3 - n.w) * (f3 - n.w);
        nn += n.x * n.x + n.y * n.y + n.z * n.z + n.w * n.w;
        __nv_bfloat162 p0 = __halves2bfloat162(h0, h1), p1 = __halves2bfloat162(h2, h3);
        uint2 pk;
        pk.x = *reinterpret_cast<uint32_t*>(&p0);
        pk.y = *reinterpret_cast<uint32_t*>(&p1);
        reinterpret_cast<uint2*>(dst_h)[c] = pk;
      }
    }
  } else {
    for (int d = lane; d < D; d += 32) { const float a = src[d]; ss += a * a; }
    ss = warp_sum(ss);
    const float den = normalize ? (sqrtf(ss) + 1e-8f) : 1.0f;
    for (int d = lane; d < D; d += 32) {
      const float n = src[d] / den;
      dst_n[d] = n;
      const __nv_bfloat16 h = __float2bfloat16_rn(n);
      const float f = __bfloat162float(h);
      hh += f * f; dd += (f - n) * (f - n); nn += n * n;
      dst_h[d] = h;
    }
  }
  // zero the pitch padding of the bf16 row
  for (int d = D + lane; d < Dp; d += 32) dst_h[d] = __float2bfloat16_rn(0.f);
  hh = warp_sum(hh); dd = warp_sum(dd); nn = warp_sum(nn);
}

__global__ void __launch_bounds__(256)
flat_build_kernel(const float* __restrict__ X, long long rows, int D, int Dp,
                  float* Xn, __nv_bfloat16* Xh, long long row0, float* stats, bool normalize) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  float mh = 0.f, md = 0.f;
  for (long long r = warp; r < rows; r += nwarps) {
    float hh, dd, nn;
    normalise_row<8>(X + r * D, D, Dp, Xn + (row0 + r) * D, Xh + (row0 + r) * Dp, lane, hh, dd, nn, normalize);
    mh = fmaxf(mh, hh); md = fmaxf(md, dd);
  }
  if (lane == 0) {
    atomic_max_nonneg(stats + 0, sqrtf(mh));
    atomic_max_nonneg(stats + 1, sqrtf(md));
  }
}

// Query preparation: qn = q/(||q||+1e-8) (fp32, used for rescoring), qh = bf16(qn) padded to
// [nq_pad, Dp] (rows >= nq are zero), eps[q] = bound on |bf16 tensor-core score - fp32 score|
// over all catalog rows:
//   |<qh,xh> - <qn,xn>| <= ||qh-qn|| * ||xh|| + ||qn|| * ||xh-xn||            (Cauchy-Schwarz)
// plus slack for the tensor core's fp32 accumulation (D * 2^-22 * ||qh|| * ||xh||).
__global__ void __launch_bounds__(256)
flat_prep_queries_kernel(const float* __restrict__ q, int nq, int nq_pad, int D, int Dp,
                         const float* __restrict__ stats,
                         float* qn, __nv_bfloat16* qh, float* eps, int* zero_me) {
  pdl_trigger();
  pdl_wait();      // the previous search on this stream may still be reading this workspace
  if (zero_me && blockIdx.x == 0 && threadIdx.x == 0) *zero_me = 0;     // n_uncertified of this call (was a memset)
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= nq_pad) return;
  if (r >= nq) {
    for (int d = lane; d < Dp; d += 32) qh[(long long)r * Dp + d] = __float2bfloat16_rn(0.f);
    return;
  }
  float hh, dd, nn;
  normalise_row<8>(q + (long long)r * D, D, Dp, qn + (long long)r * D, qh + (long long)r * Dp, lane, hh, dd, nn);
  if (lane == 0) {
    const float max_xh = stats[0], max_dx = stats[1];
    const float qh_norm = sqrtf(hh), dq = sqrtf(dd), qn_norm = sqrtf(nn);
    float e = dq * max_xh + qn_norm * max_dx + (float)D * 2.384185791015625e-07f * qh_norm * max_xh;
    e = e * 1.001f + 1e-7f;   // rounding of this very computation
    eps[r] = e;
  }
}

}  // namespace tt

using namespace tt;

extern "C" __attribute__((visibility("default"))) int64_t tt_flat_pitch(int D) { return ((int64_t)D + 63) / 64 * 64; }

extern "C" __attribute__((visibility("default"))) int tt_flat_build(const float* X, int64_t rows, int D, int normalize, float* Xn, void* Xh, int64_t row0,
                             float* stats, void* stream) {
  TT_CHECK_ARG(X && Xn && Xh && stats, "null pointer");
  TT_CHECK_ARG(rows >= 0 && D >= 1 && row0 >= 0, "need rows >= 0, D >= 1, row0 >= 0");
  if (rows == 0) return TT_OK;
  const int Dp = (int)tt_flat_pitch(D);
  const long long warps_needed = rows;
  long long grid = (warps_needed + 7) / 8;
  const long long cap = (long long)num_sms() * 32;
  if (grid > cap) grid = cap;
  flat_build_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(X, rows, D, Dp, Xn,
                                                                     reinterpret_cast<__nv_bfloat16*>(Xh), row0, stats, normalize != 0);
  TT_CHECK_LAUNCH();
  return TT_OK;
}

namespace tt {
int launch_prep_queries(const float* q, int nq, int nq_pad, int D, int Dp, const float* stats,
                        float* qn, void* qh, float* eps, int* zero_me, cudaStream_t st) {
  const int grid = (nq_pad + 7) / 8;
  count_launch();
  TT_CHECK_CUDA(launch_pdl(flat_prep_queries_kernel, dim3(grid), dim3(256), 0, st, q, nq, nq_pad, D, Dp, stats, qn,
                           reinterpret_cast<__nv_bfloat16*>(qh), eps, zero_me));
  return TT_OK;
}
}  // namespace tt
