// InfoNCE loss of the training callers (reference: src/training/losses.py:36-79, used by src/training/trainer.py:216-236):
//     logits_i = [ b_i.p_i | b_i.n_i1 .. b_i.n_iM | b_i.p_k (k != i) ] / T ,   loss = mean_i ( logsumexp(logits_i) - logits_i0 )
// Forward and backward WITHOUT the [B, B, D] expansion and the [B, 1+M+B] logits matrix the reference materialises:
// one CTA per buyer row streams the candidate rows (its positive, its M sampled negatives, the other rows' positives),
// one warp per candidate, and keeps an online logsumexp; the backward recomputes the dot products instead of storing
// them.  fp32 CUDA cores on purpose: B = 512, D = 384 is 0.2 GFLOP - the op is latency- and L2-bound, not tensor-bound.
//
//   forward : row_loss[i], lse[i]                                  (infonce_fwd_kernel)   + deterministic mean
//   backward: w_ij = (g / B) (softmax_ij - [j = 0]) / T            = dL/d(b_i . e_ij)
//             d_buyer_i = sum_j w_ij e_ij ;  d_neg_ij = w_ij b_i   (infonce_bwd_rows_kernel, also stores w for in-batch pairs)
//             d_pos_k   = w_k0 b_k + sum_{i != k} w_i,(k) b_i      (infonce_bwd_pos_kernel)
#include <math.h>
#include "tt_common.cuh"

namespace tt {

constexpr int NCE_THREADS = 256;
constexpr int NCE_WARPS = NCE_THREADS / 32;
constexpr int NCE_JMAX = 32;                 // D <= 1024: a lane holds D/32 <= 32 elements of a row

// candidate c of row i: 0 -> p_i ; 1..M -> n[i][c-1] ; M+1+k -> p_k (k != i, skipped for k == i)
__device__ __forceinline__ const float* nce_candidate(const float* __restrict__ p, const float* __restrict__ n, int i, int c,
                                                      int M, int D) {
  if (c == 0) return p + (size_t)i * D;
  if (c <= M) return n + ((size_t)i * M + (size_t)(c - 1)) * D;
  const int k = c - M - 1;
  return (k == i) ? nullptr : p + (size_t)k * D;
}

__global__ void __launch_bounds__(NCE_THREADS)
infonce_fwd_kernel(const float* __restrict__ b, const float* __restrict__ p, const float* __restrict__ n, int B, int M, int D,
                   float inv_t, float* __restrict__ row_loss, float* __restrict__ lse_out) {
  extern __shared__ float sb[];          // [D] this buyer's row
  __shared__ float wm[NCE_WARPS], ws[NCE_WARPS];
  __shared__ float s_pos;
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int d = tid; d < D; d += NCE_THREADS) sb[d] = b[(size_t)i * D + d];
  __syncthreads();
  float m = -INFINITY, s = 0.f;
  const int ncand = 1 + M + B;
  for (int c = warp; c < ncand; c += NCE_WARPS) {
    const float* e = nce_candidate(p, n, i, c, M, D);
    if (e == nullptr) continue;          // the masked diagonal of the in-batch block (losses.py:63-64)
    float dot = 0.f;
    for (int d = lane; d < D; d += 32) dot = fmaf(sb[d], __ldg(e + d), dot);
    dot = warp_sum(dot) * inv_t;
    if (c == 0 && lane == 0) s_pos = dot;
    if (dot > m) { s = s * expf(m - dot) + 1.f; m = dot; }      // exp(-inf) = 0 on the first candidate
    else s += expf(dot - m);
  }
  if (lane == 0) { wm[warp] = m; ws[warp] = s; }
  __syncthreads();
  if (tid == 0) {
    float mm = -INFINITY;
    for (int w = 0; w < NCE_WARPS; ++w) mm = fmaxf(mm, wm[w]);
    float ss = 0.f;
    for (int w = 0; w < NCE_WARPS; ++w) if (wm[w] > -INFINITY) ss += ws[w] * expf(wm[w] - mm);
    const float lse = mm + logf(ss);
    lse_out[i] = lse;
    row_loss[i] = lse - s_pos;           // F.cross_entropy with label 0 (losses.py:73-77)
  }
}

// loss = mean of row_loss, summed in a fixed order (deterministic)
__global__ void __launch_bounds__(NCE_THREADS)
infonce_mean_kernel(const float* __restrict__ row_loss, int B, float* __restrict__ loss) {
  __shared__ float red[NCE_THREADS];
  float a = 0.f;
  for (int i = threadIdx.x; i < B; i += NCE_THREADS) a += row_loss[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = NCE_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = red[0] / (float)B;
}

__global__ void __launch_bounds__(NCE_THREADS)
infonce_bwd_rows_kernel(const float* __restrict__ b, const float* __restrict__ p, const float* __restrict__ n,
                        const float* __restrict__ lse, const float* __restrict__ grad_loss, int B, int M, int D, float inv_t,
                        float* __restrict__ d_buyer, float* __restrict__ d_neg, float* __restrict__ w_inb_t,
                        float* __restrict__ w_pos) {
  extern __shared__ float sm[];          // [D] this buyer's row, then [NCE_WARPS][D] partial d_buyer
  float* sb = sm;
  float* part = sm + D;
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int d = tid; d < D; d += NCE_THREADS) sb[d] = b[(size_t)i * D + d];
  __syncthreads();
  const float scale = __ldg(grad_loss) / (float)B * inv_t;
  const float my_lse = __ldg(lse + i);
  float acc[NCE_JMAX];
#pragma unroll
  for (int j = 0; j < NCE_JMAX; ++j) acc[j] = 0.f;
  const int ncand = 1 + M + B;
  for (int c = warp; c < ncand; c += NCE_WARPS) {
    const float* e = nce_candidate(p, n, i, c, M, D);
    if (e == nullptr) continue;
    float ev[NCE_JMAX];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < NCE_JMAX; ++j) {
      const int d = lane + 32 * j;
      ev[j] = (d < D) ? __ldg(e + d) : 0.f;
      dot = fmaf((d < D) ? sb[d] : 0.f, ev[j], dot);
    }
    dot = warp_sum(dot) * inv_t;
    float w = expf(dot - my_lse);        // softmax probability of this candidate
    if (c == 0) w -= 1.f;                // label 0
    w *= scale;                          // dL/d(b_i . e)
#pragma unroll
    for (int j = 0; j < NCE_JMAX; ++j) acc[j] = fmaf(w, ev[j], acc[j]);
    if (c == 0) {
      if (lane == 0) w_pos[i] = w;
    } else if (c <= M) {
      float* dn = d_neg + ((size_t)i * M + (size_t)(c - 1)) * D;
#pragma unroll
      for (int j = 0; j < NCE_JMAX; ++j) { const int d = lane + 32 * j; if (d < D) dn[d] = w * sb[d]; }
    } else if (lane == 0) {
      w_inb_t[(size_t)(c - M - 1) * B + i] = w;      // transposed: row k holds the weights of every buyer i on p_k
    }
  }
#pragma unroll
  for (int j = 0; j < NCE_JMAX; ++j) { const int d = lane + 32 * j; if (d < D) part[(size_t)warp * D + d] = acc[j]; }
  __syncthreads();
  for (int d = tid; d < D; d += NCE_THREADS) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < NCE_WARPS; ++w) a += part[(size_t)w * D + d];
    d_buyer[(size_t)i * D + d] = a;
  }
}

__global__ void __launch_bounds__(NCE_THREADS)
infonce_bwd_pos_kernel(const float* __restrict__ b, const float* __restrict__ w_inb_t, const float* __restrict__ w_pos, int B,
                       int D, float* __restrict__ d_pos) {
  extern __shared__ float sw[];          // [B] weights of every buyer on this positive row
  const int k = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < B; i += NCE_THREADS) sw[i] = (i == k) ? __ldg(w_pos + k) : __ldg(w_inb_t + (size_t)k * B + i);
  __syncthreads();
  for (int d = tid; d < D; d += NCE_THREADS) {
    float a = 0.f;
    for (int i = 0; i < B; ++i) a = fmaf(sw[i], __ldg(b + (size_t)i * D + d), a);
    d_pos[(size_t)k * D + d] = a;
  }
}

}  // namespace tt

using namespace tt;

extern "C" __attribute__((visibility("default"))) size_t tt_infonce_workspace_bytes(int B) {
  if (B < 1) return 0;
  return align_up(((size_t)B * B + B) * sizeof(float), 256);
}

extern "C" __attribute__((visibility("default"))) int tt_infonce_forward(const float* buyer, const float* pos, const float* neg, int B, int M,
                                                            int D, float temperature, float* loss, float* row_loss,
                                                            float* lse, void* stream) {
  TT_CHECK_ARG(buyer && pos && loss && row_loss && lse, "null pointer");
  TT_CHECK_ARG(B >= 1 && M >= 0 && D >= 1 && D <= 32 * NCE_JMAX, "need B >= 1, M >= 0, 1 <= D <= 1024");
  TT_CHECK_ARG(M == 0 || neg != nullptr, "negative_embeddings is NULL with M > 0");
  TT_CHECK_ARG(temperature > 0.f, "temperature must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  infonce_fwd_kernel<<<B, NCE_THREADS, (size_t)D * sizeof(float), st>>>(buyer, pos, neg, B, M, D, 1.0f / temperature, row_loss, lse);
  TT_CHECK_LAUNCH();
  infonce_mean_kernel<<<1, NCE_THREADS, 0, st>>>(row_loss, B, loss);
  TT_CHECK_LAUNCH();
  return TT_OK;
}

extern "C" __attribute__((visibility("default"))) int tt_infonce_backward(const float* buyer, const float* pos, const float* neg,
                                                             const float* lse, const float* grad_loss, int B, int M, int D,
                                                             float temperature, float* d_buyer, float* d_pos, float* d_neg,
                                                             void* workspace, size_t workspace_bytes, void* stream) {
  TT_CHECK_ARG(buyer && pos && lse && grad_loss && d_buyer && d_pos && workspace, "null pointer");
  TT_CHECK_ARG(B >= 1 && M >= 0 && D >= 1 && D <= 32 * NCE_JMAX, "need B >= 1, M >= 0, 1 <= D <= 1024");
  TT_CHECK_ARG(M == 0 || (neg != nullptr && d_neg != nullptr), "negative_embeddings / d_neg is NULL with M > 0");
  TT_CHECK_ARG(temperature > 0.f, "temperature must be positive");
  TT_CHECK_ARG(B <= 12288, "B too large for the positive-gradient kernel's shared memory");
  if (workspace_bytes < tt_infonce_workspace_bytes(B)) {
    set_error("tt_infonce_backward: workspace too small");
    return TT_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* w_inb_t = reinterpret_cast<float*>(workspace);
  float* w_pos = w_inb_t + (size_t)B * B;
  const size_t sm1 = (size_t)(1 + NCE_WARPS) * D * sizeof(float);
  if (sm1 > 48 * 1024) TT_CHECK_CUDA(cudaFuncSetAttribute(infonce_bwd_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
  infonce_bwd_rows_kernel<<<B, NCE_THREADS, sm1, st>>>(buyer, pos, neg, lse, grad_loss, B, M, D, 1.0f / temperature, d_buyer, d_neg,
                                                       w_inb_t, w_pos);
  TT_CHECK_LAUNCH();
  const size_t sm2 = (size_t)B * sizeof(float);
  if (sm2 > 48 * 1024) TT_CHECK_CUDA(cudaFuncSetAttribute(infonce_bwd_pos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  infonce_bwd_pos_kernel<<<B, NCE_THREADS, sm2, st>>>(buyer, w_inb_t, w_pos, B, D, d_pos);
  TT_CHECK_LAUNCH();
  return TT_OK;
}
