// Buyer-tower pooling kernels (reference: src/models/buyer_tower.py:43-101).
//
// One fused kernel does, per buyer: per-event coefficients (normalised event weights, or the
// softmax of logit*weight), the coefficient-weighted sum of S item-embedding rows, and the
// final L2 normalisation.  Rows are read exactly once with coalesced 128-bit loads (a warp
// covers 512 contiguous bytes per load instruction); reductions are warp shuffles.  The
// kernel is HBM-bound: algorithmic bytes = B*S*D*4 + B*S*4 (+ B*S*8 idx, + B*S*4 logits) + B*D*4.
#include "tt_common.cuh"

namespace tt {

struct PoolParams {
  const float* x;        // dense [B,S,D] or table [N,D]
  const int64_t* idx;    // gather only: [B,S]
  const float* w;        // [B,S]
  const float* logits;   // attention: dense [B,S]; gather: per-table-row [N]
  float* out;            // [B,D]
  long long N;           // table rows (gather)
  float zero_row_logit;  // attention+gather: logit of an all-zero row
  int B, S, D;
};

template <bool GATHER, bool ATTN>
__device__ __forceinline__ float load_c(const PoolParams& p, int b, int s, long long& row) {
  // returns w (weighted) or logit*w (attention); sets row index (gather) / -1 invalid
  const long long bs = (long long)b * p.S + s;
  const float w = __ldg(p.w + bs);
  if (GATHER) {
    const long long r = __ldg((const long long*)p.idx + bs);
    row = (r >= 0 && r < p.N) ? r : -1;
  } else {
    row = bs;
  }
  if (ATTN) {
    float lg;
    if (GATHER) lg = (row >= 0) ? __ldg(p.logits + row) : p.zero_row_logit;
    else lg = __ldg(p.logits + bs);
    return lg * w;   // buyer_tower.py:89  combined_scores = attention_scores * weights
  }
  return w;
}

// NV   : float4 per lane per row (ceil(D/128))
// WPB  : warps cooperating on one buyer (1 or WARPS)
// U    : rows in flight per warp
template <int NV, int WARPS, int WPB, int U, bool GATHER, bool ATTN>
__global__ void __launch_bounds__(WARPS * 32, (WPB == 1) ? (NV <= 3 ? 7 : (NV <= 4 ? 5 : 3)) : 1)
pool_vec_kernel(const PoolParams p) {
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int buyers_per_cta = WARPS / WPB;
  const int b = blockIdx.x * buyers_per_cta + (WPB == 1 ? warp : 0);
  const int sub = (WPB == 1) ? 0 : warp;   // which slice of rows this warp takes
  const bool active = b < p.B;             // warp-uniform

  __shared__ float4 red[(WPB > 1) ? WARPS * NV * 32 : 1];

  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

  if (active) {
    const int S = p.S, D = p.D;
    // ---- pass 1: normaliser of the coefficients (tiny; w/logits stay in L1) -------------
    float m = -INFINITY, tot = 0.f;
    if (ATTN) {
      for (int s = lane; s < S; s += 32) { long long r; m = fmaxf(m, load_c<GATHER, ATTN>(p, b, s, r)); }
      m = warp_max(m);
      for (int s = lane; s < S; s += 32) { long long r; tot += expf(load_c<GATHER, ATTN>(p, b, s, r) - m); }
      tot = warp_sum(tot);
    } else {
      for (int s = lane; s < S; s += 32) tot += __ldg(p.w + (long long)b * S + s);
      tot = warp_sum(tot) + 1e-8f;   // buyer_tower.py:59
    }

    // ---- pass 2: weighted row sum ---------------------------------------------------------
    const int nvalid4 = D >> 2;   // float4 per row
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      float coef = 0.f;
      long long row = -1;
      if (s < S) {
        const float c = load_c<GATHER, ATTN>(p, b, s, row);
        coef = ATTN ? (expf(c - m) / tot)   // softmax, buyer_tower.py:92
                    : (c / tot);            // weights / weights_sum, buyer_tower.py:60
      }
      const int nrow = min(32, S - s0);
      // this warp handles rows j = sub, sub+WPB, ... of the chunk, U at a time
      for (int j0 = sub; j0 < nrow; j0 += WPB * U) {
        float4 buf[U][NV];
        float cf[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * WPB;
          const int jj = (j < nrow) ? j : 0;
          const float cj = __shfl_sync(0xffffffffu, coef, jj);
          const long long rj = __shfl_sync(0xffffffffu, row, jj);
          const bool ok = (j < nrow) && (rj >= 0);
          cf[u] = ok ? cj : 0.f;
          const float4* rp = reinterpret_cast<const float4*>(p.x + (ok ? rj : 0) * (long long)D);
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const int c4 = v * 32 + lane;
            if (ok && c4 < nvalid4) buf[u][v] = ldg_stream(rp + c4);
            else buf[u][v] = make_float4(0.f, 0.f, 0.f, 0.f);
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
  }

  if (WPB > 1) {
    // cross-warp reduction of the partial sums (small-B path: one buyer per CTA)
#pragma unroll
    for (int v = 0; v < NV; ++v) red[(warp * NV + v) * 32 + lane] = acc[v];
    __syncthreads();
    if (warp != 0) return;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float4 t = red[v * 32 + lane];
      for (int w2 = 1; w2 < WARPS; ++w2) {
        const float4 o = red[(w2 * NV + v) * 32 + lane];
        t.x += o.x; t.y += o.y; t.z += o.z; t.w += o.w;
      }
      acc[v] = t;
    }
  }
  if (!active) return;

  // ---- L2 normalisation: F.normalize(p=2, dim=1, eps=1e-12)  buyer_tower.py:66/:99 ------
  float ss = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v)
    ss += acc[v].x * acc[v].x + acc[v].y * acc[v].y + acc[v].z * acc[v].z + acc[v].w * acc[v].w;
  ss = warp_sum(ss);
  const float denom = fmaxf(sqrtf(ss), 1e-12f);
  float4* op = reinterpret_cast<float4*>(p.out + (long long)b * p.D);
  const int nvalid4 = p.D >> 2;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c4 = v * 32 + lane;
    if (c4 < nvalid4)
      op[c4] = make_float4(acc[v].x / denom, acc[v].y / denom, acc[v].z / denom, acc[v].w / denom);
  }
}

// Generic fallback for shapes the vector kernel does not take (D % 4 != 0, D > 1024 or an
// unaligned base pointer): one CTA per buyer, threads stride over d.  Still a CUDA kernel.
template <bool GATHER, bool ATTN>
__global__ void __launch_bounds__(256)
pool_generic_kernel(const PoolParams p) {
  extern __shared__ float sm[];      // coef[S] + rowoff as 2 floats each -> keep separate arrays
  float* coef = sm;
  long long* rows = reinterpret_cast<long long*>(sm + ((p.S + 1) & ~1));
  __shared__ float red[32];
  __shared__ float bc[2];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = p.S, D = p.D;

  float m = -INFINITY;
  for (int s = tid; s < S; s += blockDim.x) {
    long long r; const float c = load_c<GATHER, ATTN>(p, b, s, r);
    coef[s] = c; rows[s] = r; m = fmaxf(m, c);
  }
  if (ATTN) {
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    if (tid == 0) { float t = red[0]; for (int i = 1; i < (int)(blockDim.x >> 5); ++i) t = fmaxf(t, red[i]); bc[0] = t; }
    __syncthreads();
    m = bc[0];
  }
  __syncthreads();
  float tot = 0.f;
  for (int s = tid; s < S; s += blockDim.x) tot += ATTN ? expf(coef[s] - m) : coef[s];
  tot = warp_sum(tot);
  if (lane == 0) red[warp] = tot;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i]; bc[1] = ATTN ? t : t + 1e-8f; }
  __syncthreads();
  tot = bc[1];
  for (int s = tid; s < S; s += blockDim.x) coef[s] = ATTN ? expf(coef[s] - m) / tot : coef[s] / tot;
  __syncthreads();

  float ss = 0.f;
  for (int d = tid; d < D; d += blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) {
      const long long r = rows[s];
      if (r >= 0) a = fmaf(__ldg(p.x + r * (long long)D + d), coef[s], a);
    }
    p.out[(long long)b * D + d] = a;   // un-normalised; scaled below
    ss += a * a;
  }
  ss = warp_sum(ss);
  __syncthreads();
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i]; bc[0] = fmaxf(sqrtf(t), 1e-12f); }
  __syncthreads();
  const float denom = bc[0];
  for (int d = tid; d < D; d += blockDim.x) p.out[(long long)b * D + d] /= denom;
}

template <int NV, bool GATHER, bool ATTN>
static int launch_vec(const PoolParams& p, cudaStream_t st) {
  // Many buyers: one warp per buyer, 4 warps per CTA, 7 CTAs per SM => 28 resident warps/SM,
  // i.e. 4144 buyers in a single wave on 148 SMs.  Few buyers: 8 warps share one buyer.
  if (p.B >= 2 * num_sms()) {
    constexpr int WARPS = 4;
    const int grid = (p.B + WARPS - 1) / WARPS;
    pool_vec_kernel<NV, WARPS, 1, 2, GATHER, ATTN><<<grid, WARPS * 32, 0, st>>>(p);
  } else {
    constexpr int WARPS = 8;
    pool_vec_kernel<NV, WARPS, WARPS, 4, GATHER, ATTN><<<p.B, WARPS * 32, 0, st>>>(p);
  }
  TT_CHECK_LAUNCH();
  return TT_OK;
}

template <bool GATHER, bool ATTN>
static int launch_pool(const PoolParams& p, cudaStream_t st) {
  const bool vec_ok = (p.D % 4 == 0) && (p.D <= 1024) &&
                      ((reinterpret_cast<uintptr_t>(p.x) & 15) == 0) &&
                      ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
  if (vec_ok) {
    const int nv = (p.D + 127) / 128;
    switch (nv) {
      case 1: return launch_vec<1, GATHER, ATTN>(p, st);
      case 2: return launch_vec<2, GATHER, ATTN>(p, st);
      case 3: return launch_vec<3, GATHER, ATTN>(p, st);
      case 4: return launch_vec<4, GATHER, ATTN>(p, st);
      case 5: case 6: return launch_vec<6, GATHER, ATTN>(p, st);
      default: return launch_vec<8, GATHER, ATTN>(p, st);
    }
  }
  const size_t smem = (size_t)((p.S + 1) & ~1) * sizeof(float) + (size_t)p.S * sizeof(long long);
  if (smem > 200 * 1024) { set_error("pool: S too large for the generic kernel"); return TT_ERR_UNSUPPORTED; }
  if (smem > 48 * 1024)
    TT_CHECK_CUDA(cudaFuncSetAttribute(pool_generic_kernel<GATHER, ATTN>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  pool_generic_kernel<GATHER, ATTN><<<p.B, 256, smem, st>>>(p);
  TT_CHECK_LAUNCH();
  return TT_OK;
}

static int check_common(const void* x, const void* w, const void* out, int B, int S, int D) {
  TT_CHECK_ARG(x && w && out, "null pointer");
  TT_CHECK_ARG(B >= 0 && S >= 1 && D >= 1, "need B >= 0, S >= 1, D >= 1");
  return TT_OK;
}

}  // namespace tt

using namespace tt;

extern "C" __attribute__((visibility("default"))) int tt_pool_weighted(const float* x, const float* w, float* out, int B, int S, int D, void* stream) {
  if (int e = check_common(x, w, out, B, S, D)) return e;
  if (B == 0) return TT_OK;
  PoolParams p{}; p.x = x; p.w = w; p.out = out; p.B = B; p.S = S; p.D = D;
  return launch_pool<false, false>(p, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int tt_pool_weighted_gather(const float* table, int64_t N, const int64_t* idx, const float* w,
                                       float* out, int B, int S, int D, void* stream) {
  if (int e = check_common(table, w, out, B, S, D)) return e;
  TT_CHECK_ARG(idx != nullptr && N >= 1, "null idx or empty table");
  if (B == 0) return TT_OK;
  PoolParams p{}; p.x = table; p.idx = idx; p.w = w; p.out = out; p.N = N; p.B = B; p.S = S; p.D = D;
  return launch_pool<true, false>(p, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int tt_pool_attention(const float* x, const float* logits, const float* w, float* out,
                                 int B, int S, int D, void* stream) {
  if (int e = check_common(x, w, out, B, S, D)) return e;
  TT_CHECK_ARG(logits != nullptr, "null logits");
  if (B == 0) return TT_OK;
  PoolParams p{}; p.x = x; p.w = w; p.logits = logits; p.out = out; p.B = B; p.S = S; p.D = D;
  return launch_pool<false, true>(p, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int tt_pool_attention_gather(const float* table, int64_t N, const float* row_logits,
                                        float zero_row_logit, const int64_t* idx, const float* w,
                                        float* out, int B, int S, int D, void* stream) {
  if (int e = check_common(table, w, out, B, S, D)) return e;
  TT_CHECK_ARG(idx != nullptr && row_logits != nullptr && N >= 1, "null idx/logits or empty table");
  if (B == 0) return TT_OK;
  PoolParams p{}; p.x = table; p.idx = idx; p.w = w; p.logits = row_logits; p.zero_row_logit = zero_row_logit;
  p.out = out; p.N = N; p.B = B; p.S = S; p.D = D;
  return launch_pool<true, true>(p, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// Pooling out of a row-SHARDED item table (BASELINE config C5: the catalog, which is the item table of the
// /retrieve path, is split over the GPUs).  Owner computes: every rank reduces the history rows it owns into a
// partial record {acc[D], m, l, 0, 0} per buyer; ONE all-gather of the [B, D+4] records and a merge kernel finish the
// pooling on every rank.  weighted_avg is linear (acc = sum w_s x_s, l = sum w_s over the positions the rank owns);
// the attention softmax merges like an online softmax (m = max c_s, l = sum e^(c_s-m), acc = sum e^(c_s-m) x_s).
// A position whose index is outside [0, N_total) is an all-zero row (zero-padded history): it adds nothing to acc
// but keeps its weight / softmax mass, exactly as in the reference; the rank with owns_invalid = 1 accounts for it.
namespace tt {

struct PartialParams {
  const float* table;      // this rank's rows [N_local, D]
  const float* row_logits; // attention: logit of every LOCAL row [N_local]; NULL = weighted_avg
  const int64_t* idx;      // [B,S] GLOBAL row ids
  const float* w;          // [B,S]
  float* partial;          // [B, D+4]: acc[D], m, l, 0, 0 (rows stay 16-byte aligned)
  long long N_local, row_lo, N_total;
  float zero_row_logit;
  int owns_invalid;
  int B, S, D;
};

template <int NV, bool ATTN>
__global__ void __launch_bounds__(128)
pool_partial_kernel(const PartialParams p) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= p.B) return;
  const int S = p.S, D = p.D;
  const int nvalid4 = D >> 2;
  auto owned_c = [&](int s, long long& local_row, bool& own) -> float {
    const long long bs = (long long)b * S + s;
    const long long r = __ldg((const long long*)p.idx + bs);
    const float wv = __ldg(p.w + bs);
    const bool valid = (r >= 0 && r < p.N_total);
    const bool mine = valid && r >= p.row_lo && r < p.row_lo + p.N_local;
    own = mine || (!valid && p.owns_invalid);
    local_row = mine ? (r - p.row_lo) : -1;
    if (ATTN) return (mine ? __ldg(p.row_logits + (r - p.row_lo)) : p.zero_row_logit) * wv;
    return wv;
  };
  float m = -INFINITY, l = 0.f;
  if (ATTN) {
    for (int s = lane; s < S; s += 32) { long long lr; bool own; const float c = owned_c(s, lr, own); if (own) m = fmaxf(m, c); }
    m = warp_max(m);
  }
  float4 acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int U = 4;
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    float coef = 0.f;
    long long lrow = -1;
    if (s < S) {
      bool own;
      const float c = owned_c(s, lrow, own);
      if (own) coef = ATTN ? expf(c - m) : c;
      else lrow = -1;
    }
    l += coef;                                   // lane-private; reduced below
    const int nrow = min(32, S - s0);
    for (int j0 = 0; j0 < nrow; j0 += U) {
      float4 buf[U][NV];
      float cf[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u;
        const int jj = (j < nrow) ? j : 0;
        const float cj = __shfl_sync(0xffffffffu, coef, jj);
        const long long rj = __shfl_sync(0xffffffffu, lrow, jj);
        const bool ok = (j < nrow) && (rj >= 0);
        cf[u] = ok ? cj : 0.f;
        const float4* rp = reinterpret_cast<const float4*>(p.table + (ok ? rj : 0) * (long long)D);
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
  l = warp_sum(l);
  float4* op = reinterpret_cast<float4*>(p.partial + (long long)b * (D + 4));
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c4 = v * 32 + lane;
    if (c4 < nvalid4) op[c4] = acc[v];
  }
  if (lane == 0) op[nvalid4] = make_float4(ATTN ? m : 0.f, l, 0.f, 0.f);
}

// partials f32 [G, B, D+4] -> out [B, D], one warp per buyer.
__global__ void __launch_bounds__(128)
pool_partial_merge_kernel(const float* __restrict__ pg, int G, int attention, float* __restrict__ out, int B, int D) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  const long long stride = (long long)B * (D + 4);
  const float* base = pg + (long long)b * (D + 4);
  float M = -INFINITY;
  if (attention)
    for (int g = 0; g < G; ++g) { const float lg = base[g * stride + D + 1]; if (lg > 0.f) M = fmaxf(M, base[g * stride + D]); }
  float L = 0.f;
  for (int g = 0; g < G; ++g) {
    const float lg = base[g * stride + D + 1];
    L += attention ? ((lg > 0.f) ? lg * expf(base[g * stride + D] - M) : 0.f) : lg;
  }
  const float denom = attention ? L : (L + 1e-8f);        // buyer_tower.py:59 / :92
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) {
    float y = 0.f;
    for (int g = 0; g < G; ++g) {
      const float lg = base[g * stride + D + 1];
      const float sc = attention ? ((lg > 0.f) ? expf(base[g * stride + D] - M) : 0.f) : 1.f;
      y = fmaf(base[g * stride + d], sc, y);
    }
    y /= denom;
    out[(long long)b * D + d] = y;
    ss += y * y;
  }
  ss = warp_sum(ss);
  const float nrm = fmaxf(sqrtf(ss), 1e-12f);             // F.normalize eps, buyer_tower.py:66/:99
  __syncwarp();
  for (int d = lane; d < D; d += 32) out[(long long)b * D + d] /= nrm;
}

template <bool ATTN>
static int launch_partial(const PartialParams& p, cudaStream_t st) {
  const int grid = (p.B + 3) / 4;
  const int nv = (p.D + 127) / 128;
  switch (nv) {
    case 1: pool_partial_kernel<1, ATTN><<<grid, 128, 0, st>>>(p); break;
    case 2: pool_partial_kernel<2, ATTN><<<grid, 128, 0, st>>>(p); break;
    case 3: pool_partial_kernel<3, ATTN><<<grid, 128, 0, st>>>(p); break;
    case 4: pool_partial_kernel<4, ATTN><<<grid, 128, 0, st>>>(p); break;
    case 5: case 6: pool_partial_kernel<6, ATTN><<<grid, 128, 0, st>>>(p); break;
    default: pool_partial_kernel<8, ATTN><<<grid, 128, 0, st>>>(p); break;
  }
  TT_CHECK_LAUNCH();
  return TT_OK;
}

}  // namespace tt

extern "C" __attribute__((visibility("default"))) int tt_pool_partial_gather(const float* table, int64_t N_local, int64_t row_lo, int64_t N_total,
                                                                int owns_invalid, const float* row_logits, float zero_row_logit,
                                                                const int64_t* idx, const float* w, float* partial, int B, int S, int D,
                                                                void* stream) {
  TT_CHECK_ARG(table && idx && w && partial, "null pointer");
  TT_CHECK_ARG(B >= 0 && S >= 1 && D >= 4 && D % 4 == 0 && D <= 1024, "need B >= 0, S >= 1, D % 4 == 0, 4 <= D <= 1024");
  TT_CHECK_ARG(N_local >= 1 && row_lo >= 0 && N_total >= row_lo + N_local, "bad shard bounds");
  TT_CHECK_ARG(((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(partial)) & 15) == 0, "table and partial must be 16-byte aligned");
  if (B == 0) return TT_OK;
  PartialParams p{};
  p.table = table; p.row_logits = row_logits; p.idx = idx; p.w = w; p.partial = partial;
  p.N_local = N_local; p.row_lo = row_lo; p.N_total = N_total; p.zero_row_logit = zero_row_logit;
  p.owns_invalid = owns_invalid; p.B = B; p.S = S; p.D = D;
  return row_logits ? launch_partial<true>(p, (cudaStream_t)stream) : launch_partial<false>(p, (cudaStream_t)stream);
}

extern "C" __attribute__((visibility("default"))) int tt_pool_partial_merge(const float* partials_g, int G, int attention, float* out, int B, int D,
                                                               void* stream) {
  TT_CHECK_ARG(partials_g && out, "null pointer");
  TT_CHECK_ARG(G >= 1 && B >= 0 && D >= 1, "need G >= 1, B >= 0, D >= 1");
  if (B == 0) return TT_OK;
  pool_partial_merge_kernel<<<(B + 3) / 4, 128, 0, (cudaStream_t)stream>>>(partials_g, G, attention, out, B, D);
  TT_CHECK_LAUNCH();
  return TT_OK;
}

// ------------------------------------------------------------------------------------------
// Backward of the pooling op (training callers of the reference: src/models/two_tower.py:212,
// src/training/trainer.py:216-236).  One warp per buyer, two passes over the buyer's rows (the second one hits L2):
//   y = sum_s coef_s x_s,  out = y / max(||y||, 1e-12)                        (forward, buyer_tower.py:58-66 / :89-99)
//   dy = (g - out (out.g)) / ||y||            (dy = g / 1e-12 below the clamp, as torch.nn.functional.normalize)
//   dx_s = coef_s dy  (pooling part),  t_s = x_s . dy
//   weighted_avg: dw_s = (t_s - y.dy) / (sum w + 1e-8)
//   attention   : dc_s = coef_s (t_s - y.dy);  dlogit_s = w_s dc_s;  dw_s = logit_s dc_s
// The score MLP's own backward (dlogit -> dx, dW1, db1, dW2, db2) is two plain GEMMs and stays with the caller.
namespace tt {

struct PoolBwdParams {
  const float* x; const float* w; const float* logits; const float* g;
  float* dx; float* dw; float* dlogit;
  int B, S, D;
};

template <int NV, bool ATTN>
__global__ void __launch_bounds__(128)
pool_backward_kernel(const PoolBwdParams p) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= p.B) return;
  const int S = p.S, D = p.D, nvalid4 = D >> 2;
  const float* wb = p.w + (long long)b * S;
  const float* lb = ATTN ? p.logits + (long long)b * S : nullptr;
  const float4* xb = reinterpret_cast<const float4*>(p.x + (long long)b * S * D);
  // coefficients
  float m = -INFINITY, tot = 0.f;
  if (ATTN) {
    for (int s = lane; s < S; s += 32) m = fmaxf(m, __ldg(lb + s) * __ldg(wb + s));
    m = warp_max(m);
    for (int s = lane; s < S; s += 32) tot += expf(__ldg(lb + s) * __ldg(wb + s) - m);
    tot = warp_sum(tot);
  } else {
    for (int s = lane; s < S; s += 32) tot += __ldg(wb + s);
    tot = warp_sum(tot) + 1e-8f;
  }
  auto coef_of = [&](int s) -> float {
    return ATTN ? expf(__ldg(lb + s) * __ldg(wb + s) - m) / tot : __ldg(wb + s) / tot;
  };
  // pass 1: y
  float4 y[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) y[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < S; ++s) {
    const float c = coef_of(s);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c4 = v * 32 + lane;
      if (c4 < nvalid4) {
        const float4 xv = __ldg(xb + (long long)s * nvalid4 + c4);
        y[v].x = fmaf(xv.x, c, y[v].x); y[v].y = fmaf(xv.y, c, y[v].y);
        y[v].z = fmaf(xv.z, c, y[v].z); y[v].w = fmaf(xv.w, c, y[v].w);
      }
    }
  }
  float ss = 0.f, yg = 0.f;
  float4 gv[NV];
  const float4* gb = reinterpret_cast<const float4*>(p.g + (long long)b * D);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c4 = v * 32 + lane;
    gv[v] = (c4 < nvalid4) ? __ldg(gb + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
    ss += y[v].x * y[v].x + y[v].y * y[v].y + y[v].z * y[v].z + y[v].w * y[v].w;
    yg += y[v].x * gv[v].x + y[v].y * gv[v].y + y[v].z * gv[v].z + y[v].w * gv[v].w;
  }
  ss = warp_sum(ss);
  yg = warp_sum(yg);
  const float nrm = sqrtf(ss);
  // dy = (g - out (out.g)) / n with out = y / n;  below the eps clamp out = y / eps and dy = g / eps
  float4 dy[NV];
  float ydy = 0.f;
  const bool clamped = !(nrm > 1e-12f);
  const float inv = clamped ? 1e12f : 1.0f / nrm;
  const float proj = clamped ? 0.f : yg * inv * inv;          // (out.g)/n * (1/n) applied to y
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    dy[v].x = (gv[v].x - y[v].x * proj) * inv; dy[v].y = (gv[v].y - y[v].y * proj) * inv;
    dy[v].z = (gv[v].z - y[v].z * proj) * inv; dy[v].w = (gv[v].w - y[v].w * proj) * inv;
    ydy += y[v].x * dy[v].x + y[v].y * dy[v].y + y[v].z * dy[v].z + y[v].w * dy[v].w;
  }
  ydy = warp_sum(ydy);
  // pass 2: t_s = x_s . dy, dx_s = coef_s dy
  float4* dxb = p.dx ? reinterpret_cast<float4*>(p.dx + (long long)b * S * D) : nullptr;
  for (int s = 0; s < S; ++s) {
    const float c = coef_of(s);
    float t = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c4 = v * 32 + lane;
      if (c4 < nvalid4) {
        const float4 xv = __ldg(xb + (long long)s * nvalid4 + c4);
        t += xv.x * dy[v].x + xv.y * dy[v].y + xv.z * dy[v].z + xv.w * dy[v].w;
        if (dxb) dxb[(long long)s * nvalid4 + c4] = make_float4(c * dy[v].x, c * dy[v].y, c * dy[v].z, c * dy[v].w);
      }
    }
    t = warp_sum(t);
    if (lane == 0) {
      if (ATTN) {
        const float dc = c * (t - ydy);
        p.dlogit[(long long)b * S + s] = __ldg(wb + s) * dc;
        p.dw[(long long)b * S + s] = __ldg(lb + s) * dc;
      } else {
        p.dw[(long long)b * S + s] = (t - ydy) / tot;
      }
    }
  }
}

template <bool ATTN>
static int launch_pool_bwd(const PoolBwdParams& p, cudaStream_t st) {
  const int grid = (p.B + 3) / 4;
  switch ((p.D + 127) / 128) {
    case 1: pool_backward_kernel<1, ATTN><<<grid, 128, 0, st>>>(p); break;
    case 2: pool_backward_kernel<2, ATTN><<<grid, 128, 0, st>>>(p); break;
    case 3: pool_backward_kernel<3, ATTN><<<grid, 128, 0, st>>>(p); break;
    case 4: pool_backward_kernel<4, ATTN><<<grid, 128, 0, st>>>(p); break;
    case 5: case 6: pool_backward_kernel<6, ATTN><<<grid, 128, 0, st>>>(p); break;
    default: pool_backward_kernel<8, ATTN><<<grid, 128, 0, st>>>(p); break;
  }
  TT_CHECK_LAUNCH();
  return TT_OK;
}

}  // namespace tt

extern "C" __attribute__((visibility("default"))) int tt_pool_backward(const float* x, const float* w, const float* logits, const float* g,
                                                          float* dx, float* dw, float* dlogit, int B, int S, int D, void* stream) {
  TT_CHECK_ARG(x && w && g && dw, "null pointer");
  TT_CHECK_ARG((logits == nullptr) == (dlogit == nullptr), "logits and dlogit go together (attention mode)");
  TT_CHECK_ARG(B >= 0 && S >= 1 && D >= 4 && D % 4 == 0 && D <= 1024, "need B >= 0, S >= 1, D % 4 == 0, 4 <= D <= 1024");
  TT_CHECK_ARG(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0,
               "x, g and dx must be 16-byte aligned");
  if (B == 0) return TT_OK;
  PoolBwdParams p{};
  p.x = x; p.w = w; p.logits = logits; p.g = g; p.dx = dx; p.dw = dw; p.dlogit = dlogit; p.B = B; p.S = S; p.D = D;
  return logits ? launch_pool_bwd<true>(p, (cudaStream_t)stream) : launch_pool_bwd<false>(p, (cudaStream_t)stream);
}
