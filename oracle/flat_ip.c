/*
 * flat_ip.c — plain-C restatement of exact inner-product top-k search.  TEST INFRASTRUCTURE ONLY
 * (checker for tests/ and the timed CPU baseline of bench.py; never linked into the product).
 *
 * Restates what the reference obtains from faiss-cpu (>=1.7.4, requirements.txt:26; source NOT
 * under /root/reference) at src/inference/vector_db.py:160 / :197:
 *     scores, idx = IndexFlatIP.search(q, k)
 * i.e. for each query the fp32 inner product with every stored fp32 row, the k largest, sorted by
 * descending score.  Published algorithm being restated: linear scan + per-query binary min-heap of
 * size k whose root is replaced only by a strictly larger score (so among equal scores the lower
 * row id is kept), final heap sort.  Parity at the faiss boundary is UNPINNED (no golden vectors in
 * the reference); this file is cross-checked against oracle/flat_ip_oracle.py in tests/test_oracle.py.
 *
 * Build: see oracle/Makefile  (gcc -O3 -pthread -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

typedef struct { float s; int64_t i; } ent_t;

/* "a is worse than b": lower score, or equal score and higher id */
static inline int worse(ent_t a, ent_t b) { return a.s < b.s || (a.s == b.s && a.i > b.i); }

static void sift_down(ent_t* h, int n, int p) {
  for (;;) {
    int l = 2 * p + 1, r = l + 1, m = p;
    if (l < n && worse(h[l], h[m])) m = l;
    if (r < n && worse(h[r], h[m])) m = r;
    if (m == p) return;
    ent_t t = h[p]; h[p] = h[m]; h[m] = t; p = m;
  }
}
static void heap_push(ent_t* h, int* n, int k, ent_t e) {
  if (*n < k) {
    int c = (*n)++;
    h[c] = e;
    while (c > 0) { int p = (c - 1) / 2; if (!worse(h[c], h[p])) break; ent_t t = h[c]; h[c] = h[p]; h[p] = t; c = p; }
  } else if (worse(h[0], e)) { h[0] = e; sift_down(h, k, 0); }
}
static int cmp_desc(const void* a, const void* b) {
  const ent_t *x = (const ent_t*)a, *y = (const ent_t*)b;
  if (worse(*y, *x)) return -1;
  if (worse(*x, *y)) return 1;
  return 0;
}

static inline float dotf(const float* a, const float* b, int d) {
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int j = 0;
  for (; j + 8 <= d; j += 8)
    for (int u = 0; u < 8; ++u) acc[u] += a[j + u] * b[j + u];
  float s = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  for (; j < d; ++j) s += a[j] * b[j];
  return s;
}

typedef struct {
  const float* x; const float* q; int64_t lo, hi; int d, nb, k; ent_t* heaps; int* cnt;
} job_t;

static void* scan_rows(void* arg) {
  job_t* j = (job_t*)arg;
  for (int64_t r = j->lo; r < j->hi; ++r) {
    const float* row = j->x + r * j->d;
    for (int b = 0; b < j->nb; ++b) {
      ent_t e = { dotf(row, j->q + (int64_t)b * j->d, j->d), r };
      heap_push(j->heaps + (size_t)b * j->k, j->cnt + b, j->k, e);
    }
  }
  return NULL;
}

/* x [n,d] stored rows, q [nq,d] (both already normalised by the caller, as vector_db.py does),
 * out_s [nq,k] / out_i [nq,k]; k <= n.  Threads split the rows (pthreads).  Returns 0. */
int oracle_flat_ip_search(const float* x, int64_t n, int d, const float* q, int nq, int k,
                          float* out_s, int64_t* out_i, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  const int QB = 8;   /* queries sharing one pass over a row block */
  for (int q0 = 0; q0 < nq; q0 += QB) {
    const int nb = nq - q0 < QB ? nq - q0 : QB;
    ent_t* heaps = (ent_t*)malloc((size_t)nthreads * nb * k * sizeof(ent_t));
    int* cnt = (int*)calloc((size_t)nthreads * nb, sizeof(int));
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < nthreads; ++t) {
      job_t jb = { x, q + (int64_t)q0 * d, n * t / nthreads, n * (t + 1) / nthreads, d, nb, k,
                   heaps + (size_t)t * nb * k, cnt + t * nb };
      jobs[t] = jb;
      if (nthreads > 1) pthread_create(&th[t], NULL, scan_rows, &jobs[t]);
      else scan_rows(&jobs[t]);
    }
    if (nthreads > 1) for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    for (int b = 0; b < nb; ++b) {
      int tot = 0;
      for (int t = 0; t < nthreads; ++t) tot += cnt[t * nb + b];
      ent_t* all = (ent_t*)malloc((size_t)(tot > 0 ? tot : 1) * sizeof(ent_t));
      int m = 0;
      for (int t = 0; t < nthreads; ++t) {
        memcpy(all + m, heaps + ((size_t)t * nb + b) * k, (size_t)cnt[t * nb + b] * sizeof(ent_t));
        m += cnt[t * nb + b];
      }
      qsort(all, (size_t)tot, sizeof(ent_t), cmp_desc);
      for (int j = 0; j < k; ++j) {
        out_s[(int64_t)(q0 + b) * k + j] = j < tot ? all[j].s : -INFINITY;
        out_i[(int64_t)(q0 + b) * k + j] = j < tot ? all[j].i : -1;
      }
      free(all);
    }
    free(heaps); free(cnt);
  }
  return 0;
}

/* x / (||x|| + 1e-8) row-wise, vector_db.py:44-45 */
void oracle_normalize_rows(const float* x, int64_t n, int d, float* out) {
  for (int64_t r = 0; r < n; ++r) {
    double ss = 0;   /* numpy's pairwise fp32 sum differs from a serial fp32 sum by ~1e-7 rel; use a wide accumulator then round */
    for (int j = 0; j < d; ++j) ss += (double)x[r * d + j] * x[r * d + j];
    const float den = (float)sqrt(ss) + 1e-8f;
    for (int j = 0; j < d; ++j) out[r * d + j] = x[r * d + j] / den;
  }
}
