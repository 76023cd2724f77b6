/*
 * cpu_baseline.c -- the reference's scan as a TIMED CPU BASELINE (not a parity
 * oracle).  TEST/BENCH INFRASTRUCTURE ONLY: bench.py's cpu_baseline leg and
 * `--impl reference` are the only callers.
 *
 * Restates search_documents (src/search/mod.rs:77-120) the way it executes on
 * the host: one cosine per row (simsimd-style f32 SIMD lanes, FMA allowed),
 * the FULL result vector, a stable sort, take(top_k).  Built with
 * -O3 -march=x86-64-v3 (AVX2+FMA; portable to the GPU box host) -fopenmp, plus an AVX-512
 * kernel (function-level target attribute) picked at run time when the host has it -- simsimd
 * 6.5 dispatches the same way (its skylake / sapphire back ends), so an AVX-512 Xeon runs the
 * wider kernel here too (orc_baseline_isa() says which).  threads == 1 is the faithful configuration (the
 * reference loop is single-threaded); threads > 1 is reported separately and
 * labelled "not reference behaviour".
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#define ORC_OK 0
#define ORC_ERR_NOMEM -2
typedef struct { double d; uint64_t row; } orc_hit;

/* Stable bottom-up merge sort on distance only: equal distances keep their
 * input (row) order, which is what Rust's slice::sort_by guarantees. */
static int orc_stable_sort_hits(orc_hit *a, uint64_t n) {
  if (n < 2) return ORC_OK;
  orc_hit *tmp = (orc_hit *)malloc(sizeof(orc_hit) * n);
  if (!tmp) return ORC_ERR_NOMEM;
  orc_hit *src = a, *dst = tmp;
  for (uint64_t w = 1; w < n; w *= 2) {
    for (uint64_t lo = 0; lo < n; lo += 2 * w) {
      uint64_t mid = lo + w < n ? lo + w : n;
      uint64_t hi = lo + 2 * w < n ? lo + 2 * w : n;
      uint64_t i = lo, j = mid, k = lo;
      while (i < mid && j < hi) {
        if (src[j].d < src[i].d) dst[k++] = src[j++]; /* right only if strictly smaller */
        else dst[k++] = src[i++];
      }
      while (i < mid) dst[k++] = src[i++];
      while (j < hi) dst[k++] = src[j++];
    }
    orc_hit *s = src; src = dst; dst = s;
  }
  if (src != a) memcpy(a, src, sizeof(orc_hit) * n);
  free(tmp);
  return ORC_OK;
}

/* ------------------------------------------------- timed CPU baseline ------
 * The reference's scan as it would run on the host:
 * contiguous matrix, one cosine per row, FULL result vector, stable sort,
 * take(top_k) -- single thread, because src/search/mod.rs:84-104 is a plain
 * nested loop.  The inner product uses f32 lanes (what simsimd's SIMD backends
 * do) so the compiler can vectorise with -O3 -march=native; the selection logic
 * is identical to orc_search_documents.  Used ONLY as bench.py's cpu_baseline /
 * --impl reference timing leg, never for parity. */
#include <immintrin.h>
/* simsimd-style Haswell kernel: three 8-lane f32 FMA accumulators over the 256
 * dims, horizontal reduce, then the normalisation in f64. */
static inline double orc_cosine_simd_like(const float *a, const float *b) {
  __m256 ab = _mm256_setzero_ps(), a2 = _mm256_setzero_ps(), b2 = _mm256_setzero_ps();
  for (int i = 0; i < 256; i += 8) {
    __m256 x = _mm256_loadu_ps(a + i), y = _mm256_loadu_ps(b + i);
    ab = _mm256_fmadd_ps(x, y, ab);
    a2 = _mm256_fmadd_ps(x, x, a2);
    b2 = _mm256_fmadd_ps(y, y, b2);
  }
  float t[8];
  double sab = 0, sa2 = 0, sb2 = 0;
  _mm256_storeu_ps(t, ab); for (int l = 0; l < 8; ++l) sab += t[l];
  _mm256_storeu_ps(t, a2); for (int l = 0; l < 8; ++l) sa2 += t[l];
  _mm256_storeu_ps(t, b2); for (int l = 0; l < 8; ++l) sb2 += t[l];
  if (sa2 == 0.0 && sb2 == 0.0) return 0.0;
  if (sab == 0.0) return 1.0;
  double r = 1.0 - sab / (sqrt(sa2) * sqrt(sb2));
  return r > 0.0 ? r : 0.0;
}

/* simsimd-style Skylake-X kernel: 16-lane f32 FMA accumulators, reduce, normalise in f64. */
__attribute__((target("avx512f"))) static double orc_cosine_avx512(const float *a, const float *b) {
  __m512 ab = _mm512_setzero_ps(), a2 = _mm512_setzero_ps(), b2 = _mm512_setzero_ps();
  for (int i = 0; i < 256; i += 16) {
    __m512 x = _mm512_loadu_ps(a + i), y = _mm512_loadu_ps(b + i);
    ab = _mm512_fmadd_ps(x, y, ab);
    a2 = _mm512_fmadd_ps(x, x, a2);
    b2 = _mm512_fmadd_ps(y, y, b2);
  }
  double sab = (double)_mm512_reduce_add_ps(ab), sa2 = (double)_mm512_reduce_add_ps(a2), sb2 = (double)_mm512_reduce_add_ps(b2);
  if (sa2 == 0.0 && sb2 == 0.0) return 0.0;
  if (sab == 0.0) return 1.0;
  double r = 1.0 - sab / (sqrt(sa2) * sqrt(sb2));
  return r > 0.0 ? r : 0.0;
}

static int orc_have_avx512(void) {
  static int have = -1;
  if (have < 0) { __builtin_cpu_init(); have = __builtin_cpu_supports("avx512f") ? 1 : 0; }
  return have;
}
const char *orc_baseline_isa(void) { return orc_have_avx512() ? "avx512f" : "avx2+fma"; }

__attribute__((target("avx512f"))) static void orc_distances_avx512(const float *rows, int64_t r0, int64_t r1, const float *q, double *dist) {
  for (int64_t r = r0; r < r1; ++r) dist[r] = orc_cosine_avx512(q, rows + (uint64_t)r * 256);
}
static void orc_distances_avx2(const float *rows, int64_t r0, int64_t r1, const float *q, double *dist) {
  for (int64_t r = r0; r < r1; ++r) dist[r] = orc_cosine_simd_like(q, rows + (uint64_t)r * 256);
}

int orc_baseline_search(const float *rows, uint64_t n_rows, const float *q,
                        uint64_t top_k, int has_max, double max_distance,
                        int threads, uint64_t cap, uint64_t *out_row,
                        double *out_distance, uint64_t *out_n) {
  double thr = has_max ? max_distance : 100.0;
  orc_hit *hits = (orc_hit *)malloc(sizeof(orc_hit) * (n_rows ? n_rows : 1));
  if (!hits) return ORC_ERR_NOMEM;
  uint64_t m = 0;
  const int wide = orc_have_avx512();
  if (threads <= 1) {
    /* one cosine per row, pushed straight into the result vector (mod.rs:84-104); processed in
     * blocks of 4096 rows only so that the ISA dispatch sits outside the inner loop */
    double blk[4096];
    for (uint64_t r0 = 0; r0 < n_rows; r0 += 4096) {
      const uint64_t r1 = r0 + 4096 < n_rows ? r0 + 4096 : n_rows;
      if (wide) orc_distances_avx512(rows, (int64_t)r0, (int64_t)r1, q, blk - r0);
      else orc_distances_avx2(rows, (int64_t)r0, (int64_t)r1, q, blk - r0);
      for (uint64_t r = r0; r < r1; ++r)
        if (blk[r - r0] < thr) { hits[m].d = blk[r - r0]; hits[m].row = r; m++; }
    }
  } else {
    /* "not reference behaviour": all-cores variant.  Distances in parallel,
     * then the same sequential filter so row order (stability) is kept. */
    double *dist = (double *)malloc(sizeof(double) * (n_rows ? n_rows : 1));
    if (!dist) { free(hits); return ORC_ERR_NOMEM; }
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(threads)
#endif
    for (int64_t b = 0; b < (int64_t)((n_rows + 4095) / 4096); ++b) {
      const int64_t r0 = b * 4096, r1 = r0 + 4096 < (int64_t)n_rows ? r0 + 4096 : (int64_t)n_rows;
      if (wide) orc_distances_avx512(rows, r0, r1, q, dist);
      else orc_distances_avx2(rows, r0, r1, q, dist);
    }
    for (uint64_t r = 0; r < n_rows; ++r)
      if (dist[r] < thr) { hits[m].d = dist[r]; hits[m].row = r; m++; }
    free(dist);
  }
  int rc = orc_stable_sort_hits(hits, m);
  if (rc != ORC_OK) { free(hits); return rc; }
  uint64_t n_out = has_max ? m : (m < top_k ? m : top_k);
  *out_n = n_out;
  for (uint64_t i = 0; i < n_out && i < cap; ++i) {
    if (out_row) out_row[i] = hits[i].row;
    if (out_distance) out_distance[i] = hits[i].d;
  }
  free(hits);
  return ORC_OK;
}


int orc_baseline_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
