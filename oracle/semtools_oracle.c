/*
 * semtools_oracle.c -- CPU restatement of the semtools `search` hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it,
 * and there only as the checker (or as the timed CPU baseline).  The product
 * (semtools_b200/) never links, imports or executes anything in oracle/.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference, semtools v3.0.0).  Two third-party crates hold the
 * arithmetic and are NOT vendored in the reference tree:
 *   model2vec-rs 0.1.3 (Cargo.lock:2445)  -> orc_pool_ids / orc_prepare_ids
 *   simsimd      6.5.1 (Cargo.lock:4057)  -> orc_cosine_f32*
 * Their algorithms are restated from the published crates.  The reference holds
 * NO numeric golden vectors for them (its tests assert structure only, SURVEY
 * section 4), and no Rust toolchain exists here to run the reference itself:
 *
 *      >>>  NUMERIC PARITY FOR model2vec / simsimd IS UNPINNED.  <<<
 *
 * What IS pinned: FNV-1a-64 (public known-answer vectors + store.rs:651-661),
 * the search_documents control flow (strict `<`, stable sort, top_k vs
 * threshold, context windows; src/search/mod.rs:77-120) and the store query
 * fixture (src/workspace/store.rs:814-850); see tests/golden/.
 *
 * Build: gcc -O2 -ffp-contract=off  (contraction MUST stay off: Rust never
 * fuses a*b+c, so the f32 pooling arithmetic is mul-then-add, each rounded).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_OK 0
#define ORC_ERR_RANGE -1
#define ORC_ERR_NOMEM -2

/* ------------------------------------------------------------------ a12 ---
 * fnv1a_hash, src/workspace/store.rs:651-661: offset basis
 * 0xcbf29ce484222325, prime 0x100000001b3, xor-then-wrapping-multiply. */
uint64_t orc_fnv1a64(const uint8_t *bytes, uint64_t len) {
  uint64_t h = 0xcbf29ce484222325ULL;
  for (uint64_t i = 0; i < len; ++i) {
    h ^= (uint64_t)bytes[i];
    h *= 0x100000001b3ULL;
  }
  return h;
}

/* DocMeta::id, src/workspace/store.rs:75-80: fnv1a(path bytes). */
uint64_t orc_doc_id(const char *path, uint64_t path_len) {
  return orc_fnv1a64((const uint8_t *)path, path_len);
}

/* LineEmbedding::id, src/workspace/store.rs:82-89: fnv1a(path bytes followed by
 * the i32 line number in little-endian byte order). */
uint64_t orc_line_id(const char *path, uint64_t path_len, int32_t line_number) {
  uint64_t h = 0xcbf29ce484222325ULL;
  for (uint64_t i = 0; i < path_len; ++i) {
    h ^= (uint64_t)(uint8_t)path[i];
    h *= 0x100000001b3ULL;
  }
  uint32_t u = (uint32_t)line_number;
  for (int b = 0; b < 4; ++b) {
    h ^= (uint64_t)((u >> (8 * b)) & 0xffu);
    h *= 0x100000001b3ULL;
  }
  return h;
}

/* ------------------------------------------------------------------- a3 ---
 * Tail of StaticModel::encode_with_args (model2vec-rs 0.1.3; call sites
 * src/search/mod.rs:69 [max_length 2048], :138 [encode_single -> 512]):
 * after tokenisation, `token_ids.retain(|id| id != unk_id)` then
 * `token_ids.truncate(max_length)`.  Returns the new length; ids compacted in
 * place.  has_unk == 0 means the tokenizer has no unk token. */
uint64_t orc_prepare_ids(uint32_t *ids, uint64_t n, int has_unk, uint32_t unk_id,
                         int has_max, uint64_t max_length) {
  uint64_t w = 0;
  for (uint64_t i = 0; i < n; ++i) {
    if (has_unk && ids[i] == unk_id) continue;
    ids[w++] = ids[i];
  }
  if (has_max && w > max_length) w = max_length;
  return w;
}

/* ------------------------------------------------------------------- a4 ---
 * StaticModel::pool_ids (model2vec-rs 0.1.3; reached from
 * src/search/mod.rs:69,138,153 and src/cmds/search.rs:136,154):
 *   sum[d] += E[map(id)][d] * w(id)      sequentially over tokens, f32,
 *                                        multiply rounded, then add rounded
 *   sum[d] /= max(cnt,1) as f32
 *   if normalize: norm = sqrt(sum_d sum[d]^2)  (sequential f32 fold from 0)
 *                 sum[d] /= max(norm, 1e-12)
 * map(id) = mapping[id] if id < n_mapping else id;  w(id) = weights[id] if
 * id < n_weights else 1.0.  A row index outside the table panics upstream; here
 * it is ORC_ERR_RANGE. */
int orc_pool_ids(const float *E, uint64_t V, uint32_t D, const float *weights,
                 uint64_t n_weights, const uint32_t *mapping, uint64_t n_mapping,
                 int normalize, const uint32_t *ids, uint64_t T, float *out) {
  for (uint32_t d = 0; d < D; ++d) out[d] = 0.0f;
  uint64_t cnt = 0;
  for (uint64_t t = 0; t < T; ++t) {
    uint64_t tok = ids[t];
    uint64_t row = (mapping && tok < n_mapping) ? (uint64_t)mapping[tok] : tok;
    float scale = (weights && tok < n_weights) ? weights[tok] : 1.0f;
    if (row >= V) return ORC_ERR_RANGE;
    const float *e = E + row * (uint64_t)D;
    for (uint32_t d = 0; d < D; ++d) {
      float prod = e[d] * scale; /* rounded product, never fused */
      out[d] = out[d] + prod;
    }
    cnt += 1;
  }
  float denom = (float)(cnt > 0 ? cnt : 1);
  for (uint32_t d = 0; d < D; ++d) out[d] = out[d] / denom;
  if (normalize) {
    float ss = 0.0f;
    for (uint32_t d = 0; d < D; ++d) {
      float sq = out[d] * out[d];
      ss = ss + sq;
    }
    float norm = sqrtf(ss);
    if (!(norm > 1e-12f)) norm = 1e-12f; /* f32::max(norm, 1e-12); NaN -> 1e-12 */
    for (uint32_t d = 0; d < D; ++d) out[d] = out[d] / norm;
  }
  return ORC_OK;
}

/* Batched form over a CSR batch (offsets[n_lines+1], ids[offsets[n_lines]]):
 * row i of out = pool_ids(ids[offsets[i]..offsets[i+1]]).  Order preserved,
 * one row per input line (encode_with_args contract). */
int orc_embed_csr(const float *E, uint64_t V, uint32_t D, const float *weights,
                  uint64_t n_weights, const uint32_t *mapping, uint64_t n_mapping,
                  int normalize, const uint64_t *offsets, const uint32_t *ids,
                  uint64_t n_lines, float *out) {
  int rc = ORC_OK;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256)
#endif
  for (int64_t i = 0; i < (int64_t)n_lines; ++i) {
    int r = orc_pool_ids(E, V, D, weights, n_weights, mapping, n_mapping, normalize,
                         ids + offsets[i], offsets[i + 1] - offsets[i],
                         out + (uint64_t)i * D);
    if (r != ORC_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
      rc = r;
    }
  }
  return rc;
}

/* ------------------------------------------------------------------- a6 ---
 * f32::cosine (simsimd 6.5.1 SpatialSimilarity; call site
 * src/search/mod.rs:86).  CANONICAL ORACLE ARITHMETIC: inputs f32;
 * ab, a2, b2 accumulated in f64 in index order (each f32*f32 product is exact
 * in f64, so fused and unfused accumulation agree bit for bit);
 *   a2 == 0 && b2 == 0 -> 0;   ab == 0 -> 1;
 *   else max(0, 1 - ab / (sqrt(a2) * sqrt(b2))).   Result f64.
 * This is simsimd's `*_accurate` serial kernel shape; the SIMD backends differ
 * from it by < 1e-6 on unit-scale 256-d inputs, inside north_star's 1e-5. */
double orc_cosine_f32(const float *a, const float *b, uint64_t n) {
  double ab = 0.0, a2 = 0.0, b2 = 0.0;
  for (uint64_t i = 0; i < n; ++i) {
    double ai = (double)a[i], bi = (double)b[i];
    ab += ai * bi;
    a2 += ai * ai;
    b2 += bi * bi;
  }
  if (a2 == 0.0 && b2 == 0.0) return 0.0;
  if (ab == 0.0) return 1.0;
  double r = 1.0 - ab / (sqrt(a2) * sqrt(b2));
  return r > 0.0 ? r : 0.0;
}

/* simsimd's plain serial backend: f32 accumulators in index order and
 * 1 - ab * rsqrt(a2) * rsqrt(b2).  Kept only so tests can show that backend
 * choice moves the distance by far less than the 1e-5 tolerance. */
double orc_cosine_f32_serial32(const float *a, const float *b, uint64_t n) {
  float ab = 0.0f, a2 = 0.0f, b2 = 0.0f;
  for (uint64_t i = 0; i < n; ++i) {
    float p = a[i] * b[i];
    float pa = a[i] * a[i];
    float pb = b[i] * b[i];
    ab = ab + p;
    a2 = a2 + pa;
    b2 = b2 + pb;
  }
  if (a2 == 0.0f && b2 == 0.0f) return 0.0;
  if (ab == 0.0f) return 1.0;
  double r = 1.0 - (double)ab * (1.0 / sqrt((double)a2)) * (1.0 / sqrt((double)b2));
  return r > 0.0 ? r : 0.0;
}

/* All N distances of one query (the loop at src/search/mod.rs:84-86). */
void orc_distances(const float *rows, uint64_t n_rows, uint32_t D, const float *q,
                   double *out) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int64_t i = 0; i < (int64_t)n_rows; ++i)
    out[i] = orc_cosine_f32(rows + (uint64_t)i * D, q, D);
}

/* ------------------------------------------------------------------- a7 ---
 * search_documents, src/search/mod.rs:77-120, over a contiguous row-major
 * matrix whose rows are the lines of the documents in (doc order, line order);
 * doc_offsets[n_docs+1] gives each document's first row.
 *   :84-86   for doc, for line: distance = cosine(q, line)
 *   :88-89   keep iff distance < max_distance.unwrap_or(100.0)   (strict; a
 *            NaN distance fails the comparison and is dropped)
 *   :90-91   start = idx.saturating_sub(n_lines); end = min(len, idx+n_lines+1)
 *   :107-111 STABLE sort ascending by distance (slice::sort_by = merge sort);
 *            partial_cmp -> Equal on NaN (cannot occur after the filter)
 *   :115-119 max_distance set -> all; else take(top_k)
 * Ordering key is therefore (distance, global row).  Outputs are written up to
 * `cap` entries; *out_n is the full result count (may exceed cap). */
typedef struct {
  double d;
  uint64_t row;
} orc_hit;

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

int orc_search_documents(const float *rows, uint64_t n_rows, uint32_t D,
                         const uint64_t *doc_offsets, uint64_t n_docs,
                         const float *q, uint64_t n_lines, uint64_t top_k,
                         int has_max, double max_distance, uint64_t cap,
                         uint64_t *out_row, uint64_t *out_doc,
                         uint64_t *out_match_line, uint64_t *out_start,
                         uint64_t *out_end, double *out_distance,
                         uint64_t *out_n) {
  double thr = has_max ? max_distance : 100.0;
  orc_hit *hits = (orc_hit *)malloc(sizeof(orc_hit) * (n_rows ? n_rows : 1));
  if (!hits) return ORC_ERR_NOMEM;
  uint64_t m = 0;
  (void)n_docs;
  if (n_rows > 0 && doc_offsets[n_docs] != n_rows) { free(hits); return ORC_ERR_RANGE; }
  for (uint64_t r = 0; r < n_rows; ++r) {
    double d = orc_cosine_f32(q, rows + r * (uint64_t)D, D);
    if (d < thr) { hits[m].d = d; hits[m].row = r; m++; }
  }
  int rc = orc_stable_sort_hits(hits, m);
  if (rc != ORC_OK) { free(hits); return rc; }
  uint64_t n_out = has_max ? m : (m < top_k ? m : top_k);
  *out_n = n_out;
  uint64_t doc = 0;
  for (uint64_t i = 0; i < n_out && i < cap; ++i) {
    uint64_t r = hits[i].row;
    /* locate document: binary search in doc_offsets */
    uint64_t lo = 0, hi = n_docs;
    while (hi - lo > 1) {
      uint64_t mid = (lo + hi) / 2;
      if (doc_offsets[mid] <= r) lo = mid; else hi = mid;
    }
    doc = lo;
    uint64_t idx = r - doc_offsets[doc];
    uint64_t len = doc_offsets[doc + 1] - doc_offsets[doc];
    uint64_t start = idx > n_lines ? idx - n_lines : 0;
    uint64_t end = idx + n_lines + 1 < len ? idx + n_lines + 1 : len;
    if (out_row) out_row[i] = r;
    if (out_doc) out_doc[i] = doc;
    if (out_match_line) out_match_line[i] = idx;
    if (out_start) out_start[i] = start;
    if (out_end) out_end[i] = end;
    if (out_distance) out_distance[i] = hits[i].d;
  }
  free(hits);
  return ORC_OK;
}

/* Same scan restricted to a set of half-open row ranges (ascending, disjoint):
 * the semantic content of Store::search_line_embeddings' path filter,
 * src/workspace/store.rs:481-546:
 *   :489-491  empty subset or top_k == 0 -> empty
 *   :498-499  score_threshold = 1 - max_dist (Qdrant keeps score > threshold,
 *             i.e. distance < max_dist, strict)
 *   :517      limit 2*top_k per 1000-path chunk, then
 *   :538-543  stable sort ascending, truncate(top_k)  == global top_k over the
 *             selected rows ordered by (distance, row);  NOTE: unlike a7 the
 *             threshold does NOT lift the top_k cap here.
 *   :531      distance = 1 - score, reported as f32.
 * Qdrant's own arithmetic (pre-normalised rows, f32 SIMD dot) is third-party
 * and unpinned; the oracle uses the canonical cosine above and rounds to f32. */
int orc_store_search(const float *rows, uint64_t n_rows, uint32_t D,
                     const uint64_t *ranges, uint64_t n_ranges, const float *q,
                     uint64_t top_k, int has_max, float max_distance,
                     uint64_t cap, uint64_t *out_row, float *out_distance,
                     uint64_t *out_n) {
  *out_n = 0;
  if (n_ranges == 0 || top_k == 0) return ORC_OK;
  uint64_t total = 0;
  for (uint64_t i = 0; i < n_ranges; ++i) {
    if (ranges[2 * i + 1] < ranges[2 * i] || ranges[2 * i + 1] > n_rows) return ORC_ERR_RANGE;
    total += ranges[2 * i + 1] - ranges[2 * i];
  }
  orc_hit *hits = (orc_hit *)malloc(sizeof(orc_hit) * (total ? total : 1));
  if (!hits) return ORC_ERR_NOMEM;
  uint64_t m = 0;
  for (uint64_t i = 0; i < n_ranges; ++i) {
    for (uint64_t r = ranges[2 * i]; r < ranges[2 * i + 1]; ++r) {
      double d = orc_cosine_f32(q, rows + r * (uint64_t)D, D);
      if (has_max && !(d < (double)max_distance)) continue;
      if (d != d) continue;
      hits[m].d = d; hits[m].row = r; m++;
    }
  }
  int rc = orc_stable_sort_hits(hits, m);
  if (rc != ORC_OK) { free(hits); return rc; }
  uint64_t n_out = m < top_k ? m : top_k;
  *out_n = n_out;
  for (uint64_t i = 0; i < n_out && i < cap; ++i) {
    out_row[i] = hits[i].row;
    out_distance[i] = (float)hits[i].d;
  }
  free(hits);
  return ORC_OK;
}

/* Detects a build that contracted a*b+c into an FMA (returns 1 when the
 * arithmetic is the required unfused mul-then-add). */
int orc_selftest_unfused(void) {
  volatile float a = 1.0f + 0x1p-12f, b = 1.0f + 0x1p-12f, c = -(1.0f + 0x1p-11f);
  float x = a, y = b, z = c;
  float r = x * y + z;          /* unfused: product rounds to 1+2^-11, r == 0 */
  return r == 0.0f;
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
