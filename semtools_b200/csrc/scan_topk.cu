// K1 (+ in-kernel K4): cosine scan with running top-K', tree merge, exact re-rank.
//
// Replaces the loop/sort/take of search_documents (reference
// src/search/mod.rs:84-119; one simsimd f32::cosine per line at :86) and the
// filtered nearest query of Store::search_line_embeddings
// (src/workspace/store.rs:495-543).
//
// HBM-bound: the only large traffic is the corpus matrix, 1 KiB per row, read
// exactly once with 128-bit coalesced loads (8 lanes cover one 128-byte line of a
// row; a warp instruction touches 4 full lines).  Everything else (query, K'
// candidates per CTA, k results) is bytes.
//
// Exactness: the streaming pass ranks rows by an fp32 approximate cosine and keeps
// the best K' = 32*E per warp; the surviving K' of the whole grid are re-scored by
// one thread each in the oracle's canonical arithmetic (f64 accumulation in index
// order) and sorted by (distance,row).  The result is accepted only if the worst
// kept approximate score proves that no dropped row can reach the k-th exact
// distance (STB_SCORE_EPS); otherwise status[1]=0 and the host runs the collect
// path below, which is exact for any input.
#include <math_constants.h>

#include "common.cuh"

#ifndef STB_SCAN_U
#define STB_SCAN_U 2   // rows per 8-lane group per iteration (loads in flight = 8*U float4)
#endif
#ifndef STB_SCAN_LD
#define STB_SCAN_LD 0  // 0: ld.global.nc.L1::no_allocate  1: __ldcs  2: __ldg
#endif

__device__ __forceinline__ float4 stb_ld_stream(const float4 *p) {
#if STB_SCAN_LD == 0
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
#elif STB_SCAN_LD == 1
  return __ldcs(p);
#else
  return __ldg(p);
#endif
}

struct ScanArgs {
  const float4 *rows;        // local row 0
  uint64_t n_virtual;        // rows to scan (== n_rows when no ranges)
  const float *q;            // 256 f32 (device)
  const uint64_t *vstart;    // [n_ranges+1] virtual prefix (ranges mode)
  const uint64_t *rbegin;    // [n_ranges]   local first row of each range
  uint32_t n_ranges;
};

// virtual row -> local row (ranges mode): largest idx with vstart[idx] <= v.
__device__ __forceinline__ uint32_t stb_map_row(const ScanArgs &a, uint64_t v) {
  uint32_t lo = 0, hi = a.n_ranges;
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(a.vstart + mid) <= v) lo = mid; else hi = mid;
  }
  return (uint32_t)(__ldg(a.rbegin + lo) + (v - __ldg(a.vstart + lo)));
}

// Approximate-cosine scan.  Calls sink(score, local_row) once per 4*U-row tile with
// a warp-uniform control flow; lanes that do not represent a row pass -inf.
template <int U, bool RANGES, class Sink>
__device__ __forceinline__ void stb_scan_rows(const ScanArgs &args, Sink &sink) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 3;   // row group inside the warp
  const int j = lane & 7;    // 16-byte column slot inside the group
  // query slice of this lane: float4 index j + 8*i
  float4 q[8];
  const float4 *q4 = reinterpret_cast<const float4 *>(args.q);
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = __ldg(q4 + j + 8 * i);
  // ||q||^2 in the same fixed order as every row norm
  float qx = 0.f, qy = 0.f, qz = 0.f, qw = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    qx = fmaf(q[i].x, q[i].x, qx); qy = fmaf(q[i].y, q[i].y, qy);
    qz = fmaf(q[i].z, q[i].z, qz); qw = fmaf(q[i].w, q[i].w, qw);
  }
  float b2 = (qx + qy) + (qz + qw);
  b2 += __shfl_xor_sync(0xffffffffu, b2, 4);
  b2 += __shfl_xor_sync(0xffffffffu, b2, 2);
  b2 += __shfl_xor_sync(0xffffffffu, b2, 1);
  // b2 == 0 in fp32 is either a true zero vector or an underflowed tiny one
  bool q_any = false;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    q_any |= (q[i].x != 0.f) | (q[i].y != 0.f) | (q[i].z != 0.f) | (q[i].w != 0.f);
  q_any = __any_sync(0xffffffffu, q_any);   // all four groups hold the same query
  const bool q_zero = (b2 == 0.f) && !q_any;
  const bool q_bad = !q_zero && !(b2 >= 1e-30f && b2 <= 1e30f);  // NaN/inf/denormal/underflow
  const float rq = q_zero ? 0.f : rsqrtf(b2);

  const uint64_t tile_rows = 4 * U;
  const uint64_t n_tiles = (args.n_virtual + tile_rows - 1) / tile_rows;
  const uint64_t warps_total = (uint64_t)gridDim.x * (blockDim.x >> 5);
  const uint64_t warp_id = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);

  for (uint64_t tile = warp_id; tile < n_tiles; tile += warps_total) {
    float4 a[U][8];
    uint32_t row[U];
    bool valid[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      uint64_t v = tile * tile_rows + (uint64_t)(u * 4 + g);
      valid[u] = v < args.n_virtual;
      uint64_t vc = valid[u] ? v : (args.n_virtual - 1);
      row[u] = RANGES ? stb_map_row(args, vc) : (uint32_t)vc;
      const float4 *p = args.rows + (size_t)row[u] * STB_ROW_F4 + j;
#pragma unroll
      for (int i = 0; i < 8; ++i) a[u][i] = stb_ld_stream(p + 8 * i);
    }
    float sc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float dx = 0.f, dy = 0.f, dz = 0.f, dw = 0.f;
      float nx = 0.f, ny = 0.f, nz = 0.f, nw = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        dx = fmaf(a[u][i].x, q[i].x, dx); dy = fmaf(a[u][i].y, q[i].y, dy);
        dz = fmaf(a[u][i].z, q[i].z, dz); dw = fmaf(a[u][i].w, q[i].w, dw);
        nx = fmaf(a[u][i].x, a[u][i].x, nx); ny = fmaf(a[u][i].y, a[u][i].y, ny);
        nz = fmaf(a[u][i].z, a[u][i].z, nz); nw = fmaf(a[u][i].w, a[u][i].w, nw);
      }
      float ab = (dx + dy) + (dz + dw);
      float a2 = (nx + ny) + (nz + nw);
      ab += __shfl_xor_sync(0xffffffffu, ab, 4);
      a2 += __shfl_xor_sync(0xffffffffu, a2, 4);
      ab += __shfl_xor_sync(0xffffffffu, ab, 2);
      a2 += __shfl_xor_sync(0xffffffffu, a2, 2);
      ab += __shfl_xor_sync(0xffffffffu, ab, 1);
      a2 += __shfl_xor_sync(0xffffffffu, a2, 1);
      float s;
      if (a2 == 0.f) {
        // rare: a true zero row (simsimd rules: d = 0 vs a zero query, else d = 1) or a
        // row so small that its fp32 squared norm underflowed -> forced candidate.
        // All 8 lanes of the group hold the same a2, so the group votes together.
        bool nz = false;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          nz |= (a[u][i].x != 0.f) | (a[u][i].y != 0.f) | (a[u][i].z != 0.f) | (a[u][i].w != 0.f);
        nz = __any_sync(0xffu << (8 * g), nz);
        s = nz ? CUDART_INF_F : (q_zero ? 1.f : 0.f);
      } else if (q_bad || !(a2 >= 1e-30f && a2 <= 1e30f)) s = CUDART_INF_F;  // forced candidate
      else s = ab * rsqrtf(a2) * rq;
      sc[u] = valid[u] ? s : -CUDART_INF_F;
    }
    float s = -CUDART_INF_F;
    uint32_t r = 0;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (j == u) { s = sc[u]; r = row[u]; }
    sink(s, r);
  }
}

// ---------------------------------------------------------------- top-K' sink ---
template <int E>
struct TopSink {
  float ls[E];
  uint32_t lr[E];
  float thr;   // min score in the list (warp-uniform); -inf while not full
  int lane;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int e = 0; e < E; ++e) { ls[e] = -CUDART_INF_F; lr[e] = 0xffffffffu; }
    thr = -CUDART_INF_F;
    lane = threadIdx.x & 31;
  }
  __device__ __forceinline__ float warp_min(float m) const {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, off));
    return m;
  }
  __device__ __forceinline__ void insert(float cs, uint32_t cr) {
    float m = ls[0];
    int mi = 0;
#pragma unroll
    for (int e = 1; e < E; ++e)
      if (ls[e] < m) { m = ls[e]; mi = e; }
    unsigned owners = __ballot_sync(0xffffffffu, m == thr);
    int owner = __ffs(owners) - 1;
    if (lane == owner) {
#pragma unroll
      for (int e = 0; e < E; ++e)
        if (e == mi) { ls[e] = cs; lr[e] = cr; }
    }
    m = ls[0];
#pragma unroll
    for (int e = 1; e < E; ++e) m = fminf(m, ls[e]);
    thr = warp_min(m);
  }
  __device__ __forceinline__ void operator()(float s, uint32_t r) {
    unsigned mask = __ballot_sync(0xffffffffu, s > thr);
    while (mask) {
      int src = __ffs(mask) - 1;
      mask &= mask - 1;
      float cs = __shfl_sync(0xffffffffu, s, src);
      uint32_t cr = __shfl_sync(0xffffffffu, r, src);
      if (cs > thr) insert(cs, cr);
    }
  }
};

// Ascending bitonic sort of n (power of two <= STB_SORT_CAP) keys in shared memory.
__device__ __forceinline__ void stb_cta_sort_keys(uint64_t *keys, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        int ixj = i ^ jj;
        if (ixj > i) {
          uint64_t x = keys[i], y = keys[ixj];
          bool up = ((i & k) == 0);
          if ((x > y) == up) { keys[i] = y; keys[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
}

struct TopkArgs {
  ScanArgs scan;
  uint64_t row_base;
  uint64_t *keys;            // [gridDim.x][KP] best keys of every CTA
  unsigned int *counters;    // [0] arrival ticket
  stb_hit *out_hits;
  uint32_t *out_status;
  uint32_t top_k;
};

// Sort the first `c` keys of skeys (padded with INVALID to a power of two >= KP).
__device__ __forceinline__ int stb_pad_and_sort(uint64_t *skeys, int c, int min_n) {
  int n = min_n;
  while (n < c) n <<= 1;
  for (int i = c + threadIdx.x; i < n; i += blockDim.x) skeys[i] = STB_KEY_INVALID;
  __syncthreads();
  stb_cta_sort_keys(skeys, n);
  return n;
}

#define STB_RR_STRIDE 260   // floats per staged row (1 KiB + 16 B pad: conflict-free LDS.128)

template <int E, int U, bool RANGES>
__global__ void __launch_bounds__(STB_SCAN_THREADS, STB_SCAN_MINB)
stb_scan_topk_kernel(const TopkArgs args) {
  constexpr int KP = 32 * E;
  __shared__ uint64_t skeys[STB_SORT_CAP];
  __shared__ unsigned int s_T, s_cnt, s_ticket, s_over;
  __shared__ __align__(16) float sq[STB_D];
  __shared__ __align__(16) float srows[32 * STB_RR_STRIDE];
  __shared__ double s_d[KP];
  __shared__ uint64_t s_r[KP];
  __shared__ int s_nv[2];

  TopSink<E> sink;
  sink.init();
  stb_scan_rows<U, RANGES>(args.scan, sink);

  // ---- CTA merge ------------------------------------------------------------------
  // T = max over warps of the warp list minimum is a lower bound of the CTA's KP-th
  // best score (that warp alone holds KP keys >= its minimum), so only keys >= T can
  // matter: compact those (typically ~KP..2KP of the 8*KP) and sort the small set.
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { s_T = 0u; s_cnt = 0u; s_over = 0u; }
  __syncthreads();
  if (lane == 0) atomicMax(&s_T, stb_f2ord(sink.thr));
  __syncthreads();
  {
    const unsigned T = s_T;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      if (sink.lr[e] != 0xffffffffu && stb_f2ord(sink.ls[e]) >= T) {
        unsigned idx = atomicAdd(&s_cnt, 1u);       // <= 8*KP <= STB_SORT_CAP entries exist
        skeys[idx] = stb_make_key(sink.ls[e], sink.lr[e]);
      }
    }
  }
  __syncthreads();
  int c = (int)s_cnt;
  stb_pad_and_sort(skeys, c, KP);

  // ---- publish the CTA's best KP; last CTA to arrive finishes the query -------------
  if (gridDim.x > 1) {
    uint64_t *mine = args.keys + (size_t)blockIdx.x * KP;
    for (int i = threadIdx.x; i < KP; i += blockDim.x) mine[i] = skeys[i];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(args.counters, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    if (threadIdx.x == 0) args.counters[0] = 0u;   // re-arm for the next launch

    // Exact radix select (12+12+8 bits of the ordered score) of the KP-th best score
    // among the gridDim.x*KP published keys: distribution-independent, 4 coalesced
    // passes over <= a few hundred KB that sit in L2.  The histogram aliases the
    // re-rank staging buffer, which is not live yet.
    unsigned int *hist = reinterpret_cast<unsigned int *>(srows);
    const size_t total = (size_t)gridDim.x * KP;
    // Keys are cached in registers, STB_KPT per thread per round, all loads of a round
    // issued before any use (one L2 latency per round; one round covers 10240 keys,
    // i.e. the whole E=1 grid, which is then read exactly once).
    constexpr int KPT = 40;
    uint64_t kreg[KPT];
    const int n_rounds = (int)((total + (size_t)KPT * STB_SCAN_THREADS - 1) / ((size_t)KPT * STB_SCAN_THREADS));
    auto load_round = [&](int round) {
#pragma unroll
      for (int j = 0; j < KPT; ++j) {
        size_t i = ((size_t)round * KPT + j) * STB_SCAN_THREADS + threadIdx.x;
        kreg[j] = (i < total) ? __ldcg(args.keys + i) : STB_KEY_INVALID;
      }
    };
    if (n_rounds == 1) load_round(0);
    unsigned prefix = 0u, pmask = 0u;
    int need = KP;
    bool select_all = false;
#pragma unroll 1
    for (int pass = 0; pass < 3; ++pass) {
      const int shift = pass == 0 ? 20 : (pass == 1 ? 8 : 0);
      const int nbins = pass == 2 ? 256 : 4096;
      for (int i = threadIdx.x; i < nbins; i += blockDim.x) hist[i] = 0u;
      if (threadIdx.x == 0) { s_T = 0u; s_cnt = 0u; }
      __syncthreads();
#pragma unroll 1
      for (int round = 0; round < n_rounds; ++round) {
        if (n_rounds > 1) load_round(round);
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
          const uint64_t key = kreg[j];
          if (key != STB_KEY_INVALID) {
            unsigned o = ~(unsigned)(key >> 32);
            if ((o & pmask) == prefix) atomicAdd(&hist[(o >> shift) & (unsigned)(nbins - 1)], 1u);
          }
        }
      }
      __syncthreads();
      // largest digit d with  count(digit > d) < need <= count(digit >= d)
      if (threadIdx.x < 32) {
        const int per_lane = nbins / 32;               // bins per lane, lane 0 = top digits
        const int hi = nbins - 1 - lane * per_lane;    // this lane scans hi, hi-1, ...
        unsigned seg = 0u;
        for (int b = 0; b < per_lane; ++b) seg += hist[hi - b];
        unsigned incl = seg;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          unsigned v = __shfl_up_sync(0xffffffffu, incl, off);
          if (lane >= off) incl += v;
        }
        const unsigned before = incl - seg;
        const unsigned all = __shfl_sync(0xffffffffu, incl, 31);
        if (all < (unsigned)need) { if (lane == 0) s_T = 0xffffffffu; }    // fewer valid keys than needed
        else if (before < (unsigned)need && (unsigned)need <= incl) {
          unsigned acc = before;
          for (int b = 0; b < per_lane; ++b) {
            unsigned h = hist[hi - b];
            if (acc + h >= (unsigned)need) { s_T = (unsigned)(hi - b); s_cnt = acc; break; }
            acc += h;
          }
        }
      }
      __syncthreads();
      if (s_T == 0xffffffffu) { select_all = true; break; }
      prefix |= s_T << shift;
      pmask |= (unsigned)(nbins - 1) << shift;
      need -= (int)s_cnt;
      __syncthreads();
    }
    const unsigned Tg = select_all ? 0u : prefix;   // exact ord of the KP-th best score
    __syncthreads();
    if (threadIdx.x == 0) s_cnt = 0u;
    __syncthreads();
#pragma unroll 1
    for (int round = 0; round < n_rounds; ++round) {
      if (n_rounds > 1) load_round(round);
#pragma unroll
      for (int j = 0; j < KPT; ++j) {
        const uint64_t key = kreg[j];
        if (key != STB_KEY_INVALID && ~(unsigned)(key >> 32) >= Tg) {
          unsigned idx = atomicAdd(&s_cnt, 1u);
          if (idx < STB_SORT_CAP) skeys[idx] = key; else s_over = 1u;
        }
      }
    }
    __syncthreads();
    if (s_over) {
      // > 1024 keys share the KP-th best score (mass duplication): they cannot be
      // ranked here; the host runs the exact collect pass (status[1] = 0).
      if (threadIdx.x == 0) {
        args.out_status[0] = 0u; args.out_status[1] = 0u; args.out_status[2] = 0xffffffffu;
        args.out_status[3] = (uint32_t)KP;
      }
      return;
    }
    c = (int)s_cnt;
    stb_pad_and_sort(skeys, c, KP);
  }

  // ---- exact re-rank of the best KP in canonical arithmetic --------------------------
  // Rows are staged through shared memory (coalesced, one DRAM latency), then one
  // thread per candidate accumulates (ab, q2, r2) with f64 FMAs in index order:
  // f32 x f32 products are exact in f64, so this equals orc_cosine_f32 bit for bit.
  for (int i = threadIdx.x; i < STB_D; i += blockDim.x) sq[i] = __ldg(args.scan.q + i);
  if (threadIdx.x < 2) s_nv[threadIdx.x] = 0;
  __syncthreads();
  for (int chunk = 0; chunk < E; ++chunk) {
    {
      constexpr int PER = 32 * STB_ROW_F4 / STB_SCAN_THREADS;   // float4 per thread
      float4 v[PER];
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int t = threadIdx.x + u * STB_SCAN_THREADS;
        const uint64_t key = skeys[chunk * 32 + (t >> 6)];
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (key != STB_KEY_INVALID)
          v[u] = __ldg(args.scan.rows + (size_t)stb_key_row(key) * STB_ROW_F4 + (t & 63));
      }
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int t = threadIdx.x + u * STB_SCAN_THREADS;
        *reinterpret_cast<float4 *>(srows + (t >> 6) * STB_RR_STRIDE + (t & 63) * 4) = v[u];
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      const int ci = chunk * 32 + threadIdx.x;
      uint64_t key = skeys[ci];
      double d = CUDART_INF;
      uint64_t grow = 0xffffffffffffffffull;
      if (key != STB_KEY_INVALID) {
        const float4 *rp = reinterpret_cast<const float4 *>(srows + threadIdx.x * STB_RR_STRIDE);
        const float4 *qp = reinterpret_cast<const float4 *>(sq);
        double ab = 0.0, q2 = 0.0, r2 = 0.0;
#pragma unroll 4
        for (int i = 0; i < STB_ROW_F4; ++i) {
          float4 v = rp[i];
          float4 w = qp[i];
          ab = fma((double)w.x, (double)v.x, ab); q2 = fma((double)w.x, (double)w.x, q2); r2 = fma((double)v.x, (double)v.x, r2);
          ab = fma((double)w.y, (double)v.y, ab); q2 = fma((double)w.y, (double)w.y, q2); r2 = fma((double)v.y, (double)v.y, r2);
          ab = fma((double)w.z, (double)v.z, ab); q2 = fma((double)w.z, (double)w.z, q2); r2 = fma((double)v.z, (double)v.z, r2);
          ab = fma((double)w.w, (double)v.w, ab); q2 = fma((double)w.w, (double)w.w, q2); r2 = fma((double)v.w, (double)v.w, r2);
        }
        double dist;
        if (q2 == 0.0 && r2 == 0.0) dist = 0.0;
        else if (ab == 0.0) dist = 1.0;
        else {
          double t = 1.0 - ab / (sqrt(q2) * sqrt(r2));
          dist = t > 0.0 ? t : 0.0;
        }
        atomicAdd(&s_nv[0], 1);                     // valid candidates
        if (dist < 100.0) {                         // max_distance.unwrap_or(100.0), strict
          d = dist;
          grow = args.row_base + (uint64_t)stb_key_row(key);
          atomicAdd(&s_nv[1], 1);                   // passing
        }
      }
      s_d[ci] = d;
      s_r[ci] = grow;
    }
    __syncthreads();
  }
  // bitonic sort of the KP (distance,row) pairs
  for (int k = 2; k <= KP; k <<= 1) {
    for (int jj = k >> 1; jj > 0; jj >>= 1) {
      int i = threadIdx.x;
      if (i < KP) {
        int ixj = i ^ jj;
        if (ixj > i) {
          double da = s_d[i], db = s_d[ixj];
          uint64_t ra = s_r[i], rb = s_r[ixj];
          bool up = ((i & k) == 0);
          bool gt = stb_hit_less(db, rb, da, ra);
          if (gt == up) { s_d[i] = db; s_r[i] = rb; s_d[ixj] = da; s_r[ixj] = ra; }
        }
      }
      __syncthreads();
    }
  }
  const int n_valid = s_nv[0], n_pass = s_nv[1];
  const uint32_t k = args.top_k;
  const uint32_t n_out = min((uint32_t)n_pass, k);
  for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
    stb_hit h;
    h.distance = (i < n_out) ? s_d[i] : CUDART_INF;
    h.row = (i < n_out) ? s_r[i] : 0xffffffffffffffffull;
    args.out_hits[i] = h;
  }
  if (threadIdx.x == 0) {
    bool complete;
    if (n_valid < KP) complete = true;   // every scorable row is in the candidate set
    else {
      float s_min = stb_key_score(skeys[KP - 1]);
      complete = (n_out == k) && ((1.0 - (double)s_min - STB_SCORE_EPS) > s_d[k - 1]);
    }
    args.out_status[0] = n_out;
    args.out_status[1] = complete ? 1u : 0u;
    args.out_status[2] = (uint32_t)n_valid;
    args.out_status[3] = (uint32_t)KP;
  }
}

uint32_t stb_scan_topk_max_k(void) { return 96; }

static int stb_pick_e(uint32_t top_k) {
  if (top_k <= 16) return 1;   // K' = 32
  if (top_k <= 40) return 2;   // K' = 64
  return 4;                    // K' = 128
}

template <int E, bool RANGES>
static int stb_launch_topk_t(stb_ctx *ctx, const TopkArgs &a) {
  auto kern = stb_scan_topk_kernel<E, STB_SCAN_U, RANGES>;
  int occ = 0;
  STB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, STB_SCAN_THREADS, 0));
  if (occ < 1) occ = 1;
  uint64_t tiles = (a.scan.n_virtual + 4 * STB_SCAN_U - 1) / (4 * STB_SCAN_U);
  uint64_t want = (tiles + STB_SCAN_WARPS - 1) / STB_SCAN_WARPS;
  uint64_t grid = (uint64_t)ctx->sm_count * occ;
  if (want < grid) grid = want < 1 ? 1 : want;
  // scratch: keys for all tree levels (< 2 * grid lists), counters (< grid groups)
  size_t need_keys = (size_t)2 * grid * 32 * E + STB_SORT_CAP;
  if (need_keys > ctx->block_keys_cap || grid + 8 > ctx->counters_cap) {
    stb_set_error("scan scratch too small (grid=%llu)", (unsigned long long)grid);
    return STB_ERR_STATE;
  }
  kern<<<(unsigned)grid, STB_SCAN_THREADS, 0, ctx->stream>>>(a);
  STB_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return STB_OK;
}

int stb_launch_scan_topk(stb_ctx *ctx, const float *rows, uint64_t n_rows,
                         uint64_t row_base, const float *q_dev, uint32_t top_k,
                         const uint64_t *ranges_dev, uint32_t n_ranges,
                         uint64_t n_virtual, stb_hit *out_hits_dev,
                         uint32_t *out_status_dev) {
  (void)n_rows;
  TopkArgs a;
  a.scan.rows = reinterpret_cast<const float4 *>(rows);
  a.scan.n_virtual = n_virtual;
  a.scan.q = q_dev;
  a.scan.vstart = ranges_dev;
  a.scan.rbegin = ranges_dev ? ranges_dev + (n_ranges + 1) : nullptr;
  a.scan.n_ranges = n_ranges;
  a.row_base = row_base;
  a.keys = ctx->block_keys;
  a.counters = ctx->counters;
  a.out_hits = out_hits_dev;
  a.out_status = out_status_dev;
  a.top_k = top_k;
  const bool rg = n_ranges > 0;
  switch (stb_pick_e(top_k)) {
    case 1: return rg ? stb_launch_topk_t<1, true>(ctx, a) : stb_launch_topk_t<1, false>(ctx, a);
    case 2: return rg ? stb_launch_topk_t<2, true>(ctx, a) : stb_launch_topk_t<2, false>(ctx, a);
    default: return rg ? stb_launch_topk_t<4, true>(ctx, a) : stb_launch_topk_t<4, false>(ctx, a);
  }
}

// ------------------------------------------------------------------ collect path ---
struct CollectSink {
  float floor_;
  uint32_t *out;
  unsigned long long *count;
  uint64_t cap;
  __device__ __forceinline__ void operator()(float s, uint32_t r) {
    bool hit = (s >= floor_) && (s > -CUDART_INF_F);   // -inf marks lanes that carry no row
    unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (mask == 0) return;
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (hit) {
      unsigned long long idx = base + __popc(mask & ((1u << lane) - 1));
      if (idx < cap) out[idx] = r;
    }
  }
};

struct CollectArgs {
  ScanArgs scan;
  float cos_floor;
  uint32_t *out;
  unsigned long long *count;
  uint64_t cap;
};

template <int U, bool RANGES>
__global__ void __launch_bounds__(STB_SCAN_THREADS, STB_SCAN_MINB)
stb_scan_collect_kernel(const CollectArgs args) {
  CollectSink sink{args.cos_floor, args.out, args.count, args.cap};
  stb_scan_rows<U, RANGES>(args.scan, sink);
}

int stb_launch_scan_collect(stb_ctx *ctx, const float *rows, uint64_t n_rows,
                            const float *q_dev, float cos_floor,
                            const uint64_t *ranges_dev, uint32_t n_ranges,
                            uint64_t n_virtual) {
  (void)n_rows;
  CollectArgs a;
  a.scan.rows = reinterpret_cast<const float4 *>(rows);
  a.scan.n_virtual = n_virtual;
  a.scan.q = q_dev;
  a.scan.vstart = ranges_dev;
  a.scan.rbegin = ranges_dev ? ranges_dev + (n_ranges + 1) : nullptr;
  a.scan.n_ranges = n_ranges;
  a.cos_floor = cos_floor;
  a.out = ctx->collect_rows;
  a.count = ctx->collect_count;
  a.cap = ctx->collect_cap;
  STB_CUDA(cudaMemsetAsync(ctx->collect_count, 0, sizeof(unsigned long long), ctx->stream));
  uint64_t tiles = (n_virtual + 4 * STB_SCAN_U - 1) / (4 * STB_SCAN_U);
  uint64_t want = (tiles + STB_SCAN_WARPS - 1) / STB_SCAN_WARPS;
  uint64_t grid = (uint64_t)ctx->sm_count * STB_SCAN_MINB;
  if (want < grid) grid = want < 1 ? 1 : want;
  if (n_ranges > 0)
    stb_scan_collect_kernel<STB_SCAN_U, true><<<(unsigned)grid, STB_SCAN_THREADS, 0, ctx->stream>>>(a);
  else
    stb_scan_collect_kernel<STB_SCAN_U, false><<<(unsigned)grid, STB_SCAN_THREADS, 0, ctx->stream>>>(a);
  STB_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return STB_OK;
}

// Exact canonical distance of each collected row; one thread per row.
__global__ void stb_exact_kernel(const float4 *rows, uint64_t row_base, const float *q,
                                 const uint32_t *row_ids, uint64_t m, double limit,
                                 stb_hit *hits, uint64_t m_padded,
                                 unsigned long long *pass_count) {
  __shared__ __align__(16) float sq[STB_D];
  for (int i = threadIdx.x; i < STB_D; i += blockDim.x) sq[i] = __ldg(q + i);
  __syncthreads();
  uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m_padded) return;
  stb_hit h;
  h.distance = CUDART_INF;
  h.row = 0xffffffffffffffffull;
  if (idx < m) {
    uint32_t row = row_ids[idx];
    const float4 *rp = rows + (size_t)row * STB_ROW_F4;
    const float4 *qp = reinterpret_cast<const float4 *>(sq);
    double ab = 0.0, q2 = 0.0, r2 = 0.0;
#pragma unroll 4
    for (int i = 0; i < STB_ROW_F4; ++i) {
      float4 v = __ldg(rp + i);
      float4 w = qp[i];
      ab = fma((double)w.x, (double)v.x, ab); q2 = fma((double)w.x, (double)w.x, q2); r2 = fma((double)v.x, (double)v.x, r2);
      ab = fma((double)w.y, (double)v.y, ab); q2 = fma((double)w.y, (double)w.y, q2); r2 = fma((double)v.y, (double)v.y, r2);
      ab = fma((double)w.z, (double)v.z, ab); q2 = fma((double)w.z, (double)w.z, q2); r2 = fma((double)v.z, (double)v.z, r2);
      ab = fma((double)w.w, (double)v.w, ab); q2 = fma((double)w.w, (double)w.w, q2); r2 = fma((double)v.w, (double)v.w, r2);
    }
    double dist;
    if (q2 == 0.0 && r2 == 0.0) dist = 0.0;
    else if (ab == 0.0) dist = 1.0;
    else {
      double t = 1.0 - ab / (sqrt(q2) * sqrt(r2));
      dist = t > 0.0 ? t : 0.0;
    }
    if (dist < limit) {
      h.distance = dist;
      h.row = row_base + (uint64_t)row;
      atomicAdd(pass_count, 1ull);
    }
  }
  hits[idx] = h;
}

int stb_launch_exact(stb_ctx *ctx, const float *rows, uint64_t row_base,
                     const float *q_dev, const uint32_t *row_ids, uint64_t m,
                     double limit, stb_hit *hits, uint64_t m_padded,
                     unsigned long long *pass_count) {
  STB_CUDA(cudaMemsetAsync(pass_count, 0, sizeof(unsigned long long), ctx->stream));
  if (m_padded == 0) return STB_OK;
  unsigned blocks = (unsigned)((m_padded + 127) / 128);
  stb_exact_kernel<<<blocks, 128, 0, ctx->stream>>>(reinterpret_cast<const float4 *>(rows),
                                                     row_base, q_dev, row_ids, m, limit, hits,
                                                     m_padded, pass_count);
  STB_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return STB_OK;
}

// Global bitonic sort of hits by (distance,row): one launch per (k,j) step above
// the shared-memory span, fused steps inside a 1024-element span.
__global__ void stb_bitonic_global_step(stb_hit *h, uint64_t n, uint64_t k, uint64_t j) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t ixj = i ^ j;
  if (ixj > i) {
    stb_hit a = h[i], b = h[ixj];
    bool up = ((i & k) == 0);
    bool gt = stb_hit_less(b.distance, b.row, a.distance, a.row);
    if (gt == up) { h[i] = b; h[ixj] = a; }
  }
}

// Sorts all steps with j < 1024 of stage k inside shared memory (span 1024).
__global__ void stb_bitonic_local(stb_hit *h, uint64_t n, uint64_t k_first, uint64_t k_last) {
  __shared__ double sd[1024];
  __shared__ uint64_t sr[1024];
  const uint64_t base = (uint64_t)blockIdx.x * 1024;
  for (int t = threadIdx.x; t < 1024; t += blockDim.x) {
    stb_hit x = h[base + t];
    sd[t] = x.distance; sr[t] = x.row;
  }
  __syncthreads();
  for (uint64_t k = k_first; k <= k_last; k <<= 1) {
    uint64_t jstart = (k >> 1) < 512 ? (k >> 1) : 512;
    for (uint64_t j = jstart; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < 1024; t += blockDim.x) {
        uint64_t i = base + t, ixj = i ^ j;
        if (ixj > i) {
          int u = (int)(ixj - base);
          bool up = ((i & k) == 0);
          bool gt = stb_hit_less(sd[u], sr[u], sd[t], sr[t]);
          if (gt == up) {
            double td = sd[t]; uint64_t tr = sr[t];
            sd[t] = sd[u]; sr[t] = sr[u]; sd[u] = td; sr[u] = tr;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int t = threadIdx.x; t < 1024; t += blockDim.x) {
    stb_hit x; x.distance = sd[t]; x.row = sr[t];
    h[base + t] = x;
  }
}

int stb_launch_sort_hits(stb_ctx *ctx, stb_hit *hits, uint64_t n) {
  if (n < 2) return STB_OK;
  if (n < 1024 || (n & (n - 1))) { stb_set_error("sort size must be a power of two >= 1024"); return STB_ERR_ARG; }
  unsigned lblocks = (unsigned)(n / 1024);
  // stages k = 2..1024 entirely local
  stb_bitonic_local<<<lblocks, 256, 0, ctx->stream>>>(hits, n, 2, 1024);
  ctx->kernel_launches++;
  for (uint64_t k = 2048; k <= n; k <<= 1) {
    for (uint64_t j = k >> 1; j >= 1024; j >>= 1) {
      unsigned gblocks = (unsigned)((n + 255) / 256);
      stb_bitonic_global_step<<<gblocks, 256, 0, ctx->stream>>>(hits, n, k, j);
      ctx->kernel_launches++;
    }
    stb_bitonic_local<<<lblocks, 256, 0, ctx->stream>>>(hits, n, k, k);
    ctx->kernel_launches++;
  }
  STB_CUDA(cudaGetLastError());
  return STB_OK;
}
