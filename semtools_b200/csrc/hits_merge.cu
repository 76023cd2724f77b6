// K4: merge per-shard hit lists by (distance,row) -- the final sort_by + take of
// reference src/search/mod.rs:107-119 (and src/workspace/store.rs:538-543) applied
// across row shards.  Latency-bound: n_lists*per_list*16 bytes in, top_k*16 out.
#include <math_constants.h>

#include "common.cuh"

#define STB_MERGE_CAP 4096   // hits one CTA sorts in shared memory (64 KiB)

__global__ void __launch_bounds__(256, 1)
stb_hits_merge_kernel(const stb_hit *lists, uint32_t total, uint32_t n_sort, uint32_t top_k,
                      stb_hit *out) {
  extern __shared__ __align__(16) unsigned char smem[];
  double *sd = reinterpret_cast<double *>(smem);
  uint64_t *sr = reinterpret_cast<uint64_t *>(smem + (size_t)n_sort * sizeof(double));
  for (uint32_t i = threadIdx.x; i < n_sort; i += blockDim.x) {
    if (i < total) {
      stb_hit h = lists[i];
      // NaN distances never leave the scan; treat anything unordered as padding
      bool ok = h.distance == h.distance && h.row != 0xffffffffffffffffull;
      sd[i] = ok ? h.distance : CUDART_INF;
      sr[i] = ok ? h.row : 0xffffffffffffffffull;
    } else {
      sd[i] = CUDART_INF;
      sr[i] = 0xffffffffffffffffull;
    }
  }
  __syncthreads();
  for (uint32_t k = 2; k <= n_sort; k <<= 1) {
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < n_sort; i += blockDim.x) {
        uint32_t ixj = i ^ j;
        if (ixj > i) {
          bool up = ((i & k) == 0);
          bool gt = stb_hit_less(sd[ixj], sr[ixj], sd[i], sr[i]);
          if (gt == up) {
            double td = sd[i]; uint64_t tr = sr[i];
            sd[i] = sd[ixj]; sr[i] = sr[ixj]; sd[ixj] = td; sr[ixj] = tr;
          }
        }
      }
      __syncthreads();
    }
  }
  for (uint32_t i = threadIdx.x; i < top_k; i += blockDim.x) {
    stb_hit h;
    h.distance = (i < n_sort) ? sd[i] : CUDART_INF;
    h.row = (i < n_sort) ? sr[i] : 0xffffffffffffffffull;
    out[i] = h;
  }
}

// Batched form: lists[n_lists][nq][per_list] -> out[nq][top_k]; one CTA per query.
__global__ void __launch_bounds__(128)
stb_hits_merge_batch_kernel(const stb_hit *lists, uint32_t n_lists, uint32_t nq, uint32_t per_list,
                            uint32_t n_sort, uint32_t top_k, stb_hit *out) {
  extern __shared__ __align__(16) unsigned char smem[];
  double *sd = reinterpret_cast<double *>(smem);
  uint64_t *sr = reinterpret_cast<uint64_t *>(smem + (size_t)n_sort * sizeof(double));
  const uint32_t q = blockIdx.x, total = n_lists * per_list;
  for (uint32_t i = threadIdx.x; i < n_sort; i += blockDim.x) {
    double d = CUDART_INF;
    uint64_t r = 0xffffffffffffffffull;
    if (i < total) {
      const stb_hit h = lists[((size_t)(i / per_list) * nq + q) * per_list + (i % per_list)];
      if (h.distance == h.distance && h.row != 0xffffffffffffffffull) { d = h.distance; r = h.row; }
    }
    sd[i] = d; sr[i] = r;
  }
  __syncthreads();
  for (uint32_t k = 2; k <= n_sort; k <<= 1)
    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
      for (uint32_t i = threadIdx.x; i < n_sort; i += blockDim.x) {
        const uint32_t ixj = i ^ j;
        if (ixj > i) {
          const bool up = ((i & k) == 0);
          const bool gt = stb_hit_less(sd[ixj], sr[ixj], sd[i], sr[i]);
          if (gt == up) {
            double td = sd[i]; uint64_t tr = sr[i];
            sd[i] = sd[ixj]; sr[i] = sr[ixj]; sd[ixj] = td; sr[ixj] = tr;
          }
        }
      }
      __syncthreads();
    }
  for (uint32_t i = threadIdx.x; i < top_k; i += blockDim.x) {
    stb_hit h;
    h.distance = (i < n_sort) ? sd[i] : CUDART_INF;
    h.row = (i < n_sort) ? sr[i] : 0xffffffffffffffffull;
    out[(size_t)q * top_k + i] = h;
  }
}

int stb_launch_hits_merge_batch(stb_ctx *ctx, const stb_hit *lists_dev, uint32_t n_lists, uint32_t nq,
                                uint32_t per_list, uint32_t top_k, stb_hit *out_dev) {
  const uint64_t total = (uint64_t)n_lists * per_list;
  if (total > 2048) { stb_set_error("hits_merge_batch: %llu hits per query exceed 2048", (unsigned long long)total); return STB_ERR_ARG; }
  uint32_t n_sort = 2;
  while (n_sort < total) n_sort <<= 1;
  stb_hits_merge_batch_kernel<<<nq, 128, (size_t)n_sort * 16, ctx->stream>>>(lists_dev, n_lists, nq, per_list, n_sort, top_k, out_dev);
  STB_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return STB_OK;
}

int stb_launch_hits_merge(stb_ctx *ctx, const stb_hit *lists_dev, uint32_t n_lists,
                          uint32_t per_list, uint32_t top_k, stb_hit *out_dev) {
  uint64_t total = (uint64_t)n_lists * per_list;
  if (total > STB_MERGE_CAP) {
    stb_set_error("hits_merge: %llu hits exceed the %d-hit merge capacity",
                  (unsigned long long)total, STB_MERGE_CAP);
    return STB_ERR_ARG;
  }
  uint32_t n_sort = 2;
  while (n_sort < total) n_sort <<= 1;
  size_t smem = (size_t)n_sort * 16;
  STB_ATTR_ONCE(ctx, STB_ATTR_MERGE,
                cudaFuncSetAttribute(stb_hits_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     STB_MERGE_CAP * 16));
  stb_hits_merge_kernel<<<1, 256, smem, ctx->stream>>>(lists_dev, (uint32_t)total, n_sort, top_k,
                                                      out_dev);
  STB_CUDA(cudaGetLastError());
  ctx->kernel_launches++;
  return STB_OK;
}
