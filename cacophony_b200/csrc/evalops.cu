// Row (f) kernels — the steps right after the encoders in the reference's evaluation drivers
// (src/eval/eval_caco_torch.py:289-408, src/eval/eval_utils.py:18-66) and the HEAR timestamp pooling
// (src/eval/heareval/embeddings/audio_embedding/caco_embeddings.py:124-129):
//   topk_rows        torch.argsort(-logits, dim=-1)[:, :k]  without sorting the whole row (k <= 32)
//   retrieval_hits   the per-query body of compute_retrieval_metric: R@1 / R@5 / R@10 / AP@10 from the top-10 keys
//   avg_pool_tokens  tf.nn.avg_pool(hidden, ksize=8, strides=8, 'VALID') over the token axis
// All three are HBM/L2-bound byte shuffling: one warp per row, coalesced loads, shuffle reductions.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "caco_b200.h"
#include "common.cuh"

namespace caco {

// ---------------------------------------------------------------------------------------------------- top-k
// One warp per row.  Every element gets a 64-bit rank  (order-preserving bits of the score) << 32 | ~column : larger rank
// = better, so ties resolve to the lower column and NaN scores (mapped below -inf) rank last.  Round r picks the largest
// rank strictly below pick r-1, so no scratch memory is needed; the row (<= a few tens of KB) stays in L1/L2 across the
// k rounds.
__device__ __forceinline__ unsigned long long topk_rank(float v, int c) {
  uint32_t u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);     // monotone map of non-NaN floats onto unsigned integers
  if (v != v) u = 0u;                                  // NaN: below -inf
  return ((unsigned long long)u << 32) | (uint32_t)(0xffffffffu - (uint32_t)c);
}

__global__ void __launch_bounds__(256)
topk_rows_kernel(const float* __restrict__ x, int rows, int cols, int ldx, int k, int* __restrict__ idx_out,
                 float* __restrict__ val_out) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (size_t)warp * ldx;
  unsigned long long prev = ~0ull;
  for (int r = 0; r < k; ++r) {
    unsigned long long best = 0ull;
    for (int c = lane; c < cols; c += 32) {
      const unsigned long long rk = topk_rank(__ldg(xr + c), c);
      if (rk < prev && rk > best) best = rk;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    const int bi = (int)(0xffffffffu - (uint32_t)(best & 0xffffffffull));
    if (lane == 0) {
      idx_out[(size_t)warp * k + r] = bi;
      if (val_out) val_out[(size_t)warp * k + r] = __ldg(xr + bi);
    }
    prev = best;
  }
}

int topk_rows(const float* x, int rows, int cols, int ldx, int k, int* idx_out, float* val_out, cudaStream_t stream) {
  if (!x || !idx_out || rows <= 0 || cols <= 0 || ldx < cols || k <= 0 || k > 32 || k > cols) return CACO_ERR_ARG;
  const int warps_per_block = 8;
  const int blocks = (rows + warps_per_block - 1) / warps_per_block;
  topk_rows_kernel<<<blocks, warps_per_block * 32, 0, stream>>>(x, rows, cols, ldx, k, idx_out, val_out);
  count_launch();
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------- retrieval hits
// eval_utils.py:26-56 for one query per thread.  topk [Q, k] (k >= 10 columns used: the reference looks at indices[i, :10]).
//   mode 0 ('ta', eval_utils.py:40-41): hit_j = (key_id[topk[q, j]] == gt_id[q])
//   mode 1 ('at', eval_utils.py:28-38): hit_j = (key_id[topk[q, j]] is in the query's ground-truth set) and that key id has
//           not been counted at an earlier rank; the ground-truth sets are given as a sorted int64 array of
//           gt_id[q] * n_key_ids + key_id pairs (binary search).
// out [Q] int32: bit j set = rank j+1 is a hit (`preds` of eval_utils.py:26-41).  R@k / AP@10 (eval_utils.py:43-56) are then a
// few float64 operations per query on the host, exactly as the reference computes them.
__global__ void retrieval_hits_kernel(const int* __restrict__ topk, int ldk, int Q, const int* __restrict__ key_id,
                                      const int* __restrict__ gt_id, const long long* __restrict__ pairs, int n_pairs,
                                      long long n_key_ids, int mode, int* __restrict__ out) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  int seen[10];
  int n_seen = 0;
  bool hit[10];
  const int g = gt_id[q];
  for (int j = 0; j < 10; ++j) {
    const int idx = topk[(size_t)q * ldk + j];
    bool h = false;
    if (idx >= 0) {
      const int kid = key_id[idx];
      if (mode == 0) {
        h = (kid == g);
      } else {
        const long long want = (long long)g * n_key_ids + kid;
        int lo = 0, hi = n_pairs;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (pairs[mid] < want) lo = mid + 1; else hi = mid;
        }
        h = (lo < n_pairs && pairs[lo] == want);
        for (int s = 0; s < n_seen && h; ++s) h = (seen[s] != kid);
        if (h) seen[n_seen++] = kid;
      }
    }
    hit[j] = h;
  }
  int bits = 0;
  for (int j = 0; j < 10; ++j) bits |= hit[j] ? (1 << j) : 0;
  out[q] = bits;
}

int retrieval_hits(const int* topk, int ldk, int Q, const int* key_id, const int* gt_id, const long long* pairs, int n_pairs,
                   long long n_key_ids, int mode, int* out, cudaStream_t stream) {
  if (!topk || !key_id || !gt_id || !out || Q <= 0 || ldk < 10 || (mode != 0 && mode != 1)) return CACO_ERR_ARG;
  if (mode == 1 && (!pairs || n_pairs <= 0 || n_key_ids <= 0)) return CACO_ERR_ARG;
  retrieval_hits_kernel<<<(Q + 127) / 128, 128, 0, stream>>>(topk, ldk, Q, key_id, gt_id, pairs, n_pairs, n_key_ids, mode, out);
  count_launch();
  return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------- token pooling
// out[b, t, :] = mean_{f < group} hid[b, t*group + f, :]   for t < seq / group  ('VALID': the remainder is dropped)
__global__ void __launch_bounds__(192)
avg_pool_tokens_kernel(const float4* __restrict__ hid, int seq, int dim4, int group, int n_out, float4* __restrict__ out) {
  const int t = blockIdx.x, b = blockIdx.y;
  const float inv = 1.0f / (float)group;
  for (int c = threadIdx.x; c < dim4; c += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int f = 0; f < group; ++f) {
      const float4 v = __ldg(hid + ((size_t)b * seq + (size_t)t * group + f) * dim4 + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
    out[((size_t)b * n_out + t) * dim4 + c] = acc;
  }
}

int avg_pool_tokens(const float* hid, int batch, int seq, int dim, int group, float* out, cudaStream_t stream) {
  if (!hid || !out || batch <= 0 || seq <= 0 || dim <= 0 || (dim & 3) || group <= 0 || seq / group <= 0) return CACO_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(hid) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return CACO_ERR_ALIGN;
  const int n_out = seq / group;
  dim3 grid(n_out, batch);
  avg_pool_tokens_kernel<<<grid, 192, 0, stream>>>(reinterpret_cast<const float4*>(hid), seq, dim / 4, group, n_out,
                                                   reinterpret_cast<float4*>(out));
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace caco

extern "C" int caco_topk_rows(const float* x, int rows, int cols, int ldx, int k, int* idx_out, float* val_out, void* stream) {
  return caco::topk_rows(x, rows, cols, ldx, k, idx_out, val_out, (cudaStream_t)stream);
}
extern "C" int caco_retrieval_hits(const int* topk, int ldk, int n_queries, const int* key_id, const int* gt_id,
                                   const long long* gt_pairs, int n_pairs, long long n_key_ids, int mode, int* out,
                                   void* stream) {
  return caco::retrieval_hits(topk, ldk, n_queries, key_id, gt_id, gt_pairs, n_pairs, n_key_ids, mode, out, (cudaStream_t)stream);
}
extern "C" int caco_avg_pool_tokens(const float* hid, int batch, int seq, int dim, int group, float* out, void* stream) {
  return caco::avg_pool_tokens(hid, batch, seq, dim, group, out, (cudaStream_t)stream);
}
