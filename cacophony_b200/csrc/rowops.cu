// HBM-bound row kernels of the path: fp32->fp16 cast, LayerNorm (K4), audio position embedding,
// RoBERTa embedding + LayerNorm (K5), masked single-query attention pooling with optional fused final
// LayerNorm (K6), L2 normalisation (K7), and a small fp32 GEMM for the [batch, *] tails and the
// similarity matrix (K8).  One warp per row, 128-bit loads, warp-shuffle reductions; statistics and
// accumulation in fp32 (two-pass variance, like torch's native_layer_norm).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>

#include "caco_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace caco {

std::atomic<int64_t> g_launches{0};

constexpr int ROW_MAX_V4 = 8;  // per-lane float4 registers: dim <= 1024, dim % 128 == 0

// ------------------------------------------------------------------------------------------------ cast
__global__ void cast_f32_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(dst + i) = u;
  } else {
    for (int64_t k = i; k < n; ++k) dst[k] = __float2half_rn(src[k]);
  }
}
int cast_f32_f16(const float* src, void* dst, int64_t n, cudaStream_t stream) {
  if (!src || !dst || n <= 0) return CACO_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 7)) return CACO_ERR_ALIGN;
  const int64_t threads = (n + 3) / 4;
  cast_f32_f16_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(src, (__half*)dst, n);
  count_launch();
  return (int)cudaGetLastError();
}

// KV-cached captioning decode (engine.cu: decode_step): row b of the step's packed q | k | v GEMM output [batch, 3D] hands its
// k | v half to slot pos[b] of that sequence's cache [batch, capacity, 2D]; the first layer of a step also opens the slot in the
// key mask [batch, capacity] the step's attention launches read.  Positions live on the device so that a step can be replayed
// from a CUDA graph.
__global__ void __launch_bounds__(128)
kv_append_kernel(const __half* __restrict__ qkv, __half* __restrict__ cache, const int64_t* __restrict__ pos,
                 float* __restrict__ key_mask, int D, int capacity) {
  const int b = blockIdx.x;
  const int64_t p = pos[b];
  if (p < 0 || p >= capacity) return;
  const uint4* src = reinterpret_cast<const uint4*>(qkv + (size_t)b * 3 * D + D);
  uint4* dst = reinterpret_cast<uint4*>(cache + ((size_t)b * capacity + (size_t)p) * 2 * D);
  for (int i = threadIdx.x; i < 2 * D / 8; i += blockDim.x) dst[i] = src[i];
  if (key_mask && threadIdx.x == 0) key_mask[(size_t)b * capacity + p] = 1.0f;
}
int kv_append(const void* qkv, void* cache, const int64_t* pos, float* key_mask, int batch, int D, int capacity,
              cudaStream_t stream) {
  if (!qkv || !cache || !pos || batch <= 0 || D <= 0 || (D & 7) || capacity <= 0) return CACO_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(cache) & 15)) return CACO_ERR_ALIGN;
  kv_append_kernel<<<batch, 128, 0, stream>>>((const __half*)qkv, (__half*)cache, pos, key_mask, D, capacity);
  count_launch();
  return (int)cudaGetLastError();
}

// split-weight packing: dst[r, 0:K] = fp16(w), dst[r, K:2K] = fp16(w - fp16(w))   (K % 4 == 0)
__global__ void cast_f32_f16_split_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t rows, int64_t K) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= rows * K) return;
  const int64_t r = i / K, k = i - r * K;
  const float4 v = *reinterpret_cast<const float4*>(src + i);
  const float f[4] = {v.x, v.y, v.z, v.w};
  __half hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    hi[j] = __float2half_rn(f[j]);
    lo[j] = __float2half_rn(f[j] - __half2float(hi[j]));
  }
  *reinterpret_cast<uint2*>(dst + r * 2 * K + k) = *reinterpret_cast<const uint2*>(hi);
  *reinterpret_cast<uint2*>(dst + r * 2 * K + K + k) = *reinterpret_cast<const uint2*>(lo);
}
int cast_f32_f16_split(const float* src, void* dst, int64_t rows, int64_t K, cudaStream_t stream) {
  if (!src || !dst || rows <= 0 || K <= 0 || (K & 3)) return CACO_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 7)) return CACO_ERR_ALIGN;
  const int64_t threads = rows * K / 4;
  cast_f32_f16_split_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(src, (__half*)dst, rows, K);
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ row helpers
struct RowStats { float mean, rstd; };

// v: this lane's float4 chunks of one row (chunk c covers columns c*128 + lane*4 .. +3)
__device__ __forceinline__ RowStats row_stats(const float4 (&v)[ROW_MAX_V4], int nv, int dim, float eps) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < ROW_MAX_V4; ++c)
    if (c < nv) s += (v[c].x + v[c].y) + (v[c].z + v[c].w);
  const float mean = warp_sum(s) / (float)dim;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < ROW_MAX_V4; ++c)
    if (c < nv) {
      const float a = v[c].x - mean, b = v[c].y - mean, d = v[c].z - mean, e = v[c].w - mean;
      q += (a * a + b * b) + (d * d + e * e);
    }
  const float var = warp_sum(q) / (float)dim;
  RowStats r;
  r.mean = mean;
  r.rstd = rsqrtf(var + eps);
  return r;
}

__device__ __forceinline__ void store_row(const float4& y, float* o32, __half* o16, size_t off) {
  if (o32) *reinterpret_cast<float4*>(o32 + off) = y;
  if (o16) {
    __half2 a = __floats2half2_rn(y.x, y.y), b = __floats2half2_rn(y.z, y.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(o16 + off) = u;
  }
}

// ------------------------------------------------------------------------------------------------ K4 LayerNorm
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                 float* __restrict__ o32, __half* __restrict__ o16, int rows, int dim) {
  pdl_wait();
  pdl_launch();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nv = dim / 128;
  float4 v[ROW_MAX_V4];
  const float* xr = x + (size_t)row * dim;
#pragma unroll
  for (int c = 0; c < ROW_MAX_V4; ++c)
    if (c < nv) v[c] = *reinterpret_cast<const float4*>(xr + c * 128 + lane * 4);
  const RowStats st = row_stats(v, nv, dim, eps);
#pragma unroll
  for (int c = 0; c < ROW_MAX_V4; ++c)
    if (c < nv) {
      const int col = c * 128 + lane * 4;
      const float4 g = *reinterpret_cast<const float4*>(gamma + col);
      const float4 bt = *reinterpret_cast<const float4*>(beta + col);
      float4 y;
      y.x = (v[c].x - st.mean) * st.rstd * g.x + bt.x;
      y.y = (v[c].y - st.mean) * st.rstd * g.y + bt.y;
      y.z = (v[c].z - st.mean) * st.rstd * g.z + bt.z;
      y.w = (v[c].w - st.mean) * st.rstd * g.w + bt.w;
      store_row(y, o32, o16, (size_t)row * dim + col);
    }
}
int layernorm(const float* x, const float* gamma, const float* beta, float eps, float* out_f32, void* out_f16, int rows,
              int dim, cudaStream_t stream) {
  if (!x || !gamma || !beta || rows <= 0 || dim <= 0 || (dim % 128) || dim > 128 * ROW_MAX_V4) return CACO_ERR_ARG;
  if (!out_f32 && !out_f16) return CACO_ERR_ARG;
  cudaError_t le = launch_pdl(layernorm_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, x, gamma, beta, eps, out_f32, (__half*)out_f16, rows, dim);
  if (le != cudaSuccess) return (int)le;
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ audio pos-emb
// mae.py:102-109: angle = t * exp(2i * (-ln 1e4) / dim); x += cat[sin, cos];  mae.py:136-142: x += freq_emb[f]
__global__ void __launch_bounds__(256)
audio_add_pos_kernel(float* __restrict__ x, const float* __restrict__ time_inds, const float* __restrict__ freq_inds,
                     const float* __restrict__ freq_emb, int n_freq, int rows, int dim, int init) {
  // one warp per APOS_ROWS consecutive rows: the frequencies w_i = exp(-2 i ln(1e4) / dim) of the lane's columns are
  // computed once and reused.  init != 0: x = sincos + freq_emb (pure write: the input projection then accumulates onto it
  // in place); init == 0: x += sincos + freq_emb.
  constexpr int APOS_ROWS = 4, APOS_MAXI = 16;             // dim / 2 <= 32 * APOS_MAXI
  const int lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * APOS_ROWS;
  const int half = dim / 2;
  float w[APOS_MAXI];
#pragma unroll
  for (int k = 0; k < APOS_MAXI; ++k) {
    const int i = lane + 32 * k;
    w[k] = expf(((2.0f * (float)i) * -9.210340371976184f) / (float)dim);
  }
  for (int r = 0; r < APOS_ROWS; ++r) {
    const int row = row0 + r;
    if (row >= rows) return;
    const float t = time_inds[row];
    int f = (int)freq_inds[row];   // .long() truncation
    f = min(max(f, 0), n_freq - 1);
    float* xr = x + (size_t)row * dim;
    const float* fe = freq_emb + (size_t)f * dim;
#pragma unroll
    for (int k = 0; k < APOS_MAXI; ++k) {
      const int i = lane + 32 * k;
      if (i < half) {
        float sn, cs;
        sincosf(t * w[k], &sn, &cs);
        if (init) {
          xr[i] = sn + fe[i];
          xr[i + half] = cs + fe[i + half];
        } else {
          xr[i] = (xr[i] + sn) + fe[i];
          xr[i + half] = (xr[i + half] + cs) + fe[i + half];
        }
      }
    }
  }
}
int audio_add_pos(float* x, const float* time_inds, const float* freq_inds, const float* freq_emb, int n_freq, int rows,
                  int dim, int init, cudaStream_t stream) {
  if (!x || !time_inds || !freq_inds || !freq_emb || rows <= 0 || dim <= 0 || (dim & 1) || dim > 1024) return CACO_ERR_ARG;
  audio_add_pos_kernel<<<(rows + 31) / 32, 256, 0, stream>>>(x, time_inds, freq_inds, freq_emb, n_freq, rows, dim, init);
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ K5 text embed + LN
__global__ void __launch_bounds__(256)
text_embed_ln_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ pids, const float* __restrict__ word,
                     const float* __restrict__ pos, const float* __restrict__ type0, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, float* __restrict__ o32, __half* __restrict__ o16,
                     int rows, int T, int dim, int vocab, int max_pos) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  int64_t id = ids[row];
  int64_t pid = pids ? pids[row] : (int64_t)(row % T);   // roberta.py:292-293: arange(T)
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  pid = pid < 0 ? 0 : (pid >= max_pos ? max_pos - 1 : pid);
  const int nv = dim / 128;
  float4 v[ROW_MAX_V4];
#pragma unroll
  for (int c = 0; c < ROW_MAX_V4; ++c)
    if (c < nv) {
      const int col = c * 128 + lane * 4;
      const float4 a = *reinterpret_cast<const float4*>(word + (size_t)id * dim + col);
      const float4 p = *reinterpret_cast<const float4*>(pos + (size_t)pid * dim + col);
      const float4 ty = *reinterpret_cast<const float4*>(type0 + col);
      v[c] = make_float4((a.x + p.x) + ty.x, (a.y + p.y) + ty.y, (a.z + p.z) + ty.z, (a.w + p.w) + ty.w);
    }
  const RowStats st = row_stats(v, nv, dim, eps);
#pragma unroll
  for (int c = 0; c < ROW_MAX_V4; ++c)
    if (c < nv) {
      const int col = c * 128 + lane * 4;
      const float4 g = *reinterpret_cast<const float4*>(gamma + col);
      const float4 bt = *reinterpret_cast<const float4*>(beta + col);
      float4 y;
      y.x = (v[c].x - st.mean) * st.rstd * g.x + bt.x;
      y.y = (v[c].y - st.mean) * st.rstd * g.y + bt.y;
      y.z = (v[c].z - st.mean) * st.rstd * g.z + bt.z;
      y.w = (v[c].w - st.mean) * st.rstd * g.w + bt.w;
      store_row(y, o32, o16, (size_t)row * dim + col);
    }
}
int text_embed_ln(const int64_t* ids, const int64_t* position_ids, const float* word, const float* pos,
                  const float* type0, const float* gamma, const float* beta, float eps, float* out_f32, void* out_f16,
                  int batch, int T, int dim, int vocab, int max_pos, cudaStream_t stream) {
  if (!ids || !word || !pos || !type0 || !gamma || !beta || batch <= 0 || T <= 0) return CACO_ERR_ARG;
  if ((dim % 128) || dim > 128 * ROW_MAX_V4 || (!out_f32 && !out_f16)) return CACO_ERR_ARG;
  const int rows = batch * T;
  text_embed_ln_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(ids, position_ids, word, pos, type0, gamma, beta, eps, out_f32,
                                                           (__half*)out_f16, rows, T, dim, vocab, max_pos);
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ K6 attention pool
// One CTA (512 threads) per sample, ONE pass over the sample's rows (the two-pass form read every row twice from DRAM:
// 787 MB per audio call, L2 hit rate 0.3 % — profiles/r01_notes.md).  Each warp walks its rows with an online softmax per
// head: optional LayerNorm of the row, score against the folded query u[h], running maximum m, running sum l and a running
// weighted row sum acc[h] held in registers (dim / 32 floats per lane and head), rescaled when the maximum grows.  The 16
// warps' partial (m, l, acc) are merged through shared memory in a fixed order (deterministic).
constexpr int POOL_MAX_HEADS = 4;
constexpr int POOL_THREADS = 512, POOL_WARPS = POOL_THREADS / 32;
template <int HEADS>
__global__ void __launch_bounds__(POOL_THREADS)
attn_pool_kernel(const float* __restrict__ hid, const float* __restrict__ mask, const float* __restrict__ u,
                 const float* __restrict__ cvec, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float eps,
                 float* __restrict__ hid_out, float* __restrict__ pooled, int S, int dim, int heads_total) {
  // u / cvec / pooled already point at this launch's first head; heads_total = row stride of pooled in heads
  extern __shared__ float sm[];
  float* s_part = sm;                                   // [warps][HEADS][dim]
  __shared__ float s_m[POOL_WARPS][POOL_MAX_HEADS], s_l[POOL_WARPS][POOL_MAX_HEADS];
  __shared__ float s_scale[POOL_WARPS][POOL_MAX_HEADS], s_inv[POOL_MAX_HEADS];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* hb = hid + (size_t)b * S * dim;
  const float* mb = mask + (size_t)b * S;
  const int nv = dim / 128;
  const bool do_ln = ln_g != nullptr;

  float m[HEADS], l[HEADS];
  float4 acc[HEADS][ROW_MAX_V4];
#pragma unroll
  for (int h = 0; h < HEADS; ++h) {
    m[h] = -INFINITY;
    l[h] = 0.f;
#pragma unroll
    for (int c = 0; c < ROW_MAX_V4; ++c) acc[h][c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int j = warp; j < S; j += POOL_WARPS) {
    const bool live = mb[j] != 0.0f;
    if (!live && !(do_ln && hid_out)) continue;          // a masked row only matters if its LayerNorm output was asked for
    float4 v[ROW_MAX_V4];
#pragma unroll
    for (int c = 0; c < ROW_MAX_V4; ++c)
      if (c < nv) v[c] = *reinterpret_cast<const float4*>(hb + (size_t)j * dim + c * 128 + lane * 4);
    if (do_ln) {
      const RowStats st = row_stats(v, nv, dim, eps);
#pragma unroll
      for (int c = 0; c < ROW_MAX_V4; ++c)
        if (c < nv) {
          const int col = c * 128 + lane * 4;
          const float4 g = *reinterpret_cast<const float4*>(ln_g + col);
          const float4 bt = *reinterpret_cast<const float4*>(ln_b + col);
          v[c].x = (v[c].x - st.mean) * st.rstd * g.x + bt.x;
          v[c].y = (v[c].y - st.mean) * st.rstd * g.y + bt.y;
          v[c].z = (v[c].z - st.mean) * st.rstd * g.z + bt.z;
          v[c].w = (v[c].w - st.mean) * st.rstd * g.w + bt.w;
          if (hid_out) *reinterpret_cast<float4*>(hid_out + ((size_t)b * S + j) * dim + col) = v[c];
        }
    }
    if (!live) continue;
#pragma unroll
    for (int h = 0; h < HEADS; ++h) {
      float d = 0.f;
#pragma unroll
      for (int c = 0; c < ROW_MAX_V4; ++c)
        if (c < nv) {
          const float4 uu = *reinterpret_cast<const float4*>(u + (size_t)h * dim + c * 128 + lane * 4);
          d += (v[c].x * uu.x + v[c].y * uu.y) + (v[c].z * uu.z + v[c].w * uu.w);
        }
      d = warp_sum(d) + cvec[h];
      const float m_new = fmaxf(m[h], d);
      const float keep = expf(m[h] - m_new);              // first live row: exp(-inf) = 0
      const float p = expf(d - m_new);
      l[h] = l[h] * keep + p;
      m[h] = m_new;
#pragma unroll
      for (int c = 0; c < ROW_MAX_V4; ++c)
        if (c < nv) {
          acc[h][c].x = fmaf(p, v[c].x, acc[h][c].x * keep);
          acc[h][c].y = fmaf(p, v[c].y, acc[h][c].y * keep);
          acc[h][c].z = fmaf(p, v[c].z, acc[h][c].z * keep);
          acc[h][c].w = fmaf(p, v[c].w, acc[h][c].w * keep);
        }
    }
  }
  // ---- merge the warps' partials
#pragma unroll
  for (int h = 0; h < HEADS; ++h) {
    if (lane == 0) { s_m[warp][h] = m[h]; s_l[warp][h] = l[h]; }
#pragma unroll
    for (int c = 0; c < ROW_MAX_V4; ++c)
      if (c < nv) *reinterpret_cast<float4*>(s_part + ((size_t)warp * HEADS + h) * dim + c * 128 + lane * 4) = acc[h][c];
  }
  __syncthreads();
  if (tid < HEADS) {
    float M = -INFINITY;
    for (int w = 0; w < POOL_WARPS; ++w) M = fmaxf(M, s_m[w][tid]);
    float L = 0.f;
    for (int w = 0; w < POOL_WARPS; ++w) {
      const float sc = (s_m[w][tid] == -INFINITY) ? 0.f : expf(s_m[w][tid] - M);     // a warp without live rows adds nothing
      s_scale[w][tid] = sc;
      L += s_l[w][tid] * sc;
    }
    s_inv[tid] = 1.0f / L;                                 // no live row at all: L = 0 -> NaN row, like torch.softmax
  }
  __syncthreads();
  for (int col = tid; col < dim; col += POOL_THREADS) {
#pragma unroll
    for (int h = 0; h < HEADS; ++h) {
      float a = 0.f;
      for (int w = 0; w < POOL_WARPS; ++w) a = fmaf(s_part[((size_t)w * HEADS + h) * dim + col], s_scale[w][h], a);
      pooled[((size_t)b * heads_total + h) * dim + col] = a * s_inv[h];
    }
  }
}

template <int HEADS>
static int launch_pool(const float* hid, const float* mask, const float* u, const float* c, const float* ln_gamma,
                       const float* ln_beta, float ln_eps, float* hid_out, float* pooled, int batch, int seq, int dim,
                       int heads_total, cudaStream_t stream) {
  const size_t smem = (size_t)POOL_WARPS * HEADS * dim * sizeof(float);
  if (smem > 200 * 1024) return CACO_ERR_ARG;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(attn_pool_kernel<HEADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e) return (int)e;
    attr_once.done();
  }
  attn_pool_kernel<HEADS><<<batch, POOL_THREADS, smem, stream>>>(hid, mask, u, c, ln_gamma, ln_beta, ln_eps, hid_out, pooled, seq, dim, heads_total);
  count_launch();
  return (int)cudaGetLastError();
}

int attn_pool(const float* hid, const float* mask, const float* u, const float* c, const float* ln_gamma,
              const float* ln_beta, float ln_eps, float* hid_out, float* pooled, int batch, int seq, int heads, int dim,
              cudaStream_t stream) {
  if (!hid || !mask || !u || !c || !pooled || batch <= 0 || seq <= 0 || heads <= 0 || heads > 16) return CACO_ERR_ARG;
  if ((dim % 128) || dim > 128 * ROW_MAX_V4) return CACO_ERR_ARG;
  // up to POOL_MAX_HEADS heads share one pass over the rows (their running sums live in registers); more heads (the JAX
  // configuration's 8, caco/load_model.py:47) take one pass per group of four.  The LayerNorm-ed rows are written once.
  for (int h0 = 0; h0 < heads; h0 += POOL_MAX_HEADS) {
    const int nh = heads - h0 < POOL_MAX_HEADS ? heads - h0 : POOL_MAX_HEADS;
    const float* uu = u + (size_t)h0 * dim;
    float* pp = pooled + (size_t)h0 * dim;
    float* ho = h0 == 0 ? hid_out : nullptr;
    int rc;
    switch (nh) {
      case 1: rc = launch_pool<1>(hid, mask, uu, c + h0, ln_gamma, ln_beta, ln_eps, ho, pp, batch, seq, dim, heads, stream); break;
      case 2: rc = launch_pool<2>(hid, mask, uu, c + h0, ln_gamma, ln_beta, ln_eps, ho, pp, batch, seq, dim, heads, stream); break;
      case 3: rc = launch_pool<3>(hid, mask, uu, c + h0, ln_gamma, ln_beta, ln_eps, ho, pp, batch, seq, dim, heads, stream); break;
      default: rc = launch_pool<4>(hid, mask, uu, c + h0, ln_gamma, ln_beta, ln_eps, ho, pp, batch, seq, dim, heads, stream); break;
    }
    if (rc) return rc;
  }
  return 0;
}

// fold a pooler's key projection into its (fixed) query: u[h,i] = sum_d qs[h,d] Wk[h*dh+d, i], c[h] = sum_d qs[h,d] bk[h*dh+d]
// with qs = query * qscale (caco.py:61-62: 1/sqrt(dh); roberta.py:259: 1/sqrt(hidden)).
__global__ void fold_query_kernel(const float* __restrict__ query, const float* __restrict__ wk, const float* __restrict__ bk,
                                  float qscale, float* __restrict__ u, float* __restrict__ c, int heads, int dh, int dim) {
  const int h = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < dim) {
    float acc = 0.f;
    for (int d = 0; d < dh; ++d) acc = fmaf(query[h * dh + d] * qscale, wk[(size_t)(h * dh + d) * dim + i], acc);
    u[(size_t)h * dim + i] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float acc = 0.f;
    for (int d = 0; d < dh; ++d) acc = fmaf(query[h * dh + d] * qscale, bk[h * dh + d], acc);
    c[h] = acc;
  }
}
int fold_query(const float* query, const float* wk, const float* bk, float qscale, float* u, float* c, int heads, int dh,
               int dim, cudaStream_t stream) {
  dim3 grid((dim + 127) / 128, heads);
  fold_query_kernel<<<grid, 128, 0, stream>>>(query, wk, bk, qscale, u, c, heads, dh, dim);
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ small fp32 GEMM
// out[M,N] = alpha * A[M,K] · W[N,K]^T + bias for the [batch, 768]-sized tails.  32x32 tile, 64 threads, 4x4 micro-tile,
// K step 32, next K slab prefetched into registers while the current one is multiplied: at M = 256 the grid is
// (N/32) x 8 CTAs, so even the 256x384 value projection fills the 148 SMs.  alpha = alpha_scalar * (log_alpha ?
// exp(*log_alpha) : 1).  K % 4 == 0 (float4 loads), edges handled by zero fill.
__global__ void __launch_bounds__(64)
sgemm_nt_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw, const float* __restrict__ bias,
                float alpha, const float* __restrict__ log_alpha, float* __restrict__ out, int ldo, int M, int N, int K) {
  __shared__ float sA[32][32 + 4];   // [k][m]
  __shared__ float sW[32][32 + 4];   // [k][n]
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = tid & 7, ty = tid >> 3;        // 8 x 8 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // loader mapping: 32 rows x 8 float4 per slab = 256 float4 per operand, 4 per thread
  float4 ra[4], rw[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + q * 64, r = idx >> 3, kq = (idx & 7) * 4;
      ra[q] = (m0 + r < M && k0 + kq < K) ? *reinterpret_cast<const float4*>(A + (size_t)(m0 + r) * lda + k0 + kq)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
      rw[q] = (n0 + r < N && k0 + kq < K) ? *reinterpret_cast<const float4*>(W + (size_t)(n0 + r) * ldw + k0 + kq)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + q * 64, r = idx >> 3, kq = (idx & 7) * 4;
      sA[kq][r] = ra[q].x; sA[kq + 1][r] = ra[q].y; sA[kq + 2][r] = ra[q].z; sA[kq + 3][r] = ra[q].w;
      sW[kq][r] = rw[q].x; sW[kq + 1][r] = rw[q].y; sW[kq + 2][r] = rw[q].z; sW[kq + 3][r] = rw[q].w;
    }
    __syncthreads();
    if (k0 + 32 < K) fetch(k0 + 32);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&sA[k][ty * 4]);
      const float4 wv = *reinterpret_cast<const float4*>(&sW[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w};
      const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float al = alpha * (log_alpha ? expf(*log_alpha) : 1.0f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) out[(size_t)m * ldo + n] = al * acc[i][j] + (bias ? bias[n] : 0.f);
    }
  }
}
static int sgemm_impl(const float* A, int lda, const float* W, int ldw, const float* bias, float alpha,
                      const float* log_alpha, float* out, int ldo, int M, int N, int K, cudaStream_t stream) {
  if (!A || !W || !out || M <= 0 || N <= 0 || K <= 0) return CACO_ERR_ARG;
  if ((lda & 3) || (ldw & 3) || (K & 3) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
    return CACO_ERR_ALIGN;
  dim3 grid((N + 31) / 32, (M + 31) / 32);
  sgemm_nt_kernel<<<grid, 64, 0, stream>>>(A, lda, W, ldw, bias, alpha, log_alpha, out, ldo, M, N, K);
  count_launch();
  return (int)cudaGetLastError();
}
int sgemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias, float alpha, float* out, int ldo,
             int M, int N, int K, cudaStream_t stream) {
  return sgemm_impl(A, lda, W, ldw, bias, alpha, nullptr, out, ldo, M, N, K, stream);
}

// ------------------------------------------------------------------------------------------------ K7 / K8
__global__ void __launch_bounds__(256)
l2norm_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int dim, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = in + (size_t)row * dim;
  float s = 0.f;
  for (int i = lane; i < dim; i += 32) { const float v = x[i] + eps; s = fmaf(v, v, s); }   // caco.py:146: ||e + 1e-10||
  s = warp_sum(s);
  const float nrm = sqrtf(s);
  for (int i = lane; i < dim; i += 32) out[(size_t)row * dim + i] = x[i] / nrm;
}
int l2norm(const float* in, float* out, int rows, int dim, float eps, cudaStream_t stream) {
  if (!in || !out || rows <= 0 || dim <= 0) return CACO_ERR_ARG;
  l2norm_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(in, out, rows, dim, eps);
  count_launch();
  return (int)cudaGetLastError();
}
int sim_logits(const float* a, const float* t, const float* logit_scale, float* at, float* ta, int na, int nt, int dim,
               cudaStream_t stream) {
  if (!a || !t || !logit_scale || !at) return CACO_ERR_ARG;
  int rc = sgemm_impl(a, dim, t, dim, nullptr, 1.0f, logit_scale, at, nt, na, nt, dim, stream);
  if (rc || !ta) return rc;
  return sgemm_impl(t, dim, a, dim, nullptr, 1.0f, logit_scale, ta, na, nt, na, dim, stream);
}

}  // namespace caco

extern "C" {
int caco_version(void) { return 100; }
int caco_built_arch(void) { return 100; }
int64_t caco_launch_count(void) { return caco::g_launches.load(); }
int caco_cast_f32_f16(const float* src, void* dst, int64_t n, void* stream) { return caco::cast_f32_f16(src, dst, n, (cudaStream_t)stream); }
int caco_cast_f32_f16_split(const float* src, void* dst, int64_t rows, int64_t K, void* stream) {
  return caco::cast_f32_f16_split(src, dst, rows, K, (cudaStream_t)stream);
}
int caco_layernorm(const float* x, const float* gamma, const float* beta, float eps, float* out_f32, void* out_f16, int rows,
                   int dim, void* stream) {
  return caco::layernorm(x, gamma, beta, eps, out_f32, out_f16, rows, dim, (cudaStream_t)stream);
}
int caco_audio_add_pos(float* x, const float* time_inds, const float* freq_inds, const float* freq_emb, int n_freq, int rows,
                       int dim, void* stream) {
  return caco::audio_add_pos(x, time_inds, freq_inds, freq_emb, n_freq, rows, dim, 0, (cudaStream_t)stream);
}
int caco_text_embed_ln(const int64_t* ids, const int64_t* position_ids, const float* word, const float* pos,
                       const float* type0, const float* gamma, const float* beta, float eps, float* out_f32, void* out_f16,
                       int batch, int T, int dim, int vocab, int max_pos, void* stream) {
  return caco::text_embed_ln(ids, position_ids, word, pos, type0, gamma, beta, eps, out_f32, out_f16, batch, T, dim, vocab,
                             max_pos, (cudaStream_t)stream);
}
int caco_attn_pool(const float* hid, const float* mask, const float* u, const float* c, const float* ln_gamma,
                   const float* ln_beta, float ln_eps, float* hid_out, float* pooled, int batch, int seq, int heads, int dim,
                   void* stream) {
  return caco::attn_pool(hid, mask, u, c, ln_gamma, ln_beta, ln_eps, hid_out, pooled, batch, seq, heads, dim,
                         (cudaStream_t)stream);
}
int caco_sgemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias, float alpha, float* out, int ldo,
                  int M, int N, int K, void* stream) {
  return caco::sgemm_nt(A, lda, W, ldw, bias, alpha, out, ldo, M, N, K, (cudaStream_t)stream);
}
int caco_l2norm(const float* in, float* out, int rows, int dim, float eps, void* stream) {
  return caco::l2norm(in, out, rows, dim, eps, (cudaStream_t)stream);
}
int caco_sim_logits(const float* a, const float* t, const float* logit_scale, float* at, float* ta, int na, int nt, int dim,
                    void* stream) {
  return caco::sim_logits(a, t, logit_scale, at, ta, na, nt, dim, (cudaStream_t)stream);
}
}
