// K3 — self-attention for both towers.
//
// attention_audio: nn.MultiheadAttention semantics (mae.py:69-74,89-92 -> aten::_native_multi_head_attention):
//   packed in-proj output qkv[B*S, 3*H*dh] (q | k | v), q scaled by 1/sqrt(dh), keys with mask == 0 -> -inf,
//   fp32 softmax over keys, P·V, heads concatenated.  Flash-style: one CTA = 64 queries of one (clip, head),
//   K/V streamed in 64-key tiles through a cp.async double buffer, scores never leave the SM
//   (materialised they would be H*S^2*4 = 8 MB per clip per layer).
// attention_text: causal + key-padding attention of the RoBERTa tower (roberta.py:86-102, mask :297-310), 12 heads x 64:
//   the warp-MMA flash kernel with a causal predicate (key tiles past the diagonal are skipped).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>

#include "caco_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace caco {

// ------------------------------------------------------------------------------------------------
// small PTX helpers (legacy warp-level tensor path: fine for 10 % of the FLOPs in v1)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 -> 16 bytes of zeros
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int AT_BM = 64;   // queries per CTA
constexpr int AT_BN = 64;   // keys per tile

template <int DH>
struct AttnCfg {
  static constexpr int LD = DH + 8;                       // padded row (halfs): conflict-free ldmatrix
  static constexpr int TILE_BYTES = AT_BN * LD * 2;
  static constexpr int SMEM_BYTES = 5 * TILE_BYTES + 2 * AT_BN * 4;  // Q + 2xK + 2xV + mask bias
};

template <int DH>
__device__ __forceinline__ void load_tile(uint32_t dst, const __half* src, int row0, int n_rows, int ld_src, int tid) {
  // 64 rows x DH halfs, 16-byte chunks; rows >= n_rows are zero-filled
  constexpr int CH = DH / 8;
  for (int i = tid; i < 64 * CH; i += 128) {
    const int r = i / CH, c = i % CH;
    const int row = row0 + r;
    const bool ok = row < n_rows;
    const __half* p = src + (size_t)(ok ? row : 0) * ld_src + c * 8;
    cp_async16(dst + (r * AttnCfg<DH>::LD + c * 8) * 2, p, ok);
  }
}

// q / k / v may come from different buffers and sequences (cross-attention of the captioning decoder, roberta.py:76-83,
// 205-211: queries = text tokens, keys / values = audio tokens): Sq query rows with row pitch ldq, Skv key rows with pitch ldkv;
// self-attention passes the three column blocks of one packed qkv buffer.  mask is per KEY ([batch, Skv], 1 = keep).
template <int DH, bool CAUSAL = false>
__global__ void __launch_bounds__(128)
attention_audio_kernel(const __half* __restrict__ q_ptr, int ldq, const __half* __restrict__ k_ptr, const __half* __restrict__ v_ptr,
                       int ldkv, const float* __restrict__ mask, __half* __restrict__ out, int Sq, int S, int H, float scale_log2) {
  using Cfg = AttnCfg<DH>;
  extern __shared__ __align__(16) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + Cfg::TILE_BYTES;
  const uint32_t sV = sK + 2 * Cfg::TILE_BYTES;
  float* sBias = reinterpret_cast<float*>(smem + 5 * Cfg::TILE_BYTES);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * AT_BM, h = blockIdx.y, b = blockIdx.z;
  const int D = H * DH;
  const int ld = ldkv;
  const __half* gQ = q_ptr + (size_t)b * Sq * ldq + h * DH;
  const __half* gK = k_ptr + (size_t)b * S * ldkv + h * DH;
  const __half* gV = v_ptr + (size_t)b * S * ldkv + h * DH;
  const float* gmask = mask + (size_t)b * S;

  // causal (text tower, roberta.py:297-310): key tiles past this query tile's last row are never needed
  const int n_tiles = CAUSAL ? min((S + AT_BN - 1) / AT_BN, (q0 + AT_BM - 1) / AT_BN + 1) : (S + AT_BN - 1) / AT_BN;
  load_tile<DH>(sQ, gQ, q0, Sq, ldq, tid);
  load_tile<DH>(sK, gK, 0, S, ld, tid);
  load_tile<DH>(sV, gV, 0, S, ld, tid);
  if (tid < AT_BN) sBias[tid] = (tid < S && gmask[tid] != 0.0f) ? 0.0f : -INFINITY;
  cp_async_commit();

  uint32_t qf[DH / 16][4];
  float o[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) {
      const int nb = buf ^ 1;
      load_tile<DH>(sK + nb * Cfg::TILE_BYTES, gK, (t + 1) * AT_BN, S, ld, tid);
      load_tile<DH>(sV + nb * Cfg::TILE_BYTES, gV, (t + 1) * AT_BN, S, ld, tid);
      if (tid < AT_BN) {
        const int j = (t + 1) * AT_BN + tid;
        sBias[nb * AT_BN + tid] = (j < S && gmask[j] != 0.0f) ? 0.0f : -INFINITY;
      }
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        const int r = warp * 16 + (lane & 15), c = kk * 16 + (lane >> 4) * 8;
        ldsm_x4(sQ + (r * Cfg::LD + c) * 2, qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
      }
    }
    const uint32_t kt = sK + buf * Cfg::TILE_BYTES;
    const uint32_t vt = sV + buf * Cfg::TILE_BYTES;
    const float* bias = sBias + buf * AT_BN;

    // ---- S = Q K^T (16 x 64 per warp)
    float s[AT_BN / 8][4];
#pragma unroll
    for (int i = 0; i < AT_BN / 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < AT_BN / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        const int r = np * 16 + (lane >> 4) * 8 + (lane & 7), c = kk * 16 + ((lane >> 3) & 1) * 8;
        ldsm_x4(kt + (r * Cfg::LD + c) * 2, b0, b1, b2, b3);
        mma_16816(s[2 * np], qf[kk], b0, b1);
        mma_16816(s[2 * np + 1], qf[kk], b2, b3);
      }
    }
    // ---- scale, mask, online softmax (base-2 domain)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < AT_BN / 8; ++i) {
      const int c = i * 8 + (lane & 3) * 2;
      const float b0 = bias[c], b1 = bias[c + 1];
      s[i][0] = s[i][0] * scale_log2 + b0;
      s[i][1] = s[i][1] * scale_log2 + b1;
      s[i][2] = s[i][2] * scale_log2 + b0;
      s[i][3] = s[i][3] * scale_log2 + b1;
      if (CAUSAL) {                     // allowed(i, j) = j <= i  (on top of the key-padding bias)
        const int col = t * AT_BN + c, r0 = q0 + warp * 16 + (lane >> 2);
        if (col > r0) s[i][0] = -INFINITY;
        if (col + 1 > r0) s[i][1] = -INFINITY;
        if (col > r0 + 8) s[i][2] = -INFINITY;
        if (col + 1 > r0 + 8) s[i][3] = -INFINITY;
      }
      mx[0] = fmaxf(mx[0], fmaxf(s[i][0], s[i][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[i][2], s[i][3]));
    }
    float corr[2], m_use[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;   // fully masked so far: keep exp2(-inf - 0) = 0
      corr[r] = exp2f(m_run[r] - m_use[r]);
      m_run[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[AT_BN / 16][4];
#pragma unroll
    for (int i = 0; i < AT_BN / 8; ++i) {
      s[i][0] = exp2f(s[i][0] - m_use[0]);
      s[i][1] = exp2f(s[i][1] - m_use[0]);
      s[i][2] = exp2f(s[i][2] - m_use[1]);
      s[i][3] = exp2f(s[i][3] - m_use[1]);
      rs[0] += s[i][0] + s[i][1];
      rs[1] += s[i][2] + s[i][3];
    }
#pragma unroll
    for (int kk = 0; kk < AT_BN / 16; ++kk) {
      pf[kk][0] = pack_h2(s[2 * kk][0], s[2 * kk][1]);
      pf[kk][1] = pack_h2(s[2 * kk][2], s[2 * kk][3]);
      pf[kk][2] = pack_h2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[kk][3] = pack_h2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_run[r] = l_run[r] * corr[r] + rs[r];
    }
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0];
      o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    // ---- O += P V
#pragma unroll
    for (int kk = 0; kk < AT_BN / 16; ++kk) {
#pragma unroll
      for (int np = 0; np < DH / 16; ++np) {
        uint32_t b0, b1, b2, b3;
        const int r = kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7), c = np * 16 + (lane >> 4) * 8;
        ldsm_x4_t(vt + (r * Cfg::LD + c) * 2, b0, b1, b2, b3);
        mma_16816(o[2 * np], pf[kk], b0, b1);
        mma_16816(o[2 * np + 1], pf[kk], b2, b3);
      }
    }
    __syncthreads();  // everyone done with this K/V buffer before it is refilled
  }

  // ---- normalise, stage through smem (Q tile is dead), coalesced 16-byte stores
  const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
  __half* sO = reinterpret_cast<__half*>(smem);
  {
    const int r0 = warp * 16 + (lane >> 2);
#pragma unroll
    for (int i = 0; i < DH / 8; ++i) {
      const int c = i * 8 + (lane & 3) * 2;
      *reinterpret_cast<uint32_t*>(sO + r0 * Cfg::LD + c) = pack_h2(o[i][0] * inv0, o[i][1] * inv0);
      *reinterpret_cast<uint32_t*>(sO + (r0 + 8) * Cfg::LD + c) = pack_h2(o[i][2] * inv1, o[i][3] * inv1);
    }
  }
  __syncthreads();
  constexpr int CH = DH / 8;
  __half* gO = out + (size_t)b * Sq * D + h * DH;
  for (int i = tid; i < AT_BM * CH; i += 128) {
    const int r = i / CH, c = i % CH;
    if (q0 + r < Sq) *reinterpret_cast<uint4*>(gO + (size_t)(q0 + r) * D + c * 8) = *reinterpret_cast<const uint4*>(sO + r * Cfg::LD + c * 8);
  }
}

int attention_audio_pp(const void* qkv, const float* mask, void* out, int batch, int seq, int heads, int dh,
                       cudaStream_t stream);

int attention_audio(const void* qkv, const float* mask, void* out, int batch, int seq, int heads, int dh,
                    cudaStream_t stream) {
  if (!qkv || !mask || !out || batch <= 0 || seq <= 0 || heads <= 0) return CACO_ERR_ARG;
  // head_dim 96 (the checkpoint's audio tower): the persistent ping-pong tcgen05 kernel, up to 4096 keys (its per-item key
  // bias lives in shared memory); anything else (head_dim 64 configurations, longer sequences) takes the warp-level
  // flash kernel below.
  if (dh == 96 && seq <= 4096) return attention_audio_pp(qkv, mask, out, batch, seq, heads, dh, stream);
  if ((heads * dh) % 8) return CACO_ERR_ARG;
  const float scale_log2 = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
  dim3 grid((seq + AT_BM - 1) / AT_BM, heads, batch);
  cudaError_t e;
  if (dh == 96) {
    static PerDeviceOnce set96;
    if (set96.first()) { e = cudaFuncSetAttribute(attention_audio_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<96>::SMEM_BYTES); if (e) return (int)e; set96.done(); }
    const __half* p = (const __half*)qkv;
    const int D = heads * dh;
    attention_audio_kernel<96><<<grid, 128, AttnCfg<96>::SMEM_BYTES, stream>>>(p, 3 * D, p + D, p + 2 * D, 3 * D, mask, (__half*)out, seq, seq, heads, scale_log2);
  } else if (dh == 64) {
    static PerDeviceOnce set64;
    if (set64.first()) { e = cudaFuncSetAttribute(attention_audio_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<64>::SMEM_BYTES); if (e) return (int)e; set64.done(); }
    const __half* p = (const __half*)qkv;
    const int D = heads * dh;
    attention_audio_kernel<64><<<grid, 128, AttnCfg<64>::SMEM_BYTES, stream>>>(p, 3 * D, p + D, p + 2 * D, 3 * D, mask, (__half*)out, seq, seq, heads, scale_log2);
  } else {
    return CACO_ERR_ARG;
  }
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// text tower (roberta.py:86-102, mask :297-310): the same kernel with head_dim 64 and the causal predicate
// ------------------------------------------------------------------------------------------------
int attention_text(const void* qkv, const float* key_mask, void* out, int batch, int T, int heads, int dh,
                   cudaStream_t stream) {
  if (!qkv || !key_mask || !out || batch <= 0 || T <= 0 || heads <= 0 || dh != 64) return CACO_ERR_ARG;
  // the flash-style warp-MMA kernel with the causal predicate: one CTA per (64 query rows, head, caption).  The scalar
  // per-(caption, head) kernel above took 0.10 ms per layer at B = 256, T = 32 (latency-bound shuffle chains).
  static PerDeviceOnce attr;
  if (attr.first()) {
    cudaError_t e = cudaFuncSetAttribute(attention_audio_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<64>::SMEM_BYTES);
    if (e) return (int)e;
    attr.done();
  }
  dim3 grid((T + AT_BM - 1) / AT_BM, heads, batch);
  const __half* p = (const __half*)qkv;
  const int D = heads * dh;
  attention_audio_kernel<64, true><<<grid, 128, AttnCfg<64>::SMEM_BYTES, stream>>>(
      p, 3 * D, p + D, p + 2 * D, 3 * D, key_mask, (__half*)out, T, T, heads, (1.0f / sqrtf((float)dh)) * 1.4426950408889634f);
  count_launch();
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// cross-attention of the captioning decoder (roberta.py:76-102 with key_value_states, mask :358-361): queries from the text
// side ([batch*Tq, heads*64], row pitch ldq), keys | values from the audio side (one [batch*Skv, 2*heads*64] buffer: k | v)
// ------------------------------------------------------------------------------------------------
int attention_cross(const void* q, int ldq, const void* kv, const float* key_mask, void* out, int batch, int Tq, int Skv, int heads,
                    int dh, cudaStream_t stream) {
  if (!q || !kv || !key_mask || !out || batch <= 0 || Tq <= 0 || Skv <= 0 || heads <= 0 || dh != 64) return CACO_ERR_ARG;
  const int D = heads * dh;
  if ((ldq & 7) || (reinterpret_cast<uintptr_t>(q) & 15) || (reinterpret_cast<uintptr_t>(kv) & 15)) return CACO_ERR_ALIGN;
  static PerDeviceOnce attr;
  if (attr.first()) {
    cudaError_t e = cudaFuncSetAttribute(attention_audio_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<64>::SMEM_BYTES);
    if (e) return (int)e;
    attr.done();
  }
  dim3 grid((Tq + AT_BM - 1) / AT_BM, heads, batch);
  const __half* k = (const __half*)kv;
  attention_audio_kernel<64, false><<<grid, 128, AttnCfg<64>::SMEM_BYTES, stream>>>(
      (const __half*)q, ldq, k, k + D, 2 * D, key_mask, (__half*)out, Tq, Skv, heads, (1.0f / sqrtf((float)dh)) * 1.4426950408889634f);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace caco

extern "C" int caco_attention_audio(const void* qkv, const float* mask, void* out, int batch, int seq, int heads, int dh,
                                    void* stream) {
  return caco::attention_audio(qkv, mask, out, batch, seq, heads, dh, (cudaStream_t)stream);
}
extern "C" int caco_attention_cross(const void* q, int ldq, const void* kv, const float* key_mask, void* out, int batch, int Tq,
                                    int Skv, int heads, int dh, void* stream) {
  return caco::attention_cross(q, ldq, kv, key_mask, out, batch, Tq, Skv, heads, dh, (cudaStream_t)stream);
}
extern "C" int caco_attention_text(const void* qkv, const float* key_mask, void* out, int batch, int T, int heads, int dh,
                                   void* stream) {
  return caco::attention_text(qkv, key_mask, out, batch, T, heads, dh, (cudaStream_t)stream);
}
