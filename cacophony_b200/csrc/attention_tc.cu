// K3a on the 5th-gen tensor cores — audio self-attention (mae.py:69-74,89-92), head_dim 96.
//
// One CTA = 128 queries of one (clip, head); keys are streamed in 64-key blocks.
//   warp 0   TMA producer: Q once, then K_j / V_j through a 2-stage ring (3-D tensor maps over
//            qkv[clip][token][3*768]; rows past the clip's end are zero-filled by TMA, never read from a neighbour)
//   warp 1   tcgen05.mma issuer:  S_j = Q K_j^T  (M128 x N64 x K96: K-major operands, 64 columns in SWIZZLE_128B tiles
//            + 32 columns in SWIZZLE_64B tiles)  and  O += P_j V_j  (M128 x N96 x K64, V consumed MN-major straight
//            from its [key][d] layout);  S is double-buffered in tensor memory so S_{j+1} is computed during softmax_j
//   warps 2-5  softmax: one thread per query row, tcgen05.ld of the 64 scores, base-2 online softmax in fp32 with
//            LAZY rescaling (O and the running sum are only rescaled when the row maximum grows by more than 2^8, so
//            the common case never touches O in tensor memory), P_j -> fp16 -> swizzled smem for the next MMA
// Two CTAs are resident per SM (100 KB smem, 256 TMEM columns each) so one CTA's exp2 work overlaps the other's MMAs.
// The scores (H*S^2*4 B = 8 MB per clip per layer if materialised) never leave the SM.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>

#include "caco_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace caco {

constexpr int TA_BM = 128, TA_BN = 64, TA_DH = 96;
constexpr uint32_t TA_Q0 = 0, TA_Q1 = 16384;
constexpr int TA_KSTAGES = 3, TA_VSTAGES = 2;
constexpr uint32_t TA_K0 = 24576;   // + stage * 8192   (3 stages: K_{j+3} is requested as soon as Q K_j^T retires)
constexpr uint32_t TA_K1 = 49152;   // + stage * 4096
constexpr uint32_t TA_V = 61440;    // + stage * 16384  (two 64-column blocks, 8192 B apart)
constexpr uint32_t TA_P = 94208;
constexpr uint32_t TA_BAR = 110592;
constexpr uint32_t TA_BIAS = 110592 + 192;
constexpr uint32_t TA_Q_BYTES = 24576, TA_K_BYTES = 12288, TA_V_BYTES = 16384;
constexpr uint32_t TA_TMEM_COLS = 256, TA_S_COL = 0, TA_O_COL = 128;
constexpr float TA_RESCALE_THRESHOLD = 8.0f;   // log2 units

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&f)[16]) {
  uint32_t v[16];
  tmem_ld_32x16(taddr, v);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
}

__global__ void __launch_bounds__(192, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_q1,
                    const __grid_constant__ CUtensorMap map_kv0, const __grid_constant__ CUtensorMap map_k1,
                    const float* __restrict__ mask, __half* __restrict__ out, int S, int H, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];     // swizzled tiles need 1024-byte alignment
  const uint32_t sb = smem_u32(smem_raw);
  uint8_t* sg = smem_raw;
  const uint32_t bar = sb + TA_BAR;
  const uint32_t b_qfull = bar, b_kfull = bar + 8, b_kempty = bar + 32, b_vfull = bar + 56, b_vempty = bar + 72,
                 b_sfull = bar + 88, b_sempty = bar + 104, b_pfull = bar + 120, b_pvdone = bar + 128, tmem_slot = bar + 136;
  float* s_bias = reinterpret_cast<float*>(sg + TA_BIAS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * TA_BM, h = blockIdx.y, b = blockIdx.z;
  const int n_blocks = (S + TA_BN - 1) / TA_BN;
  const int D = H * TA_DH;

  auto load_k = [&](int j) {
    const int ks = j % TA_KSTAGES;
    mbar_expect_tx(b_kfull + 8 * ks, TA_K_BYTES);
    tma_load_3d(sb + TA_K0 + ks * 8192, &map_kv0, b_kfull + 8 * ks, D + h * TA_DH, j * TA_BN, b);
    tma_load_3d(sb + TA_K1 + ks * 4096, &map_k1, b_kfull + 8 * ks, D + h * TA_DH + 64, j * TA_BN, b);
  };
  auto load_v = [&](int j) {
    const int vs = j % TA_VSTAGES;
    mbar_expect_tx(b_vfull + 8 * vs, TA_V_BYTES);
    tma_load_3d(sb + TA_V + vs * 16384, &map_kv0, b_vfull + 8 * vs, 2 * D + h * TA_DH, j * TA_BN, b);
    tma_load_3d(sb + TA_V + vs * 16384 + 8192, &map_kv0, b_vfull + 8 * vs, 2 * D + h * TA_DH + 64, j * TA_BN, b);
  };
  if (tid == 0) {
    if ((sb & 1023u) != 0) __trap();
    tma_prefetch_desc(&map_q0); tma_prefetch_desc(&map_q1); tma_prefetch_desc(&map_kv0); tma_prefetch_desc(&map_k1);
    mbar_init(b_qfull, 1);
    for (int s = 0; s < TA_KSTAGES; ++s) { mbar_init(b_kfull + 8 * s, 1); mbar_init(b_kempty + 8 * s, 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(b_vfull + 8 * s, 1);
      mbar_init(b_vempty + 8 * s, 1);
      mbar_init(b_sfull + 8 * s, 1);
      mbar_init(b_sempty + 8 * s, 4);
    }
    mbar_init(b_pfull, 4);
    mbar_init(b_pvdone, 1);
    fence_mbar_init();
    // prologue loads go out before the CTA-wide barrier so their latency overlaps the mask fetch and the TMEM allocation
    mbar_expect_tx(b_qfull, TA_Q_BYTES);
    tma_load_3d(sb + TA_Q0, &map_q0, b_qfull, h * TA_DH, q0, b);
    tma_load_3d(sb + TA_Q1, &map_q1, b_qfull, h * TA_DH + 64, q0, b);
    for (int j = 0; j < TA_KSTAGES && j < n_blocks; ++j) load_k(j);
    for (int j = 0; j < TA_VSTAGES && j < n_blocks; ++j) load_v(j);
  }
  if (warp == 1) {
    tmem_alloc<1>(tmem_slot, TA_TMEM_COLS);
    tmem_relinquish<1>();
  }
  // additive key bias: 0 for keys the clip's mask keeps, -inf for masked keys and for the padding past S
  for (int j = tid; j < n_blocks * TA_BN; j += blockDim.x)
    s_bias[j] = (j < S && mask[(size_t)b * S + j] != 0.0f) ? 0.0f : -INFINITY;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sg + TA_BAR + 136);

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer (steady state)
    // V_j's slot frees when P V_{j-2} retires, K_{j+3}'s slot when Q K_j^T retires: both in tensor-pipe order
    for (int j = 0; j < n_blocks; ++j) {
      if (j >= TA_VSTAGES) {
        mbar_wait(b_vempty + 8 * (j % TA_VSTAGES), ((j / TA_VSTAGES) + 1) & 1);
        if (lane == 0) load_v(j);
      }
      const int jk = j + TA_KSTAGES;
      if (jk < n_blocks) {
        mbar_wait(b_kempty + 8 * (jk % TA_KSTAGES), ((jk / TA_KSTAGES) + 1) & 1);
        if (lane == 0) load_k(jk);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    constexpr uint32_t idesc_qk = umma_idesc_f16(TA_BM, TA_BN);
    constexpr uint32_t idesc_pv = umma_idesc_f16(TA_BM, TA_DH, false, true);   // B = V is MN-major ([key][d])
    auto issue_qk = [&](int j) {
      const int s = j & 1, ks_ = j % TA_KSTAGES;
      const uint64_t a0 = umma_desc_kmajor_sw128(sb + TA_Q0), a1 = umma_desc_kmajor_sw64(sb + TA_Q1);
      const uint64_t k0 = umma_desc_kmajor_sw128(sb + TA_K0 + ks_ * 8192), k1 = umma_desc_kmajor_sw64(sb + TA_K1 + ks_ * 4096);
      const uint32_t d = tmem_base + TA_S_COL + s * TA_BN;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_f16<1>(d, a0 + 2 * ks, k0 + 2 * ks, idesc_qk, ks ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) umma_f16<1>(d, a1 + 2 * ks, k1 + 2 * ks, idesc_qk, 1u);
      umma_commit<1>(b_sfull + 8 * s);
      umma_commit<1>(b_kempty + 8 * ks_);
    };
    mbar_wait(b_qfull, 0);
    mbar_wait(b_kfull, 0);
    tc_fence_after();
    if (lane == 0) issue_qk(0);
    __syncwarp();
    for (int j = 0; j < n_blocks; ++j) {
      if (j + 1 < n_blocks) {
        const int s1 = (j + 1) & 1;
        mbar_wait(b_kfull + 8 * ((j + 1) % TA_KSTAGES), ((j + 1) / TA_KSTAGES) & 1);
        if (j + 1 >= 2) mbar_wait(b_sempty + 8 * s1, (((j + 1) >> 1) + 1) & 1);
        tc_fence_after();
        if (lane == 0) issue_qk(j + 1);
        __syncwarp();
      }
      const int s = j & 1;
      mbar_wait(b_vfull + 8 * s, (j >> 1) & 1);
      mbar_wait(b_pfull, j & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t pa = umma_desc_kmajor_sw128(sb + TA_P);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t vb = umma_desc_mnmajor_sw128(sb + TA_V + s * 16384 + ks * 2048, 8192);
          umma_f16<1>(tmem_base + TA_O_COL, pa + 2 * ks, vb, idesc_pv, (j | ks) ? 1u : 0u);
        }
        umma_commit<1>(b_pvdone);
        umma_commit<1>(b_vempty + 8 * s);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- softmax warps (thread = query row)
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t(quarter * 32) << 16);
    uint8_t* p_row = sg + TA_P + row * 128;
    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_blocks; ++j) {
      const int s = j & 1;
      mbar_wait(b_sfull + 8 * s, (j >> 1) & 1);
      tc_fence_after();
      uint32_t v0[32], v1[32];
      tmem_ld_32x32(t_lane + TA_S_COL + s * TA_BN, v0);
      tmem_ld_32x32(t_lane + TA_S_COL + s * TA_BN + 32, v1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_sempty + 8 * s);
      float t[64];
      float mx = -INFINITY;
      const float4* bias4 = reinterpret_cast<const float4*>(s_bias + j * TA_BN);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float4 bb = bias4[c];
        const uint32_t* src = (c < 8) ? &v0[4 * c] : &v1[4 * (c - 8)];
        t[4 * c + 0] = fmaf(__uint_as_float(src[0]), scale_log2, bb.x);
        t[4 * c + 1] = fmaf(__uint_as_float(src[1]), scale_log2, bb.y);
        t[4 * c + 2] = fmaf(__uint_as_float(src[2]), scale_log2, bb.z);
        t[4 * c + 3] = fmaf(__uint_as_float(src[3]), scale_log2, bb.w);
        mx = fmaxf(mx, fmaxf(fmaxf(t[4 * c], t[4 * c + 1]), fmaxf(t[4 * c + 2], t[4 * c + 3])));
      }
      // lazy rescale: only when this row's maximum grew by more than the threshold since the reference was taken
      const bool need = mx > m_ref + TA_RESCALE_THRESHOLD;      // m_ref == -inf: true iff the block has a live key
      if (__any_sync(0xffffffffu, need)) {
        const float factor = need ? exp2f(m_ref - mx) : 1.0f;   // exp2(-inf) = 0 on the first live block
        if (j > 0) {
          mbar_wait(b_pvdone, (j - 1) & 1);                     // O must hold every P V issued so far
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < TA_DH / 16; ++c) {
            uint32_t o[16];
            tmem_ld_32x16(t_lane + TA_O_COL + c * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * factor);
            tmem_st_32x16(t_lane + TA_O_COL + c * 16, o);
          }
          tmem_st_wait();
          tc_fence_before();
        }
        l_run *= factor;
        if (need) m_ref = mx;
      }
      const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
      uint32_t ph[32];
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float p0 = fast_exp2(t[2 * c] - m_use), p1 = fast_exp2(t[2 * c + 1] - m_use);
        sum += p0 + p1;
        __half2 hh = __floats2half2_rn(p0, p1);
        ph[c] = *reinterpret_cast<uint32_t*>(&hh);
      }
      l_run += sum;
      if (j > 0) mbar_wait(b_pvdone, (j - 1) & 1);              // P buffer is free once P_{j-1} V_{j-1} retired
#pragma unroll
      for (int cc = 0; cc < 8; ++cc) {
        uint4 u = make_uint4(ph[4 * cc], ph[4 * cc + 1], ph[4 * cc + 2], ph[4 * cc + 3]);
        *reinterpret_cast<uint4*>(p_row + ((cc ^ (row & 7)) << 4)) = u;   // 128-byte swizzle, as UMMA expects
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_pfull);
    }
    // ---- epilogue: O / l -> fp16
    mbar_wait(b_pvdone, (n_blocks - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l_run;                             // l == 0 (no live key) -> NaN row, like torch.softmax
    const int q = q0 + row;
    __half* dst = out + ((size_t)b * S + q) * D + h * TA_DH;
#pragma unroll
    for (int c = 0; c < TA_DH / 16; ++c) {
      float f[16];
      tmem_ld16(t_lane + TA_O_COL + c * 16, f);
      if (q < S) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __half2 hh = __floats2half2_rn(f[2 * i] * inv, f[2 * i + 1] * inv);
          pk[i] = *reinterpret_cast<uint32_t*>(&hh);
        }
        *reinterpret_cast<uint4*>(dst + c * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(dst + c * 16 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem_base, TA_TMEM_COLS);
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}
// qkv viewed as [batch][seq][ld] fp16; box = [1][box_rows][box_cols]
static int make_map3(CUtensorMap* m, const void* base, int batch, int seq, int ld, int box_cols, int box_rows, bool sw128) {
  PFN_encodeTiled enc = encode_fn();
  if (!enc) return CACO_ERR_DRIVER;
  cuuint64_t dims[3] = {(cuuint64_t)ld, (cuuint64_t)seq, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)seq};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : CACO_ERR_DRIVER;
}

int attention_audio_tc(const void* qkv, const float* mask, void* out, int batch, int seq, int heads, int dh,
                       cudaStream_t stream) {
  if (!qkv || !mask || !out || batch <= 0 || seq <= 0 || heads <= 0 || dh != TA_DH) return CACO_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return CACO_ERR_ALIGN;
  const int n_blocks = (seq + TA_BN - 1) / TA_BN;
  const size_t smem = TA_BIAS + (size_t)n_blocks * TA_BN * 4;
  if (smem > 115 * 1024) return CACO_ERR_ARG;    // two CTAs per SM (seq <= 1728)
  const int ld = 3 * heads * dh;
  CUtensorMap mq0, mq1, mkv0, mk1;
  int rc;
  if ((rc = make_map3(&mq0, qkv, batch, seq, ld, 64, TA_BM, true))) return rc;
  if ((rc = make_map3(&mq1, qkv, batch, seq, ld, 32, TA_BM, false))) return rc;
  if ((rc = make_map3(&mkv0, qkv, batch, seq, ld, 64, TA_BN, true))) return rc;
  if ((rc = make_map3(&mk1, qkv, batch, seq, ld, 32, TA_BN, false))) return rc;
  static PerDeviceMax smem_max;
  if (smem_max.need(smem)) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return (int)e;
    smem_max.set(smem);
  }
  dim3 grid((seq + TA_BM - 1) / TA_BM, heads, batch);
  attention_tc_kernel<<<grid, 192, smem, stream>>>(mq0, mq1, mkv0, mk1, mask, (__half*)out, seq, heads,
                                                  (1.0f / sqrtf((float)dh)) * 1.4426950408889634f);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace caco
