// K3a — audio self-attention (mae.py:69-74,89-92), head_dim 96: persistent, tensor-memory-resident P, double-buffered S.
//
// Lessons built in (clock64 traces of the issuing thread and of one softmax warp, profiles/r01_attention_notes.md):
//   * with P staged through shared memory the M128 tcgen05.mma instructions are bound by their smem operand fetch and the
//     single issuing thread blocks on it (tc2: ~1900 cycles of issue per 64-key block of two tiles)  ->  P is written as
//     packed fp16 into the tensor-memory columns of the S tile it came from and P V takes its A operand from tensor memory;
//   * with one S buffer per tile (tc3, 128-key blocks) a tile's chain softmax -> P V -> Q K^T -> softmax is serial and
//     exposes ~1400 cycles of tensor work per block  ->  S is double-buffered per tile (64-key blocks: 4 x 64 columns), so
//     Q K^T of block g+2 is issued right behind P V of block g and is long finished when the softmax warps come back;
//   * tcgen05.ld takes several hundred cycles while the tensor pipe is busy  ->  both 32-column loads of a block are
//     issued back to back and the 64 scores stay in registers between the max pass and the exp2 pass;
//   * per-row strided 16-byte global stores from the epilogue cost ~2800 cycles per item  ->  O is staged (swizzled) in the
//     item's dead Q tile and leaves through one TMA store per warp.
// Work item = 256 queries (tiles A, B) of one (clip, head); every CTA walks a static item list as ONE flat pipeline over
// 64-key blocks.  K and V ride 4-stage rings fed by two producer warps; Q and the key-mask bias are double-buffered per
// item by a third.  All tensor-memory hazards (P over S, next S over P, next item's O) are ordered by the in-order MMA pipe.
//   warps 0 / 10 / 11  TMA producers (K, V, Q + bias; 3-D maps over qkv[clip][token][3*768], OOB rows read as zero)
//   warp 1             tcgen05.mma issuer: S_x(g) = Q_x K_g^T (M128 N64 K96: SW128 + SW64 K-major tiles),
//                      O_x += P_x(g) V_g (M128 N96 K64: A from tensor memory, V MN-major from smem)
//   warps 2-5, 6-9     softmax warpgroups A, B (thread = query row): base-2 online softmax, lazy rescaling (threshold 2^8)
// Tensor memory: S_A0 S_A1 S_B0 S_B1 (4 x 64 columns, P aliases the first 32 of each) + O_A O_B (2 x 96) = 448 of 512.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "caco_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace caco {

namespace t4 {
constexpr int BM = 128, BN = 64, DH = 96, NST = 4;
constexpr uint32_t Q_TILE = 24576;                  // 16 KB SW128 (cols 0..63) + 8 KB SW64 (cols 64..95)
constexpr uint32_t K_TILE = 12288;                  // 8 KB SW128 + 4 KB SW64
constexpr uint32_t V_TILE = 16384;                  // two 64-column blocks of 64 keys, 8192 B apart
constexpr uint32_t OFF_Q = 0;                       // [buf 2][tile 2] x Q_TILE
constexpr uint32_t OFF_K = 4 * Q_TILE;              // NST x K_TILE
constexpr uint32_t OFF_V = OFF_K + NST * K_TILE;    // NST x V_TILE
constexpr uint32_t OFF_BAR = OFF_V + NST * V_TILE;  // 256 B: mbarriers, tmem slot
constexpr uint32_t OFF_FLAG = OFF_BAR + 256;        // [buf 2][64] per-block "has a masked key" flags
constexpr uint32_t OFF_BIAS = OFF_FLAG + 512;       // [buf 2] x max_keys floats
constexpr uint32_t TM_S = 0, TM_O = 256, TM_COLS = 512;
constexpr float RESCALE_T = 8.0f;
constexpr uint32_t B_QFULL = 0, B_ITEMDONE = 16, B_BIASFULL = 32, B_KFULL = 48, B_KEMPTY = 80, B_VFULL = 112, B_VEMPTY = 144,
                   B_SFULL = 176 /* [tile][buf] */, B_PFULL = 208, B_PVDONE = 224, B_TMEMSLOT = 240;
}  // namespace t4

__device__ long long* g_attn4_trace = nullptr;
#define TC4_STAMP(role, blk, ev)                                                              \
  do {                                                                                        \
    if (trace != nullptr && (blk) < 64) trace[((role) * 64 + (blk)) * 8 + (ev)] = clock64(); \
  } while (0)

// mbarrier wait with a watchdog: a protocol bug traps with a message instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity, int id, int g) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
      printf("attention_tc4: barrier timeout id=%d g=%d parity=%u cta=%d warp=%d lane=%d\n", id, g, parity, (int)blockIdx.x,
             (int)(threadIdx.x >> 5), (int)(threadIdx.x & 31));
      __trap();
    }
  }
}

struct Attn4Args {
  const float* mask;
  int S, H, B;
  int n_items, qpairs, n_blocks, max_keys;
  float scale_log2;
};

__global__ void __launch_bounds__(384, 1)
attention_tc4_kernel(const __grid_constant__ CUtensorMap map_q64, const __grid_constant__ CUtensorMap map_q32,
                     const __grid_constant__ CUtensorMap map_k64, const __grid_constant__ CUtensorMap map_k32,
                     const __grid_constant__ CUtensorMap map_o64, const __grid_constant__ CUtensorMap map_o32, const Attn4Args a) {
  using namespace t4;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar = sb + OFF_BAR;
  int* s_flag = reinterpret_cast<int*>(smem + OFF_FLAG);
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.H * DH, nb = a.n_blocks;
  const int n_local = (a.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = n_local * nb;
  long long* trace = (blockIdx.x == 0 && lane == 0 && (warp == 1 || warp == 2)) ? g_attn4_trace : nullptr;

  if (tid == 0) {
    if ((sb & 1023u) != 0) __trap();
    tma_prefetch_desc(&map_q64); tma_prefetch_desc(&map_q32); tma_prefetch_desc(&map_k64); tma_prefetch_desc(&map_k32);
    tma_prefetch_desc(&map_o64); tma_prefetch_desc(&map_o32);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar + B_QFULL + 8 * i, 1);
      mbar_init(bar + B_ITEMDONE + 8 * i, 8);
      mbar_init(bar + B_BIASFULL + 8 * i, 1);
      mbar_init(bar + B_PFULL + 8 * i, 4);
      mbar_init(bar + B_PVDONE + 8 * i, 1);
    }
    for (int i = 0; i < NST; ++i) {
      mbar_init(bar + B_KFULL + 8 * i, 1); mbar_init(bar + B_KEMPTY + 8 * i, 1);
      mbar_init(bar + B_VFULL + 8 * i, 1); mbar_init(bar + B_VEMPTY + 8 * i, 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(bar + B_SFULL + 8 * i, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<1>(bar + B_TMEMSLOT, TM_COLS);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + B_TMEMSLOT);

  auto decode = [&](int it, int& b, int& h, int& q0) {
    const int item = (int)blockIdx.x + it * (int)gridDim.x;
    const int qp = item % a.qpairs;
    const int bh = item / a.qpairs;
    h = bh % a.H;
    b = bh / a.H;
    q0 = qp * 2 * BM;
  };

  if (warp == 0) {
    // ================================================================ K producer (slot of K_g frees when Q K_{g-4}^T retires)
    for (int it = 0; it < n_local; ++it) {
      int b, h, q0;
      decode(it, b, h, q0);
      for (int j = 0; j < nb; ++j) {
        const int g = it * nb + j, st = g % NST;
        if (g >= NST) mbar_wait_wd(bar + B_KEMPTY + 8 * st, ((g / NST) + 1) & 1, 1, g);
        if (lane == 0) {
          const uint32_t kf = bar + B_KFULL + 8 * st;
          mbar_expect_tx(kf, K_TILE);
          tma_load_3d(sb + OFF_K + st * K_TILE, &map_k64, kf, D + h * DH, j * BN, b);
          tma_load_3d(sb + OFF_K + st * K_TILE + 8192, &map_k32, kf, D + h * DH + 64, j * BN, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 10) {
    // ================================================================ V producer (slot of V_g frees when P V_{g-4} retires)
    for (int it = 0; it < n_local; ++it) {
      int b, h, q0;
      decode(it, b, h, q0);
      for (int j = 0; j < nb; ++j) {
        const int g = it * nb + j, st = g % NST;
        if (g >= NST) mbar_wait_wd(bar + B_VEMPTY + 8 * st, ((g / NST) + 1) & 1, 2, g);
        if (lane == 0) {
          const uint32_t vf = bar + B_VFULL + 8 * st;
          mbar_expect_tx(vf, V_TILE);
          tma_load_3d(sb + OFF_V + st * V_TILE, &map_k64, vf, 2 * D + h * DH, j * BN, b);
          tma_load_3d(sb + OFF_V + st * V_TILE + 8192, &map_k64, vf, 2 * D + h * DH + 64, j * BN, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 11) {
    // ================================================================ Q + key-mask bias producer (per item, double-buffered)
    for (int it = 0; it < n_local; ++it) {
      int b, h, q0;
      decode(it, b, h, q0);
      const int ib = it & 1;
      if (it >= 2) mbar_wait_wd(bar + B_ITEMDONE + 8 * ib, ((it >> 1) + 1) & 1, 3, it);
      if (lane == 0) {
        const uint32_t qf = bar + B_QFULL + 8 * ib;
        mbar_expect_tx(qf, 2 * Q_TILE);
        for (int x = 0; x < 2; ++x) {
          const uint32_t dst = sb + OFF_Q + (ib * 2 + x) * Q_TILE;
          tma_load_3d(dst, &map_q64, qf, h * DH, q0 + x * BM, b);
          tma_load_3d(dst + 16384, &map_q32, qf, h * DH + 64, q0 + x * BM, b);
        }
      }
      // additive key bias (0 = live key, -inf = masked key or padding past S) + per-block "any masked" flag
      for (int j0 = 0; j0 < a.max_keys; j0 += BN) {
        bool any = false;
        for (int j = j0 + lane; j < j0 + BN; j += 32) {
          const bool live = (j < a.S) && (__ldg(a.mask + (size_t)b * a.S + j) != 0.0f);
          s_bias[ib * a.max_keys + j] = live ? 0.0f : -INFINITY;
          any |= !live;
        }
        any = __any_sync(0xffffffffu, any);
        if (lane == 0) s_flag[ib * 64 + j0 / BN] = any ? 1 : 0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar + B_BIASFULL + 8 * ib);
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    constexpr uint32_t idesc_qk = umma_idesc_f16(BM, BN);
    constexpr uint32_t idesc_pv = umma_idesc_f16(BM, DH, false, true);
    auto issue_qk = [&](int g, int x) {      // lane 0: S_x[g&1] = Q_x K_g^T
      const int ib = (g / nb) & 1, st = g % NST, sbuf = g & 1;
      const uint32_t k = sb + OFF_K + st * K_TILE, q = sb + OFF_Q + (ib * 2 + x) * Q_TILE;
      const uint64_t k0 = umma_desc_kmajor_sw128(k), k1 = umma_desc_kmajor_sw64(k + 8192);
      const uint64_t a0 = umma_desc_kmajor_sw128(q), a1 = umma_desc_kmajor_sw64(q + 16384);
      const uint32_t d = tmem_base + TM_S + (x * 2 + sbuf) * BN;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_f16<1>(d, a0 + 2 * ks, k0 + 2 * ks, idesc_qk, ks ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) umma_f16<1>(d, a1 + 2 * ks, k1 + 2 * ks, idesc_qk, 1u);
      umma_commit<1>(bar + B_SFULL + 8 * (x * 2 + sbuf));
      if (x == 1) umma_commit<1>(bar + B_KEMPTY + 8 * st);
    };
    auto wait_qk_inputs = [&](int g) {
      const int it = g / nb;
      if (g % nb == 0) mbar_wait_wd(bar + B_QFULL + 8 * (it & 1), (it >> 1) & 1, 4, it);
      mbar_wait_wd(bar + B_KFULL + 8 * (g % NST), (g / NST) & 1, 5, g);
      tc_fence_after();
    };
    for (int g = 0; g < 2 && g < total; ++g) {    // prologue: both S buffers of both tiles
      wait_qk_inputs(g);
      if (lane == 0) { issue_qk(g, 0); issue_qk(g, 1); }
      __syncwarp();
    }
    for (int g = 0; g < total; ++g) {
      TC4_STAMP(0, g, 0);
      const int st = g % NST, sbuf = g & 1;
      const uint32_t acc0 = (g % nb) ? 1u : 0u;
      mbar_wait_wd(bar + B_VFULL + 8 * st, (g / NST) & 1, 6, g);
      if (g + 2 < total) wait_qk_inputs(g + 2);
      TC4_STAMP(0, g, 1);
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        mbar_wait_wd(bar + B_PFULL + 8 * x, g & 1, 7, g);
        TC4_STAMP(0, g, 2 + 2 * x);
        tc_fence_after();
        if (lane == 0) {
          // O_x += P_x V_g : A = P_x from tensor memory (packed fp16, 8 columns per 16 keys), B = V_g MN-major from smem
#pragma unroll
          for (int ks = 0; ks < BN / 16; ++ks) {
            const uint64_t vb = umma_desc_mnmajor_sw128(sb + OFF_V + st * V_TILE + ks * 2048, 8192);
            umma_f16_ts<1>(tmem_base + TM_O + x * DH, tmem_base + TM_S + (x * 2 + sbuf) * BN + ks * 8, vb, idesc_pv,
                           (acc0 | (uint32_t)ks) ? 1u : 0u);
          }
          umma_commit<1>(bar + B_PVDONE + 8 * x);
          if (x == 1) umma_commit<1>(bar + B_VEMPTY + 8 * st);
          if (g + 2 < total) issue_qk(g + 2, x);      // reuses S_x[sbuf] (= P_x(g)): ordered behind P V by the in-order pipe
        }
        __syncwarp();
        TC4_STAMP(0, g, 3 + 2 * x);
      }
    }
  } else if (warp < 10) {
    // ================================================================ softmax warpgroups (thread = query row)
    const int x = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + (uint32_t(quarter * 32) << 16);
    const uint32_t t_o = t_lane + TM_O + x * DH;
    const uint32_t b_sfull = bar + B_SFULL + 8 * (x * 2), b_pfull = bar + B_PFULL + 8 * x, b_pvdone = bar + B_PVDONE + 8 * x;
    float m_ref = -INFINITY, l_run = 0.f;
    int b = 0, h = 0, q0 = 0;
    for (int g = 0; g < total; ++g) {
      const int it = g / nb, j = g - it * nb, ib = it & 1, sbuf = g & 1;
      if (j == 0) {
        decode(it, b, h, q0);
        mbar_wait_wd(bar + B_BIASFULL + 8 * ib, (it >> 1) & 1, 8, it);
        m_ref = -INFINITY;
        l_run = 0.f;
      }
      const bool masked = s_flag[ib * 64 + j] != 0;           // warp-uniform
      const float* bias = s_bias + ib * a.max_keys + j * BN;
      const uint32_t t_s = t_lane + TM_S + (x * 2 + sbuf) * BN;
      TC4_STAMP(1, g, 0);
      mbar_wait_wd(b_sfull + 8 * sbuf, (g >> 1) & 1, 9, g);
      TC4_STAMP(1, g, 1);
      tc_fence_after();
      uint32_t v[2][32];
      tmem_ld_32x32(t_s, v[0]);
      tmem_ld_32x32(t_s + 32, v[1]);
      tmem_ld_wait();
      // ---- pass 1: row maximum (scale > 0, so max commutes with the scaling); masked blocks add the 0 / -inf key bias
      float mx = -INFINITY;
      if (masked) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float t = __uint_as_float(v[c][i]) + bias[c * 32 + i];
            v[c][i] = __float_as_uint(t);
            mx = fmaxf(mx, t);
          }
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[c][i]));
      }
      mx *= a.scale_log2;                                     // -inf stays -inf
      TC4_STAMP(1, g, 2);
      // ---- lazy rescale of O_x and l (only when the maximum grew by more than 2^8 since the reference was taken)
      const bool need = mx > m_ref + RESCALE_T;               // m_ref == -inf: true iff this block has a live key
      if (__any_sync(0xffffffffu, need)) {
        const float factor = need ? exp2f(m_ref - mx) : 1.0f;
        if (j > 0) {
          mbar_wait_wd(b_pvdone, (g - 1) & 1, 10, g);                   // O_x must hold every P V of this item issued so far
          tc_fence_after();
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {
            uint32_t o[3][16];
#pragma unroll
            for (int c = 0; c < 3; ++c) tmem_ld_32x16(t_o + (hc * 3 + c) * 16, o[c]);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 3; ++c) {
#pragma unroll
              for (int i = 0; i < 16; ++i) o[c][i] = __float_as_uint(__uint_as_float(o[c][i]) * factor);
              tmem_st_32x16(t_o + (hc * 3 + c) * 16, o[c]);
            }
          }
        }
        l_run *= factor;
        if (need) m_ref = mx;
      }
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      // ---- pass 2 (from registers): p = exp2(s*scale - m), row sum, pack to fp16, write P over the first 32 columns of S
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t ph[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(v[c][2 * i]), a.scale_log2, neg_m));       // -inf * scale = -inf -> 0
          const float p1 = fast_exp2(fmaf(__uint_as_float(v[c][2 * i + 1]), a.scale_log2, neg_m));
          sum += p0 + p1;
          __half2 hh = __floats2half2_rn(p0, p1);
          ph[i] = *reinterpret_cast<uint32_t*>(&hh);
        }
        tmem_st_32x16(t_s + c * 16, ph);                      // keys 32c..32c+31 -> columns 16c..16c+15
      }
      l_run += sum;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_pfull);
      TC4_STAMP(1, g, 3);
      if (j == nb - 1) {
        // ---- item epilogue: O / l -> fp16 (the next item's first P V needs this warp's next P, so O_x is safe to read)
        mbar_wait_wd(b_pvdone, g & 1, 10, g);
        TC4_STAMP(1, g, 4);
        tc_fence_after();
        const float inv = 1.0f / l_run;                        // l == 0 (no live key): NaN row, like torch.softmax
        // O tile -> fp16 -> this item's (now dead) Q_x buffer in the swizzled layouts of the two output tensor maps ->
        // one TMA store per warp (32 rows x 64 + 32 columns); rows past the clip's end are clipped by the hardware
        uint8_t* stage = smem + OFF_Q + (ib * 2 + x) * Q_TILE;
        uint8_t* r0 = stage + row * 128;
        uint8_t* r1 = stage + 16384 + row * 64;
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
          uint32_t o[3][16];
#pragma unroll
          for (int c = 0; c < 3; ++c) tmem_ld_32x16(t_o + (hc * 3 + c) * 16, o[c]);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __half2 hh = __floats2half2_rn(__uint_as_float(o[c][2 * i]) * inv, __uint_as_float(o[c][2 * i + 1]) * inv);
              pk[i] = *reinterpret_cast<uint32_t*>(&hh);
            }
            const int col16 = (hc * 3 + c) * 2;                // index of the first of two 16-byte chunks (8 fp16 each)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int ch = col16 + k;                        // 0..11
              const uint4 u = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
              if (ch < 8) *reinterpret_cast<uint4*>(r0 + ((ch ^ (row & 7)) << 4)) = u;                       // SWIZZLE_128B
              else *reinterpret_cast<uint4*>(r1 + (((ch - 8) ^ ((row >> 1) & 3)) << 4)) = u;                 // SWIZZLE_64B
            }
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int qrow = q0 + x * BM + quarter * 32;
          tma_store_3d(&map_o64, sb + OFF_Q + (ib * 2 + x) * Q_TILE + quarter * 32 * 128, h * DH, qrow, b);
          tma_store_3d(&map_o32, sb + OFF_Q + (ib * 2 + x) * Q_TILE + 16384 + quarter * 32 * 64, h * DH + 64, qrow, b);
          tma_store_commit();
          tma_store_wait_read<0>();                            // smem may be refilled with the next-but-one item's Q
        }
        TC4_STAMP(1, g, 5);
        __syncwarp();
        if (lane == 0) mbar_arrive(bar + B_ITEMDONE + 8 * ib);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem_base, TM_COLS);
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled4 encode_fn4() {
  static PFN_encodeTiled4 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled4>(p);
  }
  return fn;
}
static int make_map3d(CUtensorMap* m, const void* base, int batch, int seq, int ld, int box_cols, int box_rows, bool sw128) {
  PFN_encodeTiled4 enc = encode_fn4();
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

int attention_audio_tc4(const void* qkv, const float* mask, void* out, int batch, int seq, int heads, int dh,
                        cudaStream_t stream) {
  using namespace t4;
  if (!qkv || !mask || !out || batch <= 0 || seq <= 0 || heads <= 0 || dh != DH) return CACO_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return CACO_ERR_ALIGN;
  Attn4Args a;
  a.mask = mask; a.S = seq; a.H = heads; a.B = batch;
  a.qpairs = (seq + 2 * BM - 1) / (2 * BM);
  a.n_items = batch * heads * a.qpairs;
  a.n_blocks = (seq + BN - 1) / BN;
  a.max_keys = a.n_blocks * BN;
  a.scale_log2 = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
  const size_t smem = OFF_BIAS + 2 * (size_t)a.max_keys * 4;
  if (smem > 232448 || a.n_blocks > 64) return CACO_ERR_ARG;
  const int ld = 3 * heads * dh;
  CUtensorMap q64, q32, k64, k32, o64, o32;
  int rc;
  if ((rc = make_map3d(&q64, qkv, batch, seq, ld, 64, BM, true))) return rc;
  if ((rc = make_map3d(&q32, qkv, batch, seq, ld, 32, BM, false))) return rc;
  if ((rc = make_map3d(&k64, qkv, batch, seq, ld, 64, BN, true))) return rc;
  if ((rc = make_map3d(&k32, qkv, batch, seq, ld, 32, BN, false))) return rc;
  if ((rc = make_map3d(&o64, out, batch, seq, heads * dh, 64, 32, true))) return rc;      // per-warp store boxes: 32 rows
  if ((rc = make_map3d(&o32, out, batch, seq, heads * dh, 32, 32, false))) return rc;
  static size_t cur = 0;
  if (smem > cur) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return (int)e;
    cur = smem;
  }
  int grid = num_sms();
  if (grid > a.n_items) grid = a.n_items;
  attention_tc4_kernel<<<grid, 384, smem, stream>>>(q64, q32, k64, k32, o64, o32, a);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace caco

extern "C" int caco_attn4_trace(void* dev_buf) {
  long long* p = (long long*)dev_buf;
  return (int)cudaMemcpyToSymbol(caco::g_attn4_trace, &p, sizeof(p));
}
