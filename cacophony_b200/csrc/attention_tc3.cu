// K3a, persistent ping-pong version — audio self-attention (mae.py:69-74,89-92), head_dim 96, one CTA per SM.
//
// Why this shape (measured on the two earlier kernels with a clock64 trace of the MMA-issuing thread): with 64-key blocks
// and P staged through shared memory, a block of two 128-query tiles costs ~1900 cycles of tcgen05.mma *issue* time — the
// M128 x N64 / N96 instructions are bound by their shared-memory operand fetch (4 KB of A per instruction), not by math, and
// the single issuing thread blocks on it; tensor pipe 22 % active, MUFU 29 %.  Here
//   * Q K^T runs on 128-key blocks (M128 x N128 x K16: operand bytes per FLOP halved),
//   * P never touches shared memory: the softmax warps write it as packed fp16 into the tensor-memory columns its S tile
//     just vacated and P V reads its A operand from tensor memory (tcgen05.mma with A in TMEM), so P V only fetches V,
//   * the two query tiles A and B of a work item ping-pong: while warpgroup A does exp2 on S_A the tensor core runs
//     P_B V and Q_B K^T, and vice versa (S is single-buffered per tile; all TMEM hazards are ordered by the in-order MMA pipe).
// Work item = 256 queries of one (clip, head); every CTA walks a static list of items as ONE flat pipeline over 128-key
// blocks; K/V ride 2-stage rings fed by two producer warps, Q and the key-mask bias are double-buffered per item by a third.
//   warp 0 / 10 / 11   TMA producers: K ring, V ring, Q + mask bias (3-D tensor maps over qkv[clip][token][3*768], OOB = 0)
//   warp 1             tcgen05.mma issuer
//   warps 2-5, 6-9     softmax warpgroups A, B (thread = query row): two passes over the scores in tensor memory (max, then
//                      exp2 / sum / pack), lazy rescaling of O (threshold 2^8), per-item epilogue O / l -> fp16 -> HBM
// Tensor memory: S_A S_B (2 x 128 columns; P_X aliases the first 64 columns of S_X) + O_A O_B (2 x 96) = 448 of 512 columns.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>

#include "caco_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace caco {

namespace t3 {
constexpr int BM = 128, BN = 128, DH = 96, NST = 2;
constexpr uint32_t Q_TILE = 24576;                  // 16 KB SW128 (cols 0..63) + 8 KB SW64 (cols 64..95); K tiles alike
constexpr uint32_t OFF_Q = 0;                       // [buf 2][tile 2] x Q_TILE
constexpr uint32_t OFF_K = 4 * Q_TILE;              // NST x Q_TILE
constexpr uint32_t OFF_V = OFF_K + NST * Q_TILE;    // NST x 32768 (two 64-column blocks of 128 keys, 16384 B apart)
constexpr uint32_t OFF_BAR = OFF_V + NST * 32768;   // 256 B: mbarriers, tmem slot
constexpr uint32_t OFF_FLAG = OFF_BAR + 256;        // [buf 2][32] per-block "has a masked key" flags
constexpr uint32_t OFF_BIAS = OFF_FLAG + 256;       // [buf 2] x max_keys floats
constexpr uint32_t K_BYTES = Q_TILE, V_BYTES = 32768;
constexpr uint32_t TM_S = 0, TM_O = 256, TM_COLS = 512;
constexpr float RESCALE_T = 8.0f;
constexpr uint32_t B_QFULL = 0, B_ITEMDONE = 16, B_BIASFULL = 32, B_KFULL = 48, B_KEMPTY = 64, B_VFULL = 80, B_VEMPTY = 96,
                   B_SFULL = 112, B_PFULL = 128, B_PVDONE = 144, B_TMEMSLOT = 160;
}  // namespace t3

__device__ long long* g_attn3_trace = nullptr;
#define TC3_STAMP(role, blk, ev)                                                              \
  do {                                                                                        \
    if (trace != nullptr && (blk) < 64) trace[((role) * 64 + (blk)) * 8 + (ev)] = clock64(); \
  } while (0)

struct Attn3Args {
  const float* mask;
  __half* out;
  int S, H, B;
  int n_items, qpairs, n_blocks, max_keys;
  float scale_log2;
};

// 2^x on the FMA/ALU pipes (Cody-Waite: x = n + f, |f| <= 1/2; degree-3 minimax of 2^f, relative error 7.5e-5 — a sixth of
// the fp16 rounding P gets anyway; 2^n through the exponent field).  Used for a compile-time share of the scores so that the
// two softmax warps of an SM sub-partition do not queue on its one MUFU (16 ex2 / clk / SM is this kernel's tightest floor).
__device__ __forceinline__ float exp2_poly3(float x) {
  x = fmaxf(x, -125.0f);                               // masked keys (-inf): 2^-125, which rounds to 0 in fp16 and in the row sum
  const float r = x + 12582912.0f;                     // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float f = x - (r - 12582912.0f);
  float p = fmaf(f, 5.517137796e-02f, 2.426112145e-01f);
  p = fmaf(p, f, 6.932610273e-01f);
  p = fmaf(p, f, 9.999280572e-01f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}
// packed fp32 pairs (sm_100: FFMA2 / FADD2 issue two fp32 operations per instruction; FMNMX3 is a three-input maximum): the
// softmax warps are instruction-issue-bound (measured: every instruction added to the exp2 pass lengthens the kernel by its
// issue slot), so the passes are written in the fewest instructions per score the ISA offers
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// which of the 32 scores of a chunk take the polynomial: POLY 0 none, 1 = 1/4, 2 = 1/2, 3 = 3/8
template <int POLY>
__device__ __forceinline__ constexpr bool use_poly(int e) {
  return POLY == 1 ? (e & 3) == 3 : POLY == 2 ? (e & 1) == 1 : POLY == 3 ? ((e & 7) == 1 || (e & 7) == 4 || (e & 7) == 7) : false;
}

template <int POLY>
__global__ void __launch_bounds__(384, 1)
attention_tc3_kernel(const __grid_constant__ CUtensorMap map_64, const __grid_constant__ CUtensorMap map_32,
                     const __grid_constant__ CUtensorMap map_o64, const __grid_constant__ CUtensorMap map_o32, const Attn3Args a) {
  using namespace t3;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar = sb + OFF_BAR;
  int* s_flag = reinterpret_cast<int*>(smem + OFF_FLAG);
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.H * DH, nb = a.n_blocks;
  const int n_local = (a.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = n_local * nb;
  long long* trace = (blockIdx.x == 0 && lane == 0 && (warp == 1 || warp == 2)) ? g_attn3_trace : nullptr;

  if (tid == 0) {
    if ((sb & 1023u) != 0) __trap();
    tma_prefetch_desc(&map_64); tma_prefetch_desc(&map_32); tma_prefetch_desc(&map_o64); tma_prefetch_desc(&map_o32);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar + B_QFULL + 8 * i, 1);
      mbar_init(bar + B_ITEMDONE + 8 * i, 8);
      mbar_init(bar + B_BIASFULL + 8 * i, 1);
      mbar_init(bar + B_KFULL + 8 * i, 1); mbar_init(bar + B_KEMPTY + 8 * i, 1);
      mbar_init(bar + B_VFULL + 8 * i, 1); mbar_init(bar + B_VEMPTY + 8 * i, 1);
      mbar_init(bar + B_SFULL + 8 * i, 1);
      mbar_init(bar + B_PFULL + 8 * i, 4);
      mbar_init(bar + B_PVDONE + 8 * i, 1);
    }
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
  pdl_wait();      // the QKV projection must have completed before any q/k/v tile is fetched
  pdl_launch();

  auto decode = [&](int it, int& b, int& h, int& q0) {
    const int item = (int)blockIdx.x + it * (int)gridDim.x;
    const int qp = item % a.qpairs;
    const int bh = item / a.qpairs;
    h = bh % a.H;
    b = bh / a.H;
    q0 = qp * 2 * BM;
  };

  if (warp == 0) {
    // ================================================================ K producer
    for (int it = 0; it < n_local; ++it) {
      int b, h, q0;
      decode(it, b, h, q0);
      for (int j = 0; j < nb; ++j) {
        const int g = it * nb + j, st = g % NST;
        if (g >= NST) mbar_wait(bar + B_KEMPTY + 8 * st, ((g / NST) + 1) & 1);
        if (lane == 0) {
          const uint32_t kf = bar + B_KFULL + 8 * st;
          mbar_expect_tx(kf, K_BYTES);
          tma_load_3d(sb + OFF_K + st * Q_TILE, &map_64, kf, D + h * DH, j * BN, b);
          tma_load_3d(sb + OFF_K + st * Q_TILE + 16384, &map_32, kf, D + h * DH + 64, j * BN, b);
        }
        __syncwarp();
      }
    }
  } else if (warp == 10) {
    // ================================================================ V producer
    for (int it = 0; it < n_local; ++it) {
      int b, h, q0;
      decode(it, b, h, q0);
      for (int j = 0; j < nb; ++j) {
        const int g = it * nb + j, st = g % NST;
        if (g >= NST) mbar_wait(bar + B_VEMPTY + 8 * st, ((g / NST) + 1) & 1);
        if (lane == 0) {
          const uint32_t vf = bar + B_VFULL + 8 * st;
          mbar_expect_tx(vf, V_BYTES);
          tma_load_3d(sb + OFF_V + st * 32768, &map_64, vf, 2 * D + h * DH, j * BN, b);
          tma_load_3d(sb + OFF_V + st * 32768 + 16384, &map_64, vf, 2 * D + h * DH + 64, j * BN, b);
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
      if (it >= 2) mbar_wait(bar + B_ITEMDONE + 8 * ib, ((it >> 1) + 1) & 1);
      if (lane == 0) {
        const uint32_t qf = bar + B_QFULL + 8 * ib;
        mbar_expect_tx(qf, 2 * Q_TILE);
        for (int x = 0; x < 2; ++x) {
          const uint32_t dst = sb + OFF_Q + (ib * 2 + x) * Q_TILE;
          tma_load_3d(dst, &map_64, qf, h * DH, q0 + x * BM, b);
          tma_load_3d(dst + 16384, &map_32, qf, h * DH + 64, q0 + x * BM, b);
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
        if (lane == 0) s_flag[ib * 32 + j0 / BN] = any ? 1 : 0;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar + B_BIASFULL + 8 * ib);
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer
    constexpr uint32_t idesc_qk = umma_idesc_f16(BM, BN);
    constexpr uint32_t idesc_pv = umma_idesc_f16(BM, DH, false, true);
    auto issue_qk = [&](int g, int x) {      // lane 0: S_x = Q_x K_g^T
      const int ib = (g / nb) & 1, st = g % NST;
      const uint32_t k = sb + OFF_K + st * Q_TILE, q = sb + OFF_Q + (ib * 2 + x) * Q_TILE;
      const uint64_t k0 = umma_desc_kmajor_sw128(k), k1 = umma_desc_kmajor_sw64(k + 16384);
      const uint64_t a0 = umma_desc_kmajor_sw128(q), a1 = umma_desc_kmajor_sw64(q + 16384);
      const uint32_t d = tmem_base + TM_S + x * BN;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_f16<1>(d, a0 + 2 * ks, k0 + 2 * ks, idesc_qk, ks ? 1u : 0u);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) umma_f16<1>(d, a1 + 2 * ks, k1 + 2 * ks, idesc_qk, 1u);
      umma_commit<1>(bar + B_SFULL + 8 * x);
      if (x == 1) umma_commit<1>(bar + B_KEMPTY + 8 * st);
    };
    auto wait_qk_inputs = [&](int g) {
      const int it = g / nb;
      if (g % nb == 0) mbar_wait(bar + B_QFULL + 8 * (it & 1), (it >> 1) & 1);
      mbar_wait(bar + B_KFULL + 8 * (g % NST), (g / NST) & 1);
      tc_fence_after();
    };
    if (total > 0) {
      wait_qk_inputs(0);
      if (lane == 0) { issue_qk(0, 0); issue_qk(0, 1); }
      __syncwarp();
    }
    for (int g = 0; g < total; ++g) {
      TC3_STAMP(0, g, 0);
      const int st = g % NST;
      const uint32_t acc0 = (g % nb) ? 1u : 0u;
      mbar_wait(bar + B_VFULL + 8 * st, (g / NST) & 1);
      if (g + 1 < total) wait_qk_inputs(g + 1);
      TC3_STAMP(0, g, 1);
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        mbar_wait(bar + B_PFULL + 8 * x, g & 1);
        TC3_STAMP(0, g, 2 + 2 * x);
        tc_fence_after();
        if (lane == 0) {
          // O_x += P_x V_g : A = P_x from tensor memory (packed fp16, 8 columns per 16 keys), B = V_g MN-major from smem
#pragma unroll
          for (int ks = 0; ks < BN / 16; ++ks) {
            const uint64_t vb = umma_desc_mnmajor_sw128(sb + OFF_V + st * 32768 + ks * 2048, 16384);
            umma_f16_ts<1>(tmem_base + TM_O + x * DH, tmem_base + TM_S + x * BN + ks * 8, vb, idesc_pv, (acc0 | (uint32_t)ks) ? 1u : 0u);
          }
          umma_commit<1>(bar + B_PVDONE + 8 * x);
          if (x == 1) umma_commit<1>(bar + B_VEMPTY + 8 * st);
          if (g + 1 < total) issue_qk(g + 1, x);      // overwrites S_x / P_x: ordered after P_x V by the in-order pipe
        }
        __syncwarp();
        TC3_STAMP(0, g, 3 + 2 * x);
      }
    }
  } else if (warp < 10) {
    // ================================================================ softmax warpgroups (thread = query row)
    const int x = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_s = tmem_base + (uint32_t(quarter * 32) << 16) + TM_S + x * BN;
    const uint32_t t_o = tmem_base + (uint32_t(quarter * 32) << 16) + TM_O + x * DH;
    const uint32_t b_sfull = bar + B_SFULL + 8 * x, b_pfull = bar + B_PFULL + 8 * x, b_pvdone = bar + B_PVDONE + 8 * x;
    float m_ref = -INFINITY, l_run = 0.f;
    int b = 0, h = 0, q0 = 0;
    for (int g = 0; g < total; ++g) {
      const int it = g / nb, j = g - it * nb, ib = it & 1;
      if (j == 0) {
        decode(it, b, h, q0);
        mbar_wait(bar + B_BIASFULL + 8 * ib, (it >> 1) & 1);
        m_ref = -INFINITY;
        l_run = 0.f;
      }
      const bool masked = s_flag[ib * 32 + j] != 0;           // warp-uniform
      const float* bias = s_bias + ib * a.max_keys + j * BN;
      TC3_STAMP(1, g, 0);
      mbar_wait(b_sfull, g & 1);                              // S_x(g) complete => every earlier MMA (incl. P_x V of g-1) retired
      TC3_STAMP(1, g, 1);
      tc_fence_after();
      // ---- pass 1: row maximum of the raw scores (scale > 0, so max commutes with the scaling).  tcgen05.ld is slow while
      // the tensor pipe is busy (several hundred cycles, measured), so the next chunk is always in flight during the math.
      uint32_t va[32], vb[32];
      float mx = -INFINITY;
      tmem_ld_32x32(t_s, va);
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        tmem_ld_wait();
        uint32_t* cur = (c & 1) ? vb : va;
        uint32_t* nxt = (c & 1) ? va : vb;
        if (c + 1 < BN / 32) tmem_ld_32x32(t_s + (c + 1) * 32, *reinterpret_cast<uint32_t(*)[32]>(nxt));
        else tmem_ld_32x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(nxt));      // first chunk of pass 2
        if (masked) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint64_t sb2 = add_f32x2(pack_f32x2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1])),
                                           *reinterpret_cast<const uint64_t*>(bias + c * 32 + 2 * i));
            float s0, s1;
            unpack_f32x2(sb2, s0, s1);
            mx = max3(mx, s0, s1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = max3(mx, __uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1]));
        }
      }
      mx *= a.scale_log2;                                     // -inf stays -inf
      TC3_STAMP(1, g, 2);
      // ---- lazy rescale of O_x and l (only when the maximum grew by more than 2^8 since the reference was taken)
      const bool need = mx > m_ref + RESCALE_T;               // m_ref == -inf: true iff this block has a live key
      if (__any_sync(0xffffffffu, need)) {
        const float factor = need ? exp2f(m_ref - mx) : 1.0f;
        if (j > 0) {
          tmem_ld_wait();                                     // the pass-2 prefetch shares the wait group: drain it first
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
      // ---- pass 2: p = exp2(s*scale - m), row sum, pack to fp16 and write P over the S columns already consumed.
      // BN/32 is even, so pass 2 starts in buffer va (prefetched above) and alternates like pass 1.
      // per PAIR of scores: one FFMA2 (scale, subtract the reference), two MUFU.EX2, one FADD2 (two running sums), one F2FP
      uint64_t sum2 = pack_f32x2(0.f, 0.f);
      const uint64_t scale2 = pack_f32x2(a.scale_log2, a.scale_log2), negm2 = pack_f32x2(neg_m, neg_m);
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        tmem_ld_wait();
        uint32_t* cur = (c & 1) ? vb : va;
        uint32_t* nxt = (c & 1) ? va : vb;
        if (c + 1 < BN / 32) tmem_ld_32x32(t_s + (c + 1) * 32, *reinterpret_cast<uint32_t(*)[32]>(nxt));
        uint32_t ph[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          uint64_t t2 = fma_f32x2(pack_f32x2(__uint_as_float(cur[2 * i]), __uint_as_float(cur[2 * i + 1])), scale2, negm2);
          if (masked) t2 = add_f32x2(t2, *reinterpret_cast<const uint64_t*>(bias + c * 32 + 2 * i));
          float t0, t1;
          unpack_f32x2(t2, t0, t1);
          const float p0 = use_poly<POLY>(2 * i) ? exp2_poly3(t0) : fast_exp2(t0);
          const float p1 = use_poly<POLY>(2 * i + 1) ? exp2_poly3(t1) : fast_exp2(t1);
          sum2 = add_f32x2(sum2, pack_f32x2(p0, p1));
          __half2 hh = __floats2half2_rn(p0, p1);
          ph[i] = *reinterpret_cast<uint32_t*>(&hh);
        }
        // keys 32c..32c+31 -> columns 16c..16c+15: never ahead of the columns already read; chunk c+1 (columns 32c+32..)
        // in flight is above them too
        tmem_st_32x16(t_s + c * 16, ph);
      }
      {
        float s0, s1;
        unpack_f32x2(sum2, s0, s1);
        l_run += s0 + s1;
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(b_pfull);
      TC3_STAMP(1, g, 3);
      if (j == nb - 1) {
        // ---- item epilogue: O / l -> fp16 (the next item's first P V needs this warp's next P, so O_x is safe to read)
        mbar_wait(b_pvdone, g & 1);
        TC3_STAMP(1, g, 4);
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
        TC3_STAMP(1, g, 5);
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
typedef CUresult (*PFN_encodeTiled3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled3 encode_fn3() {
  static PFN_encodeTiled3 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled3>(p);
  }
  return fn;
}
static int make_map3c(CUtensorMap* m, const void* base, int batch, int seq, int ld, int box_cols, int box_rows, bool sw128) {
  PFN_encodeTiled3 enc = encode_fn3();
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

template <int POLY>
static int launch_tc3(int grid, size_t smem, cudaStream_t stream, const CUtensorMap& m64, const CUtensorMap& m32,
                      const CUtensorMap& o64, const CUtensorMap& o32, const Attn3Args& a) {
  static PerDeviceMax smem_max;
  if (smem_max.need(smem)) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc3_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return (int)e;
    smem_max.set(smem);
  }
  cudaError_t le = launch_pdl(attention_tc3_kernel<POLY>, dim3(grid), dim3(384), smem, stream, m64, m32, o64, o32, a);
  if (le != cudaSuccess) return (int)le;
  count_launch();
  return (int)cudaGetLastError();
}

int attention_audio_tc3(const void* qkv, const float* mask, void* out, int batch, int seq, int heads, int dh,
                        cudaStream_t stream) {
  using namespace t3;
  if (!qkv || !mask || !out || batch <= 0 || seq <= 0 || heads <= 0 || dh != DH) return CACO_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return CACO_ERR_ALIGN;
  Attn3Args a;
  a.mask = mask; a.out = (__half*)out; a.S = seq; a.H = heads; a.B = batch;
  a.qpairs = (seq + 2 * BM - 1) / (2 * BM);
  a.n_items = batch * heads * a.qpairs;
  a.n_blocks = (seq + BN - 1) / BN;
  a.max_keys = a.n_blocks * BN;
  a.scale_log2 = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
  const size_t smem = OFF_BIAS + 2 * (size_t)a.max_keys * 4;
  if (smem > 232448 || a.n_blocks > 32) return CACO_ERR_ARG;
  const int ld = 3 * heads * dh;
  CUtensorMap m64, m32, o64, o32;
  int rc;
  if ((rc = make_map3c(&m64, qkv, batch, seq, ld, 64, BM, true))) return rc;
  if ((rc = make_map3c(&m32, qkv, batch, seq, ld, 32, BM, false))) return rc;
  if ((rc = make_map3c(&o64, out, batch, seq, heads * dh, 64, 32, true))) return rc;      // per-warp store boxes: 32 rows
  if ((rc = make_map3c(&o32, out, batch, seq, heads * dh, 32, 32, false))) return rc;
  int grid = num_sms();
  if (grid > a.n_items) grid = a.n_items;
  switch (opts().attn_poly) {
    case 1: return launch_tc3<1>(grid, smem, stream, m64, m32, o64, o32, a);
    case 2: return launch_tc3<2>(grid, smem, stream, m64, m32, o64, o32, a);
    case 3: return launch_tc3<3>(grid, smem, stream, m64, m32, o64, o32, a);
    default: return launch_tc3<0>(grid, smem, stream, m64, m32, o64, o32, a);
  }
}

}  // namespace caco

extern "C" int caco_attn3_trace(void* dev_buf) {
  long long* p = (long long*)dev_buf;
  return (int)cudaMemcpyToSymbol(caco::g_attn3_trace, &p, sizeof(p));
}
