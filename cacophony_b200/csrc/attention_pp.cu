// K3a — audio self-attention (mae.py:69-74,89-92), head_dim 96, one persistent CTA per SM, two 128-query tiles in ping-pong.
//
// What the measurements said about the previous kernel (profiles/r02_attention_notes.md):
//   * tcgen05.mma is throughput-, not latency-bound: 65 cycles per M128 x N128 x K16 step (K-major operands), 70 per step for
//     P V with V read MN-major whatever N is — a tile's 14 MMAs per 128-key block cost ~950 cycles + ~300 of commit / wake-up;
//   * the softmax leg (2600 cycles) was bound by tensor-memory LOAD latency: eight 32-column tcgen05.ld per block, one in
//     flight at a time, each several hundred cycles while the tensor pipe is busy; moving exp2 work to the FMA pipe made it slower.
// So here
//   * a softmax thread reads its whole 128-score row ONCE (four tcgen05.ld in flight together) and keeps it in registers
//     for the max and the exp2 pass (setmaxnreg gives the two softmax warpgroups 208 registers, the producer / MMA warpgroup 88),
//   * the passes use packed FFMA2 / FADD2 / FMNMX3 (half the instructions per score),
//   * ONE thread issues every tcgen05.mma and probes the next step's barriers (non-blocking mbarrier.test_wait) between MMA
//     issues, so their ~90-cycle latencies hide under the MMAs in flight.
// Measured and dropped (r02 notes): P published in 32- or 64-key chunks with P V issued under the exp2 pass (every extra
// barrier operation on the issuing thread costs more than the overlap returns), one issuing thread per tile (15 % slower),
// an FMA-pipe exp2 polynomial for 25-50 % of the scores (slower in every kernel version: a tile's exp2 pass is one warp per
// SM sub-partition issuing in order, so pipe times add up instead of overlapping).
// Second pass over the timeline (tile B traced too, wall-clock anchors: 1.67 GHz cold, 1.35-1.40 GHz under sustained load):
//   * an item's epilogue is deferred under the next item's first block (it used to keep a tile's warps away from the next S
//     for ~3300 cycles per 4-block item): O_x stays in tensor memory, is read into registers just before that block's P is
//     published (whose P V overwrites it), scaled / converted / stored by TMA after the publish, and the store's read of the
//     staging buffer is checked one block later;
//   * the row maximum runs as eight independent FMNMX3 chains (one chain of 64: ~375 cycles of latency), blocks with and
//     without masked keys are separate instantiations (a shared body cost the common path 128 register moves per block), the
//     key bias is applied per 32-key chunk that has a masked key, and no thread on the critical path divides by a runtime
//     value any more (integer division goes through the MUFU pipe the exp2 passes saturate);
//   * the exp2 pass is at its floor: scripts/micro/xu_pipe.cu measures 10.3 cycles per MUFU.EX2 for one warp per sub-partition
//     (1320 for a row of 128; the pass with its FFMA2 / FADD2 / F2FP: 1436) — two tiles per SM sub-partition need 2640 of a
//     block's ~3700 cycles; a tile's chain (pass 1436 + load / max / hand-offs ~700 + its MMA leg ~1050) is what remains.
// Outputs are bit-identical to the previous form; 0.333 -> 0.316 ms per layer under sustained load (0.272 -> 0.265 cold).
// Unchanged: Q K^T on 128-key blocks into tensor memory, P written as packed fp16 over the S columns and consumed from tensor
// memory (tcgen05.mma with A in TMEM), V consumed MN-major from its natural layout, lazy rescaling of O (threshold 2^8), K / V
// 2-stage TMA rings, Q and the additive key-mask bias double-buffered per work item (256 queries of one (clip, head)), per-warp
// TMA stores of the normalised output.
//   warps 0 / 3       TMA producers: K + V rings, Q + mask bias (3-D tensor maps over qkv[clip][token][3*768], OOB = 0)
//   warps 1 / 2       tcgen05.mma issuers of tile A / tile B
//   warps 4-7, 8-11   softmax warpgroups A, B (thread = query row)
// Tensor memory: S_A S_B (2 x 128 columns; P_X aliases the first 64 columns of S_X) + O_A O_B (2 x 96) = 448 of 512 columns.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>

#include <type_traits>

#include "caco_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace caco {

namespace pp {
constexpr int BM = 128, BN = 128, DH = 96, NST = 2, NCH = 4;   // the exp2 pass walks a block in NCH chunks of 32 keys
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
                   B_SFULL = 112, B_PVDONE = 128, B_TMEMSLOT = 144, B_PFULL = 160;   // B_PFULL: [tile 2] x 8 B
constexpr int REGS_SOFTMAX = 208, REGS_OTHER = 88;  // 256 x 208 + 128 x 88 = 64512 <= 65536
}  // namespace pp

__device__ long long* g_attn_trace = nullptr;
#define PP_STAMP(role, blk, ev)                                                               \
  do {                                                                                        \
    if (trace != nullptr && (blk) < 64) trace[((role) * 64 + (blk)) * 8 + (ev)] = clock64(); \
  } while (0)

struct AttnArgs {
  const float* mask;
  __half* out;
  int S, H, B;
  int n_items, qpairs, n_blocks, max_keys;
  float scale_log2;
};

// packed fp32 pairs (sm_100: FFMA2 / FADD2 issue two fp32 operations per instruction; FMNMX3 is a three-input maximum)
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
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }

__global__ void __launch_bounds__(384, 1)
attention_pp_kernel(const __grid_constant__ CUtensorMap map_64, const __grid_constant__ CUtensorMap map_32,
                    const __grid_constant__ CUtensorMap map_o64, const __grid_constant__ CUtensorMap map_o32, const AttnArgs a) {
  using namespace pp;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar = sb + OFF_BAR;
  int* s_flag = reinterpret_cast<int*>(smem + OFF_FLAG);
  float* s_bias = reinterpret_cast<float*>(smem + OFF_BIAS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = a.H * DH, nb = a.n_blocks;
  const int n_local = (a.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int total = n_local * nb;
  long long* trace = (blockIdx.x == 0 && lane == 0 && (warp == 1 || warp == 4 || warp == 8)) ? g_attn_trace : nullptr;

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
      mbar_init(bar + B_PVDONE + 8 * i, 1);
      mbar_init(bar + B_PFULL + 8 * i, 4);
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

  if (warp < 4) {
    reg_dealloc<REGS_OTHER>();
    if (warp == 0) {
      // ================================================================ K and V producer (separate 2-stage rings: a K stage is
      // released after both tiles' Q K^T of its block, a V stage only after both tiles' P V)
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
    } else if (warp == 3) {
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
        for (int j0 = 0, blk = 0; j0 < a.max_keys; j0 += BN, ++blk) {
          int chunks = 0;                                      // bit c: the block's 32-key chunk c has a masked key
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            const int j = j0 + c * 32 + lane;
            const bool live = (j < a.S) && (__ldg(a.mask + (size_t)b * a.S + j) != 0.0f);
            s_bias[ib * a.max_keys + j] = live ? 0.0f : -INFINITY;
            if (__any_sync(0xffffffffu, !live)) chunks |= 1 << c;
          }
          if (lane == 0) s_flag[ib * 32 + blk] = chunks;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar + B_BIASFULL + 8 * ib);
      }
    } else if (warp == 1 && lane == 0) {
      // ================================================================ MMA issuer: ONE thread for both tiles.
      // Measured (profiles/r02_attention_notes.md): this thread is the kernel's critical resource — tcgen05.mma issue is
      // synchronous with execution (a shallow queue: ~87 cycles per MMA here) and every mbarrier wait costs ~90 cycles even
      // when the phase completed long ago.  So all barriers of the NEXT step are probed with the non-blocking
      // mbarrier.test_wait between MMA issues (their latency hides under the MMAs in flight) and a blocking wait is taken
      // only if a probe came back negative.  A second issuing thread (one per tile) was measured 15 % slower.
      constexpr uint32_t idesc_qk = umma_idesc_f16(BM, BN);
      constexpr uint32_t idesc_pv = umma_idesc_f16(BM, DH, false, true);
      // (block g = block j of item it: the pairs are counted along with g — a runtime integer division goes through the MUFU
      // pipe, which the softmax warps of this SM sub-partition keep saturated with exp2)
      struct Pos { int it, j; };
      auto next = [&](Pos p) -> Pos { return (p.j + 1 == nb) ? Pos{p.it + 1, 0} : Pos{p.it, p.j + 1}; };
      auto issue_qk = [&](int g, Pos p, int x) {      // S_x = Q_x K_g^T
        const int ib = p.it & 1, st = g % NST;
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
      auto issue_pv = [&](int g, Pos p, int x, int ks0, int ks1) {   // O_x += P_x V_g, k-steps [ks0, ks1): A = P_x from tensor
        const int st = g % NST;                                       // memory (packed fp16, 8 columns per 16 keys), B = V_g MN-major
        const uint32_t acc0 = p.j ? 1u : 0u;
#pragma unroll
        for (int ks = ks0; ks < ks1; ++ks) {
          const uint64_t vb = umma_desc_mnmajor_sw128(sb + OFF_V + st * 32768 + ks * 2048, 16384);
          umma_f16_ts<1>(tmem_base + TM_O + x * DH, tmem_base + TM_S + x * BN + ks * 8, vb, idesc_pv,
                         (acc0 | (uint32_t)ks) ? 1u : 0u);
        }
      };
      // barriers of Q K^T(g): Q (first block of an item) and K
      auto probe_qk_inputs = [&](int g, Pos p) -> bool {
        bool ok = mbar_test(bar + B_KFULL + 8 * (g % NST), (g / NST) & 1);
        if (p.j == 0) ok = mbar_test(bar + B_QFULL + 8 * (p.it & 1), (p.it >> 1) & 1) && ok;
        return ok;
      };
      auto wait_qk_inputs = [&](int g, Pos p) {
        if (p.j == 0) mbar_wait(bar + B_QFULL + 8 * (p.it & 1), (p.it >> 1) & 1);
        mbar_wait(bar + B_KFULL + 8 * (g % NST), (g / NST) & 1);
      };
      bool ok_v = false, ok_qk = false, ok_pa = false, ok_pb = false;
      Pos p0{0, 0};
      if (total > 0) {
        wait_qk_inputs(0, p0);
        tc_fence_after();
        issue_qk(0, p0, 0);
        issue_qk(0, p0, 1);
      }
      for (int g = 0; g < total; ++g) {
        PP_STAMP(0, g, 0);
        if (trace != nullptr && (g == 0 || g == 63)) {         // wall-clock anchors: SM clock = cycles / ns between them
          unsigned long long ns;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
          trace[(0 * 64 + g) * 8 + 5] = (long long)ns;
        }
        const int st = g % NST;
        const bool more = g + 1 < total;
        const Pos p1 = next(p0), p2 = next(p1);
        if (!ok_v) mbar_wait(bar + B_VFULL + 8 * st, (g / NST) & 1);
        if (more && !ok_qk) wait_qk_inputs(g + 1, p1);
        if (!ok_pa) mbar_wait(bar + B_PFULL + 8 * 0, g & 1);
        PP_STAMP(0, g, 1);
        tc_fence_after();
        // ---- tile A
        issue_pv(g, p0, 0, 0, 4);
        ok_pb = mbar_test(bar + B_PFULL + 8 * 1, g & 1);                   // tile B's P of this block
        issue_pv(g, p0, 0, 4, 8);
        umma_commit<1>(bar + B_PVDONE + 8 * 0);
        if (more) issue_qk(g + 1, p1, 0);              // overwrites S_A / P_A: ordered after P_A V by the in-order pipe
        PP_STAMP(0, g, 2);
        if (!ok_pb) mbar_wait(bar + B_PFULL + 8 * 1, g & 1);
        PP_STAMP(0, g, 3);
        tc_fence_after();
        // ---- tile B, with the probes for block g + 1 between its MMAs
        issue_pv(g, p0, 1, 0, 4);
        ok_v = more && mbar_test(bar + B_VFULL + 8 * ((g + 1) % NST), ((g + 1) / NST) & 1);
        ok_qk = (g + 2 < total) && probe_qk_inputs(g + 2, p2);
        issue_pv(g, p0, 1, 4, 8);
        umma_commit<1>(bar + B_PVDONE + 8 * 1);
        umma_commit<1>(bar + B_VEMPTY + 8 * st);
        if (more) {
          issue_qk(g + 1, p1, 1);
          ok_pa = mbar_test(bar + B_PFULL + 8 * 0, (g + 1) & 1);            // tile A's P of the next block
        }
        PP_STAMP(0, g, 4);
        p0 = p1;
      }
    }
  } else {
    // ================================================================ softmax warpgroups (thread = query row)
    reg_alloc<REGS_SOFTMAX>();
    const int x = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t t_s = tmem_base + (uint32_t(quarter * 32) << 16) + TM_S + x * BN;
    const uint32_t t_o = tmem_base + (uint32_t(quarter * 32) << 16) + TM_O + x * DH;
    const uint32_t b_sfull = bar + B_SFULL + 8 * x, b_pfull = bar + B_PFULL + 8 * x, b_pvdone = bar + B_PVDONE + 8 * x;
    float m_ref = -INFINITY, l_run = 0.f;
    int b = 0, h = 0, q0 = 0;
    // Item epilogue, deferred (measured on the CTA-0 timeline: done right after an item's last block it kept the tile's warps
    // away from the next item's first S for ~3300 cycles — waiting for the last P V, 1600 cycles of O -> fp16 -> smem -> TMA
    // store — while the issuing thread idled; an item is only 4 blocks of ~3200).  Now the finished item's O_x stays in tensor
    // memory while the warps run the NEXT item's first block; its P V may not be issued before O_x has been read (it
    // overwrites O_x), so the rows are read into registers just before that block's P is published, and scaling, conversion and
    // the TMA store happen after the publish, under the issuing thread's MMAs.
    bool pend = false;                 // a finished item of this tile still has its O_x in tensor memory
    float pend_l = 0.f;
    int pend_ib = 0, pend_b = 0, pend_h = 0, pend_q0 = 0, pend_par = 0;
    auto drain_o = [&](uint32_t (&o)[6][16], int par) {       // O_x rows -> registers, once the item's last P V has retired
      mbar_wait(b_pvdone, par);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 6; ++c) tmem_ld_32x16(t_o + c * 16, o[c]);
      tmem_ld_wait();
    };
    auto store_item = [&](const uint32_t (&o)[6][16], float l, int ib, int bb, int hh, int qq0) {
      const float inv = 1.0f / l;                              // l == 0 (no live key): NaN row, like torch.softmax
      // O rows / l -> fp16 -> the item's (dead) Q_x buffer in the swizzled layouts of the two output tensor maps -> one TMA
      // store per warp (32 rows x 64 + 32 columns); rows past the clip's end are clipped by the hardware
      uint8_t* stage = smem + OFF_Q + (ib * 2 + x) * Q_TILE;
      uint8_t* r0 = stage + row * 128;
      uint8_t* r1 = stage + 16384 + row * 64;
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __half2 hf = __floats2half2_rn(__uint_as_float(o[c][2 * i]) * inv, __uint_as_float(o[c][2 * i + 1]) * inv);
          pk[i] = *reinterpret_cast<uint32_t*>(&hf);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int ch = c * 2 + k;                            // 16-byte chunk (8 fp16) 0..11 of the 96-column row
          const uint4 u = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
          if (ch < 8) *reinterpret_cast<uint4*>(r0 + ((ch ^ (row & 7)) << 4)) = u;                       // SWIZZLE_128B
          else *reinterpret_cast<uint4*>(r1 + (((ch - 8) ^ ((row >> 1) & 3)) << 4)) = u;                 // SWIZZLE_64B
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const int qrow = qq0 + x * BM + quarter * 32;
        tma_store_3d(&map_o64, sb + OFF_Q + (ib * 2 + x) * Q_TILE + quarter * 32 * 128, hh * DH, qrow, bb);
        tma_store_3d(&map_o32, sb + OFF_Q + (ib * 2 + x) * Q_TILE + 16384 + quarter * 32 * 64, hh * DH + 64, qrow, bb);
        tma_store_commit();
      }
      __syncwarp();
    };
    // the store's read of the staging buffer is not waited for in line (~1000 cycles on the timeline): the issuing lane checks
    // it one block later, when it has long completed, and only then hands the buffer back to the Q producer
    int store_ib = -1;
    auto finish_store = [&]() {
      if (store_ib >= 0) {
        if (lane == 0) {
          tma_store_wait_read<0>();                            // smem may be refilled with the next-but-one item's Q
          mbar_arrive(bar + B_ITEMDONE + 8 * store_ib);
        }
        __syncwarp();
        store_ib = -1;
      }
    };
    int it = 0, j = 0;                 // item and block within it (counted, not divided: integer division runs on the MUFU pipe
                                       // the other tile's exp2 pass is saturating)
    for (int g = 0; g < total; ++g) {
      const int ib = it & 1;
      if (j == 0) {
        decode(it, b, h, q0);
        mbar_wait(bar + B_BIASFULL + 8 * ib, (it >> 1) & 1);
        m_ref = -INFINITY;
        l_run = 0.f;
      }
      const int mchunks = s_flag[ib * 32 + j];                // warp-uniform: which 32-key chunks of the block have masked keys
      const bool masked = mchunks != 0;
      const float* bias = s_bias + ib * a.max_keys + j * BN;
      PP_STAMP(1 + x, g, 0);
      mbar_wait(b_sfull, g & 1);                              // S_x(g) complete => every earlier MMA (incl. P_x V of g-1) retired
      PP_STAMP(1 + x, g, 1);
      tc_fence_after();
      // The block's softmax, instantiated separately for blocks with and without masked keys (only a clip's last block has any):
      // with one body and an in-place "s += bias" on the masked path the two paths' register assignments of s[] had to be
      // merged, which cost the common unmasked path 128 register moves per block.  Leaves P_x stored (not yet published).
      auto softmax_block = [&](auto masked_c) {
        constexpr bool MASKED = decltype(masked_c)::value;
        // ---- the whole row of scores, once: four loads in flight together
        uint32_t s[BN];
#pragma unroll
        for (int c = 0; c < NCH; ++c) tmem_ld_32x32(t_s + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
        tmem_ld_wait();
        PP_STAMP(1 + x, g, 2);
        if (MASKED) {                                           // key bias (0 / -inf) onto the chunks that have masked keys
#pragma unroll
          for (int c = 0; c < NCH; ++c)
            if ((mchunks >> c) & 1) {
#pragma unroll
              for (int i = c * 16; i < c * 16 + 16; ++i) {
                float s0, s1;
                unpack_f32x2(add_f32x2(pack_f32x2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])),
                                       *reinterpret_cast<const uint64_t*>(bias + 2 * i)), s0, s1);
                s[2 * i] = __float_as_uint(s0);
                s[2 * i + 1] = __float_as_uint(s1);
              }
            }
        }
        // ---- row maximum of the raw scores (scale > 0, so max commutes with the scaling): eight independent FMNMX3 chains (one
        // chain of 64 dependent maxima was ~375 cycles of pure ALU latency on the tile's critical path); the maximum is exact, so
        // the grouping does not change the result
        float mx8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mx8[i] = fmaxf(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]));
#pragma unroll
        for (int i = 8; i < BN / 2; ++i) mx8[i & 7] = max3(mx8[i & 7], __uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]));
        float mx = fmaxf(max3(mx8[0], mx8[1], mx8[2]), max3(mx8[3], mx8[4], mx8[5]));
        mx = max3(mx, mx8[6], mx8[7]);
        mx *= a.scale_log2;                                     // -inf stays -inf
        // ---- lazy rescale of O_x and l (only when the maximum grew by more than 2^8 since the reference was taken)
        const bool need = mx > m_ref + RESCALE_T;               // m_ref == -inf: true iff this block has a live key
        if (j == 0) {                                           // first block of an item: no O_x, l == 0, the reference
          m_ref = mx;                                           // maximum is this block's (what the general path reduces to)
        } else if (__any_sync(0xffffffffu, need)) {
          const float factor = need ? exp2f(m_ref - mx) : 1.0f;
          {
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
        PP_STAMP(1 + x, g, 3);
        // ---- p = exp2(s*scale - m) per PAIR of scores: one FFMA2, two MUFU.EX2, one FADD2 (two running sums), one F2FP; a chunk
        // of 32 keys is written as packed fp16 over S columns 16c..16c+15
        // The pass is software-pipelined by hand in groups of four pairs (the eight exponentials of group q+1 are issued before
        // the results of group q are consumed; volatile asm keeps that order).  scripts/micro/xu_pipe.cu: one warp per
        // sub-partition issues a MUFU.EX2 every ~10.3 cycles (128 of them: 1320 cycles), F2FP goes to another pipe at ~7 cycles,
        // and the pass in exactly this form takes 1436 cycles — what the timeline shows for it here (1450-1650 when the other
        // tile's pass overlaps), i.e. the pass is at the MUFU floor; moving a third of the pairs to an FMA-pipe polynomial makes
        // it 1622.
        uint64_t sum2 = pack_f32x2(0.f, 0.f);
        const uint64_t scale2 = pack_f32x2(a.scale_log2, a.scale_log2), negm2 = pack_f32x2(neg_m, neg_m);
        auto ex2_group = [&](int q) {                          // s[8q .. 8q+7] <- exp2(s * scale - m)
#pragma unroll
          for (int i = 4 * q; i < 4 * q + 4; ++i) {
            const uint64_t t2 = fma_f32x2(pack_f32x2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])), scale2, negm2);
            float t0, t1, p0, p1;
            unpack_f32x2(t2, t0, t1);
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(t0));
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(t1));
            s[2 * i] = __float_as_uint(p0);
            s[2 * i + 1] = __float_as_uint(p1);
          }
        };
        uint32_t ph[16];
        auto use_group = [&](int q) {                          // row sum and packed fp16 of group q's probabilities
#pragma unroll
          for (int i = 4 * q; i < 4 * q + 4; ++i) {
            const float p0 = __uint_as_float(s[2 * i]), p1 = __uint_as_float(s[2 * i + 1]);
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(sum2) : "l"(pack_f32x2(p0, p1)));
            asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(ph[i & 15]) : "f"(p1), "f"(p0));
          }
        };
        ex2_group(0);
#pragma unroll
        for (int q = 0; q < BN / 8; ++q) {
          if (q + 1 < BN / 8) ex2_group(q + 1);
          use_group(q);
          if ((q & 3) == 3) tmem_st_32x16(t_s + (q >> 2) * 16, ph);
        }
        float s0, s1;
        unpack_f32x2(sum2, s0, s1);
        l_run += s0 + s1;
        tmem_st_wait();
      };
      if (masked) softmax_block(std::true_type{});
      else softmax_block(std::false_type{});
      if (pend) {
        // the previous item's O_x: out of tensor memory before this block's P V (which overwrites it) can be issued
        uint32_t o[6][16];
        drain_o(o, pend_par);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_pfull);
        PP_STAMP(1 + x, g, 4);
        store_item(o, pend_l, pend_ib, pend_b, pend_h, pend_q0);
        store_ib = pend_ib;
        PP_STAMP(1 + x, g, 6);
        pend = false;
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_pfull);
        PP_STAMP(1 + x, g, 4);
        finish_store();
      }
      if (j == nb - 1) {
        if (g + 1 < total) {           // defer: O_x is complete once P_x V of this block retires (b_pvdone, phase g)
          finish_store();              // (single-block items: the previous store has had this block's time)
          pend = true;
          pend_l = l_run; pend_ib = ib; pend_b = b; pend_h = h; pend_q0 = q0; pend_par = g & 1;
        } else {                       // the CTA's last item
          finish_store();
          uint32_t o[6][16];
          drain_o(o, g & 1);
          PP_STAMP(1 + x, g, 5);
          store_item(o, l_run, ib, b, h, q0);
          store_ib = ib;
          finish_store();
        }
        j = 0;
        ++it;
      } else {
        ++j;
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

int attention_audio_pp(const void* qkv, const float* mask, void* out, int batch, int seq, int heads, int dh,
                       cudaStream_t stream) {
  using namespace pp;
  if (!qkv || !mask || !out || batch <= 0 || seq <= 0 || heads <= 0 || dh != DH) return CACO_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return CACO_ERR_ALIGN;
  AttnArgs a;
  a.mask = mask; a.out = (__half*)out; a.S = seq; a.H = heads; a.B = batch;
  a.qpairs = (seq + 2 * BM - 1) / (2 * BM);
  a.n_items = batch * heads * a.qpairs;
  a.n_blocks = (seq + BN - 1) / BN;
  a.max_keys = a.n_blocks * BN;
  a.scale_log2 = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
  const size_t smem = OFF_BIAS + 2 * (size_t)a.max_keys * 4;
  if (smem > 232448 || a.n_blocks > 32) return CACO_ERR_ARG;   // at most 4096 keys (per-item key bias in shared memory)
  const int ld = 3 * heads * dh;
  CUtensorMap m64, m32, o64, o32;
  int rc;
  if ((rc = make_map3c(&m64, qkv, batch, seq, ld, 64, BM, true))) return rc;
  if ((rc = make_map3c(&m32, qkv, batch, seq, ld, 32, BM, false))) return rc;
  if ((rc = make_map3c(&o64, out, batch, seq, heads * dh, 64, 32, true))) return rc;      // per-warp store boxes: 32 rows
  if ((rc = make_map3c(&o32, out, batch, seq, heads * dh, 32, 32, false))) return rc;
  int grid = num_sms();
  if (grid > a.n_items) grid = a.n_items;
  static PerDeviceMax smem_max;
  if (smem_max.need(smem)) {
    cudaError_t e = cudaFuncSetAttribute(attention_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e) return (int)e;
    smem_max.set(smem);
  }
  cudaError_t le = launch_pdl(attention_pp_kernel, dim3(grid), dim3(384), smem, stream, m64, m32, o64, o32, a);
  if (le != cudaSuccess) return (int)le;
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace caco

extern "C" int caco_attn_trace(void* dev_buf) {
  long long* p = (long long*)dev_buf;
  return (int)cudaMemcpyToSymbol(caco::g_attn_trace, &p, sizeof(p));
}
