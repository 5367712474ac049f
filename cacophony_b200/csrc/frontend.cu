// K1 — waveform -> |STFT| -> HTK mel -> log -> 16x16 patches (+ time/freq indices and mask).
// Replaces compute_mel_spectrogram + spectrogram_to_patches + prepare_audio_batch
// (src/eval/eval_caco_torch.py:41-105, :108-151, :181-206), batched and without the host numpy hop.
//
// One CTA per (clip, 16-frame group) = one patch row t: it stages the 2912 samples the 16 frames touch with TMA (twelve
// 256-sample tiles of a 2-D tensor map over wave[batch][samples]; samples past the clip's end are zero-filled by the
// hardware, which IS the reference's tail padding) and the 9.3 KB of constant tables with one bulk copy, all completing on
// one mbarrier; unaligned or ragged-with-garbage inputs take a cp.async path.  Then 16 real 512-point FFTs (each a 256-point complex FFT done as two radix-16 passes in the
// registers of 16 threads, one padded shared-memory exchange between them, fp32, host-computed twiddles) + real split,
// applies the sparse HTK filterbank (505 non-zeros), log(x+1e-5)*0.2+0.9, and writes the 8 patches of that row — 8 KB
// contiguous in the [B, max_patches, 256] layout — with coalesced 128-byte stores.  Algorithmic HBM traffic: 1.158 MB per
// 10 s clip.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <mutex>
#include <vector>

#include "caco_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace caco {

constexpr int FE_HOP = 160, FE_WIN = 400, FE_NFFT = 512, FE_NMEL = 128, FE_NFREQ = 257;
constexpr int FE_WOFF = (FE_NFFT - FE_WIN) / 2;  // torch.stft centres the 400-tap window in the 512 frame
constexpr int FE_MAXW = 16;                      // max non-zeros of one mel filter
constexpr int FE_FRAMES = 16;                    // frames per CTA (= patch height)
constexpr int FE_SAMPLES = (FE_FRAMES - 1) * FE_HOP + FE_NFFT;  // 2912
constexpr int FE_MEL_LD = 144;                   // padded row of the staged log-mel tile

// Constant tables, packed so that every CTA pulls them into shared memory with 16-byte cp.async copies (9.3 KB):
struct FeTables {
  float2 w256t[256];        // [k1][n2] = e^{-2 pi i n2 k1 / 256}: twiddles between the two radix-16 passes
  float2 w512[256];         // e^{-2 pi i k / 512}, k < 256: real-split twiddles (computed in double on the host)
  float win[FE_WIN];        // periodic Hann(400)
  int mel_start[FE_NMEL];   // first FFT bin of each HTK filter
  int mel_count[FE_NMEL];   // number of non-zero taps (<= FE_MAXW)
  int mel_woff[FE_NMEL];    // offset of the filter's taps in mel_w
  float mel_w[512];         // the 505 non-zero filterbank weights, filter after filter
};
static_assert(sizeof(FeTables) % 16 == 0, "tables are copied in 16-byte pieces");
__device__ __align__(16) FeTables g_fe;

// torch.linspace(start, end, steps) in fp32: symmetric evaluation from both ends.
static void linspace_f32(float start, float end, int steps, std::vector<float>& out) {
  out.resize(steps);
  const float step = (end - start) / static_cast<float>(steps - 1);
  const int half = steps / 2;
  for (int i = 0; i < steps; ++i)
    out[i] = (i < half) ? (start + step * static_cast<float>(i)) : (end - step * static_cast<float>(steps - i - 1));
}

// torchaudio.functional.melscale_fbanks(257, 0, 8000, 128, 16000, norm=None, mel_scale="htk") in fp32
// (called at eval_caco_torch.py:94-101).  out: [257][128].
void mel_filterbank_host(float* out) {
  std::vector<float> all_freqs, m_pts;
  linspace_f32(0.0f, 8000.0f, FE_NFREQ, all_freqs);
  const double m_min = 2595.0 * log10(1.0 + 0.0 / 700.0);
  const double m_max = 2595.0 * log10(1.0 + 8000.0 / 700.0);
  linspace_f32(static_cast<float>(m_min), static_cast<float>(m_max), FE_NMEL + 2, m_pts);
  std::vector<float> f_pts(FE_NMEL + 2);
  for (int i = 0; i < FE_NMEL + 2; ++i) f_pts[i] = 700.0f * (powf(10.0f, m_pts[i] / 2595.0f) - 1.0f);
  for (int k = 0; k < FE_NFREQ; ++k)
    for (int m = 0; m < FE_NMEL; ++m) {
      const float down = (-1.0f * (f_pts[m] - all_freqs[k])) / (f_pts[m + 1] - f_pts[m]);
      const float up = (f_pts[m + 2] - all_freqs[k]) / (f_pts[m + 2] - f_pts[m + 1]);
      const float v = fminf(down, up);
      out[k * FE_NMEL + m] = v > 0.0f ? v : 0.0f;
    }
}

static int frontend_init() {
  // host tables once per process, the __device__ copy once per device
  static FeTables t;
  static int host_rc = -100;
  static std::once_flag once;
  std::call_once(once, [] {
    std::vector<float> fb(FE_NFREQ * FE_NMEL);
    mel_filterbank_host(fb.data());
    host_rc = 0;
    int woff = 0;
    for (int m = 0; m < FE_NMEL; ++m) {
      int lo = -1, hi = -1;
      for (int k = 0; k < FE_NFREQ; ++k)
        if (fb[k * FE_NMEL + m] != 0.0f) { if (lo < 0) lo = k; hi = k; }
      t.mel_start[m] = lo < 0 ? 0 : lo;
      t.mel_count[m] = lo < 0 ? 0 : hi - lo + 1;
      t.mel_woff[m] = woff;
      if (t.mel_count[m] > FE_MAXW || woff + t.mel_count[m] > 512) { host_rc = CACO_ERR_STATE; return; }
      for (int j = 0; j < t.mel_count[m]; ++j) t.mel_w[woff++] = fb[(t.mel_start[m] + j) * FE_NMEL + m];
    }
    for (; woff < 512; ++woff) t.mel_w[woff] = 0.0f;
    for (int k = 0; k < 256; ++k) {
      const double ang = -2.0 * M_PI * (double)k / (double)FE_NFFT;
      t.w512[k] = make_float2((float)cos(ang), (float)sin(ang));
    }
    for (int k1 = 0; k1 < 16; ++k1)
      for (int n2 = 0; n2 < 16; ++n2) {
        const double ang = -2.0 * M_PI * (double)((n2 * k1) % 256) / 256.0;
        t.w256t[k1 * 16 + n2] = make_float2((float)cos(ang), (float)sin(ang));
      }
    for (int i = 0; i < FE_WIN; ++i) t.win[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)i / (double)FE_WIN));   // torch.hann_window(400)
  });
  if (host_rc) return host_rc;
  static PerDeviceOnce uploaded;
  if (uploaded.first()) {
    cudaError_t e = cudaMemcpyToSymbol(g_fe, &t, sizeof(t));
    if (e != cudaSuccess) return (int)e;
    uploaded.done();
  }
  return 0;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
// 4-byte copy, zero-filled when !valid (src-size 0)
__device__ __forceinline__ void cp_async4_zfill(void* smem_dst, const void* gsrc, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// single-instruction MUFU forms: sqrt.approx and lg2.approx are accurate to ~1 ulp / 2^-22 absolute, three orders of magnitude
// inside the log-mel tolerance (tests/util.py), and save ~17 instructions per magnitude / log over sqrtf / logf — the kernel is
// instruction-issue-bound (profiles/r02_notes.md)
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_ln(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * 0.69314718055994530942f;
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_negi(float2 a) { return make_float2(a.y, -a.x); }   // -i * a

// 4-point forward DFT in place: (a, b, c, d) -> (X0, X1, X2, X3)
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = mul_negi(csub(b, d));
  a = cadd(s0, s2);
  b = cadd(s1, s3);
  c = csub(s0, s2);
  d = csub(s1, s3);
}

// 16-point forward DFT of v[0..15] in registers, natural order in and out: n = 4a + b, k = c + 4d,
// X[c + 4d] = sum_b W16^{bc} W4^{bd} sum_a v[4a + b] W4^{ac}  (two radix-4 stages, compile-time twiddles)
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  constexpr float C1 = 0.92387953251128673848f, S1 = 0.38268343236508978178f, R2 = 0.70710678118654752440f;
#pragma unroll
  for (int b = 0; b < 4; ++b) dft4(v[b], v[4 + b], v[8 + b], v[12 + b]);      // v[4c + b] = u_b[c]
  // u_b[c] *= W16^{bc}
  v[4 + 1] = cmul(v[4 + 1], make_float2(C1, -S1));     // b=1,c=1: W^1
  v[8 + 1] = cmul(v[8 + 1], make_float2(R2, -R2));     // b=1,c=2: W^2
  v[12 + 1] = cmul(v[12 + 1], make_float2(S1, -C1));   // b=1,c=3: W^3
  v[4 + 2] = cmul(v[4 + 2], make_float2(R2, -R2));     // b=2,c=1: W^2
  v[8 + 2] = mul_negi(v[8 + 2]);                       // b=2,c=2: W^4 = -i
  v[12 + 2] = cmul(v[12 + 2], make_float2(-R2, -R2));  // b=2,c=3: W^6
  v[4 + 3] = cmul(v[4 + 3], make_float2(S1, -C1));     // b=3,c=1: W^3
  v[8 + 3] = cmul(v[8 + 3], make_float2(-R2, -R2));    // b=3,c=2: W^6
  v[12 + 3] = cmul(v[12 + 3], make_float2(-C1, S1));   // b=3,c=3: W^9
#pragma unroll
  for (int c = 0; c < 4; ++c) dft4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);   // v[4c + d] = X[c + 4d]
  // transpose the 4x4 index so that v[k] = X[k]
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int d = c + 1; d < 4; ++d) {
      const float2 t = v[4 * c + d];
      v[4 * c + d] = v[4 * d + c];
      v[4 * d + c] = t;
    }
}

constexpr int FE_ZLD = 17;                               // padded row of the 16x16 exchange buffer (float2 units)
constexpr int FE_ZFRAME = 16 * FE_ZLD;                   // 272 float2 per frame (>= 257 floats for the magnitudes)
constexpr int FE_TMA_BOX = 256;                          // samples per TMA tile (the box limit of a tensor-map dimension)
constexpr int FE_TMA_TILES = (FE_SAMPLES + FE_TMA_BOX - 1) / FE_TMA_BOX;   // 12 tiles = 3072 samples staged (2912 used)
constexpr int FE_OFF_Z = FE_TMA_TILES * FE_TMA_BOX * 4;                    // 12288 (128-byte aligned: TMA destination)
constexpr int FE_OFF_TAB = FE_OFF_Z + FE_FRAMES * FE_ZFRAME * 8;           // 46464
constexpr int FE_OFF_MBAR = FE_OFF_TAB + (int)sizeof(FeTables);
constexpr int FE_SMEM_BYTES = FE_OFF_MBAR + 16;                            // 56.4 KB -> four CTAs per SM
static_assert(FE_OFF_Z % 128 == 0 && FE_OFF_TAB % 16 == 0 && FE_OFF_MBAR % 8 == 0, "TMA / cp.async / mbarrier alignment");

// 1-D bulk copy global -> shared, completing on an mbarrier (the constant tables)
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}

// 16 threads per frame, 16 frames per CTA.  The 512-point real FFT of a frame is a 256-point complex FFT of
// z[n] = x[2n] + i x[2n+1], done as 16 x 16: thread n2 transforms z[16 n1 + n2] over n1 in registers, multiplies by
// W256^{n2 k1}, the frame's 16 threads exchange through one padded (conflict-free) shared-memory tile, thread k1 transforms
// over n2 and owns X[k1 + 16 k2]; the conjugate partner Z[256 - k] for the real split comes from a second pass through the
// same tile.  Two shared-memory round trips per frame instead of the eight of a radix-4 Stockham, and only warp-level
// synchronisation (a frame lives in half a warp).
// USE_TMA: samples staged by tensor-map tile loads (needs a 16-byte aligned base and row pitch); MASK_TAIL: samples past a
// ragged clip's own length are forced to zero at the window stage (the tensor map only knows the common row length).
template <bool USE_TMA, bool MASK_TAIL>
__global__ void __launch_bounds__(256)
frontend_kernel(const __grid_constant__ CUtensorMap map_wave, const float* __restrict__ wave, const int* __restrict__ lengths,
                int stride, int n_samples_u, int max_patches, float* __restrict__ patches, __half* __restrict__ patches_f16,
                float* __restrict__ time_inds, float* __restrict__ freq_inds, float* __restrict__ mask,
                float* __restrict__ log_mel) {
  extern __shared__ __align__(128) uint8_t fe_smem[];
  float* s_x = reinterpret_cast<float*>(fe_smem);                                   // 2912 samples; later the log-mel tile
  float2* s_z = reinterpret_cast<float2*>(fe_smem + FE_OFF_Z);                      // [16 frames][16][17]
  const FeTables& tab = *reinterpret_cast<const FeTables*>(fe_smem + FE_OFF_TAB);
  float (*s_out)[FE_MEL_LD] = reinterpret_cast<float (*)[FE_MEL_LD]>(s_x);          // [16][144] = 2304 floats <= 2912

  const int t = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  // ragged batches: clip b holds lengths[b] samples in a row of `stride` floats (uniform batches: lengths == nullptr)
  int n_samples = n_samples_u;
  if (lengths != nullptr) {
    n_samples = __ldg(lengths + b);
    n_samples = n_samples < 0 ? 0 : (n_samples > stride ? stride : n_samples);
  }
  const int n_frames = (n_samples + FE_HOP - 1) / FE_HOP;  // eval_caco_torch.py:67
  const int nt_valid = n_frames / 16;                      // eval_caco_torch.py:116-117
  const int frame0 = t * FE_FRAMES;
  const int T_out = (max_patches + 7) / 8;
  const int n_valid_tokens = nt_valid * 8;
  const bool have_frames = frame0 < n_frames;
  const bool want_patch = (t < nt_valid) && (t < T_out);
  const bool compute = have_frames && (want_patch || log_mel != nullptr);

  // indices + mask for this row's 8 tokens (eval_caco_torch.py:132-144: padded slots carry 0)
  if (t < T_out && tid < 8) {
    const int p = t * 8 + tid;
    if (p < max_patches) {
      const bool live = p < n_valid_tokens;
      const size_t o = (size_t)b * max_patches + p;
      mask[o] = live ? 1.0f : 0.0f;
      time_inds[o] = live ? (float)t : 0.0f;
      freq_inds[o] = live ? (float)tid : 0.0f;
    }
  }
  if (!compute) {
    if (t < T_out) {  // zero padding rows
      for (int i = tid; i < 2048; i += 256) {
        const int p = t * 8 + (i >> 8);
        if (p < max_patches) {
          const size_t o = ((size_t)b * max_patches + p) * 256 + (i & 255);
          if (patches) patches[o] = 0.0f;
          if (patches_f16) patches_f16[o] = __float2half_rn(0.0f);
        }
      }
    }
    return;
  }

  // ---- stage the tables and the samples (zero tail pad, eval_caco_torch.py:72-78)
  if constexpr (USE_TMA) {
    // one thread: 12 tile loads of 256 samples (coordinates past the row's end are zero-filled by the TMA unit) + one bulk
    // copy of the tables, 21.6 KB in flight on one mbarrier; nobody else issues a load instruction
    const uint32_t bar = smem_u32(fe_smem + FE_OFF_MBAR);
    if (tid == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
      mbar_expect_tx(bar, FE_TMA_TILES * FE_TMA_BOX * 4 + (uint32_t)sizeof(FeTables));
      bulk_copy_g2s(smem_u32(fe_smem + FE_OFF_TAB), &g_fe, (uint32_t)sizeof(FeTables), bar);
      const int s0 = frame0 * FE_HOP;
#pragma unroll
      for (int i = 0; i < FE_TMA_TILES; ++i)
        tma_load_2d(smem_u32(fe_smem) + i * FE_TMA_BOX * 4, &map_wave, bar, s0 + i * FE_TMA_BOX, b);
    }
    __syncthreads();                      // the barrier's init is visible to every waiter
    mbar_wait(bar, 0);
  } else {
    // cp.async: every copy of the CTA is in flight at once instead of one exposed global-load latency per loop trip
    const uint8_t* src = reinterpret_cast<const uint8_t*>(&g_fe);
    uint8_t* dst = fe_smem + FE_OFF_TAB;
    for (int i = tid; i < (int)sizeof(FeTables) / 16; i += 256) cp_async16(dst + 16 * i, src + 16 * i);
    const float* wv = wave + (size_t)b * stride;
    const int s0 = frame0 * FE_HOP;
    const bool aligned = ((reinterpret_cast<uintptr_t>(wv + s0) & 15) == 0);
    for (int i = tid; i < FE_SAMPLES / 4; i += 256) {
      const int s = s0 + 4 * i;
      if (aligned && s + 3 < n_samples) {
        cp_async16(s_x + 4 * i, wv + s);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) cp_async4_zfill(s_x + 4 * i + e, wv + (s + e < n_samples ? s + e : 0), s + e < n_samples);
      }
    }
    cp_async_wait_all();
    __syncthreads();
  }

  const int fr = tid >> 4;       // frame within the CTA
  const int q = tid & 15;        // n2 in pass 1, k1 in pass 2
  float2* zf = s_z + fr * FE_ZFRAME;
  float2 v[16];
  {
    // z[16 n1 + q] = (x[2n] w[2n], x[2n+1] w[2n+1]), n = 16 n1 + q; window support is [56, 456) of the 512-sample frame
    const float* xf = s_x + fr * FE_HOP;
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const int i0 = 2 * (16 * n1 + q), i1 = i0 + 1;
      const float w0 = (i0 >= FE_WOFF && i0 < FE_WOFF + FE_WIN) ? tab.win[i0 - FE_WOFF] : 0.0f;
      const float w1 = (i1 >= FE_WOFF && i1 < FE_WOFF + FE_WIN) ? tab.win[i1 - FE_WOFF] : 0.0f;
      float2 xx = *reinterpret_cast<const float2*>(xf + i0);
      if constexpr (MASK_TAIL) {           // ragged clip shorter than the staged row: its tail is zero, whatever the buffer holds
        const int g0 = (frame0 + fr) * FE_HOP + i0;
        if (g0 >= n_samples) xx.x = 0.0f;
        if (g0 + 1 >= n_samples) xx.y = 0.0f;
      }
      v[n1] = make_float2(xx.x * w0, xx.y * w1);
    }
  }
  dft16(v);                                                 // over n1 -> index k1
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) v[k1] = cmul(v[k1], tab.w256t[k1 * 16 + q]);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) zf[k1 * FE_ZLD + q] = v[k1];
  __syncwarp();
#pragma unroll
  for (int n2 = 0; n2 < 16; ++n2) v[n2] = zf[q * FE_ZLD + n2];
  dft16(v);                                                 // over n2 -> v[k2] = Z[q + 16 k2]
  __syncwarp();
  // second exchange: natural order Z[k] so that every thread can fetch the conjugate partners Z[256 - k]
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) zf[q + 16 * k2] = v[k2];
  __syncwarp();
  float mag[16];
  float mag_nyq = 0.0f;
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) {
    // real split: X[k] = E + W512^k O,  E = (Z[k] + conj Z[256-k]) / 2,  O = -i (Z[k] - conj Z[256-k]) / 2
    const int k = q + 16 * k2;
    const float2 zk = v[k2];
    const float2 zc = zf[(256 - k) & 255];
    const float er = 0.5f * (zk.x + zc.x), ei = 0.5f * (zk.y - zc.y);
    const float dr = zk.x - zc.x, di = zk.y + zc.y;
    const float orr = 0.5f * di, oi = -0.5f * dr;
    const float2 w = tab.w512[k];
    const float xr = er + (w.x * orr - w.y * oi);
    const float xi = ei + (w.x * oi + w.y * orr);
    mag[k2] = fast_sqrt(xr * xr + xi * xi);
    if (k == 0) {                                           // X[256] = E - O at k = 0 (W512^256 = -1)
      const float nr = er - orr, ni = ei - oi;
      mag_nyq = fast_sqrt(nr * nr + ni * ni);
    }
  }
  __syncwarp();                                             // every partner read is done: the tile becomes the magnitudes
  float* s_mag = reinterpret_cast<float*>(zf);              // [257]
#pragma unroll
  for (int k2 = 0; k2 < 16; ++k2) s_mag[q + 16 * k2] = mag[k2];
  if (q == 0) s_mag[256] = mag_nyq;
  __syncthreads();                                          // all frames have consumed s_x: it becomes the log-mel tile
  // ---- sparse mel + log (eval_caco_torch.py:103-104): 8 filters per thread
#pragma unroll
  for (int i = 0; i < FE_NMEL / 16; ++i) {
    const int m = q + 16 * i;
    const int st = tab.mel_start[m], cnt = tab.mel_count[m];
    const float* mw = tab.mel_w + tab.mel_woff[m];
    float acc = 0.0f;
    for (int c = 0; c < cnt; ++c) acc = fmaf(s_mag[st + c], mw[c], acc);
    s_out[fr][m] = fast_ln(acc + 1e-5f) * 0.2f + 0.9f;
  }
  __syncthreads();
  // ---- optional raw log-mel [B, n_frames, 128]
  if (log_mel != nullptr) {
    for (int i = tid; i < FE_FRAMES * FE_NMEL; i += 256) {
      const int dt = i >> 7, m = i & 127;
      if (frame0 + dt < n_frames) log_mel[((size_t)b * n_frames + frame0 + dt) * FE_NMEL + m] = s_out[dt][m];
    }
  }
  // ---- patches: token p = 8t+f, element dt*16+df = mel[16t+dt, 16f+df] (eval_caco_torch.py:124-129)
  if (want_patch) {
    for (int i = tid; i < 2048; i += 256) {
      const int f = i >> 8, e = i & 255;
      const int p = t * 8 + f;
      if (p < max_patches) {
        const float val = s_out[e >> 4][16 * f + (e & 15)];
        const size_t o = ((size_t)b * max_patches + p) * 256 + e;
        if (patches) patches[o] = val;
        if (patches_f16) patches_f16[o] = __float2half_rn(val);
      }
    }
  } else if (t < T_out) {
    for (int i = tid; i < 2048; i += 256) {
      const int p = t * 8 + (i >> 8);
      if (p < max_patches) {
        const size_t o = ((size_t)b * max_patches + p) * 256 + (i & 255);
        if (patches) patches[o] = 0.0f;
        if (patches_f16) patches_f16[o] = __float2half_rn(0.0f);
      }
    }
  }
}

// 2-D tensor map over wave[batch][n_samples] fp32, box = 256 samples of one clip, no swizzle, zero fill out of bounds
static int make_tmap_wave(CUtensorMap* map, const float* wave, int batch, int n_samples) {
  typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static PFN enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return CACO_ERR_DRIVER;
    enc = reinterpret_cast<PFN>(p);
  }
  cuuint64_t dims[2] = {(cuuint64_t)n_samples, (cuuint64_t)batch};
  cuuint64_t strides[1] = {(cuuint64_t)n_samples * 4};
  cuuint32_t box[2] = {FE_TMA_BOX, 1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(wave), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : CACO_ERR_DRIVER;
}

int frontend(const float* wave, const int* lengths, int batch, int n_samples, int max_patches, float* patches,
             void* patches_f16, float* time_inds, float* freq_inds, float* mask, float* log_mel, cudaStream_t stream) {
  if (!wave || (!patches && !patches_f16) || !time_inds || !freq_inds || !mask || batch <= 0 || n_samples <= 0 || max_patches <= 0)
    return CACO_ERR_ARG;
  if (lengths != nullptr && log_mel != nullptr) return CACO_ERR_ARG;   // the raw log-mel output is per-length: uniform batches only
  int rc = frontend_init();
  if (rc) return rc;
  const int n_frames = (n_samples + FE_HOP - 1) / FE_HOP;
  const int T_out = (max_patches + 7) / 8;
  int gx = T_out;
  if (log_mel != nullptr) gx = max(gx, (n_frames + FE_FRAMES - 1) / FE_FRAMES);
  dim3 grid(gx, batch);
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaError_t e = cudaFuncSetAttribute(frontend_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FE_SMEM_BYTES);
    if (!e) e = cudaFuncSetAttribute(frontend_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FE_SMEM_BYTES);
    if (!e) e = cudaFuncSetAttribute(frontend_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FE_SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_once.done();
  }
  // TMA staging needs a 16-byte aligned base and row pitch (every real batch: torch allocations, 160000-sample rows); the
  // tensor map's inner extent is the common row length, so OOB zero fill reproduces the reference's tail padding
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  bool use_tma = ((reinterpret_cast<uintptr_t>(wave) & 15) == 0) && ((n_samples & 3) == 0) && batch <= 65535;
  if (use_tma) use_tma = make_tmap_wave(&map, wave, batch, n_samples) == 0;
  __half* p16 = reinterpret_cast<__half*>(patches_f16);
  if (use_tma && lengths == nullptr)
    frontend_kernel<true, false><<<grid, 256, FE_SMEM_BYTES, stream>>>(map, wave, lengths, n_samples, n_samples, max_patches, patches,
                                                                       p16, time_inds, freq_inds, mask, log_mel);
  else if (use_tma)
    frontend_kernel<true, true><<<grid, 256, FE_SMEM_BYTES, stream>>>(map, wave, lengths, n_samples, n_samples, max_patches, patches,
                                                                      p16, time_inds, freq_inds, mask, log_mel);
  else
    frontend_kernel<false, false><<<grid, 256, FE_SMEM_BYTES, stream>>>(map, wave, lengths, n_samples, n_samples, max_patches, patches,
                                                                        p16, time_inds, freq_inds, mask, log_mel);
  count_launch();
  return (int)cudaGetLastError();
}

}  // namespace caco

extern "C" int caco_frontend(const float* wave, int batch, int n_samples, int max_patches, float* patches,
                             void* patches_f16, float* time_inds, float* freq_inds, float* mask, float* log_mel,
                             void* stream) {
  return caco::frontend(wave, nullptr, batch, n_samples, max_patches, patches, patches_f16, time_inds, freq_inds, mask, log_mel,
                        (cudaStream_t)stream);
}

// ragged batch: clip b = wave[b*stride : b*stride + lengths[b]] (lengths: DEVICE int32 [batch]); every clip gets the
// reference's per-clip treatment (its own frame count, valid-patch count, zero padding, mask) — eval_caco_torch.py:181-206
// applied clip by clip, in one launch
extern "C" int caco_frontend_ragged(const float* wave, const int* lengths, int batch, int stride, int max_patches,
                                    float* patches, void* patches_f16, float* time_inds, float* freq_inds, float* mask,
                                    void* stream) {
  if (!lengths) return CACO_ERR_ARG;
  return caco::frontend(wave, lengths, batch, stride, max_patches, patches, patches_f16, time_inds, freq_inds, mask, nullptr,
                        (cudaStream_t)stream);
}

// host-only helper (no GPU needed): the fp32 HTK filterbank the kernel uses, [257][128]
extern "C" int caco_mel_filterbank(float* out_257x128) {
  if (!out_257x128) return CACO_ERR_ARG;
  caco::mel_filterbank_host(out_257x128);
  return 0;
}
