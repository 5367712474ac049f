// K1 — waveform -> |STFT| -> HTK mel -> log -> 16x16 patches (+ time/freq indices and mask).
// Replaces compute_mel_spectrogram + spectrogram_to_patches + prepare_audio_batch
// (src/eval/eval_caco_torch.py:41-105, :108-151, :181-206), batched and without the host numpy hop.
//
// One CTA per (clip, 16-frame group) = one patch row t: it stages the 2912 samples the 16 frames touch,
// runs 16 real 512-point FFTs in shared memory (256-point complex radix-4 Stockham + real split, fp32,
// host-computed twiddle table), applies the sparse HTK filterbank (505 non-zeros), log(x+1e-5)*0.2+0.9, and
// writes the 8 patches of that row — 8 KB contiguous in the [B, max_patches, 256] layout — with
// coalesced 128-byte stores.  HBM-bound by design: 1.158 MB per 10 s clip.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
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

struct MelTable {
  int start[FE_NMEL];
  int count[FE_NMEL];
  float w[FE_NMEL][FE_MAXW];
};
__device__ MelTable g_mel;
__device__ float2 g_twiddle[FE_NFFT];   // e^{-2 pi i k / 512}, computed in double on the host
__device__ float g_window[FE_WIN];      // periodic Hann(400)

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
  static int rc = -100;
  static std::once_flag once;
  std::call_once(once, [] {
    std::vector<float> fb(FE_NFREQ * FE_NMEL);
    mel_filterbank_host(fb.data());
    MelTable t;
    rc = 0;
    for (int m = 0; m < FE_NMEL; ++m) {
      int lo = -1, hi = -1;
      for (int k = 0; k < FE_NFREQ; ++k)
        if (fb[k * FE_NMEL + m] != 0.0f) { if (lo < 0) lo = k; hi = k; }
      t.start[m] = lo < 0 ? 0 : lo;
      t.count[m] = lo < 0 ? 0 : hi - lo + 1;
      if (t.count[m] > FE_MAXW) { rc = CACO_ERR_STATE; return; }
      for (int j = 0; j < FE_MAXW; ++j) t.w[m][j] = (j < t.count[m]) ? fb[(t.start[m] + j) * FE_NMEL + m] : 0.0f;
    }
    cudaError_t e = cudaMemcpyToSymbol(g_mel, &t, sizeof(t));
    if (e != cudaSuccess) { rc = (int)e; return; }
    std::vector<float2> tw(FE_NFFT);
    for (int k = 0; k < FE_NFFT; ++k) {
      const double ang = -2.0 * M_PI * (double)k / (double)FE_NFFT;
      tw[k] = make_float2((float)cos(ang), (float)sin(ang));
    }
    e = cudaMemcpyToSymbol(g_twiddle, tw.data(), sizeof(float2) * FE_NFFT);
    if (e != cudaSuccess) { rc = (int)e; return; }
    std::vector<float> win(FE_WIN);
    for (int i = 0; i < FE_WIN; ++i) win[i] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)i / (double)FE_WIN));   // torch.hann_window(400)
    e = cudaMemcpyToSymbol(g_window, win.data(), sizeof(float) * FE_WIN);
    if (e != cudaSuccess) rc = (int)e;
  });
  return rc;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__global__ void __launch_bounds__(256)
frontend_kernel(const float* __restrict__ wave, const int* __restrict__ lengths, int stride, int n_samples_u, int max_patches,
                float* __restrict__ patches, __half* __restrict__ patches_f16, float* __restrict__ time_inds,
                float* __restrict__ freq_inds, float* __restrict__ mask, float* __restrict__ log_mel) {
  __shared__ float s_x[FE_SAMPLES];
  __shared__ float2 s_tw[FE_NFFT];                 // e^{-2 pi i k / 512}
  __shared__ float s_win[FE_WIN];
  __shared__ float2 s_buf[2][4][256];              // ping-pong, 4 frames in flight
  __shared__ float s_mag[4][FE_NFREQ + 3];
  __shared__ float s_out[FE_FRAMES][FE_MEL_LD];

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

  // ---- stage samples (zero tail pad, eval_caco_torch.py:72-78), twiddles, window, mel table
  const float* wv = wave + (size_t)b * stride;
  const int s0 = frame0 * FE_HOP;
  for (int i = tid; i < FE_SAMPLES; i += 256) {
    const int s = s0 + i;
    s_x[i] = (s < n_samples) ? __ldg(wv + s) : 0.0f;
  }
  for (int i = tid; i < FE_NFFT; i += 256) s_tw[i] = g_twiddle[i];
  for (int i = tid; i < FE_WIN; i += 256) s_win[i] = g_window[i];
  __syncthreads();

  const int slot = tid >> 6;  // frame slot 0..3
  const int j = tid & 63;     // radix-4 butterfly index
  for (int round = 0; round < FE_FRAMES / 4; ++round) {
    const int dt = round * 4 + slot;
    const float* xf = s_x + dt * FE_HOP;
    int cur = 0;
#pragma unroll
    for (int stage = 0; stage < 4; ++stage) {
      const int Ns = 1 << (2 * stage);
      float2 v[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = j + r * 64;
        if (stage == 0) {
          // z[n] = xw[2n] + i xw[2n+1], window support is [56, 456)
          const int i0 = 2 * n, i1 = 2 * n + 1;
          const float w0 = (i0 >= FE_WOFF && i0 < FE_WOFF + FE_WIN) ? s_win[i0 - FE_WOFF] : 0.0f;
          const float w1 = (i1 >= FE_WOFF && i1 < FE_WOFF + FE_WIN) ? s_win[i1 - FE_WOFF] : 0.0f;
          v[r] = make_float2(xf[i0] * w0, xf[i1] * w1);
        } else {
          v[r] = s_buf[cur][slot][n];
          if (r > 0) v[r] = cmul(v[r], s_tw[((j & (Ns - 1)) * r * (128 / Ns)) & 511]);
        }
      }
      const float2 a0 = make_float2(v[0].x + v[2].x, v[0].y + v[2].y);
      const float2 a1 = make_float2(v[0].x - v[2].x, v[0].y - v[2].y);
      const float2 a2 = make_float2(v[1].x + v[3].x, v[1].y + v[3].y);
      const float2 d = make_float2(v[1].x - v[3].x, v[1].y - v[3].y);
      const float2 a3 = make_float2(d.y, -d.x);  // -i * d
      const int idx = (j / Ns) * Ns * 4 + (j & (Ns - 1));
      float2* dst = s_buf[stage == 0 ? 0 : (cur ^ 1)][slot];
      dst[idx] = make_float2(a0.x + a2.x, a0.y + a2.y);
      dst[idx + Ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
      dst[idx + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
      dst[idx + 3 * Ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
      if (stage > 0) cur ^= 1;
      named_bar_sync(1 + slot, 64);   // only the two warps working on this frame
    }
    // ---- real split: X[k] = E + W^k O, |X[k]|, k = 0..256
    const float2* Z = s_buf[cur][slot];
    for (int k = j; k <= 256; k += 64) {
      const float2 zk = Z[k & 255];
      const float2 zc = Z[(256 - k) & 255];
      const float er = 0.5f * (zk.x + zc.x), ei = 0.5f * (zk.y - zc.y);      // E = (Zk + conj Zc)/2
      const float dr = zk.x - zc.x, di = zk.y + zc.y;                        // Zk - conj Zc
      const float orr = 0.5f * di, oi = -0.5f * dr;                          // O = -i/2 (Zk - conj Zc)
      const float2 w = (k < 256) ? s_tw[k] : make_float2(-1.0f, 0.0f);
      const float xr = er + (w.x * orr - w.y * oi);
      const float xi = ei + (w.x * oi + w.y * orr);
      s_mag[slot][k] = sqrtf(xr * xr + xi * xi);
    }
    named_bar_sync(1 + slot, 64);   // only the two warps working on this frame
    // ---- sparse mel + log (eval_caco_torch.py:103-104)
    for (int m = j; m < FE_NMEL; m += 64) {
      const int st = g_mel.start[m], cnt = g_mel.count[m];  // 9 KB table, L1/L2 resident
      float acc = 0.0f;
      for (int q = 0; q < cnt; ++q) acc = fmaf(s_mag[slot][st + q], g_mel.w[m][q], acc);
      s_out[dt][m] = logf(acc + 1e-5f) * 0.2f + 0.9f;
    }
    named_bar_sync(1 + slot, 64);   // only the two warps working on this frame
  }

  __syncthreads();   // all four 64-thread groups have filled their rows of s_out
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
        const float v = s_out[e >> 4][16 * f + (e & 15)];
        const size_t o = ((size_t)b * max_patches + p) * 256 + e;
        if (patches) patches[o] = v;
        if (patches_f16) patches_f16[o] = __float2half_rn(v);
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
  frontend_kernel<<<grid, 256, 0, stream>>>(wave, lengths, n_samples, n_samples, max_patches, patches,
                                            reinterpret_cast<__half*>(patches_f16), time_inds, freq_inds, mask, log_mel);
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
