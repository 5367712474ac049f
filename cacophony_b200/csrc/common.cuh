// Shared host-side helpers for the library's translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace caco {

extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms();

// Function attributes (max dynamic shared memory) and __device__ tables are per DEVICE: one-time work is tracked per
// device, so a process that drives several GPUs (model.to("cuda:1") next to "cuda:0") initialises each of them.
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev & 63;
}
struct PerDeviceOnce {
  std::atomic<uint64_t> mask{0};
  bool first() {                       // true exactly once per device (benign if two threads race: the work is idempotent)
    const uint64_t bit = 1ull << current_device();
    return !(mask.load(std::memory_order_acquire) & bit);
  }
  void done() { mask.fetch_or(1ull << current_device(), std::memory_order_release); }
};
struct PerDeviceMax {                  // largest value already configured on each device (for growing smem opt-ins)
  size_t cur[64] = {};
  bool need(size_t v) const { return v > cur[current_device()]; }
  void set(size_t v) { cur[current_device()] = v; }
};

// Execution options.  A model handle owns one set (caco_model_set_option) and installs it for the duration of each of its
// calls on the calling thread; op-level C-ABI calls outside a handle use the library defaults (caco_set_default_option).
struct Options {
  int pdl = 1;              // programmatic dependent launch of the tower kernels (ptx.cuh: pdl_wait / pdl_launch)
  int gemm_variant = 0;     // 0 = auto, else CACO_GEMM_*
  int resid_red = 1;        // in-place residual GEMMs add through the L2 (red.global.add.v4.f32)
  int audio_chunk_rows = 131072;   // token rows per pass of the audio tower (larger batches are chunked)
  int text_chunk_rows = 131072;
  int split_weights = 0;    // precision mode: GEMM weights as fp16 hi + lo (two accumulating MMA passes)
};
extern Options g_default_opts;
extern thread_local const Options* tl_opts;
inline const Options& opts() { return tl_opts ? *tl_opts : g_default_opts; }
struct OptionsScope {       // RAII: a handle's options are current on this thread while one of its calls enqueues work
  const Options* prev;
  explicit OptionsScope(const Options* o) : prev(tl_opts) { tl_opts = o; }
  ~OptionsScope() { tl_opts = prev; }
};
int set_option(Options& o, const char* name, int value);   // 0 ok, CACO_ERR_ARG unknown name / bad value
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = opts().pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// op-level entry points implemented across the .cu files (C++ side of the C ABI)
int gemm_f16(const void* A, int lda, const void* W, int ldw, const float* bias, const float* resid, int ldr, void* out,
             int ldo, int M, int N, int K, int epi, int variant, int max_ctas, cudaStream_t stream, int a_k = 0);
int frontend(const float* wave, const int* lengths, int batch, int n_samples, int max_patches, float* patches,
             void* patches_f16, float* time_inds, float* freq_inds, float* mask, float* log_mel, cudaStream_t stream);
int cast_f32_f16(const float* src, void* dst, int64_t n, cudaStream_t stream);
int cast_f32_f16_split(const float* src, void* dst, int64_t rows, int64_t K, cudaStream_t stream);
int kv_append(const void* qkv, void* cache, const int64_t* pos, float* key_mask, int batch, int D, int capacity,
              cudaStream_t stream);
int topk_rows(const float* x, int rows, int cols, int ldx, int k, int* idx_out, float* val_out, cudaStream_t stream);
int layernorm(const float* x, const float* gamma, const float* beta, float eps, float* out_f32, void* out_f16, int rows,
              int dim, cudaStream_t stream);
int audio_add_pos(float* x, const float* time_inds, const float* freq_inds, const float* freq_emb, int n_freq, int rows,
                  int dim, int init, cudaStream_t stream);
int attention_audio(const void* qkv, const float* mask, void* out, int batch, int seq, int heads, int dh,
                    cudaStream_t stream);
int attention_text(const void* qkv, const float* key_mask, void* out, int batch, int T, int heads, int dh,
                   cudaStream_t stream);
int attention_cross(const void* q, int ldq, const void* kv, const float* key_mask, void* out, int batch, int Tq, int Skv,
                    int heads, int dh, cudaStream_t stream);
int text_embed_ln(const int64_t* ids, const int64_t* position_ids, const float* word, const float* pos,
                  const float* type0, const float* gamma, const float* beta, float eps, float* out_f32, void* out_f16,
                  int batch, int T, int dim, int vocab, int max_pos, cudaStream_t stream);
int attn_pool(const float* hid, const float* mask, const float* u, const float* c, const float* ln_gamma,
              const float* ln_beta, float ln_eps, float* hid_out, float* pooled, int batch, int seq, int heads, int dim,
              cudaStream_t stream);
int sgemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias, float alpha, float* out, int ldo,
             int M, int N, int K, cudaStream_t stream);
int l2norm(const float* in, float* out, int rows, int dim, float eps, cudaStream_t stream);
int sim_logits(const float* a, const float* t, const float* logit_scale, float* at, float* ta, int na, int nt, int dim,
               cudaStream_t stream);

}  // namespace caco
