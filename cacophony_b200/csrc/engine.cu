// Model handle: the C++ side of CACO (src/caco_torch/caco.py:82-261).  Owns fp16-packed GEMM weights, the folded
// pooler vectors and an activation workspace; borrows everything else (biases, LayerNorm parameters, embedding
// tables, inputs, outputs) from the caller.  The two towers are straight-line sequences of the kernels in
// gemm.cu / attention.cu / rowops.cu on ONE stream — at batch 256 every launch is milliseconds long, so there is
// nothing for a graph to save; the host cost is ~100 launches per tower.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <unordered_map>
#include <vector>

#include "caco_b200.h"
#include "common.cuh"

namespace caco {
int fold_query(const float* query, const float* wk, const float* bk, float qscale, float* u, float* c, int heads, int dh,
               int dim, cudaStream_t stream);
static thread_local char g_err[256] = "";
static void set_err(const char* fmt, const char* a) { snprintf(g_err, sizeof(g_err), fmt, a); }
}  // namespace caco

struct caco_model {
  caco_config cfg;
  struct T { const float* p; int64_t n; };
  std::unordered_map<std::string, T> w;
  bool packed = false;

  struct ALayer {
    const __half *qkv_w, *out_w, *fc1_w, *fc2_w;
    const float *qkv_b, *out_b, *fc1_b, *fc2_b, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  };
  struct TLayer {
    const __half *qkv_w, *o_w, *fc1_w, *fc2_w;
    const float *qkv_b, *o_b, *fc1_b, *fc2_b, *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  };
  struct DLayer {          // captioning decoder layer (roberta.py:181-215 with crossattention): self-attention block as in TLayer,
    TLayer self_ffn;       // + cross-attention: q from the text side, k | v (stacked [2D, D]) from the audio hidden state
    const __half *cq_w, *ckv_w, *co_w;
    const float *cq_b, *ckv_b, *co_b, *lnc_g, *lnc_b;
  };
  std::vector<ALayer> al;
  std::vector<TLayer> tl;
  std::vector<DLayer> dl;
  const __half* dproj_w = nullptr;       // decoder_proj [vocab_pad, D] (rows past the vocabulary are zero)
  float* dproj_b = nullptr;              // [vocab_pad] in arena32
  int dec_vocab = 0, dec_vocab_pad = 0;
  const __half* in_w = nullptr;
  const float *in_b = nullptr, *freq_emb = nullptr, *an_g = nullptr, *an_b = nullptr;
  const float *ap_vw = nullptr, *ap_vb = nullptr, *ap_ow = nullptr, *ap_ob = nullptr;   // audio pooler value rows / out_proj
  const float *word = nullptr, *pos = nullptr, *type0 = nullptr, *te_g = nullptr, *te_b = nullptr;
  const float *tp_vw = nullptr, *tp_vb = nullptr, *proj_w = nullptr, *proj_b = nullptr;
  const float* logit_scale = nullptr;

  __half* arena16 = nullptr;   // all fp16 GEMM weights
  float* arena32 = nullptr;    // packed text qkv biases + folded pooler vectors
  float *a_u = nullptr, *a_c = nullptr, *t_u = nullptr, *t_c = nullptr;

  void* ws[3] = {nullptr, nullptr, nullptr};   // [0] audio tower, [1] text tower (they may run on different streams), [2] decoder
  size_t ws_bytes[3] = {0, 0, 0};
  int device = -1;                       // the device the arenas / workspaces live on (a handle follows its model's .to())
  int packed_split = 0;                  // layout of arena16: 0 = [N, K] fp16, 1 = [N, 2K] fp16 hi | lo (split_weights)
  uint64_t generation = 0;               // bumped whenever a pointer a captured CUDA graph may have baked in is released
  caco::Options opt;                     // per-handle execution options (caco_model_set_option)
};

namespace caco {

static const float* need(caco_model* m, const std::string& key, int64_t numel, int* rc) {
  auto it = m->w.find(key);
  if (it == m->w.end()) { set_err("missing tensor %s", key.c_str()); *rc = CACO_ERR_STATE; return nullptr; }
  if (it->second.n != numel) { set_err("wrong size for %s", key.c_str()); *rc = CACO_ERR_STATE; return nullptr; }
  return it->second.p;
}

static void free_ws(caco_model* m, int which) {
  if (!m->ws[which]) return;
  cudaDeviceSynchronize();
  cudaFree(m->ws[which]);
  m->ws[which] = nullptr;
  m->ws_bytes[which] = 0;
  ++m->generation;
}
static int ensure_ws(caco_model* m, int which, size_t bytes) {
  if (bytes <= m->ws_bytes[which]) return 0;
  // A caller that grows its request a little every call (a decode loop re-running get_decoder_logits on a prefix that gets one
  // token longer each step) would otherwise synchronise, free and re-allocate every step: grow the text / decoder workspaces
  // geometrically (at most 256 MB beyond the request).  The audio workspace (gigabytes at bench size) is sized exactly.
  size_t want = bytes;
  if (which != 0 && m->ws_bytes[which] > 0) {
    const size_t geo = m->ws_bytes[which] * 2, cap = bytes + ((size_t)256 << 20);
    if (geo > want) want = geo < cap ? geo : cap;
  }
  free_ws(m, which);
  cudaError_t e = cudaMalloc(&m->ws[which], want);
  if (e != cudaSuccess && want > bytes) {
    (void)cudaGetLastError();          // clear the failed attempt: the launch wrappers report cudaGetLastError()
    want = bytes;
    e = cudaMalloc(&m->ws[which], want);
  }
  if (e != cudaSuccess) return (int)e;
  m->ws_bytes[which] = want;
  return 0;
}
// every entry point: the handle must be used on the device it was packed on (its arenas and workspaces live there)
static int check_device(const caco_model* m) {
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != m->device) { set_err("model packed on another device than the current one%s", ""); return CACO_ERR_STATE; }
  return 0;
}

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

#define CK(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

static int pack(caco_model* m, cudaStream_t st) {
  const caco_config& c = m->cfg;
  const int64_t D = c.hidden, F = c.ffn, P = c.patch_dim;
  int rc = 0;
  // ---- sizes
  const int split = m->opt.split_weights ? 1 : 0;
  const int64_t per_layer = 3 * D * D + D * D + 2 * F * D;
  // captioning head (optional): present iff decoder_module.decoder_proj.weight was registered
  int dec_layers = 0;
  int64_t dec_vocab = 0, dec_vocab_pad = 0;
  {
    auto it = m->w.find("decoder_module.decoder_proj.weight");
    if (it != m->w.end()) {
      dec_vocab = it->second.n / D;
      dec_vocab_pad = (dec_vocab + 7) & ~(int64_t)7;
      while (m->w.count("decoder_module.encoder.layers." + std::to_string(dec_layers) + ".intermediate.dense.weight")) ++dec_layers;
    }
  }
  const int64_t dec_per_layer = per_layer + D * D + 2 * D * D + D * D;      // + cross q, k | v, out
  const int64_t n16 = (P * D + per_layer * (c.audio_layers + c.text_layers) + dec_per_layer * dec_layers + dec_vocab_pad * D) *
                      (split ? 2 : 1);
  const int64_t n32 = (int64_t)c.text_layers * 3 * D + (int64_t)(c.pool_heads + 1) * D + 4 * 64 +
                      (int64_t)dec_layers * 6 * D + dec_vocab_pad;
  int cur = 0;
  cudaGetDevice(&cur);
  if (m->device != cur) {                // first pack, or the model moved to another GPU: nothing of the old device is kept
    free_ws(m, 0);
    free_ws(m, 1);
    free_ws(m, 2);
    m->device = cur;
  }
  m->packed = false;
  if (m->arena16 || m->arena32) { cudaDeviceSynchronize(); ++m->generation; }
  if (m->arena16) cudaFree(m->arena16);
  if (m->arena32) cudaFree(m->arena32);
  m->arena16 = nullptr;
  m->arena32 = nullptr;
  cudaError_t e = cudaMalloc(&m->arena16, n16 * sizeof(__half));
  if (e) return (int)e;
  e = cudaMalloc(&m->arena32, n32 * sizeof(float));
  if (e) return (int)e;
  __half* p16 = m->arena16;
  float* p32 = m->arena32;
  // nn.Linear.weight [rows, K] fp32 -> fp16 [rows, K], or in split mode [rows, 2K] = fp16(w) | fp16(w - fp16(w))
  auto take16 = [&](const float* src, int64_t n, int64_t K) -> const __half* {
    __half* dst = p16;
    p16 += n * (split ? 2 : 1);
    if (src && rc == 0) rc = split ? cast_f32_f16_split(src, dst, n / K, K, st) : cast_f32_f16(src, dst, n, st);
    return dst;
  };
  // ---- audio tower (key names: SURVEY.md §8b)
  const std::string A = "audio_module.";
  m->in_w = take16(need(m, A + "input_proj.weight", D * P, &rc), D * P, P);
  m->in_b = need(m, A + "input_proj.bias", D, &rc);
  m->freq_emb = need(m, A + "freq_positional_embedding", (int64_t)c.n_freq * D, &rc);
  m->an_g = need(m, A + "norm.weight", D, &rc);
  m->an_b = need(m, A + "norm.bias", D, &rc);
  m->al.resize(c.audio_layers);
  for (int i = 0; i < c.audio_layers && rc == 0; ++i) {
    const std::string L = A + "layers." + std::to_string(i) + ".";
    caco_model::ALayer& l = m->al[i];
    l.ln1_g = need(m, L + "norm1.weight", D, &rc);
    l.ln1_b = need(m, L + "norm1.bias", D, &rc);
    l.ln2_g = need(m, L + "norm2.weight", D, &rc);
    l.ln2_b = need(m, L + "norm2.bias", D, &rc);
    l.qkv_w = take16(need(m, L + "attn.in_proj_weight", 3 * D * D, &rc), 3 * D * D, D);
    l.qkv_b = need(m, L + "attn.in_proj_bias", 3 * D, &rc);
    l.out_w = take16(need(m, L + "attn.out_proj.weight", D * D, &rc), D * D, D);
    l.out_b = need(m, L + "attn.out_proj.bias", D, &rc);
    l.fc1_w = take16(need(m, L + "mlp.fc1.weight", F * D, &rc), F * D, D);
    l.fc1_b = need(m, L + "mlp.fc1.bias", F, &rc);
    l.fc2_w = take16(need(m, L + "mlp.fc2.weight", D * F, &rc), D * F, F);
    l.fc2_b = need(m, L + "mlp.fc2.bias", D, &rc);
  }
  if (rc) return rc;
  // ---- audio pooler (caco.py:24-79): fold kv_proj's key half into the query
  {
    const std::string Q = "audio_attention_pool.";
    const float* query = need(m, Q + "query", D, &rc);
    const float* kvw = need(m, Q + "kv_proj.weight", 2 * D * D, &rc);
    const float* kvb = need(m, Q + "kv_proj.bias", 2 * D, &rc);
    m->ap_ow = need(m, Q + "out_proj.weight", D * D, &rc);
    m->ap_ob = need(m, Q + "out_proj.bias", D, &rc);
    if (rc) return rc;
    m->ap_vw = kvw + D * D;
    m->ap_vb = kvb + D;
    m->a_u = p32; p32 += (int64_t)c.pool_heads * D;
    m->a_c = p32; p32 += 64;   // keep later carves 16-byte aligned (they are read with float4 loads)
    const int dh = (int)D / c.pool_heads;
    CK(fold_query(query, kvw, kvb, 1.0f / sqrtf((float)dh), m->a_u, m->a_c, c.pool_heads, dh, (int)D, st));
  }
  // ---- text tower
  const std::string Tm = "text_module.";
  m->word = need(m, Tm + "embeddings.word_embeddings.weight", (int64_t)c.vocab * D, &rc);
  m->pos = need(m, Tm + "embeddings.position_embeddings.weight", (int64_t)c.max_pos * D, &rc);
  m->type0 = need(m, Tm + "embeddings.token_type_embeddings.weight", D, &rc);
  m->te_g = need(m, Tm + "embeddings.LayerNorm.weight", D, &rc);
  m->te_b = need(m, Tm + "embeddings.LayerNorm.bias", D, &rc);
  m->tl.resize(c.text_layers);
  for (int i = 0; i < c.text_layers && rc == 0; ++i) {
    const std::string L = Tm + "encoder.layers." + std::to_string(i) + ".";
    caco_model::TLayer& l = m->tl[i];
    // q | k | v stacked into one [3D, D] weight so the three projections are one GEMM (roberta.py:77-83)
    __half* qkv = p16;
    float* qkvb = p32;
    const char* nm[3] = {"query", "key", "value"};
    for (int j = 0; j < 3; ++j) {
      take16(need(m, L + "attention.self." + nm[j] + ".weight", D * D, &rc), D * D, D);
      const float* b = need(m, L + "attention.self." + nm[j] + ".bias", D, &rc);
      if (rc) return rc;
      e = cudaMemcpyAsync(p32, b, D * sizeof(float), cudaMemcpyDeviceToDevice, st);
      if (e) return (int)e;
      p32 += D;
    }
    l.qkv_w = qkv;
    l.qkv_b = qkvb;
    l.o_w = take16(need(m, L + "attention.output.dense.weight", D * D, &rc), D * D, D);
    l.o_b = need(m, L + "attention.output.dense.bias", D, &rc);
    l.ln1_g = need(m, L + "attention.output.LayerNorm.weight", D, &rc);
    l.ln1_b = need(m, L + "attention.output.LayerNorm.bias", D, &rc);
    l.fc1_w = take16(need(m, L + "intermediate.dense.weight", F * D, &rc), F * D, D);
    l.fc1_b = need(m, L + "intermediate.dense.bias", F, &rc);
    l.fc2_w = take16(need(m, L + "output.dense.weight", D * F, &rc), D * F, F);
    l.fc2_b = need(m, L + "output.dense.bias", D, &rc);
    l.ln2_g = need(m, L + "output.LayerNorm.weight", D, &rc);
    l.ln2_b = need(m, L + "output.LayerNorm.bias", D, &rc);
  }
  if (rc) return rc;
  {
    const std::string Q = Tm + "pooler.";
    const float* query = need(m, Q + "attention_pool_query", D, &rc);
    const float* kw = need(m, Q + "key_proj.weight", D * D, &rc);
    const float* kb = need(m, Q + "key_proj.bias", D, &rc);
    m->tp_vw = need(m, Q + "value_proj.weight", D * D, &rc);
    m->tp_vb = need(m, Q + "value_proj.bias", D, &rc);
    m->proj_w = need(m, "text_proj.weight", D * D, &rc);
    m->proj_b = need(m, "text_proj.bias", D, &rc);
    m->logit_scale = need(m, "logit_scale", 1, &rc);
    if (rc) return rc;
    m->t_u = p32; p32 += D;
    m->t_c = p32; p32 += 64;
    CK(fold_query(query, kw, kb, 1.0f / sqrtf((float)D), m->t_u, m->t_c, 1, (int)D, (int)D, st));   // roberta.py:259
  }
  // ---- captioning decoder (roberta.py:329-373), only when its tensors were registered
  m->dl.clear();
  m->dl.resize(dec_layers);
  m->dec_vocab = (int)dec_vocab;
  m->dec_vocab_pad = (int)dec_vocab_pad;
  for (int i = 0; i < dec_layers && rc == 0; ++i) {
    const std::string L = "decoder_module.encoder.layers." + std::to_string(i) + ".";
    caco_model::DLayer& d = m->dl[i];
    caco_model::TLayer& l = d.self_ffn;
    auto stack = [&](const std::string& blk, const char* const* names, int n, const __half** w_out, const float** b_out) {
      __half* w0 = p16;
      float* b0 = p32;
      for (int j = 0; j < n; ++j) {
        take16(need(m, L + blk + ".self." + names[j] + ".weight", D * D, &rc), D * D, D);
        const float* b = need(m, L + blk + ".self." + names[j] + ".bias", D, &rc);
        if (rc) return;
        if (cudaMemcpyAsync(p32, b, D * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) { rc = CACO_ERR_STATE; return; }
        p32 += D;
      }
      *w_out = w0;
      *b_out = b0;
    };
    const char* qkv_n[3] = {"query", "key", "value"};
    const char* q_n[1] = {"query"};
    const char* kv_n[2] = {"key", "value"};
    stack("attention", qkv_n, 3, &l.qkv_w, &l.qkv_b);
    if (rc) return rc;
    l.o_w = take16(need(m, L + "attention.output.dense.weight", D * D, &rc), D * D, D);
    l.o_b = need(m, L + "attention.output.dense.bias", D, &rc);
    l.ln1_g = need(m, L + "attention.output.LayerNorm.weight", D, &rc);
    l.ln1_b = need(m, L + "attention.output.LayerNorm.bias", D, &rc);
    stack("crossattention", q_n, 1, &d.cq_w, &d.cq_b);
    stack("crossattention", kv_n, 2, &d.ckv_w, &d.ckv_b);
    if (rc) return rc;
    d.co_w = take16(need(m, L + "crossattention.output.dense.weight", D * D, &rc), D * D, D);
    d.co_b = need(m, L + "crossattention.output.dense.bias", D, &rc);
    d.lnc_g = need(m, L + "crossattention.output.LayerNorm.weight", D, &rc);
    d.lnc_b = need(m, L + "crossattention.output.LayerNorm.bias", D, &rc);
    l.fc1_w = take16(need(m, L + "intermediate.dense.weight", F * D, &rc), F * D, D);
    l.fc1_b = need(m, L + "intermediate.dense.bias", F, &rc);
    l.fc2_w = take16(need(m, L + "output.dense.weight", D * F, &rc), D * F, F);
    l.fc2_b = need(m, L + "output.dense.bias", D, &rc);
    l.ln2_g = need(m, L + "output.LayerNorm.weight", D, &rc);
    l.ln2_b = need(m, L + "output.LayerNorm.bias", D, &rc);
  }
  if (rc) return rc;
  m->dproj_w = nullptr;
  if (dec_layers > 0) {
    // vocabulary projection, rows padded to a multiple of 8 (50265 -> 50272) with zeros so that the GEMM's 16-byte stores apply
    __half* w0 = p16;
    const int64_t el = split ? 2 : 1;
    e = cudaMemsetAsync(w0, 0, dec_vocab_pad * D * el * sizeof(__half), st);
    if (e) return (int)e;
    take16(need(m, "decoder_module.decoder_proj.weight", dec_vocab * D, &rc), dec_vocab * D, D);
    p16 = w0 + dec_vocab_pad * D * el;
    const float* pb = need(m, "decoder_module.decoder_proj.bias", dec_vocab, &rc);
    if (rc) return rc;
    m->dproj_b = p32;
    e = cudaMemsetAsync(p32, 0, dec_vocab_pad * sizeof(float), st);
    if (!e) e = cudaMemcpyAsync(p32, pb, dec_vocab * sizeof(float), cudaMemcpyDeviceToDevice, st);
    if (e) return (int)e;
    p32 += dec_vocab_pad;
    m->dproj_w = w0;
  }
  m->packed = true;
  m->packed_split = split;
  return 0;
}

// y = epilogue(A[M,K] · W^T) with W as packed by pack(): [N, K], or [N, 2K] hi | lo in split-weight mode (A re-read per term)
static int lin(const caco_model* m, const __half* A, int lda, const __half* W, int K, const float* bias, const float* resid,
               int ldr, void* out, int ldo, int M, int N, int epi, cudaStream_t st) {
  const int t = m->packed_split ? 2 : 1;
  return gemm_f16(A, lda, W, K * t, bias, resid, ldr, out, ldo, M, N, K * t, epi, 0, 0, st, K);
}
// entry-point prologue: right device, weights packed in the layout the handle's options ask for
static int ready(caco_model* m, cudaStream_t st) {
  if (!m || !m->packed) return CACO_ERR_STATE;
  CK(check_device(m));
  if ((m->opt.split_weights ? 1 : 0) != m->packed_split) CK(pack(m, st));
  return 0;
}

// ---------------------------------------------------------------------------------------- audio tower
// patches (fp32) or patches16 (fp16 operand copy already made by the frontend): exactly one is non-null
static int audio_chunk(caco_model* m, const float* patches, const __half* patches16, const float* t_inds, const float* f_inds,
                       const float* mask, int B, int S, int normalize, float* emb_out, float* hidden_out, cudaStream_t st) {
  const caco_config& c = m->cfg;
  const int D = c.hidden, F = c.ffn, P = c.patch_dim;
  const size_t R = (size_t)B * S;
  // workspace carve-up
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += al256(bytes); return o; };
  const size_t o_x = carve(R * D * 4), o_h = carve(R * D * 2), o_qkv = carve(R * 3 * D * 2), o_att = carve(R * D * 2);
  const size_t o_mlp = carve(R * (size_t)F * 2), o_p16 = carve(R * P * 2);
  const size_t o_pool = carve((size_t)B * c.pool_heads * D * 4), o_o = carve((size_t)B * D * 4), o_e = carve((size_t)B * D * 4);
  CK(ensure_ws(m, 0, off));
  uint8_t* w = (uint8_t*)m->ws[0];
  float* x = (float*)(w + o_x);
  __half* h16 = (__half*)(w + o_h);
  __half* qkv = (__half*)(w + o_qkv);
  __half* att = (__half*)(w + o_att);
  __half* mlp = (__half*)(w + o_mlp);
  __half* p16 = (__half*)(w + o_p16);
  float* pooled = (float*)(w + o_pool);
  float* ov = (float*)(w + o_o);
  float* eraw = (float*)(w + o_e);
  const int Ri = (int)R;

  // input projection + position embeddings (mae.py:133-142)
  if (patches16 == nullptr) CK(cast_f32_f16(patches, p16, (int64_t)R * P, st));
  // x = pos(t) + freq_emb[f] written first (no read), then the projection accumulates onto it in place (L2 reductions):
  // one pass over x less than project-then-add
  CK(audio_add_pos(x, t_inds, f_inds, m->freq_emb, c.n_freq, Ri, D, 1, st));
  CK(lin(m, patches16 ? patches16 : p16, P, m->in_w, P, m->in_b, x, D, x, D, Ri, D, CACO_EPI_BIAS_RESID_F32, st));
  // pre-LN blocks (mae.py:80-99)
  for (int i = 0; i < c.audio_layers; ++i) {
    const caco_model::ALayer& l = m->al[i];
    CK(layernorm(x, l.ln1_g, l.ln1_b, c.audio_ln_eps, nullptr, h16, Ri, D, st));
    CK(lin(m, h16, D, l.qkv_w, D, l.qkv_b, nullptr, 0, qkv, 3 * D, Ri, 3 * D, CACO_EPI_BIAS_F16, st));
    CK(attention_audio(qkv, mask, att, B, S, c.audio_heads, D / c.audio_heads, st));
    CK(lin(m, att, D, l.out_w, D, l.out_b, x, D, x, D, Ri, D, CACO_EPI_BIAS_RESID_F32, st));
    CK(layernorm(x, l.ln2_g, l.ln2_b, c.audio_ln_eps, nullptr, h16, Ri, D, st));
    CK(lin(m, h16, D, l.fc1_w, D, l.fc1_b, nullptr, 0, mlp, F, Ri, F, CACO_EPI_BIAS_SILU_F16, st));
    CK(lin(m, mlp, F, l.fc2_w, F, l.fc2_b, x, D, x, D, Ri, D, CACO_EPI_BIAS_RESID_F32, st));
  }
  // final LN (mae.py:147) fused into the pooler's pass over the rows (caco.py:41-79)
  CK(attn_pool(x, mask, m->a_u, m->a_c, m->an_g, m->an_b, c.audio_ln_eps, hidden_out, pooled, B, S, c.pool_heads, D, st));
  const int dh = D / c.pool_heads;
  for (int h = 0; h < c.pool_heads; ++h)
    CK(sgemm_nt(pooled + (size_t)h * D, c.pool_heads * D, m->ap_vw + (size_t)h * dh * D, D, m->ap_vb + h * dh, 1.0f,
                ov + h * dh, D, B, dh, D, st));
  float* dst = normalize ? eraw : emb_out;
  CK(sgemm_nt(ov, D, m->ap_ow, D, m->ap_ob, 1.0f, dst, D, B, D, D, st));
  if (normalize) CK(l2norm(eraw, emb_out, B, D, 1e-10f, st));
  return 0;
}

static int audio_embedding(caco_model* m, const float* patches, const __half* patches16, const float* t_inds,
                           const float* f_inds, const float* mask, int B, int S, int normalize, float* emb_out,
                           float* hidden_out, cudaStream_t st) {
  CK(ready(m, st));
  OptionsScope scope(&m->opt);
  if ((!patches && !patches16) || !t_inds || !f_inds || !mask || !emb_out || B <= 0 || S <= 0) return CACO_ERR_ARG;
  const caco_config& c = m->cfg;
  // bound the workspace: at most audio_chunk_rows token rows per pass (default 131072 = 256 clips of 500 tokens)
  int chunk = m->opt.audio_chunk_rows / S;
  if (chunk < 1) chunk = 1;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = (B - b0 < chunk) ? (B - b0) : chunk;
    const size_t r0 = (size_t)b0 * S;
    CK(audio_chunk(m, patches ? patches + r0 * c.patch_dim : nullptr, patches16 ? patches16 + r0 * c.patch_dim : nullptr,
                   t_inds + r0, f_inds + r0, mask + r0, nb, S, normalize,
                   emb_out + (size_t)b0 * c.hidden, hidden_out ? hidden_out + r0 * c.hidden : nullptr, st));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------- text tower
static int text_embedding(caco_model* m, const int64_t* ids, const float* mask, const int64_t* pids, int B, int T,
                          int normalize, float* emb_out, float* hidden_out, cudaStream_t st) {
  CK(ready(m, st));
  OptionsScope scope(&m->opt);
  if (!ids || !mask || !emb_out || B <= 0 || T <= 0) return CACO_ERR_ARG;
  const caco_config& c = m->cfg;
  const int D = c.hidden, F = c.ffn;
  int chunk = m->opt.text_chunk_rows / T;
  if (chunk < 1) chunk = 1;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = (B - b0 < chunk) ? (B - b0) : chunk;
    const size_t R = (size_t)nb * T;
    const int Ri = (int)R;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += al256(bytes); return o; };
    const size_t o_x = carve(R * D * 4), o_x16 = carve(R * D * 2), o_a = carve(R * D * 4), o_a16 = carve(R * D * 2);
    const size_t o_qkv = carve(R * 3 * D * 2), o_att = carve(R * D * 2), o_mlp = carve(R * (size_t)F * 2);
    const size_t o_pool = carve((size_t)nb * D * 4), o_v = carve((size_t)nb * D * 4), o_e = carve((size_t)nb * D * 4);
    CK(ensure_ws(m, 1, off));
    uint8_t* w = (uint8_t*)m->ws[1];
    float* x = (float*)(w + o_x);
    __half* x16 = (__half*)(w + o_x16);
    float* a = (float*)(w + o_a);
    __half* a16 = (__half*)(w + o_a16);
    __half* qkv = (__half*)(w + o_qkv);
    __half* att = (__half*)(w + o_att);
    __half* mlp = (__half*)(w + o_mlp);
    float* pooled = (float*)(w + o_pool);
    float* val = (float*)(w + o_v);
    float* eraw = (float*)(w + o_e);
    const int64_t* ids_c = ids + (size_t)b0 * T;
    const int64_t* pids_c = pids ? pids + (size_t)b0 * T : nullptr;
    const float* mask_c = mask + (size_t)b0 * T;

    CK(text_embed_ln(ids_c, pids_c, m->word, m->pos, m->type0, m->te_g, m->te_b, c.ln_eps, x, x16, nb, T, D, c.vocab,
                     c.max_pos, st));
    for (int i = 0; i < c.text_layers; ++i) {   // post-LN blocks (roberta.py:191-215)
      const caco_model::TLayer& l = m->tl[i];
      CK(lin(m, x16, D, l.qkv_w, D, l.qkv_b, nullptr, 0, qkv, 3 * D, Ri, 3 * D, CACO_EPI_BIAS_F16, st));
      CK(attention_text(qkv, mask_c, att, nb, T, c.text_heads, D / c.text_heads, st));
      // the residual operand is dead after the add (post-LN), so the sum is accumulated in place (L2 reductions)
      CK(lin(m, att, D, l.o_w, D, l.o_b, x, D, x, D, Ri, D, CACO_EPI_BIAS_RESID_F32, st));
      CK(layernorm(x, l.ln1_g, l.ln1_b, c.ln_eps, a, a16, Ri, D, st));
      CK(lin(m, a16, D, l.fc1_w, D, l.fc1_b, nullptr, 0, mlp, F, Ri, F, CACO_EPI_BIAS_GELU_F16, st));
      CK(lin(m, mlp, F, l.fc2_w, F, l.fc2_b, a, D, a, D, Ri, D, CACO_EPI_BIAS_RESID_F32, st));
      CK(layernorm(a, l.ln2_g, l.ln2_b, c.ln_eps, x, x16, Ri, D, st));
    }
    if (hidden_out) {
      cudaError_t e = cudaMemcpyAsync(hidden_out + (size_t)b0 * T * D, x, R * D * 4, cudaMemcpyDeviceToDevice, st);
      if (e) return (int)e;
    }
    CK(attn_pool(x, mask_c, m->t_u, m->t_c, nullptr, nullptr, 0.f, nullptr, pooled, nb, T, 1, D, st));   // roberta.py:253-271
    CK(sgemm_nt(pooled, D, m->tp_vw, D, m->tp_vb, 1.0f, val, D, nb, D, D, st));
    float* dst = normalize ? eraw : emb_out + (size_t)b0 * D;
    CK(sgemm_nt(val, D, m->proj_w, D, m->proj_b, 1.0f, dst, D, nb, D, D, st));                            // caco.py:169
    if (normalize) CK(l2norm(eraw, emb_out + (size_t)b0 * D, nb, D, 1e-10f, st));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------- captioning decoder
// CACO.get_decoder_logits minus the text tower (caco.py:214-240 -> RobertaDecoder.forward, roberta.py:337-373): per layer
// causal self-attention over the text hidden state, cross-attention to the audio hidden state (audio key mask), GELU MLP, all
// post-LN; then the vocabulary projection.  text_hidden [B,T,D] f32 is the text tower's final hidden state, audio_hidden
// [B,S,D] f32 the audio tower's LayerNorm-ed output.  logits_out [B,T,vocab] f32.
static int decoder_logits(caco_model* m, const float* text_hidden, const float* text_mask, const float* audio_hidden,
                          const float* audio_mask, int B, int T, int S, float* logits_out, cudaStream_t st) {
  CK(ready(m, st));
  OptionsScope scope(&m->opt);
  if (m->dl.empty() || !m->dproj_w) { set_err("decoder_module tensors were not registered%s", ""); return CACO_ERR_STATE; }
  if (!text_hidden || !text_mask || !audio_hidden || !audio_mask || !logits_out || B <= 0 || T <= 0 || S <= 0) return CACO_ERR_ARG;
  const caco_config& c = m->cfg;
  const int D = c.hidden, F = c.ffn, V = m->dec_vocab, Vp = m->dec_vocab_pad;
  const size_t R = (size_t)B * T, RA = (size_t)B * S;
  if (R > (1u << 20) || RA > (1u << 22)) return CACO_ERR_ARG;
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += al256(bytes); return o; };
  const size_t o_x = carve(R * D * 4), o_x16 = carve(R * D * 2), o_a = carve(R * D * 4), o_a16 = carve(R * D * 2);
  const size_t o_c = carve(R * D * 4), o_c16 = carve(R * D * 2), o_qkv = carve(R * 3 * D * 2), o_att = carve(R * D * 2);
  const size_t o_q = carve(R * D * 2), o_mlp = carve(R * (size_t)F * 2), o_ah = carve(RA * D * 2), o_kv = carve(RA * 2 * D * 2);
  const size_t o_log = carve(R * (size_t)Vp * 4);
  CK(ensure_ws(m, 2, off));
  uint8_t* w = (uint8_t*)m->ws[2];
  float *x = (float*)(w + o_x), *a = (float*)(w + o_a), *cc = (float*)(w + o_c), *logp = (float*)(w + o_log);
  __half *x16 = (__half*)(w + o_x16), *a16 = (__half*)(w + o_a16), *c16 = (__half*)(w + o_c16), *qkv = (__half*)(w + o_qkv);
  __half *att = (__half*)(w + o_att), *q = (__half*)(w + o_q), *mlp = (__half*)(w + o_mlp), *ah16 = (__half*)(w + o_ah);
  __half* kv = (__half*)(w + o_kv);
  const int Ri = (int)R, RAi = (int)RA;
  cudaError_t e = cudaMemcpyAsync(x, text_hidden, R * D * 4, cudaMemcpyDeviceToDevice, st);
  if (e) return (int)e;
  CK(cast_f32_f16(text_hidden, x16, (int64_t)R * D, st));
  CK(cast_f32_f16(audio_hidden, ah16, (int64_t)RA * D, st));
  for (size_t i = 0; i < m->dl.size(); ++i) {
    const caco_model::DLayer& d = m->dl[i];
    const caco_model::TLayer& l = d.self_ffn;
    // self-attention block (causal + text padding, roberta.py:346-355)
    CK(lin(m, x16, D, l.qkv_w, D, l.qkv_b, nullptr, 0, qkv, 3 * D, Ri, 3 * D, CACO_EPI_BIAS_F16, st));
    CK(attention_text(qkv, text_mask, att, B, T, c.text_heads, D / c.text_heads, st));
    CK(lin(m, att, D, l.o_w, D, l.o_b, x, D, x, D, Ri, D, CACO_EPI_BIAS_RESID_F32, st));
    CK(layernorm(x, l.ln1_g, l.ln1_b, c.ln_eps, a, a16, Ri, D, st));
    // cross-attention to the audio tokens (roberta.py:205-211; key mask = audio mask, :358-361)
    CK(lin(m, a16, D, d.cq_w, D, d.cq_b, nullptr, 0, q, D, Ri, D, CACO_EPI_BIAS_F16, st));
    CK(lin(m, ah16, D, d.ckv_w, D, d.ckv_b, nullptr, 0, kv, 2 * D, RAi, 2 * D, CACO_EPI_BIAS_F16, st));
    CK(attention_cross(q, D, kv, audio_mask, att, B, T, S, c.text_heads, D / c.text_heads, st));
    CK(lin(m, att, D, d.co_w, D, d.co_b, a, D, a, D, Ri, D, CACO_EPI_BIAS_RESID_F32, st));
    CK(layernorm(a, d.lnc_g, d.lnc_b, c.ln_eps, cc, c16, Ri, D, st));
    // MLP
    CK(lin(m, c16, D, l.fc1_w, D, l.fc1_b, nullptr, 0, mlp, F, Ri, F, CACO_EPI_BIAS_GELU_F16, st));
    CK(lin(m, mlp, F, l.fc2_w, F, l.fc2_b, cc, D, cc, D, Ri, D, CACO_EPI_BIAS_RESID_F32, st));
    CK(layernorm(cc, l.ln2_g, l.ln2_b, c.ln_eps, x, x16, Ri, D, st));
  }
  // vocabulary projection into the padded buffer, then the [R, vocab] rows out (50265 is odd: no aligned row pitch exists)
  CK(lin(m, x16, D, m->dproj_w, D, m->dproj_b, nullptr, 0, logp, Vp, Ri, Vp, CACO_EPI_BIAS_F32, st));
  e = cudaMemcpy2DAsync(logits_out, (size_t)V * 4, logp, (size_t)Vp * 4, (size_t)V * 4, R, cudaMemcpyDeviceToDevice, st);
  return (int)e;
}

// ---------------------------------------------------------------------------------------- KV-cached captioning decode
// The reference's decode loop (eval_caco_torch.py:411-472) re-runs text tower + decoder on the whole prefix for every new
// token.  Both are causal, so the keys / values of earlier tokens never change: a step only has to push ONE token per sequence
// through the 12 text layers and the 4 decoder layers, attending to cached keys / values ("KV-cached
// sampling"; the JAX twin caches too, caco/caco.py:154-230).  The cache is caller-owned memory laid out by DecodeCache:
//   text layers       Lt x [batch, capacity, 2D] f16   k | v of every token pushed so far
//   decoder self      Ld x [batch, capacity, 2D] f16
//   decoder cross     Ld x [batch * S, 2D] f16         k | v of the audio tokens, computed once by decode_begin
//   key mask          [batch, capacity] f32            1 for the slots filled so far (what a causal mask reduces to for the
//                                                      newest query), audio mask [batch, S] f32 (copied: the caller's may go away)
// Same kernels and the same per-row arithmetic as the full-prefix path: masked slots contribute exp2(-inf) = 0 exactly.
struct DecodeCache {
  size_t self_bytes, cross_bytes, o_text, o_dself, o_cross, o_kmask, o_amask, total;
  DecodeCache(const caco_model* m, int B, int S, int cap) {
    const size_t D = m->cfg.hidden;
    self_bytes = al256((size_t)B * cap * 2 * D * sizeof(__half));
    cross_bytes = al256((size_t)B * S * 2 * D * sizeof(__half));
    size_t off = 0;
    o_text = off;  off += self_bytes * (size_t)m->cfg.text_layers;
    o_dself = off; off += self_bytes * m->dl.size();
    o_cross = off; off += cross_bytes * m->dl.size();
    o_kmask = off; off += al256((size_t)B * cap * sizeof(float));
    o_amask = off; off += al256((size_t)B * S * sizeof(float));
    total = off;
  }
};

static int decode_args_ok(caco_model* m, const void* cache, int B, int S, int cap) {
  if (m->dl.empty() || !m->dproj_w) { set_err("decoder_module tensors were not registered%s", ""); return CACO_ERR_STATE; }
  if (!cache || B <= 0 || S <= 0 || cap <= 0 || cap > m->cfg.max_pos || (size_t)B * S > (1u << 22) || B > (1 << 16))
    return CACO_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(cache) & 255) return CACO_ERR_ALIGN;
  return 0;
}

static int decode_begin(caco_model* m, void* cache, size_t cache_bytes, const float* audio_hidden, const float* audio_mask,
                        int B, int S, int cap, cudaStream_t st) {
  CK(ready(m, st));
  OptionsScope scope(&m->opt);
  CK(decode_args_ok(m, cache, B, S, cap));
  if (!audio_hidden || !audio_mask) return CACO_ERR_ARG;
  const DecodeCache L(m, B, S, cap);
  if (cache_bytes < L.total) return CACO_ERR_ARG;
  const int D = m->cfg.hidden;
  const size_t RA = (size_t)B * S;
  CK(ensure_ws(m, 2, al256(RA * D * sizeof(__half))));
  __half* ah16 = (__half*)m->ws[2];
  uint8_t* c = (uint8_t*)cache;
  // empty self caches (masked slots are multiplied by an exact 0, so they must hold finite numbers) and an all-closed key mask
  cudaError_t e = cudaMemsetAsync(c + L.o_text, 0, L.o_cross - L.o_text, st);
  if (!e) e = cudaMemsetAsync(c + L.o_kmask, 0, (size_t)B * cap * sizeof(float), st);
  if (!e) e = cudaMemcpyAsync(c + L.o_amask, audio_mask, RA * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (e) return (int)e;
  CK(cast_f32_f16(audio_hidden, ah16, (int64_t)RA * D, st));
  for (size_t i = 0; i < m->dl.size(); ++i)      // roberta.py:76-83 with key_value_states = the audio hidden state
    CK(lin(m, ah16, D, m->dl[i].ckv_w, D, m->dl[i].ckv_b, nullptr, 0, c + L.o_cross + i * L.cross_bytes, 2 * D, (int)RA, 2 * D,
           CACO_EPI_BIAS_F16, st));
  return 0;
}

// One token per sequence: ids [B] (the newest token), pos [B] (its position = number of tokens already cached).  Writes the
// next-token logits [B, vocab] (logits_out, may be NULL) and / or their arg-max (next_out int32 [B], may be NULL).
static int decode_step(caco_model* m, void* cache, const int64_t* ids, const int64_t* pos, int B, int S, int cap,
                       float* logits_out, int* next_out, cudaStream_t st) {
  CK(ready(m, st));
  OptionsScope scope(&m->opt);
  CK(decode_args_ok(m, cache, B, S, cap));
  if (!ids || !pos || (!logits_out && !next_out)) return CACO_ERR_ARG;
  const caco_config& cf = m->cfg;
  const DecodeCache L(m, B, S, cap);
  const int D = cf.hidden, F = cf.ffn, V = m->dec_vocab, Vp = m->dec_vocab_pad, H = cf.text_heads, dh = D / cf.text_heads;
  const size_t R = (size_t)B;
  size_t off = 0;
  auto carve = [&](size_t bytes) { size_t o = off; off += al256(bytes); return o; };
  const size_t o_x = carve(R * D * 4), o_x16 = carve(R * D * 2), o_a = carve(R * D * 4), o_a16 = carve(R * D * 2);
  const size_t o_c = carve(R * D * 4), o_c16 = carve(R * D * 2), o_qkv = carve(R * 3 * D * 2), o_att = carve(R * D * 2);
  const size_t o_q = carve(R * D * 2), o_mlp = carve(R * (size_t)F * 2), o_log = carve(R * (size_t)Vp * 4);
  CK(ensure_ws(m, 2, off));
  uint8_t* w = (uint8_t*)m->ws[2];
  float *x = (float*)(w + o_x), *a = (float*)(w + o_a), *cc = (float*)(w + o_c), *logp = (float*)(w + o_log);
  __half *x16 = (__half*)(w + o_x16), *a16 = (__half*)(w + o_a16), *c16 = (__half*)(w + o_c16), *qkv = (__half*)(w + o_qkv);
  __half *att = (__half*)(w + o_att), *q = (__half*)(w + o_q), *mlp = (__half*)(w + o_mlp);
  uint8_t* c = (uint8_t*)cache;
  float* kmask = (float*)(c + L.o_kmask);
  const float* amask = (const float*)(c + L.o_amask);

  // text tower on the new token (roberta.py:35-53, 191-215): position ids are the cache positions
  CK(text_embed_ln(ids, pos, m->word, m->pos, m->type0, m->te_g, m->te_b, cf.ln_eps, x, x16, B, 1, D, cf.vocab, cf.max_pos, st));
  // causal self-attention block against a cache: the new token's k | v go into slot pos[b], its query sees slots 0..pos[b]
  auto self_block = [&](const caco_model::TLayer& l, void* kv, bool open_slot) -> int {
    CK(lin(m, x16, D, l.qkv_w, D, l.qkv_b, nullptr, 0, qkv, 3 * D, B, 3 * D, CACO_EPI_BIAS_F16, st));
    CK(kv_append(qkv, kv, pos, open_slot ? kmask : nullptr, B, D, cap, st));
    CK(attention_cross(qkv, 3 * D, kv, kmask, att, B, 1, cap, H, dh, st));
    CK(lin(m, att, D, l.o_w, D, l.o_b, x, D, x, D, B, D, CACO_EPI_BIAS_RESID_F32, st));
    return layernorm(x, l.ln1_g, l.ln1_b, cf.ln_eps, a, a16, B, D, st);
  };
  // GELU MLP + LayerNorm of `in` (f32 + f16 copy) back into x / x16
  auto mlp_block = [&](const caco_model::TLayer& l, float* in, const __half* in16) -> int {
    CK(lin(m, in16, D, l.fc1_w, D, l.fc1_b, nullptr, 0, mlp, F, B, F, CACO_EPI_BIAS_GELU_F16, st));
    CK(lin(m, mlp, F, l.fc2_w, F, l.fc2_b, in, D, in, D, B, D, CACO_EPI_BIAS_RESID_F32, st));
    return layernorm(in, l.ln2_g, l.ln2_b, cf.ln_eps, x, x16, B, D, st);
  };
  for (int i = 0; i < cf.text_layers; ++i) {
    CK(self_block(m->tl[i], c + L.o_text + (size_t)i * L.self_bytes, i == 0));
    CK(mlp_block(m->tl[i], a, a16));
  }
  // captioning decoder on the text tower's hidden state (roberta.py:337-373)
  for (size_t i = 0; i < m->dl.size(); ++i) {
    const caco_model::DLayer& d = m->dl[i];
    CK(self_block(d.self_ffn, c + L.o_dself + i * L.self_bytes, false));
    CK(lin(m, a16, D, d.cq_w, D, d.cq_b, nullptr, 0, q, D, B, D, CACO_EPI_BIAS_F16, st));
    CK(attention_cross(q, D, c + L.o_cross + i * L.cross_bytes, amask, att, B, 1, S, H, dh, st));
    CK(lin(m, att, D, d.co_w, D, d.co_b, a, D, a, D, B, D, CACO_EPI_BIAS_RESID_F32, st));
    CK(layernorm(a, d.lnc_g, d.lnc_b, cf.ln_eps, cc, c16, B, D, st));
    CK(mlp_block(d.self_ffn, cc, c16));
  }
  CK(lin(m, x16, D, m->dproj_w, D, m->dproj_b, nullptr, 0, logp, Vp, B, Vp, CACO_EPI_BIAS_F32, st));
  if (next_out) CK(topk_rows(logp, B, V, Vp, 1, next_out, nullptr, st));
  if (logits_out) {
    cudaError_t e = cudaMemcpy2DAsync(logits_out, (size_t)V * 4, logp, (size_t)Vp * 4, (size_t)V * 4, R, cudaMemcpyDeviceToDevice, st);
    if (e) return (int)e;
  }
  return 0;
}

}  // namespace caco

extern "C" {

size_t caco_model_decode_cache_bytes(const caco_model* m, int batch, int S, int capacity) {
  if (!m || !m->packed || m->dl.empty() || batch <= 0 || S <= 0 || capacity <= 0) return 0;
  return caco::DecodeCache(m, batch, S, capacity).total;
}
int caco_model_decode_begin(caco_model* m, void* cache, size_t cache_bytes, const float* audio_hidden, const float* audio_mask,
                            int batch, int S, int capacity, void* stream) {
  if (!m) return CACO_ERR_ARG;
  return caco::decode_begin(m, cache, cache_bytes, audio_hidden, audio_mask, batch, S, capacity, (cudaStream_t)stream);
}
int caco_model_decode_step(caco_model* m, void* cache, const int64_t* ids, const int64_t* positions, int batch, int S,
                           int capacity, float* logits_out, int* next_out, void* stream) {
  if (!m) return CACO_ERR_ARG;
  return caco::decode_step(m, cache, ids, positions, batch, S, capacity, logits_out, next_out, (cudaStream_t)stream);
}

int caco_model_decoder_logits(caco_model* m, const float* text_hidden, const float* text_mask, const float* audio_hidden,
                              const float* audio_mask, int batch, int T, int S, float* logits_out, void* stream) {
  if (!m) return CACO_ERR_ARG;
  return caco::decoder_logits(m, text_hidden, text_mask, audio_hidden, audio_mask, batch, T, S, logits_out, (cudaStream_t)stream);
}
int caco_model_decoder_vocab(const caco_model* m) { return (m && m->packed) ? m->dec_vocab : 0; }

const char* caco_last_error(void) { return caco::g_err; }

int caco_model_create(const caco_config* cfg, caco_model** out) {
  if (!cfg || !out) return CACO_ERR_ARG;
  if (cfg->hidden % 128 || cfg->hidden > 1024 || cfg->hidden % cfg->audio_heads || cfg->hidden % cfg->text_heads ||
      cfg->hidden % cfg->pool_heads || cfg->pool_heads > 8)      // caco/load_model.py:47 uses 8 pooler heads, caco.py:292 uses 2
    return CACO_ERR_ARG;
  const int adh = cfg->hidden / cfg->audio_heads;
  if ((adh != 96 && adh != 64) || cfg->hidden / cfg->text_heads != 64) return CACO_ERR_ARG;
  caco_model* m = new caco_model();
  m->cfg = *cfg;
  if (m->cfg.audio_ln_eps <= 0.f) m->cfg.audio_ln_eps = 1e-5f;   // nn.LayerNorm's default (mae.py:68,76,123)
  m->opt = caco::g_default_opts;
  *out = m;
  return 0;
}

void caco_model_destroy(caco_model* m) {
  if (!m) return;
  cudaDeviceSynchronize();
  if (m->arena16) cudaFree(m->arena16);
  if (m->arena32) cudaFree(m->arena32);
  for (int i = 0; i < 3; ++i)
    if (m->ws[i]) cudaFree(m->ws[i]);
  delete m;
}

int caco_model_set_tensor(caco_model* m, const char* key, const float* dev_ptr, int64_t numel) {
  if (!m || !key || !dev_ptr || numel <= 0) return CACO_ERR_ARG;
  m->w[key] = caco_model::T{dev_ptr, numel};     // decoder_module.* included: packed when the whole head is registered
  m->packed = false;
  return 0;
}

int caco_model_pack(caco_model* m, void* stream) {
  if (!m) return CACO_ERR_ARG;
  return caco::pack(m, (cudaStream_t)stream);
}

int caco_model_set_option(caco_model* m, const char* name, int value) {
  if (!m) return CACO_ERR_ARG;
  return caco::set_option(m->opt, name, value);
}

uint64_t caco_model_generation(const caco_model* m) { return m ? m->generation : 0; }

int caco_model_audio_embedding(caco_model* m, const float* patches, const float* time_inds, const float* freq_inds,
                               const float* mask, int batch, int seq, int normalize, float* emb_out, float* hidden_out,
                               void* stream) {
  return caco::audio_embedding(m, patches, nullptr, time_inds, freq_inds, mask, batch, seq, normalize, emb_out, hidden_out,
                               (cudaStream_t)stream);
}

int caco_model_text_embedding(caco_model* m, const int64_t* ids, const float* mask, const int64_t* position_ids, int batch,
                              int T, int normalize, float* emb_out, float* hidden_out, void* stream) {
  return caco::text_embedding(m, ids, mask, position_ids, batch, T, normalize, emb_out, hidden_out, (cudaStream_t)stream);
}

// waveform-in path: frontend (uniform or ragged) + audio tower.  The frontend outputs live in a side allocation so the
// tower's workspace carve-up stays independent; the tower only needs the fp16 operand copy of the patches, so the fp32
// patches are never written on this path.
static int encode_audio_impl(caco_model* m, const float* wave, const int* lengths, int batch, int stride, int max_patches,
                             int normalize, float* emb_out, float* hidden_out, float* mask_out, cudaStream_t st) {
  CK(caco::ready(m, st));
  caco::OptionsScope scope(&m->opt);
  if (!wave || !emb_out || batch <= 0 || stride <= 0 || max_patches <= 0) return CACO_ERR_ARG;
  const size_t R = (size_t)batch * max_patches;
  uint8_t* buf = nullptr;
  const size_t p16_bytes = (R * 256 * sizeof(__half) + 255) & ~(size_t)255;
  cudaError_t e = cudaMallocAsync((void**)&buf, p16_bytes + 3 * R * sizeof(float), st);
  if (e) return (int)e;
  __half* p16 = (__half*)buf;
  float *ti = (float*)(buf + p16_bytes), *fi = ti + R, *mk = fi + R;
  int rc = caco::frontend(wave, lengths, batch, stride, max_patches, nullptr, p16, ti, fi, mk, nullptr, st);
  if (!rc) rc = caco::audio_embedding(m, nullptr, p16, ti, fi, mk, batch, max_patches, normalize, emb_out, hidden_out, st);
  if (!rc && mask_out) rc = (int)cudaMemcpyAsync(mask_out, mk, R * sizeof(float), cudaMemcpyDeviceToDevice, st);
  cudaFreeAsync(buf, st);
  return rc;
}

int caco_model_encode_audio(caco_model* m, const float* wave, int batch, int n_samples, int max_patches, int normalize,
                            float* emb_out, void* stream) {
  return encode_audio_impl(m, wave, nullptr, batch, n_samples, max_patches, normalize, emb_out, nullptr, nullptr,
                           (cudaStream_t)stream);
}

int caco_model_encode_audio_ex(caco_model* m, const float* wave, const int* lengths, int batch, int stride,
                               int max_patches, int normalize, float* emb_out, float* hidden_out, float* mask_out,
                               void* stream) {
  return encode_audio_impl(m, wave, lengths, batch, stride, max_patches, normalize, emb_out, hidden_out, mask_out,
                           (cudaStream_t)stream);
}

const float* caco_model_logit_scale(caco_model* m) { return m ? m->logit_scale : nullptr; }
}
