/* caco_b200.h — C ABI of libcaco_b200.so, the sm_100a implementation of Cacophony's inference hot path.
 *
 * The reference (gzhu06/Cacophony, src/caco_torch + the frontend functions of
 * src/eval/eval_caco_torch.py) is pure Python over torch ops and has no FFI layer of its own; this
 * header is therefore the boundary a reference maintainer would bind with ctypes (INTEGRATION.md shows
 * the stub).  Every entry point cites the reference code it replaces (paths relative to the reference
 * repo root).  Conventions:
 *   - plain C: raw DEVICE pointers + sizes + a cudaStream_t passed as void*; no torch types;
 *   - the caller owns all memory; the library only borrows pointers for the duration of the call
 *     (work is enqueued on `stream`, so buffers must stay alive until the stream has drained);
 *   - all matrices are dense row-major; "f16" = IEEE binary16, "f32" = binary32;
 *   - return 0 on success, a CACO_ERR_* (negative) or a cudaError_t (positive) otherwise; nothing
 *     throws across the boundary and there is NO CPU fallback.
 */
#ifndef CACO_B200_H_
#define CACO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CACO_ERR_ARG (-1)     /* bad shape / null pointer / unsupported size */
#define CACO_ERR_ALIGN (-2)   /* pointer or leading dimension not 16-byte aligned */
#define CACO_ERR_DRIVER (-3)  /* cuTensorMapEncodeTiled unavailable or failed */
#define CACO_ERR_STATE (-4)   /* model handle used before weights were packed, etc. */

/* GEMM epilogues (what is fused after A·Wᵀ) */
#define CACO_EPI_BIAS_F16 0        /* out_f16 = acc + bias                      (QKV projections)            */
#define CACO_EPI_BIAS_SILU_F16 1   /* out_f16 = silu(acc + bias)                (mae.py:56-57 fc1+SiLU)      */
#define CACO_EPI_BIAS_GELU_F16 2   /* out_f16 = gelu_erf(acc + bias)            (roberta.py:156-157)         */
#define CACO_EPI_BIAS_F32 3        /* out_f32 = acc + bias                      (input_proj mae.py:133)      */
#define CACO_EPI_BIAS_RESID_F32 4  /* out_f32 = acc + bias + resid_f32          (mae.py:93,97; roberta.py:122,176) */

/* GEMM kernel variants (0 = library default) */
#define CACO_GEMM_CG1_N256 1  /* one CTA per 128x256 tile, tcgen05.mma.cta_group::1 */
#define CACO_GEMM_CG1_N128 2  /* one CTA per 128x128 tile */
#define CACO_GEMM_CG2_N256 3  /* CTA pair per 256x256 tile, tcgen05.mma.cta_group::2 */
#define CACO_GEMM_CG2_N256_E16 4 /* same, 16 epilogue warps / 4 smem stages (activation-heavy epilogues) */

/* Library/version probe; also reports the compute capability the kernels were built for (100). */
int caco_version(void);
int caco_built_arch(void);

/* ---- K1: waveform -> log-mel -> 16x16 patches  (src/eval/eval_caco_torch.py:41-151, :181-206) ----
 * wave        [batch, n_samples] f32 (all clips the same length; ragged batches = one call per length)
 * patches     [batch, max_patches, 256] f32   token p = 8*t+f, element dt*16+df = mel[16t+dt, 16f+df]
 * patches_f16 same layout in f16 or NULL (operand copy for the input-projection GEMM)
 * time_inds / freq_inds / mask  [batch, max_patches] f32 (padded slots: index 0, mask 0)
 * log_mel     optional [batch, ceil(n/160), 128] f32 (compute_mel_spectrogram output) or NULL          */
int caco_frontend(const float* wave, int batch, int n_samples, int max_patches, float* patches, void* patches_f16,
                  float* time_inds, float* freq_inds, float* mask, float* log_mel, void* stream);

/* Ragged batch: clip b = wave[b*stride : b*stride + lengths[b]], lengths = DEVICE int32 [batch].  Each
 * clip gets eval_caco_torch.py:181-206's per-clip treatment (own frame count ceil(L/160), own valid-patch count
 * floor(frames/16)*8 truncated to max_patches, zero padding rows, mask) in one launch. */
int caco_frontend_ragged(const float* wave, const int* lengths, int batch, int stride, int max_patches, float* patches,
                         void* patches_f16, float* time_inds, float* freq_inds, float* mask, void* stream);

/* ---- K2: out = epilogue(A[M,K] f16 · W[N,K]ᵀ f16), fp32 accumulate on tcgen05 tensor cores.
 * Replaces every nn.Linear on the path (mae.py:51-52,69,116; roberta.py:62-64,110,153,164; caco.py:35,37).
 * lda/ldw/ldo/ldr are leading dimensions in ELEMENTS.  N % 4 == 0, K % 8 == 0.                       */
int caco_gemm_f16(const void* A, int lda, const void* W, int ldw, const float* bias, const float* resid, int ldr,
                  void* out, int ldo, int M, int N, int K, int epi, int variant, void* stream);
/* Split-weight form (precision escape hatch): W2 [N, 2K] f16 holds fp16(w) | fp16(w - fp16(w)) per row; the
 * kernel accumulates A·hi^T + A·lo^T in one pass over a 2K-long reduction (A's k-blocks are re-read), so the weights enter
 * with ~22 mantissa bits.  Twice the tensor work of caco_gemm_f16. */
int caco_gemm_f16_wsplit(const void* A, int lda, const void* W2, int ldw, const float* bias, const float* resid, int ldr,
                         void* out, int ldo, int M, int N, int K, int epi, int variant, void* stream);
/* f32 [rows, K] -> f16 [rows, 2K] hi | lo (the packing caco_gemm_f16_wsplit reads). */
int caco_cast_f32_f16_split(const float* src, void* dst, int64_t rows, int64_t K, void* stream);
/* Execution options.  Every model handle owns a set (caco_model_set_option); op-level calls made outside a handle use the
 * library defaults changed here.  Names (values): "pdl" (1 = programmatic dependent launch of the tower kernels: a kernel's
 * prologue overlaps its predecessor's tail, griddepcontrol.wait before touching data), "gemm_variant" (0 = auto, CACO_GEMM_*),
 * "resid_red" (1 = in-place residual GEMMs add through the L2 with red.global.add.v4.f32, 0 = load/add/store in the SM; same
 * fp32 result), "audio_chunk_rows" / "text_chunk_rows" (token rows per pass of a tower), "split_weights" (0/1: GEMM
 * weights as fp16 hi + lo, see caco_gemm_f16_wsplit).  Returns 0, or CACO_ERR_ARG for an unknown name / bad value. */
int caco_set_default_option(const char* name, int value);
/* fp16 range guard: number of 4-element fp16 stores of GEMM epilogues that had to clamp a value to +-65504 since the last
 * reset (0 = every operand copy was in range).  Synchronises the device. */
unsigned int caco_saturation_count(int reset);
/* live profiling for bench.py: CUDA events around every GEMM launch on its stream.  caco_gemm_profile(1) resets and
 * starts recording; caco_gemm_profile_read synchronises the device and returns the launch count, summed device
 * time (ms) and summed algorithmic FLOPs (2*M*N*K). */
void caco_gemm_profile(int enable);
int caco_gemm_profile_read(double* total_ms, double* total_flops);

/* f32 -> f16 round-to-nearest copy (weight packing / operand copies). */
int caco_cast_f32_f16(const float* src, void* dst, int64_t n, void* stream);

/* ---- K4: LayerNorm over the last dim (nn.LayerNorm eps=1e-5: mae.py:68,76,123; roberta.py:32,111,165).
 * y = (x-mean)/sqrt(var+eps)*gamma+beta ; writes y as f32 (out_f32, may be NULL) and/or f16 (out_f16, may be NULL). */
int caco_layernorm(const float* x, const float* gamma, const float* beta, float eps, float* out_f32, void* out_f16,
                   int rows, int dim, void* stream);

/* ---- audio position embedding (mae.py:102-109,135-142): x[m,:] += cat[sin(t·w), cos(t·w)] + freq_emb[f[m]]. */
int caco_audio_add_pos(float* x, const float* time_inds, const float* freq_inds, const float* freq_emb, int n_freq,
                       int rows, int dim, void* stream);

/* ---- K3a: audio self-attention, nn.MultiheadAttention semantics (mae.py:69-74,89-92):
 * qkv [batch*seq, 3*heads*dh] f16 (q|k|v packed in-proj output), q scaled by 1/sqrt(dh) inside,
 * keys with mask==0 get -inf, softmax in fp32, out [batch*seq, heads*dh] f16.  dh = 96 (seq <= 4096): the persistent
 * ping-pong tcgen05 kernel (attention_pp.cu); dh = 64, or longer sequences: the warp-level flash kernel (attention.cu). */
int caco_attention_audio(const void* qkv, const float* mask, void* out, int batch, int seq, int heads, int dh,
                         void* stream);


/* ---- K3b: causal text self-attention (roberta.py:86-102, mask from roberta.py:297-310):
 * qkv [batch*T, 3*heads*64] f16, key_mask [batch, T] f32 (1 = keep); allowed(i,j) = j<=i && key_mask[j]. */
int caco_attention_text(const void* qkv, const float* key_mask, void* out, int batch, int T, int heads, int dh,
                        void* stream);

/* ---- K5: RoBERTa embeddings + LayerNorm (roberta.py:35-53): LN(word[id] + pos[pid] + type[0]).
 * position_ids may be NULL (= arange(T), roberta.py:292-293).                                          */
int caco_text_embed_ln(const int64_t* ids, const int64_t* position_ids, const float* word, const float* pos,
                       const float* type0, const float* gamma, const float* beta, float eps, float* out_f32,
                       void* out_f16, int batch, int T, int dim, int vocab, int max_pos, void* stream);

/* ---- K6: masked single-query attention pooling over tokens.
 * scores[b,h,j] = dot(u[h,:], hid[b,j,:]) + c[h]  (the pooler's key projection folded into u, c at pack time),
 * w = softmax_j(mask(scores)), pooled[b,h,:] = sum_j w[b,h,j] * hid[b,j,:].
 * If ln_gamma != NULL the rows of `hid` are LayerNorm-ed on the fly (final encoder LN, mae.py:147) and the
 * normalised rows are also written to hid_out (may be NULL).
 * Used for AudioAttentionPooler (caco.py:41-79, heads=2) and AttentionPooler (roberta.py:253-271, heads=1). */
int caco_attn_pool(const float* hid, const float* mask, const float* u, const float* c, const float* ln_gamma,
                   const float* ln_beta, float ln_eps, float* hid_out, float* pooled, int batch, int seq, int heads,
                   int dim, void* stream);

/* ---- small fp32 GEMM for the [batch, *] tails (pooler value/out projections, text_proj, similarity):
 * out[M,N] = alpha * A[M,K] · W[N,K]ᵀ + bias   (fp32 FMA, exact-order independent of tensor cores). */
int caco_sgemm_nt(const float* A, int lda, const float* W, int ldw, const float* bias, float alpha, float* out,
                  int ldo, int M, int N, int K, void* stream);

/* ---- K7: e / ||e + 1e-10||_2 per row (caco.py:146,173). in == out allowed. */
int caco_l2norm(const float* in, float* out, int rows, int dim, float eps, void* stream);

/* ---- K7 fused with the path's one exchange: out-of-place L2 normalisation whose result rows are stored
 * straight into every rank's gathered embedding matrix through NVLink peer mappings, then a release flag per peer.
 * peer_base_dev: DEVICE array [world] of base pointers of the ranks' symmetric buffers (this rank's own included); the rows go
 * to peer_base[p] + dst_byte_offset (the caller puts this rank's row block there), the flag to the uint32 array at
 * peer_base[p] + flag_byte_offset, slot flag_index, value `epoch` (monotonically increasing per slot).  ticket: one zeroed
 * device uint32 per concurrently used stream.  world <= 16. */
int caco_l2norm_scatter(const float* in, int rows, int dim, float eps, void* const* peer_base_dev, long long dst_byte_offset,
                        long long flag_byte_offset, int flag_index, unsigned int epoch, int world, unsigned int* ticket,
                        void* stream);
/* One-CTA kernel that returns once flags[0..n) >= epoch (acquire at system scope): what the similarity launch is ordered
 * behind.  After timeout_ms (default 10 s) it gives up and writes 1 + the late slot to *status (DEVICE int, may be NULL)
 * instead of hanging the GPU. */
int caco_wait_flags(const unsigned int* flags, int n, unsigned int epoch, int timeout_ms, int* status, void* stream);

/* ---- K8: at = exp(logit_scale)·A·Tᵀ, ta = atᵀ (caco.py:208-210).  A [na,dim], T [nt,dim] f32.
 * logit_scale is a DEVICE pointer to the scalar parameter.  ta may be NULL.                          */
int caco_sim_logits(const float* a, const float* t, const float* logit_scale, float* at, float* ta, int na, int nt,
                    int dim, void* stream);

/* ---- row (f-2): the device side of the evaluation drivers (eval_caco_torch.py:289-408, eval_utils.py:18-66).
 * caco_topk_rows: idx_out[r, :k] = argsort(-x[r, :])[:k] (ties: lower column first, NaN last), k <= 32; val_out may be NULL.
 * Replaces torch.argsort(-logits, dim=-1) at eval_caco_torch.py:331,401,406 (only the first k <= 10 ranks are ever used). */
int caco_topk_rows(const float* x, int rows, int cols, int ldx, int k, int* idx_out, float* val_out, void* stream);
/* caco_retrieval_hits: `preds` of compute_retrieval_metric (eval_utils.py:26-41) for every query: out[q] bit j = rank j+1 hit.
 * topk [n_queries, ldk >= 10] int32 key indices; key_id [n_keys] int32; gt_id [n_queries] int32.
 * mode 0 ('ta'): hit = key_id[idx] == gt_id[q].   mode 1 ('at'): hit = (gt_id[q] * n_key_ids + key_id[idx]) is in the sorted
 * int64 array gt_pairs AND that key_id was not already counted at an earlier rank. */
int caco_retrieval_hits(const int* topk, int ldk, int n_queries, const int* key_id, const int* gt_id,
                        const long long* gt_pairs, int n_pairs, long long n_key_ids, int mode, int* out, void* stream);
/* ---- row (f-4): HEAR timestamp embeddings (caco_embeddings.py:124-129): out[b, t, :] = mean of hid[b, t*group .. +group-1, :],
 * t < seq / group ('VALID'). */
int caco_avg_pool_tokens(const float* hid, int batch, int seq, int dim, int group, float* out, void* stream);

/* ======================================= model handle ==========================================
 * A handle owns fp16-packed copies of the encoder weights, folded pooler vectors and an activation
 * workspace (cudaMalloc); inputs/outputs stay caller-owned.  Mirrors CACO (src/caco_torch/caco.py:82-261). */
typedef struct caco_model caco_model;

typedef struct {
  int hidden;        /* 768  */
  int ffn;           /* 3072 */
  int patch_dim;     /* 256  */
  int audio_layers;  /* 12   */
  int audio_heads;   /* 8    */
  int n_freq;        /* 8    */
  int pool_heads;    /* 2  (CACOConfig.num_attention_pool_heads, caco.py:20) */
  int text_layers;   /* 12   */
  int text_heads;    /* 12   */
  int vocab;         /* 50265 */
  int max_pos;       /* 514  */
  float ln_eps;      /* 1e-5  RobertaConfig.layer_norm_eps (roberta.py:22) */
  float audio_ln_eps; /* 1e-5 nn.LayerNorm default of the audio tower (mae.py:68,76,123); <= 0 means 1e-5 */
} caco_config;

int caco_model_create(const caco_config* cfg, caco_model** out);
void caco_model_destroy(caco_model* m);
/* Register one f32 DEVICE tensor of the reference state_dict by its key, e.g.
 * "audio_module.layers.3.mlp.fc1.weight".  The pointer must stay valid until caco_model_pack returns
 * for GEMM weights (they are copied to f16) and for the model's lifetime for everything else
 * (biases, LayerNorm params, embeddings are used in place).  decoder_module.* tensors are optional: when the whole captioning
 * head is registered it is packed too (caco_model_decoder_logits). */
int caco_model_set_tensor(caco_model* m, const char* key, const float* dev_ptr, int64_t numel);
int caco_model_pack(caco_model* m, void* stream);
/* Per-handle execution option (names as in caco_set_default_option; a new handle starts from the library defaults).
 * "split_weights" takes effect at the next call (the handle re-packs its weight arena from the registered tensors). */
int caco_model_set_option(caco_model* m, const char* name, int value);
/* Counter bumped whenever the handle releases device memory a previously enqueued or captured call may reference (workspace
 * growth, re-pack, move to another device): a CUDA graph captured from this handle is stale once it changes. */
uint64_t caco_model_generation(const caco_model* m);

/* CACO.get_audio_embedding (caco.py:123-150).  hidden_out [batch, seq, hidden] f32 or NULL. */
int caco_model_audio_embedding(caco_model* m, const float* patches, const float* time_inds, const float* freq_inds,
                               const float* mask, int batch, int seq, int normalize, float* emb_out, float* hidden_out,
                               void* stream);
/* CACO.get_text_embedding (caco.py:152-177).  mask [batch, T] f32; position_ids may be NULL. */
int caco_model_text_embedding(caco_model* m, const int64_t* ids, const float* mask, const int64_t* position_ids,
                              int batch, int T, int normalize, float* emb_out, float* hidden_out, void* stream);
/* CACO.get_decoder_logits minus the text tower (caco.py:214-240, RobertaDecoder.forward roberta.py:337-373; SURVEY.md 8 row
 * f-4): text_hidden [batch, T, hidden] f32 = the text tower's final hidden state (hidden_out of caco_model_text_embedding),
 * audio_hidden [batch, S, hidden] f32 = the audio tower's hidden_out, masks f32 (1 = keep).  Causal self-attention, cross-attention
 * to the audio tokens, GELU MLP per layer (post-LN), then decoder_proj.  logits_out [batch, T, vocab] f32.  Needs the
 * decoder_module.* tensors registered before caco_model_pack (CACO_ERR_STATE otherwise).  caco_model_decoder_vocab: the
 * vocabulary size of the packed head (0 = none). */
int caco_model_decoder_logits(caco_model* m, const float* text_hidden, const float* text_mask, const float* audio_hidden,
                              const float* audio_mask, int batch, int T, int S, float* logits_out, void* stream);
int caco_model_decoder_vocab(const caco_model* m);
/* KV-cached captioning decode ("KV-cached sampling").  The reference's loop (eval_caco_torch.py:411-472)
 * re-runs CACO.get_decoder_logits on the whole prefix for every token; text tower and decoder are causal, so a step here pushes
 * ONE new token per sequence through them against cached keys / values (what the JAX twin does, caco/caco.py:154-230) — the same
 * next-token logits as the full-prefix call.
 *   cache     caller-owned DEVICE buffer of caco_model_decode_cache_bytes(m, batch, S, capacity) bytes, 256-byte aligned;
 *             capacity = the longest sequence (prompt + generated tokens) it can hold, <= max_pos
 *   begin     empties the cache and stores the cross-attention keys / values of audio_hidden [batch, S, hidden] f32 (the audio
 *             tower's hidden_out) and a copy of audio_mask [batch, S] f32
 *   step      ids [batch] int64 = each sequence's newest token, positions [batch] int64 = its index in the sequence (DEVICE
 *             arrays, so a step can be captured in a CUDA graph and replayed); logits_out [batch, vocab] f32 and / or next_out
 *             [batch] int32 = arg-max of the logits (either may be NULL).  Steps must be issued in position order from 0; a
 *             token whose position is outside [0, capacity) is not cached (nothing is written out of bounds). */
size_t caco_model_decode_cache_bytes(const caco_model* m, int batch, int S, int capacity);
int caco_model_decode_begin(caco_model* m, void* cache, size_t cache_bytes, const float* audio_hidden, const float* audio_mask,
                            int batch, int S, int capacity, void* stream);
int caco_model_decode_step(caco_model* m, void* cache, const int64_t* ids, const int64_t* positions, int batch, int S,
                           int capacity, float* logits_out, int* next_out, void* stream);
/* cross-attention core used by it: q [batch*Tq, heads*64] f16 (row pitch ldq), kv [batch*Skv, 2*heads*64] f16 (k | v),
 * key_mask [batch, Skv] f32, out [batch*Tq, heads*64] f16 (roberta.py:76-102 with key_value_states). */
int caco_attention_cross(const void* q, int ldq, const void* kv, const float* key_mask, void* out, int batch, int Tq, int Skv,
                         int heads, int dh, void* stream);
/* waveform-in convenience: frontend + get_audio_embedding in one call (encode_audio). */
int caco_model_encode_audio(caco_model* m, const float* wave, int batch, int n_samples, int max_patches, int normalize,
                            float* emb_out, void* stream);
/* general form: lengths (DEVICE int32 [batch]) or NULL for a uniform batch of `stride` samples; hidden_out
 * [batch, max_patches, hidden] f32 (the LayerNorm-ed encoder output, what HEAR timestamp embeddings pool) or NULL;
 * mask_out [batch, max_patches] f32 or NULL. */
int caco_model_encode_audio_ex(caco_model* m, const float* wave, const int* lengths, int batch, int stride,
                               int max_patches, int normalize, float* emb_out, float* hidden_out, float* mask_out,
                               void* stream);
/* DEVICE pointer to the registered logit_scale scalar (caco.py:116), for caco_sim_logits. */
const float* caco_model_logit_scale(caco_model* m);
/* number of kernels the library has launched since load (bench.py's gpu_launches). */
int64_t caco_launch_count(void);
/* last error detail of this thread (e.g. which state_dict key was missing), "" if none. */
const char* caco_last_error(void);
/* host-only: the fp32 HTK mel filterbank the frontend uses, out[257][128] (no GPU needed). */
int caco_mel_filterbank(float* out_257x128);

#ifdef __cplusplus
}
#endif
#endif /* CACO_B200_H_ */
