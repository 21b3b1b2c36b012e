/* teochat_b200 — C-ABI of the B200-native TEOChat inference hot path.
 *
 * The reference (ermongroup/TEOChat) has no FFI: its hot path is Python calling
 * transformers==4.31 / PyTorch.  This header is the boundary a maintainer binds
 * instead (ctypes stub in INTEGRATION.md): every entry point names the reference code it
 * replaces (paths relative to the reference root).
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name says `host`; tensors are dense row-major;
 *     bf16 = 16-bit brain float; `stream` is a cudaStream_t passed as void*.
 *   - calls only enqueue work on `stream` (async w.r.t. the host); the caller synchronises.
 *   - nothing here allocates or frees caller memory: scratch is passed in, sized by the
 *     matching `*_workspace_bytes` query.
 *   - return value: TEO_OK (0) or a negative TEO_ERR_*; text via teo_last_error() (thread-local).
 *   - a handle is bound to one device and one host thread (one process per GPU).
 */
#ifndef TEOCHAT_B200_H_
#define TEOCHAT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TEO_OK 0
#define TEO_ERR_BAD_ARG (-1)
#define TEO_ERR_CUDA (-2)
#define TEO_ERR_WORKSPACE (-3)
#define TEO_ERR_UNSUPPORTED (-4)

#define TEO_ACT_NONE 0
#define TEO_ACT_QUICK_GELU 1 /* x*sigmoid(1.702x): CLIP default, configuration_image.py:191 */
#define TEO_ACT_GELU 2       /* erf GELU: nn.GELU(), multimodal_projector/builder.py:44 */
/* SwiGLU fused into the gate/up projection (HF LlamaMLP: act_fn(gate_proj(x)) * up_proj(x)): W [N,K] holds the gate and
 * up rows interleaved in blocks of 32 (rows 64b..64b+31 = gate rows 32b.., rows 64b+32..64b+63 = up rows 32b..), and C
 * gets N/2 columns: C[m, 32b+j] = bf16(silu(bf16(g)) * bf16(u)).  Tiled schedule only (M > 128), bf16 output,
 * N % 128 == 0, no bias / residual. */
#define TEO_ACT_SWIGLU_PAIRS 3

#define TEO_IMAGE_TOKEN_INDEX (-200) /* videollava/constants.py:9 */

typedef struct teo_handle teo_handle;

/* ---- lifecycle ------------------------------------------------------------------------- */
int teo_create(int device_id, teo_handle** out);
int teo_destroy(teo_handle* h);
const char* teo_last_error(void);
int teo_abi_version(void);
/* digest of the sources this binary was compiled from (teochat_b200/build.py:source_digest); the Python loader refuses a
 * library whose digest differs from the sources beside it */
const char* teo_build_digest(void);
/* kernels launched through this handle so far (bench.py's gpu_launches) */
unsigned long long teo_launch_count(const teo_handle* h);

/* ---- synthetic ("random-init") parameters ---------------------------------------------- */
/* value_i = mean + (sum of the four u16 fields of splitmix64(seed + (i+1)*golden) - 131070) * scale;
 * replaces the initialisers at languagebind/image/modeling_image.py:179-230 and HF Llama
 * _init_weights with a device-reproducible stream.  out: bf16 [n]. */
int teo_init_normal_hash_bf16(void* out, size_t n, uint64_t seed, float scale, float mean, void* stream);
int teo_init_normal_hash_f32(void* out, size_t n, uint64_t seed, float scale, float mean, void* stream);
/* out: u8 [n] = top byte of splitmix64(seed + (i+1)*golden) — synthetic frames */
int teo_init_u8_hash(void* out, size_t n, uint64_t seed, void* stream);

/* ---- GEMM core ------------------------------------------------------------------------- */
/* C[M,N] = act(A[M,K] · W[N,K]^T + bias[N]) + residual[M,N]   (nn.Linear semantics; replaces the
 * cuBLAS calls under every nn.Linear of HF CLIPAttention/CLIPMLP (modeling_image.py:11-12),
 * the projector (multimodal_projector/builder.py:41-48) and HF LlamaAttention/LlamaMLP/lm_head
 * (llava_llama.py:48,88)).  A, W, bias, residual bf16; fp32 accumulation in TMEM; C bf16, or
 * fp32 when out_fp32 != 0.  lda/ldw/ldc/ldr in elements; K % 8 == 0, N % 8 == 0, lda % 8 == 0.
 * bias / residual may be NULL.  residual may alias C. */
size_t teo_gemm_workspace_bytes(int M, int N, int K);
int teo_gemm_bf16(teo_handle* h, const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N,
                  int K, const void* bias, const void* residual, int ldr, int act, int out_fp32, void* workspace,
                  size_t workspace_bytes, void* stream);

/* The same GEMM with the full option set.  LayerNorm FOLDED into the linear that consumes it (HF CLIPEncoderLayer:
 * layer_norm1 -> q/k/v proj, layer_norm2 -> fc1, modeling_image.py:136-151): A holds the UN-normalised rows x [M,K], W the
 * weights pre-multiplied by the LayerNorm gain, W'[n,k] = bf16(gamma_k * W[n,k]), and
 *     C[m,n] = act( rstd_m * (sum_k x[m,k] W'[n,k] - mean_m * ln_c[n]) + ln_bias[n] )
 * with ln_c[n] = sum_k W'[n,k] and ln_bias[n] = sum_k beta_k W[n,k] + b[n] (both f32 [N], 16-byte aligned) — i.e.
 * act(LayerNorm(x) W^T + b) without the normalised tensor ever reaching HBM.  mean_m / rstd_m come from ln_stats, f32
 * [M][ln_slots][2] partial (sum x, sum x^2) per row, which the GEMM that PRODUCED x writes when given stats_out (f32
 * [M][teo_gemm_stats_slots(M,N,K)][2], statistics of the bf16 values it stores), or teo_row_stats computes.  Tiled schedule
 * only (M > 128), bf16 output, bias == NULL together with ln_stats.  Zero-initialise the struct for plain behaviour. */
typedef struct {
    const void* bias;      /* bf16 [N] or NULL */
    const void* residual;  /* bf16 [M, ldr] or NULL */
    int ldr, act, out_fp32, w_blocked;
    const void* ln_stats;  /* f32 [M][ln_slots][2] or NULL */
    const void* ln_c;      /* f32 [N] */
    const void* ln_bias;   /* f32 [N] */
    int ln_slots;
    float ln_eps;
    void* stats_out;       /* f32 [M][teo_gemm_stats_slots(M,N,K)][2] or NULL */
} teo_gemm_opts;
int teo_gemm_stats_slots(int M, int N, int K);
int teo_gemm_bf16_ex(teo_handle* h, const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                     const teo_gemm_opts* opts, void* workspace, size_t workspace_bytes, void* stream);
/* stats f32 [rows][slots][2]: slot 0 = (sum x, sum x^2) of the bf16 row x[r, 0..d), the other slots zero */
int teo_row_stats(const void* x, void* stats, int rows, int d, int slots, void* stream);

/* Blocked weight layout.  A weight matrix W[N,K] (N % 128 == 0, K % 64 == 0) may be stored as
 * bf16 [N/128][K/64][128][64]: each 128-row × 64-column operand tile is 16 KiB CONTIGUOUS in HBM, so the TMA
 * producer streams whole DRAM bursts instead of 128 separate 128-byte row pieces 2·K bytes apart (what bounds the
 * weight-streaming decode GEMMs).  teo_weight_to_blocked converts out of place; teo_gemm_bf16_wblocked is
 * teo_gemm_bf16 for such a W (same results). */
int teo_weight_to_blocked(const void* w_rowmajor, void* w_blocked, int N, int K, void* stream);
int teo_gemm_bf16_wblocked(teo_handle* h, const void* A, int lda, const void* W_blocked, void* C, int ldc, int M, int N,
                           int K, const void* bias, const void* residual, int ldr, int act, int out_fp32, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---- vision tower pieces --------------------------------------------------------------- */
/* ToTensor+Normalize (processing_image.py:18,22) fused with im2col for the 14x14/14 patch conv
 * (HF CLIPVisionEmbeddings.patch_embedding): frames u8 [n,H,W,3] NHWC → patches bf16 [n*g*g, kpad],
 * column (c*P+ky)*P+kx, columns >= 3*P*P zero. */
int teo_patchify_u8_nhwc(const void* frames_u8, void* patches, int n_frames, int image, int patch, int kpad,
                         void* stream);
/* same for already-normalised float frames f32 [n,3,H,W] (the reference's pixel_values,
 * eval/inference.py:52-53) */
int teo_patchify_f32_nchw(const void* pixel_values, void* patches, int n_frames, int image, int patch, int kpad,
                          void* stream);
/* The processor's transform chain for ONE image of any size (processing_image.py:15-25: ToTensor → Resize(short
 * side → S, bicubic, antialias) → CenterCrop(S) → Normalize): src u8 [H,W,3] → dst f32 [3,S,S], the reference's
 * pixel_values.  The caller passes the geometry torchvision derives on the host: nh x nw = size after Resize
 * (short side S, long side int(S*long/short)), (top,left) = int(round((n - S)/2)) crop origin.  Rows up to 68 k
 * pixels wide.  Workspace: teo_resize_workspace_bytes(H, W, S). */
size_t teo_resize_workspace_bytes(int H, int W, int S);
int teo_resize_crop_normalize_u8(const void* src_u8, int H, int W, int nh, int nw, int top, int left, int S,
                                 const float* mean3, const float* std3, void* dst_f32, void* workspace,
                                 size_t workspace_bytes, void* stream);
/* [CLS; patches] + position_embedding → pre_layrnorm (modeling_image.py:645-649):
 * patch_out bf16 [n*np, d] → hidden bf16 [n*(np+1), d] */
int teo_vit_assemble_preln(const void* patch_out, const void* cls, const void* pos, const void* ln_w,
                           const void* ln_b, void* hidden, int n_frames, int n_patches, int d, float eps,
                           void* stream);
/* LayerNorm over the last dim (bf16 in/out, fp32 statistics) */
int teo_layernorm(const void* x, const void* w, const void* b, void* y, int rows, int d, float eps, void* stream);
/* drop CLS (languagebind/__init__.py:123-124): hidden [n,(np+1),d] → feats [n,np,d] */
int teo_vit_drop_cls(const void* hidden, void* feats, int n_frames, int n_patches, int d, void* stream);

/* ---- attention ------------------------------------------------------------------------- */
/* Variable-length multi-head attention, softmax(scale·QK^T [+causal mask])·V, flash-style
 * (replaces HF CLIPAttention's bmm/softmax/bmm and HF LlamaAttention's prefill branch).
 * q/k/v: bf16, token-major, row strides ldq/ldk/ldv elements, head h at column h*head_dim;
 * sequence b owns rows [cu_seqlens[b], cu_seqlens[b+1]).  out: bf16 [tokens, ldo].
 * head_dim ∈ {64, 128}.  cu_seqlens: int32 [n_seqs+1] on device. */
int teo_flash_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                        const void* cu_seqlens, int n_seqs, int max_seqlen, int n_heads, int head_dim, float scale,
                        int causal, void* stream);
/* The same contraction on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA-fed K/V ring) — the path
 * teo_vit_encode / teo_llama_prefill take.  Needs the handle (TMA descriptors are cached there) and the total
 * number of token rows behind q/k/v (= cu_seqlens[n_seqs]).  q_offset ∈ {0,1}: with 1, row 0 of every sequence is
 * computed by a spare warp of the CTA from the K/V tiles in shared memory instead of a 128-row tile (ViT: 257
 * tokens = CLS + two tiles); q_offset must be 0 when causal.  Same numerics contract as teo_flash_attention (P rounded to bf16). */
int teo_flash_attention_tc(teo_handle* h, const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out,
                           int ldo, const void* cu_seqlens, int n_seqs, int max_seqlen, int total_tokens, int n_heads,
                           int head_dim, float scale, int causal, int q_offset, void* stream);

/* KV pages: one pool per layer, bf16 [n_pages][2 (K,V)][n_heads][page_size][head_dim]
 * (replaces HF's torch.cat KV growth).  block_table int32 [n_seqs, max_pages]. */
/* Page management (host side; the pools are caller-owned device memory, one per layer, the same page id indexes every
 * layer's pool).  teo_kv_pool_bytes: bytes of ONE layer's pool.  teo_kv_plan: pages each sequence needs to hold its
 * prompt plus max_new_tokens, their maximum (= block-table row length) and sum (= pool size for a fresh batch).
 * teo_kv_create/alloc/free/destroy: a free-list allocator over n_pages page ids — teo_kv_alloc hands a sequence
 * ceil(n_tokens/page_size) ids (lowest free first) and returns their count, teo_kv_free returns the pages of a retired
 * sequence to the pool (continuous batching / sequence retirement). */
typedef struct teo_kv_allocator teo_kv_allocator;
size_t teo_kv_pool_bytes(int n_pages, int n_heads, int page_size, int head_dim, int exact);
int teo_kv_plan(const int* host_seq_lens, int n_seqs, int max_new_tokens, int page_size, int* host_pages_per_seq,
                int* max_pages_out, int* total_pages_out);
int teo_kv_create(int n_pages, teo_kv_allocator** out);
int teo_kv_destroy(teo_kv_allocator* a);
int teo_kv_available(const teo_kv_allocator* a);
int teo_kv_alloc(teo_kv_allocator* a, int n_tokens, int page_size, int* host_pages_out, int max_pages);
int teo_kv_free(teo_kv_allocator* a, const int* host_pages, int n_pages);
/* RoPE (rotate-half, HF LlamaRotaryEmbedding) on q and k in place inside a fused qkv buffer
 * bf16 [tokens, 3*n_heads*head_dim], then scatter k,v rows into the pages.
 * positions int32 [tokens]; seq_ids int32 [tokens] (row of block_table; NULL → token index).
 * rope_cos / rope_sin: f32 [max_pos, head_dim/2] tables built on the host like HF's cos_cached /
 * sin_cached (so angles are bit-identical to the PyTorch formula). */
int teo_rope_kv_write(void* qkv, const void* positions, const void* seq_ids, void* kv_pages, const void* block_table,
                      int max_pages, int tokens, int n_heads, int head_dim, int page_size, const void* rope_cos,
                      const void* rope_sin, void* stream);
/* One-token-per-sequence attention over the paged cache (HF LlamaAttention decode branch).
 * q: bf16 [n_seqs, n_heads*head_dim] (row stride ldq); seq_lens int32 [n_seqs] = cached tokens
 * including the current one; out bf16 [n_seqs, n_heads*head_dim]. */
size_t teo_decode_attention_workspace_bytes(int n_seqs, int n_heads, int head_dim, int max_splits);
int teo_decode_attention(const void* q, int ldq, const void* kv_pages, const void* block_table, int max_pages,
                         const void* seq_lens, void* out, int n_seqs, int n_heads, int head_dim, int page_size,
                         int max_seq_len, float scale, void* workspace, size_t workspace_bytes, void* stream);
/* Same call with the handle: takes the tensor-core kernel (mma.sync over TMA-swizzled page tiles, TMA descriptor of
 * the pool cached in the handle) for head_dim 128 / page_size 64 — the form teo_llama_decode_step uses; other shapes
 * run the CUDA-core kernel.  Rows of a page past the sequence length may hold anything (they are never read into
 * the result). */
int teo_decode_attention_h(teo_handle* h, const void* q, int ldq, const void* kv_pages, const void* block_table, int max_pages,
                           const void* seq_lens, void* out, int n_seqs, int n_heads, int head_dim, int page_size,
                           int max_seq_len, float scale, void* workspace, size_t workspace_bytes, void* stream);

/* ---- LLaMA pieces ---------------------------------------------------------------------- */
/* y = x * rsqrt(mean(x^2)+eps) * w  (HF LlamaRMSNorm; fp32 statistics, one bf16 rounding) */
int teo_rmsnorm(const void* x, const void* w, void* y, int rows, int d, float eps, void* stream);
/* out[r, i] = silu(gate_up[r, i]) * gate_up[r, inter + i]  (HF LlamaMLP act_fn(gate)*up) */
int teo_swiglu(const void* gate_up, void* out, int rows, int inter, void* stream);
/* the multimodal splice (llava_arch.py:251-293) as a gather: for flattened token row t,
 * src[t] >= 0 → embed_tokens[src[t]]; src[t] < 0 → image_feats[-(src[t]+1)] (row of the
 * [n_images*tokens_per_image, d] projector output).  src int32 [tokens]. */
int teo_splice_embed(const void* embed_tokens, const void* image_feats, const void* src, void* out, int tokens, int d,
                     void* stream);
/* greedy step (HF greedy_search: argmax(logits[:, -1]); first index wins ties) + the eval
 * path's stopping rule (mm_utils.py:73-104 with keywords ["</s>"] ≡ last id == eos):
 * logits f32 [n_seqs, vocab]; finished u8 [n_seqs] (in/out); tokens int32 [n_seqs, max_new];
 * step = column to write; next_ids int32 [n_seqs] (eos for finished rows). */
int teo_argmax_step(const void* logits, int vocab, void* finished, void* tokens, int max_new, int step,
                    void* next_ids, int n_seqs, int eos_id, void* stream);

/* sampling step (HF sample(): TemperatureLogitsWarper + TopKLogitsWarper + multinomial — the branch
 * eval/inference.py:64-72 takes with do_sample=True, temperature=0.2 and HF's default top_k=50).
 * Same state arguments as teo_argmax_step; u comes from a counter-based generator keyed by
 * (seed, step, sequence): reproducible and graph-replayable, not torch's RNG stream.  top_k <= 0 → no filter. */
int teo_sample_step(const void* logits, int vocab, float temperature, int top_k, uint64_t seed, void* finished,
                    void* tokens, int max_new, int step, void* next_ids, int n_seqs, int eos_id, void* stream);
/* select how teo_llama_decode_step picks the next token for this handle: temperature <= 0 → greedy */
int teo_set_sampling(teo_handle* h, float temperature, int top_k, uint64_t seed);

/* the decode step reads the sampling seed from device memory (u64 [1]) instead of the value given to teo_set_sampling, so
 * that one captured CUDA graph of the step serves every seed; NULL switches back to the host value */
int teo_set_sampling_seed_device(teo_handle* h, const void* seed_u64_device);

/* teo_llama_decode_step can run the GEMMs between two attention calls (o_proj, gate/up, down, next qkv or lm_head, with their
 * reductions) as ONE persistent kernel per layer with grid barriers between the phases (csrc/decode_chain.cu) instead of one
 * kernel per GEMM + one per reduction.  Both produce bit-identical logits and ids; measured on B200 the chain is the slower
 * of the two (profiles/r02_decode_chain.txt), so it is OFF by default: enabled != 0 (or TEO_DEC_CHAIN=1) selects it. */
int teo_set_decode_chain(teo_handle* h, int enabled);

/* programmatic dependent launch between the kernels of teo_llama_decode_step (default on): each kernel's launch,
 * prologue and — for the GEMMs — weight prefetch overlap the tail of its predecessor.  Results are identical. */
int teo_set_pdl(teo_handle* h, int enabled);

/* ---- Exact (parity) mode ---------------------------------------------------------------- */
/* The north-star parity clause asks for bit-exact greedy ids "with fp32 accumulation".  With bf16 activation storage a
 * 32-layer random-init LLaMA sits 3-4 % from an fp32 run whatever the kernels do, so the whole-model entry points below
 * have a second mode, selected per model struct (`exact != 0`): activations, residual stream and KV pages are f32 in
 * HBM; every nn.Linear still runs on the tensor cores, its f32 input split into three bf16 terms (hi + mid + lo == x)
 * that are contracted against the same bf16 weights with fp32 accumulation (teo_gemm_bf16x3); norms, RoPE, softmax and
 * activations run in fp32 with the HF-4.31 op order.  Buffers that change type: teo_vit_encode feats f32;
 * teo_projector_mlp2x feats/out f32; teo_llama_prefill x f32 (build it with teo_splice_embed_f32); kv_pages f32
 * [n_pages][2][n_heads][page_size][head_dim].  The *_workspace_bytes queries follow the struct's flag. */
int teo_splice_embed_f32(const void* embed_tokens, const void* image_feats_f32, const void* src, void* out_f32, int tokens,
                         int d, void* stream);
/* x f32 [rows,K] -> bf16 [rows,3K] = hi | mid | lo planes (K % 4 == 0) */
int teo_split_f32_bf16x3(const void* x_f32, void* planes_bf16, int rows, int K, void* stream);
/* C f32 [M,N] = X[M,K]·W[N,K]^T + bias[N] (+ residual f32 [M,N], may alias C); X as the planes above (K % 64 == 0);
 * W bf16 row-major [N,K] or blocked (w_blocked != 0).  Workspace: teo_gemm_workspace_bytes(M, N, 3*K). */
int teo_gemm_bf16x3(teo_handle* h, const void* planes_bf16, const void* W, int w_blocked, void* C_f32, int M, int N, int K,
                    const void* bias, const void* residual_f32, void* workspace, size_t workspace_bytes, void* stream);

/* ---- whole-model entry points ---------------------------------------------------------- */
typedef struct {
    const void *ln1_w, *ln1_b;   /* [d] */
    const void *qkv_w, *qkv_b;   /* [3d, d], [3d]  (q;k;v rows stacked) */
    const void *out_w, *out_b;   /* [d, d], [d] */
    const void *ln2_w, *ln2_b;   /* [d] */
    const void *fc1_w, *fc1_b;   /* [inter, d], [inter] */
    const void *fc2_w, *fc2_b;   /* [d, inter], [d] */
    /* optional (all six or none): LayerNorm folded into the two linears that consume it (teo_gemm_bf16_ex):
     * qkv_wf = bf16(ln1_w * qkv_w) [3d,d], qkv_c = rowsum(qkv_wf) f32 [3d], qkv_bf = qkv_w·ln1_b + qkv_b f32 [3d]; fc1_* likewise
     * with ln2.  When present (and the batch is large enough for the tiled GEMM) teo_vit_encode runs no LayerNorm kernels:
     * the out-proj / fc2 GEMMs emit the row statistics of the residual stream they write. */
    const void *qkv_wf, *qkv_c, *qkv_bf;
    const void *fc1_wf, *fc1_c, *fc1_bf;
} teo_vit_layer;

typedef struct {
    int hidden, inter, heads, image, patch, kpad, act, layers_run;
    float eps;
    int w_blocked;               /* != 0: every GEMM weight below is in the blocked layout (teo_weight_to_blocked) */
    int exact;                   /* != 0: exact (parity) mode, see "Exact mode" below: feats is f32 */
    const void* patch_w;         /* [hidden, kpad] bf16, columns (c*P+ky)*P+kx, zero padded */
    const void *cls, *pos;       /* [hidden], [np+1, hidden] */
    const void *pre_ln_w, *pre_ln_b;
    const teo_vit_layer* layers; /* HOST array [layers_run] */
} teo_vit_model;

/* CLIPVisionTransformer.forward up to hidden_states[select_layer] + feature_select
 * (modeling_image.py:610-672, languagebind/__init__.py:121-146).  Exactly one of frames_u8
 * (u8 NHWC) / pixel_values (f32 NCHW, already normalised) is non-NULL.
 * feats: bf16 [n_frames, np, hidden] (CLS dropped). */
size_t teo_vit_workspace_bytes(const teo_vit_model* m, int n_frames);
int teo_vit_encode(teo_handle* h, const teo_vit_model* m, const void* frames_u8, const void* pixel_values,
                   int n_frames, void* feats, void* workspace, size_t workspace_bytes, void* stream);

typedef struct {
    int in_dim, hidden;
    int w_blocked;       /* != 0: w0 / w2 in the blocked layout */
    int exact;           /* != 0: exact mode — feats and out are f32 */
    const void *w0, *b0; /* [hidden, in_dim], [hidden] */
    const void *w2, *b2; /* [hidden, hidden], [hidden] */
} teo_projector;
/* mlp2x_gelu (multimodal_projector/builder.py:41-48): feats [rows,in_dim] → out [rows,hidden] */
size_t teo_projector_workspace_bytes(const teo_projector* p, int rows);
int teo_projector_mlp2x(teo_handle* h, const teo_projector* p, const void* feats, int rows, void* out, void* workspace,
                        size_t workspace_bytes, void* stream);

typedef struct {
    const void* in_norm;    /* [h] */
    const void* qkv_w;      /* [3h, h] (q;k;v) */
    const void* o_w;        /* [h, h] */
    const void* post_norm;  /* [h] */
    const void* gate_up_w;  /* [2*inter, h] (gate rows then up rows, or interleaved: teo_llama_model.gate_up_interleaved) */
    const void* down_w;     /* [h, inter] */
    void* kv_pages;         /* this layer's page pool */
} teo_llama_layer;

typedef struct {
    int hidden, inter, heads, layers, vocab, page_size, rope_max_pos;
    float eps;
    int w_blocked;                   /* != 0: qkv_w / o_w / gate_up_w / down_w / lm_head in the blocked layout (embed stays row-major) */
    int gate_up_interleaved;         /* != 0: gate_up_w rows interleaved in blocks of 32 (TEO_ACT_SWIGLU_PAIRS), inter % 32 == 0;
                                        0: gate rows then up rows */
    int exact;                       /* != 0: exact mode — x (prefill) is f32 [tokens,h] and every layer's kv_pages pool is f32 */
    const void *rope_cos, *rope_sin; /* f32 [rope_max_pos, head_dim/2] */
    const void* embed;      /* [vocab, h] */
    const void* final_norm; /* [h] */
    const void* lm_head;    /* [vocab, h] */
    const teo_llama_layer* layer; /* HOST array [layers] */
} teo_llama_model;

/* Ragged batched prefill (HF LlamaModel.forward over inputs_embeds, llava_llama.py:88-99) that
 * writes K/V into the pages and returns only last-position logits f32 [n_seqs, vocab].
 * x: bf16 [tokens, h] spliced embeddings (overwritten: used as the residual stream);
 * cu_seqlens int32 [n_seqs+1]; positions/seq_ids int32 [tokens]; last_rows int32 [n_seqs]. */
size_t teo_llama_prefill_workspace_bytes(const teo_llama_model* m, int tokens, int n_seqs);
int teo_llama_prefill(teo_handle* h, const teo_llama_model* m, void* x, int tokens, const void* cu_seqlens,
                      const void* positions, const void* seq_ids, const void* last_rows, int n_seqs, int max_seqlen,
                      const void* block_table, int max_pages, void* logits, void* workspace, size_t workspace_bytes,
                      void* stream);

/* One greedy decode step for n_seqs sequences (HF generate loop body, SURVEY.md §3.2 hot loop
 * #1): embed next_ids → 32 layers over the paged cache → logits → argmax/eos.  All state lives
 * on the device so the step can be captured in a CUDA graph:
 *   next_ids int32 [n_seqs] (in: token to feed; out: token produced)
 *   seq_lens int32 [n_seqs] (in: cached length before this step; out: +1)
 *   finished u8 [n_seqs], tokens int32 [n_seqs, max_new], step_ptr int32 [1] (column, +1). */
size_t teo_llama_decode_workspace_bytes(const teo_llama_model* m, int n_seqs, int max_seq_len);
int teo_llama_decode_step(teo_handle* h, const teo_llama_model* m, void* next_ids, void* seq_lens, void* finished,
                          void* tokens, int max_new, void* step_ptr, int n_seqs, int max_seq_len,
                          const void* block_table, int max_pages, void* logits, int eos_id, void* workspace,
                          size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TEOCHAT_B200_H_ */
