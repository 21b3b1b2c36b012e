// Launch sequences of the exact ("parity") mode — see exact.cu.  Same C-ABI entry points as the bf16 path (model.cu
// dispatches here when teo_*_model.exact != 0); what changes is the storage type of every activation buffer:
//   teo_vit_encode       feats   f32 [n, np, d]
//   teo_projector_mlp2x  feats / out f32
//   teo_llama_prefill    x       f32 [tokens, h]   (teo_splice_embed_f32), KV pages f32
//   teo_llama_decode_step        KV pages f32
// Weights, biases and norm gains stay bf16 (they are bf16 values in the checkpoint / oracle as well).
#include <math.h>

#include <algorithm>

#include "arena.h"
#include "common.h"

namespace teo {

int x_split(const float* in, bf16* out3, float* out_f32, int rows, int K, int mode, int n_patches, int interleaved, cudaStream_t s);
int x_norm(int mode, const float* in, const bf16* w, const bf16* b, float* out_f32, bf16* out3, int rows, int d, float eps,
           const bf16* cls, const bf16* pos, int n_patches, cudaStream_t s);
int x_splice_embed(const bf16* embed, const float* feats, const int* src, float* out, int tokens, int d, cudaStream_t s);
int x_gather_rows(const float* in, const int* rows, float* out, int n, int d, cudaStream_t s);
int x_rope_kv_write(float* qkv, const int* positions, const int* seq_ids, float* kv_pages, const int* block_table, int max_pages,
                    int tokens, int n_heads, int head_dim, int page_size, const float* rope_cos, const float* rope_sin, cudaStream_t s);
int x_attention(bool paged, const float* qkv, int ld, float* out, int ldo, const int* cu_seqlens, const int* positions, const int* seq_ids,
                const float* kv_pages, const int* block_table, int max_pages, int rows, int n_heads, int head_dim, int page_size,
                int max_keys, float scale, cudaStream_t s);
int x_fill_seq_ids(int* seq_ids, int rows, int len, cudaStream_t s);
int x_patchify(bool u8, const void* in, void* patches3, int n_frames, int image, int patch, int kpad, cudaStream_t stream);

// C f32 [M,N] = A3 (three bf16 planes of an fp32 [M,K]) · W[N,K]^T + bias (+ fp32 residual, may alias C)
static int gemm_x(teo_handle* h, const bf16* a3, const void* W, float* C, int M, int N, int K, const void* bias, const float* residual,
                  void* ws, size_t ws_bytes, cudaStream_t stream, int w_blocked) {
    GemmEpilogue ep;
    ep.bias = static_cast<const bf16*>(bias);
    ep.out_fp32 = 1;
    ep.k_planes = 3;
    if (residual) {
        ep.residual = reinterpret_cast<const bf16*>(residual);
        ep.residual_f32 = 1;
        ep.ldr = N;
    }
    return launch_gemm(h, a3, 3 * K, static_cast<const bf16*>(W), K, C, N, M, N, K, ep, ws, ws_bytes, stream, w_blocked);
}
static size_t gemm_x_ws(int M, int N, int K) { return teo_gemm_workspace_bytes(M, N, 3 * K); }

// ------------------------------------------------------------------------------ ViT
struct VitXWs {
    bf16 *patches3, *a3, *m3;
    float *patch_out, *hidden, *qkv, *attn, *mlp;
    int* seq_ids;
    int* cu;
    uint8_t* gws;
    size_t gws_bytes;
};
static size_t vit_x_layout(const teo_vit_model* m, int n, Arena& A, VitXWs* w) {
    const int g = m->image / m->patch, np = g * g, d = m->hidden;
    const size_t rows = static_cast<size_t>(n) * (np + 1);
    VitXWs t;
    t.patches3 = A.take<bf16>(static_cast<size_t>(n) * np * 3 * m->kpad);
    t.patch_out = A.take<float>(static_cast<size_t>(n) * np * d);
    t.hidden = A.take<float>(rows * d);
    t.a3 = A.take<bf16>(rows * 3 * d);
    t.qkv = A.take<float>(rows * 3 * d);
    t.attn = A.take<float>(rows * d);
    t.mlp = A.take<float>(rows * m->inter);
    t.m3 = A.take<bf16>(rows * 3 * m->inter);
    t.seq_ids = A.take<int>(rows);
    t.cu = A.take<int>(n + 1);
    size_t gb = gemm_x_ws(n * np, d, m->kpad);
    gb = std::max(gb, gemm_x_ws(static_cast<int>(rows), 3 * d, d));
    gb = std::max(gb, gemm_x_ws(static_cast<int>(rows), m->inter, d));
    gb = std::max(gb, gemm_x_ws(static_cast<int>(rows), d, m->inter));
    t.gws_bytes = gb;
    t.gws = A.take<uint8_t>(gb);
    if (w) *w = t;
    return A.off;
}
size_t vit_exact_workspace_bytes(const teo_vit_model* m, int n) {
    Arena A(nullptr, 0);
    return vit_x_layout(m, n, A, nullptr);
}

__global__ void fill_cu_kernel(int* cu, int n, int len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) cu[i] = i * len;
}

int vit_encode_exact(teo_handle* h, const teo_vit_model* m, const void* frames_u8, const void* pixel_values, int n, void* feats_,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    const int g = m->image / m->patch, np = g * g, d = m->hidden, hd = d / m->heads;
    const int rows = n * (np + 1);
    Arena A(workspace, workspace_bytes);
    VitXWs w;
    vit_x_layout(m, n, A, &w);
    if (!A.ok) {
        set_error("vit_encode (exact): workspace too small (%zu needed, %zu given)", A.off, workspace_bytes);
        return TEO_ERR_WORKSPACE;
    }
    float* feats = static_cast<float*>(feats_);
    TEO_TRY(x_patchify(frames_u8 != nullptr, frames_u8 ? frames_u8 : pixel_values, w.patches3, n, m->image, m->patch, m->kpad, stream));
    TEO_TRY(gemm_x(h, w.patches3, m->patch_w, w.patch_out, n * np, d, m->kpad, nullptr, nullptr, w.gws, w.gws_bytes, stream, m->w_blocked));
    TEO_TRY(x_norm(1, w.patch_out, static_cast<const bf16*>(m->pre_ln_w), static_cast<const bf16*>(m->pre_ln_b), w.hidden, nullptr, rows, d,
                   m->eps, static_cast<const bf16*>(m->cls), static_cast<const bf16*>(m->pos), np, stream));
    TEO_TRY(x_fill_seq_ids(w.seq_ids, rows, np + 1, stream));
    fill_cu_kernel<<<(n + 256) / 256, 256, 0, stream>>>(w.cu, n, np + 1);
    TEO_LAUNCH_CHECK("fill_cu_kernel");
    h->launches += 4;
    const float scale = 1.0f / sqrtf(static_cast<float>(hd));
    const int xact = m->act == TEO_ACT_QUICK_GELU ? 1 : (m->act == TEO_ACT_GELU ? 2 : 0);
    for (int l = 0; l < m->layers_run; ++l) {
        const teo_vit_layer& L = m->layers[l];
        TEO_TRY(x_norm(0, w.hidden, static_cast<const bf16*>(L.ln1_w), static_cast<const bf16*>(L.ln1_b), nullptr, w.a3, rows, d, m->eps,
                       nullptr, nullptr, 0, stream));
        TEO_TRY(gemm_x(h, w.a3, L.qkv_w, w.qkv, rows, 3 * d, d, L.qkv_b, nullptr, w.gws, w.gws_bytes, stream, m->w_blocked));
        TEO_TRY(x_attention(false, w.qkv, 3 * d, w.attn, d, w.cu, nullptr, w.seq_ids, nullptr, nullptr, 0, rows, m->heads, hd, 1, np + 1, scale,
                            stream));
        TEO_TRY(x_split(w.attn, w.a3, nullptr, rows, d, 0, 0, 0, stream));
        TEO_TRY(gemm_x(h, w.a3, L.out_w, w.hidden, rows, d, d, L.out_b, w.hidden, w.gws, w.gws_bytes, stream, m->w_blocked));
        TEO_TRY(x_norm(0, w.hidden, static_cast<const bf16*>(L.ln2_w), static_cast<const bf16*>(L.ln2_b), nullptr, w.a3, rows, d, m->eps,
                       nullptr, nullptr, 0, stream));
        TEO_TRY(gemm_x(h, w.a3, L.fc1_w, w.mlp, rows, m->inter, d, L.fc1_b, nullptr, w.gws, w.gws_bytes, stream, m->w_blocked));
        TEO_TRY(x_split(w.mlp, w.m3, nullptr, rows, m->inter, xact, 0, 0, stream));
        TEO_TRY(gemm_x(h, w.m3, L.fc2_w, w.hidden, rows, d, m->inter, L.fc2_b, w.hidden, w.gws, w.gws_bytes, stream, m->w_blocked));
        h->launches += 5;
    }
    TEO_TRY(x_split(w.hidden, nullptr, feats, n * np, d, 0, np, 0, stream));      // drop CLS (languagebind/__init__.py:123-124)
    h->launches += 1;
    return TEO_OK;
}

// ------------------------------------------------------------------------------ projector
size_t projector_exact_workspace_bytes(const teo_projector* p, int rows) {
    const size_t r = rows;
    return al256(r * 3 * p->in_dim * sizeof(bf16)) + al256(r * p->hidden * sizeof(float)) + al256(r * 3 * p->hidden * sizeof(bf16)) +
           al256(std::max(gemm_x_ws(rows, p->hidden, p->in_dim), gemm_x_ws(rows, p->hidden, p->hidden)));
}
int projector_exact(teo_handle* h, const teo_projector* p, const void* feats, int rows, void* out, void* workspace, size_t workspace_bytes,
                    cudaStream_t stream) {
    Arena A(workspace, workspace_bytes);
    const size_t r = rows;
    bf16* f3 = A.take<bf16>(r * 3 * p->in_dim);
    float* mid = A.take<float>(r * p->hidden);
    bf16* m3 = A.take<bf16>(r * 3 * p->hidden);
    const size_t gb = std::max(gemm_x_ws(rows, p->hidden, p->in_dim), gemm_x_ws(rows, p->hidden, p->hidden));
    uint8_t* gw = A.take<uint8_t>(gb);
    if (!A.ok) {
        set_error("projector (exact): workspace too small (%zu needed, %zu given)", A.off, workspace_bytes);
        return TEO_ERR_WORKSPACE;
    }
    TEO_TRY(x_split(static_cast<const float*>(feats), f3, nullptr, rows, p->in_dim, 0, 0, 0, stream));
    TEO_TRY(gemm_x(h, f3, p->w0, mid, rows, p->hidden, p->in_dim, p->b0, nullptr, gw, gb, stream, p->w_blocked));
    TEO_TRY(x_split(mid, m3, nullptr, rows, p->hidden, 2, 0, 0, stream));           // nn.GELU() (erf)
    TEO_TRY(gemm_x(h, m3, p->w2, static_cast<float*>(out), rows, p->hidden, p->hidden, p->b2, nullptr, gw, gb, stream, p->w_blocked));
    h->launches += 2;
    return TEO_OK;
}

// ------------------------------------------------------------------------------ LLaMA
struct LlamaXWs {
    bf16 *a3, *act3;
    float *x, *qkv, *attn, *gate_up, *last_x;
    uint8_t* gws;
    size_t gws_bytes;
};
// rows = tokens of the pass (prefill: all tokens; decode: n_seqs); own_x: the decode step keeps its residual stream here
static size_t llama_x_layout(const teo_llama_model* m, int rows, int n_seqs, bool own_x, Arena& A, LlamaXWs* w) {
    const size_t T = rows, h = m->hidden, I = m->inter;
    LlamaXWs t;
    t.x = own_x ? A.take<float>(T * h) : nullptr;
    t.a3 = A.take<bf16>(T * 3 * h);
    t.qkv = A.take<float>(T * 3 * h);
    t.attn = A.take<float>(T * h);
    t.gate_up = A.take<float>(T * 2 * I);
    t.act3 = A.take<bf16>(T * 3 * I);
    t.last_x = A.take<float>(static_cast<size_t>(n_seqs) * h);
    size_t gb = gemm_x_ws(n_seqs, m->vocab, m->hidden);
    gb = std::max(gb, gemm_x_ws(rows, 3 * m->hidden, m->hidden));
    gb = std::max(gb, gemm_x_ws(rows, 2 * m->inter, m->hidden));
    gb = std::max(gb, gemm_x_ws(rows, m->hidden, m->inter));
    t.gws_bytes = gb;
    t.gws = A.take<uint8_t>(gb);
    if (w) *w = t;
    return A.off;
}
size_t llama_prefill_exact_workspace_bytes(const teo_llama_model* m, int tokens, int n_seqs) {
    Arena A(nullptr, 0);
    return llama_x_layout(m, tokens, n_seqs, false, A, nullptr);
}
size_t llama_decode_exact_workspace_bytes(const teo_llama_model* m, int n_seqs) {
    Arena A(nullptr, 0);
    return llama_x_layout(m, n_seqs, n_seqs, true, A, nullptr);
}

// One pass of all layers over `rows` token rows (x f32 in place): the HF-4.31 LlamaDecoderLayer op order in fp32.
static int llama_layers_exact(teo_handle* h, const teo_llama_model* m, float* x, int rows, const int* positions, const int* seq_ids,
                              const int* block_table, int max_pages, int max_keys, const LlamaXWs& w, cudaStream_t stream) {
    const int hdim = m->hidden, hd = hdim / m->heads, I = m->inter;
    const float scale = 1.0f / sqrtf(static_cast<float>(hd));
    for (int l = 0; l < m->layers; ++l) {
        const teo_llama_layer& L = m->layer[l];
        float* pages = static_cast<float*>(L.kv_pages);
        TEO_TRY(x_norm(2, x, static_cast<const bf16*>(L.in_norm), nullptr, nullptr, w.a3, rows, hdim, m->eps, nullptr, nullptr, 0, stream));
        TEO_TRY(gemm_x(h, w.a3, L.qkv_w, w.qkv, rows, 3 * hdim, hdim, nullptr, nullptr, w.gws, w.gws_bytes, stream, m->w_blocked));
        TEO_TRY(x_rope_kv_write(w.qkv, positions, seq_ids, pages, block_table, max_pages, rows, m->heads, hd, m->page_size,
                                static_cast<const float*>(m->rope_cos), static_cast<const float*>(m->rope_sin), stream));
        TEO_TRY(x_attention(true, w.qkv, 3 * hdim, w.attn, hdim, nullptr, positions, seq_ids, pages, block_table, max_pages, rows, m->heads, hd,
                            m->page_size, max_keys, scale, stream));
        TEO_TRY(x_split(w.attn, w.a3, nullptr, rows, hdim, 0, 0, 0, stream));
        TEO_TRY(gemm_x(h, w.a3, L.o_w, x, rows, hdim, hdim, nullptr, x, w.gws, w.gws_bytes, stream, m->w_blocked));
        TEO_TRY(x_norm(2, x, static_cast<const bf16*>(L.post_norm), nullptr, nullptr, w.a3, rows, hdim, m->eps, nullptr, nullptr, 0, stream));
        TEO_TRY(gemm_x(h, w.a3, L.gate_up_w, w.gate_up, rows, 2 * I, hdim, nullptr, nullptr, w.gws, w.gws_bytes, stream, m->w_blocked));
        TEO_TRY(x_split(w.gate_up, w.act3, nullptr, rows, I, 3, 0, m->gate_up_interleaved, stream));
        TEO_TRY(gemm_x(h, w.act3, L.down_w, x, rows, hdim, I, nullptr, x, w.gws, w.gws_bytes, stream, m->w_blocked));
        h->launches += 6;
    }
    return TEO_OK;
}

int llama_prefill_exact(teo_handle* h, const teo_llama_model* m, void* x_, int tokens, const void* positions, const void* seq_ids,
                        const void* last_rows, int n_seqs, int max_seqlen, const void* block_table, int max_pages, void* logits,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    Arena A(workspace, workspace_bytes);
    LlamaXWs w;
    llama_x_layout(m, tokens, n_seqs, false, A, &w);
    if (!A.ok) {
        set_error("llama_prefill (exact): workspace too small (%zu needed, %zu given)", A.off, workspace_bytes);
        return TEO_ERR_WORKSPACE;
    }
    float* x = static_cast<float*>(x_);
    TEO_TRY(llama_layers_exact(h, m, x, tokens, static_cast<const int*>(positions), static_cast<const int*>(seq_ids),
                               static_cast<const int*>(block_table), max_pages, max_seqlen, w, stream));
    TEO_TRY(x_gather_rows(x, static_cast<const int*>(last_rows), w.last_x, n_seqs, m->hidden, stream));
    TEO_TRY(x_norm(2, w.last_x, static_cast<const bf16*>(m->final_norm), nullptr, nullptr, w.a3, n_seqs, m->hidden, m->eps, nullptr, nullptr, 0,
                   stream));
    TEO_TRY(gemm_x(h, w.a3, m->lm_head, static_cast<float*>(logits), n_seqs, m->vocab, m->hidden, nullptr, nullptr, w.gws, w.gws_bytes, stream,
                   m->w_blocked));
    h->launches += 2;
    return TEO_OK;
}

int llama_decode_step_exact(teo_handle* h, const teo_llama_model* m, void* next_ids, void* seq_lens, void* finished, void* tokens, int max_new,
                            void* step_ptr, int n_seqs, int max_seq_len, const void* block_table, int max_pages, void* logits, int eos_id,
                            void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    Arena A(workspace, workspace_bytes);
    LlamaXWs w;
    llama_x_layout(m, n_seqs, n_seqs, true, A, &w);
    if (!A.ok) {
        set_error("llama_decode_step (exact): workspace too small (%zu needed, %zu given)", A.off, workspace_bytes);
        return TEO_ERR_WORKSPACE;
    }
    TEO_TRY(x_splice_embed(static_cast<const bf16*>(m->embed), nullptr, static_cast<const int*>(next_ids), w.x, n_seqs, m->hidden, stream));
    // position of the new token = tokens cached so far; it attends keys 0 … position
    TEO_TRY(llama_layers_exact(h, m, w.x, n_seqs, static_cast<const int*>(seq_lens), nullptr, static_cast<const int*>(block_table), max_pages,
                               max_seq_len, w, stream));
    TEO_TRY(x_norm(2, w.x, static_cast<const bf16*>(m->final_norm), nullptr, nullptr, w.a3, n_seqs, m->hidden, m->eps, nullptr, nullptr, 0, stream));
    TEO_TRY(gemm_x(h, w.a3, m->lm_head, static_cast<float*>(logits), n_seqs, m->vocab, m->hidden, nullptr, nullptr, w.gws, w.gws_bytes, stream,
                   m->w_blocked));
    if (h->temperature > 0.f)
        TEO_TRY(launch_sample_step(static_cast<const float*>(logits), m->vocab, h->temperature, h->top_k, h->sample_seed,
                                   static_cast<uint8_t*>(finished), static_cast<int*>(tokens), max_new, 0, static_cast<int*>(step_ptr),
                                   static_cast<int*>(next_ids), static_cast<int*>(seq_lens), n_seqs, eos_id, stream, h->sample_seed_ptr));
    else
        TEO_TRY(launch_argmax_step(static_cast<const float*>(logits), m->vocab, static_cast<uint8_t*>(finished), static_cast<int*>(tokens),
                                   max_new, 0, static_cast<int*>(step_ptr), static_cast<int*>(next_ids), static_cast<int*>(seq_lens), n_seqs,
                                   eos_id, stream));
    h->launches += 4;
    return TEO_OK;
}

}  // namespace teo

using namespace teo;

/* the multimodal splice in fp32 (exact mode): embed bf16 rows widened, image_feats f32 rows copied */
extern "C" int teo_splice_embed_f32(const void* embed_tokens, const void* image_feats_f32, const void* src, void* out_f32, int tokens, int d,
                                    void* stream) {
    TEO_CHECK_ARG(embed_tokens && src && out_f32, "splice_embed_f32: null pointer");
    TEO_CHECK_ARG(tokens > 0 && d > 0, "splice_embed_f32: tokens=%d d=%d", tokens, d);
    return x_splice_embed(static_cast<const bf16*>(embed_tokens), static_cast<const float*>(image_feats_f32), static_cast<const int*>(src),
                          static_cast<float*>(out_f32), tokens, d, static_cast<cudaStream_t>(stream));
}

/* fp32 activation [rows, K] -> three bf16 planes [rows, 3K] (hi | mid | lo; hi + mid + lo == x) */
extern "C" int teo_split_f32_bf16x3(const void* x_f32, void* planes_bf16, int rows, int K, void* stream) {
    TEO_CHECK_ARG(x_f32 && planes_bf16, "split_f32_bf16x3: null pointer");
    return x_split(static_cast<const float*>(x_f32), static_cast<bf16*>(planes_bf16), nullptr, rows, K, 0, 0, 0, static_cast<cudaStream_t>(stream));
}

/* C f32 [M,N] = X[M,K] · W[N,K]^T + bias (+ residual f32), X given as the three bf16 planes of teo_split_f32_bf16x3:
 * an fp32-input nn.Linear on the bf16 tensor cores (fp32 accumulation in TMEM). */
extern "C" int teo_gemm_bf16x3(teo_handle* h, const void* planes_bf16, const void* W, int w_blocked, void* C_f32, int M, int N, int K,
                               const void* bias, const void* residual_f32, void* workspace, size_t workspace_bytes, void* stream) {
    TEO_CHECK_ARG(h && planes_bf16 && W && C_f32, "gemm_bf16x3: null pointer");
    return gemm_x(h, static_cast<const bf16*>(planes_bf16), W, static_cast<float*>(C_f32), M, N, K, bias, static_cast<const float*>(residual_f32),
                  workspace, workspace_bytes, static_cast<cudaStream_t>(stream), w_blocked);
}
