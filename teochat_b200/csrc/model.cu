// C-ABI lifecycle + whole-model entry points (ViT encode, projector, LLaMA prefill, decode step):
// fixed launch sequences over the kernels in gemm.cu / attention.cu / kernels_misc.cu.  No
// allocation, no host synchronisation — every call only enqueues on the caller's stream, so the
// decode step can be captured in a CUDA graph.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <new>

#include "arena.h"
#include "common.h"

namespace teo {

static thread_local char g_err[512] = "";
static thread_local bool g_pdl = false;
bool pdl_enabled() { return g_pdl; }
void set_pdl(bool on) { g_pdl = on; }
int pdl_mask() {
    static int mask = [] {
        const char* e = getenv("TEO_PDL_MASK");
        return e ? atoi(e) : 7;
    }();
    return mask;
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int launch_flash_attention(const bf16* q, int ldq, const bf16* k, int ldk, const bf16* v, int ldv, bf16* out, int ldo,
                           const int* cu_seqlens, int n_seqs, int max_seqlen, int n_heads, int head_dim, float scale,
                           int causal, cudaStream_t stream);
int launch_flash_attention_tc(teo_handle* h, const bf16* q, int ldq, const bf16* k, int ldk, const bf16* v, int ldv, bf16* out, int ldo,
                              const int* cu_seqlens, int n_seqs, int max_seqlen, int total_tokens, int n_heads, int head_dim,
                              float scale, int causal, int q_offset, cudaStream_t stream);
// TEO_FLASH=mma selects the mma.sync kernels of attention.cu (A/B measurements); default is the tcgen05 path.
static bool flash_use_tc() {
    static const bool tc = [] {
        const char* e = getenv("TEO_FLASH");
        return !(e && e[0] == 'm');
    }();
    return tc;
}
int launch_decode_attention(teo_handle* h, const bf16* q, int ldq, const bf16* kv_pages, const int* block_table, int max_pages,
                            const int* seq_lens, int len_bias, bf16* out, int n_seqs, int n_heads, int head_dim, int page_size,
                            int max_seq_len, float scale, void* workspace, size_t workspace_bytes, cudaStream_t stream);
int launch_rope_kv_write(void* qkv, const int* positions, const int* seq_ids, void* kv_pages, const int* block_table, int max_pages,
                         int tokens, int n_heads, int head_dim, int page_size, const float* rope_cos, const float* rope_sin,
                         cudaStream_t stream);
int launch_reduce_residual_rmsnorm(const PartialInfo& pi, bf16* x, const bf16* w, bf16* y, int rows, int d, float eps, cudaStream_t stream);
int launch_reduce_swiglu(const PartialInfo& pi, bf16* act, int rows, int inter, int interleaved, cudaStream_t stream);
int launch_swiglu(const void* gate_up, void* out, int rows, int inter, int interleaved, cudaStream_t stream);
int launch_reduce_rope_kv_write(const PartialInfo& pi, void* qkv, const int* positions, void* kv_pages,
                                const int* block_table, int max_pages, int n_seqs, int n_heads, int head_dim, int page_size,
                                const float* rope_cos, const float* rope_sin, cudaStream_t stream);

__global__ void fill_cu_seqlens_kernel(int* cu, int n, int len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) cu[i] = i * len;
}

}  // namespace teo

namespace teo {
// exact (parity) mode launch sequences, model_exact.cu
size_t vit_exact_workspace_bytes(const teo_vit_model* m, int n);
int vit_encode_exact(teo_handle* h, const teo_vit_model* m, const void* frames_u8, const void* pixel_values, int n, void* feats,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t projector_exact_workspace_bytes(const teo_projector* p, int rows);
int projector_exact(teo_handle* h, const teo_projector* p, const void* feats, int rows, void* out, void* workspace, size_t workspace_bytes,
                    cudaStream_t stream);
size_t llama_prefill_exact_workspace_bytes(const teo_llama_model* m, int tokens, int n_seqs);
size_t llama_decode_exact_workspace_bytes(const teo_llama_model* m, int n_seqs);
int llama_prefill_exact(teo_handle* h, const teo_llama_model* m, void* x, int tokens, const void* positions, const void* seq_ids,
                        const void* last_rows, int n_seqs, int max_seqlen, const void* block_table, int max_pages, void* logits,
                        void* workspace, size_t workspace_bytes, cudaStream_t stream);
int llama_decode_step_exact(teo_handle* h, const teo_llama_model* m, void* next_ids, void* seq_lens, void* finished, void* tokens, int max_new,
                            void* step_ptr, int n_seqs, int max_seq_len, const void* block_table, int max_pages, void* logits, int eos_id,
                            void* workspace, size_t workspace_bytes, cudaStream_t stream);
}  // namespace teo

using namespace teo;

// ------------------------------------------------------------------------------ lifecycle
extern "C" const char* teo_last_error(void) { return g_err; }
extern "C" int teo_abi_version(void) { return 4; }
#ifndef TEO_BUILD_DIGEST
#define TEO_BUILD_DIGEST "unknown"
#endif
extern "C" const char* teo_build_digest(void) { return TEO_BUILD_DIGEST; }

extern "C" int teo_create(int device_id, teo_handle** out) {
    TEO_CHECK_ARG(out != nullptr, "teo_create: null out");
    *out = nullptr;
    int count = 0;
    TEO_CUDA(cudaGetDeviceCount(&count));
    TEO_CHECK_ARG(device_id >= 0 && device_id < count, "teo_create: device %d not in [0,%d)", device_id, count);
    TEO_CUDA(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    TEO_CUDA(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major != 10) {
        set_error("teo_create: device %d is sm_%d%d; this library is built for sm_100a only", device_id, prop.major, prop.minor);
        return TEO_ERR_UNSUPPORTED;
    }
    teo_handle* h = new (std::nothrow) teo_handle();
    TEO_CHECK_ARG(h != nullptr, "teo_create: out of host memory");
    h->device = device_id;
    h->num_sms = prop.multiProcessorCount;
    h->decode_chain = decode_chain_enabled();
    *out = h;
    return TEO_OK;
}
extern "C" int teo_destroy(teo_handle* h) {
    if (h && h->chain_sync) cudaFree(h->chain_sync);
    if (h && h->sk_flags) cudaFree(h->sk_flags);
    delete h;
    return TEO_OK;
}
extern "C" int teo_set_sampling(teo_handle* h, float temperature, int top_k, uint64_t seed) {
    TEO_CHECK_ARG(h != nullptr, "teo_set_sampling: null handle");
    h->temperature = temperature;
    h->top_k = top_k;
    h->sample_seed = seed;
    return TEO_OK;
}
extern "C" int teo_set_sampling_seed_device(teo_handle* h, const void* seed_u64_device) {
    TEO_CHECK_ARG(h != nullptr, "teo_set_sampling_seed_device: null handle");
    h->sample_seed_ptr = static_cast<const unsigned long long*>(seed_u64_device);
    return TEO_OK;
}
extern "C" int teo_set_decode_chain(teo_handle* h, int enabled) {
    TEO_CHECK_ARG(h != nullptr, "teo_set_decode_chain: null handle");
    h->decode_chain = enabled != 0;
    return TEO_OK;
}
extern "C" int teo_set_pdl(teo_handle* h, int enabled) {
    TEO_CHECK_ARG(h != nullptr, "teo_set_pdl: null handle");
    h->pdl = enabled != 0;
    return TEO_OK;
}
extern "C" unsigned long long teo_launch_count(const teo_handle* h) { return h ? h->launches : 0ULL; }

// ------------------------------------------------------------------------------ ViT
static size_t vit_gemm_ws_bytes(const teo_vit_model* m, int n) {
    const int g = m->image / m->patch, np = g * g, rows = n * (np + 1), d = m->hidden;
    size_t b = teo_gemm_workspace_bytes(n * np, d, m->kpad);
    b = std::max(b, teo_gemm_workspace_bytes(rows, 3 * d, d));
    b = std::max(b, teo_gemm_workspace_bytes(rows, d, d));
    b = std::max(b, teo_gemm_workspace_bytes(rows, m->inter, d));
    b = std::max(b, teo_gemm_workspace_bytes(rows, d, m->inter));
    return b;
}

extern "C" int teo_gemm_stats_slots(int M, int N, int K);
extern "C" int teo_row_stats(const void* x, void* stats, int rows, int d, int slots, void* stream);

static size_t vit_ws_layout(const teo_vit_model* m, int n, Arena* a, bf16** patches, bf16** patch_out, bf16** hidden, bf16** ln_out,
                            bf16** qkv, bf16** attn, bf16** mlp, int** cu, uint8_t** gemm_ws = nullptr, float** stats = nullptr) {
    const int g = m->image / m->patch, np = g * g;
    const size_t rows = static_cast<size_t>(n) * (np + 1);
    Arena local(nullptr, 0);
    Arena& A = a ? *a : local;
    bf16* p;
    p = A.take<bf16>(static_cast<size_t>(n) * np * m->kpad); if (patches) *patches = p;
    p = A.take<bf16>(static_cast<size_t>(n) * np * m->hidden); if (patch_out) *patch_out = p;
    p = A.take<bf16>(rows * m->hidden); if (hidden) *hidden = p;
    p = A.take<bf16>(rows * m->hidden); if (ln_out) *ln_out = p;
    p = A.take<bf16>(rows * 3 * m->hidden); if (qkv) *qkv = p;
    p = A.take<bf16>(rows * m->hidden); if (attn) *attn = p;
    p = A.take<bf16>(rows * m->inter); if (mlp) *mlp = p;
    int* c = A.take<int>(n + 1); if (cu) *cu = c;
    uint8_t* gw = A.take<uint8_t>(vit_gemm_ws_bytes(m, n)); if (gemm_ws) *gemm_ws = gw;   // small batches run the small-M GEMM schedule
    float* st = A.take<float>(rows * 2 * std::max(1, teo_gemm_stats_slots(static_cast<int>(rows), m->hidden, m->hidden)));   // folded-LayerNorm row statistics
    if (stats) *stats = st;
    return A.off;
}

extern "C" size_t teo_vit_workspace_bytes(const teo_vit_model* m, int n_frames) {
    if (!m || n_frames <= 0) return 0;
    if (m->exact) return vit_exact_workspace_bytes(m, n_frames);
    return vit_ws_layout(m, n_frames, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
}

extern "C" int teo_vit_encode(teo_handle* h, const teo_vit_model* m, const void* frames_u8, const void* pixel_values, int n_frames,
                              void* feats, void* workspace, size_t workspace_bytes, void* stream_) {
    TEO_CHECK_ARG(h && m && feats, "vit_encode: null pointer");
    TEO_CHECK_ARG((frames_u8 != nullptr) != (pixel_values != nullptr), "vit_encode: pass exactly one of frames_u8 / pixel_values");
    TEO_CHECK_ARG(n_frames > 0, "vit_encode: n_frames=%d", n_frames);
    TEO_CHECK_ARG(m->hidden % m->heads == 0 && m->layers_run >= 0 && m->layers != nullptr, "vit_encode: bad model");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (m->exact) return vit_encode_exact(h, m, frames_u8, pixel_values, n_frames, feats, workspace, workspace_bytes, stream);
    const int g = m->image / m->patch, np = g * g, d = m->hidden, hd = d / m->heads;
    const int rows = n_frames * (np + 1);
    Arena A(workspace, workspace_bytes);
    bf16 *patches, *patch_out, *hidden, *ln_out, *qkv, *attn, *mlp;
    int* cu;
    uint8_t* gws;
    float* stats;
    vit_ws_layout(m, n_frames, &A, &patches, &patch_out, &hidden, &ln_out, &qkv, &attn, &mlp, &cu, &gws, &stats);
    const size_t gws_bytes = vit_gemm_ws_bytes(m, n_frames);
    if (!A.ok) {
        set_error("vit_encode: workspace too small (%zu needed, %zu given)", A.off, workspace_bytes);
        return TEO_ERR_WORKSPACE;
    }
    if (frames_u8) TEO_TRY(teo_patchify_u8_nhwc(frames_u8, patches, n_frames, m->image, m->patch, m->kpad, stream));
    else TEO_TRY(teo_patchify_f32_nchw(pixel_values, patches, n_frames, m->image, m->patch, m->kpad, stream));
    GemmEpilogue none;
    TEO_TRY(launch_gemm(h, patches, m->kpad, static_cast<const bf16*>(m->patch_w), m->kpad, patch_out, d, n_frames * np, d, m->kpad,
                        none, gws, gws_bytes, stream, m->w_blocked));
    TEO_TRY(teo_vit_assemble_preln(patch_out, m->cls, m->pos, m->pre_ln_w, m->pre_ln_b, hidden, n_frames, np, d, m->eps, stream));
    fill_cu_seqlens_kernel<<<(n_frames + 256) / 256, 256, 0, stream>>>(cu, n_frames, np + 1);
    TEO_LAUNCH_CHECK("fill_cu_seqlens_kernel");
    h->launches += 3;
    const float scale = 1.0f / sqrtf(static_cast<float>(hd));
    // LayerNorm folded into the q/k/v and fc1 linears (teo_gemm_bf16_ex): no LayerNorm kernels, the normalised tensor never reaches HBM;
    // the out-proj / fc2 GEMMs emit the row statistics of the residual stream they write.  Needs the tiled GEMM schedule (rows > 128).
    const int st_slots = teo_gemm_stats_slots(rows, d, d);
    bool folded = rows > 128 && st_slots > 0 && m->layers_run > 0;
    for (int l = 0; l < m->layers_run && folded; ++l) {
        const teo_vit_layer& L = m->layers[l];
        folded = L.qkv_wf && L.qkv_c && L.qkv_bf && L.fc1_wf && L.fc1_c && L.fc1_bf;
    }
    if (folded) {
        TEO_TRY(teo_row_stats(hidden, stats, rows, d, st_slots, stream));
        h->launches += 1;
    }
    for (int l = 0; l < m->layers_run; ++l) {
        const teo_vit_layer& L = m->layers[l];
        if (folded) {
            GemmEpilogue e1;
            e1.ln_stats = stats; e1.ln_slots = st_slots; e1.ln_eps = m->eps;
            e1.ln_c = static_cast<const float*>(L.qkv_c); e1.ln_bias = static_cast<const float*>(L.qkv_bf);
            TEO_TRY(launch_gemm(h, hidden, d, static_cast<const bf16*>(L.qkv_wf), d, qkv, 3 * d, rows, 3 * d, d, e1, gws, gws_bytes, stream, m->w_blocked));
            TEO_TRY(launch_flash_attention_tc(h, qkv, 3 * d, qkv + d, 3 * d, qkv + 2 * d, 3 * d, attn, d, cu, n_frames, np + 1, rows, m->heads, hd,
                                              scale, 0, (np % 128 == 0 && flash_use_tc() && (hd == 64 || hd == 128)) ? 1 : 0, stream));
            GemmEpilogue e2;
            e2.bias = static_cast<const bf16*>(L.out_b);
            e2.residual = hidden; e2.ldr = d;
            e2.stats_out = stats;
            TEO_TRY(launch_gemm(h, attn, d, static_cast<const bf16*>(L.out_w), d, hidden, d, rows, d, d, e2, gws, gws_bytes, stream, m->w_blocked));
            GemmEpilogue e3;
            e3.ln_stats = stats; e3.ln_slots = st_slots; e3.ln_eps = m->eps;
            e3.ln_c = static_cast<const float*>(L.fc1_c); e3.ln_bias = static_cast<const float*>(L.fc1_bf);
            e3.act = m->act;
            TEO_TRY(launch_gemm(h, hidden, d, static_cast<const bf16*>(L.fc1_wf), d, mlp, m->inter, rows, m->inter, d, e3, gws, gws_bytes, stream, m->w_blocked));
            GemmEpilogue e4;
            e4.bias = static_cast<const bf16*>(L.fc2_b);
            e4.residual = hidden; e4.ldr = d;
            e4.stats_out = stats;
            TEO_TRY(launch_gemm(h, mlp, m->inter, static_cast<const bf16*>(L.fc2_w), m->inter, hidden, d, rows, d, m->inter, e4, gws, gws_bytes,
                                stream, m->w_blocked));
            h->launches += 1;
            continue;
        }
        TEO_TRY(teo_layernorm(hidden, L.ln1_w, L.ln1_b, ln_out, rows, d, m->eps, stream));
        GemmEpilogue e1;
        e1.bias = static_cast<const bf16*>(L.qkv_b);
        TEO_TRY(launch_gemm(h, ln_out, d, static_cast<const bf16*>(L.qkv_w), d, qkv, 3 * d, rows, 3 * d, d, e1, gws, gws_bytes, stream, m->w_blocked));
        if (flash_use_tc() && (hd == 64 || hd == 128))      // CLS row separately: 257 tokens = 1 + two 128-row query tiles
            TEO_TRY(launch_flash_attention_tc(h, qkv, 3 * d, qkv + d, 3 * d, qkv + 2 * d, 3 * d, attn, d, cu, n_frames, np + 1, rows,
                                              m->heads, hd, scale, 0, (np % 128 == 0) ? 1 : 0, stream));
        else
            TEO_TRY(launch_flash_attention(qkv, 3 * d, qkv + d, 3 * d, qkv + 2 * d, 3 * d, attn, d, cu, n_frames, np + 1, m->heads, hd,
                                           scale, 0, stream));
        GemmEpilogue e2;
        e2.bias = static_cast<const bf16*>(L.out_b);
        e2.residual = hidden;
        e2.ldr = d;
        TEO_TRY(launch_gemm(h, attn, d, static_cast<const bf16*>(L.out_w), d, hidden, d, rows, d, d, e2, gws, gws_bytes, stream, m->w_blocked));
        TEO_TRY(teo_layernorm(hidden, L.ln2_w, L.ln2_b, ln_out, rows, d, m->eps, stream));
        GemmEpilogue e3;
        e3.bias = static_cast<const bf16*>(L.fc1_b);
        e3.act = m->act;
        TEO_TRY(launch_gemm(h, ln_out, d, static_cast<const bf16*>(L.fc1_w), d, mlp, m->inter, rows, m->inter, d, e3, gws, gws_bytes, stream, m->w_blocked));
        GemmEpilogue e4;
        e4.bias = static_cast<const bf16*>(L.fc2_b);
        e4.residual = hidden;
        e4.ldr = d;
        TEO_TRY(launch_gemm(h, mlp, m->inter, static_cast<const bf16*>(L.fc2_w), m->inter, hidden, d, rows, d, m->inter, e4, gws, gws_bytes,
                            stream, m->w_blocked));
        h->launches += 3;
    }
    TEO_TRY(teo_vit_drop_cls(hidden, feats, n_frames, np, d, stream));
    h->launches += 1;
    return TEO_OK;
}

// ------------------------------------------------------------------------------ projector
extern "C" size_t teo_projector_workspace_bytes(const teo_projector* p, int rows) {
    if (!p || rows <= 0) return 0;
    if (p->exact) return projector_exact_workspace_bytes(p, rows);
    return al256(static_cast<size_t>(rows) * p->hidden * sizeof(bf16)) + al256(teo_gemm_workspace_bytes(rows, p->hidden, p->hidden));
}

extern "C" int teo_projector_mlp2x(teo_handle* h, const teo_projector* p, const void* feats, int rows, void* out, void* workspace,
                                   size_t workspace_bytes, void* stream_) {
    TEO_CHECK_ARG(h && p && feats && out, "projector: null pointer");
    TEO_CHECK_ARG(rows > 0, "projector: rows=%d", rows);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (p->exact) return projector_exact(h, p, feats, rows, out, workspace, workspace_bytes, stream);
    Arena A(workspace, workspace_bytes);
    bf16* mid = A.take<bf16>(static_cast<size_t>(rows) * p->hidden);
    const size_t gws = teo_gemm_workspace_bytes(rows, p->hidden, p->hidden);
    uint8_t* gw = A.take<uint8_t>(gws);
    if (!A.ok) {
        set_error("projector: workspace too small (%zu needed, %zu given)", A.off, workspace_bytes);
        return TEO_ERR_WORKSPACE;
    }
    GemmEpilogue e0;
    e0.bias = static_cast<const bf16*>(p->b0);
    e0.act = TEO_ACT_GELU;
    TEO_TRY(launch_gemm(h, static_cast<const bf16*>(feats), p->in_dim, static_cast<const bf16*>(p->w0), p->in_dim, mid, p->hidden, rows,
                        p->hidden, p->in_dim, e0, gw, gws, stream, p->w_blocked));
    GemmEpilogue e1;
    e1.bias = static_cast<const bf16*>(p->b2);
    TEO_TRY(launch_gemm(h, mid, p->hidden, static_cast<const bf16*>(p->w2), p->hidden, out, p->hidden, rows, p->hidden, p->hidden, e1, gw,
                        gws, stream, p->w_blocked));
    return TEO_OK;
}

// ------------------------------------------------------------------------------ LLaMA prefill
struct PrefillWs {
    bf16 *norm_out, *qkv, *attn, *gate_up, *act, *last_x, *last_norm;
    uint8_t* gemm_ws;
    size_t gemm_ws_bytes;
};
static size_t prefill_ws_layout(const teo_llama_model* m, int T, int B, Arena& A, PrefillWs* w) {
    const size_t h = m->hidden, I = m->inter;
    PrefillWs t;
    t.norm_out = A.take<bf16>(static_cast<size_t>(T) * h);
    t.qkv = A.take<bf16>(static_cast<size_t>(T) * 3 * h);
    t.attn = A.take<bf16>(static_cast<size_t>(T) * h);
    t.gate_up = A.take<bf16>(static_cast<size_t>(T) * 2 * I);
    t.act = A.take<bf16>(static_cast<size_t>(T) * I);
    t.last_x = A.take<bf16>(static_cast<size_t>(B) * h);
    t.last_norm = A.take<bf16>(static_cast<size_t>(B) * h);
    size_t g = teo_gemm_workspace_bytes(B, m->vocab, m->hidden);
    g = std::max(g, teo_gemm_workspace_bytes(T, 3 * m->hidden, m->hidden));
    g = std::max(g, teo_gemm_workspace_bytes(T, 2 * m->inter, m->hidden));
    g = std::max(g, teo_gemm_workspace_bytes(T, m->hidden, m->inter));
    t.gemm_ws_bytes = g;
    t.gemm_ws = A.take<uint8_t>(g);
    if (w) *w = t;
    return A.off;
}
extern "C" size_t teo_llama_prefill_workspace_bytes(const teo_llama_model* m, int tokens, int n_seqs) {
    if (!m || tokens <= 0 || n_seqs <= 0) return 0;
    if (m->exact) return llama_prefill_exact_workspace_bytes(m, tokens, n_seqs);
    Arena A(nullptr, 0);
    return prefill_ws_layout(m, tokens, n_seqs, A, nullptr);
}

static int llama_layer_mlp(teo_handle* h, const teo_llama_model* m, const teo_llama_layer& L, bf16* x, int rows, bf16* norm_out,
                           bf16* gate_up, bf16* act, void* gws, size_t gws_bytes, cudaStream_t stream) {
    const int hd = m->hidden, I = m->inter;
    TEO_TRY(teo_rmsnorm(x, L.post_norm, norm_out, rows, hd, m->eps, stream));
    if (m->gate_up_interleaved && rows > 128) {
        // tiled GEMM: SwiGLU in the epilogue, the [rows, 2I] gate/up activations never reach HBM
        GemmEpilogue pairs;
        pairs.act = TEO_ACT_SWIGLU_PAIRS;
        TEO_TRY(launch_gemm(h, norm_out, hd, static_cast<const bf16*>(L.gate_up_w), hd, act, I, rows, 2 * I, hd, pairs, gws, gws_bytes, stream,
                            m->w_blocked));
    } else {
        GemmEpilogue none;
        TEO_TRY(launch_gemm(h, norm_out, hd, static_cast<const bf16*>(L.gate_up_w), hd, gate_up, 2 * I, rows, 2 * I, hd, none, gws, gws_bytes,
                            stream, m->w_blocked));
        TEO_TRY(launch_swiglu(gate_up, act, rows, I, m->gate_up_interleaved, stream));
        h->launches++;
    }
    GemmEpilogue res;
    res.residual = x;
    res.ldr = hd;
    TEO_TRY(launch_gemm(h, act, I, static_cast<const bf16*>(L.down_w), I, x, hd, rows, hd, I, res, gws, gws_bytes, stream, m->w_blocked));
    h->launches += 1;                  // the RMSNorm (GEMMs and the SwiGLU kernel count themselves)
    return TEO_OK;
}

extern "C" int teo_llama_prefill(teo_handle* h, const teo_llama_model* m, void* x_, int tokens, const void* cu_seqlens,
                                 const void* positions, const void* seq_ids, const void* last_rows, int n_seqs, int max_seqlen,
                                 const void* block_table, int max_pages, void* logits, void* workspace, size_t workspace_bytes,
                                 void* stream_) {
    TEO_CHECK_ARG(h && m && x_ && cu_seqlens && positions && seq_ids && last_rows && block_table && logits, "llama_prefill: null pointer");
    TEO_CHECK_ARG(tokens > 0 && n_seqs > 0 && max_seqlen > 0, "llama_prefill: bad sizes");
    TEO_CHECK_ARG(m->hidden % m->heads == 0 && m->layer != nullptr, "llama_prefill: bad model");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (m->exact)
        return llama_prefill_exact(h, m, x_, tokens, positions, seq_ids, last_rows, n_seqs, max_seqlen, block_table, max_pages, logits,
                                   workspace, workspace_bytes, stream);
    bf16* x = static_cast<bf16*>(x_);
    const int hdim = m->hidden, hd = hdim / m->heads;
    Arena A(workspace, workspace_bytes);
    PrefillWs w;
    prefill_ws_layout(m, tokens, n_seqs, A, &w);
    if (!A.ok) {
        set_error("llama_prefill: workspace too small (%zu needed, %zu given)", A.off, workspace_bytes);
        return TEO_ERR_WORKSPACE;
    }
    const float scale = 1.0f / sqrtf(static_cast<float>(hd));
    for (int l = 0; l < m->layers; ++l) {
        const teo_llama_layer& L = m->layer[l];
        TEO_TRY(teo_rmsnorm(x, L.in_norm, w.norm_out, tokens, hdim, m->eps, stream));
        GemmEpilogue none;
        TEO_TRY(launch_gemm(h, w.norm_out, hdim, static_cast<const bf16*>(L.qkv_w), hdim, w.qkv, 3 * hdim, tokens, 3 * hdim, hdim, none,
                            w.gemm_ws, w.gemm_ws_bytes, stream, m->w_blocked));
        TEO_TRY(launch_rope_kv_write(w.qkv, static_cast<const int*>(positions), static_cast<const int*>(seq_ids), L.kv_pages,
                                     static_cast<const int*>(block_table), max_pages, tokens, m->heads, hd, m->page_size,
                                     static_cast<const float*>(m->rope_cos), static_cast<const float*>(m->rope_sin), stream));
        if (flash_use_tc() && (hd == 64 || hd == 128))
            TEO_TRY(launch_flash_attention_tc(h, w.qkv, 3 * hdim, w.qkv + hdim, 3 * hdim, w.qkv + 2 * hdim, 3 * hdim, w.attn, hdim,
                                              static_cast<const int*>(cu_seqlens), n_seqs, max_seqlen, tokens, m->heads, hd, scale, 1, 0,
                                              stream));
        else
            TEO_TRY(launch_flash_attention(w.qkv, 3 * hdim, w.qkv + hdim, 3 * hdim, w.qkv + 2 * hdim, 3 * hdim, w.attn, hdim,
                                           static_cast<const int*>(cu_seqlens), n_seqs, max_seqlen, m->heads, hd, scale, 1, stream));
        GemmEpilogue res;
        res.residual = x;
        res.ldr = hdim;
        TEO_TRY(launch_gemm(h, w.attn, hdim, static_cast<const bf16*>(L.o_w), hdim, x, hdim, tokens, hdim, hdim, res, w.gemm_ws,
                            w.gemm_ws_bytes, stream, m->w_blocked));
        h->launches += 3;
        TEO_TRY(llama_layer_mlp(h, m, L, x, tokens, w.norm_out, w.gate_up, w.act, w.gemm_ws, w.gemm_ws_bytes, stream));
    }
    // only the last position of every sequence feeds lm_head (the reference
    // materialises logits for all positions and discards them)
    TEO_TRY(teo_splice_embed(x, nullptr, last_rows, w.last_x, n_seqs, hdim, stream));
    TEO_TRY(teo_rmsnorm(w.last_x, m->final_norm, w.last_norm, n_seqs, hdim, m->eps, stream));
    GemmEpilogue lg;
    lg.out_fp32 = 1;
    TEO_TRY(launch_gemm(h, w.last_norm, hdim, static_cast<const bf16*>(m->lm_head), hdim, logits, m->vocab, n_seqs, m->vocab, hdim, lg,
                        w.gemm_ws, w.gemm_ws_bytes, stream, m->w_blocked));
    h->launches += 2;
    return TEO_OK;
}

// ------------------------------------------------------------------------------ LLaMA decode step
struct DecodeWs {
    bf16 *x, *norm_out, *qkv, *attn, *gate_up, *act;
    uint8_t *gemm_ws, *attn_ws, *chain_ws;
    size_t gemm_ws_bytes, attn_ws_bytes, chain_ws_bytes;
};
static bool decode_chain_usable(const teo_llama_model* m, int B) {
    return m->w_blocked && decode_chain_shape_ok(B, m->hidden, m->inter, m->vocab, m->heads);
}
static size_t decode_ws_layout(const teo_llama_model* m, int B, Arena& A, DecodeWs* w) {
    const size_t h = m->hidden, I = m->inter;
    DecodeWs t;
    t.x = A.take<bf16>(static_cast<size_t>(B) * h);
    t.norm_out = A.take<bf16>(static_cast<size_t>(B) * h);
    t.qkv = A.take<bf16>(static_cast<size_t>(B) * 3 * h);
    t.attn = A.take<bf16>(static_cast<size_t>(B) * h);
    t.gate_up = A.take<bf16>(static_cast<size_t>(B) * 2 * I);
    t.act = A.take<bf16>(static_cast<size_t>(B) * I);
    size_t g = teo_gemm_workspace_bytes(B, m->vocab, m->hidden);
    g = std::max(g, teo_gemm_workspace_bytes(B, 3 * m->hidden, m->hidden));
    g = std::max(g, teo_gemm_workspace_bytes(B, 2 * m->inter, m->hidden));
    g = std::max(g, teo_gemm_workspace_bytes(B, m->hidden, m->inter));
    t.gemm_ws_bytes = g;
    t.gemm_ws = A.take<uint8_t>(g);
    t.attn_ws_bytes = teo_decode_attention_workspace_bytes(B, m->heads, m->hidden / m->heads, 32);
    t.attn_ws = A.take<uint8_t>(t.attn_ws_bytes);
    t.chain_ws_bytes = 0;
    if (decode_chain_usable(m, B)) {       // partial regions of the four phases of one chain launch (148-CTA split: the upper bound)
        const int hh = m->hidden, II = m->inter;
        t.chain_ws_bytes = decode_chain_workspace_bytes(B, hh, hh, 148) + decode_chain_workspace_bytes(B, 2 * II, hh, 148) +
                           decode_chain_workspace_bytes(B, hh, II, 148) +
                           std::max(decode_chain_workspace_bytes(B, 3 * hh, hh, 148), decode_chain_workspace_bytes(B, m->vocab, hh, 148));
    }
    t.chain_ws = A.take<uint8_t>(t.chain_ws_bytes);
    if (w) *w = t;
    return A.off;
}
extern "C" size_t teo_llama_decode_workspace_bytes(const teo_llama_model* m, int n_seqs, int max_seq_len) {
    (void)max_seq_len;
    if (!m || n_seqs <= 0) return 0;
    if (m->exact) return llama_decode_exact_workspace_bytes(m, n_seqs);
    Arena A(nullptr, 0);
    return decode_ws_layout(m, n_seqs, A, nullptr);
}

// In-kernel reduction of the decode gate/up GEMM (gemm.cu, SkFuse): TEO_SK_FUSE=0|1 (A/B measurements)
static bool sk_fuse_enabled() {
    static const bool on = [] {
        const char* e = getenv("TEO_SK_FUSE");
        return e != nullptr && e[0] == '1';
    }();
    return on;
}

extern "C" int teo_llama_decode_step(teo_handle* h, const teo_llama_model* m, void* next_ids, void* seq_lens, void* finished,
                                     void* tokens, int max_new, void* step_ptr, int n_seqs, int max_seq_len, const void* block_table,
                                     int max_pages, void* logits, int eos_id, void* workspace, size_t workspace_bytes, void* stream_) {
    TEO_CHECK_ARG(h && m && next_ids && seq_lens && finished && tokens && step_ptr && block_table && logits, "llama_decode_step: null pointer");
    TEO_CHECK_ARG(n_seqs > 0 && max_seq_len > 0 && max_new > 0, "llama_decode_step: bad sizes");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (m->exact)
        return llama_decode_step_exact(h, m, next_ids, seq_lens, finished, tokens, max_new, step_ptr, n_seqs, max_seq_len, block_table,
                                       max_pages, logits, eos_id, workspace, workspace_bytes, stream);
    const int hdim = m->hidden, hd = hdim / m->heads;
    Arena A(workspace, workspace_bytes);
    DecodeWs w;
    decode_ws_layout(m, n_seqs, A, &w);
    if (!A.ok) {
        set_error("llama_decode_step: workspace too small (%zu needed, %zu given)", A.off, workspace_bytes);
        return TEO_ERR_WORKSPACE;
    }
    const float scale = 1.0f / sqrtf(static_cast<float>(hd));
    const int I = m->inter;
    struct PdlScope {           // programmatic dependent launch for every kernel of the step (see common.h)
        explicit PdlScope(bool on) { set_pdl(on); }
        ~PdlScope() { set_pdl(false); }
    } pdl_scope(h->pdl);
    const bool fused = n_seqs <= 128 && hdim >= 256 && hdim % 32 == 0 && hdim <= 8192 && I % 4 == 0;   // swap-AB split-K regime of the GEMM
    TEO_TRY(teo_splice_embed(m->embed, nullptr, next_ids, w.x, n_seqs, hdim, stream));
    TEO_TRY(teo_rmsnorm(w.x, m->layer[0].in_norm, w.norm_out, n_seqs, hdim, m->eps, stream));
    h->launches += 2;
    if (fused && h->decode_chain && decode_chain_usable(m, n_seqs) && h->num_sms <= 148) {
        // ---- persistent chain: per layer ONE kernel for o → norm → gate/up → SwiGLU → down → norm → next qkv → RoPE/KV (or lm_head),
        //      then the attention kernel(s); 3 launches per layer instead of 10 (decode_chain.cu)
        const int* lens = static_cast<const int*>(seq_lens);
        const int* bt = static_cast<const int*>(block_table);
        const float* rc = static_cast<const float*>(m->rope_cos);
        const float* rs = static_cast<const float*>(m->rope_sin);
        auto qkv_spec = [&](int l) {
            ChainSpec s{};
            s.W = m->layer[l].qkv_w; s.N = 3 * hdim; s.K = hdim; s.A = w.norm_out; s.lda = hdim; s.reduce = 2;
            s.qkv = w.qkv; s.kv_pages = static_cast<bf16*>(m->layer[l].kv_pages);
            return s;
        };
        {
            ChainSpec first = qkv_spec(0);
            TEO_TRY(launch_decode_chain(h, &first, 1, n_seqs, lens, bt, max_pages, m->heads, hd, m->page_size, I, m->gate_up_interleaved, rc, rs,
                                        m->eps, w.chain_ws, w.chain_ws_bytes, stream));
        }
        for (int l = 0; l < m->layers; ++l) {
            const teo_llama_layer& L = m->layer[l];
            const bool last = l + 1 == m->layers;
            TEO_TRY(launch_decode_attention(h, w.qkv, 3 * hdim, static_cast<const bf16*>(L.kv_pages), bt, max_pages, lens, 1, w.attn, n_seqs,
                                            m->heads, hd, m->page_size, max_seq_len, scale, w.attn_ws, w.attn_ws_bytes, stream));
            ChainSpec sp[4] = {};
            sp[0].W = L.o_w; sp[0].N = hdim; sp[0].K = hdim; sp[0].A = w.attn; sp[0].lda = hdim; sp[0].reduce = 0;
            sp[0].x = w.x; sp[0].norm_w = static_cast<const bf16*>(L.post_norm); sp[0].y = w.norm_out;
            sp[1].W = L.gate_up_w; sp[1].N = 2 * I; sp[1].K = hdim; sp[1].A = w.norm_out; sp[1].lda = hdim; sp[1].reduce = 1;
            sp[1].act = w.act;
            sp[2].W = L.down_w; sp[2].N = hdim; sp[2].K = I; sp[2].A = w.act; sp[2].lda = I; sp[2].reduce = 0;
            sp[2].x = w.x; sp[2].norm_w = static_cast<const bf16*>(last ? m->final_norm : m->layer[l + 1].in_norm); sp[2].y = w.norm_out;
            if (!last) {
                sp[3] = qkv_spec(l + 1);
            } else {
                sp[3].W = m->lm_head; sp[3].N = m->vocab; sp[3].K = hdim; sp[3].A = w.norm_out; sp[3].lda = hdim; sp[3].reduce = 3;
                sp[3].logits = static_cast<float*>(logits);
            }
            TEO_TRY(launch_decode_chain(h, sp, 4, n_seqs, lens, bt, max_pages, m->heads, hd, m->page_size, I, m->gate_up_interleaved, rc, rs,
                                        m->eps, w.chain_ws, w.chain_ws_bytes, stream));
        }
        if (h->temperature > 0.f)
            TEO_TRY(launch_sample_step(static_cast<const float*>(logits), m->vocab, h->temperature, h->top_k, h->sample_seed,
                                       static_cast<uint8_t*>(finished), static_cast<int*>(tokens), max_new, 0, static_cast<int*>(step_ptr),
                                       static_cast<int*>(next_ids), static_cast<int*>(seq_lens), n_seqs, eos_id, stream, h->sample_seed_ptr));
        else
            TEO_TRY(launch_argmax_step(static_cast<const float*>(logits), m->vocab, static_cast<uint8_t*>(finished), static_cast<int*>(tokens),
                                       max_new, 0, static_cast<int*>(step_ptr), static_cast<int*>(next_ids), static_cast<int*>(seq_lens), n_seqs,
                                       eos_id, stream));
        h->launches += 2;
        return TEO_OK;
    }
    for (int l = 0; l < m->layers; ++l) {
        const teo_llama_layer& L = m->layer[l];
        const void* next_norm = (l + 1 < m->layers) ? m->layer[l + 1].in_norm : m->final_norm;
        if (fused) {
            PartialInfo pi{};
            // qkv: GEMM partials → reduce + RoPE + KV-page write (position of the new token = tokens cached so far)
            TEO_TRY(launch_gemm_partials(h, w.norm_out, hdim, static_cast<const bf16*>(L.qkv_w), hdim, n_seqs, 3 * hdim, hdim, w.gemm_ws,
                                         w.gemm_ws_bytes, &pi, stream, m->w_blocked));
            TEO_TRY(launch_reduce_rope_kv_write(pi, w.qkv, static_cast<const int*>(seq_lens),
                                                L.kv_pages, static_cast<const int*>(block_table), max_pages, n_seqs, m->heads, hd,
                                                m->page_size, static_cast<const float*>(m->rope_cos), static_cast<const float*>(m->rope_sin),
                                                stream));
            TEO_TRY(launch_decode_attention(h, w.qkv, 3 * hdim, static_cast<const bf16*>(L.kv_pages), static_cast<const int*>(block_table),
                                            max_pages, static_cast<const int*>(seq_lens), 1, w.attn, n_seqs, m->heads, hd, m->page_size,
                                            max_seq_len, scale, w.attn_ws, w.attn_ws_bytes, stream));
            // o_proj partials → reduce + residual + post-attention RMSNorm
            TEO_TRY(launch_gemm_partials(h, w.attn, hdim, static_cast<const bf16*>(L.o_w), hdim, n_seqs, hdim, hdim, w.gemm_ws, w.gemm_ws_bytes,
                                         &pi, stream, m->w_blocked));
            TEO_TRY(launch_reduce_residual_rmsnorm(pi, w.x, static_cast<const bf16*>(L.post_norm),
                                                   w.norm_out, n_seqs, hdim, m->eps, stream));
            // gate/up partials → reduce + SwiGLU
            if (sk_fuse_enabled() && m->gate_up_interleaved && I % 64 == 0) {
                // … with the reduction + SwiGLU inside the GEMM (the CTA holding slot 0 of a weight tile reduces it)
                SkFuse fz;
                fz.kind = 1;
                fz.out = w.act;
                fz.inter = I;
                TEO_TRY(launch_gemm_partials(h, w.norm_out, hdim, static_cast<const bf16*>(L.gate_up_w), hdim, n_seqs, 2 * I, hdim, w.gemm_ws,
                                             w.gemm_ws_bytes, &pi, stream, m->w_blocked, &fz));
                h->launches -= 1;
            } else {
                TEO_TRY(launch_gemm_partials(h, w.norm_out, hdim, static_cast<const bf16*>(L.gate_up_w), hdim, n_seqs, 2 * I, hdim, w.gemm_ws,
                                             w.gemm_ws_bytes, &pi, stream, m->w_blocked));
                TEO_TRY(launch_reduce_swiglu(pi, w.act, n_seqs, I, m->gate_up_interleaved, stream));
            }
            // down_proj partials → reduce + residual + the NEXT layer's input RMSNorm (or the final norm)
            TEO_TRY(launch_gemm_partials(h, w.act, I, static_cast<const bf16*>(L.down_w), I, n_seqs, hdim, I, w.gemm_ws, w.gemm_ws_bytes, &pi,
                                         stream, m->w_blocked));
            TEO_TRY(launch_reduce_residual_rmsnorm(pi, w.x, static_cast<const bf16*>(next_norm),
                                                   w.norm_out, n_seqs, hdim, m->eps, stream));
            h->launches += 4;
        } else {
            GemmEpilogue none;
            TEO_TRY(launch_gemm(h, w.norm_out, hdim, static_cast<const bf16*>(L.qkv_w), hdim, w.qkv, 3 * hdim, n_seqs, 3 * hdim, hdim, none,
                                w.gemm_ws, w.gemm_ws_bytes, stream, m->w_blocked));
            TEO_TRY(launch_rope_kv_write(w.qkv, static_cast<const int*>(seq_lens), nullptr, L.kv_pages, static_cast<const int*>(block_table),
                                         max_pages, n_seqs, m->heads, hd, m->page_size, static_cast<const float*>(m->rope_cos),
                                         static_cast<const float*>(m->rope_sin), stream));
            TEO_TRY(launch_decode_attention(h, w.qkv, 3 * hdim, static_cast<const bf16*>(L.kv_pages), static_cast<const int*>(block_table),
                                            max_pages, static_cast<const int*>(seq_lens), 1, w.attn, n_seqs, m->heads, hd, m->page_size,
                                            max_seq_len, scale, w.attn_ws, w.attn_ws_bytes, stream));
            GemmEpilogue res;
            res.residual = w.x;
            res.ldr = hdim;
            TEO_TRY(launch_gemm(h, w.attn, hdim, static_cast<const bf16*>(L.o_w), hdim, w.x, hdim, n_seqs, hdim, hdim, res, w.gemm_ws,
                                w.gemm_ws_bytes, stream, m->w_blocked));
            h->launches += 1;
            TEO_TRY(llama_layer_mlp(h, m, L, w.x, n_seqs, w.norm_out, w.gate_up, w.act, w.gemm_ws, w.gemm_ws_bytes, stream));
            TEO_TRY(teo_rmsnorm(w.x, next_norm, w.norm_out, n_seqs, hdim, m->eps, stream));
            h->launches += 1;
        }
    }
    GemmEpilogue lg;
    lg.out_fp32 = 1;
    TEO_TRY(launch_gemm(h, w.norm_out, hdim, static_cast<const bf16*>(m->lm_head), hdim, logits, m->vocab, n_seqs, m->vocab, hdim, lg,
                        w.gemm_ws, w.gemm_ws_bytes, stream, m->w_blocked));
    if (h->temperature > 0.f)
        TEO_TRY(launch_sample_step(static_cast<const float*>(logits), m->vocab, h->temperature, h->top_k, h->sample_seed,
                                   static_cast<uint8_t*>(finished), static_cast<int*>(tokens), max_new, 0, static_cast<int*>(step_ptr),
                                   static_cast<int*>(next_ids), static_cast<int*>(seq_lens), n_seqs, eos_id, stream, h->sample_seed_ptr));
    else
        TEO_TRY(launch_argmax_step(static_cast<const float*>(logits), m->vocab, static_cast<uint8_t*>(finished), static_cast<int*>(tokens),
                                   max_new, 0, static_cast<int*>(step_ptr), static_cast<int*>(next_ids), static_cast<int*>(seq_lens), n_seqs,
                                   eos_id, stream));
    h->launches += 3;
    return TEO_OK;
}
