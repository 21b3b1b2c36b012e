// Host-side shared declarations for the C-ABI implementation (internal; the public surface is
// include/teochat_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <unordered_map>

#include "../../include/teochat_b200.h"

namespace teo {

typedef __nv_bfloat16 bf16;

void set_error(const char* fmt, ...);

// Programmatic dependent launch (PDL) switch for the launch helpers below: while it is on (decode step), kernels
// are launched with cudaLaunchAttributeProgrammaticStreamSerialization so that a kernel's launch, prologue and
// (for the GEMM) weight prefetch overlap the tail of its predecessor; every kernel calls griddepcontrol.wait
// before touching data its predecessor produced.
bool pdl_enabled();
void set_pdl(bool on);
// which kernel classes get the attribute (bit 0 GEMM, bit 1 attention, bit 2 small glue kernels); TEO_PDL_MASK env
int pdl_mask();
constexpr int PDL_GEMM = 1, PDL_ATTN = 2, PDL_SMALL = 4;

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kc(int kind, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_enabled() && (pdl_mask() & kind)) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    return launch_kc(PDL_SMALL, kernel, grid, block, smem, stream, static_cast<Args&&>(args)...);
}

#define TEO_CHECK_ARG(cond, ...)            \
    do {                                    \
        if (!(cond)) {                      \
            ::teo::set_error(__VA_ARGS__);  \
            return TEO_ERR_BAD_ARG;         \
        }                                   \
    } while (0)

#define TEO_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            ::teo::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return TEO_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define TEO_LAUNCH_CHECK(name)                                                                 \
    do {                                                                                       \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess) {                                                              \
            ::teo::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));        \
            return TEO_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

#define TEO_TRY(call)              \
    do {                           \
        int r__ = (call);          \
        if (r__ != TEO_OK) return r__; \
    } while (0)

struct TmapKey {
    const void* ptr;
    uint64_t rows, cols, ld;
    uint32_t box_rows;
    bool operator==(const TmapKey& o) const {
        return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
    }
};
struct TmapKeyHash {
    size_t operator()(const TmapKey& k) const {
        uint64_t h = reinterpret_cast<uint64_t>(k.ptr);
        h = h * 0x9E3779B97F4A7C15ULL ^ k.rows;
        h = h * 0x9E3779B97F4A7C15ULL ^ k.cols;
        h = h * 0x9E3779B97F4A7C15ULL ^ k.ld;
        h = h * 0x9E3779B97F4A7C15ULL ^ k.box_rows;
        return static_cast<size_t>(h);
    }
};

}  // namespace teo

// The opaque per-device handle of the C-ABI.
struct teo_handle {
    int device = 0;
    int num_sms = 148;
    unsigned long long launches = 0;   // kernels launched through this handle (bench's gpu_launches)
    float temperature = 0.f;           // > 0 → teo_llama_decode_step samples (teo_set_sampling)
    int top_k = 50;
    unsigned long long sample_seed = 0;
    const unsigned long long* sample_seed_ptr = nullptr;   // device-resident seed (teo_set_sampling_seed_device); wins over sample_seed
    bool pdl = true;                   // programmatic dependent launch inside teo_llama_decode_step (teo_set_pdl)
    bool decode_chain = true;          // teo_llama_decode_step uses the persistent chain kernel (teo_set_decode_chain; TEO_DEC_CHAIN=0)
    void* chain_sync = nullptr;        // 256 B of device memory: grid-barrier counters of the decode chain kernel (decode_chain.cu)
    int* sk_flags = nullptr;           // 4 KiB of device memory: per-tile arrival counters of the stream-K GEMM's in-kernel reduction (gemm.cu)
    std::unordered_map<teo::TmapKey, CUtensorMap, teo::TmapKeyHash> tmaps;
};

namespace teo {

// Row-major bf16 matrix [rows, cols] with leading dimension ld (elements) → TMA descriptor with
// a {64 cols, box_rows} box under SWIZZLE_128B (cached per handle; *out is a copy, valid whatever the cache does later).
int get_tmap_bf16(teo_handle* h, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                  CUtensorMap* out);

struct GemmEpilogue {
    const bf16* bias = nullptr;       // [N]
    const bf16* residual = nullptr;   // [M, ldr]
    int ldr = 0;
    int act = TEO_ACT_NONE;
    int out_fp32 = 0;                 // C is float instead of bf16
    int residual_f32 = 0;             // residual is float [M, ldr] (needs out_fp32; may alias C)
    // LayerNorm folded into this GEMM (tiled schedule, bf16 output): A holds the UN-normalised rows x, W the weights pre-multiplied
    // by the LayerNorm gain, W'[n,k] = bf16(γ_k·W[n,k]);  C = act(rstd_m·(x·W'ᵀ − mean_m·ln_c) + ln_bias) with
    // ln_c[n] = Σ_k W'[n,k], ln_bias[n] = Σ_k β_k·W[n,k] + b[n]  (= LayerNorm(x)·Wᵀ + b up to the rounding of W').
    // ln_stats: f32 [M][ln_slots][2] partial (Σx, Σx²) per input row, written by the GEMM that produced x (stats_out) or by
    // teo_row_stats.  `bias` must be null when ln_stats is set.
    const float* ln_stats = nullptr;
    const float* ln_c = nullptr;
    const float* ln_bias = nullptr;
    int ln_slots = 0;
    float ln_eps = 0.f;
    float* stats_out = nullptr;       // f32 [M][teo_gemm_stats_slots(M,N,K)][2]: (Σ, Σ²) partials of the bf16 rows this GEMM writes
    int k_planes = 1;                 // exact mode: A is [M, k_planes·K] — an fp32 activation split into k_planes bf16 terms
                                      // (hi | mid | lo) stored side by side; every plane is multiplied by the same W[N,K]
};

// C[M,N] = epilogue(A[M,K] · W[N,K]^T).  Chooses the swap-AB / split-K schedule for small M.
// workspace: teo_gemm_workspace_bytes(M,N,K) bytes (only used by the split-K schedule).
int launch_gemm(teo_handle* h, const bf16* A, int lda, const bf16* W, int ldw, void* C, int ldc, int M, int N, int K,
                const GemmEpilogue& ep, void* workspace, size_t workspace_bytes, cudaStream_t stream, int w_blocked = 0);

// fp32 partial sums left by the small-M (stream-K) GEMM schedule: P[slot][row][col], slot < partial_count(col).
// CTA c of that schedule owns k-blocks [c·q, (c+1)·q) of the flattened (128-column tile, k-block) space, so the
// column tile t = col/128 was touched by CTAs (t·kb)/q … ((t+1)·kb − 1)/q, one slot each.
struct PartialInfo {
    const float* P;
    long long stride;     // elements between slots (= rows · cols)
    int kb, q, grid;
};
__host__ __device__ inline int partial_count(const PartialInfo& pi, int col) {
    const int t = col >> 7;                       // tiles·kb < 2^31 for every shape here (≤ 250 tiles × 172 k-blocks)
    const int first = (t * pi.kb) / pi.q;
    int last = ((t + 1) * pi.kb - 1) / pi.q;
    if (last > pi.grid - 1) last = pi.grid - 1;
    return last - first + 1;
}

// next-token kernels (kernels_misc.cu); step_ptr / seed_ptr != nullptr: column / seed read from device memory (graph replay)
int launch_argmax_step(const float* logits, int vocab, uint8_t* finished, int* tokens, int max_new, int step_host, int* step_ptr,
                       int* next_ids, int* seq_lens, int n_seqs, int eos_id, cudaStream_t stream);
int launch_sample_step(const float* logits, int vocab, float temperature, int top_k, unsigned long long seed, uint8_t* finished, int* tokens,
                       int max_new, int step_host, int* step_ptr, int* next_ids, int* seq_lens, int n_seqs, int eos_id, cudaStream_t stream,
                       const unsigned long long* seed_ptr = nullptr);

// Persistent decode chain kernel (decode_chain.cu): one launch runs up to four weight-streaming GEMM phases with their fused
// reductions.  One ChainSpec per phase: weights (blocked layout), activations [B, K] (row stride lda), reduction + operands.
struct ChainSpec {
    const void* W;
    int N, K;
    const bf16* A;
    int lda;
    int reduce;                 // 0 residual + RMSNorm, 1 SwiGLU, 2 RoPE + KV write, 3 logits
    bf16* x;
    const bf16* norm_w;
    bf16* y;
    bf16* act;
    bf16* qkv;
    bf16* kv_pages;
    float* logits;
};
bool decode_chain_enabled();
bool decode_chain_shape_ok(int B, int hidden, int inter, int vocab, int n_heads);
size_t decode_chain_workspace_bytes(int B, int N, int K, int grid);
int launch_decode_chain(teo_handle* h, const ChainSpec* specs, int n_phases, int B, const int* positions, const int* block_table, int max_pages,
                        int n_heads, int head_dim, int page_size, int inter, int interleaved, const float* rope_cos, const float* rope_sin,
                        float eps, void* ws, size_t ws_bytes, cudaStream_t stream);

// Small-M (decode) GEMM that stops at the fp32 partials; the consumer kernel reduces them (fixed slot order).
// With an SkFuse the GEMM reduces them itself: the CTA that holds slot 0 of a 128-row weight tile waits for the other slots'
// arrival flags and runs the same reduction code as the stand-alone glue kernel on that tile (bit-identical results) — one kernel
// and two grid-completion hand-offs less on the decode step's dependency chain.  kind 1: SwiGLU over interleaved gate/up rows.
struct SkFuse {
    int kind = 0;
    bf16* out = nullptr;       // act [M, inter]
    int inter = 0;
};
int launch_gemm_partials(teo_handle* h, const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, void* workspace,
                         size_t workspace_bytes, PartialInfo* info, cudaStream_t stream, int w_blocked = 0, const SkFuse* fuse = nullptr);

}  // namespace teo
