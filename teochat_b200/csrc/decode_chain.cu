// Persistent decode CHAIN kernel: the four weight-streaming GEMMs between two attention calls of the decode step —
//   o_proj → (+residual, RMSNorm) → gate/up → (SwiGLU) → down → (+residual, RMSNorm) → next layer's qkv (→ RoPE + KV write)
// (HF-4.31 LlamaDecoderLayer, called per token through llava_llama.py:88-99) — as ONE launch instead of eight.
//
// Why: at decode batch sizes the GEMMs are pure weight streaming (13.2 GB per step), and with one kernel
// per GEMM plus one per reduction every matrix paid ≈ 13 µs of launch + dependency latency on 15–35 µs of streaming
// (115 µs per layer against 63 µs of HBM time).  Here the grid is persistent (one CTA per SM), the phases
// are separated by grid-wide barriers through L2 instead of kernel boundaries, and the TMA producer keeps streaming the
// NEXT matrix's weight tiles into the shared-memory ring while the current phase's partials are reduced — only the small
// activation operand waits for the barrier.
//
// Per phase p (weights W_p [N_p, K_p] in the blocked layout, activations A_p [B, K_p] bf16, B ≤ 128):
//   warp 0      TMA producer: stream-K share of the flattened (128-row weight tile, k-block) space — CTA c owns k-blocks
//               [c·q_p, (c+1)·q_p) — W tiles first (they are constants), A tiles once the phase's input is complete
//   warp 1      MMA issuer: tcgen05.mma 128 × BN × 16 (swap-AB: weight rows on M, batch on N), fp32 accumulation in TMEM
//   warp 2      TMEM allocator
//   warps 4-11  epilogue: TMEM → fp32 partials P_p[slot][B][N_p] (slot = CTA − first CTA of the tile);
//               grid barrier "partials complete"; then ALL CTAs' epilogue warps run the phase's reduction (decode_reduce.cuh,
//               the same code and summation order as the stand-alone kernels → bit-identical results);
//               grid barrier "next input complete" (the producer waits on it before its A loads).
// Barriers are monotone counters in a small device buffer owned by the handle (zero between launches: the last CTA to
// leave resets them); every CTA arrives exactly once per barrier whether or not it had work in the phase.
// Co-residency: the grid never exceeds the SM count and a CTA takes a whole SM (≈ 200 KB of shared memory), so every CTA
// is resident before any can wait; waits are bounded (trap after ≈ 4 s) like the mbarrier waits.
#include <cuda.h>
#include <stdlib.h>

#include <algorithm>

#include "common.h"
#include "decode_reduce.cuh"
#include "gemm_common.cuh"
#include "ptx.cuh"

namespace teo {

constexpr int CH_MAX_PHASES = 4;
constexpr int CH_SYNC_WORDS = 2 * CH_MAX_PHASES + 1;          // two barriers per phase + the exit counter
enum ChainReduce { CH_RESID_NORM = 0, CH_SWIGLU = 1, CH_ROPE_KV = 2, CH_LOGITS = 3 };

struct ChainPhase {
    int N, K;                   // weight rows (UMMA M dimension), contraction length
    int q, work_ctas;           // stream-K: k-blocks per CTA, CTAs that own any
    int reduce;                 // ChainReduce
    float* partials;            // [slots][B][N]
    // CH_RESID_NORM: x[r] = bf16(Σ + x[r]); y[r] = RMSNorm(x[r])·norm_w      CH_SWIGLU: act = silu(g)·u
    // CH_ROPE_KV: q → qkv buffer, k/v → kv_pages                              CH_LOGITS: logits f32
    bf16* x;
    const bf16* norm_w;
    bf16* y;
    bf16* act;
    bf16* qkv;
    bf16* kv_pages;
    float* logits;
};
struct ChainArgs {
    int n_phases, B;
    ChainPhase ph[CH_MAX_PHASES];
    unsigned int* sync;         // CH_SYNC_WORDS counters
    const int* positions;       // seq_lens: position of the new token
    const int* block_table;
    int max_pages, n_heads, head_dim, page_size, inter, interleaved;
    const float* rope_cos;
    const float* rope_sin;
    float eps;
    int l2_prefetch;            // weight k-blocks (16 KiB each) per CTA the producer prefetches into L2 while it waits for a phase's input
    unsigned long long* trace;  // development only (teo_dbg_chain_trace): per CTA and phase 8 %globaltimer stamps, else nullptr
};
__device__ __forceinline__ void chain_stamp(const ChainArgs& a, int phase, int slot) {
    if (a.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        a.trace[(static_cast<size_t>(blockIdx.x) * CH_MAX_PHASES + phase) * 8 + slot] = t;
    }
}

template <int BN>
struct ChainCfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = (204800 / STAGE_BYTES) > 10 ? 10 : (204800 / STAGE_BYTES);     // 10 / 8 / 6 for BN = 32 / 64 / 128
    static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers + scratch*/;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }      // the 8 epilogue warps

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// one thread: wait until all `target` CTAs have arrived on the counter
__device__ __forceinline__ void grid_wait(const unsigned int* ctr, unsigned int target) {
    if (ld_acquire_gpu(ctr) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire_gpu(ctr) < target) {
        __nanosleep(32);
        if (clock64() - t0 > 8000000000LL) {
            printf("teochat_b200: decode chain grid barrier timed out (block %d, counter %u of %u)\n", blockIdx.x, ld_acquire_gpu(ctr), target);
            __trap();
        }
    }
}
// one thread, after the CTA's contributing threads have synchronised: publish this CTA's writes and arrive
__device__ __forceinline__ void grid_arrive(unsigned int* ctr) {
    __threadfence();
    atomicAdd(ctr, 1u);
}

template <int BN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
decode_chain_kernel(const __grid_constant__ CUtensorMap tw0, const __grid_constant__ CUtensorMap tw1, const __grid_constant__ CUtensorMap tw2,
                    const __grid_constant__ CUtensorMap tw3, const __grid_constant__ CUtensorMap ta0, const __grid_constant__ CUtensorMap ta1,
                    const __grid_constant__ CUtensorMap ta2, const __grid_constant__ CUtensorMap ta3, const ChainArgs a) {
    using Cfg = ChainCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;                                             // weight tiles  [STAGES][128][64]
    uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;                    // activations   [STAGES][BN][64]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* red_scratch = reinterpret_cast<float*>(tmem_slot + 2);       // 8 floats (row RMSNorm)

    pdl_trigger();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned int G = gridDim.x;
    const CUtensorMap* tw[CH_MAX_PHASES] = {&tw0, &tw1, &tw2, &tw3};
    const CUtensorMap* ta[CH_MAX_PHASES] = {&ta0, &ta1, &ta2, &ta3};

    if (warp == 0 && lane == 0) {
        for (int p = 0; p < a.n_phases; ++p) {
            tma_prefetch_desc(tw[p]);
            tma_prefetch_desc(ta[p]);
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], EPI_WARPS * 32);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this CTA's stream-K share of phase p: k-blocks [k0, k1) of the flattened (tile, k-block) space
    auto share = [&](int p, long long& k0, long long& k1, int& total_kb) {
        const ChainPhase& ph = a.ph[p];
        total_kb = ph.K / BK;
        const long long total = static_cast<long long>(ph.N / BM) * total_kb;
        k0 = min(total, static_cast<long long>(blockIdx.x) * ph.q);
        k1 = min(total, k0 + ph.q);
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int s = 0;
            uint32_t ph_bit = 0;
            for (int p = 0; p < a.n_phases; ++p) {
                long long k0, k1;
                int total_kb;
                share(p, k0, k1, total_kb);
                auto load_w = [&](int st, long long k) {
                    tma_load_4d(smem_a + st * A_STAGE_BYTES, tw[p], &full_bar[st], 0, 0, static_cast<int>(k % total_kb), static_cast<int>(k / total_kb));
                };
                auto load_act = [&](int st, long long k) {
                    tma_load_2d(smem_b + st * Cfg::B_STAGE_BYTES, ta[p], &full_bar[st], static_cast<int>(k % total_kb) * BK, 0);
                };
                // weights of the first stages before the phase's input exists (previous kernel / previous phase's reduction)
                const int s0 = s;
                int pre = 0;
                for (long long k = k0; k < k1 && pre < STAGES; ++k, ++pre) {
                    mbar_wait(&empty_bar[s], ph_bit ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                    load_w(s, k);
                    if (++s == STAGES) { s = 0; ph_bit ^= 1; }
                }
                chain_stamp(a, p, 0);            // ring filled with this phase's first weight tiles
                // HBM would idle until the phase's input exists (the grid barrier + the reduction before it): pull the next
                // weight k-blocks of this CTA's share into L2 meanwhile, so that the stream restarts at L2 speed
                if (p > 0 || a.l2_prefetch < 0) {
                    const int npf = a.l2_prefetch < 0 ? -a.l2_prefetch : a.l2_prefetch;
                    for (long long k = k0 + pre; k < k1 && k < k0 + pre + npf; ++k)
                        tma_prefetch_l2_4d(tw[p], 0, 0, static_cast<int>(k % total_kb), static_cast<int>(k / total_kb));
                }
                chain_stamp(a, p, 1);            // L2 prefetches issued
                if (p == 0) {
                    pdl_wait();
                } else {
                    grid_wait(&a.sync[2 * (p - 1) + 1], G);
                    asm volatile("fence.proxy.async;" ::: "memory");        // other CTAs' generic-proxy stores → this thread's TMA reads
                }
                chain_stamp(a, p, 2);            // input of the phase complete
                {
                    int st = s0;
                    for (int i = 0; i < pre; ++i) {
                        load_act(st, k0 + i);
                        if (++st == STAGES) st = 0;
                    }
                }
                for (long long k = k0 + pre; k < k1; ++k) {
                    mbar_wait(&empty_bar[s], ph_bit ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                    load_w(s, k);
                    load_act(s, k);
                    if (++s == STAGES) { s = 0; ph_bit ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            int s = 0, as = 0;
            uint32_t ph_bit = 0, aph = 0;
            for (int p = 0; p < a.n_phases; ++p) {
                long long k0, k1;
                int total_kb;
                share(p, k0, k1, total_kb);
                long long k = k0;
                while (k < k1) {
                    const long long tile_end = min(k1, (k / total_kb + 1) * total_kb);     // this item: k-blocks [k, tile_end) of one tile
                    mbar_wait(&tempty_bar[as], aph ^ 1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + as * BN;
                    for (long long kk = k; kk < tile_end; ++kk) {
                        mbar_wait(&full_bar[s], ph_bit);
                        tc_fence_after();
                        const uint64_t a_desc = umma_desc_k_sw128(smem_u32(smem_a + s * A_STAGE_BYTES));
                        const uint64_t b_desc = umma_desc_k_sw128(smem_u32(smem_b + s * Cfg::B_STAGE_BYTES));
#pragma unroll
                        for (int j = 0; j < BK / UMMA_K; ++j) umma_bf16(d_tmem, a_desc + 2 * j, b_desc + 2 * j, idesc, (kk > k || j > 0) ? 1u : 0u);
                        umma_commit(&empty_bar[s]);
                        if (kk == tile_end - 1) umma_commit(&tfull_bar[as]);
                        if (++s == STAGES) { s = 0; ph_bit ^= 1; }
                    }
                    if (++as == 2) { as = 0; aph ^= 1; }
                    k = tile_end;
                }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue + reductions
        pdl_wait();                              // partial workspace, x, activations belong to the previous kernels until now
        const int ew = warp - 4;
        const int qd = ew & 3;                   // TMEM lane quadrant of this warp
        const int hsel = ew >> 2;
        const int t256 = threadIdx.x - 128;
        int as = 0;
        uint32_t aph = 0;
        for (int p = 0; p < a.n_phases; ++p) {
            const ChainPhase& ph = a.ph[p];
            long long k0, k1;
            int total_kb;
            share(p, k0, k1, total_kb);
            const long long stride = static_cast<long long>(a.B) * ph.N;
            long long k = k0;
            while (k < k1) {
                const int tile = static_cast<int>(k / total_kb);
                const long long tile_end = min(k1, static_cast<long long>(tile + 1) * total_kb);
                const int slot = static_cast<int>(blockIdx.x) - static_cast<int>((static_cast<long long>(tile) * total_kb) / ph.q);
                mbar_wait(&tfull_bar[as], aph);
                tc_fence_after();
                const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(qd * 32) << 16) + as * BN;
                const int m = tile * BM + qd * 32 + lane;          // weight row = output column
                float* P = ph.partials + static_cast<long long>(slot) * stride;
#pragma unroll 1
                for (int c0 = hsel * 32; c0 < BN; c0 += 64) {
                    if (c0 >= a.B) break;                          // warp-uniform
                    uint32_t v[32];
                    tmem_ld_32x32(t_acc + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c0 + j < a.B) P[static_cast<long long>(c0 + j) * ph.N + m] = __uint_as_float(v[j]);
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[as]);
                if (++as == 2) { as = 0; aph ^= 1; }
                k = tile_end;
            }
            // ---- all partials of the phase complete, everywhere
            epi_bar_sync();
            if (t256 == 0) {
                chain_stamp(a, p, 3);            // this CTA's partials stored
                grid_arrive(&a.sync[2 * p]);
                grid_wait(&a.sync[2 * p], G);
                chain_stamp(a, p, 4);            // everyone's partials stored
            }
            epi_bar_sync();
            // ---- the phase's reduction, spread over the epilogue warps of the whole grid
            const PartialInfo pi{ph.partials, stride, total_kb, ph.q, ph.work_ctas};
            if (ph.reduce == CH_RESID_NORM) {
                for (int row = blockIdx.x; row < a.B; row += G)
                    reduce_residual_rmsnorm_row(pi, ph.x, ph.norm_w, ph.y, row, ph.N, a.eps, t256, red_scratch, [] { epi_bar_sync(); });
            } else if (ph.reduce == CH_SWIGLU) {
                reduce_swiglu_part(pi, ph.act, a.B, a.inter, a.interleaved, static_cast<long long>(blockIdx.x) * 256 + t256,
                                   static_cast<long long>(G) * 256);
            } else if (ph.reduce == CH_ROPE_KV) {
                for (int gw = blockIdx.x * EPI_WARPS + ew; gw < a.B * a.n_heads; gw += G * EPI_WARPS)
                    reduce_rope_kv_warp(pi, ph.qkv, a.positions, ph.kv_pages, a.block_table, a.max_pages, a.B, a.n_heads, a.head_dim,
                                        a.page_size, a.rope_cos, a.rope_sin, gw, lane);
            } else {
                reduce_logits_part(pi, ph.logits, a.B, ph.N, static_cast<long long>(blockIdx.x) * 256 + t256, static_cast<long long>(G) * 256);
            }
            // ---- the next phase's input complete, everywhere (the producer warps wait on this counter)
            if (p + 1 < a.n_phases) {
                asm volatile("fence.proxy.async;" ::: "memory");   // this thread's generic-proxy stores → visible to the TMA (async proxy) reads
                epi_bar_sync();
                if (t256 == 0) {
                    chain_stamp(a, p, 5);        // this CTA's part of the reduction done
                    grid_arrive(&a.sync[2 * p + 1]);
                }
            } else if (t256 == 0) {
                chain_stamp(a, p, 5);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    if (threadIdx.x == 0) {
        // the last CTA to leave zeroes the counters for the next launch (every CTA has passed every barrier by now)
        __threadfence();
        if (atomicAdd(&a.sync[CH_SYNC_WORDS - 1], 1u) == G - 1) {
            for (int i = 0; i < CH_SYNC_WORDS; ++i) a.sync[i] = 0u;
            __threadfence();
        }
    }
}

int get_tmap_wblocked(teo_handle* h, const void* ptr, uint64_t N, uint64_t K, uint32_t nblocks, CUtensorMap* out);

// development hooks (tools/chain_trace.py): a device buffer of `launches` × 148 × 4 × 8 u64 stamped by the next chain launches
// (used as a ring), and the L2-prefetch depth (k-blocks per CTA and phase; < 0: also before the first phase)
static unsigned long long* g_chain_trace = nullptr;
static int g_chain_trace_cap = 0;
static long long g_chain_trace_n = 0;
extern "C" long long teo_dbg_chain_trace(void* device_buffer, int launches) {
    const long long n = g_chain_trace_n;
    g_chain_trace = static_cast<unsigned long long*>(device_buffer);
    g_chain_trace_cap = device_buffer ? launches : 0;
    g_chain_trace_n = 0;
    return n;
}
static int g_chain_pf = [] {
    const char* e = getenv("TEO_CHAIN_PF");
    return e ? atoi(e) : 0;          // measured: 24 / 48 k-blocks of L2 prefetch per stall made the step 1.5 / 2.8 % slower
}();
extern "C" void teo_dbg_chain_prefetch(int kblocks) { g_chain_pf = kblocks; }

// Default OFF: measured on B200 (profiles/r02_decode_chain.txt) the chain is bit-identical to the per-GEMM sequence but 4–10 %
// SLOWER per decode step — the grid barriers wait for the slowest of 148 weight streams in every phase (finish-time skew of
// 12–44 µs per phase) and the in-kernel reductions run on 148 × 256 threads instead of the wide glue grids.  TEO_DEC_CHAIN=1 /
// teo_set_decode_chain(h, 1) select it.
bool decode_chain_enabled() {
    static const bool on = [] {
        const char* e = getenv("TEO_DEC_CHAIN");
        return e && e[0] == '1';
    }();
    return on;
}

// upper bound on fp32 partial slots of one 128-row tile under the chain's stream-K split over `grid` CTAs
static int chain_slots(int N, int K, int grid, int* q_out, int* work_out) {
    const int tiles = N / BM, total_kb = K / BK;
    const long long total = static_cast<long long>(tiles) * total_kb;
    const int q = static_cast<int>((total + grid - 1) / grid);
    if (q_out) *q_out = q;
    if (work_out) *work_out = static_cast<int>((total + q - 1) / q);
    return (total_kb + q - 2) / q + 1;
}

size_t decode_chain_workspace_bytes(int B, int N, int K, int grid) {
    return (static_cast<size_t>(chain_slots(N, K, grid, nullptr, nullptr)) * B * N * sizeof(float) + 255) & ~static_cast<size_t>(255);
}

bool decode_chain_shape_ok(int B, int hidden, int inter, int vocab, int n_heads) {
    const int hd = hidden / n_heads;
    return B >= 1 && B <= 128 && hidden % 128 == 0 && hidden % 64 == 0 && inter % 64 == 0 && (2 * inter) % 128 == 0 && vocab % 128 == 0 &&
           hidden <= 8192 && hidden % 32 == 0 && inter % 4 == 0 && hd % 4 == 0 && hd <= 256;
}

template <int BN>
static int launch_chain_bn(teo_handle* h, const CUtensorMap* tw, const CUtensorMap* ta, const ChainArgs& a, cudaStream_t stream) {
    using Cfg = ChainCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        TEO_CUDA(cudaFuncSetAttribute(decode_chain_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    TEO_CUDA(launch_kc(PDL_GEMM, decode_chain_kernel<BN>, dim3(h->num_sms), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, tw[0], tw[1], tw[2], tw[3],
                       ta[0], ta[1], ta[2], ta[3], a));
    TEO_LAUNCH_CHECK("decode_chain_kernel");
    h->launches++;
    return TEO_OK;
}

int launch_decode_chain(teo_handle* h, const ChainSpec* specs, int n_phases, int B, const int* positions, const int* block_table, int max_pages,
                        int n_heads, int head_dim, int page_size, int inter, int interleaved, const float* rope_cos, const float* rope_sin,
                        float eps, void* ws, size_t ws_bytes, cudaStream_t stream) {
    TEO_CHECK_ARG(h && specs && n_phases >= 1 && n_phases <= CH_MAX_PHASES && B >= 1 && B <= 128, "decode_chain: bad arguments");
    if (h->chain_sync == nullptr) {
        // CH_SYNC_WORDS barrier counters: the one piece of device memory the library owns (256 B per handle)
        TEO_CUDA(cudaMalloc(&h->chain_sync, 256));
        TEO_CUDA(cudaMemset(h->chain_sync, 0, 256));
    }
    const int bn = B <= 32 ? 32 : (B <= 64 ? 64 : 128);
    ChainArgs a{};
    a.n_phases = n_phases;
    a.B = B;
    a.sync = static_cast<unsigned int*>(h->chain_sync);
    a.positions = positions;
    a.block_table = block_table;
    a.max_pages = max_pages;
    a.n_heads = n_heads;
    a.head_dim = head_dim;
    a.page_size = page_size;
    a.inter = inter;
    a.interleaved = interleaved;
    a.rope_cos = rope_cos;
    a.rope_sin = rope_sin;
    a.eps = eps;
    a.l2_prefetch = g_chain_pf;
    a.trace = (g_chain_trace && g_chain_trace_cap > 0)
                  ? g_chain_trace + (g_chain_trace_n++ % g_chain_trace_cap) * (148ull * CH_MAX_PHASES * 8) : nullptr;
    CUtensorMap tw[CH_MAX_PHASES], ta[CH_MAX_PHASES];
    size_t off = 0;
    for (int p = 0; p < n_phases; ++p) {
        const ChainSpec& sp = specs[p];
        TEO_CHECK_ARG(sp.W && sp.A && sp.N % BM == 0 && sp.K % BK == 0 && sp.N > 0 && sp.K > 0, "decode_chain: phase %d has a bad shape (N=%d K=%d)", p,
                      sp.N, sp.K);
        ChainPhase& ph = a.ph[p];
        ph.N = sp.N;
        ph.K = sp.K;
        chain_slots(sp.N, sp.K, h->num_sms, &ph.q, &ph.work_ctas);
        const size_t need = decode_chain_workspace_bytes(B, sp.N, sp.K, h->num_sms);
        if (ws == nullptr || off + need > ws_bytes) {
            set_error("decode_chain: workspace too small (phase %d needs %zu at offset %zu of %zu)", p, need, off, ws_bytes);
            return TEO_ERR_WORKSPACE;
        }
        ph.partials = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + off);
        off += need;
        ph.reduce = sp.reduce;
        ph.x = sp.x;
        ph.norm_w = sp.norm_w;
        ph.y = sp.y;
        ph.act = sp.act;
        ph.qkv = sp.qkv;
        ph.kv_pages = sp.kv_pages;
        ph.logits = sp.logits;
        TEO_TRY(get_tmap_wblocked(h, sp.W, sp.N, sp.K, 1, &tw[p]));
        TEO_TRY(get_tmap_bf16(h, sp.A, B, sp.K, sp.lda, bn, &ta[p]));
    }
    for (int p = n_phases; p < CH_MAX_PHASES; ++p) {
        tw[p] = tw[0];
        ta[p] = ta[0];
    }
    switch (bn) {
        case 32: return launch_chain_bn<32>(h, tw, ta, a, stream);
        case 64: return launch_chain_bn<64>(h, tw, ta, a, stream);
        default: return launch_chain_bn<128>(h, tw, ta, a, stream);
    }
}

}  // namespace teo
