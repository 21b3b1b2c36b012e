// Pieces shared by the two tcgen05 GEMM kernels — the single-CTA kernel (gemm.cu: every schedule) and the CTA-pair kernel
// (gemm_pair.cu: cta_group::2, 256-row tiles, the large tiled GEMMs of prefill and the ViT): tile constants, the kernel
// argument block and the staged TMA-store epilogue.
#pragma once
#include <cuda.h>

#include <type_traits>

#include "common.h"
#include "ptx.cuh"

#ifndef EPI_STAMP        // gemm_pair.cu's development build (TEO_PAIR_TRACE) stamps clock64 at these points of a tile's first chunk
#define EPI_STAMP(slot) do { } while (0)
#endif

namespace teo {

constexpr int BM = 128;         // UMMA M
constexpr int BK = 64;          // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_STAGE_BYTES = BM * BK * 2;   // 16 KiB
constexpr int GEMM_THREADS = 384;     // 4 control warps + 8 epilogue warps
constexpr int EPI_WARPS = 8;
// Epilogue staging: per warp EPI_BUFS buffers of 32 rows × 64 bf16 (128 B rows, swizzled).  With two, a chunk is converted and staged
// while the previous chunk's TMA store is still reading its buffer (cp.async.bulk.wait_group.read 1 instead of 0): the epilogue of a tile
// no longer pays one store-drain latency per chunk — what made the un-overlapped epilogue of the 512-row pair tiles expensive.
#ifndef TEO_EPI_BUFS
#define TEO_EPI_BUFS 2
#endif
constexpr int EPI_BUFS = TEO_EPI_BUFS;
constexpr int STAGING_BYTES = EPI_WARPS * 4096 * EPI_BUFS;
constexpr int RES_BARS = EPI_WARPS * EPI_BUFS;       // one "residual tile landed" barrier per staging buffer
constexpr int GEMM_BAR_BYTES = 512;                  // barrier area at the end of shared memory (≤ 16 + 4 + RES_BARS barriers + the TMEM slot)
constexpr int GROUP_M = 16;     // rasterisation group (tiles along M sharing W tiles in L2)

struct GemmArgs {
    int M, N, K;                // GEMM-space sizes (swap-AB: M = weight rows, N = batch rows)
    void* C;
    long long ldc;
    const bf16* bias;
    const bf16* residual;
    long long ldr;
    int act;
    int out_fp32;
    int transposed;             // store C[n*ldc + m], bias indexed by m (swap-AB)
    int streamk;                // small-M schedule: every CTA takes an equal contiguous share of the flattened
                                // (tile, k-block) space; fp32 partials per (tile, slot), no bias/act/residual here
    int sk_q;                   // k-blocks per CTA in that schedule
    long long split_stride;     // elements between partial slots
    int tma_epi;                // bf16 row-major output through the staged TMA-store epilogue
    int w_is_a;                 // operand A holds the (constant) weights: may be fetched before griddepcontrol.wait
    int w_blocked;              // weights stored tile-blocked [N/128][K/64][128][64] (4-D tensor map)
    int producers;              // stream-K schedule: 2 = a second TMA producer warp shares the k-blocks (gemm.cu)
    int k_wrap;                 // > 0: the WEIGHT operand has only k_wrap k-blocks and k-block kb reads kb % k_wrap — the
                                // activation operand then holds K/k_wrap/64 bf16 planes side by side (exact mode: an fp32
                                // activation split into hi | mid | lo bf16 terms, all multiplied by the same weights)
    int residual_f32;           // residual is float (direct-store epilogue only)
    int group_n, raster, l2_hint;   // CTA-pair kernel: supertile width (tile columns), supertile order, L2 eviction hints (gemm_pair.cu)
    // LayerNorm folded into the GEMM that consumes it (staged epilogue only; see GemmEpilogue in common.h):
    //   C[m,n] = act( rstd_m · (acc[m,n] − mean_m · ln_c[n]) + ln_bias[n] ),  mean / rstd from the row statistics ln_stats
    const float* ln_stats;      // f32 [M][ln_slots][2]: partial (Σx, Σx²) of the INPUT rows
    const float* ln_c;          // f32 [N]: Σ_k W'[n,k]
    const float* ln_bias;       // f32 [N]: Σ_k β_k·W[n,k] + b[n]
    int ln_slots;
    float ln_inv_d, ln_eps;     // 1 / K of the normalised dimension, epsilon
    // … and the row statistics of the OUTPUT (the residual stream this GEMM writes) for the next folded LayerNorm:
    float* stats_out;           // f32 [M][stats_slots][2]; slot = 2·(column tile) + (column-chunk parity of the writing warp)
    int stats_slots;
    // stream-K schedule: reduction of the partials + SwiGLU inside the GEMM (gemm.cu, launch_gemm_partials with an SkFuse)
    int fuse;                   // 0 | 1: the CTA holding slot 0 of a weight tile reduces the tile's partials once the other slots have arrived
    int* sk_flags;              // [tiles] arrival counters, zero between launches (reset by the reducing CTA)
    bf16* fuse_out;             // act [batch, fuse_inter]
    int fuse_inter, fuse_interleaved;
    unsigned long long* trace;  // development only (teo_dbg_gemm_trace): per CTA 8 %globaltimer stamps, else nullptr
};

// SK = the small-M stream-K schedule (decode).  Its partial outputs go straight from registers to global memory, so the epilogue
// staging buffer could go to the ring instead (TEO_SK_DEEP: 11 / 9 / 7 stages for BN = 32 / 64 / 128 instead of 8 / 8 / 6).  Built and
// A/B-measured on one box, alternating runs (build.build_variant("skdeep", ["TEO_SK_DEEP"]), scripts/gpu_r02_ring.sh): the deeper
// ring is 1.0 % SLOWER on the decode step at bs=32 (2401 vs 2378 ms per 255 steps, 3 runs each), 2.3 % at bs=2 — more weight
// requests in flight per SM do not raise the 5.1 TB/s the stream reaches (profiles/r02_dec_gemm_skew.txt).  Default: round-1 depth.
template <int BN, bool SK = false>
struct GemmCfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
#ifdef TEO_SK_DEEP
    static constexpr int STAGING = SK ? 0 : STAGING_BYTES;
    static constexpr int BUDGET = SK ? 229376 : 196608;
    static constexpr int MAX_STAGES = SK ? 11 : 8;
#else
    static constexpr int STAGING = SK ? 0 : STAGING_BYTES;                 // (the stream-K epilogue stores straight from registers)
    static constexpr int BUDGET = SK ? 196608 : 229376 - STAGING_BYTES;    // ring: 8 / 8 / 6 stages (SK), 8 / 6 / 5 / 3 (tiled, BN 32 … 256)
    static constexpr int MAX_STAGES = 8;
#endif
    static constexpr int STAGES = (BUDGET / STAGE_BYTES) > MAX_STAGES ? MAX_STAGES : (BUDGET / STAGE_BYTES);
    static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;   // power of two for BN ∈ {32,64,128,256}
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING + 1024 /*align slack*/ + GEMM_BAR_BYTES;
};

__device__ __forceinline__ void trace_stamp(const GemmArgs& g, int slot) {
    if (g.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        g.trace[blockIdx.x * 8 + slot] = t;
    }
}

__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == TEO_ACT_QUICK_GELU) {
        // x·σ(1.702x) = x·(½ + ½·tanh(0.851x)): one MUFU op (tanh.approx, rel. error 2^-11 ≪ the bf16 rounding that follows)
        // instead of ex2 + rcp — the epilogue of the ViT fc1 GEMM (K = 1024) is otherwise MUFU-bound
        float t;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
        return x * fmaf(0.5f, t, 0.5f);
    }
    if (act == TEO_ACT_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
    return x;
}

// One 64-column chunk: 2 × 32 accumulator columns of this thread's row → (+ bias) → activation → (+ residual, read from the
// staging buffer the TMA load put it in) → bf16 → 128-byte-swizzled staging row.  Compile-time specialised: written as ONE loop
// with run-time tests of g.bias / g.act / g.ln_stats / g.stats_out / g.residual the chunk body compiled to ≈ 2 900 SASS
// instructions (64 inlined erff expansions, the folded-LayerNorm loads, the statistics, …), 46 KB of code that eight warps walked
// through taking branches around almost all of it — instruction fetch, not arithmetic, made a chunk cost ≈ 2.2 k cycles
// (clock64 stamps inside the epilogue, profiles/r02_pair_epilogue.txt).  Each specialisation is ≈ 200 instructions.
template <int ACT, bool HAS_BIAS, bool HAS_RES>
__device__ __forceinline__ void convert_chunk(const uint32_t (&v0)[32], const uint32_t (&v1)[32], const uint4 (&bvec)[8], uint8_t* stg,
                                              int lane) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {              // eight 16-byte groups of 8 columns
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(c < 4 ? v0[c * 8 + j] : v1[(c - 4) * 8 + j]);
        if constexpr (HAS_BIAS) {              // (columns past N carry zeros in bvec; what is computed for them is clipped by the TMA store)
            const uint32_t bw[4] = {bvec[c].x, bvec[c].y, bvec[c].z, bvec[c].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { x[2 * j] += bf16_lo(bw[j]); x[2 * j + 1] += bf16_hi(bw[j]); }
        }
        if constexpr (ACT != TEO_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = apply_act(x[j], ACT);
        }
        uint4* slot = reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4));
        if constexpr (HAS_RES) {
            const uint4 rv = *slot;
            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { x[2 * j] += bf16_lo(rw[j]); x[2 * j + 1] += bf16_hi(rw[j]); }
        }
        *slot = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
    }
}
template <int ACT>
__device__ __forceinline__ void convert_chunk_act(bool has_bias, bool has_res, const uint32_t (&v0)[32], const uint32_t (&v1)[32],
                                                  const uint4 (&bvec)[8], uint8_t* stg, int lane) {
    if (has_bias) {
        if (has_res) convert_chunk<ACT, true, true>(v0, v1, bvec, stg, lane);
        else convert_chunk<ACT, true, false>(v0, v1, bvec, stg, lane);
    } else {
        if (has_res) convert_chunk<ACT, false, true>(v0, v1, bvec, stg, lane);
        else convert_chunk<ACT, false, false>(v0, v1, bvec, stg, lane);
    }
}

// Staged epilogue of the MT row sub-tiles (128 × BN accumulators each) of one output tile, bf16 row-major output: this warp's 32 TMEM
// lanes × its 64-column chunks → bias / activation / residual → bf16 → 128-byte-swizzled staging → TMA store.  Shared by the
// single-CTA kernel (gemm.cu, MT = 1) and the CTA-pair kernel (gemm_pair.cu, whose CTAs each own 128 rows of every 256-row sub-tile).
//   t_acc + sub·acc_stride   TMEM address of sub-tile `sub` (lane quadrant included);  its rows are block m_blk0 + sub·m_stride
//   stg_base / rbars         this warp's EPI_BUFS staging buffers (4 KiB each) and their "residual landed" barriers
//   nchunk                   chunks this warp has staged since the kernel started: selects buffer and barrier phase
//   wait_full()              waits for the accumulators (called once, AFTER the first residual chunk has been requested)
//   release()                called exactly once, right after this warp's last TMEM read of the tile
// The residual tile of a chunk is fetched by TMA into the staging buffer the chunk is then converted in.  With two buffers the
// request for chunk i+1 is issued while chunk i is read from TMEM and converted, and the first chunk of a tile is requested before
// the accumulators are even waited for — the L2 round trip of the residual used to be paid once per chunk, un-overlapped, which
// is what made the epilogue of the 512-row pair tiles (not overlapped by a second accumulator stage) cost ≈ 12 k cycles per tile.
template <int BN, int MT, typename WaitFull, typename Release>
__device__ __forceinline__ void staged_epilogue(const GemmArgs& g, const CUtensorMap* tma_c, const CUtensorMap* tma_r, uint32_t t_acc,
                                                uint32_t acc_stride, int m_blk0, int m_stride, int n_blk, uint8_t* stg_base, uint64_t* rbars,
                                                uint32_t& nchunk, int lane, int q, int hsel, WaitFull wait_full, Release release) {
    if (g.act == TEO_ACT_SWIGLU_PAIRS) {
        // ---- SwiGLU fused into the gate/up projection: the weight rows come interleaved in blocks of 32
        // (…| gate 32 | up 32 |…), so accumulator columns [c, c+32) and [c+32, c+64) are the gate and the up
        // projection of the SAME 32 outputs.  A warp turns two such 64-column chunks into one 64-column
        // output chunk — bf16(silu(bf16(g)) · bf16(u)), the rounding points of the unfused chain — and
        // stores it with the usual TMA box; C has N/2 columns.
        constexpr int CPW = (BN / 128 + 1) / 2;                       // output chunks per warp and sub-tile
        int cv = 0;
#pragma unroll
        for (int i = 0; i < CPW; ++i) cv += (2 * (n_blk * (BN / 2) + (hsel + 2 * i) * 64) < g.N && hsel + 2 * i < BN / 128) ? 1 : 0;
        const int m = MT * cv;
        wait_full();
        if (m == 0) release();                                         // this warp owned no chunk of the tile
#pragma unroll 1
        for (int i = 0; i < m; ++i) {
            const int sub = i / cv, oc = hsel + 2 * (i - sub * cv);
            const int row0 = (m_blk0 + sub * m_stride) * BM + q * 32;
            const int n_out0 = n_blk * (BN / 2) + oc * 64;
            uint8_t* stg = stg_base + (nchunk % EPI_BUFS) * 4096;
            ++nchunk;
            if (lane == 0) tma_store_wait_read<EPI_BUFS - 1>();   // the store that last used this staging buffer has drained it
            __syncwarp();
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {
                const int col = oc * 128 + hc * 64;
                uint32_t vg[32], vu[32];
                if (n_blk * BN + col < g.N) {
                    tmem_ld_32x32(t_acc + sub * acc_stride + col, vg);
                    tmem_ld_32x32(t_acc + sub * acc_stride + col + 32, vu);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) vg[j] = vu[j] = 0u;
                }
                if (hc == 1 && i == m - 1) release();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float x[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float gg = __bfloat162float(__float2bfloat16_rn(__uint_as_float(vg[c * 8 + j])));
                        const float uu = __bfloat162float(__float2bfloat16_rn(__uint_as_float(vu[c * 8 + j])));
                        x[j] = silu_mul_fast(gg, uu);
                    }
                    const int cc = hc * 4 + c;
                    *reinterpret_cast<uint4*>(stg + lane * 128 + ((cc ^ (lane & 7)) << 4)) =
                        make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
                }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(tma_c, stg, n_out0, row0);
                tma_store_commit();
            }
        }
        return;
    }
    // ---- staged path: 64-column chunks → swizzled smem → TMA store
    constexpr int CPW = (BN / 64 + 1) / 2;                            // chunks per warp and sub-tile
    constexpr bool PIPE = EPI_BUFS >= 2;                               // residual of chunk i+1 requested while chunk i is converted
    const bool has_res = g.residual != nullptr;
    const bool generic = g.ln_stats != nullptr || g.stats_out != nullptr;      // folded LayerNorm / row statistics: the one-loop-does-all path
    int cv = 0;
#pragma unroll
    for (int i = 0; i < CPW; ++i) cv += (n_blk * BN + (hsel + 2 * i) * 64 < g.N && hsel + 2 * i < BN / 64) ? 1 : 0;
    const int m = MT * cv;
    auto chunk_at = [&](int i, int& sub, int& n0, int& row0) {
        sub = i / cv;
        n0 = n_blk * BN + (hsel + 2 * (i - sub * cv)) * 64;
        row0 = (m_blk0 + sub * m_stride) * BM + q * 32;
    };
    // lane 0: request the residual tile of chunk i (staging buffer / barrier of chunk number `no`); `pending` = TMA stores that may
    // still be reading their buffers — the one that last used THIS buffer must not be among them
    auto request_res = [&](int i, uint32_t no, auto pending) {
        int sub, n0, row0;
        chunk_at(i, sub, n0, row0);
        tma_store_wait_read<decltype(pending)::value>();
        mbar_arrive_expect_tx(&rbars[no % EPI_BUFS], 4096);
        tma_load_2d(stg_base + (no % EPI_BUFS) * 4096, tma_r, &rbars[no % EPI_BUFS], n0, row0);
    };
    if (has_res && m > 0 && lane == 0) request_res(0, nchunk, std::integral_constant<int, EPI_BUFS - 1>{});
    wait_full();
    if (m == 0) release();                                             // this warp owned no chunk of the tile
    float ln_mu = 0.f, ln_rstd = 1.f;
    float st1 = 0.f, st2 = 0.f;            // Σ, Σ² of the bf16 values this warp writes for its row (current sub-tile)
#pragma unroll 1
    for (int i = 0; i < m; ++i) {
        int sub, n0, row0;
        chunk_at(i, sub, n0, row0);
        const bool sub_first = (i - sub * cv) == 0, sub_last = (i - sub * cv) == cv - 1;
        const uint32_t no = nchunk++;
        uint8_t* stg = stg_base + (no % EPI_BUFS) * 4096;
        if (sub_first && g.ln_stats != nullptr) {
            // folded LayerNorm: this thread's row statistics (fixed summation order over the producer's slots → deterministic)
            ln_mu = 0.f; ln_rstd = 1.f;
            if (row0 + lane < g.M) {
                const float* st = g.ln_stats + static_cast<long long>(row0 + lane) * g.ln_slots * 2;
                float s1 = 0.f, s2 = 0.f;
                for (int k = 0; k < g.ln_slots; ++k) { s1 += st[2 * k]; s2 += st[2 * k + 1]; }
                ln_mu = s1 * g.ln_inv_d;
                ln_rstd = rsqrtf(fmaxf(s2 * g.ln_inv_d - ln_mu * ln_mu, 0.f) + g.ln_eps);
            }
        }
        if (i == 0) EPI_STAMP(10);
        if (lane == 0) {
            if (!has_res) {
                tma_store_wait_read<EPI_BUFS - 1>();   // the store that last used this staging buffer has drained it
            } else if (!PIPE && i > 0) {
                request_res(i, no, std::integral_constant<int, EPI_BUFS - 1>{});
            }
        }
        __syncwarp();
        if (i == 0) EPI_STAMP(11);
        // The chunk's 64 bias values first (8 × 16 bytes, the same addresses in every lane): issued together, their L1 / L2
        // latency overlaps the TMEM read and the residual tile's arrival.  Loaded one group at a time inside the loop below they
        // were eight dependent round trips per chunk — on the K = 1024 GEMMs of the ViT (short main loop, every linear has a
        // bias) that made the epilogue, not the tensor pipe, the critical path.
        uint4 bvec[8];
        if (g.bias != nullptr && g.ln_stats == nullptr) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                bvec[c] = (n0 + c * 8 < g.N) ? *reinterpret_cast<const uint4*>(g.bias + n0 + c * 8) : make_uint4(0u, 0u, 0u, 0u);
        }
        uint32_t v0[32], v1[32];
        const uint32_t t_chunk = t_acc + sub * acc_stride + static_cast<uint32_t>(n0 - n_blk * BN);
        tmem_ld_32x32(t_chunk, v0);
        tmem_ld_32x32(t_chunk + 32, v1);
        if (PIPE && has_res && i + 1 < m && lane == 0)      // the next chunk's residual: its buffer was last read by the store of chunk no − 1
            request_res(i + 1, no + 1, std::integral_constant<int, (EPI_BUFS >= 2 ? EPI_BUFS - 2 : 0)>{});
        __syncwarp();
        tmem_ld_wait();
        if (i == 0) EPI_STAMP(12);
        if (i == m - 1) release();                          // last chunk of this warp: the accumulators are free
        if (has_res) mbar_wait(&rbars[no % EPI_BUFS], (no / EPI_BUFS) & 1);
        if (i == 0) EPI_STAMP(13);
        if (!generic) {
            // the specialisations (warp-uniform dispatch, once per chunk)
            const bool hb = g.bias != nullptr;
            if (g.act == TEO_ACT_NONE) convert_chunk_act<TEO_ACT_NONE>(hb, has_res, v0, v1, bvec, stg, lane);
            else if (g.act == TEO_ACT_QUICK_GELU) convert_chunk_act<TEO_ACT_QUICK_GELU>(hb, has_res, v0, v1, bvec, stg, lane);
            else convert_chunk_act<TEO_ACT_GELU>(hb, has_res, v0, v1, bvec, stg, lane);
        } else
#pragma unroll
        for (int c = 0; c < 8; ++c) {              // generic path (folded LayerNorm / row statistics): eight 16-byte groups of 8 columns
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(c < 4 ? v0[c * 8 + j] : v1[(c - 4) * 8 + j]);
            const int n = n0 + c * 8;
            if (g.ln_stats != nullptr && n < g.N) {
                const float4 c0 = *reinterpret_cast<const float4*>(g.ln_c + n), c1 = *reinterpret_cast<const float4*>(g.ln_c + n + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(g.ln_bias + n), b1 = *reinterpret_cast<const float4*>(g.ln_bias + n + 4);
                const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = fmaf(ln_rstd, x[j] - ln_mu * cc[j], bb[j]);
            } else if (g.bias && n < g.N) {
                const uint4 bv = bvec[c];
                const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { x[2 * j] += bf16_lo(bw[j]); x[2 * j + 1] += bf16_hi(bw[j]); }
            }
            if (g.act != TEO_ACT_NONE) {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = apply_act(x[j], g.act);
            }
            uint4* slot = reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4));
            if (has_res) {
                const uint4 rv = *slot;
                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) { x[2 * j] += bf16_lo(rw[j]); x[2 * j + 1] += bf16_hi(rw[j]); }
            }
            const uint4 packed = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                                            pack_bf16x2(x[6], x[7]));
            *slot = packed;
            if (g.stats_out != nullptr) {          // statistics of the values AS STORED (bf16), columns past N hold zeros
                const uint32_t pw[4] = {packed.x, packed.y, packed.z, packed.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float lo = bf16_lo(pw[j]), hi = bf16_hi(pw[j]);
                    st1 += lo + hi;
                    st2 = fmaf(lo, lo, fmaf(hi, hi, st2));
                }
            }
        }
        if (i == 0) EPI_STAMP(14);
        fence_proxy_async();                       // generic-proxy writes → visible to the TMA engine
        __syncwarp();
        if (i == 0) EPI_STAMP(15);
        if (lane == 0) {
            tma_store_2d(tma_c, stg, n0, row0);
            tma_store_commit();
        }
        if (sub_last && g.stats_out != nullptr) {
            if (row0 + lane < g.M) {
                float* so = g.stats_out + (static_cast<long long>(row0 + lane) * g.stats_slots + n_blk * 2 + hsel) * 2;
                so[0] = st1;
                so[1] = st2;
            }
            st1 = st2 = 0.f;
        }
    }
    if (m == 0 && g.stats_out != nullptr) {        // zeros from a warp that owned no chunk: the consumer sums all slots
#pragma unroll
        for (int sub = 0; sub < MT; ++sub) {
            const int row0 = (m_blk0 + sub * m_stride) * BM + q * 32;
            if (row0 + lane < g.M) {
                float* so = g.stats_out + (static_cast<long long>(row0 + lane) * g.stats_slots + n_blk * 2 + hsel) * 2;
                so[0] = 0.f;
                so[1] = 0.f;
            }
        }
    }
}

}  // namespace teo
