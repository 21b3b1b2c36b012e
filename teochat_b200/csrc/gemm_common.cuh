// Pieces shared by the two tcgen05 GEMM kernels — the single-CTA kernel (gemm.cu: every schedule) and the CTA-pair kernel
// (gemm_pair.cu: cta_group::2, 256-row tiles, the large tiled GEMMs of prefill and the ViT): tile constants, the kernel
// argument block and the staged TMA-store epilogue.
#pragma once
#include <cuda.h>

#include "common.h"
#include "ptx.cuh"

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
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING + 1024 /*align slack*/ + 256 /*barriers*/;
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

// Staged epilogue of one 128 × BN accumulator tile (bf16 row-major output): this warp's 32 TMEM lanes × its 64-column
// chunks → bias / activation / residual (residual tile fetched by TMA into the staging buffer) → bf16 → 128-byte-swizzled
// staging → TMA store.  `release()` is called exactly once, right after this warp's last TMEM read of the tile (it frees
// the accumulator stage for the MMA warp).  Shared by the single-CTA kernel (gemm.cu) and the CTA-pair kernel
// (gemm_pair.cu), whose CTAs each own 128 rows of a 256-row tile.
template <int BN, typename Release>
__device__ __forceinline__ void staged_epilogue_tile(const GemmArgs& g, const CUtensorMap* tma_c, const CUtensorMap* tma_r, uint32_t t_acc,
                                                     int m_blk, int n_blk, uint8_t* stg_base, uint64_t* rbar, uint32_t& rph, int lane, int q,
                                                     int hsel, Release release) {
    // rph carries two counters: bit 0 = phase of the residual barrier, bits 1.. = chunks staged so far (selects the staging buffer)
    auto next_buf = [&]() -> uint8_t* {
        uint8_t* b = stg_base + ((rph >> 1) % EPI_BUFS) * 4096;
        rph += 2;
        return b;
    };
    if (g.act == TEO_ACT_SWIGLU_PAIRS) {
        // ---- SwiGLU fused into the gate/up projection: the weight rows come interleaved in blocks of 32
        // (…| gate 32 | up 32 |…), so accumulator columns [c, c+32) and [c+32, c+64) are the gate and the up
        // projection of the SAME 32 outputs.  A warp turns two such 64-column chunks into one 64-column
        // output chunk — bf16(silu(bf16(g)) · bf16(u)), the rounding points of the unfused chain — and
        // stores it with the usual TMA box; C has N/2 columns.
        const int row0 = m_blk * BM + q * 32;
#pragma unroll 1
        for (int oc = hsel; oc < BN / 128; oc += 2) {
            const int n_out0 = n_blk * (BN / 2) + oc * 64;
            if (2 * n_out0 >= g.N) break;              // warp-uniform
            uint8_t* stg = next_buf();
            if (lane == 0) tma_store_wait_read<EPI_BUFS - 1>();   // the store that last used this staging buffer has drained it
            __syncwarp();
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {
                const int col = oc * 128 + hc * 64;
                uint32_t vg[32], vu[32];
                if (n_blk * BN + col < g.N) {
                    tmem_ld_32x32(t_acc + col, vg);
                    tmem_ld_32x32(t_acc + col + 32, vu);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) vg[j] = vu[j] = 0u;
                }
                if (hc == 1 && (oc + 2 >= BN / 128 || 2 * (n_out0 + 128) >= g.N)) {   // last chunk of this warp
                    release();
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float x[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float gg = __bfloat162float(__float2bfloat16_rn(__uint_as_float(vg[c * 8 + j])));
                        const float uu = __bfloat162float(__float2bfloat16_rn(__uint_as_float(vu[c * 8 + j])));
                        x[j] = (gg / (1.0f + expf(-gg))) * uu;
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
        if (hsel >= BN / 128 || 2 * (n_blk * (BN / 2) + hsel * 64) >= g.N) {   // this warp owned no chunk of the tile
            release();
        }
    } else {
        // ---- staged path: 64-column chunks → swizzled smem → TMA store
        const int row0 = m_blk * BM + q * 32;
        const bool has_res = g.residual != nullptr;
        // folded LayerNorm: this thread's row statistics (fixed summation order over the producer's slots → deterministic)
        float ln_mu = 0.f, ln_rstd = 1.f;
        if (g.ln_stats != nullptr && row0 + lane < g.M) {
            const float* st = g.ln_stats + static_cast<long long>(row0 + lane) * g.ln_slots * 2;
            float s1 = 0.f, s2 = 0.f;
            for (int i = 0; i < g.ln_slots; ++i) { s1 += st[2 * i]; s2 += st[2 * i + 1]; }
            ln_mu = s1 * g.ln_inv_d;
            ln_rstd = rsqrtf(fmaxf(s2 * g.ln_inv_d - ln_mu * ln_mu, 0.f) + g.ln_eps);
        }
        float st1 = 0.f, st2 = 0.f;            // Σ, Σ² of the bf16 values this warp writes for its row
#pragma unroll 1
        for (int cj = hsel; cj < BN / 64; cj += 2) {
            const int n0 = n_blk * BN + cj * 64;
            if (n0 >= g.N) break;                      // warp-uniform
            uint8_t* stg = next_buf();
            if (lane == 0) {
                tma_store_wait_read<EPI_BUFS - 1>();   // the store that last used this staging buffer has drained it
                if (has_res) {
                    mbar_arrive_expect_tx(rbar, 4096);
                    tma_load_2d(stg, tma_r, rbar, n0, row0);
                }
            }
            __syncwarp();
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
            tmem_ld_32x32(t_acc + cj * 64, v0);
            tmem_ld_32x32(t_acc + cj * 64 + 32, v1);
            tmem_ld_wait();
            if (cj + 2 >= BN / 64 || n0 + 128 >= g.N) {   // last chunk of this warp: accumulator stage is free
                release();
            }
            if (has_res) {
                mbar_wait(rbar, rph & 1);
                rph ^= 1;
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) {              // eight 16-byte groups of 8 columns
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
            fence_proxy_async();                       // generic-proxy writes → visible to the TMA engine
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(tma_c, stg, n0, row0);
                tma_store_commit();
            }
        }
        if (hsel >= BN / 64 || n_blk * BN + hsel * 64 >= g.N) {   // this warp owned no chunk of the tile
            release();
        }
        if (g.stats_out != nullptr && row0 + lane < g.M) {         // (zeros from a warp that owned no chunk: the consumer sums all slots)
            float* so = g.stats_out + (static_cast<long long>(row0 + lane) * g.stats_slots + n_blk * 2 + hsel) * 2;
            so[0] = st1;
            so[1] = st2;
        }
    }
}

}  // namespace teo
