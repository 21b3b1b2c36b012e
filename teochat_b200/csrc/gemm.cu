// bf16 × bf16 → fp32 GEMM on the 5th-gen tensor cores: C[M,N] = epi(A[M,K] · W[N,K]^T).
//
// One persistent, warp-specialised kernel:
//   warp 0      TMA producer   cp.async.bulk.tensor 2-D tiles of A and W (both K-major, 128-byte
//                              swizzle) into a STAGES-deep shared-memory ring, mbarrier-signalled
//   warp 1      MMA issuer     one thread issues tcgen05.mma (UMMA 128×BN×16, cta_group::1),
//                              accumulating in TMEM; tcgen05.commit releases ring slots
//   warp 2      TMEM allocator (2 accumulator stages × BN fp32 columns)
//   warps 4-11  epilogue       two warps per TMEM lane quadrant (column halves).  bf16 outputs:
//                              tcgen05.ld 32 lanes × 64 columns → bias / activation / residual
//                              (residual tile fetched by TMA into the staging buffer) → bf16 →
//                              128-byte-swizzled staging in shared memory → TMA store (full
//                              128-byte lines, edges clipped by the tensor map); overlaps the next
//                              tile's MMAs through the second TMEM accumulator stage.
//                              fp32 / split-K partial / transposed outputs use direct stores.
//
// Schedules:
//   normal      A = activations [M,K], W = weights [N,K]; tiles 128 × 256
//   swap-AB     small M (decode: M = batch ≤ 128): the weight matrix takes the UMMA M dimension
//               (128 rows of W per tile), the activations the N dimension (BN = 32/64/128), the
//               epilogue stores transposed.  Weight streaming is HBM-bound, so tiles are also
//               split along K to put ≥ 2 work units on every SM; fp32 partials go to the
//               workspace and a small kernel reduces them in a fixed order (deterministic).
//
// Algorithmic work: 2·M·N·K flop; bytes 2·(M·K + N·K + M·N).
#include <cuda.h>
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>

#include "common.h"
#include "decode_reduce.cuh"
#include "gemm_common.cuh"
#include "ptx.cuh"

namespace teo {

// Work items of one CTA.  Normal schedule: whole tiles, strided over the persistent grid, rasterised in groups of
// GROUP_M tiles along M.  Stream-K schedule (small M, weight streaming): CTA c owns k-blocks [c·Q, (c+1)·Q) of the
// flattened (tile, k-block) space, so every SM streams the same number of weight bytes (no wave quantisation);
// a tile touched by several CTAs gets one fp32 partial per CTA, slot = c − first CTA of the tile.
struct WorkItem {
    int m_blk, n_blk, kb0, kb1, slot;
};
struct Scheduler {
    int num_m, num_n, total_kb;
    int unit, num_units;          // normal
    long long k, k_end;           // stream-K
    __device__ __forceinline__ void init(const GemmArgs& g, int num_m_, int num_n_, int total_kb_) {
        num_m = num_m_; num_n = num_n_; total_kb = total_kb_;
        unit = blockIdx.x;
        num_units = num_m * num_n;
        if (g.streamk) {
            const long long total = static_cast<long long>(num_units) * total_kb;
            k = static_cast<long long>(blockIdx.x) * g.sk_q;
            k_end = min(total, k + g.sk_q);
        }
    }
    __device__ __forceinline__ bool next(const GemmArgs& g, WorkItem& it) {
        if (g.streamk) {
            if (k >= k_end) return false;
            const int tile = static_cast<int>(k / total_kb);
            it.kb0 = static_cast<int>(k - static_cast<long long>(tile) * total_kb);
            it.kb1 = static_cast<int>(min(static_cast<long long>(total_kb), it.kb0 + (k_end - k)));
            it.m_blk = tile / num_n;
            it.n_blk = tile % num_n;
            it.slot = static_cast<int>(blockIdx.x) - static_cast<int>((static_cast<long long>(tile) * total_kb) / g.sk_q);
            k += it.kb1 - it.kb0;
            return true;
        }
        if (unit >= num_units) return false;
        const int group_sz = GROUP_M * num_n;
        const int grp = unit / group_sz;
        const int first_m = grp * GROUP_M;
        const int gm = min(GROUP_M, num_m - first_m);
        const int in_grp = unit - grp * group_sz;
        it.m_blk = first_m + in_grp % gm;
        it.n_blk = in_grp / gm;
        it.kb0 = 0;
        it.kb1 = total_kb;
        it.slot = 0;
        unit += gridDim.x;
        return true;
    }
};

__device__ __forceinline__ void named_bar_sync_epi() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory"); }

template <int BN, bool SK>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_r, const GemmArgs g) {
    using Cfg = GemmCfg<BN, SK>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
    uint8_t* staging = smem + STAGES * Cfg::STAGE_BYTES;               // 1024-byte aligned (stage sizes are)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* res_bar = tempty_bar + 2;                                 // [EPI_WARPS][EPI_BUFS] residual tile landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + RES_BARS);

    pdl_trigger();                               // let the next kernel's launch + prologue overlap this one
    if (threadIdx.x == 0) {
        if (g.trace) g.trace[blockIdx.x * 8 + 4] = 0;
        trace_stamp(g, 0);                       // CTA entered
    }
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m = (g.M + BM - 1) / BM;
    const int num_n = (g.N + BN - 1) / BN;
    const int total_kb = (g.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        if (g.tma_epi) {
            tma_prefetch_desc(&tma_c);
            tma_prefetch_desc(&tma_r);
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
        for (int s = 0; s < RES_BARS; ++s) mbar_init(&res_bar[s], 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0 && g.trace) {            // slot 1: the SM this CTA runs on (tools/dec_gemm_skew.py)
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g.trace[blockIdx.x * 8 + 1] = smid;
    }

    if (warp == 0 || (warp == 3 && g.producers == 2)) {
        // ------------------------------------------------------------------ TMA producer(s)
        // g.producers == 2 (stream-K decode GEMMs, TEO_SK_PRODUCERS=2): warp 3, otherwise idle, is a second producer — the two take
        // alternate k-blocks of the CTA's share (same schedule, same ring positions), doubling the rate at which one SM can issue
        // TMA requests; each full barrier is still armed and fed by exactly one of them.
        const int pid = warp == 3 ? 1 : 0, np = g.producers == 2 ? 2 : 1;
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            // operand loads: weights may live in the blocked layout (contiguous 16 KiB tiles → full-rate HBM bursts)
            auto load_a = [&](int st, int kb, int m_blk) {
                if (g.w_is_a && g.k_wrap) kb %= g.k_wrap;
                if (g.w_is_a && g.w_blocked) tma_load_4d(smem_a + st * A_STAGE_BYTES, &tma_a, &full_bar[st], 0, 0, kb, m_blk);
                else tma_load_2d(smem_a + st * A_STAGE_BYTES, &tma_a, &full_bar[st], kb * BK, m_blk * BM);
            };
            auto load_b = [&](int st, int kb, int n_blk) {
                if (!g.w_is_a && g.k_wrap) kb %= g.k_wrap;
                if (!g.w_is_a && g.w_blocked) tma_load_4d(smem_b + st * Cfg::B_STAGE_BYTES, &tma_b, &full_bar[st], 0, 0, kb, n_blk * (BN / 128));
                else tma_load_2d(smem_b + st * Cfg::B_STAGE_BYTES, &tma_b, &full_bar[st], kb * BK, n_blk * BN);
            };
            // Under programmatic dependent launch the activations are produced by the previous kernel but the
            // weights are constant: fill the ring with weight tiles first, wait for the dependency, then add
            // the activation tiles of those stages (each full barrier expects both).
            int pre = 0;
            {
                Scheduler sc;
                sc.init(g, num_m, num_n, total_kb);
                WorkItem it;
                int cnt = 0;
                while (cnt < STAGES && sc.next(g, it)) {
                    for (int kb = it.kb0; kb < it.kb1 && cnt < STAGES; ++kb, ++cnt) {
                        if (cnt % np != pid) continue;
                        mbar_arrive_expect_tx(&full_bar[cnt], Cfg::STAGE_BYTES);
                        if (g.w_is_a) load_a(cnt, kb, it.m_blk);
                        else load_b(cnt, kb, it.n_blk);
                    }
                }
                pre = cnt;
            }
            if (pid == 0) trace_stamp(g, 2);     // ring filled with weight tiles
            pdl_wait();
            if (pid == 0) trace_stamp(g, 3);     // predecessor grid complete
            int done = 0;
            Scheduler sc;
            sc.init(g, num_m, num_n, total_kb);
            WorkItem it;
            while (sc.next(g, it)) {
                for (int kb = it.kb0; kb < it.kb1; ++kb, ++done) {
                    if (done % np == pid) {
                        if (done < pre) {               // weights already in flight: add the other operand
                            if (g.w_is_a) load_b(s, kb, it.n_blk);
                            else load_a(s, kb, it.m_blk);
                        } else {
                            mbar_wait(&empty_bar[s], ph ^ 1);
                            mbar_arrive_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                            load_a(s, kb, it.m_blk);
                            load_b(s, kb, it.n_blk);
                        }
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            int s = 0, as = 0;
            uint32_t ph = 0, aph = 0;
            Scheduler sc;
            sc.init(g, num_m, num_n, total_kb);
            WorkItem it;
            while (sc.next(g, it)) {
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = it.kb0; kb < it.kb1; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    if (g.trace && g.trace[blockIdx.x * 8 + 4] == 0) trace_stamp(g, 4);     // first operand pair landed
                    const uint64_t a_desc = umma_desc_k_sw128(smem_u32(smem_a + s * A_STAGE_BYTES));
                    const uint64_t b_desc = umma_desc_k_sw128(smem_u32(smem_b + s * Cfg::B_STAGE_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advancing K by 16 bf16 = 32 bytes inside the swizzle row: +2 in the
                        // (address >> 4) field of the descriptor
                        umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > it.kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);          // frees the ring slot when the MMAs retire
                    if (kb == it.kb1 - 1) umma_commit(&tfull_bar[as]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                if (++as == 2) { as = 0; aph ^= 1; }
            }
            trace_stamp(g, 5);                   // last MMA issued
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue
        pdl_wait();                              // outputs / residual / partial workspace belong to the previous kernel until now
        const int ew = warp - 4;                 // 0..7
        const int q = ew & 3;                    // TMEM lane quadrant this warp may access (= warp % 4)
        const int hsel = ew >> 2;                // which column chunks of the tile this warp owns
        uint8_t* stg = staging + ew * 4096 * EPI_BUFS;
        uint64_t* rbars = &res_bar[ew * EPI_BUFS];
        uint32_t nchunk = 0;
        int as = 0;
        uint32_t aph = 0;
        Scheduler sc;
        sc.init(g, num_m, num_n, total_kb);
        WorkItem it;
        while (sc.next(g, it)) {
            const int m_blk = it.m_blk, n_blk = it.n_blk, split = it.slot;
            const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
            if (!SK && g.tma_epi) {             // (the stream-K instantiation has no staging buffer and never runs this epilogue)
                staged_epilogue<BN, 1>(g, &tma_c, &tma_r, t_acc, 0u, m_blk, 0, n_blk, stg, rbars, nchunk, lane, q, hsel,
                                       [&] {
                                           mbar_wait(&tfull_bar[as], aph);
                                           tc_fence_after();
                                       },
                                       [&] {
                                           tc_fence_before();
                                           mbar_arrive(&tempty_bar[as]);
                                       });
            } else {
                mbar_wait(&tfull_bar[as], aph);
                tc_fence_after();
                // ---- direct path: fp32 / split-K partial / transposed (swap-AB) outputs
                const int m = m_blk * BM + q * 32 + lane;
                const bool m_ok = m < g.M;
                const bool partial = g.streamk != 0;
#pragma unroll 1
                for (int c0 = hsel * 32; c0 < BN; c0 += 64) {
                    const int n0 = n_blk * BN + c0;
                    if (n0 >= g.N) break;            // warp-uniform
                    uint32_t v[32];
                    tmem_ld_32x32(t_acc + c0, v);
                    tmem_ld_wait();
                    if (g.transposed) {
                        // C[n, m]: lanes hold consecutive m → coalesced along m for each n
                        if (partial) {
                            float* P = reinterpret_cast<float*>(g.C) + static_cast<long long>(split) * g.split_stride;
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (m_ok && n0 + j < g.N) P[static_cast<long long>(n0 + j) * g.ldc + m] = __uint_as_float(v[j]);
                        } else {
                            const float bm = (g.bias && m_ok) ? __bfloat162float(g.bias[m]) : 0.f;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                if (m_ok && n0 + j < g.N) {
                                    float x = apply_act(__uint_as_float(v[j]) + bm, g.act);
                                    if (g.residual) x += __bfloat162float(g.residual[static_cast<long long>(n0 + j) * g.ldr + m]);
                                    const long long o = static_cast<long long>(n0 + j) * g.ldc + m;
                                    if (g.out_fp32) reinterpret_cast<float*>(g.C)[o] = x;
                                    else reinterpret_cast<bf16*>(g.C)[o] = __float2bfloat16_rn(x);
                                }
                            }
                        }
                    } else if (m_ok) {
                        if (partial) {
                            float* P = reinterpret_cast<float*>(g.C) + static_cast<long long>(split) * g.split_stride +
                                       static_cast<long long>(m) * g.ldc + n0;
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                if (n0 + j < g.N)
                                    *reinterpret_cast<float4*>(P + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                                    __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                        } else {
#pragma unroll
                            for (int j0 = 0; j0 < 32; j0 += 8) {
                                if (n0 + j0 >= g.N) break;
                                float x[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[j0 + j]);
                                if (g.bias) {
                                    const uint4 bv = *reinterpret_cast<const uint4*>(g.bias + n0 + j0);
                                    const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                                    for (int j = 0; j < 4; ++j) { x[2 * j] += bf16_lo(bw[j]); x[2 * j + 1] += bf16_hi(bw[j]); }
                                }
                                if (g.act != TEO_ACT_NONE) {
#pragma unroll
                                    for (int j = 0; j < 8; ++j) x[j] = apply_act(x[j], g.act);
                                }
                                if (g.residual && g.residual_f32) {
                                    const float* rp = reinterpret_cast<const float*>(g.residual) + static_cast<long long>(m) * g.ldr + n0 + j0;
                                    const float4 r0 = *reinterpret_cast<const float4*>(rp), r1 = *reinterpret_cast<const float4*>(rp + 4);
                                    x[0] += r0.x; x[1] += r0.y; x[2] += r0.z; x[3] += r0.w;
                                    x[4] += r1.x; x[5] += r1.y; x[6] += r1.z; x[7] += r1.w;
                                } else if (g.residual) {
                                    const uint4 rv = *reinterpret_cast<const uint4*>(g.residual + static_cast<long long>(m) * g.ldr + n0 + j0);
                                    const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                                    for (int j = 0; j < 4; ++j) { x[2 * j] += bf16_lo(rw[j]); x[2 * j + 1] += bf16_hi(rw[j]); }
                                }
                                if (g.out_fp32) {
                                    float* o = reinterpret_cast<float*>(g.C) + static_cast<long long>(m) * g.ldc + n0 + j0;
                                    *reinterpret_cast<float4*>(o) = make_float4(x[0], x[1], x[2], x[3]);
                                    *reinterpret_cast<float4*>(o + 4) = make_float4(x[4], x[5], x[6], x[7]);
                                } else {
                                    bf16* o = reinterpret_cast<bf16*>(g.C) + static_cast<long long>(m) * g.ldc + n0 + j0;
                                    *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]),
                                                                              pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
                                }
                            }
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&tempty_bar[as]);
                if constexpr (SK) {
                    if (g.fuse) {
                        // ---- in-kernel reduction of this weight tile (all eight epilogue warps).  Every slot's partial is in global
                        // memory first (fixed summation order s = 0,1,… over the slots, like the glue kernel: same code, same bits);
                        // slots > 0 then announce themselves, slot 0 — for a tile shared with other CTAs always this CTA's LAST work
                        // item: its range ends inside the tile — waits for them and reduces.
                        const int tile = m_blk * num_n + n_blk;
                        const int et = static_cast<int>(threadIdx.x) - 128;
                        __threadfence();
                        named_bar_sync_epi();
                        if (split != 0) {
                            if (et == 0) atomicAdd(&g.sk_flags[tile], 1);
                        } else {
                            const PartialInfo pi{reinterpret_cast<const float*>(g.C), g.split_stride, total_kb, g.sk_q, static_cast<int>(gridDim.x)};
                            const int others = partial_count(pi, tile * BM) - 1;
                            if (others > 0) {
                                if (et == 0) {
                                    const long long t0 = clock64();
                                    while (*reinterpret_cast<volatile int*>(&g.sk_flags[tile]) < others) {
                                        if (clock64() - t0 > 8000000000LL) {
                                            printf("teochat_b200: stream-K fix-up timed out (block %d tile %d)\n", blockIdx.x, tile);
                                            __trap();
                                        }
                                    }
                                    g.sk_flags[tile] = 0;              // ready for the next launch
                                    __threadfence();
                                }
                                named_bar_sync_epi();
                            }
                            reduce_swiglu_cols(pi, g.fuse_out, g.N, g.fuse_inter, 1, tile * (BM / 2), BM / 2, et, EPI_WARPS * 32);
                        }
                    }
                }
            }
            if (++as == 2) { as = 0; aph ^= 1; }
        }
        if (g.tma_epi && lane == 0) tma_store_wait_all<0>();   // all output tiles written before the CTA retires
        if (warp == 4 && lane == 0) trace_stamp(g, 6);         // epilogue done
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    if (threadIdx.x == 0) trace_stamp(g, 7);                   // CTA leaving
}

// out[r,c] = act(Σ_slots partial[slot][r,c] + bias[c]) + residual[r,c] — fixed summation order (deterministic).
__global__ void splitk_reduce_kernel(PartialInfo pi, void* out, long long ldo, const bf16* __restrict__ bias,
                                     const bf16* residual, long long ldr, int act, int out_fp32, int res_f32, int rows, int cols) {
    pdl_trigger();
    pdl_wait();
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<long long>(rows) * cols) return;
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
    const int n = partial_count(pi, c);
    float acc = 0.f;
    for (int s = 0; s < n; ++s) acc += pi.P[s * pi.stride + i];
    if (bias) acc += __bfloat162float(bias[c]);
    acc = apply_act(acc, act);
    if (residual) acc += res_f32 ? reinterpret_cast<const float*>(residual)[r * ldr + c] : __bfloat162float(residual[r * ldr + c]);
    if (out_fp32) reinterpret_cast<float*>(out)[r * ldo + c] = acc;
    else reinterpret_cast<bf16*>(out)[r * ldo + c] = __float2bfloat16_rn(acc);
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int get_tmap_bf16(teo_handle* h, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                  CUtensorMap* out) {
    TmapKey key{ptr, rows, cols, ld, box_rows};
    auto it = h->tmaps.find(key);
    if (it != h->tmaps.end()) {
        *out = it->second;
        return TEO_OK;
    }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return TEO_ERR_CUDA;
    }
    TEO_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand %p not 16-byte aligned", ptr);
    TEO_CHECK_ARG((ld * 2) % 16 == 0, "TMA operand leading dimension %llu not a multiple of 8 elements",
                  (unsigned long long)ld);
    CUtensorMap m;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %llu cols %llu ld %llu box %u)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
        return TEO_ERR_CUDA;
    }
    if (h->tmaps.size() > 4096) h->tmaps.clear();   // bounded cache (callers hold copies, never references)
    h->tmaps.emplace(key, m);
    *out = m;
    return TEO_OK;
}

// Weights in the blocked layout [N/128][K/64][128][64] (bf16): 4-D map, box = `nblocks` consecutive 128-row tiles
// of one k block, i.e. nblocks × 16 KiB contiguous in HBM, written to shared memory as (nblocks·128) rows × 128 B
// under the same 128-byte swizzle as the 2-D path.
int get_tmap_wblocked(teo_handle* h, const void* ptr, uint64_t N, uint64_t K, uint32_t nblocks, CUtensorMap* out) {
    TmapKey key{ptr, N, K, 0xB10CB10CULL, nblocks};
    auto it = h->tmaps.find(key);
    if (it != h->tmaps.end()) {
        *out = it->second;
        return TEO_OK;
    }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
        return TEO_ERR_CUDA;
    }
    TEO_CHECK_ARG(N % 128 == 0 && K % 64 == 0, "blocked weight layout needs N %% 128 == 0 and K %% 64 == 0 (N=%llu K=%llu)",
                  (unsigned long long)N, (unsigned long long)K);
    TEO_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 127) == 0, "blocked weights must be 128-byte aligned");
    CUtensorMap m;
    cuuint64_t dims[4] = {64, 128, K / 64, N / 128};
    cuuint64_t strides[3] = {128, 16384, 16384ULL * (K / 64)};
    cuuint32_t box[4] = {64, 128, 1, nblocks};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (blocked weights) failed with CUresult %d (N %llu K %llu)", (int)r, (unsigned long long)N,
                  (unsigned long long)K);
        return TEO_ERR_CUDA;
    }
    if (h->tmaps.size() > 4096) h->tmaps.clear();
    h->tmaps.emplace(key, m);
    *out = m;
    return TEO_OK;
}

// development hook (tools/dec_gemm_bench.py): device buffer of 8 u64 per CTA that the next GEMM launches stamp with
// %globaltimer at their phase boundaries; nullptr (the default) compiles to one predicated-off branch per stamp
static unsigned long long* g_gemm_trace = nullptr;
static int g_gemm_trace_cap = 0;         // launches the buffer holds (used as a ring)
static long long g_gemm_trace_n = 0;     // launches stamped so far
static unsigned long long* next_trace_slot() {
    if (!g_gemm_trace || g_gemm_trace_cap <= 0) return nullptr;
    return g_gemm_trace + (g_gemm_trace_n++ % g_gemm_trace_cap) * (148 * 8);
}

bool gemm_pair_enabled();
int launch_gemm_pair(teo_handle* h, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tr,
                     const GemmArgs& g, cudaStream_t stream);

// TMA producer threads per CTA in the stream-K (decode) schedule: TEO_SK_PRODUCERS=1|2
static int sk_producers() {
    static const int n = [] {
        const char* e = getenv("TEO_SK_PRODUCERS");
        return (e && e[0] == '2') ? 2 : 1;
    }();
    return n;
}

struct GemmPlan {
    bool swap;
    int bn;
    int sk_q, sk_grid, sk_smax;      // stream-K schedule (swap mode)
};

static GemmPlan plan_gemm(int M, int N, int K, int num_sms) {
    GemmPlan p{};
    const int total_kb = (K + BK - 1) / BK;
    p.swap = (M <= 128) && (N >= 256);
    if (p.swap) {
        p.bn = M <= 32 ? 32 : (M <= 64 ? 64 : 128);
        const int tiles = (N + BM - 1) / BM;
        const long long total = static_cast<long long>(tiles) * total_kb;
        long long grid = std::min<long long>(num_sms, static_cast<long long>(tiles) * 16);   // ≤ ~17 partial slots per tile
        grid = std::max<long long>(1, std::min(grid, total));
        p.sk_q = static_cast<int>((total + grid - 1) / grid);
        p.sk_grid = static_cast<int>((total + p.sk_q - 1) / p.sk_q);
        p.sk_smax = (total_kb + p.sk_q - 2) / p.sk_q + 1;
    } else {
        p.bn = N >= 256 ? 256 : (N > 64 ? 128 : 64);
    }
    return p;
}

}  // namespace teo

using namespace teo;

extern "C" size_t teo_gemm_workspace_bytes(int M, int N, int K) {
    GemmPlan p = plan_gemm(M, N, K, 148);
    if (!p.swap) return 0;
    // upper bound on partial slots per tile for any SM count (plan_gemm caps CTAs at 16 per tile)
    return static_cast<size_t>(19) * static_cast<size_t>(M) * static_cast<size_t>(N) * sizeof(float);
}

template <int BN, bool SK = false>
static int launch_cfg(teo_handle* h, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tr,
                      const GemmArgs& g, int units, cudaStream_t stream) {
    using Cfg = GemmCfg<BN, SK>;
    static bool attr_set = false;
    if (!attr_set) {
        TEO_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<BN, SK>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int grid = std::min(units, h->num_sms);
    TEO_CUDA(launch_kc(PDL_GEMM, gemm_tn_kernel<BN, SK>, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, ta, tb, tc, tr, g));
    TEO_LAUNCH_CHECK("gemm_tn_kernel");
    h->launches++;
    return TEO_OK;
}

// Small-M GEMM that stops at the fp32 split-K partials: P[s][M][N] in `workspace`, s < *splits_out.
// The caller's next kernel reduces them (fused with its own work) in the fixed order s = 0,1,...
int teo::launch_gemm_partials(teo_handle* h, const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, void* workspace,
                              size_t workspace_bytes, PartialInfo* info, cudaStream_t stream, int w_blocked, const SkFuse* fuse) {
    TEO_CHECK_ARG(h != nullptr && info != nullptr, "gemm_partials: null handle");
    TEO_CHECK_ARG(M > 0 && M <= 128 && N >= 256 && K > 0 && K % 8 == 0, "gemm_partials: needs 0 < M <= 128, N >= 256, K %% 8 == 0");
    const GemmPlan p = plan_gemm(M, N, K, h->num_sms);
    const size_t need = static_cast<size_t>(p.sk_smax) * M * N * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("gemm_partials: need %zu workspace bytes, got %zu", need, workspace_bytes);
        return TEO_ERR_WORKSPACE;
    }
    GemmArgs g{};
    g.trace = next_trace_slot();
    g.M = N;
    g.N = M;
    g.K = K;
    g.transposed = 1;
    g.streamk = 1;
    g.sk_q = p.sk_q;
    g.producers = sk_producers();
    g.w_is_a = 1;
    g.C = workspace;
    g.ldc = N;
    g.split_stride = static_cast<long long>(M) * N;
    if (fuse != nullptr && fuse->kind != 0) {
        TEO_CHECK_ARG(fuse->kind == 1 && fuse->out && N == 2 * fuse->inter && fuse->inter % 64 == 0 && N % BM == 0 && N / BM <= 1024,
                      "gemm_partials: the fused SwiGLU reduction needs interleaved gate/up rows, N = 2·inter, inter %% 64 == 0, ≤ 1024 tiles");
        if (h->sk_flags == nullptr) {                 // (first call is the eager step that precedes graph capture)
            TEO_CUDA(cudaMalloc(&h->sk_flags, 1024 * sizeof(int)));
            TEO_CUDA(cudaMemset(h->sk_flags, 0, 1024 * sizeof(int)));
        }
        g.fuse = fuse->kind;
        g.sk_flags = h->sk_flags;
        g.fuse_out = fuse->out;
        g.fuse_inter = fuse->inter;
        g.fuse_interleaved = 1;
    }
    CUtensorMap ta, tb;
    g.w_blocked = w_blocked ? 1 : 0;
    if (w_blocked) TEO_TRY(get_tmap_wblocked(h, W, N, K, 1, &ta));
    else TEO_TRY(get_tmap_bf16(h, W, N, K, ldw, BM, &ta));
    TEO_TRY(get_tmap_bf16(h, A, M, K, lda, p.bn, &tb));
    int rc;
    switch (p.bn) {
        case 32: rc = launch_cfg<32, true>(h, ta, tb, ta, ta, g, p.sk_grid, stream); break;
        case 64: rc = launch_cfg<64, true>(h, ta, tb, ta, ta, g, p.sk_grid, stream); break;
        default: rc = launch_cfg<128, true>(h, ta, tb, ta, ta, g, p.sk_grid, stream); break;
    }
    TEO_TRY(rc);
    info->P = static_cast<const float*>(workspace);
    info->stride = static_cast<long long>(M) * N;
    info->kb = (K + BK - 1) / BK;
    info->q = p.sk_q;
    info->grid = p.sk_grid;
    return TEO_OK;
}

int teo::launch_gemm(teo_handle* h, const bf16* A, int lda, const bf16* W, int ldw, void* C, int ldc, int M, int N,
                     int K, const GemmEpilogue& ep, void* workspace, size_t workspace_bytes, cudaStream_t stream, int w_blocked) {
    TEO_CHECK_ARG(h != nullptr, "null handle");
    TEO_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: non-positive size M=%d N=%d K=%d", M, N, K);
    TEO_CHECK_ARG(K % 8 == 0 && N % 8 == 0, "gemm: K (%d) and N (%d) must be multiples of 8", K, N);
    TEO_CHECK_ARG(lda >= K && (w_blocked || ldw >= K) && ldc >= (ep.act == TEO_ACT_SWIGLU_PAIRS ? N / 2 : N),
                  "gemm: leading dimension too small");
    TEO_CHECK_ARG(!w_blocked || (N % 128 == 0 && K % 64 == 0), "gemm: blocked weights need N %% 128 == 0 and K %% 64 == 0");
    TEO_CHECK_ARG(ldc % (ep.out_fp32 ? 4 : 8) == 0, "gemm: ldc (%d) breaks 16-byte row alignment", ldc);
    TEO_CHECK_ARG(ep.residual == nullptr || ep.ldr % 8 == 0, "gemm: ldr (%d) must be a multiple of 8", ep.ldr);
    TEO_CHECK_ARG((reinterpret_cast<uintptr_t>(C) & 15) == 0, "gemm: C not 16-byte aligned");
    TEO_CHECK_ARG(ep.k_planes >= 1 && ep.k_planes <= 3 && (ep.k_planes == 1 || (K % BK == 0 && ep.out_fp32)),
                  "gemm: split operands (k_planes=%d) need K %% 64 == 0 and fp32 output", ep.k_planes);
    TEO_CHECK_ARG(!ep.residual_f32 || (ep.out_fp32 && ep.ldr % 4 == 0), "gemm: an fp32 residual needs fp32 output and ldr %% 4 == 0");
    const int Kw = K;                 // columns of W
    K *= ep.k_planes;                 // contraction length seen by the kernel (= columns of A)
    TEO_CHECK_ARG(lda >= K, "gemm: lda (%d) smaller than k_planes * K (%d)", lda, K);
    TEO_CHECK_ARG(ep.bias == nullptr || (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0, "gemm: bias not 16-byte aligned");
    TEO_CHECK_ARG(ep.residual == nullptr || (reinterpret_cast<uintptr_t>(ep.residual) & 15) == 0,
                  "gemm: residual not 16-byte aligned");
    const GemmPlan p = plan_gemm(M, N, K, h->num_sms);
    if (ep.act == TEO_ACT_SWIGLU_PAIRS) {
        TEO_CHECK_ARG(!ep.out_fp32 && !ep.bias && !ep.residual && N % 128 == 0 && N >= 256,
                      "gemm: the SwiGLU-pairs epilogue needs bf16 output, no bias/residual and N %% 128 == 0 (N=%d)", N);
        if (p.swap) {
            set_error("gemm: the SwiGLU-pairs epilogue runs on the tiled schedule only (M=%d <= 128 takes the swap-AB path)", M);
            return TEO_ERR_UNSUPPORTED;
        }
    }
    const bool ln_any = ep.ln_stats != nullptr || ep.stats_out != nullptr;
    if (ln_any) {
        TEO_CHECK_ARG(!p.swap && !ep.out_fp32 && ep.act != TEO_ACT_SWIGLU_PAIRS,
                      "gemm: folded LayerNorm / row statistics need the tiled schedule (M > 128 or N < 256) with bf16 output");
        TEO_CHECK_ARG(ep.ln_stats == nullptr || (ep.ln_c && ep.ln_bias && ep.ln_slots > 0 && ep.bias == nullptr && ep.k_planes == 1 &&
                                                 (reinterpret_cast<uintptr_t>(ep.ln_c) & 15) == 0 && (reinterpret_cast<uintptr_t>(ep.ln_bias) & 15) == 0),
                      "gemm: folded LayerNorm needs ln_c, ln_bias (16-byte aligned), ln_slots > 0 and no separate bias");
    }
    GemmArgs g{};
    g.trace = next_trace_slot();
    g.ln_stats = ep.ln_stats;
    g.ln_c = ep.ln_c;
    g.ln_bias = ep.ln_bias;
    g.ln_slots = ep.ln_slots;
    g.ln_inv_d = 1.0f / static_cast<float>(Kw);
    g.ln_eps = ep.ln_eps;
    g.stats_out = ep.stats_out;
    g.stats_slots = 2 * ((N + p.bn - 1) / p.bn);
    g.act = ep.act;
    g.out_fp32 = ep.out_fp32;
    g.bias = ep.bias;
    g.residual = ep.residual;
    g.ldr = ep.ldr;
    g.residual_f32 = ep.residual_f32;
    g.k_wrap = ep.k_planes > 1 ? Kw / BK : 0;
    g.C = C;
    g.ldc = ldc;
    g.K = K;
    CUtensorMap ta, tb, tc, tr;
    if (p.swap) {
        g.M = N;   // weight rows on the UMMA M dimension
        g.w_is_a = 1;
            g.N = M;
        g.transposed = 1;
        g.w_blocked = w_blocked ? 1 : 0;
        if (w_blocked) TEO_TRY(get_tmap_wblocked(h, W, N, Kw, 1, &ta));
        else TEO_TRY(get_tmap_bf16(h, W, N, Kw, ldw, BM, &ta));
        TEO_TRY(get_tmap_bf16(h, A, M, K, lda, p.bn, &tb));
        {
            const size_t need = static_cast<size_t>(p.sk_smax) * M * N * sizeof(float);
            if (workspace == nullptr || workspace_bytes < need) {
                set_error("gemm: small-M schedule needs %zu workspace bytes, got %zu", need, workspace_bytes);
                return TEO_ERR_WORKSPACE;
            }
            g.streamk = 1;
            g.sk_q = p.sk_q;
            g.producers = sk_producers();
            g.C = workspace;
            g.ldc = N;                       // partial layout [slot][M_act][N_out]
            g.split_stride = static_cast<long long>(M) * N;
        }
    } else {
        g.M = M;
        g.N = N;
        g.transposed = 0;
        TEO_TRY(get_tmap_bf16(h, A, M, K, lda, BM, &ta));
        if (p.bn == 256 && !ep.out_fp32 && gemm_pair_enabled()) {
            // large tiled GEMM with bf16 output: the CTA-pair kernel (gemm_pair.cu), 256 x 256 tiles, each CTA loads 128 W rows
            g.w_blocked = w_blocked ? 1 : 0;
            g.tma_epi = 1;
            if (w_blocked) TEO_TRY(get_tmap_wblocked(h, W, N, Kw, 1, &tb));
            else TEO_TRY(get_tmap_bf16(h, W, N, Kw, ldw, BM, &tb));
            TEO_TRY(get_tmap_bf16(h, C, M, ep.act == TEO_ACT_SWIGLU_PAIRS ? N / 2 : N, ldc, 32, &tc));
            tr = tc;
            if (ep.residual) TEO_TRY(get_tmap_bf16(h, ep.residual, M, N, ep.ldr, 32, &tr));
            return launch_gemm_pair(h, ta, tb, tc, tr, g, stream);
        }
        g.w_blocked = (w_blocked && p.bn >= 128) ? 1 : 0;
        TEO_CHECK_ARG(!w_blocked || p.bn >= 128, "gemm: blocked weights need N >= 128");
        if (g.w_blocked) TEO_TRY(get_tmap_wblocked(h, W, N, Kw, p.bn / 128, &tb));
        else TEO_TRY(get_tmap_bf16(h, W, N, Kw, ldw, p.bn, &tb));
        if (!ep.out_fp32) {           // staged TMA-store epilogue: 32-row × 64-column boxes of C (and of the residual)
            g.tma_epi = 1;
            TEO_TRY(get_tmap_bf16(h, C, M, ep.act == TEO_ACT_SWIGLU_PAIRS ? N / 2 : N, ldc, 32, &tc));
            tr = tc;
            if (ep.residual) TEO_TRY(get_tmap_bf16(h, ep.residual, M, N, ep.ldr, 32, &tr));
        }
    }
    if (!g.tma_epi) tc = tr = ta;     // unused by the direct-store epilogue
    const int units = p.swap ? p.sk_grid : ((g.M + BM - 1) / BM) * ((g.N + p.bn - 1) / p.bn);
    int rc;
    if (p.swap) {                     // stream-K schedule: deeper ring, no staging buffer
        switch (p.bn) {
            case 32: rc = launch_cfg<32, true>(h, ta, tb, tc, tr, g, units, stream); break;
            case 64: rc = launch_cfg<64, true>(h, ta, tb, tc, tr, g, units, stream); break;
            default: rc = launch_cfg<128, true>(h, ta, tb, tc, tr, g, units, stream); break;
        }
    } else {
        switch (p.bn) {
            case 32: rc = launch_cfg<32>(h, ta, tb, tc, tr, g, units, stream); break;
            case 64: rc = launch_cfg<64>(h, ta, tb, tc, tr, g, units, stream); break;
            case 128: rc = launch_cfg<128>(h, ta, tb, tc, tr, g, units, stream); break;
            default: rc = launch_cfg<256>(h, ta, tb, tc, tr, g, units, stream); break;
        }
    }
    TEO_TRY(rc);
    if (p.swap) {
        const long long total = static_cast<long long>(M) * N;
        const int threads = 256;
        const int blocks = static_cast<int>((total + threads - 1) / threads);
        PartialInfo pi{reinterpret_cast<const float*>(workspace), total, (K + BK - 1) / BK, p.sk_q, p.sk_grid};
        TEO_CUDA(launch_k(splitk_reduce_kernel, dim3(blocks), dim3(threads), 0, stream, pi, C, static_cast<long long>(ldc), ep.bias, ep.residual,
                          static_cast<long long>(ep.ldr), ep.act, ep.out_fp32, ep.residual_f32, M, N));
        TEO_LAUNCH_CHECK("splitk_reduce_kernel");
        h->launches++;
    }
    return TEO_OK;
}

extern "C" int teo_gemm_stats_slots(int M, int N, int K) {
    const GemmPlan p = plan_gemm(M, N, K, 148);
    return p.swap ? 0 : 2 * ((N + p.bn - 1) / p.bn);
}

extern "C" int teo_gemm_bf16_ex(teo_handle* h, const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M, int N, int K,
                                const teo_gemm_opts* o, void* workspace, size_t workspace_bytes, void* stream) {
    TEO_CHECK_ARG(A && W && C && o, "gemm_ex: null operand");
    TEO_CHECK_ARG(o->act >= TEO_ACT_NONE && o->act <= TEO_ACT_SWIGLU_PAIRS, "gemm_ex: unknown activation %d", o->act);
    GemmEpilogue ep;
    ep.bias = static_cast<const bf16*>(o->bias);
    ep.residual = static_cast<const bf16*>(o->residual);
    ep.ldr = o->ldr;
    ep.act = o->act;
    ep.out_fp32 = o->out_fp32;
    ep.ln_stats = static_cast<const float*>(o->ln_stats);
    ep.ln_c = static_cast<const float*>(o->ln_c);
    ep.ln_bias = static_cast<const float*>(o->ln_bias);
    ep.ln_slots = o->ln_slots;
    ep.ln_eps = o->ln_eps;
    ep.stats_out = static_cast<float*>(o->stats_out);
    return launch_gemm(h, static_cast<const bf16*>(A), lda, static_cast<const bf16*>(W), o->w_blocked ? K : ldw, C, ldc, M, N, K, ep, workspace,
                       workspace_bytes, static_cast<cudaStream_t>(stream), o->w_blocked);
}

extern "C" int teo_gemm_bf16(teo_handle* h, const void* A, int lda, const void* W, int ldw, void* C, int ldc, int M,
                             int N, int K, const void* bias, const void* residual, int ldr, int act, int out_fp32,
                             void* workspace, size_t workspace_bytes, void* stream) {
    TEO_CHECK_ARG(A && W && C, "gemm: null operand");
    TEO_CHECK_ARG(act >= TEO_ACT_NONE && act <= TEO_ACT_SWIGLU_PAIRS, "gemm: unknown activation %d", act);
    GemmEpilogue ep;
    ep.bias = static_cast<const bf16*>(bias);
    ep.residual = static_cast<const bf16*>(residual);
    ep.ldr = ldr;
    ep.act = act;
    ep.out_fp32 = out_fp32;
    return launch_gemm(h, static_cast<const bf16*>(A), lda, static_cast<const bf16*>(W), ldw, C, ldc, M, N, K, ep,
                       workspace, workspace_bytes, static_cast<cudaStream_t>(stream), 0);
}

extern "C" int teo_gemm_bf16_wblocked(teo_handle* h, const void* A, int lda, const void* W_blocked, void* C, int ldc, int M, int N, int K,
                                      const void* bias, const void* residual, int ldr, int act, int out_fp32, void* workspace,
                                      size_t workspace_bytes, void* stream) {
    TEO_CHECK_ARG(A && W_blocked && C, "gemm: null operand");
    TEO_CHECK_ARG(act >= TEO_ACT_NONE && act <= TEO_ACT_SWIGLU_PAIRS, "gemm: unknown activation %d", act);
    GemmEpilogue ep;
    ep.bias = static_cast<const bf16*>(bias);
    ep.residual = static_cast<const bf16*>(residual);
    ep.ldr = ldr;
    ep.act = act;
    ep.out_fp32 = out_fp32;
    return launch_gemm(h, static_cast<const bf16*>(A), lda, static_cast<const bf16*>(W_blocked), K, C, ldc, M, N, K, ep, workspace,
                       workspace_bytes, static_cast<cudaStream_t>(stream), 1);
}

// row-major [N,K] → blocked [N/128][K/64][128][64]
__global__ void block_weight_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int N, int K) {
    const long long total = static_cast<long long>(N) * K / 8;
    const int k8 = K / 8;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long n = i / k8;
        const int c8 = static_cast<int>(i % k8);               // 16-byte chunk along K
        const long long nb = n / 128, r = n % 128;
        const int kb = c8 / 8, cc = c8 % 8;
        dst[((nb * (K / 64) + kb) * 128 + r) * 8 + cc] = src[i];
    }
}
extern "C" int teo_weight_to_blocked(const void* w_rowmajor, void* w_blocked, int N, int K, void* stream) {
    TEO_CHECK_ARG(w_rowmajor && w_blocked && w_rowmajor != w_blocked, "weight_to_blocked: bad pointers (out of place only)");
    TEO_CHECK_ARG(N > 0 && K > 0 && N % 128 == 0 && K % 64 == 0, "weight_to_blocked: N %% 128 and K %% 64 must be 0 (N=%d K=%d)", N, K);
    block_weight_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4*>(w_rowmajor),
                                                                               static_cast<uint4*>(w_blocked), N, K);
    TEO_LAUNCH_CHECK("block_weight_kernel");
    return TEO_OK;
}

// Development hook, not part of include/teochat_b200.h: see g_gemm_trace.  `launches` slots of 148 x 8 u64, used as a ring;
// returns the number of launches stamped since the previous call.
extern "C" long long teo_dbg_gemm_trace(void* device_buffer, int launches) {
    const long long n = g_gemm_trace_n;
    g_gemm_trace = static_cast<unsigned long long*>(device_buffer);
    g_gemm_trace_cap = device_buffer ? launches : 0;
    g_gemm_trace_n = 0;
    return n;
}
