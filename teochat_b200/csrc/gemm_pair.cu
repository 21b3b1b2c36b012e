// CTA-pair GEMM (tcgen05 cta_group::2) for the large tiled GEMMs of the hot path — LLaMA prefill (qkv / o / gate_up /
// down over ~68 k tokens) and the ViT / projector linears: C[M,N] = epi(A[M,K] · W[N,K]^T), bf16 in, fp32 accumulate in
// TMEM, bf16 out through the staged TMA-store epilogue of gemm_common.cuh.
//
// A cluster of two CTAs (the two SMs of a TPC) owns a 256 × 256 output tile.  Per 64-wide k-block each CTA stages its
// own 128 rows of A and HALF of the W tile (128 of the 256 W rows): 32 KiB per stage instead of the single-CTA kernel's
// 48 KiB, so the ring is 6 deep and every SM reads a third less shared memory per flop.  The leader CTA (cluster rank 0)
// issues tcgen05.mma.cta_group::2 (UMMA 256 × 256 × 16); the hardware reads A rows 128..255 and W rows 128..255 from the
// peer's shared memory at the same offsets and writes each CTA's 128 accumulator rows into that CTA's own TMEM.
//
//   warp 0      TMA producer (both CTAs): cp.async.bulk.tensor …cta_group::2, completion bytes credited to the LEADER's
//               full barrier (which expects both CTAs' 2 × 32 KiB)
//   warp 1      MMA issuer (leader only): waits full → 4 UMMAs per k-block → tcgen05.commit …multicast frees the ring
//               slot in BOTH CTAs; after the last k-block a multicast commit publishes the accumulator stage to both
//   warp 2      TMEM allocator (cta_group::2 alloc / dealloc, the same warp in both CTAs)
//   warps 4-11  epilogue (both CTAs, each on its own 128 rows): staged_epilogue; the accumulator stage is released
//               by arriving on the leader's tempty barrier (remote arrive from the peer)
// Persistent: pair p walks tiles p, p + pairs, … rasterised in groups of 16 tile rows (4096 rows of A stay in L2 while W
// streams).  Selected by launch_gemm (gemm.cu) for M > 128, N ≥ 256, bf16 output; TEO_GEMM_PAIR=0 keeps the single-CTA
// kernel (A/B measurements).
#include <stdlib.h>

#include <algorithm>

#include "common.h"
#ifdef TEO_PAIR_TRACE
// (development build) stamps inside the shared epilogue: warp 4 / lane 0 of this translation unit's kernel only
namespace teo {
__device__ unsigned long long* g_pair_trace = nullptr;
__device__ int g_pair_trace_tiles = 0;
__device__ int g_pair_trace_tile_no[148];          // tile the epilogue of each CTA is working on (for the stamps inside staged_epilogue)
}  // namespace teo
#define EPI_STAMP(slot)                                                                                                                  \
    do {                                                                                                                                 \
        if (teo::g_pair_trace != nullptr && (threadIdx.x == 128) && teo::g_pair_trace_tile_no[blockIdx.x] < teo::g_pair_trace_tiles)                    \
            teo::g_pair_trace[(static_cast<long long>(blockIdx.x) * teo::g_pair_trace_tiles + teo::g_pair_trace_tile_no[blockIdx.x]) * 16 + (slot)] = clock64(); \
    } while (0)
#endif
#include "gemm_common.cuh"
#include "ptx.cuh"

namespace teo {

constexpr int PAIR_BN = 256;
constexpr int PAIR_HALF_B_BYTES = (PAIR_BN / 2) * BK * 2;              // 16 KiB: this CTA's half of the W tile
constexpr int PAIR_TMEM_COLS = 2 * PAIR_BN;                             // MT = 1: two accumulator stages; MT = 2: the two row sub-tiles
constexpr int PAIR_GROUP_M = 16;                                        // 256-row tile rows per rasterisation group (4096 rows of A)
// MT = row sub-tiles per pair tile.  MT = 1: 256 × 256 tiles, two TMEM accumulator stages (the epilogue of a tile overlaps the next
// tile's MMAs).  MT = 2: 512 × 256 tiles — both accumulators belong to ONE tile, every W k-block is used by two MMAs, so the
// operand bytes fetched per flop drop by a quarter (48 KiB per 2·128·256·64 MACs instead of 32 KiB per 128·256·64): the main loop
// of the 256 × 256 tiling is bound by what the L2 slices can deliver, at the price of an epilogue that no longer
// overlaps (long-K GEMMs only).
template <int MT>
struct PairCfg {
    static constexpr int STAGE_BYTES = MT * A_STAGE_BYTES + PAIR_HALF_B_BYTES;    // 32 / 48 KiB
    static constexpr int STAGES = (229376 - STAGING_BYTES) / STAGE_BYTES;          // 6 / 4 with one staging buffer per warp, 5 / 3 with two
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + GEMM_BAR_BYTES;
};

#ifdef TEO_PAIR_TRACE
// Development build only (build.build_variant("pairtrace", ["TEO_PAIR_TRACE"]), tools/pair_trace.py): clock64 stamps per tile of every CTA,
// u64 [n_ctas][max_tiles][16]: 0 issuer past tempty · 1 issuer has k-block 0 · 2 issuer committed tfull · 3 epilogue saw tfull ·
// 4 epilogue released the accumulators · 5 epilogue done · 6 producer issued k-block 0 · 7 producer issued the last k-block · 8 / 9 producer saw the ring slot of k-block 0 / STAGES free again
#define PAIR_STAMP(tile, slot)                                                                                         \
    do {                                                                                                               \
        if (g_pair_trace != nullptr && (tile) < g_pair_trace_tiles)                                                    \
            g_pair_trace[(static_cast<long long>(blockIdx.x) * g_pair_trace_tiles + (tile)) * 16 + (slot)] = clock64(); \
    } while (0)
#else
#define PAIR_STAMP(tile, slot) do { } while (0)
#endif

struct PairTile {
    int m2, n_blk;
};
// Tile `unit` of the (num_m2 × num_n) grid.  Supertiles of group_m tile rows × group_n tile columns, M fastest inside one.
//   raster 0  supertiles walk along N inside a band of group_m rows, band after band: the band's A panel (group_m · 256 rows)
//             is what should stay in L2 while W streams past it (group_n ≥ num_n is the one-dimensional grouping of round 1)
//   raster 1  supertiles walk along M inside a strip of group_n columns, strip after strip: the strip's W panel stays, A streams
__device__ __forceinline__ PairTile pair_tile(int unit, int num_m2, int num_n, int group_m, int group_n, int raster) {
    if (raster == 0) {
        const int band_sz = group_m * num_n;
        const int band = unit / band_sz;
        const int first_m = band * group_m;
        const int gm = min(group_m, num_m2 - first_m);
        const int in_band = unit - band * band_sz;
        const int per_sup = gm * group_n;
        const int sn = in_band / per_sup;
        const int first_n = sn * group_n;
        const int in_sup = in_band - sn * per_sup;
        return {first_m + in_sup % gm, first_n + in_sup / gm};
    }
    const int strip_sz = group_n * num_m2;
    const int strip = unit / strip_sz;
    const int first_n = strip * group_n;
    const int gn = min(group_n, num_n - first_n);
    const int in_strip = unit - strip * strip_sz;
    const int per_sup = group_m * gn;
    const int sm = in_strip / per_sup;
    const int first_m = sm * group_m;
    const int gm = min(group_m, num_m2 - first_m);
    const int in_sup = in_strip - sm * per_sup;
    return {first_m + in_sup % gm, first_n + in_sup / gm};
}

template <int MT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_r, const GemmArgs g) {
    constexpr int BN = PAIR_BN;
    constexpr int STAGES = PairCfg<MT>::STAGES;
    constexpr int A_BYTES = MT * A_STAGE_BYTES;              // this CTA's A rows of one k-block: MT sub-tiles of 128 rows
    constexpr int ACC_STAGES = MT == 1 ? 2 : 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_BYTES;
    uint8_t* staging = smem + STAGES * PairCfg<MT>::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + STAGING_BYTES);     // used in the leader only
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;                                          // used in the leader only
    uint64_t* res_bar = tempty_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + RES_BARS);

    pdl_trigger();
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = static_cast<int>(cluster_ctarank());
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int num_m2 = (g.M + MT * 2 * BM - 1) / (MT * 2 * BM);          // pair tiles along M (MT · 256 rows each)
    const int num_n = (g.N + BN - 1) / BN;
    const int total_kb = (g.K + BK - 1) / BK;
    const int units = num_m2 * num_n;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        tma_prefetch_desc(&tma_c);
        tma_prefetch_desc(&tma_r);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 2 * EPI_WARPS * 32);       // the epilogue threads of both CTAs
        }
        for (int s = 0; s < RES_BARS; ++s) mbar_init(&res_bar[s], 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_pair<PAIR_TMEM_COLS>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();                  // the peer's barriers exist before anything is signalled across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            pdl_wait();
            int s = 0;
            uint32_t ph = 0;
            // L2 eviction priorities (l2_hint 1): the operand whose panel is meant to stay resident across the supertiles of a band /
            // strip is fetched evict_last, the operand that streams past it evict_first
            // l2_hint 1: resident evict_last + streaming evict_first; 2: resident evict_last only; 3: streaming evict_first only
            const uint64_t h_stay = (g.l2_hint == 1 || g.l2_hint == 2) ? L2_EVICT_LAST : L2_EVICT_NORMAL;
            const uint64_t h_stream = (g.l2_hint == 1 || g.l2_hint == 3) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
            const uint64_t hint_a = g.raster == 0 ? h_stay : h_stream;
            const uint64_t hint_w = g.raster == 0 ? h_stream : h_stay;
            int tile_no = 0;
            for (int unit = pair; unit < units; unit += n_pairs, ++tile_no) {
                const PairTile t = pair_tile(unit, num_m2, num_n, g.sk_q, g.group_n, g.raster);
                for (int kb = 0; kb < total_kb; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    if (kb == 0) PAIR_STAMP(tile_no, 6);
                    if (kb == total_kb - 1) PAIR_STAMP(tile_no, 7);
                    if (kb == STAGES) PAIR_STAMP(tile_no, 8);
                    if (kb == 2 * STAGES) PAIR_STAMP(tile_no, 9);
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * PairCfg<MT>::STAGE_BYTES);
#pragma unroll
                    for (int sub = 0; sub < MT; ++sub) {             // this CTA's 128 rows of every 256-row sub-tile
                        const int m_blk = (t.m2 * MT + sub) * 2 + rank;
                        uint8_t* dst = smem_a + s * A_BYTES + sub * A_STAGE_BYTES;
                        if (g.l2_hint) tma_load_2d_pair_hint(dst, &tma_a, &full_bar[s], kb * BK, m_blk * BM, hint_a);
                        else tma_load_2d_pair(dst, &tma_a, &full_bar[s], kb * BK, m_blk * BM);
                    }
                    if (g.l2_hint) {
                        if (g.w_blocked) tma_load_4d_pair_hint(smem_b + s * PAIR_HALF_B_BYTES, &tma_b, &full_bar[s], 0, 0, kb, t.n_blk * 2 + rank, hint_w);
                        else tma_load_2d_pair_hint(smem_b + s * PAIR_HALF_B_BYTES, &tma_b, &full_bar[s], kb * BK, t.n_blk * BN + rank * (BN / 2), hint_w);
                    } else {
                        if (g.w_blocked) tma_load_4d_pair(smem_b + s * PAIR_HALF_B_BYTES, &tma_b, &full_bar[s], 0, 0, kb, t.n_blk * 2 + rank);
                        else tma_load_2d_pair(smem_b + s * PAIR_HALF_B_BYTES, &tma_b, &full_bar[s], kb * BK, t.n_blk * BN + rank * (BN / 2));
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
            int s = 0, as = 0;
            uint32_t ph = 0, aph = 0;
            int tile_no = 0;
            for (int unit = pair; unit < units; unit += n_pairs, ++tile_no) {
                mbar_wait(&tempty_bar[as], aph ^ 1);             // both CTAs' epilogues have drained this accumulator stage
                tc_fence_after();
                PAIR_STAMP(tile_no, 0);
                const uint32_t d_tmem = tmem_base + (MT == 1 ? as * BN : 0);
                for (int kb = 0; kb < total_kb; ++kb) {
                    mbar_wait(&full_bar[s], ph);                 // both CTAs' halves of the stage have landed
                    tc_fence_after();
                    if (kb == 0) PAIR_STAMP(tile_no, 1);
                    const uint64_t b_desc = umma_desc_k_sw128(smem_u32(smem_b + s * PAIR_HALF_B_BYTES));
#pragma unroll
                    for (int sub = 0; sub < MT; ++sub) {         // MT = 2: the W k-block feeds both row sub-tiles (accumulators sub · 256)
                        const uint64_t a_desc = umma_desc_k_sw128(smem_u32(smem_a + s * A_BYTES + sub * A_STAGE_BYTES));
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_bf16_pair(d_tmem + sub * BN, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit_pair(&empty_bar[s], 3);          // frees the ring slot in both CTAs when the MMAs retire
                    if (kb == total_kb - 1) {
                        umma_commit_pair(&tfull_bar[as], 3);
                        PAIR_STAMP(tile_no, 2);
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                if (++as == ACC_STAGES) { as = 0; aph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
        pdl_wait();
        const int ew = warp - 4;
        const int q = ew & 3;
        const int hsel = ew >> 2;
        uint8_t* stg = staging + ew * 4096 * EPI_BUFS;
        uint64_t* rbars = &res_bar[ew * EPI_BUFS];
        uint32_t nchunk = 0;
        int as = 0;
        uint32_t aph = 0;
        int tile_no = 0;
        for (int unit = pair; unit < units; unit += n_pairs, ++tile_no) {
            const PairTile t = pair_tile(unit, num_m2, num_n, g.sk_q, g.group_n, g.raster);
#ifdef TEO_PAIR_TRACE
            if (threadIdx.x == 128) g_pair_trace_tile_no[blockIdx.x] = tile_no;
#endif
            const uint32_t lanes = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
            // MT = 1: this accumulator stage's 128 x 256 tile; MT = 2: the two row sub-tiles one after the other (accumulators at columns
            // 0 and 256), handed back after the LAST read of the second.  The residual of the first chunk is requested before the wait.
            staged_epilogue<BN, MT>(g, &tma_c, &tma_r, lanes + (MT == 1 ? as * BN : 0), static_cast<uint32_t>(BN), (t.m2 * MT) * 2 + rank, 2, t.n_blk,
                                    stg, rbars, nchunk, lane, q, hsel,
                                    [&] {
                                        // (One polling warp + bar.sync for the other seven, with a suspend-time hint on the wait, was built and
                                        // A/B-measured: 3/4 fewer instructions executed, no change in sustained rate — scripts/gpu_r02_epi6.sh.)
                                        mbar_wait(&tfull_bar[as], aph);
                                        tc_fence_after();
                                        if (ew == 0 && lane == 0) PAIR_STAMP(tile_no, 3);
                                    },
                                    [&] {
                                        tc_fence_before();
                                        mbar_arrive_leader(&tempty_bar[as]);
                                        if (ew == 0 && lane == 0) PAIR_STAMP(tile_no, 4);
                                    });
            if (ew == 0 && lane == 0) PAIR_STAMP(tile_no, 5);
            if (++as == ACC_STAGES) { as = 0; aph ^= 1; }
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    cluster_sync_all();                  // neither CTA leaves while the peer may still read its shared memory or signal it
    if (warp == 2) tmem_dealloc_pair<PAIR_TMEM_COLS>(tmem_base);
}

bool gemm_pair_enabled() {
    static const bool on = [] {
        const char* e = getenv("TEO_GEMM_PAIR");
        return !(e && e[0] == '0');
    }();
    return on;
}

// development hook (tools/pair_sweep.py): overrides of group_m, group_n, raster, l2_hint for the next launches; <= 0 / < 0 = default
static int g_pair_cfg[4] = {0, 0, -1, -1};
extern "C" void teo_dbg_pair_cfg(int group_m, int group_n, int raster, int l2_hint) {
    g_pair_cfg[0] = group_m;
    g_pair_cfg[1] = group_n;
    g_pair_cfg[2] = raster;
    g_pair_cfg[3] = l2_hint;
}

#ifdef TEO_PAIR_TRACE
extern "C" int teo_dbg_pair_trace(void* device_buffer, int max_tiles) {
    unsigned long long* p = static_cast<unsigned long long*>(device_buffer);
    if (cudaMemcpyToSymbol(g_pair_trace, &p, sizeof(p)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(g_pair_trace_tiles, &max_tiles, sizeof(int)) != cudaSuccess) return -1;
    return 0;
}
#endif

// Called by launch_gemm with the tensor maps already built: ta = A boxes of 128 rows, tb = W boxes of 128 rows (or the
// 4-D blocked map with one 128-row block per box), tc / tr = 32-row boxes of C and of the residual.
int launch_gemm_pair(teo_handle* h, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tr,
                     const GemmArgs& g, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        TEO_CUDA(cudaFuncSetAttribute(gemm_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<1>::SMEM_BYTES));
        TEO_CUDA(cudaFuncSetAttribute(gemm_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<2>::SMEM_BYTES));
        attr_set = true;
    }
    // 512-row pair tiles for long contractions over many rows; TEO_PAIR_MT=1|2 forces one, a value >= 1024 sets the K threshold (A/B)
    static const int env_mt = [] {
        const char* e = getenv("TEO_PAIR_MT");
        return e ? atoi(e) : 0;
    }();
    // Rule: MT = 2 wherever the un-overlapped epilogue of the two sub-tiles is small against the main loop: K >= 4096, i.e. all four
    // prefill GEMMs.  (Round 2, first version: only the down projection, K = 11008, gained — the residual epilogue paid one L2 round
    // trip per chunk and the SwiGLU epilogue ~20 instructions per output; with the residual prefetched a chunk ahead and
    // ex2/rcp.approx SwiGLU (gemm_common.cuh) the same-box A/B reads prefill 786 -> 766 ms for K >= 4096 against K >= 8192,
    // scripts/gpu_r02_epi3.sh.)  At K = 1024 (ViT) the 512-row tiles lose 10-30 %: TEO_PAIR_MT=1024 shows it.
    const int mt = env_mt == 1 || env_mt == 2 ? env_mt : ((g.K >= (env_mt >= 1024 ? env_mt : 4096) && g.M >= 8192) ? 2 : 1);
    const int units = ((g.M + mt * 2 * BM - 1) / (mt * 2 * BM)) * ((g.N + PAIR_BN - 1) / PAIR_BN);
    const int pairs = std::max(1, std::min(units, h->num_sms / 2));
    // Rasterisation group: W is re-read from HBM once per group of tile rows, so bigger groups mean less DRAM traffic (and
    // power — the prefill GEMMs run power-capped) until the A panels of a group no longer stay in L2 beside the W stream.
    // Measured in the full step on one box: 8 rows 869 ms of prefill, 12 849, 16 841, 20 / 24 / 32 slower again; a K-dependent
    // rule (≈ 32 MiB of A panels) was no better than the constant.  TEO_PAIR_GROUP_M=n overrides (A/B measurements).
    static const int env_group_m = [] {
        const char* e = getenv("TEO_PAIR_GROUP_M");
        return e ? std::max(1, atoi(e)) : PAIR_GROUP_M;
    }();
    static const int env_group_n = [] {
        const char* e = getenv("TEO_PAIR_GROUP_N");
        return e ? std::max(1, atoi(e)) : (1 << 20);
    }();
    static const int env_raster = [] {
        const char* e = getenv("TEO_PAIR_RASTER");
        return e ? atoi(e) : 0;
    }();
    static const int env_hint = [] {
        const char* e = getenv("TEO_PAIR_L2_HINT");
        return e ? atoi(e) : 0;
    }();
    GemmArgs ga = g;
    ga.sk_q = g_pair_cfg[0] > 0 ? g_pair_cfg[0] : std::max(1, env_group_m / mt);      // (stream-K field, unused by this kernel: carries the group size in pair tiles)
    ga.group_n = g_pair_cfg[1] > 0 ? g_pair_cfg[1] : env_group_n;
    ga.raster = g_pair_cfg[2] >= 0 ? g_pair_cfg[2] : env_raster;
    ga.l2_hint = g_pair_cfg[3] >= 0 ? g_pair_cfg[3] : env_hint;
    if (mt == 2) TEO_CUDA(launch_kc(PDL_GEMM, gemm_pair_kernel<2>, dim3(2 * pairs), dim3(GEMM_THREADS), PairCfg<2>::SMEM_BYTES, stream, ta, tb, tc, tr, ga));
    else TEO_CUDA(launch_kc(PDL_GEMM, gemm_pair_kernel<1>, dim3(2 * pairs), dim3(GEMM_THREADS), PairCfg<1>::SMEM_BYTES, stream, ta, tb, tc, tr, ga));
    TEO_LAUNCH_CHECK("gemm_pair_kernel");
    h->launches++;
    return TEO_OK;
}

}  // namespace teo
