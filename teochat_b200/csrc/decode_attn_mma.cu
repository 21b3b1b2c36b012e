// Paged decode attention, tensor-core form (head_dim 128, page_size 64) — the kernel teo_llama_decode_step runs.
//
// One new query token per sequence against its paged K/V cache (HF LlamaAttention's decode branch over the cat'ed
// cache): pure HBM streaming, 16 384·S bytes per sequence per layer.  The CUDA-core kernel in
// attention.cu spends ≈ 4 400 cycles of dependent shared-memory loads, shuffles and FMAs on every 32 KiB K/V page
// pair, so its throughput is set by resident CTAs × that latency (≈ 6.3 TB/s, and fewer-but-deeper CTAs are slower),
// not by the memory system, whose read-only ceiling measures 7.4 TB/s here (tools/hbm_read_peak.cu).  This kernel
// shrinks the per-page work to ≈ 300 warp instructions by putting both contractions on mma.sync m16n8k16 with the
// query as row 0 of the 16-row A tile (MHA: one query row per KV head, so 15/16 of the tile is idle — the tensor
// pipe has nothing else to do), K/V pages arriving as 128-byte-swizzled TMA tiles that ldmatrix reads conflict-free:
//   warp 2        producer: persistent over (sequence, head, split) items, K and V page tiles (2 × 16 KiB) into a
//                 2-stage ring, running ahead across item boundaries
//   warps 0, 1    consumers: page n of the CTA's stream belongs to warp n & 1 (= ring stage); each keeps its own
//                 online-softmax state for the item and the two are merged through shared memory at item end
// Three CTAs per SM keep up to 192 KiB of page requests in flight.  Numerics as the other attention kernels: fp32
// scores and statistics, P rounded to bf16 before P·V, fp32 accumulation; split partials go to decode_combine_kernel.
#include <stdlib.h>

#include <cmath>

#include "common.h"
#include "ptx.cuh"

namespace teo {

constexpr int DM_THREADS = 96;
constexpr int DM_HD = 128, DM_PAGE = 64;
constexpr int DM_HALF = DM_PAGE * 128;             // [64 keys][64 bf16] swizzled half tile
constexpr int DM_TILE = 2 * DM_HALF;               // one K (or V) page slice of one head: 16 KiB
constexpr int DM_STAGE = 2 * DM_TILE;              // K + V
constexpr int DM_SMEM = 2 * DM_STAGE + 2 * (DM_HD + 2) * 4 + 64 + 1024;

struct DmArgs {
    const bf16* q;
    long long ldq;
    const int* block_table;
    const int* seq_lens;
    bf16* out;
    float* o_part;
    float* ml_part;
    int max_pages, len_bias, n_heads, n_splits, n_items;
    int dbg_skip;               // measurement only (TEO_DEC_DBG_SKIP=1): consumers release pages without computing;
                                // 2: additionally the page slices arrive as two linear 16 KiB bulk copies instead of four swizzled 8 KiB boxes
    const bf16* kv;             // pool base (dbg_skip = 2 only)
    float scale_log2;
};

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ≤ 168 registers: the register file is split four ways (one slice per SM sub-partition), three resident CTAs put up to
// three warps on a slice, and 3 × 32 × 168 is what a slice holds — at 187 registers only two CTAs were resident.
__global__ void __launch_bounds__(DM_THREADS, 4)
decode_attn_mma_kernel(const __grid_constant__ CUtensorMap tkv, const DmArgs g) {
    constexpr int HD = DM_HD, PAGE = DM_PAGE, KS = HD / 16, DT = HD / 8;
    extern __shared__ uint8_t dm_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dm_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sKV = smem;                                          // [2 stages][K | V][2 halves][64][128 B]
    float* sM = reinterpret_cast<float*>(smem + 2 * DM_STAGE);    // [2 item parities][HD + 2]: warp 1's partial state
    uint64_t* full = reinterpret_cast<uint64_t*>(sM + 2 * (HD + 2));
    uint64_t* empty = full + 2;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    pdl_wait();                                  // q, the new K/V rows and seq_lens come from the previous kernels
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tkv);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        fence_barrier_init();
    }
    __syncthreads();

    // item → (sequence, head, split) and its page range; splits fastest
    auto item_range = [&](int item, int& seq, int& head, int& split, int& len, int& p0, int& p1) {
        split = item % g.n_splits;
        const int sh = item / g.n_splits;
        head = sh % g.n_heads;
        seq = sh / g.n_heads;
        len = g.seq_lens[seq] + g.len_bias;
        const int n_pages = (len + PAGE - 1) / PAGE;
        const int pps = (n_pages + g.n_splits - 1) / g.n_splits;
        p0 = split * pps;
        p1 = min(n_pages, p0 + pps);
    };

    if (warp == 2) {
        // ------------------------------------------------------------------ producer (whole warp; lane 0 issues)
        // Page indices are fetched 32 at a time, one per lane, ONE ITEM AHEAD (with the item's page range): a block_table
        // or seq_lens load in front of every TMA request would add its ≈ 1 µs to each page's round trip, and with one
        // or two requests in flight per CTA that latency, not HBM, would set the throughput.
        uint32_t n = 0;
        int seq, head, split, len, p0 = 0, p1 = 0, cur = 0;
        if (blockIdx.x < g.n_items) {
            item_range(blockIdx.x, seq, head, split, len, p0, p1);
            if (p0 + lane < p1) cur = g.block_table[static_cast<long long>(seq) * g.max_pages + p0 + lane];
        }
        for (int item = blockIdx.x; item < g.n_items; item += gridDim.x) {
            int nseq = 0, nhead = 0, nsplit, nlen, np0 = 0, np1 = 0, nxt = 0;
            if (item + static_cast<int>(gridDim.x) < g.n_items) {
                item_range(item + gridDim.x, nseq, nhead, nsplit, nlen, np0, np1);
                if (np0 + lane < np1) nxt = g.block_table[static_cast<long long>(nseq) * g.max_pages + np0 + lane];
            }
            for (int p = p0; p < p1; ++p, ++n) {
                const int k = p - p0;
                if (k > 0 && (k & 31) == 0 && p + lane < p1)          // items longer than 32 pages: next group (exposed)
                    cur = g.block_table[static_cast<long long>(seq) * g.max_pages + p + lane];
                const int page = __shfl_sync(0xffffffffu, cur, k & 31);
                if (lane == 0) {
                    const int st = n & 1;
                    mbar_wait(&empty[st], ((n >> 1) & 1) ^ 1);
                    uint8_t* dst = sKV + st * DM_STAGE;
                    const int krow = ((page * 2 + 0) * g.n_heads + head) * PAGE;
                    const int vrow = ((page * 2 + 1) * g.n_heads + head) * PAGE;
                    mbar_arrive_expect_tx(&full[st], DM_STAGE);
                    if (g.dbg_skip == 2) {
                        bulk_load_1d(dst, g.kv + static_cast<long long>(krow) * DM_HD, DM_TILE, &full[st]);
                        bulk_load_1d(dst + DM_TILE, g.kv + static_cast<long long>(vrow) * DM_HD, DM_TILE, &full[st]);
                    } else {
                        tma_load_2d(dst, &tkv, &full[st], 0, krow);
                        tma_load_2d(dst + DM_HALF, &tkv, &full[st], 64, krow);
                        tma_load_2d(dst + DM_TILE, &tkv, &full[st], 0, vrow);
                        tma_load_2d(dst + DM_TILE + DM_HALF, &tkv, &full[st], 64, vrow);
                    }
                }
                __syncwarp();
            }
            seq = nseq; head = nhead; p0 = np0; p1 = np1; cur = nxt;
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers (warp = ring stage)
    const int gq = lane >> 2, tq = lane & 3;
    struct Range { int seq, head, split, len, p0, p1; };
    auto load_q = [&](int seq, int head, uint32_t (&qf)[KS][2]) {    // A fragments of row 0 (rows 1..15 stay zero)
        const bf16* qp = g.q + static_cast<long long>(seq) * g.ldq + head * HD;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            qf[ks][0] = gq == 0 ? *reinterpret_cast<const uint32_t*>(qp + ks * 16 + 2 * tq) : 0u;
            qf[ks][1] = gq == 0 ? *reinterpret_cast<const uint32_t*>(qp + ks * 16 + 8 + 2 * tq) : 0u;
        }
    };
    Range rn{};
    if (blockIdx.x < g.n_items) item_range(blockIdx.x, rn.seq, rn.head, rn.split, rn.len, rn.p0, rn.p1);
    uint32_t n = 0;                                               // pages of this CTA's stream so far (all items)
    uint32_t item_no = 0;
    const uint8_t* sK = sKV + warp * DM_STAGE;
    const uint8_t* sV = sK + DM_TILE;
    for (int item = blockIdx.x; item < g.n_items; item += gridDim.x, ++item_no) {
        const int seq = rn.seq, head = rn.head, split = rn.split, len = rn.len, p0 = rn.p0, p1 = rn.p1;
        uint32_t qf[KS][2];
        load_q(seq, head, qf);
        if (item + static_cast<int>(gridDim.x) < g.n_items)                                  // the next item's range, one item ahead
            item_range(item + gridDim.x, rn.seq, rn.head, rn.split, rn.len, rn.p0, rn.p1);
        float o[DT][4];
#pragma unroll
        for (int i = 0; i < DT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;
        const bool w0_first = (n & 1) == 0;                        // which consumer gets the item's first, third, … page
        for (int p = p0; p < p1; ++p, ++n) {
            if ((n & 1) != static_cast<uint32_t>(warp)) continue;  // the other consumer's page
            mbar_wait(&full[warp], (n >> 1) & 1);
            const int valid = g.dbg_skip ? 0 : min(PAGE, len - p * PAGE);
            if (valid < PAGE && !g.dbg_skip) {
                // cache rows past the sequence end are uninitialised memory: P is 0 there, but 0 · NaN is not — clear V
                for (int i = valid * 16 + lane; i < PAGE * 16; i += 32) {
                    const int r = i >> 4, c = i & 15;              // row, 16-byte chunk (position inside the row is irrelevant)
                    *reinterpret_cast<uint4*>(const_cast<uint8_t*>(sV) + (c >> 3) * DM_HALF + r * 128 + ((c & 7) << 4)) = make_uint4(0, 0, 0, 0);
                }
                fence_proxy_async();                               // ordered before the TMA refill of this stage
                __syncwarp();
            }
#pragma unroll 1
            for (int sub = 0; sub * 32 < valid; ++sub) {
                // ---- S[0, 32 keys] = q · Kᵀ
                float sc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i) sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const uint32_t a[4] = {qf[ks][0], 0u, qf[ks][1], 0u};
#pragma unroll
                    for (int nt = 0; nt < 4; nt += 2) {
                        uint32_t kf[4];
                        const int r = sub * 32 + nt * 8 + (lane & 7) + ((lane >> 4) & 1) * 8;
                        const int c = 2 * ks + ((lane >> 3) & 1);
                        ldmatrix_x4(kf, sK + (c >> 3) * DM_HALF + r * 128 + (((c & 7) ^ (r & 7)) << 4));
                        mma_bf16_16816(sc[nt], a, kf[0], kf[1]);
                        mma_bf16_16816(sc[nt + 1], a, kf[2], kf[3]);
                    }
                }
                // ---- online softmax of row 0 (lanes 0-3 hold it; the other lanes carry zero rows)
                float mx = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (sub * 32 + nt * 8 + 2 * tq + c >= valid) sc[nt][c] = -INFINITY;
                        mx = fmaxf(mx, sc[nt][c]);
                    }
                }
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                const float m_new = fmaxf(m_run, mx);           // sub·32 < valid ⇒ key sub·32 is live ⇒ finite
                const float msub = m_new * g.scale_log2;
                const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run * g.scale_log2 - msub);
                m_run = m_new;
                float rs = 0.f;
                uint32_t pf[2][4];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const float p0v = exp2f(fmaf(sc[nt][0], g.scale_log2, -msub));   // 0 for masked keys
                    const float p1v = exp2f(fmaf(sc[nt][1], g.scale_log2, -msub));
                    rs += p0v + p1v;
                    pf[nt >> 1][(nt & 1) * 2] = pack_bf16x2(p0v, p1v);
                    pf[nt >> 1][(nt & 1) * 2 + 1] = 0u;             // rows 8..15: nothing
                }
                rs += __shfl_xor_sync(0xffffffffu, rs, 1);
                rs += __shfl_xor_sync(0xffffffffu, rs, 2);
                l_run = l_run * alpha + rs;
#pragma unroll
                for (int dt = 0; dt < DT; ++dt) { o[dt][0] *= alpha; o[dt][1] *= alpha; }
                // ---- O[0, :] += P · V
#pragma unroll
                for (int ks2 = 0; ks2 < 2; ++ks2) {
#pragma unroll
                    for (int dt = 0; dt < DT; dt += 2) {
                        uint32_t vf[4];
                        const int r = sub * 32 + ks2 * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                        const int c = dt + (lane >> 4);
                        ldmatrix_x4_trans(vf, sV + (c >> 3) * DM_HALF + r * 128 + (((c & 7) ^ (r & 7)) << 4));
                        mma_bf16_16816(o[dt], pf[ks2], vf[0], vf[1]);
                        mma_bf16_16816(o[dt + 1], pf[ks2], vf[2], vf[3]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[warp]);              // this warp's ring stage may be refilled
        }
        // ---- merge the two consumers' partial states (row 0 lives in lanes 0-3), write result / split partial
        float* mb = sM + (item_no & 1) * (HD + 2);
        if (warp == 1 && gq == 0) {
#pragma unroll
            for (int dt = 0; dt < DT; ++dt) *reinterpret_cast<float2*>(mb + dt * 8 + 2 * tq) = make_float2(o[dt][0], o[dt][1]);
            if (tq == 0) {
                mb[HD] = m_run;
                mb[HD + 1] = l_run;
            }
        }
        named_bar_sync(1, 64);
        if (warp == 0 && gq == 0) {
            // The item's pages alternate between the two consumers starting with whichever owns the current ring stage, so
            // WHICH warp holds the state of pages p0, p0+2, … depends on the CTA's history.  The merge is written in terms of
            // (first-page state A, second-page state B) with a fixed operation order, so that the result of an item does not
            // depend on where in a CTA's stream it ran (a sequence gives bit-identical output wherever it sits in the batch).
            const float m1 = mb[HD], l1 = mb[HD + 1];
            const float mA = w0_first ? m_run : m1, mB = w0_first ? m1 : m_run;
            const float lA = w0_first ? l_run : l1, lB = w0_first ? l1 : l_run;
            const float M = fmaxf(mA, mB);
            const float aA = (mA == -INFINITY) ? 0.f : exp2f((mA - M) * g.scale_log2);
            const float aB = (mB == -INFINITY) ? 0.f : exp2f((mB - M) * g.scale_log2);
            const float L = fmaf(lB, aB, lA * aA);
            const long long base = static_cast<long long>(seq) * g.n_heads + head;
            auto merged = [&](float x0, float x1) {               // x0: this warp's value, x1: warp 1's
                const float xA = w0_first ? x0 : x1, xB = w0_first ? x1 : x0;
                return fmaf(xB, aB, xA * aA);
            };
            if (g.n_splits == 1) {
                const float inv = 1.0f / L;
                bf16* op = g.out + base * HD + 2 * tq;
#pragma unroll
                for (int dt = 0; dt < DT; ++dt) {
                    const float2 o1 = *reinterpret_cast<const float2*>(mb + dt * 8 + 2 * tq);
                    *reinterpret_cast<uint32_t*>(op + dt * 8) = pack_bf16x2(merged(o[dt][0], o1.x) * inv, merged(o[dt][1], o1.y) * inv);
                }
            } else {
                float* op = g.o_part + (base * g.n_splits + split) * HD + 2 * tq;
#pragma unroll
                for (int dt = 0; dt < DT; ++dt) {
                    const float2 o1 = *reinterpret_cast<const float2*>(mb + dt * 8 + 2 * tq);
                    *reinterpret_cast<float2*>(op + dt * 8) = make_float2(merged(o[dt][0], o1.x), merged(o[dt][1], o1.y));
                }
                if (tq == 0) {
                    float* ml = g.ml_part + (base * g.n_splits + split) * 2;
                    ml[0] = M;          // -inf when this split had no pages
                    ml[1] = L;
                }
            }
        }
    }
}

int get_tmap_bf16(teo_handle* h, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, CUtensorMap* out);

// grid: persistent, 3 CTAs per SM (66 KiB of shared memory each)
int launch_decode_attention_mma(teo_handle* h, const bf16* q, int ldq, const bf16* kv_pages, const int* block_table, int max_pages,
                                const int* seq_lens, int len_bias, bf16* out, int n_seqs, int n_heads, int splits, float scale,
                                float* o_part, float* ml_part, bool* merged, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        TEO_CUDA(cudaFuncSetAttribute(decode_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DM_SMEM));
        attr_set = true;
    }
    // The pool is addressed as one 2-D tensor [page·2·heads·64 rows][128]: the row bound is only an out-of-bounds clip,
    // and block_table never points past the caller's pool, so a generous bound is safe without knowing the pool size.
    CUtensorMap tkv;
    TEO_TRY(get_tmap_bf16(h, kv_pages, 1ull << 31, DM_HD, DM_HD, DM_PAGE, &tkv));
    DmArgs g{};
    g.q = q;
    g.ldq = ldq;
    g.block_table = block_table;
    g.seq_lens = seq_lens;
    g.out = out;
    g.o_part = o_part;
    g.ml_part = ml_part;
    *merged = false;            // split partials are merged by decode_combine_kernel (an in-kernel last-arriver merge measured slower:
                                // its fence + atomic per item stalls the consumer warp, 10.03 vs 9.54 ms per decode step)
    g.max_pages = max_pages;
    g.len_bias = len_bias;
    g.n_heads = n_heads;
    g.n_splits = splits;
    g.n_items = splits * n_heads * n_seqs;
    g.scale_log2 = scale * 1.4426950408889634f;
    static const int dbg_skip = [] { const char* e = getenv("TEO_DEC_DBG_SKIP"); return e ? atoi(e) : 0; }();
    g.dbg_skip = dbg_skip;
    g.kv = kv_pages;
    const int grid = std::min(g.n_items, h->num_sms * 3);
    TEO_CUDA(launch_kc(PDL_ATTN, decode_attn_mma_kernel, dim3(grid), dim3(DM_THREADS), DM_SMEM, stream, tkv, g));
    TEO_LAUNCH_CHECK("decode_attn_mma_kernel");
    return TEO_OK;
}

}  // namespace teo
