// Flash attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) for the two dense attention
// shapes of the hot path:
//   ViT blocks      non-causal, 257 tokens per frame, head_dim 64   (HF CLIPAttention: softmax(q·s kᵀ) v)
//   LLaMA prefill   causal, ragged batch of ~2.1k-token sequences, head_dim 128 (HF LlamaAttention, fp32 softmax)
//
// Persistent CTAs walk a list of work items; one item = TWO adjacent 128-row query tiles of one (sequence, head),
// whose key/value blocks (128 keys at head_dim 128, 64 keys at head_dim 64) are streamed once:
//   warp 0        TMA producer: the item's Q tiles, then its K and V blocks into two mbarrier rings (128-byte swizzle;
//                 head_dim 128 = two 64-column halves per tile; tail blocks as 32-row boxes)
//   warp 1        MMA issuer (one thread): S_t = Q_t·Kᵀ (both operands K-major in shared memory) into TMEM,
//                 O_t += P_t·V with P_t read from TMEM (A operand) and V MN-major straight from its TMA tile
//   warp 2        TMEM allocator (S0 | S1 | O0 | O1; P_t overlays the first half of S_t)
//   warp 3        q_offset = 1 only: row 0 of the sequence (the ViT CLS query) on CUDA cores from the same K/V tiles
//   warps 4-7     softmax of tile 0, one thread per query row (= TMEM lane): row max, lazy rescale of O
//   warps 8-11    softmax of tile 1        (only when the running max grows by > 2^8), exp2, bf16 P → TMEM
// The two tiles ping-pong: while one tile's rows are in softmax, the tensor core runs the other tile's
// P·V and next Q·Kᵀ.  tcgen05.mma retires in issue order, so the commit that publishes S_t(j+1) also
// guarantees P_t(j)·V(j) is complete — the softmax warps may rescale O_t right after that wait, and the
// epilogue's read of O_t is ordered before the next item's first P·V by the p_full arrival that follows it.
//
// Numerics match flash_fwd_kernel (attention.cu): fp32 scores and statistics, P rounded to bf16 before P·V,
// fp32 accumulation, one rounding of the output.  Algorithmic work 4·S²·d flop (half if causal).
#include <cmath>

#include "common.h"
#include "ptx.cuh"

namespace teo {

constexpr int FA_THREADS = 384;
constexpr int FA_TILE_BYTES = 128 * 128;      // [128 rows][64 bf16] under the 128-byte swizzle
constexpr float FA_RESCALE_LOG2 = 8.0f;       // O is rescaled only when the row max grew by more than 2^8

// Tile configuration per head_dim.  head_dim 128: 128-key blocks, one CTA per SM (TMEM 2·128 + 2·128 = 512 columns,
// 192 KiB shared memory), the score row of a block held in registers.  head_dim 64 (ViT): 64-key blocks, so that TWO
// CTAs fit on an SM (TMEM 2·64 + 2·64 = 256 columns each, 80 KiB shared memory, ≤ 80 registers) — four query tiles in
// flight per SM keep the MUFU (exp2) pipe busy across the softmax → MMA → softmax round trips of each tile.
template <int HD>
struct FaCfg {
    static constexpr int BN = HD == 128 ? 128 : 64;             // keys per block
    static constexpr int CTAS_PER_SM = HD == 128 ? 1 : 2;
    static constexpr bool REG_RESIDENT = HD == 128;             // whole score row in registers (one TMEM read)
    static constexpr int HALVES = HD / 64;
    static constexpr int HALF_BYTES = BN * 128;                 // [BN keys][64 bf16] under the 128-byte swizzle
    static constexpr int KV_TILE = HALVES * HALF_BYTES;
    static constexpr int Q_SET = 2 * HALVES * FA_TILE_BYTES;    // both query tiles of one work item
    static constexpr int Q_SETS = 1;
    static constexpr int STAGES = HD == 128 ? 2 : 3;
    static constexpr int TMEM_COLS = (2 * BN + 2 * HD <= 256) ? 256 : 512;
    static constexpr int SMEM = Q_SETS * Q_SET + 2 * STAGES * KV_TILE + 1024 /*align*/ + 256 /*barriers*/;
};

struct FaArgs {
    bf16* O;
    long long ldo;
    const bf16* Q;              // raw query pointer / row stride: the q_offset row is read directly (not through TMA)
    long long ldq;
    const int* cu_seqlens;
    int n_heads, n_sh;          // n_sh = sequences × heads
    int n_pairs;                // 256-row query groups per sequence; work items = n_pairs × n_sh
    int q_offset;               // 0 | 1: row 0 of every sequence is not tiled but computed by warp 3 from the K/V tiles
                                // in shared memory (ViT: 257 tokens = CLS row + exactly two 128-row query tiles)
    int chunk;                  // (sequence, head) pairs scheduled together, heavy causal groups first
    float scale_log2;
};

// One work item = two adjacent 128-row query tiles of one (sequence, head).  Items are ordered in chunks of
// (sequence, head) pairs; inside a chunk the heaviest (last) query groups come first, so the K/V of a chunk stay
// in L2 while the persistent CTAs (item i → CTA i mod grid) each get a similar mix of heavy and light items.
struct FaItem {
    int head, seq_start, seqlen, nq, m0, n_tiles, nb[2], nb_max;
    bool valid;
};
template <bool CAUSAL, int BN>
__device__ __forceinline__ FaItem fa_decode(const FaArgs& g, int idx) {
    FaItem it;
    const int per = g.chunk * g.n_pairs;
    const int ch = idx / per, rem = idx - ch * per;
    const int csz = min(g.chunk, g.n_sh - ch * g.chunk);
    const int pair = g.n_pairs - 1 - rem / csz;
    const int sh = ch * g.chunk + rem % csz;
    const int seq = sh / g.n_heads;
    it.head = sh - seq * g.n_heads;
    it.seq_start = g.cu_seqlens[seq];
    it.seqlen = g.cu_seqlens[seq + 1] - it.seq_start;
    it.nq = it.seqlen - g.q_offset;
    it.m0 = pair * 256;
    it.valid = it.m0 < it.nq || (g.q_offset == 1 && it.m0 == 0 && it.seqlen > 0);   // a 1-token sequence is only its row 0
    it.n_tiles = (it.nq - it.m0 > 128) ? 2 : (it.nq - it.m0 > 0 ? 1 : 0);
    const int nb_all = (it.seqlen + BN - 1) / BN;
    it.nb[0] = CAUSAL ? min(nb_all, (it.m0 + 127) / BN + 1) : nb_all;          // blocks up to the tile's last row
    it.nb[1] = it.n_tiles < 2 ? 0 : (CAUSAL ? min(nb_all, (it.m0 + 255) / BN + 1) : nb_all);
    it.nb_max = max(it.nb[0], it.nb[1]);
    return it;
}
template <int BN>
__device__ __forceinline__ int fa_block_cols(int seqlen, int j) { return min(BN, (seqlen - j * BN + 31) & ~31); }

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(FA_THREADS, FaCfg<HD>::CTAS_PER_SM)
flash_tc_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk, const __grid_constant__ CUtensorMap tv,
                const __grid_constant__ CUtensorMap tk32, const __grid_constant__ CUtensorMap tv32, const FaArgs g) {
    using Cfg = FaCfg<HD>;
    constexpr int HALVES = Cfg::HALVES, STAGES = Cfg::STAGES, KV_TILE = Cfg::KV_TILE, Q_SETS = Cfg::Q_SETS, BN = Cfg::BN;
    constexpr int HALF_BYTES = Cfg::HALF_BYTES;
    constexpr uint32_t TM_S = 0, TM_O = 2 * BN;                 // TMEM column bases: S_t at BN·t, O_t at 2·BN + HD·t
    const int n_items = g.n_sh * g.n_pairs;

    extern __shared__ uint8_t fa_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fa_smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                                         // [Q_SETS][2 tiles][HALVES][128][64]
    uint8_t* sK = sQ + Q_SETS * Cfg::Q_SET;                     // [STAGES][HALVES][128][64]
    uint8_t* sV = sK + STAGES * KV_TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + STAGES * KV_TILE);
    uint64_t* q_full = bars;                                    // [Q_SETS]
    uint64_t* q_empty = q_full + Q_SETS;
    uint64_t* k_full = q_empty + Q_SETS;
    uint64_t* k_empty = k_full + STAGES;
    uint64_t* v_full = k_empty + STAGES;
    uint64_t* v_empty = v_full + STAGES;
    uint64_t* s_full = v_empty + STAGES;                        // [2] S_t(j) complete (MMA → softmax)
    uint64_t* p_full = s_full + 2;                              // [2] P_t(j) in TMEM, O_t rescaled (softmax → MMA)
    uint64_t* o_full = p_full + 2;                              // [2] last P·V of tile t complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tq);
        tma_prefetch_desc(&tk);
        tma_prefetch_desc(&tv);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < Q_SETS; ++s) {
            mbar_init(&q_full[s], 1);
            mbar_init(&q_empty[s], 1);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1 + g.q_offset);           // MMA commit (+ the row-0 warp)
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1 + g.q_offset);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(&s_full[t], 1);
            mbar_init(&p_full[t], 128);
            mbar_init(&o_full[t], 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t kv_cnt = 0, q_cnt = 0;
            for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
                const FaItem it = fa_decode<CAUSAL, BN>(g, idx);
                if (!it.valid) continue;
                const int col0 = it.head * HD;
                const int qs = q_cnt % Q_SETS;
                mbar_wait(&q_empty[qs], ((q_cnt / Q_SETS) & 1) ^ 1);
                mbar_arrive_expect_tx(&q_full[qs], it.n_tiles * HALVES * FA_TILE_BYTES);
                for (int t = 0; t < it.n_tiles; ++t)
                    for (int h = 0; h < HALVES; ++h)
                        tma_load_2d(sQ + qs * Cfg::Q_SET + (t * HALVES + h) * FA_TILE_BYTES, &tq, &q_full[qs], col0 + 64 * h,
                                    it.seq_start + g.q_offset + it.m0 + 128 * t);
                ++q_cnt;
                auto load_block = [&](uint8_t* dst, const CUtensorMap* big, const CUtensorMap* small, uint64_t* bar, int j) {
                    const int cols = fa_block_cols<BN>(it.seqlen, j), row = it.seq_start + j * BN;
                    if (cols == BN) {
                        mbar_arrive_expect_tx(bar, KV_TILE);
                        for (int h = 0; h < HALVES; ++h) tma_load_2d(dst + h * HALF_BYTES, big, bar, col0 + 64 * h, row);
                    } else {                                     // tail block: 32-row boxes, only the rows in use
                        const int n32 = cols >> 5;
                        mbar_arrive_expect_tx(bar, HALVES * n32 * 4096);
                        for (int h = 0; h < HALVES; ++h)
                            for (int i = 0; i < n32; ++i)
                                tma_load_2d(dst + h * HALF_BYTES + i * 4096, small, bar, col0 + 64 * h, row + 32 * i);
                    }
                };
                for (int j = 0; j < it.nb_max; ++j, ++kv_cnt) {
                    const int s = kv_cnt % STAGES;
                    const uint32_t ph = (kv_cnt / STAGES) & 1;
                    mbar_wait(&k_empty[s], ph ^ 1);
                    load_block(sK + s * KV_TILE, &tk, &tk32, &k_full[s], j);
                    mbar_wait(&v_empty[s], ph ^ 1);
                    load_block(sV + s * KV_TILE, &tv, &tv32, &v_full[s], j);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        // The WHOLE warp walks the schedule in convergent control flow and only the tcgen05 instructions are issued by one elected
        // lane.  Every value the descriptors are built from is made provably warp-uniform (item fields are broadcast from lane 0),
        // so the compiler keeps descriptors and loop state in uniform registers.  Issued from inside `if (lane == 0)` each of the
        // 16 MMAs of a (tile, key block) cost a ≈ 16-instruction ELECT / R2UR.BROADCAST / BRA.U.ANY loop, and the single issuing
        // thread — not the tensor pipe, not the softmax — bounded the kernel: with the softmax removed it ran at 842 TFLOP/s
        // (`profiles/r02_flash_issue.txt`).
        {
            constexpr uint32_t idesc_pv = umma_idesc_bf16(128, HD) | UMMA_IDESC_B_MN_MAJOR;
            const uint32_t sQ_u = smem_u32(sQ), sK_u = smem_u32(sK), sV_u = smem_u32(sV);
            uint32_t kv_cnt = 0, q_cnt = 0, n_t[2] = {0, 0};
            for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
                FaItem it = fa_decode<CAUSAL, BN>(g, idx);
                it.valid = __shfl_sync(0xffffffffu, static_cast<int>(it.valid), 0) != 0;
                if (!it.valid) continue;
                it.seqlen = __shfl_sync(0xffffffffu, it.seqlen, 0);
                it.n_tiles = __shfl_sync(0xffffffffu, it.n_tiles, 0);
                it.nb[0] = __shfl_sync(0xffffffffu, it.nb[0], 0);
                it.nb[1] = __shfl_sync(0xffffffffu, it.nb[1], 0);
                it.nb_max = max(it.nb[0], it.nb[1]);
                const int qs = q_cnt % Q_SETS;
                const uint32_t sQi = sQ_u + qs * Cfg::Q_SET;
                auto issue_s = [&](int t, int stage, int cols) {           // S_t = Q_t · K(stage)ᵀ   [128 × cols]
                    const uint32_t idesc = umma_idesc_bf16(128, cols);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < HD / 16; ++ks) {
                            const uint64_t a = umma_desc_k_sw128(sQi + (t * HALVES + ks / 4) * FA_TILE_BYTES) + 2 * (ks & 3);
                            const uint64_t b = umma_desc_k_sw128(sK_u + stage * KV_TILE + (ks / 4) * HALF_BYTES) + 2 * (ks & 3);
                            umma_bf16(tmem_base + TM_S + BN * t, a, b, idesc, ks > 0 ? 1u : 0u);
                        }
                    }
                    __syncwarp();
                };
                auto issue_pv = [&](int t, int stage, int cols, bool first) {   // O_t (+)= P_t · V(stage)   [128 × HD]
                    const uint32_t vbase = sV_u + stage * KV_TILE;
                    if (elect_one()) {
                        if (cols == BN) {
#pragma unroll
                            for (int ks = 0; ks < BN / 16; ++ks) {
                                const uint64_t b = umma_desc_mn_sw128(vbase + ks * 2048, HALF_BYTES, 1024);
                                umma_bf16_ts(tmem_base + TM_O + HD * t, tmem_base + TM_S + BN * t + 8 * ks, b, idesc_pv, (first && ks == 0) ? 0u : 1u);
                            }
                        } else {
                            for (int ks = 0; ks < cols / 16; ++ks) {
                                const uint64_t b = umma_desc_mn_sw128(vbase + ks * 2048, HALF_BYTES, 1024);
                                umma_bf16_ts(tmem_base + TM_O + HD * t, tmem_base + TM_S + BN * t + 8 * ks, b, idesc_pv, (first && ks == 0) ? 0u : 1u);
                            }
                        }
                    }
                    __syncwarp();
                };
                auto commit = [&](uint64_t* bar) {
                    if (elect_one()) umma_commit(bar);
                    __syncwarp();
                };
                mbar_wait(&q_full[qs], (q_cnt / Q_SETS) & 1);
                mbar_wait(&k_full[kv_cnt % STAGES], (kv_cnt / STAGES) & 1);
                tc_fence_after();
                for (int t = 0; t < it.n_tiles; ++t) {
                    issue_s(t, kv_cnt % STAGES, fa_block_cols<BN>(it.seqlen, 0));
                    commit(&s_full[t]);
                }
                commit(&k_empty[kv_cnt % STAGES]);
                if (it.nb_max == 1) commit(&q_empty[qs]);
                for (int j = 0; j < it.nb_max; ++j) {
                    const uint32_t cv = kv_cnt + j, ck = cv + 1;
                    const int sv = cv % STAGES, sk = ck % STAGES;
                    const int cols = fa_block_cols<BN>(it.seqlen, j);
                    bool v_ready = false;
                    for (int t = 0; t < it.n_tiles; ++t) {
                        if (j >= it.nb[t]) continue;
                        mbar_wait(&p_full[t], n_t[t] & 1);
                        ++n_t[t];
                        if (!v_ready) {
                            mbar_wait(&v_full[sv], (cv / STAGES) & 1);
                            v_ready = true;
                        }
                        tc_fence_after();
                        issue_pv(t, sv, cols, j == 0);
                        if (j == it.nb[t] - 1) commit(&o_full[t]);
                        if (j + 1 < it.nb[t]) {
                            mbar_wait(&k_full[sk], (ck / STAGES) & 1);
                            tc_fence_after();
                            issue_s(t, sk, fa_block_cols<BN>(it.seqlen, j + 1));
                            commit(&s_full[t]);
                        }
                    }
                    commit(&v_empty[sv]);
                    if (j + 1 < it.nb_max) commit(&k_empty[sk]);
                    if (j + 2 == it.nb_max) commit(&q_empty[qs]);     // the item's last Q·Kᵀ has been issued
                }
                kv_cnt += it.nb_max;
                ++q_cnt;
            }
        }
        __syncwarp();
    } else if (warp == 3 && g.q_offset == 1) {
        // ------------------------------------------------------------------ row 0 of every sequence (one warp, mma.sync)
        // A 16-row flash-attention warp whose only live row is the sequence's row 0: walks the same K/V ring as the
        // MMA warp, 32 keys at a time — q·kᵀ and p·v on mma.sync m16n8k16 with ldmatrix straight from the 128-byte-
        // swizzled TMA tiles, online softmax in registers.  ~130 instructions per 32 keys, far off the critical path.
        // Every block of every item is acknowledged on k_empty / v_empty (count 2), also for items whose row 0
        // belongs to another item (m0 > 0), so that the barrier phases stay in step.
        // (Round 2: splitting the row over warps 2 and 3 — alternate 32-key slices, states merged through shared memory — was
        // built and measured: 284.9 vs 282.7 µs per ViT layer, no gain, reverted; this warp is not what the 257th token costs.)
        constexpr int KS = HD / 16, DT = HD / 8;
        const int gq = lane >> 2, tq = lane & 3;
        uint32_t kv_cnt = 0;
        for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
            const FaItem it = fa_decode<CAUSAL, BN>(g, idx);
            if (!it.valid) continue;
            const bool mine = it.m0 == 0;
            uint32_t qf[KS][2];                                   // A fragments of row 0 (rows 1..15 are zero)
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) qf[ks][0] = qf[ks][1] = 0u;
            if (mine && gq == 0) {
                const bf16* qp = g.Q + static_cast<long long>(it.seq_start) * g.ldq + it.head * HD;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    qf[ks][0] = *reinterpret_cast<const uint32_t*>(qp + ks * 16 + 2 * tq);
                    qf[ks][1] = *reinterpret_cast<const uint32_t*>(qp + ks * 16 + 8 + 2 * tq);
                }
            }
            float o[DT][4];
#pragma unroll
            for (int i = 0; i < DT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
            float m_run = -INFINITY, l_run = 0.f;
            for (int j = 0; j < it.nb_max; ++j, ++kv_cnt) {
                const int st = kv_cnt % STAGES;
                const uint32_t ph = (kv_cnt / STAGES) & 1;
                const int kv0 = j * BN;
                const int n_valid = min(BN, it.seqlen - kv0);
                mbar_wait(&k_full[st], ph);
                mbar_wait(&v_full[st], ph);
                if (mine) {
                    const uint8_t* sKb = sK + st * KV_TILE;
                    const uint8_t* sVb = sV + st * KV_TILE;
                    for (int sub = 0; sub * 32 < n_valid; ++sub) {
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
                                ldmatrix_x4(kf, sKb + (c >> 3) * HALF_BYTES + r * 128 + (((c & 7) ^ (r & 7)) << 4));
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
                                if (sub * 32 + nt * 8 + 2 * tq + c >= n_valid) sc[nt][c] = -INFINITY;
                                mx = fmaxf(mx, sc[nt][c]);
                            }
                        }
                        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                        const float m_new = fmaxf(m_run, mx);           // sub·32 < n_valid ⇒ key sub·32 is live ⇒ finite
                        const float msub = m_new * g.scale_log2;
                        const float alpha = (m_run == -INFINITY) ? 0.f : ex2_approx(m_run * g.scale_log2 - msub);
                        m_run = m_new;
                        float rs = 0.f;
                        uint32_t pf[2][4];
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) {
                            const float p0 = ex2_approx(fmaf(sc[nt][0], g.scale_log2, -msub));   // 0 for masked keys
                            const float p1 = ex2_approx(fmaf(sc[nt][1], g.scale_log2, -msub));
                            rs += p0 + p1;
                            pf[nt >> 1][(nt & 1) * 2] = pack_bf16x2(p0, p1);
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
                                ldmatrix_x4_trans(vf, sVb + (c >> 3) * HALF_BYTES + r * 128 + (((c & 7) ^ (r & 7)) << 4));
                                mma_bf16_16816(o[dt], pf[ks2], vf[0], vf[1]);
                                mma_bf16_16816(o[dt + 1], pf[ks2], vf[2], vf[3]);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&k_empty[st]);
                    mbar_arrive(&v_empty[st]);
                }
            }
            if (mine && gq == 0) {
                const float inv = 1.0f / l_run;
                bf16* op = g.O + static_cast<long long>(it.seq_start) * g.ldo + it.head * HD + 2 * tq;
#pragma unroll
                for (int dt = 0; dt < DT; ++dt) *reinterpret_cast<uint32_t*>(op + dt * 8) = pack_bf16x2(o[dt][0] * inv, o[dt][1] * inv);
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ softmax + epilogue (thread = query row)
        const int t = (warp - 4) >> 2;
        const int quad = warp & 3;                               // TMEM lane quadrant this warp may access
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        const uint32_t t_s = lane_base + TM_S + BN * t;
        const uint32_t t_o = lane_base + TM_O + HD * t;
        uint32_t n_blk = 0, n_out = 0;                           // phase counters of s_full[t] / o_full[t]
        for (int idx = blockIdx.x; idx < n_items; idx += gridDim.x) {
            const FaItem it = fa_decode<CAUSAL, BN>(g, idx);
            if (!it.valid || t >= it.n_tiles) continue;
            const int seqlen = it.seqlen;
            const int row = it.m0 + 128 * t + quad * 32 + lane;  // among the tiled rows of this sequence
            const int qpos = g.q_offset + row;                   // position in the sequence (causal mask)
            float m_used = -INFINITY, l_run = 0.f;
            const int nbt = it.nb[t];
            for (int j = 0; j < nbt; ++j, ++n_blk) {
                const int cols = fa_block_cols<BN>(seqlen, j);
                const int kv0 = j * BN;
                // keys past the sequence end, or (causal) past the first row of this tile: element-wise mask needed
                const bool need_mask = (kv0 + cols > seqlen) || (CAUSAL && kv0 + cols - 1 > it.m0 + 128 * t);
                mbar_wait(&s_full[t], n_blk & 1);
                tc_fence_after();
                // The element-wise mask must stay a (warp-uniform) BRANCH: only the one or two blocks per query tile on the diagonal or
                // at the sequence end need it.  Written as plain conditional assignments the compiler if-converted the compares and
                // selects into EVERY block (≈ 5 of 10 instructions per score element); the empty asm statements forbid that.
                const int lim = (CAUSAL ? min(seqlen - 1, qpos) : seqlen - 1) - kv0;     // last live column of this row in the block
#if defined(TEO_FA_DBG) && TEO_FA_DBG == 3      // timing decomposition builds only (results are garbage): no softmax at all
                if (lim > -1000000) {
                    tc_fence_before();
                    mbar_arrive(&p_full[t]);
                    continue;
                }
#endif
                constexpr int NV = Cfg::REG_RESIDENT ? BN : 32;
                uint32_t v[NV];
                // chunk c (32 score columns) → registers, masked; REG_RESIDENT keeps all chunks live for the second pass
                auto fetch = [&](int c, uint32_t(&x)[32]) {
                    tmem_ld_32x32(t_s + c * 32, x);
                };
                auto mask = [&](int c, uint32_t(&x)[32]) {
                    if (need_mask) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            if (c * 32 + i > lim) x[i] = 0xFF800000u;   // -inf: past the sequence end or (causal) past this row
                            asm volatile("" : "+r"(x[i]));              // keeps this a branch (see above)
                        }
                    }
                };
                float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                auto rowmax = [&](const uint32_t(&x)[32]) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
#pragma unroll
                        for (int a = 0; a < 4; ++a)
                            mx4[a] = fmaxf(mx4[a], fmaxf(__uint_as_float(x[i + 2 * a]), __uint_as_float(x[i + 2 * a + 1])));
                    }
                };
                // ---- pass 1: row max of the block
                if constexpr (Cfg::REG_RESIDENT) {
#pragma unroll
                    for (int c = 0; c < BN / 32; ++c)
                        if (c * 32 < cols) fetch(c, reinterpret_cast<uint32_t(&)[32]>(v[c * 32]));
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < BN / 32; ++c) {
                        if (c * 32 < cols) {
                            mask(c, reinterpret_cast<uint32_t(&)[32]>(v[c * 32]));
                            rowmax(reinterpret_cast<uint32_t(&)[32]>(v[c * 32]));
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < BN / 32; ++c) {
                        if (c * 32 < cols) {
                            fetch(c, reinterpret_cast<uint32_t(&)[32]>(v[0]));
                            tmem_ld_wait();
                            mask(c, reinterpret_cast<uint32_t(&)[32]>(v[0]));
                            rowmax(reinterpret_cast<uint32_t(&)[32]>(v[0]));
                        }
                    }
                }
                const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
                // ---- lazy rescale of O (warp-uniform decision: tcgen05.ld/st are warp-collective)
                if (j == 0) {
                    m_used = mx;
                } else {
                    const bool grow = (mx - m_used) * g.scale_log2 > FA_RESCALE_LOG2;
                    if (__any_sync(0xffffffffu, grow)) {
                        const float m_new = fmaxf(m_used, mx);
                        const float alpha = ex2_approx((m_used - m_new) * g.scale_log2);
                        m_used = m_new;
                        l_run *= alpha;
#pragma unroll 1
                        for (int c = 0; c < HD; c += 32) {
                            uint32_t o[32];
                            tmem_ld_32x32(t_o + c, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st_32x32(t_o + c, o);
                        }
                    }
                }
                const float neg_ms = -m_used * g.scale_log2;
                // ---- pass 2: P = exp2(S·scale − m·scale) → bf16 pairs → TMEM, over the columns of S already consumed
                float rs4[4] = {0.f, 0.f, 0.f, 0.f};
                auto expstore = [&](int c, const uint32_t(&x)[32]) {
                    uint32_t pk[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
#if defined(TEO_FA_DBG) && TEO_FA_DBG == 1      // no MUFU
                        const float p0 = fmaf(__uint_as_float(x[i]), g.scale_log2, neg_ms);
                        const float p1 = fmaf(__uint_as_float(x[i + 1]), g.scale_log2, neg_ms);
                        rs4[(i >> 1) & 3] += p0 + p1;
                        pk[i >> 1] = pack_bf16x2(p0, p1);
#elif defined(TEO_FA_DBG) && TEO_FA_DBG == 2    // no pass-2 arithmetic at all: TMEM traffic + pass 1 only
                        pk[i >> 1] = x[i] ^ x[i + 1];
#else
                        const float p0 = ex2_approx(fmaf(__uint_as_float(x[i]), g.scale_log2, neg_ms));
                        const float p1 = ex2_approx(fmaf(__uint_as_float(x[i + 1]), g.scale_log2, neg_ms));
                        rs4[(i >> 1) & 3] += p0 + p1;
                        pk[i >> 1] = pack_bf16x2(p0, p1);
#endif
                    }
                    tmem_st_32x16(t_s + c * 16, pk);
                };
#pragma unroll
                for (int c = 0; c < BN / 32; ++c) {
                    if (c * 32 < cols) {
                        if constexpr (Cfg::REG_RESIDENT) {
                            expstore(c, reinterpret_cast<uint32_t(&)[32]>(v[c * 32]));
                        } else {
                            fetch(c, reinterpret_cast<uint32_t(&)[32]>(v[0]));
                            tmem_ld_wait();
                            mask(c, reinterpret_cast<uint32_t(&)[32]>(v[0]));
                            expstore(c, reinterpret_cast<uint32_t(&)[32]>(v[0]));
                        }
                    }
                }
                l_run += (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_full[t]);
            }
            // ---- epilogue: O / l → bf16 → global (rows of this sequence only)
            mbar_wait(&o_full[t], n_out & 1);
            ++n_out;
            tc_fence_after();
            const bool row_ok = row < it.nq;
            const float inv = 1.0f / l_run;
            bf16* dst = g.O + static_cast<long long>(it.seq_start + qpos) * g.ldo + it.head * HD;
#pragma unroll 1
            for (int c = 0; c < HD; c += 32) {
                uint32_t o[32];
                tmem_ld_32x32(t_o + c, o);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        *reinterpret_cast<uint4*>(dst + c + i) =
                            make_uint4(pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv),
                                       pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv),
                                       pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv),
                                       pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

template <int HD, bool CAUSAL>
static int launch_flash_tc_t(teo_handle* h, const bf16* q, int ldq, const bf16* k, int ldk, const bf16* v, int ldv, bf16* out, int ldo,
                             const int* cu, int n_seqs, int max_seqlen, int total_tokens, int n_heads, float scale, int q_offset,
                             cudaStream_t stream) {
    using Cfg = FaCfg<HD>;
    static bool attr_set = false;
    if (!attr_set) {
        TEO_CUDA(cudaFuncSetAttribute(flash_tc_kernel<HD, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_set = true;
    }
    CUtensorMap tq, tk, tv, tk32, tv32;
    const uint64_t cols = static_cast<uint64_t>(n_heads) * HD;
    TEO_TRY(get_tmap_bf16(h, q, total_tokens, cols, ldq, 128, &tq));
    TEO_TRY(get_tmap_bf16(h, k, total_tokens, cols, ldk, Cfg::BN, &tk));
    TEO_TRY(get_tmap_bf16(h, v, total_tokens, cols, ldv, Cfg::BN, &tv));
    TEO_TRY(get_tmap_bf16(h, k, total_tokens, cols, ldk, 32, &tk32));
    TEO_TRY(get_tmap_bf16(h, v, total_tokens, cols, ldv, 32, &tv32));
    FaArgs g{};
    g.O = out;
    g.ldo = ldo;
    g.Q = q;
    g.ldq = ldq;
    g.cu_seqlens = cu;
    g.n_heads = n_heads;
    g.n_sh = n_seqs * n_heads;
    g.n_pairs = (max_seqlen - q_offset + 255) / 256;
    g.q_offset = q_offset;
    g.chunk = std::min(g.n_sh, 64);          // ≈ 64 × (K+V of one head) stays well inside the 126 MB L2
    g.scale_log2 = scale * 1.4426950408889634f;
    const long long items = static_cast<long long>(g.n_sh) * g.n_pairs;
    TEO_CHECK_ARG(items < (1LL << 30), "flash_attention_tc: too many work items");
    const int grid = static_cast<int>(std::min<long long>(items, h->num_sms * Cfg::CTAS_PER_SM));     // persistent CTAs
    flash_tc_kernel<HD, CAUSAL><<<grid, FA_THREADS, Cfg::SMEM, stream>>>(tq, tk, tv, tk32, tv32, g);
    TEO_LAUNCH_CHECK("flash_tc_kernel");
    h->launches++;
    return TEO_OK;
}

// q_offset = 1: row 0 of every sequence is computed by a spare warp of the same CTA from the K/V tiles in shared memory
// instead of a (nearly empty) third 128-row tile: the ViT's 257 tokens are the CLS row + exactly two query tiles.
int launch_flash_attention_tc(teo_handle* h, const bf16* q, int ldq, const bf16* k, int ldk, const bf16* v, int ldv, bf16* out, int ldo,
                              const int* cu_seqlens, int n_seqs, int max_seqlen, int total_tokens, int n_heads, int head_dim,
                              float scale, int causal, int q_offset, cudaStream_t stream) {
    TEO_CHECK_ARG(h && q && k && v && out && cu_seqlens, "flash_attention_tc: null pointer");
    TEO_CHECK_ARG(n_seqs > 0 && max_seqlen > 0 && n_heads > 0 && total_tokens > 0, "flash_attention_tc: bad sizes");
    TEO_CHECK_ARG((q_offset == 0 || q_offset == 1) && q_offset < max_seqlen && !(causal && q_offset),
                  "flash_attention_tc: q_offset must be 0 or 1 (and 0 when causal), got %d", q_offset);
    TEO_CHECK_ARG(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "flash_attention_tc: output rows must be 16-byte aligned");
    if (head_dim == 64 && !causal)
        return launch_flash_tc_t<64, false>(h, q, ldq, k, ldk, v, ldv, out, ldo, cu_seqlens, n_seqs, max_seqlen, total_tokens, n_heads, scale, q_offset, stream);
    if (head_dim == 64 && causal)
        return launch_flash_tc_t<64, true>(h, q, ldq, k, ldk, v, ldv, out, ldo, cu_seqlens, n_seqs, max_seqlen, total_tokens, n_heads, scale, q_offset, stream);
    if (head_dim == 128 && !causal)
        return launch_flash_tc_t<128, false>(h, q, ldq, k, ldk, v, ldv, out, ldo, cu_seqlens, n_seqs, max_seqlen, total_tokens, n_heads, scale, q_offset, stream);
    if (head_dim == 128 && causal)
        return launch_flash_tc_t<128, true>(h, q, ldq, k, ldk, v, ldv, out, ldo, cu_seqlens, n_seqs, max_seqlen, total_tokens, n_heads, scale, q_offset, stream);
    set_error("flash_attention_tc: head_dim %d unsupported (64 or 128)", head_dim);
    return TEO_ERR_UNSUPPORTED;
}

}  // namespace teo

extern "C" int teo_flash_attention_tc(teo_handle* h, const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out,
                                      int ldo, const void* cu_seqlens, int n_seqs, int max_seqlen, int total_tokens, int n_heads,
                                      int head_dim, float scale, int causal, int q_offset, void* stream) {
    return teo::launch_flash_attention_tc(h, static_cast<const teo::bf16*>(q), ldq, static_cast<const teo::bf16*>(k), ldk,
                                          static_cast<const teo::bf16*>(v), ldv, static_cast<teo::bf16*>(out), ldo,
                                          static_cast<const int*>(cu_seqlens), n_seqs, max_seqlen, total_tokens, n_heads, head_dim, scale,
                                          causal, q_offset, static_cast<cudaStream_t>(stream));
}
