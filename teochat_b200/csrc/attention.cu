// Attention kernels of the hot path that do NOT use tcgen05 — kept as the fallbacks / cross-checks of the tensor-core
// versions (attention_tc.cu: flash_tc_kernel for the ViT and prefill; decode_attn_mma.cu: the decode step's kernel) and
// for shapes those do not cover.
//
// (1) flash_fwd_kernel — variable-length multi-head attention (ViT: non-causal, 257 tokens, head_dim 64; LLaMA prefill:
//     causal, head_dim 128).  Online-softmax tiles of BLOCK_M queries × 64 keys; Q/K/V tiles staged with cp.async into
//     XOR-swizzled shared memory, contractions on mma.sync m16n8k16 (bf16 in, fp32 accumulate).  Selected with
//     TEO_FLASH=mma; GPU tests compare flash_tc_kernel against it.
//
// (2) decode_attn_kernel / decode_attn_persist_kernel — one query token per sequence over the paged KV cache on CUDA
//     cores.  Pure HBM streaming (0.5 MiB per cached token per sequence): whole [page_size × head_dim] K and V
//     page slices arrive by cp.async.bulk (16 KiB contiguous at page_size 64, head_dim 128) into a shared-memory ring
//     signalled by mbarriers; one CTA per (sequence, head, KV split), or persistent CTAs walking those items as one page
//     stream.  Used for (head_dim, page_size) other than (128, 64), through teo_decode_attention (no handle), and with
//     TEO_DEC_ATTN=cuda|v1.  Split partials of all decode kernels are merged by decode_combine_kernel.
#include <stdlib.h>

#include <cmath>
#include <type_traits>

#include "common.h"
#include "ptx.cuh"

namespace teo {

// =============================================================================== flash (prefill / ViT)
constexpr int FL_BN = 64;

template <int HD>
__device__ __forceinline__ uint32_t swz(int row, int chunk) {   // byte offset of 16-byte chunk in a [rows][HD] bf16 tile
    return static_cast<uint32_t>(row * (HD * 2) + ((chunk ^ (row & 7)) << 4));
}

template <int HD, int ROWS, int THREADS>
__device__ __forceinline__ void load_tile_async(uint8_t* smem_tile, const bf16* __restrict__ gbase, long long ld, int row0,
                                                int row_end) {
    constexpr int CPR = HD / 8;   // 16-byte chunks per row
#pragma unroll
    for (int i = threadIdx.x; i < ROWS * CPR; i += THREADS) {
        const int r = i / CPR, c = i % CPR;
        const bool ok = row0 + r < row_end;
        const bf16* src = gbase + static_cast<long long>(ok ? row0 + r : row0) * ld + c * 8;
        cp_async_16(smem_tile + swz<HD>(r, c), src, ok);
    }
}

template <int HD, int BLOCK_M, bool CAUSAL>
__global__ void __launch_bounds__(BLOCK_M * 2)
flash_fwd_kernel(const bf16* __restrict__ Q, long long ldq, const bf16* __restrict__ K, long long ldk,
                 const bf16* __restrict__ V, long long ldv, bf16* __restrict__ O, long long ldo,
                 const int* __restrict__ cu_seqlens, float scale_log2) {
    constexpr int THREADS = BLOCK_M * 2;      // one warp per 16 query rows
    constexpr int KSTEPS = HD / 16;
    constexpr int DT = HD / 8;                // output n-tiles
    extern __shared__ uint8_t fl_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fl_smem_raw) + 127) & ~uintptr_t(127));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + BLOCK_M * HD * 2;
    uint8_t* sV = sK + 2 * FL_BN * HD * 2;

    const int seq = blockIdx.z, head = blockIdx.y;
    const int m_blk = gridDim.x - 1 - blockIdx.x;     // heavy (late) causal blocks first
    const int seq_start = cu_seqlens[seq];
    const int seqlen = cu_seqlens[seq + 1] - seq_start;
    const int m0 = m_blk * BLOCK_M;
    if (m0 >= seqlen) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;

    const bf16* qb = Q + static_cast<long long>(seq_start) * ldq + head * HD;
    const bf16* kb = K + static_cast<long long>(seq_start) * ldk + head * HD;
    const bf16* vb = V + static_cast<long long>(seq_start) * ldv + head * HD;

    const int kv_end = CAUSAL ? min(seqlen, m0 + BLOCK_M) : seqlen;
    const int n_blocks = (kv_end + FL_BN - 1) / FL_BN;

    load_tile_async<HD, BLOCK_M, THREADS>(sQ, qb, ldq, m0, seqlen);
    load_tile_async<HD, FL_BN, THREADS>(sK, kb, ldk, 0, seqlen);
    load_tile_async<HD, FL_BN, THREADS>(sV, vb, ldv, 0, seqlen);
    cp_async_commit();

    uint32_t qf[KSTEPS][4];
    float o[DT][4];
#pragma unroll
    for (int i = 0; i < DT; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const int qrow0 = m0 + warp * 16 + g;     // this thread's two query rows: qrow0, qrow0 + 8

    for (int j = 0; j < n_blocks; ++j) {
        const int buf = j & 1;
        cp_async_wait<0>();
        __syncthreads();
        if (j + 1 < n_blocks) {
            load_tile_async<HD, FL_BN, THREADS>(sK + (buf ^ 1) * FL_BN * HD * 2, kb, ldk, (j + 1) * FL_BN, seqlen);
            load_tile_async<HD, FL_BN, THREADS>(sV + (buf ^ 1) * FL_BN * HD * 2, vb, ldv, (j + 1) * FL_BN, seqlen);
            cp_async_commit();
        }
        if (j == 0) {
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldmatrix_x4(qf[ks], sQ + swz<HD>(r, 2 * ks + (lane >> 4)));
            }
        }
        const uint8_t* sKb = sK + buf * FL_BN * HD * 2;
        const uint8_t* sVb = sV + buf * FL_BN * HD * 2;

        // ---- S = Q K^T (16 × 64 per warp)
        float s[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
            for (int nt = 0; nt < 8; nt += 2) {
                uint32_t kf[4];
                const int r = nt * 8 + (lane & 7) + ((lane >> 4) & 1) * 8;
                ldmatrix_x4(kf, sKb + swz<HD>(r, 2 * ks + ((lane >> 3) & 1)));
                mma_bf16_16816(s[nt], qf[ks], kf[0], kf[1]);
                mma_bf16_16816(s[nt + 1], qf[ks], kf[2], kf[3]);
            }
        }
        // ---- mask + online softmax
        const int kv0 = j * FL_BN;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int kv = kv0 + nt * 8 + 2 * t + (c & 1);
                const int qr = qrow0 + (c >> 1) * 8;
                const bool masked = kv >= seqlen || (CAUSAL && kv > qr);
                if (masked) s[nt][c] = -INFINITY;
                mx[c >> 1] = fmaxf(mx[c >> 1], s[nt][c]);
            }
        }
        float alpha[2], msub[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
            mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
            const float m_new = fmaxf(m_run[h], mx[h]);
            msub[h] = (m_new == -INFINITY) ? 0.f : m_new * scale_log2;
            alpha[h] = (m_run[h] == -INFINITY) ? 0.f : exp2f(m_run[h] * scale_log2 - msub[h]);
            m_run[h] = m_new;
        }
        float rs[2] = {0.f, 0.f};
        uint32_t pf[4][4];    // P as A fragments for the 4 k-steps of 16 keys
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float p[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                p[c] = exp2f(s[nt][c] * scale_log2 - msub[c >> 1]);   // exp2(-inf) = 0 for masked
                rs[c >> 1] += p[c];
            }
            pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p[0], p[1]);
            pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p[2], p[3]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            rs[h] += __shfl_xor_sync(0xffffffffu, rs[h], 1);
            rs[h] += __shfl_xor_sync(0xffffffffu, rs[h], 2);
            l_run[h] = l_run[h] * alpha[h] + rs[h];
        }
#pragma unroll
        for (int dt = 0; dt < DT; ++dt) {
            o[dt][0] *= alpha[0]; o[dt][1] *= alpha[0];
            o[dt][2] *= alpha[1]; o[dt][3] *= alpha[1];
        }
        // ---- O += P V
#pragma unroll
        for (int ks2 = 0; ks2 < 4; ++ks2) {
#pragma unroll
            for (int dt = 0; dt < DT; dt += 2) {
                uint32_t vf[4];
                const int r = ks2 * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldmatrix_x4_trans(vf, sVb + swz<HD>(r, dt + (lane >> 4)));
                mma_bf16_16816(o[dt], pf[ks2], vf[0], vf[1]);
                mma_bf16_16816(o[dt + 1], pf[ks2], vf[2], vf[3]);
            }
        }
    }
    // ---- normalise and store
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int qr = qrow0 + h * 8;
        if (qr < seqlen) {
            const float inv = 1.0f / l_run[h];
            bf16* dst = O + static_cast<long long>(seq_start + qr) * ldo + head * HD;
#pragma unroll
            for (int dt = 0; dt < DT; ++dt)
                *reinterpret_cast<uint32_t*>(dst + dt * 8 + 2 * t) = pack_bf16x2(o[dt][2 * h] * inv, o[dt][2 * h + 1] * inv);
        }
    }
}

template <int HD, int BLOCK_M, bool CAUSAL>
static int launch_flash(const bf16* q, int ldq, const bf16* k, int ldk, const bf16* v, int ldv, bf16* out, int ldo,
                        const int* cu, int n_seqs, int max_seqlen, int n_heads, float scale, cudaStream_t stream) {
    constexpr int SMEM = BLOCK_M * HD * 2 + 4 * FL_BN * HD * 2 + 128;
    static bool attr_set = false;
    if (!attr_set) {
        TEO_CUDA(cudaFuncSetAttribute(flash_fwd_kernel<HD, BLOCK_M, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set = true;
    }
    dim3 grid((max_seqlen + BLOCK_M - 1) / BLOCK_M, n_heads, n_seqs);
    flash_fwd_kernel<HD, BLOCK_M, CAUSAL><<<grid, BLOCK_M * 2, SMEM, stream>>>(q, ldq, k, ldk, v, ldv, out, ldo, cu,
                                                                              scale * 1.4426950408889634f);
    TEO_LAUNCH_CHECK("flash_fwd_kernel");
    return TEO_OK;
}

int launch_flash_attention(const bf16* q, int ldq, const bf16* k, int ldk, const bf16* v, int ldv, bf16* out, int ldo,
                           const int* cu_seqlens, int n_seqs, int max_seqlen, int n_heads, int head_dim, float scale,
                           int causal, cudaStream_t stream) {
    TEO_CHECK_ARG(q && k && v && out && cu_seqlens, "flash_attention: null pointer");
    TEO_CHECK_ARG(n_seqs > 0 && max_seqlen > 0 && n_heads > 0, "flash_attention: bad sizes");
    TEO_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "flash_attention: row strides must keep 16-byte alignment");
    TEO_CHECK_ARG(((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0,
                  "flash_attention: q/k/v must be 16-byte aligned");
    if (head_dim == 64) {
        return causal ? launch_flash<64, 64, true>(q, ldq, k, ldk, v, ldv, out, ldo, cu_seqlens, n_seqs, max_seqlen, n_heads, scale, stream)
                      : launch_flash<64, 64, false>(q, ldq, k, ldk, v, ldv, out, ldo, cu_seqlens, n_seqs, max_seqlen, n_heads, scale, stream);
    }
    if (head_dim == 128) {
        return causal ? launch_flash<128, 128, true>(q, ldq, k, ldk, v, ldv, out, ldo, cu_seqlens, n_seqs, max_seqlen, n_heads, scale, stream)
                      : launch_flash<128, 128, false>(q, ldq, k, ldk, v, ldv, out, ldo, cu_seqlens, n_seqs, max_seqlen, n_heads, scale, stream);
    }
    set_error("flash_attention: head_dim %d unsupported (64 or 128)", head_dim);
    return TEO_ERR_UNSUPPORTED;
}

// =============================================================================== paged decode
constexpr int DEC_THREADS = 128;
// K+V page pairs in the persistent kernel's ring (TEO_DEC_STAGES=2|3|4 overrides, for measurements)
static int dec_stages() {
    static const int n = [] {
        const char* e = getenv("TEO_DEC_STAGES");
        const int v = e ? atoi(e) : 2;
        return (v >= 2 && v <= 4) ? v : 2;
    }();
    return n;
}

// grid (splits, heads, seqs).  Workspace (splits > 1): o_part f32 [seq][head][split][HD], ml f32 [seq][head][split][2].
template <int HD, int PAGE>
__global__ void __launch_bounds__(DEC_THREADS)
decode_attn_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ kv_pages, const int* __restrict__ block_table,
                   int max_pages, const int* __restrict__ seq_lens, int len_bias, bf16* __restrict__ out, float* __restrict__ o_part,
                   float* __restrict__ ml_part, int n_heads, float scale_log2) {
    static_assert(HD % 16 == 0 && PAGE % 8 == 0 && PAGE * 2 <= DEC_THREADS * 8, "decode tile shape");
    constexpr int PAGE_BYTES = PAGE * HD * 2;
    constexpr int TPT = DEC_THREADS / PAGE >= 2 ? 2 : 1;   // threads cooperating on one key (QK phase)
    constexpr int KEYS_PER_PASS = DEC_THREADS / TPT;
    constexpr int CH = HD / 8 / TPT;                       // 16-byte chunks per thread per key
    constexpr int DCH = HD / 8;                            // 16-byte chunks per V row
    constexpr int TG = DEC_THREADS / DCH;                  // key groups in the PV phase
    extern __shared__ uint8_t dec_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dec_smem_raw) + 127) & ~uintptr_t(127));
    uint8_t* sK = smem;                        // [2][PAGE][HD] bf16
    uint8_t* sV = sK + 2 * PAGE_BYTES;         // [2][PAGE][HD] bf16
    float* sQ = reinterpret_cast<float*>(sV + 2 * PAGE_BYTES);   // [HD]
    float* sS = sQ + HD;                       // [PAGE] scores, then probabilities
    float* sO = sS + PAGE;                     // [TG][HD] cross-group reduction
    uint64_t* bar = reinterpret_cast<uint64_t*>(sO + TG * HD);    // [2]

    const int split = blockIdx.x, n_splits = gridDim.x, head = blockIdx.y, seq = blockIdx.z;
    const int tid = threadIdx.x, lane = tid & 31;
    pdl_trigger();
    pdl_wait();                                  // q, the new K/V rows and seq_lens come from the previous kernels
    const int len = seq_lens[seq] + len_bias;
    const int n_pages = (len + PAGE - 1) / PAGE;
    const int pps = (n_pages + n_splits - 1) / n_splits;
    const int p0 = split * pps, p1 = min(n_pages, p0 + pps);

    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    for (int i = tid; i < HD; i += DEC_THREADS) sQ[i] = __bfloat162float(q[static_cast<long long>(seq) * ldq + head * HD + i]);
    __syncthreads();

    const int* bt = block_table + static_cast<long long>(seq) * max_pages;
    auto issue = [&](int p, int buf) {
        const int page = bt[p];
        const bf16* kp = kv_pages + ((static_cast<long long>(page) * 2 + 0) * n_heads + head) * (PAGE * HD);
        const bf16* vp = kv_pages + ((static_cast<long long>(page) * 2 + 1) * n_heads + head) * (PAGE * HD);
        mbar_arrive_expect_tx(&bar[buf], 2 * PAGE_BYTES);
        bulk_load_1d(sK + buf * PAGE_BYTES, kp, PAGE_BYTES, &bar[buf]);
        bulk_load_1d(sV + buf * PAGE_BYTES, vp, PAGE_BYTES, &bar[buf]);
    };
    if (tid == 0 && p0 < p1) issue(p0, 0);

    float m_run = -INFINITY, l_run = 0.f;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int dc = tid % DCH, tg = tid / DCH;   // PV phase: 8 dims [dc*8, dc*8+8) of keys ≡ tg (mod TG)

    for (int p = p0; p < p1; ++p) {
        const int buf = (p - p0) & 1;
        if (tid == 0 && p + 1 < p1) issue(p + 1, buf ^ 1);   // buffer buf^1 was released by the trailing __syncthreads
        mbar_wait(&bar[buf], static_cast<uint32_t>(((p - p0) >> 1) & 1));
        const uint8_t* kb = sK + buf * PAGE_BYTES;
        const uint8_t* vb = sV + buf * PAGE_BYTES;
        const int valid = min(PAGE, len - p * PAGE);
        // ---- scores: TPT threads per key, rotated chunk order (bank-conflict-free 16-byte reads)
        for (int key0 = 0; key0 < PAGE; key0 += KEYS_PER_PASS) {
            const int key = key0 + tid / TPT, part = tid % TPT;
            float sacc = 0.f;
            if (key < PAGE) {
                const int rot = key * TPT + part;
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int c = part * CH + ((i + rot) % CH);
                    const uint4 kv = *reinterpret_cast<const uint4*>(kb + key * (HD * 2) + c * 16);
                    const float4 q0 = *reinterpret_cast<const float4*>(sQ + c * 8);
                    const float4 q1 = *reinterpret_cast<const float4*>(sQ + c * 8 + 4);
                    sacc += bf16_lo(kv.x) * q0.x + bf16_hi(kv.x) * q0.y + bf16_lo(kv.y) * q0.z + bf16_hi(kv.y) * q0.w +
                            bf16_lo(kv.z) * q1.x + bf16_hi(kv.z) * q1.y + bf16_lo(kv.w) * q1.z + bf16_hi(kv.w) * q1.w;
                }
            }
            if (TPT == 2) sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
            if (key < PAGE && part == 0) sS[key] = key < valid ? sacc : -INFINITY;
        }
        __syncthreads();
        // ---- online softmax (every warp recomputes the page max: no extra barrier)
        float mx = -INFINITY;
        for (int i = lane; i < PAGE; i += 32) mx = fmaxf(mx, sS[i]);
        mx = warp_max(mx);
        const float m_new = fmaxf(m_run, mx);
        const float msub = m_new * scale_log2;      // valid ≥ 1 ⇒ finite
        const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run * scale_log2 - msub);
        m_run = m_new;
        float psum = 0.f;
        for (int i = lane; i < PAGE; i += 32) psum += exp2f(sS[i] * scale_log2 - msub);
        psum = warp_sum(psum);
        l_run = l_run * alpha + psum;
        // ---- O += P V : thread owns 8 dims of the keys in its group
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= alpha;
        for (int key = tg; key < valid; key += TG) {
            float pr = exp2f(sS[key] * scale_log2 - msub);
            pr = __bfloat162float(__float2bfloat16_rn(pr));    // P is bf16 in the prefill path too
            const uint4 vv = *reinterpret_cast<const uint4*>(vb + key * (HD * 2) + dc * 16);
            acc[0] += pr * bf16_lo(vv.x); acc[1] += pr * bf16_hi(vv.x);
            acc[2] += pr * bf16_lo(vv.y); acc[3] += pr * bf16_hi(vv.y);
            acc[4] += pr * bf16_lo(vv.z); acc[5] += pr * bf16_hi(vv.z);
            acc[6] += pr * bf16_lo(vv.w); acc[7] += pr * bf16_hi(vv.w);
        }
        __syncthreads();     // everyone is done with sS and buffer `buf`
    }
    // ---- reduce the TG key groups, write result / partial
#pragma unroll
    for (int j = 0; j < 8; ++j) sO[tg * HD + dc * 8 + j] = acc[j];
    __syncthreads();
    for (int d = tid; d < HD; d += DEC_THREADS) {
        float v = 0.f;
#pragma unroll
        for (int gI = 0; gI < TG; ++gI) v += sO[gI * HD + d];
        if (n_splits == 1) {
            out[(static_cast<long long>(seq) * n_heads + head) * HD + d] = __float2bfloat16_rn(v / l_run);
        } else {
            o_part[((static_cast<long long>(seq) * n_heads + head) * n_splits + split) * HD + d] = v;
        }
    }
    if (n_splits > 1 && tid == 0) {
        float* ml = ml_part + ((static_cast<long long>(seq) * n_heads + head) * n_splits + split) * 2;
        ml[0] = m_run;      // -inf when this split had no pages
        ml[1] = l_run;
    }
}

// Persistent form of decode_attn_kernel (the default): gridDim.x CTAs walk the (sequence, head, split) items
// i = blockIdx.x, blockIdx.x + gridDim.x, …  and treat their pages as ONE stream — thread 0 keeps the 2-deep page
// ring full across item boundaries (the next item's first pages are in flight while this item's last page is being
// reduced), and the next item's query is fetched into registers one item ahead.  With one CTA per item every CTA
// paid launch + barrier init + q load + first-page latency (≈ 2 µs of a ≈ 20 µs life) with its share of the HBM pipe
// idle; here that is paid once per kernel.  Same arithmetic, same order of operations as decode_attn_kernel.
template <int HD, int PAGE, int NST>
__global__ void __launch_bounds__(DEC_THREADS)
decode_attn_persist_kernel(const bf16* __restrict__ q, long long ldq, const bf16* __restrict__ kv_pages, const int* __restrict__ block_table,
                           int max_pages, const int* __restrict__ seq_lens, int len_bias, bf16* __restrict__ out,
                           float* __restrict__ o_part, float* __restrict__ ml_part, int n_heads, int n_splits, int n_items,
                           float scale_log2) {
    static_assert(HD % 16 == 0 && PAGE % 8 == 0 && PAGE * 2 <= DEC_THREADS * 8 && HD <= DEC_THREADS, "decode tile shape");
    constexpr int PAGE_BYTES = PAGE * HD * 2;
    constexpr int TPT = DEC_THREADS / PAGE >= 2 ? 2 : 1;
    constexpr int KEYS_PER_PASS = DEC_THREADS / TPT;
    constexpr int CH = HD / 8 / TPT;
    constexpr int DCH = HD / 8;
    constexpr int TG = DEC_THREADS / DCH;
    extern __shared__ uint8_t dec_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dec_smem_raw) + 127) & ~uintptr_t(127));
    uint8_t* sK = smem;
    uint8_t* sV = sK + NST * PAGE_BYTES;
    float* sQ = reinterpret_cast<float*>(sV + NST * PAGE_BYTES);
    float* sS = sQ + HD;
    float* sO = sS + PAGE;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sO + TG * HD);

    const int tid = threadIdx.x, lane = tid & 31;
    pdl_trigger();
    pdl_wait();                                  // q, the new K/V rows and seq_lens come from the previous kernels
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) mbar_init(&bar[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    // item → (sequence, head, split) and its page range; splits fastest, like the one-CTA-per-item grid
    auto item_range = [&](int item, int& seq, int& head, int& split, int& len, int& p0, int& p1) {
        split = item % n_splits;
        const int sh = item / n_splits;
        head = sh % n_heads;
        seq = sh / n_heads;
        len = seq_lens[seq] + len_bias;
        const int n_pages = (len + PAGE - 1) / PAGE;
        const int pps = (n_pages + n_splits - 1) / n_splits;
        p0 = split * pps;
        p1 = min(n_pages, p0 + pps);
    };
    // ---- loader cursor (thread 0): next page to request, possibly items ahead of the compute cursor
    int l_item = blockIdx.x, l_seq = 0, l_head = 0, l_p = 0, l_p1 = 0;
    uint32_t l_cnt = 0;
    auto loader_next = [&]() {
        while (l_item < n_items && l_p >= l_p1) {
            l_item += gridDim.x;
            if (l_item < n_items) {
                int split, len;
                item_range(l_item, l_seq, l_head, split, len, l_p, l_p1);
            }
        }
        if (l_item >= n_items) return;
        const int buf = l_cnt % NST;
        const int page = block_table[static_cast<long long>(l_seq) * max_pages + l_p];
        const bf16* kp = kv_pages + ((static_cast<long long>(page) * 2 + 0) * n_heads + l_head) * (PAGE * HD);
        const bf16* vp = kv_pages + ((static_cast<long long>(page) * 2 + 1) * n_heads + l_head) * (PAGE * HD);
        mbar_arrive_expect_tx(&bar[buf], 2 * PAGE_BYTES);
        bulk_load_1d(sK + buf * PAGE_BYTES, kp, PAGE_BYTES, &bar[buf]);
        bulk_load_1d(sV + buf * PAGE_BYTES, vp, PAGE_BYTES, &bar[buf]);
        ++l_p;
        ++l_cnt;
    };
    if (tid == 0) {
        if (l_item < n_items) {
            int split, len;
            item_range(l_item, l_seq, l_head, split, len, l_p, l_p1);
        }
        for (int s = 0; s < NST; ++s) loader_next();
    }

    const int dc = tid % DCH, tg = tid / DCH;   // PV phase: 8 dims [dc*8, dc*8+8) of keys ≡ tg (mod TG)
    uint32_t c_cnt = 0;                          // pages consumed by this CTA so far (ring stage / phase)
    float q_next = 0.f;                          // this thread's element of the NEXT item's query
    if (blockIdx.x < n_items && tid < HD) {
        int seq, head, split, len, p0, p1;
        item_range(blockIdx.x, seq, head, split, len, p0, p1);
        q_next = __bfloat162float(q[static_cast<long long>(seq) * ldq + head * HD + tid]);
    }
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int seq, head, split, len, p0, p1;
        item_range(item, seq, head, split, len, p0, p1);
        if (tid < HD) sQ[tid] = q_next;
        if (item + static_cast<int>(gridDim.x) < n_items && tid < HD) {
            int s2, h2, sp2, l2, a2, b2;
            item_range(item + gridDim.x, s2, h2, sp2, l2, a2, b2);
            q_next = __bfloat162float(q[static_cast<long long>(s2) * ldq + h2 * HD + tid]);
        }
        __syncthreads();                         // sQ ready; the previous item's readers of sO / sS are done

        float m_run = -INFINITY, l_run = 0.f;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int p = p0; p < p1; ++p, ++c_cnt) {
            const int buf = c_cnt % NST;
            mbar_wait(&bar[buf], (c_cnt / NST) & 1);
            const uint8_t* kb = sK + buf * PAGE_BYTES;
            const uint8_t* vb = sV + buf * PAGE_BYTES;
            const int valid = min(PAGE, len - p * PAGE);
            // ---- scores: TPT threads per key, rotated chunk order (bank-conflict-free 16-byte reads)
            for (int key0 = 0; key0 < PAGE; key0 += KEYS_PER_PASS) {
                const int key = key0 + tid / TPT, part = tid % TPT;
                float sacc = 0.f;
                if (key < PAGE) {
                    const int rot = key * TPT + part;
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const int c = part * CH + ((i + rot) % CH);
                        const uint4 kv = *reinterpret_cast<const uint4*>(kb + key * (HD * 2) + c * 16);
                        const float4 q0 = *reinterpret_cast<const float4*>(sQ + c * 8);
                        const float4 q1 = *reinterpret_cast<const float4*>(sQ + c * 8 + 4);
                        sacc += bf16_lo(kv.x) * q0.x + bf16_hi(kv.x) * q0.y + bf16_lo(kv.y) * q0.z + bf16_hi(kv.y) * q0.w +
                                bf16_lo(kv.z) * q1.x + bf16_hi(kv.z) * q1.y + bf16_lo(kv.w) * q1.z + bf16_hi(kv.w) * q1.w;
                    }
                }
                if (TPT == 2) sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
                if (key < PAGE && part == 0) sS[key] = key < valid ? sacc : -INFINITY;
            }
            __syncthreads();
            // ---- online softmax (every warp recomputes the page max: no extra barrier)
            float mx = -INFINITY;
            for (int i = lane; i < PAGE; i += 32) mx = fmaxf(mx, sS[i]);
            mx = warp_max(mx);
            const float m_new = fmaxf(m_run, mx);
            const float msub = m_new * scale_log2;      // valid ≥ 1 ⇒ finite
            const float alpha = (m_run == -INFINITY) ? 0.f : exp2f(m_run * scale_log2 - msub);
            m_run = m_new;
            float psum = 0.f;
            for (int i = lane; i < PAGE; i += 32) psum += exp2f(sS[i] * scale_log2 - msub);
            psum = warp_sum(psum);
            l_run = l_run * alpha + psum;
            // ---- O += P V : thread owns 8 dims of the keys in its group
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] *= alpha;
            for (int key = tg; key < valid; key += TG) {
                float pr = exp2f(sS[key] * scale_log2 - msub);
                pr = __bfloat162float(__float2bfloat16_rn(pr));    // P is bf16 in the prefill path too
                const uint4 vv = *reinterpret_cast<const uint4*>(vb + key * (HD * 2) + dc * 16);
                acc[0] += pr * bf16_lo(vv.x); acc[1] += pr * bf16_hi(vv.x);
                acc[2] += pr * bf16_lo(vv.y); acc[3] += pr * bf16_hi(vv.y);
                acc[4] += pr * bf16_lo(vv.z); acc[5] += pr * bf16_hi(vv.z);
                acc[6] += pr * bf16_lo(vv.w); acc[7] += pr * bf16_hi(vv.w);
            }
            __syncthreads();     // everyone is done with sS and ring slot `buf` …
            if (tid == 0) loader_next();         // … which takes the next page of this CTA's stream (maybe a later item's)
        }
        // ---- reduce the TG key groups, write result / partial
#pragma unroll
        for (int j = 0; j < 8; ++j) sO[tg * HD + dc * 8 + j] = acc[j];
        __syncthreads();
        for (int d = tid; d < HD; d += DEC_THREADS) {
            float v = 0.f;
#pragma unroll
            for (int gI = 0; gI < TG; ++gI) v += sO[gI * HD + d];
            if (n_splits == 1) {
                out[(static_cast<long long>(seq) * n_heads + head) * HD + d] = __float2bfloat16_rn(v / l_run);
            } else {
                o_part[((static_cast<long long>(seq) * n_heads + head) * n_splits + split) * HD + d] = v;
            }
        }
        if (n_splits > 1 && tid == 0) {
            float* ml = ml_part + ((static_cast<long long>(seq) * n_heads + head) * n_splits + split) * 2;
            ml[0] = m_run;      // -inf when this split had no pages
            ml[1] = l_run;
        }
    }
}

template <int HD>
__global__ void decode_combine_kernel(const float* __restrict__ o_part, const float* __restrict__ ml_part, bf16* __restrict__ out,
                                      int n_heads, int n_splits, float scale_log2) {
    const int head = blockIdx.x, seq = blockIdx.y;
    const long long base = static_cast<long long>(seq) * n_heads + head;
    pdl_trigger();
    pdl_wait();
    float M = -INFINITY;
    for (int s = 0; s < n_splits; ++s) M = fmaxf(M, ml_part[(base * n_splits + s) * 2]);
    float L = 0.f;
    for (int s = 0; s < n_splits; ++s) {
        const float m = ml_part[(base * n_splits + s) * 2];
        if (m != -INFINITY) L += ml_part[(base * n_splits + s) * 2 + 1] * exp2f((m - M) * scale_log2);
    }
    for (int d = threadIdx.x; d < HD; d += blockDim.x) {
        float v = 0.f;
        for (int s = 0; s < n_splits; ++s) {
            const float m = ml_part[(base * n_splits + s) * 2];
            if (m != -INFINITY) v += o_part[(base * n_splits + s) * HD + d] * exp2f((m - M) * scale_log2);
        }
        out[base * HD + d] = __float2bfloat16_rn(v / L);
    }
}

// KV splits per (sequence, head).  All CTAs of a launch carry about the same work, so the launch
// runs in whole waves of (SMs × resident CTAs): pick the smallest split count whose last wave is
// ≥ 95 % full (bs=32 × 32 heads = 1024 CTAs over 444 slots is 2.31 waves → 77 % efficient; 3 splits
// give 6.92 waves → 99 %), keeping ≥ 4 pages per split.
static int decode_splits(int n_seqs, int n_heads, int max_seq_len, int page_size, int num_sms, int ctas_per_sm) {
    const int ctas = n_seqs * n_heads;
    const int pages = (max_seq_len + page_size - 1) / page_size;
    const int slots = num_sms * ctas_per_sm;
    const int max_s = std::max(1, std::min(32, pages / 4));
    int best = 1;
    double best_eff = 0.0;
    for (int s = 1; s <= max_s; ++s) {
        const double waves = static_cast<double>(ctas) * s / slots;
        const double eff = waves / std::ceil(waves);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
        if (eff >= 0.95) { best = s; break; }
    }
    return best;
}

template <int HD, int PAGE>
static int launch_decode_t(const bf16* q, int ldq, const bf16* kv_pages, const int* block_table, int max_pages, const int* seq_lens,
                           int len_bias, bf16* out, int n_seqs, int n_heads, int splits, float scale, float* o_part, float* ml_part,
                           int resident_ctas, cudaStream_t stream) {
    constexpr int TG = DEC_THREADS / (HD / 8);
    constexpr int SMEM = 4 * PAGE * HD * 2 + (HD + PAGE + TG * HD) * 4 + 16 + 128;
    static bool attr_set = false;
    if (!attr_set) {
        TEO_CUDA(cudaFuncSetAttribute(decode_attn_kernel<HD, PAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set = true;
    }
    const float sl2 = scale * 1.4426950408889634f;
    static const bool persistent = [] {
        const char* e = getenv("TEO_DEC_ATTN");           // "v1": one CTA per (split, head, sequence) (A/B measurements)
        return !(e && e[0] == 'v' && e[1] == '1');
    }();
    if (persistent) {
        const int n_items = splits * n_heads * n_seqs;
        const int grid_p = std::min(n_items, resident_ctas);
        auto go = [&](auto nst_tag) -> int {
            constexpr int NST = decltype(nst_tag)::value;
            constexpr int SMEM_P = 2 * NST * PAGE * HD * 2 + (HD + PAGE + TG * HD) * 4 + 8 * NST + 128;
            static bool attr2_set = false;
            if (!attr2_set) {
                TEO_CUDA(cudaFuncSetAttribute(decode_attn_persist_kernel<HD, PAGE, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_P));
                attr2_set = true;
            }
            TEO_CUDA(launch_kc(PDL_ATTN, decode_attn_persist_kernel<HD, PAGE, NST>, dim3(grid_p), dim3(DEC_THREADS), SMEM_P, stream, q,
                               static_cast<long long>(ldq), kv_pages, block_table, max_pages, seq_lens, len_bias, out, o_part, ml_part,
                               n_heads, splits, n_items, sl2));
            return TEO_OK;
        };
        const int nst = dec_stages();
        TEO_TRY(nst == 3 ? go(std::integral_constant<int, 3>{}) : nst == 4 ? go(std::integral_constant<int, 4>{})
                                                                         : go(std::integral_constant<int, 2>{}));
        TEO_LAUNCH_CHECK("decode_attn_persist_kernel");
    } else {
        dim3 grid(splits, n_heads, n_seqs);
        TEO_CUDA(launch_kc(PDL_ATTN, decode_attn_kernel<HD, PAGE>, grid, dim3(DEC_THREADS), SMEM, stream, q, static_cast<long long>(ldq), kv_pages,
                           block_table, max_pages, seq_lens, len_bias, out, o_part, ml_part, n_heads, sl2));
        TEO_LAUNCH_CHECK("decode_attn_kernel");
    }
    if (splits > 1) {
        TEO_CUDA(launch_k(decode_combine_kernel<HD>, dim3(n_heads, n_seqs), dim3(HD), 0, stream, static_cast<const float*>(o_part),
                          static_cast<const float*>(ml_part), out, n_heads, splits, sl2));
        TEO_LAUNCH_CHECK("decode_combine_kernel");
    }
    return TEO_OK;
}

int launch_decode_attention_mma(teo_handle* h, const bf16* q, int ldq, const bf16* kv_pages, const int* block_table, int max_pages,
                                const int* seq_lens, int len_bias, bf16* out, int n_seqs, int n_heads, int splits, float scale,
                                float* o_part, float* ml_part, bool* merged, cudaStream_t stream);

int launch_decode_attention(teo_handle* h, const bf16* q, int ldq, const bf16* kv_pages, const int* block_table, int max_pages,
                            const int* seq_lens, int len_bias, bf16* out, int n_seqs, int n_heads, int head_dim, int page_size,
                            int max_seq_len, float scale, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    TEO_CHECK_ARG(q && kv_pages && block_table && seq_lens && out, "decode_attention: null pointer");
    TEO_CHECK_ARG(n_seqs > 0 && n_heads > 0 && max_seq_len > 0, "decode_attention: bad sizes");
    const int num_sms = h ? h->num_sms : 148;
    // Tensor-core kernel (decode_attn_mma.cu) whenever a handle is there to cache the pool's TMA descriptor and the
    // shape is LLaMA's; TEO_DEC_ATTN=cuda|v1 keeps the CUDA-core kernels of this file (A/B measurements).
    static const bool allow_mma = [] {
        const char* e = getenv("TEO_DEC_ATTN");
        return e == nullptr || (e[0] != 'c' && e[0] != 'v');
    }();
    if (h && allow_mma && head_dim == 128 && page_size == 64 && (ldq % 2) == 0 && (reinterpret_cast<uintptr_t>(kv_pages) & 15) == 0) {
        const int splits = decode_splits(n_seqs, n_heads, max_seq_len, page_size, num_sms, 3);
        float *o_part = nullptr, *ml_part = nullptr;
        if (splits > 1) {
            const size_t need = teo_decode_attention_workspace_bytes(n_seqs, n_heads, head_dim, splits);
            if (workspace == nullptr || workspace_bytes < need) {
                set_error("decode_attention: %d splits need %zu workspace bytes, got %zu", splits, need, workspace_bytes);
                return TEO_ERR_WORKSPACE;
            }
            o_part = static_cast<float*>(workspace);
            ml_part = o_part + static_cast<size_t>(n_seqs) * n_heads * splits * head_dim;
        }
        bool merged = false;
        TEO_TRY(launch_decode_attention_mma(h, q, ldq, kv_pages, block_table, max_pages, seq_lens, len_bias, out, n_seqs, n_heads, splits,
                                            scale, o_part, ml_part, &merged, stream));
        if (splits > 1 && !merged) {
            TEO_CUDA(launch_k(decode_combine_kernel<128>, dim3(n_heads, n_seqs), dim3(128), 0, stream, static_cast<const float*>(o_part),
                              static_cast<const float*>(ml_part), out, n_heads, splits, scale * 1.4426950408889634f));
            TEO_LAUNCH_CHECK("decode_combine_kernel");
        }
        h->launches += (splits > 1 && !merged) ? 2 : 1;
        return TEO_OK;
    }
    const int smem_per_cta = 2 * dec_stages() * page_size * head_dim * 2 + (head_dim + page_size + (DEC_THREADS / (head_dim / 8)) * head_dim) * 4 + 8 * dec_stages() + 128;
    const int ctas_per_sm = std::max(1, std::min(16, (227 * 1024) / (smem_per_cta + 1024)));
    int splits = decode_splits(n_seqs, n_heads, max_seq_len, page_size, num_sms, ctas_per_sm);
    float *o_part = nullptr, *ml_part = nullptr;
    if (splits > 1) {
        const size_t need = teo_decode_attention_workspace_bytes(n_seqs, n_heads, head_dim, splits);
        if (workspace == nullptr || workspace_bytes < need) {
            set_error("decode_attention: %d splits need %zu workspace bytes, got %zu", splits, need, workspace_bytes);
            return TEO_ERR_WORKSPACE;
        }
        o_part = static_cast<float*>(workspace);
        ml_part = o_part + static_cast<size_t>(n_seqs) * n_heads * splits * head_dim;
    }
    int rc;
    if (head_dim == 128 && page_size == 64)
        rc = launch_decode_t<128, 64>(q, ldq, kv_pages, block_table, max_pages, seq_lens, len_bias, out, n_seqs, n_heads, splits, scale, o_part, ml_part, num_sms * ctas_per_sm, stream);
    else if (head_dim == 128 && page_size == 16)
        rc = launch_decode_t<128, 16>(q, ldq, kv_pages, block_table, max_pages, seq_lens, len_bias, out, n_seqs, n_heads, splits, scale, o_part, ml_part, num_sms * ctas_per_sm, stream);
    else if (head_dim == 64 && page_size == 64)
        rc = launch_decode_t<64, 64>(q, ldq, kv_pages, block_table, max_pages, seq_lens, len_bias, out, n_seqs, n_heads, splits, scale, o_part, ml_part, num_sms * ctas_per_sm, stream);
    else {
        set_error("decode_attention: (head_dim %d, page_size %d) unsupported; built: (128,64) (128,16) (64,64)", head_dim, page_size);
        return TEO_ERR_UNSUPPORTED;
    }
    if (rc == TEO_OK && h) h->launches += splits > 1 ? 2 : 1;
    return rc;
}

}  // namespace teo

using namespace teo;

extern "C" int teo_flash_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out, int ldo,
                                   const void* cu_seqlens, int n_seqs, int max_seqlen, int n_heads, int head_dim, float scale,
                                   int causal, void* stream) {
    return launch_flash_attention(static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(k), ldk, static_cast<const bf16*>(v), ldv,
                                  static_cast<bf16*>(out), ldo, static_cast<const int*>(cu_seqlens), n_seqs, max_seqlen, n_heads,
                                  head_dim, scale, causal, static_cast<cudaStream_t>(stream));
}

extern "C" size_t teo_decode_attention_workspace_bytes(int n_seqs, int n_heads, int head_dim, int max_splits) {
    if (max_splits < 1) max_splits = 32;
    return static_cast<size_t>(n_seqs) * n_heads * max_splits * (head_dim + 2) * sizeof(float);
}

extern "C" int teo_decode_attention_h(teo_handle* h, const void* q, int ldq, const void* kv_pages, const void* block_table, int max_pages,
                                      const void* seq_lens, void* out, int n_seqs, int n_heads, int head_dim, int page_size,
                                      int max_seq_len, float scale, void* workspace, size_t workspace_bytes, void* stream) {
    TEO_CHECK_ARG(h != nullptr, "decode_attention_h: null handle");
    return launch_decode_attention(h, static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(kv_pages),
                                   static_cast<const int*>(block_table), max_pages, static_cast<const int*>(seq_lens), 0,
                                   static_cast<bf16*>(out), n_seqs, n_heads, head_dim, page_size, max_seq_len, scale, workspace,
                                   workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int teo_decode_attention(const void* q, int ldq, const void* kv_pages, const void* block_table, int max_pages,
                                    const void* seq_lens, void* out, int n_seqs, int n_heads, int head_dim, int page_size,
                                    int max_seq_len, float scale, void* workspace, size_t workspace_bytes, void* stream) {
    return launch_decode_attention(nullptr, static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(kv_pages),
                                   static_cast<const int*>(block_table), max_pages, static_cast<const int*>(seq_lens), 0,
                                   static_cast<bf16*>(out), n_seqs, n_heads, head_dim, page_size, max_seq_len, scale, workspace,
                                   workspace_bytes, static_cast<cudaStream_t>(stream));
}
