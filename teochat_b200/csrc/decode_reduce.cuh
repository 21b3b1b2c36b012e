// Reductions of the decode GEMMs' fp32 stream-K partials fused with the next element-wise stage (SwiGLU, RoPE + KV-page write,
// residual + RMSNorm, logits), as __device__ functions: the stand-alone glue kernels of kernels_misc.cu and the persistent
// decode chain kernel (decode_chain.cu) run the SAME code, so both decode paths produce bit-identical results.
// Fixed summation order s = 0,1,… over the partial slots everywhere (deterministic).  Partials are read with ld.global.cg (L2 only):
// inside the chain kernel they were written by OTHER SMs during the same launch, and a line left in this SM's L1 by an earlier
// launch over the same workspace must not be served instead.
#pragma once
#include "common.h"
#include "ptx.cuh"

namespace teo {

// Column of gate value i of a [rows, 2·inter] gate/up row: [gate | up] halves, or interleaved in blocks of 32
// (| gate 32 | up 32 |, TEO_ACT_SWIGLU_PAIRS layout); the matching up value sits `up_off` columns further.
__device__ __forceinline__ long long gate_col(long long c, int interleaved) { return interleaved ? (c / 32) * 64 + (c % 32) : c; }

// Σ_s P[s][idx] in the fixed order s = 0,1,…; loads are issued four at a time so the L2 latencies overlap.
__device__ __forceinline__ float sum_partials_n(const float* __restrict__ P, long long stride, int splits, long long idx) {
    float acc = 0.f;
    int s = 0;
    for (; s + 4 <= splits; s += 4) {
        const float p0 = __ldcg(P + (s + 0) * stride + idx), p1 = __ldcg(P + (s + 1) * stride + idx);
        const float p2 = __ldcg(P + (s + 2) * stride + idx), p3 = __ldcg(P + (s + 3) * stride + idx);
        acc += p0; acc += p1; acc += p2; acc += p3;
    }
    for (; s < splits; ++s) acc += __ldcg(P + s * stride + idx);
    return acc;
}
__device__ __forceinline__ float4 sum_partials4_n(const float* __restrict__ P, long long stride, int splits, long long idx) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0;
    for (; s + 4 <= splits; s += 4) {
        const float4 p0 = __ldcg(reinterpret_cast<const float4*>(P + (s + 0) * stride + idx));
        const float4 p1 = __ldcg(reinterpret_cast<const float4*>(P + (s + 1) * stride + idx));
        const float4 p2 = __ldcg(reinterpret_cast<const float4*>(P + (s + 2) * stride + idx));
        const float4 p3 = __ldcg(reinterpret_cast<const float4*>(P + (s + 3) * stride + idx));
        acc.x += p0.x; acc.y += p0.y; acc.z += p0.z; acc.w += p0.w;
        acc.x += p1.x; acc.y += p1.y; acc.z += p1.z; acc.w += p1.w;
        acc.x += p2.x; acc.y += p2.y; acc.z += p2.z; acc.w += p2.w;
        acc.x += p3.x; acc.y += p3.y; acc.z += p3.z; acc.w += p3.w;
    }
    for (; s < splits; ++s) {
        const float4 p = __ldcg(reinterpret_cast<const float4*>(P + s * stride + idx));
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    return acc;
}

// idx = row·cols + col; the slot count depends on the 128-column tile of `col`
__device__ __forceinline__ float sum_partials(const PartialInfo& pi, long long idx, int col) {
    return sum_partials_n(pi.P, pi.stride, partial_count(pi, col), idx);
}
__device__ __forceinline__ float4 sum_partials4(const PartialInfo& pi, long long idx, int col) {
    return sum_partials4_n(pi.P, pi.stride, partial_count(pi, col), idx);
}

// act[r, i] = bf16( silu(g) * u ),  g = bf16(Σ partials[r, i]),  u = bf16(Σ partials[r, inter + i]); thread `tid` of `nthreads`
// (outputs [c_begin, c_begin + n_cols) of every row: the whole row for the stand-alone kernel, one 128-column weight tile = 64 outputs
// for the stream-K GEMM's in-kernel fix-up)
__device__ __forceinline__ void reduce_swiglu_cols(const PartialInfo& pi, bf16* __restrict__ act, int rows, int inter, int interleaved,
                                                   int c_begin, int n_cols, long long tid, long long nthreads) {
    const int i4 = n_cols / 4;
    const int up_off = interleaved ? 32 : inter;
    const long long total = static_cast<long long>(rows) * i4;
    for (long long i = tid; i < total; i += nthreads) {
        const long long r = i / i4, c = c_begin + (i % i4) * 4;
        const long long gc = gate_col(c, interleaved);
        // gate and up partials of the same slots are fetched together (one L2 round trip per four slots instead of two); each
        // element is still summed in the order s = 0,1,…
        const long long ig = r * 2 * inter + gc, iu = ig + up_off;
        const int cg = partial_count(pi, static_cast<int>(gc)), cu = partial_count(pi, static_cast<int>(gc + up_off));
        float gf[4] = {0.f, 0.f, 0.f, 0.f}, uf[4] = {0.f, 0.f, 0.f, 0.f};
        for (int s0 = 0; s0 < max(cg, cu); s0 += 4) {
            float4 pg[4], pu[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (s0 + j < cg) pg[j] = __ldcg(reinterpret_cast<const float4*>(pi.P + (s0 + j) * pi.stride + ig));
                if (s0 + j < cu) pu[j] = __ldcg(reinterpret_cast<const float4*>(pi.P + (s0 + j) * pi.stride + iu));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (s0 + j < cg) { gf[0] += pg[j].x; gf[1] += pg[j].y; gf[2] += pg[j].z; gf[3] += pg[j].w; }
                if (s0 + j < cu) { uf[0] += pu[j].x; uf[1] += pu[j].y; uf[2] += pu[j].z; uf[3] += pu[j].w; }
            }
        }
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gg = __bfloat162float(__float2bfloat16_rn(gf[j])), uu = __bfloat162float(__float2bfloat16_rn(uf[j]));
            o[j] = __fmul_rn(__fdiv_rn(gg, __fadd_rn(1.0f, expf(-gg))), uu);
        }
        *reinterpret_cast<uint2*>(act + r * inter + c) = make_uint2(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]));
    }
}
__device__ __forceinline__ void reduce_swiglu_part(const PartialInfo& pi, bf16* __restrict__ act, int rows, int inter, int interleaved,
                                                   long long tid, long long nthreads) {
    reduce_swiglu_cols(pi, act, rows, inter, interleaved, 0, inter, tid, nthreads);
}

// One warp per (sequence, head) — `gw` of n_seqs·n_heads: reduce the q/k/v partials, RoPE q and k, write q into the qkv buffer
// and k, v into the KV page of position positions[seq].
__device__ __forceinline__ void reduce_rope_kv_warp(const PartialInfo& pi, bf16* __restrict__ qkv, const int* __restrict__ positions,
                                                    bf16* __restrict__ kv_pages, const int* __restrict__ block_table, int max_pages,
                                                    int n_seqs, int n_heads, int head_dim, int page_size,
                                                    const float* __restrict__ rope_cos, const float* __restrict__ rope_sin, int gw, int lane) {
    if (gw >= n_seqs * n_heads) return;
    const int seq = gw / n_heads, head = gw % n_heads;
    const int hidden = n_heads * head_dim, half = head_dim / 2;
    const int pos = positions[seq];
    const int page = block_table[static_cast<size_t>(seq) * max_pages + pos / page_size];
    const int slot = pos % page_size;
    const long long rowbase = static_cast<long long>(seq) * 3 * hidden + head * head_dim;
    bf16* q = qkv + rowbase;
    bf16* kdst = kv_pages + (((static_cast<size_t>(page) * 2 + 0) * n_heads + head) * page_size + slot) * head_dim;
    bf16* vdst = kv_pages + (((static_cast<size_t>(page) * 2 + 1) * n_heads + head) * page_size + slot) * head_dim;
    const float* cs = rope_cos + static_cast<size_t>(pos) * half;
    const float* sn = rope_sin + static_cast<size_t>(pos) * half;
    const float* P = pi.P;
    const long long stride = pi.stride;
    const long long row0 = static_cast<long long>(seq) * 3 * hidden;
    // head_dim-aligned head slices never straddle a 128-column tile when head_dim divides 128 or is a multiple of it
    const int col_q = static_cast<int>(rowbase - row0);
    const int cnt_q = partial_count(pi, col_q), cnt_k = partial_count(pi, col_q + hidden);
    const bool uniform = (head_dim <= 128) && (128 % head_dim == 0);
    auto r2 = [&](long long off, float& a, float& b, int cnt_hint) {   // two adjacent reduced values, rounded to bf16 like the GEMM output
        float x0 = 0.f, x1 = 0.f;
        const int splits = uniform ? cnt_hint : partial_count(pi, static_cast<int>(off - row0));
        int sp = 0;
        for (; sp + 4 <= splits; sp += 4) {
            const float2 p0 = __ldcg(reinterpret_cast<const float2*>(P + (sp + 0) * stride + off));
            const float2 p1 = __ldcg(reinterpret_cast<const float2*>(P + (sp + 1) * stride + off));
            const float2 p2 = __ldcg(reinterpret_cast<const float2*>(P + (sp + 2) * stride + off));
            const float2 p3 = __ldcg(reinterpret_cast<const float2*>(P + (sp + 3) * stride + off));
            x0 += p0.x; x1 += p0.y; x0 += p1.x; x1 += p1.y; x0 += p2.x; x1 += p2.y; x0 += p3.x; x1 += p3.y;
        }
        for (; sp < splits; ++sp) {
            const float2 p = __ldcg(reinterpret_cast<const float2*>(P + sp * stride + off));
            x0 += p.x; x1 += p.y;
        }
        a = __bfloat162float(__float2bfloat16_rn(x0));
        b = __bfloat162float(__float2bfloat16_rn(x1));
    };
    // All loads of an iteration are issued before its first store (the compiler cannot prove that the stores into qkv /
    // the KV page do not alias the partials, so interleaving them would serialise five L2 round trips per thread).
    for (int i = lane * 2; i < half; i += 64) {
        const float c0 = cs[i], c1 = cs[i + 1], s0 = sn[i], s1 = sn[i + 1];
        float qa0, qa1, qb0, qb1, ka0, ka1, kb0, kb1;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int vi = 2 * i;                                   // lane·4: this lane's four v columns of the same pass
        const bool has_v = vi < head_dim;
        if (uniform) {
            // the five reductions of this lane (q lo/hi, k lo/hi, v) share ONE slot loop: four slots × five loads in flight per
            // round trip instead of five loops one after the other (this kernel is pure L2 latency); per element the order is
            // still s = 0,1,…
            const int cnt_v = has_v ? partial_count(pi, col_q + 2 * hidden) : 0;
            const long long oq = rowbase + i, ok = rowbase + hidden + i, ov = rowbase + 2 * hidden + vi;
            float xq[4] = {0.f, 0.f, 0.f, 0.f}, xk[4] = {0.f, 0.f, 0.f, 0.f};      // {lo.x, lo.y, hi.x, hi.y}
            for (int sp = 0; sp < max(max(cnt_q, cnt_k), cnt_v); sp += 4) {
                float2 pq[4][2], pk[4][2];
                float4 pv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float* Ps = P + (sp + j) * stride;
                    if (sp + j < cnt_q) { pq[j][0] = __ldcg(reinterpret_cast<const float2*>(Ps + oq)); pq[j][1] = __ldcg(reinterpret_cast<const float2*>(Ps + oq + half)); }
                    if (sp + j < cnt_k) { pk[j][0] = __ldcg(reinterpret_cast<const float2*>(Ps + ok)); pk[j][1] = __ldcg(reinterpret_cast<const float2*>(Ps + ok + half)); }
                    if (sp + j < cnt_v) pv[j] = __ldcg(reinterpret_cast<const float4*>(Ps + ov));
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (sp + j < cnt_q) { xq[0] += pq[j][0].x; xq[1] += pq[j][0].y; xq[2] += pq[j][1].x; xq[3] += pq[j][1].y; }
                    if (sp + j < cnt_k) { xk[0] += pk[j][0].x; xk[1] += pk[j][0].y; xk[2] += pk[j][1].x; xk[3] += pk[j][1].y; }
                    if (sp + j < cnt_v) { v.x += pv[j].x; v.y += pv[j].y; v.z += pv[j].z; v.w += pv[j].w; }
                }
            }
            qa0 = __bfloat162float(__float2bfloat16_rn(xq[0])); qa1 = __bfloat162float(__float2bfloat16_rn(xq[1]));
            qb0 = __bfloat162float(__float2bfloat16_rn(xq[2])); qb1 = __bfloat162float(__float2bfloat16_rn(xq[3]));
            ka0 = __bfloat162float(__float2bfloat16_rn(xk[0])); ka1 = __bfloat162float(__float2bfloat16_rn(xk[1]));
            kb0 = __bfloat162float(__float2bfloat16_rn(xk[2])); kb1 = __bfloat162float(__float2bfloat16_rn(xk[3]));
        } else {
            r2(rowbase + i, qa0, qa1, cnt_q);
            r2(rowbase + i + half, qb0, qb1, cnt_q);
            r2(rowbase + hidden + i, ka0, ka1, cnt_k);
            r2(rowbase + hidden + i + half, kb0, kb1, cnt_k);
            if (has_v) v = sum_partials4(pi, rowbase + 2 * hidden + vi, static_cast<int>(rowbase - row0) + 2 * hidden + vi);
        }
        // explicit mul / add roundings (x·cos ± y·sin as HF computes it: two products, one sum): no FMA contraction, so that
        // this code gives the same bits in whichever kernel it is inlined (stand-alone glue kernel, chain kernel)
        auto rot_lo = [](float x, float y, float c, float s) { return __fsub_rn(__fmul_rn(x, c), __fmul_rn(y, s)); };
        auto rot_hi = [](float x, float y, float c, float s) { return __fadd_rn(__fmul_rn(y, c), __fmul_rn(x, s)); };
        *reinterpret_cast<uint32_t*>(q + i) = pack_bf16x2(rot_lo(qa0, qb0, c0, s0), rot_lo(qa1, qb1, c1, s1));
        *reinterpret_cast<uint32_t*>(q + i + half) = pack_bf16x2(rot_hi(qa0, qb0, c0, s0), rot_hi(qa1, qb1, c1, s1));
        *reinterpret_cast<uint32_t*>(kdst + i) = pack_bf16x2(rot_lo(ka0, kb0, c0, s0), rot_lo(ka1, kb1, c1, s1));
        *reinterpret_cast<uint32_t*>(kdst + i + half) = pack_bf16x2(rot_hi(ka0, kb0, c0, s0), rot_hi(ka1, kb1, c1, s1));
        if (has_v) *reinterpret_cast<uint2*>(vdst + vi) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
    }
}

// x[row] = bf16(Σ partials + x[row]);  y[row] = RMSNorm(x[row]) · w — one row by 256 threads (8 warps) of one CTA, reproducing
// BIT FOR BIT the summation tree of reduce_residual_rmsnorm_kernel (kernels_misc.cu: a cluster of 8 CTAs × 128 threads per
// row, CTA r owning columns [r·d/8, (r+1)·d/8)): warp w here plays CTA rank w, and walks that CTA's four warps one after the
// other — per-thread sums over its ≤ 2 float4 groups, xor-butterfly over the 32 lanes, the four warp sums added in order,
// the eight CTA sums added in order.  `t256` = thread index in the 256-thread group, `cta_part` = 8 floats of shared memory;
// `sync256` synchronises exactly those 256 threads.
template <typename Sync>
__device__ __forceinline__ void reduce_residual_rmsnorm_row(const PartialInfo& pi, bf16* __restrict__ x, const bf16* __restrict__ w,
                                                            bf16* __restrict__ y, int row, int d, float eps, int t256, float* cta_part,
                                                            Sync sync256) {
    constexpr int VC = 8, VT = 128, MAXV = 2;            // the stand-alone kernel's RN_CLUSTER, RN_THREADS, RN_MAXV
    const int rank = t256 >> 5, lane = t256 & 31;
    const int cols = d / VC;
    const long long base = static_cast<long long>(row) * d + static_cast<long long>(rank) * cols;
    float vals[4][MAXV][4];
    uint2 wreg[4][MAXV], xreg[4][MAXV];
    int cnt[4][MAXV];
    int max_cnt = 0;
    // all loads of the row slice first: x, w, then the partial slots with the slot loop OUTERMOST, so that the thread's (up to
    // eight) column groups are in flight together — one L2 round trip per slot instead of one per (group, slot)
#pragma unroll
    for (int vw = 0; vw < 4; ++vw) {
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            const int c = (v * VT + vw * 32 + lane) * 4;
            cnt[vw][v] = 0;
            vals[vw][v][0] = vals[vw][v][1] = vals[vw][v][2] = vals[vw][v][3] = 0.f;
            if (c < cols) {
                xreg[vw][v] = *reinterpret_cast<const uint2*>(x + base + c);
                wreg[vw][v] = *reinterpret_cast<const uint2*>(w + static_cast<long long>(rank) * cols + c);
                cnt[vw][v] = partial_count(pi, rank * cols + c);
                max_cnt = max(max_cnt, cnt[vw][v]);
            }
        }
    }
#pragma unroll 2
    for (int s = 0; s < max_cnt; ++s) {
#pragma unroll
        for (int vw = 0; vw < 4; ++vw) {
#pragma unroll
            for (int v = 0; v < MAXV; ++v) {
                if (s < cnt[vw][v]) {                    // same order s = 0,1,… per element as sum_partials4_n
                    const int c = (v * VT + vw * 32 + lane) * 4;
                    const float4 p = __ldcg(reinterpret_cast<const float4*>(pi.P + s * pi.stride + base + c));
                    vals[vw][v][0] += p.x; vals[vw][v][1] += p.y; vals[vw][v][2] += p.z; vals[vw][v][3] += p.w;
                }
            }
        }
    }
    float parts[4];
#pragma unroll
    for (int vw = 0; vw < 4; ++vw) {
        float sq = 0.f;
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            if (cnt[vw][v] > 0) {
                const uint2 r = xreg[vw][v];
                vals[vw][v][0] = __bfloat162float(__float2bfloat16_rn(vals[vw][v][0] + bf16_lo(r.x)));
                vals[vw][v][1] = __bfloat162float(__float2bfloat16_rn(vals[vw][v][1] + bf16_hi(r.x)));
                vals[vw][v][2] = __bfloat162float(__float2bfloat16_rn(vals[vw][v][2] + bf16_lo(r.y)));
                vals[vw][v][3] = __bfloat162float(__float2bfloat16_rn(vals[vw][v][3] + bf16_hi(r.y)));
#pragma unroll
                for (int j = 0; j < 4; ++j) sq = __fmaf_rn(vals[vw][v][j], vals[vw][v][j], sq);
            }
        }
        parts[vw] = warp_sum(sq);
    }
    if (lane == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) t += parts[i];
        cta_part[rank] = t;
    }
    sync256();
    float tot = 0.f;
#pragma unroll
    for (int r = 0; r < VC; ++r) tot += cta_part[r];
    sync256();                                           // cta_part may be rewritten by the next row
    const float rstd = __frcp_rn(__fsqrt_rn(__fadd_rn(__fdiv_rn(tot, static_cast<float>(d)), eps)));
#pragma unroll
    for (int vw = 0; vw < 4; ++vw) {
        const int vt = vw * 32 + lane;
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            const int c = (v * VT + vt) * 4;
            if (cnt[vw][v] > 0) {
                const uint2 wv = wreg[vw][v];
                const float wf[4] = {bf16_lo(wv.x), bf16_hi(wv.x), bf16_lo(wv.y), bf16_hi(wv.y)};
                const float* vv = vals[vw][v];
                *reinterpret_cast<uint2*>(x + base + c) = make_uint2(pack_bf16x2(vv[0], vv[1]), pack_bf16x2(vv[2], vv[3]));
                *reinterpret_cast<uint2*>(y + base + c) = make_uint2(pack_bf16x2((vv[0] * rstd) * wf[0], (vv[1] * rstd) * wf[1]),
                                                                     pack_bf16x2((vv[2] * rstd) * wf[2], (vv[3] * rstd) * wf[3]));
            }
        }
    }
}

// out[i] = Σ_s P[s][i]  (fp32 logits: the reduction splitk_reduce_kernel does for the lm_head GEMM); thread `tid` of `nthreads`
__device__ __forceinline__ void reduce_logits_part(const PartialInfo& pi, float* __restrict__ out, int rows, int cols, long long tid,
                                                   long long nthreads) {
    const long long total = static_cast<long long>(rows) * cols;
    for (long long i = tid; i < total; i += nthreads) {
        const int c = static_cast<int>(i % cols);
        const int n = partial_count(pi, c);
        float acc = 0.f;
        for (int s = 0; s < n; ++s) acc += __ldcg(pi.P + s * pi.stride + i);
        out[i] = acc;
    }
}

}  // namespace teo
