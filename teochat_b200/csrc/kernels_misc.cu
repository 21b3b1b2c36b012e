// HBM-bound helper kernels of the hot path: synthetic init, patchify (normalise + im2col),
// embeddings + LayerNorm, RMSNorm, SwiGLU, RoPE + KV-page scatter, splice gather, greedy argmax.
// All are one-pass, 16-byte-vectorised where the layout allows, fp32 math on bf16 storage.
#include <cooperative_groups.h>
#include <stdlib.h>

#include <algorithm>

#include "common.h"
#include "decode_reduce.cuh"
#include "ptx.cuh"

namespace teo {

// ------------------------------------------------------------------------------ synthetic init
__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float hash_normal_at(uint64_t seed, size_t i, float scale, float mean) {
    const uint64_t x = splitmix64(seed + (static_cast<uint64_t>(i) + 1ULL) * 0x9E3779B97F4A7C15ULL);
    const int s = static_cast<int>(x & 0xFFFF) + static_cast<int>((x >> 16) & 0xFFFF) +
                  static_cast<int>((x >> 32) & 0xFFFF) + static_cast<int>((x >> 48) & 0xFFFF);
    return __fadd_rn(mean, __fmul_rn(static_cast<float>(s - 131070), scale));   // no FMA contraction
}
template <typename T>
__global__ void init_normal_hash_kernel(T* out, size_t n, uint64_t seed, float scale, float mean) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float v = hash_normal_at(seed, i, scale, mean);
        if constexpr (sizeof(T) == 2) out[i] = __float2bfloat16_rn(v);
        else out[i] = v;
    }
}
__global__ void init_u8_hash_kernel(uint8_t* out, size_t n, uint64_t seed) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = static_cast<uint8_t>(splitmix64(seed + (static_cast<uint64_t>(i) + 1ULL) * 0x9E3779B97F4A7C15ULL) >> 56);
}

// ------------------------------------------------------------------------------ patchify
// One block per (frame, patch row): stage the P image rows in shared memory with coalesced loads,
// then emit g patches × kpad bf16 columns, column index (c*P + ky)*P + kx.
// PLANES = 3 (exact mode, exact.cu): every value is written as three bf16 terms hi | mid | lo, kpad columns apart, in rows
// of 3·kpad — the split-operand form of the exact-mode GEMM.
template <bool U8, int PLANES>
__global__ void patchify_kernel(const void* __restrict__ in, bf16* __restrict__ patches, int image, int patch, int kpad) {
    extern __shared__ float tile[];   // [3][P][image] normalised values
    const int g = image / patch;
    const int frame = blockIdx.x / g, py = blockIdx.x % g;
    const int row_elems = image * 3;
    const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
    const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
    if constexpr (U8) {
        const uint8_t* src = static_cast<const uint8_t*>(in) + (static_cast<size_t>(frame) * image + static_cast<size_t>(py) * patch) * row_elems;
        for (int i = threadIdx.x; i < patch * row_elems; i += blockDim.x) {
            const int ky = i / row_elems, r = i % row_elems, x = r / 3, c = r % 3;
            const float v = __fdiv_rn(static_cast<float>(src[i]), 255.0f);      // ToTensor
            tile[(c * patch + ky) * image + x] = __fdiv_rn(v - mean[c], stdv[c]);   // Normalize
        }
    } else {
        const float* src = static_cast<const float*>(in) + static_cast<size_t>(frame) * 3 * image * image;
        for (int i = threadIdx.x; i < 3 * patch * image; i += blockDim.x) {
            const int c = i / (patch * image), r = i % (patch * image), ky = r / image, x = r % image;
            tile[(c * patch + ky) * image + x] = src[(static_cast<size_t>(c) * image + py * patch + ky) * image + x];
        }
    }
    __syncthreads();
    const int pdim = 3 * patch * patch;
    bf16* dst = patches + (static_cast<size_t>(frame) * g * g + static_cast<size_t>(py) * g) * kpad * PLANES;
    for (int i = threadIdx.x; i < g * kpad; i += blockDim.x) {
        const int px = i / kpad, col = i % kpad;
        float v = 0.f;
        if (col < pdim) {
            const int c = col / (patch * patch), r = col % (patch * patch), ky = r / patch, kx = r % patch;
            v = tile[(c * patch + ky) * image + px * patch + kx];
        }
        bf16* o = dst + static_cast<size_t>(px) * kpad * PLANES + col;
        const bf16 hi = __float2bfloat16_rn(v);
        o[0] = hi;
        if constexpr (PLANES == 3) {
            float r = v - __bfloat162float(hi);
            const bf16 mid = __float2bfloat16_rn(r);
            r -= __bfloat162float(mid);
            o[kpad] = mid;
            o[2 * kpad] = __float2bfloat16_rn(r);
        }
    }
}

// The uint8 NHWC path of the bench (PLANES = 1), vectorised: the P image rows of a patch row are one contiguous run of bytes —
// staged with 16-byte loads, kept as BYTES in shared memory — and every thread emits eight consecutive im2col columns as one
// 16-byte store.  ToTensor + Normalize is a per-channel table of the 256 possible results, built with the same two IEEE divisions
// as the scalar kernel (bit-identical values).  The scalar kernel above moved one byte in and one bf16 out per thread and iteration,
// with two IEEE divisions per pixel value: 264 µs for 256 frames, 7 % of what the 122 MB it touches cost at HBM speed.
__global__ void __launch_bounds__(256) patchify_u8_vec_kernel(const uint8_t* __restrict__ in, bf16* __restrict__ patches, int image, int patch,
                                                              int kpad) {
    extern __shared__ __align__(16) uint8_t raw[];          // [patch rows][image][3] bytes, then the table
    const int g = image / patch;
    const int frame = blockIdx.x / g, py = blockIdx.x % g;
    const int row_elems = image * 3, n_bytes = patch * row_elems;
    float* lut = reinterpret_cast<float*>(raw + ((n_bytes + 15) & ~15));     // [3][256]
    const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};
    const float stdv[3] = {0.26862954f, 0.26130258f, 0.27577711f};
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
        const int c = i >> 8;
        const float v = __fdiv_rn(static_cast<float>(i & 255), 255.0f);      // ToTensor
        lut[i] = __fdiv_rn(v - mean[c], stdv[c]);                            // Normalize
    }
    const uint8_t* src = in + (static_cast<size_t>(frame) * image + static_cast<size_t>(py) * patch) * row_elems;
    for (int i = threadIdx.x; i < n_bytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(raw)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
    __syncthreads();
    const int pdim = 3 * patch * patch, pp = patch * patch, k8 = kpad / 8;
    bf16* dst = patches + (static_cast<size_t>(frame) * g * g + static_cast<size_t>(py) * g) * kpad;
    for (int i = threadIdx.x; i < g * k8; i += blockDim.x) {
        const int px = i / k8, col0 = (i % k8) * 8;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = col0 + j;
            v[j] = 0.f;
            if (col < pdim) {
                const int c = col / pp, r = col - c * pp, ky = r / patch, kx = r - ky * patch;
                v[j] = lut[c * 256 + raw[ky * row_elems + (px * patch + kx) * 3 + c]];
            }
        }
        *reinterpret_cast<uint4*>(dst + static_cast<size_t>(px) * kpad + col0) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
}

// ------------------------------------------------------------------------------ row-wise norms
// TPR threads cooperate on one row; each holds up to MAXV 8-element vectors in registers.
template <int TPR>
__device__ __forceinline__ float row_sum(float v, float* smem_red) {
    v = warp_sum(v);
    if constexpr (TPR > 32) {
        const int w = (threadIdx.x % TPR) >> 5, row_in_block = threadIdx.x / TPR;
        constexpr int WPR = TPR / 32;
        __syncthreads();
        if ((threadIdx.x & 31) == 0) smem_red[row_in_block * WPR + w] = v;
        __syncthreads();
        v = 0.f;
#pragma unroll
        for (int i = 0; i < WPR; ++i) v += smem_red[row_in_block * WPR + i];
    }
    return v;
}

constexpr int NORM_MAXV = 8;

__device__ __forceinline__ void load8(const bf16* p, float (&x)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { x[2 * j] = bf16_lo(w[j]); x[2 * j + 1] = bf16_hi(w[j]); }
}
__device__ __forceinline__ void store8(bf16* p, const float (&x)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]),
                                              pack_bf16x2(x[6], x[7]));
}

// MODE 0: LayerNorm(x)            (x row = in[row])
// MODE 1: ViT embeddings: x row = (tok==0 ? cls : patch_out[frame*np + tok-1]) + pos[tok], then LayerNorm
// MODE 2: RMSNorm(x)
// MAXV = 8-element vectors a thread may hold: 4 covers d ≤ 4·8·TPR (CLIP-L's 1024 at TPR 32, LLaMA-2-7B's 4096 at TPR 128) at half the
// registers of the general 8 — more resident blocks, more bytes in flight on a kernel that does nothing but stream its rows.
template <int TPR, int MODE, int MAXV>
__global__ void norm_kernel(const bf16* __restrict__ in, const bf16* __restrict__ w, const bf16* __restrict__ b,
                            bf16* __restrict__ out, int rows, int d, float eps, const bf16* __restrict__ cls,
                            const bf16* __restrict__ pos, int n_patches) {
    __shared__ float red[32];
    pdl_trigger();
    pdl_wait();
    constexpr int RPB = 128 / TPR;   // rows per 128-thread block
    const int row = blockIdx.x * RPB + threadIdx.x / TPR;
    const int t = threadIdx.x % TPR;
    const bool active = row < rows;
    float x[MAXV][8];
    float sum = 0.f;
    const bf16* src = in;
    const bf16* posrow = nullptr;
    if (active) {
        if constexpr (MODE == 1) {
            const int tok = row % (n_patches + 1), frame = row / (n_patches + 1);
            src = tok == 0 ? cls : in + (static_cast<size_t>(frame) * n_patches + tok - 1) * d;
            posrow = pos + static_cast<size_t>(tok) * d;
        } else {
            src = in + static_cast<size_t>(row) * d;
        }
    }
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
        const int c = (v * TPR + t) * 8;
        if (active && c < d) {
            load8(src + c, x[v]);
            if constexpr (MODE == 1) {
                float p[8];
                load8(posrow + c, p);
#pragma unroll
                for (int j = 0; j < 8; ++j) x[v][j] += p[j];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += (MODE == 2) ? x[v][j] * x[v][j] : x[v][j];
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[v][j] = 0.f;
        }
    }
    sum = row_sum<TPR>(sum, red);
    float mean = 0.f, rstd;
    if constexpr (MODE == 2) {
        rstd = 1.0f / sqrtf(sum / static_cast<float>(d) + eps);
    } else {
        mean = sum / static_cast<float>(d);
        float sq = 0.f;
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            const int c = (v * TPR + t) * 8;
            if (active && c < d) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float dlt = x[v][j] - mean; sq += dlt * dlt; }
            }
        }
        sq = row_sum<TPR>(sq, red);
        rstd = 1.0f / sqrtf(sq / static_cast<float>(d) + eps);
    }
    if (!active) return;
    bf16* dst = out + static_cast<size_t>(row) * d;
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
        const int c = (v * TPR + t) * 8;
        if (c < d) {
            float wv[8], y[8];
            load8(w + c, wv);
            if constexpr (MODE == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = (x[v][j] * rstd) * wv[j];
            } else {
                float bv[8];
                load8(b + c, bv);
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = (x[v][j] - mean) * rstd * wv[j] + bv[j];
            }
            store8(dst + c, y);
        }
    }
}

template <int MODE>
static int launch_norm(const bf16* in, const bf16* w, const bf16* b, bf16* out, int rows, int d, float eps, const bf16* cls,
                       const bf16* pos, int n_patches, cudaStream_t stream) {
    TEO_CHECK_ARG(rows > 0 && d > 0 && d % 8 == 0, "norm: rows=%d d=%d (d must be a positive multiple of 8)", rows, d);
    TEO_CHECK_ARG(d <= 128 * 8 * NORM_MAXV, "norm: d=%d exceeds %d", d, 128 * 8 * NORM_MAXV);
    if (d <= 32 * 8 * 4) {
        TEO_CUDA(launch_k(norm_kernel<32, MODE, 4>, dim3((rows + 3) / 4), dim3(128), 0, stream, in, w, b, out, rows, d, eps, cls, pos, n_patches));
    } else if (d <= 128 * 8 * 4) {
        TEO_CUDA(launch_k(norm_kernel<128, MODE, 4>, dim3(rows), dim3(128), 0, stream, in, w, b, out, rows, d, eps, cls, pos, n_patches));
    } else {
        TEO_CUDA(launch_k(norm_kernel<128, MODE, NORM_MAXV>, dim3(rows), dim3(128), 0, stream, in, w, b, out, rows, d, eps, cls, pos, n_patches));
    }
    TEO_LAUNCH_CHECK("norm_kernel");
    return TEO_OK;
}

// (Σx, Σx²) of every bf16 row into slot 0 of stats[row][slots][2] (other slots zero): the row statistics a folded LayerNorm
// (gemm_common.cuh) reads, for rows that no GEMM epilogue produced.  One warp per row.
__global__ void row_stats_kernel(const bf16* __restrict__ x, float* __restrict__ stats, int rows, int d, int slots) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const bf16* src = x + static_cast<size_t>(row) * d;
    float s1 = 0.f, s2 = 0.f;
    for (int c = lane * 8; c < d; c += 256) {
        float v[8];
        load8(src + c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { s1 += v[j]; s2 = fmaf(v[j], v[j], s2); }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    float* so = stats + static_cast<size_t>(row) * slots * 2;
    for (int i = lane; i < 2 * slots; i += 32) so[i] = i == 0 ? s1 : (i == 1 ? s2 : 0.f);
}

// ------------------------------------------------------------------------------ small copies / elementwise
__global__ void drop_cls_kernel(const uint4* __restrict__ hidden, uint4* __restrict__ feats, int n_patches, int d8, size_t total) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t row = i / d8, c = i % d8;
        const size_t frame = row / n_patches, tok = row % n_patches;
        feats[i] = hidden[(frame * (n_patches + 1) + tok + 1) * d8 + c];
    }
}

__global__ void swiglu_kernel(const bf16* __restrict__ gate_up, bf16* __restrict__ out, int rows, int inter, int interleaved) {
    const int i8 = inter / 8;
    const int up_off = interleaved ? 32 : inter;
    const size_t total = static_cast<size_t>(rows) * i8;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = i / i8, c = (i % i8) * 8;
        const size_t gc = static_cast<size_t>(gate_col(static_cast<long long>(c), interleaved));
        float g[8], u[8], y[8];
        load8(gate_up + r * 2 * inter + gc, g);
        load8(gate_up + r * 2 * inter + gc + up_off, u);
#pragma unroll
        for (int j = 0; j < 8; ++j) y[j] = silu_mul_fast(g[j], u[j]);
        store8(out + r * inter + c, y);
    }
}

__global__ void splice_embed_kernel(const bf16* __restrict__ embed, const bf16* __restrict__ feats, const int* __restrict__ src,
                                    bf16* __restrict__ out, int tokens, int d8) {
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x;
    if (row >= tokens) return;
    const int s = src[row];
    const uint4* from = reinterpret_cast<const uint4*>(s >= 0 ? embed + static_cast<size_t>(s) * d8 * 8
                                                              : feats + static_cast<size_t>(-(s + 1)) * d8 * 8);
    uint4* to = reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * d8 * 8);
    for (int c = threadIdx.x; c < d8; c += blockDim.x) to[c] = from[c];
}

// ------------------------------------------------------------------------------ RoPE + KV scatter
// One warp per (token, head).  rotate-half: out[i] = x[i]cos_i - x[i+h/2]sin_i; out[i+h/2] = x[i+h/2]cos_i + x[i]sin_i.
// cos/sin tables f32 [max_pos, hd/2] are built on the host exactly like HF's cached tables.
__global__ void rope_kv_write_kernel(bf16* __restrict__ qkv, const int* __restrict__ positions, const int* __restrict__ seq_ids,
                                     bf16* __restrict__ kv_pages, const int* __restrict__ block_table, int max_pages, int tokens,
                                     int n_heads, int head_dim, int page_size, const float* __restrict__ rope_cos,
                                     const float* __restrict__ rope_sin) {
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= tokens * n_heads) return;
    const int tok = gw / n_heads, head = gw % n_heads;
    const int hidden = n_heads * head_dim, half = head_dim / 2;
    const int pos = positions[tok];
    const int seq = seq_ids ? seq_ids[tok] : tok;
    const int page = block_table[static_cast<size_t>(seq) * max_pages + pos / page_size];
    const int slot = pos % page_size;
    bf16* q = qkv + static_cast<size_t>(tok) * 3 * hidden + head * head_dim;
    bf16* k = q + hidden;
    const bf16* v = k + hidden;
    // page layout [page][2][head][slot][dim]
    bf16* kdst = kv_pages + (((static_cast<size_t>(page) * 2 + 0) * n_heads + head) * page_size + slot) * head_dim;
    bf16* vdst = kv_pages + (((static_cast<size_t>(page) * 2 + 1) * n_heads + head) * page_size + slot) * head_dim;
    const float* cs = rope_cos + static_cast<size_t>(pos) * half;
    const float* sn = rope_sin + static_cast<size_t>(pos) * half;
    for (int i = lane * 2; i < half; i += 64) {
        const float c0 = cs[i], c1 = cs[i + 1], s0 = sn[i], s1 = sn[i + 1];
        {
            const uint32_t lo = *reinterpret_cast<const uint32_t*>(q + i), hi = *reinterpret_cast<const uint32_t*>(q + i + half);
            const float a0 = bf16_lo(lo), a1 = bf16_hi(lo), b0 = bf16_lo(hi), b1 = bf16_hi(hi);
            *reinterpret_cast<uint32_t*>(q + i) = pack_bf16x2(a0 * c0 - b0 * s0, a1 * c1 - b1 * s1);
            *reinterpret_cast<uint32_t*>(q + i + half) = pack_bf16x2(b0 * c0 + a0 * s0, b1 * c1 + a1 * s1);
        }
        {
            const uint32_t lo = *reinterpret_cast<const uint32_t*>(k + i), hi = *reinterpret_cast<const uint32_t*>(k + i + half);
            const float a0 = bf16_lo(lo), a1 = bf16_hi(lo), b0 = bf16_lo(hi), b1 = bf16_hi(hi);
            const uint32_t o_lo = pack_bf16x2(a0 * c0 - b0 * s0, a1 * c1 - b1 * s1);
            const uint32_t o_hi = pack_bf16x2(b0 * c0 + a0 * s0, b1 * c1 + a1 * s1);
            *reinterpret_cast<uint32_t*>(k + i) = o_lo;
            *reinterpret_cast<uint32_t*>(k + i + half) = o_hi;
            *reinterpret_cast<uint32_t*>(kdst + i) = o_lo;
            *reinterpret_cast<uint32_t*>(kdst + i + half) = o_hi;
        }
    }
    for (int i = lane * 2; i < head_dim; i += 64)
        *reinterpret_cast<uint32_t*>(vdst + i) = *reinterpret_cast<const uint32_t*>(v + i);
}

// Same operation with 16-byte accesses (head_dim % 16 == 0, the prefill path: 68 k tokens × 32 heads per layer): one thread
// owns 8 adjacent dims of the low half and the matching 8 of the high half of one (token, head), rotates q and k, and
// copies 16 dims of v.  All ten loads are issued before the first store.
__device__ __forceinline__ void rope8(const uint4& lo, const uint4& hi, const float* c, const float* s, uint4& o_lo, uint4& o_hi) {
    const uint32_t l[4] = {lo.x, lo.y, lo.z, lo.w}, h[4] = {hi.x, hi.y, hi.z, hi.w};
    uint32_t ol[4], oh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a0 = bf16_lo(l[j]), a1 = bf16_hi(l[j]), b0 = bf16_lo(h[j]), b1 = bf16_hi(h[j]);
        const float c0 = c[2 * j], c1 = c[2 * j + 1], s0 = s[2 * j], s1 = s[2 * j + 1];
        ol[j] = pack_bf16x2(a0 * c0 - b0 * s0, a1 * c1 - b1 * s1);
        oh[j] = pack_bf16x2(b0 * c0 + a0 * s0, b1 * c1 + a1 * s1);
    }
    o_lo = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    o_hi = make_uint4(oh[0], oh[1], oh[2], oh[3]);
}
__global__ void __launch_bounds__(256)
rope_kv_write_vec_kernel(bf16* __restrict__ qkv, const int* __restrict__ positions, const int* __restrict__ seq_ids,
                         bf16* __restrict__ kv_pages, const int* __restrict__ block_table, int max_pages, long long items,
                         int n_heads, int head_dim, int page_size, const float* __restrict__ rope_cos,
                         const float* __restrict__ rope_sin) {
    const int half = head_dim / 2, tpi = half / 8;                 // threads per (token, head)
    const long long gt = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long item = gt / tpi;
    if (item >= items) return;
    const int part = static_cast<int>(gt % tpi);
    const int tok = static_cast<int>(item / n_heads), head = static_cast<int>(item % n_heads);
    const int hidden = n_heads * head_dim;
    const int pos = positions[tok];
    const int seq = seq_ids ? seq_ids[tok] : tok;
    const int page = block_table[static_cast<size_t>(seq) * max_pages + pos / page_size];
    const int slot = pos % page_size;
    bf16* q = qkv + static_cast<size_t>(tok) * 3 * hidden + head * head_dim + part * 8;
    bf16* k = q + hidden;
    const bf16* v = qkv + static_cast<size_t>(tok) * 3 * hidden + 2 * hidden + head * head_dim + part * 16;
    bf16* kdst = kv_pages + (((static_cast<size_t>(page) * 2 + 0) * n_heads + head) * page_size + slot) * head_dim + part * 8;
    bf16* vdst = kv_pages + (((static_cast<size_t>(page) * 2 + 1) * n_heads + head) * page_size + slot) * head_dim + part * 16;
    const float4* cs = reinterpret_cast<const float4*>(rope_cos + static_cast<size_t>(pos) * half + part * 8);
    const float4* sn = reinterpret_cast<const float4*>(rope_sin + static_cast<size_t>(pos) * half + part * 8);
    const uint4 q_lo = *reinterpret_cast<const uint4*>(q), q_hi = *reinterpret_cast<const uint4*>(q + half);
    const uint4 k_lo = *reinterpret_cast<const uint4*>(k), k_hi = *reinterpret_cast<const uint4*>(k + half);
    const uint4 v0 = *reinterpret_cast<const uint4*>(v), v1 = *reinterpret_cast<const uint4*>(v + 8);
    const float4 c0 = cs[0], c1 = cs[1], s0 = sn[0], s1 = sn[1];
    const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    uint4 o_lo, o_hi;
    rope8(q_lo, q_hi, c, sv, o_lo, o_hi);
    *reinterpret_cast<uint4*>(q) = o_lo;
    *reinterpret_cast<uint4*>(q + half) = o_hi;
    rope8(k_lo, k_hi, c, sv, o_lo, o_hi);
    *reinterpret_cast<uint4*>(k) = o_lo;
    *reinterpret_cast<uint4*>(k + half) = o_hi;
    *reinterpret_cast<uint4*>(kdst) = o_lo;
    *reinterpret_cast<uint4*>(kdst + half) = o_hi;
    *reinterpret_cast<uint4*>(vdst) = v0;
    *reinterpret_cast<uint4*>(vdst + 8) = v1;
}

// ------------------------------------------------------------------------------ decode-step fusions
// The decode GEMMs (M = batch) leave fp32 split-K partials P[s][rows][n]; these kernels reduce them in the
// fixed order s = 0,1,… and do the next element-wise stage in the same pass (same rounding points as the
// unfused chain: the reduced value is rounded to bf16 exactly where the GEMM epilogue would have).
// x[row] = bf16(Σ partials + x[row]);  y[row] = RMSNorm(x[row]) * w      (o_proj / down_proj → next norm)
// A thread-block cluster of RN_CLUSTER CTAs shares one row (bs=32 rows alone would occupy 32 of 148 SMs): each CTA
// reduces d/RN_CLUSTER columns with 16-byte loads, the per-CTA sums of squares are exchanged through distributed
// shared memory, then every CTA normalises its own columns.
constexpr int RN_CLUSTER = 8;
constexpr int RN_THREADS = 128;
constexpr int RN_MAXV = 2;               // float4 groups per thread: d ≤ RN_CLUSTER · RN_THREADS · 4 · RN_MAXV = 8192
__global__ void __cluster_dims__(RN_CLUSTER, 1, 1) __launch_bounds__(RN_THREADS)
reduce_residual_rmsnorm_kernel(PartialInfo pi, bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ y, int d, float eps) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float warp_part[RN_THREADS / 32];
    __shared__ float cta_part;
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x / RN_CLUSTER;
    const int rank = static_cast<int>(cluster.block_rank());
    const int cols = d / RN_CLUSTER;                 // columns owned by this CTA (multiple of 4)
    const long long base = static_cast<long long>(row) * d + static_cast<long long>(rank) * cols;
    float vals[RN_MAXV][4];
    uint2 wreg[RN_MAXV];                             // norm weights: fetched with the first wave of loads, used after the syncs
    float sq = 0.f;
#pragma unroll
    for (int v = 0; v < RN_MAXV; ++v) {
        const int c = (v * RN_THREADS + threadIdx.x) * 4;
        if (c < cols) {
            const uint2 r = *reinterpret_cast<const uint2*>(x + base + c);
            wreg[v] = *reinterpret_cast<const uint2*>(w + static_cast<long long>(rank) * cols + c);
            const float4 acc = sum_partials4(pi, base + c, rank * cols + c);
            // the residual stream is stored in bf16: round before the statistics, like the unfused chain
            vals[v][0] = __bfloat162float(__float2bfloat16_rn(acc.x + bf16_lo(r.x)));
            vals[v][1] = __bfloat162float(__float2bfloat16_rn(acc.y + bf16_hi(r.x)));
            vals[v][2] = __bfloat162float(__float2bfloat16_rn(acc.z + bf16_lo(r.y)));
            vals[v][3] = __bfloat162float(__float2bfloat16_rn(acc.w + bf16_hi(r.y)));
#pragma unroll
            for (int j = 0; j < 4; ++j) sq = __fmaf_rn(vals[v][j], vals[v][j], sq);      // explicit: same bits as reduce_residual_rmsnorm_row
        }
    }
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < RN_THREADS / 32; ++i) t += warp_part[i];
        cta_part = t;
    }
    cluster.sync();
    float tot = 0.f;
#pragma unroll
    for (int r = 0; r < RN_CLUSTER; ++r) tot += *cluster.map_shared_rank(&cta_part, r);     // fixed order: deterministic
    cluster.sync();                                   // nobody leaves while its shared memory may still be read
    const float rstd = __frcp_rn(__fsqrt_rn(__fadd_rn(__fdiv_rn(tot, static_cast<float>(d)), eps)));
#pragma unroll
    for (int v = 0; v < RN_MAXV; ++v) {
        const int c = (v * RN_THREADS + threadIdx.x) * 4;
        if (c < cols) {
            const uint2 wv = wreg[v];
            const float wf[4] = {bf16_lo(wv.x), bf16_hi(wv.x), bf16_lo(wv.y), bf16_hi(wv.y)};
            *reinterpret_cast<uint2*>(x + base + c) = make_uint2(pack_bf16x2(vals[v][0], vals[v][1]), pack_bf16x2(vals[v][2], vals[v][3]));
            *reinterpret_cast<uint2*>(y + base + c) = make_uint2(pack_bf16x2((vals[v][0] * rstd) * wf[0], (vals[v][1] * rstd) * wf[1]),
                                                                 pack_bf16x2((vals[v][2] * rstd) * wf[2], (vals[v][3] * rstd) * wf[3]));
        }
    }
}

// act[r, i] = bf16( silu(g) * u ),  g = bf16(Σ partials[r, i]),  u = bf16(Σ partials[r, inter + i])
// One group of four outputs per thread, as many CTAs as that takes: capping the grid at two CTAs per SM (so that the next
// GEMM's CTAs find room beside this kernel's) measured SLOWER, 9.72 → 9.93 ms per decode step.
__global__ void reduce_swiglu_kernel(PartialInfo pi, bf16* __restrict__ act, int rows, int inter, int interleaved) {
    pdl_trigger();
    pdl_wait();
    reduce_swiglu_part(pi, act, rows, inter, interleaved, static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x,
                       static_cast<long long>(gridDim.x) * blockDim.x);
}

// One warp per (sequence, head): reduce the q/k/v partials, RoPE q and k, write q into the qkv buffer and k, v
// into the KV page of position seq_lens[seq].
__global__ void reduce_rope_kv_write_kernel(PartialInfo pi, bf16* __restrict__ qkv,
                                            const int* __restrict__ positions, bf16* __restrict__ kv_pages,
                                            const int* __restrict__ block_table, int max_pages, int n_seqs, int n_heads, int head_dim,
                                            int page_size, const float* __restrict__ rope_cos, const float* __restrict__ rope_sin) {
    pdl_trigger();
    pdl_wait();
    reduce_rope_kv_warp(pi, qkv, positions, kv_pages, block_table, max_pages, n_seqs, n_heads, head_dim, page_size, rope_cos, rope_sin,
                        (blockIdx.x * blockDim.x + threadIdx.x) >> 5, threadIdx.x & 31);
}

// ------------------------------------------------------------------------------ greedy argmax + stop rule
// One block per sequence.  Ties → lowest index (torch.argmax semantics).  When step_ptr != NULL
// the column is read from device memory (graph replay) and block 0 bumps it afterwards.
__global__ void argmax_step_kernel(const float* __restrict__ logits, int vocab, uint8_t* finished, int* tokens, int max_new,
                                   int step_host, int* step_ptr, int* next_ids, int* seq_lens, int eos_id) {
    __shared__ float bv[32];
    __shared__ int bi[32];
    pdl_trigger();
    pdl_wait();
    const int seq = blockIdx.x;
    const float* row = logits + static_cast<size_t>(seq) * vocab;
    float best = -INFINITY;
    int idx = 0x7fffffff;
    for (int i = threadIdx.x; i < vocab; i += blockDim.x) {
        const float v = row[i];
        if (v > best || (v == best && i < idx)) { best = v; idx = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > best || (ov == best && oi < idx)) { best = ov; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { bv[threadIdx.x >> 5] = best; bi[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w)
            if (bv[w] > best || (bv[w] == best && bi[w] < idx)) { best = bv[w]; idx = bi[w]; }
        const int step = step_ptr ? *step_ptr : step_host;
        const bool was_done = finished[seq] != 0;
        const int tok = was_done ? eos_id : idx;
        if (step < max_new) tokens[static_cast<size_t>(seq) * max_new + step] = was_done ? -1 : tok;
        if (!was_done && tok == eos_id) finished[seq] = 1;
        next_ids[seq] = tok;
        if (seq_lens) seq_lens[seq] += 1;
    }
}
// Temperature / top-k sampling step (HF `sample()` with TemperatureLogitsWarper + TopKLogitsWarper, the path
// eval/inference.py:64-72 takes with do_sample=True, temperature=0.2 and HF's default top_k=50), one block per
// sequence:  z = logits / T;  keep z >= (k-th largest z) (ties kept, like HF's `scores < kth` mask);  softmax;
// inverse-CDF draw in index order with u from a counter-based generator keyed by (seed, step, sequence) — so the
// step is replayable from a CUDA graph and reproducible, but not bit-compatible with torch.multinomial's stream.
__device__ __forceinline__ uint32_t float_order_key(float f) {          // monotone float → uint
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__global__ void __launch_bounds__(256) sample_step_kernel(const float* __restrict__ logits, int vocab, float inv_temp, int top_k,
                                                          unsigned long long seed, const unsigned long long* seed_ptr, uint8_t* finished,
                                                          int* tokens, int max_new, int step_host, const int* step_ptr, int* next_ids,
                                                          int* seq_lens, int eos_id) {
    __shared__ uint32_t hist[256];
    __shared__ float redf[8];
    __shared__ uint32_t sh_prefix, sh_remaining;
    __shared__ float sh_scan[256];
    pdl_trigger();
    pdl_wait();
    const int seq = blockIdx.x, tid = threadIdx.x;
    const float* row = logits + static_cast<size_t>(seq) * vocab;
    // ---- k-th largest key by 4 radix passes (most significant byte first)
    uint32_t prefix = 0, remaining = static_cast<uint32_t>(min(max(top_k, 1), vocab));
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        hist[tid] = 0;
        __syncthreads();
        const uint32_t mask_hi = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (int i = tid; i < vocab; i += 256) {
            const uint32_t key = float_order_key(row[i] * inv_temp);
            if ((key & mask_hi) == (prefix & mask_hi)) atomicAdd(&hist[(key >> shift) & 0xFF], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t rem = remaining;
            int b = 255;
            for (; b > 0; --b) {
                if (hist[b] >= rem) break;
                rem -= hist[b];
            }
            sh_prefix = prefix | (static_cast<uint32_t>(b) << shift);
            sh_remaining = rem;
        }
        __syncthreads();
        prefix = sh_prefix;
        remaining = sh_remaining;
    }
    const uint32_t kth_key = prefix;        // keep keys >= kth_key
    // ---- max and normaliser over the kept set
    float mx = -INFINITY;
    for (int i = tid; i < vocab; i += 256) mx = fmaxf(mx, row[i] * inv_temp);
    mx = warp_max(mx);
    if ((tid & 31) == 0) redf[tid >> 5] = mx;
    __syncthreads();
    mx = redf[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, redf[i]);
    __syncthreads();
    // each thread owns a contiguous index range so the inverse-CDF walk is in index order
    const int per = (vocab + 255) / 256, lo = tid * per, hi = min(vocab, lo + per);
    float mine = 0.f;
    for (int i = lo; i < hi; ++i) {
        const float z = row[i] * inv_temp;
        if (float_order_key(z) >= kth_key) mine += expf(z - mx);
    }
    sh_scan[tid] = mine;
    __syncthreads();
    if (tid == 0) {
        float total = 0.f;
        for (int i = 0; i < 256; ++i) total += sh_scan[i];
        const int step = step_ptr ? *step_ptr : step_host;
        if (seed_ptr) seed = *seed_ptr;           // device-resident seed: a captured decode graph is reusable across seeds
        const uint64_t r = splitmix64(seed + (static_cast<uint64_t>(step) * 0x100000001B3ULL + static_cast<uint64_t>(seq) + 1ULL) * 0x9E3779B97F4A7C15ULL);
        const float u = static_cast<float>(r >> 40) * (1.0f / 16777216.0f);      // [0,1)
        float target = u * total, acc = 0.f;
        int owner = 255;
        for (int i = 0; i < 256; ++i) {
            if (acc + sh_scan[i] > target) { owner = i; break; }
            acc += sh_scan[i];
        }
        // walk the owner's range
        int pick = -1;
        const int olo = owner * per, ohi = min(vocab, olo + per);
        for (int i = olo; i < ohi; ++i) {
            const float z = row[i] * inv_temp;
            if (float_order_key(z) >= kth_key) {
                pick = i;                                   // last kept index is the fallback for rounding at the tail
                acc += expf(z - mx);
                if (acc > target) break;
            }
        }
        if (pick < 0) {                                     // numerical corner: fall back to the arg-max
            float best = -INFINITY;
            for (int i = 0; i < vocab; ++i) if (row[i] > best) { best = row[i]; pick = i; }
        }
        const bool was_done = finished[seq] != 0;
        const int tok = was_done ? eos_id : pick;
        if (step < max_new) tokens[static_cast<size_t>(seq) * max_new + step] = was_done ? -1 : tok;
        if (!was_done && tok == eos_id) finished[seq] = 1;
        next_ids[seq] = tok;
        if (seq_lens) seq_lens[seq] += 1;
    }
}

__global__ void bump_step_kernel(int* step_ptr) {
    pdl_trigger();
    pdl_wait();
    *step_ptr += 1;
}

}  // namespace teo

using namespace teo;

static inline int grid_for(size_t n, int threads) {
    size_t b = (n + threads - 1) / threads;
    return static_cast<int>(b > 148 * 32 ? 148 * 32 : (b == 0 ? 1 : b));
}

extern "C" int teo_init_normal_hash_bf16(void* out, size_t n, uint64_t seed, float scale, float mean, void* stream) {
    TEO_CHECK_ARG(out || n == 0, "init: null output");
    if (n == 0) return TEO_OK;
    init_normal_hash_kernel<bf16><<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<bf16*>(out), n, seed, scale, mean);
    TEO_LAUNCH_CHECK("init_normal_hash_kernel");
    return TEO_OK;
}
extern "C" int teo_init_normal_hash_f32(void* out, size_t n, uint64_t seed, float scale, float mean, void* stream) {
    TEO_CHECK_ARG(out || n == 0, "init: null output");
    if (n == 0) return TEO_OK;
    init_normal_hash_kernel<float><<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float*>(out), n, seed, scale, mean);
    TEO_LAUNCH_CHECK("init_normal_hash_kernel");
    return TEO_OK;
}
extern "C" int teo_init_u8_hash(void* out, size_t n, uint64_t seed, void* stream) {
    TEO_CHECK_ARG(out || n == 0, "init: null output");
    if (n == 0) return TEO_OK;
    init_u8_hash_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<uint8_t*>(out), n, seed);
    TEO_LAUNCH_CHECK("init_u8_hash_kernel");
    return TEO_OK;
}

static int patchify_common(bool u8, const void* in, void* patches, int n_frames, int image, int patch, int kpad, void* stream, int planes = 1) {
    TEO_CHECK_ARG(in && patches, "patchify: null pointer");
    TEO_CHECK_ARG(n_frames > 0 && image > 0 && patch > 0 && image % patch == 0, "patchify: bad geometry image=%d patch=%d", image, patch);
    TEO_CHECK_ARG(kpad >= 3 * patch * patch && kpad % 8 == 0, "patchify: kpad=%d must be >= %d and a multiple of 8", kpad, 3 * patch * patch);
    const int g = image / patch;
    const size_t smem = static_cast<size_t>(3) * patch * image * sizeof(float);
    TEO_CHECK_ARG(smem <= 48 * 1024, "patchify: image row tile (%zu B) exceeds 48 KiB", smem);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (u8 && planes == 1 && (patch * image * 3) % 16 == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(patches) & 15) == 0) {
        const size_t vsmem = ((static_cast<size_t>(patch) * image * 3 + 15) & ~size_t(15)) + 768 * sizeof(float);
        patchify_u8_vec_kernel<<<n_frames * g, 256, vsmem, s>>>(static_cast<const uint8_t*>(in), static_cast<bf16*>(patches), image, patch, kpad);
        TEO_LAUNCH_CHECK("patchify_u8_vec_kernel");
        return TEO_OK;
    }
    if (planes == 3) {
        if (u8) patchify_kernel<true, 3><<<n_frames * g, 256, smem, s>>>(in, static_cast<bf16*>(patches), image, patch, kpad);
        else patchify_kernel<false, 3><<<n_frames * g, 256, smem, s>>>(in, static_cast<bf16*>(patches), image, patch, kpad);
    } else {
        if (u8) patchify_kernel<true, 1><<<n_frames * g, 256, smem, s>>>(in, static_cast<bf16*>(patches), image, patch, kpad);
        else patchify_kernel<false, 1><<<n_frames * g, 256, smem, s>>>(in, static_cast<bf16*>(patches), image, patch, kpad);
    }
    TEO_LAUNCH_CHECK("patchify_kernel");
    return TEO_OK;
}
extern "C" int teo_patchify_u8_nhwc(const void* frames_u8, void* patches, int n_frames, int image, int patch, int kpad, void* stream) {
    return patchify_common(true, frames_u8, patches, n_frames, image, patch, kpad, stream);
}
extern "C" int teo_patchify_f32_nchw(const void* pixel_values, void* patches, int n_frames, int image, int patch, int kpad, void* stream) {
    return patchify_common(false, pixel_values, patches, n_frames, image, patch, kpad, stream);
}
namespace teo {
// exact mode: patches bf16 [n*g*g, 3*kpad] (hi | mid | lo planes of the normalised fp32 pixel values)
int x_patchify(bool u8, const void* in, void* patches3, int n_frames, int image, int patch, int kpad, cudaStream_t stream) {
    return patchify_common(u8, in, patches3, n_frames, image, patch, kpad, stream, 3);
}
}  // namespace teo

extern "C" int teo_vit_assemble_preln(const void* patch_out, const void* cls, const void* pos, const void* ln_w, const void* ln_b,
                                      void* hidden, int n_frames, int n_patches, int d, float eps, void* stream) {
    TEO_CHECK_ARG(patch_out && cls && pos && ln_w && ln_b && hidden, "vit_assemble_preln: null pointer");
    TEO_CHECK_ARG(n_frames > 0 && n_patches > 0, "vit_assemble_preln: bad sizes");
    return launch_norm<1>(static_cast<const bf16*>(patch_out), static_cast<const bf16*>(ln_w), static_cast<const bf16*>(ln_b),
                          static_cast<bf16*>(hidden), n_frames * (n_patches + 1), d, eps, static_cast<const bf16*>(cls),
                          static_cast<const bf16*>(pos), n_patches, static_cast<cudaStream_t>(stream));
}
extern "C" int teo_row_stats(const void* x, void* stats, int rows, int d, int slots, void* stream) {
    TEO_CHECK_ARG(x && stats && rows > 0 && d > 0 && d % 8 == 0 && slots > 0, "row_stats: bad arguments (rows=%d d=%d slots=%d)", rows, d, slots);
    row_stats_kernel<<<(rows * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(x), static_cast<float*>(stats),
                                                                                             rows, d, slots);
    TEO_LAUNCH_CHECK("row_stats_kernel");
    return TEO_OK;
}
extern "C" int teo_layernorm(const void* x, const void* w, const void* b, void* y, int rows, int d, float eps, void* stream) {
    TEO_CHECK_ARG(x && w && b && y, "layernorm: null pointer");
    return launch_norm<0>(static_cast<const bf16*>(x), static_cast<const bf16*>(w), static_cast<const bf16*>(b), static_cast<bf16*>(y),
                          rows, d, eps, nullptr, nullptr, 0, static_cast<cudaStream_t>(stream));
}
extern "C" int teo_rmsnorm(const void* x, const void* w, void* y, int rows, int d, float eps, void* stream) {
    TEO_CHECK_ARG(x && w && y, "rmsnorm: null pointer");
    return launch_norm<2>(static_cast<const bf16*>(x), static_cast<const bf16*>(w), nullptr, static_cast<bf16*>(y), rows, d, eps,
                          nullptr, nullptr, 0, static_cast<cudaStream_t>(stream));
}
extern "C" int teo_vit_drop_cls(const void* hidden, void* feats, int n_frames, int n_patches, int d, void* stream) {
    TEO_CHECK_ARG(hidden && feats, "vit_drop_cls: null pointer");
    TEO_CHECK_ARG(n_frames > 0 && n_patches > 0 && d > 0 && d % 8 == 0, "vit_drop_cls: bad sizes");
    const size_t total = static_cast<size_t>(n_frames) * n_patches * (d / 8);
    drop_cls_kernel<<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(hidden), static_cast<uint4*>(feats), n_patches, d / 8, total);
    TEO_LAUNCH_CHECK("drop_cls_kernel");
    return TEO_OK;
}
namespace teo {
int launch_swiglu(const void* gate_up, void* out, int rows, int inter, int interleaved, cudaStream_t stream) {
    TEO_CHECK_ARG(gate_up && out, "swiglu: null pointer");
    TEO_CHECK_ARG(rows > 0 && inter > 0 && inter % 8 == 0, "swiglu: rows=%d inter=%d", rows, inter);
    TEO_CHECK_ARG(!interleaved || inter % 32 == 0, "swiglu: the interleaved layout needs inter %% 32 == 0 (inter=%d)", inter);
    const size_t total = static_cast<size_t>(rows) * (inter / 8);
    swiglu_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const bf16*>(gate_up), static_cast<bf16*>(out), rows, inter,
                                                           interleaved ? 1 : 0);
    TEO_LAUNCH_CHECK("swiglu_kernel");
    return TEO_OK;
}
}  // namespace teo
extern "C" int teo_swiglu(const void* gate_up, void* out, int rows, int inter, void* stream) {
    return teo::launch_swiglu(gate_up, out, rows, inter, 0, static_cast<cudaStream_t>(stream));
}
extern "C" int teo_splice_embed(const void* embed_tokens, const void* image_feats, const void* src, void* out, int tokens, int d,
                                void* stream) {
    TEO_CHECK_ARG(embed_tokens && src && out, "splice_embed: null pointer");
    TEO_CHECK_ARG(tokens > 0 && d > 0 && d % 8 == 0, "splice_embed: tokens=%d d=%d", tokens, d);
    TEO_CUDA(launch_k(splice_embed_kernel, dim3(tokens), dim3(128), 0, static_cast<cudaStream_t>(stream), static_cast<const bf16*>(embed_tokens),
                      static_cast<const bf16*>(image_feats), static_cast<const int*>(src), static_cast<bf16*>(out), tokens, d / 8));
    TEO_LAUNCH_CHECK("splice_embed_kernel");
    return TEO_OK;
}

namespace teo {
int launch_rope_kv_write(void* qkv, const int* positions, const int* seq_ids, void* kv_pages, const int* block_table, int max_pages,
                         int tokens, int n_heads, int head_dim, int page_size, const float* rope_cos, const float* rope_sin,
                         cudaStream_t stream) {
    TEO_CHECK_ARG(qkv && positions && kv_pages && block_table && rope_cos && rope_sin, "rope_kv_write: null pointer");
    TEO_CHECK_ARG(tokens > 0 && n_heads > 0 && head_dim > 0 && head_dim % 4 == 0 && page_size > 0, "rope_kv_write: bad sizes");
    if (head_dim % 16 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(kv_pages) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(rope_cos) & 15) == 0 && (reinterpret_cast<uintptr_t>(rope_sin) & 15) == 0 && (n_heads * head_dim) % 8 == 0) {
        const long long items = static_cast<long long>(tokens) * n_heads;
        const long long nthreads = items * (head_dim / 16);
        rope_kv_write_vec_kernel<<<static_cast<unsigned>((nthreads + 255) / 256), 256, 0, stream>>>(
            static_cast<bf16*>(qkv), positions, seq_ids, static_cast<bf16*>(kv_pages), block_table, max_pages, items, n_heads, head_dim,
            page_size, rope_cos, rope_sin);
        TEO_LAUNCH_CHECK("rope_kv_write_vec_kernel");
        return TEO_OK;
    }
    const long long warps = static_cast<long long>(tokens) * n_heads;
    const int threads = 256;
    const long long blocks = (warps * 32 + threads - 1) / threads;
    rope_kv_write_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(static_cast<bf16*>(qkv), positions, seq_ids,
                                                                               static_cast<bf16*>(kv_pages), block_table, max_pages, tokens,
                                                                               n_heads, head_dim, page_size, rope_cos, rope_sin);
    TEO_LAUNCH_CHECK("rope_kv_write_kernel");
    return TEO_OK;
}
int launch_reduce_residual_rmsnorm(const PartialInfo& pi, bf16* x, const bf16* w, bf16* y, int rows, int d, float eps, cudaStream_t stream) {
    TEO_CHECK_ARG(pi.P && x && w && y && rows > 0, "reduce_residual_rmsnorm: bad arguments");
    TEO_CHECK_ARG(d > 0 && d % (RN_CLUSTER * 4) == 0 && d <= RN_CLUSTER * RN_THREADS * 4 * RN_MAXV,
                  "reduce_residual_rmsnorm: d=%d must be a multiple of %d and <= %d", d, RN_CLUSTER * 4, RN_CLUSTER * RN_THREADS * 4 * RN_MAXV);
    TEO_CUDA(launch_k(reduce_residual_rmsnorm_kernel, dim3(rows * RN_CLUSTER), dim3(RN_THREADS), 0, stream, pi, x, w, y, d, eps));
    TEO_LAUNCH_CHECK("reduce_residual_rmsnorm_kernel");
    return TEO_OK;
}
int launch_reduce_swiglu(const PartialInfo& pi, bf16* act, int rows, int inter, int interleaved, cudaStream_t stream) {
    TEO_CHECK_ARG(pi.P && act && rows > 0 && inter > 0, "reduce_swiglu: bad arguments");
    TEO_CHECK_ARG(inter % 4 == 0 && (!interleaved || inter % 32 == 0), "reduce_swiglu: inter %% 4 != 0 (or %% 32 for the interleaved layout)");
    const long long total = static_cast<long long>(rows) * (inter / 4);
    TEO_CUDA(launch_k(reduce_swiglu_kernel, dim3(static_cast<unsigned>((total + 127) / 128)), dim3(128), 0, stream, pi, act, rows, inter, interleaved ? 1 : 0));
    TEO_LAUNCH_CHECK("reduce_swiglu_kernel");
    return TEO_OK;
}
int launch_reduce_rope_kv_write(const PartialInfo& pi, void* qkv, const int* positions, void* kv_pages,
                                const int* block_table, int max_pages, int n_seqs, int n_heads, int head_dim, int page_size,
                                const float* rope_cos, const float* rope_sin, cudaStream_t stream) {
    TEO_CHECK_ARG(pi.P && qkv && positions && kv_pages && block_table && rope_cos && rope_sin, "reduce_rope_kv_write: null pointer");
    const long long warps = static_cast<long long>(n_seqs) * n_heads;
    TEO_CUDA(launch_k(reduce_rope_kv_write_kernel, dim3(static_cast<unsigned>((warps * 32 + 255) / 256)), dim3(256), 0, stream, pi,
                      static_cast<bf16*>(qkv), positions, static_cast<bf16*>(kv_pages), block_table, max_pages, n_seqs, n_heads, head_dim, page_size,
                      rope_cos, rope_sin));
    TEO_LAUNCH_CHECK("reduce_rope_kv_write_kernel");
    return TEO_OK;
}
int launch_sample_step(const float* logits, int vocab, float temperature, int top_k, unsigned long long seed, uint8_t* finished, int* tokens,
                       int max_new, int step_host, int* step_ptr, int* next_ids, int* seq_lens, int n_seqs, int eos_id, cudaStream_t stream,
                       const unsigned long long* seed_ptr) {
    TEO_CHECK_ARG(logits && finished && tokens && next_ids, "sample_step: null pointer");
    TEO_CHECK_ARG(n_seqs > 0 && vocab > 0 && max_new > 0 && temperature > 0.f, "sample_step: bad sizes / temperature");
    if (top_k <= 0 || top_k > vocab) top_k = vocab;
    TEO_CUDA(launch_k(sample_step_kernel, dim3(n_seqs), dim3(256), 0, stream, logits, vocab, 1.0f / temperature, top_k, seed, seed_ptr, finished, tokens, max_new,
                      step_host, static_cast<const int*>(step_ptr), next_ids, seq_lens, eos_id));
    TEO_LAUNCH_CHECK("sample_step_kernel");
    if (step_ptr) {
        TEO_CUDA(launch_k(bump_step_kernel, dim3(1), dim3(1), 0, stream, step_ptr));
        TEO_LAUNCH_CHECK("bump_step_kernel");
    }
    return TEO_OK;
}
int launch_argmax_step(const float* logits, int vocab, uint8_t* finished, int* tokens, int max_new, int step_host, int* step_ptr,
                       int* next_ids, int* seq_lens, int n_seqs, int eos_id, cudaStream_t stream) {
    TEO_CHECK_ARG(logits && finished && tokens && next_ids, "argmax_step: null pointer");
    TEO_CHECK_ARG(n_seqs > 0 && vocab > 0 && max_new > 0, "argmax_step: bad sizes");
    TEO_CUDA(launch_k(argmax_step_kernel, dim3(n_seqs), dim3(1024), 0, stream, logits, vocab, finished, tokens, max_new, step_host, step_ptr, next_ids,
                      seq_lens, eos_id));
    TEO_LAUNCH_CHECK("argmax_step_kernel");
    if (step_ptr) {
        TEO_CUDA(launch_k(bump_step_kernel, dim3(1), dim3(1), 0, stream, step_ptr));
        TEO_LAUNCH_CHECK("bump_step_kernel");
    }
    return TEO_OK;
}
}  // namespace teo

extern "C" int teo_rope_kv_write(void* qkv, const void* positions, const void* seq_ids, void* kv_pages, const void* block_table,
                                 int max_pages, int tokens, int n_heads, int head_dim, int page_size, const void* rope_cos,
                                 const void* rope_sin, void* stream) {
    return launch_rope_kv_write(qkv, static_cast<const int*>(positions), static_cast<const int*>(seq_ids), kv_pages,
                                static_cast<const int*>(block_table), max_pages, tokens, n_heads, head_dim, page_size,
                                static_cast<const float*>(rope_cos), static_cast<const float*>(rope_sin), static_cast<cudaStream_t>(stream));
}
extern "C" int teo_sample_step(const void* logits, int vocab, float temperature, int top_k, uint64_t seed, void* finished, void* tokens,
                               int max_new, int step, void* next_ids, int n_seqs, int eos_id, void* stream) {
    return launch_sample_step(static_cast<const float*>(logits), vocab, temperature, top_k, seed, static_cast<uint8_t*>(finished),
                              static_cast<int*>(tokens), max_new, step, nullptr, static_cast<int*>(next_ids), nullptr, n_seqs, eos_id,
                              static_cast<cudaStream_t>(stream));
}
extern "C" int teo_argmax_step(const void* logits, int vocab, void* finished, void* tokens, int max_new, int step, void* next_ids,
                               int n_seqs, int eos_id, void* stream) {
    return launch_argmax_step(static_cast<const float*>(logits), vocab, static_cast<uint8_t*>(finished), static_cast<int*>(tokens), max_new,
                              step, nullptr, static_cast<int*>(next_ids), nullptr, n_seqs, eos_id, static_cast<cudaStream_t>(stream));
}
