// Exact ("parity") mode of the hot path: every activation, the residual stream and the KV pages are fp32 in HBM, and
// the contractions still run on the 5th-gen tensor cores — each fp32 activation x is split into three bf16 terms
// hi = bf16(x), mid = bf16(x − hi), lo = bf16(x − hi − mid) (hi + mid + lo == x exactly: 3 × 8 significand bits), stored
// side by side as [rows, 3K]; the tcgen05 GEMM (gemm.cu, GemmArgs::k_wrap) contracts all three planes against the same
// bf16 weights with fp32 accumulation in TMEM, i.e. it computes the fp32 product Σ_k x_k·w_k up to summation order.
// Attention, norms, RoPE, activations run in fp32 on the CUDA cores with the HF-4.31 op order.
//
// Purpose: the north-star parity clause ("bit-exact token ids under greedy decode with fp32 accumulation").  With bf16
// activation storage the 32-layer random-init network sits 3–4 % from the fp32 oracle whatever the kernels do (bf16
// GEMM inputs alone: 1.9 %); in this mode the full-size configs[0] logits agree with the fp32 oracle to
// ~1e-5 and all greedy ids match.  Throughput is irrelevant here (≈ 3× the tensor work, CUDA-core attention).
#include <math.h>

#include "common.h"
#include "ptx.cuh"

namespace teo {

__device__ __forceinline__ void split3(float x, bf16& hi, bf16& mid, bf16& lo) {
    hi = __float2bfloat16_rn(x);
    float r = x - __bfloat162float(hi);          // exact: x − bf16(x) needs ≤ 16 significand bits
    mid = __float2bfloat16_rn(r);
    r -= __bfloat162float(mid);                  // exact, ≤ 8 bits left
    lo = __float2bfloat16_rn(r);
}
__device__ __forceinline__ void store_split4(bf16* row3, int K, int c, const float (&y)[4]) {
    bf16 hi[4], mid[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split3(y[j], hi[j], mid[j], lo[j]);
    *reinterpret_cast<uint2*>(row3 + c) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(row3 + K + c) = *reinterpret_cast<const uint2*>(mid);
    *reinterpret_cast<uint2*>(row3 + 2 * K + c) = *reinterpret_cast<const uint2*>(lo);
}

__device__ __forceinline__ float block_sum(float v, float* red) {      // all threads get the total; red: ≥ 32 floats
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < static_cast<int>((blockDim.x + 31) >> 5); ++i) t += red[i];
    return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = -INFINITY;
    for (int i = 0; i < static_cast<int>((blockDim.x + 31) >> 5); ++i) t = fmaxf(t, red[i]);
    return t;
}

// ------------------------------------------------------------------------------ element-wise → split planes
// MODE 0: y = x                      in [rows, K]   (n_patches > 0: input row = frame·(np+1) + tok + 1, i.e. CLS dropped)
// MODE 1: y = x·σ(1.702x)            quick_gelu
// MODE 2: y = ½x(1 + erf(x/√2))      nn.GELU()
// MODE 3: y = silu(g)·u              in [rows, 2K] = gate/up (halves, or interleaved in blocks of 32)
// out3: bf16 [rows, 3K] planes, or (out_f32 != nullptr) a plain fp32 copy [rows, K] instead.
template <int MODE>
__global__ void x_split_kernel(const float* __restrict__ in, bf16* __restrict__ out3, float* __restrict__ out_f32, int rows, int K,
                               int n_patches, int interleaved) {
    const int k4 = K / 4;
    const long long total = static_cast<long long>(rows) * k4;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / k4;
        const int c = static_cast<int>(i % k4) * 4;
        float y[4];
        if constexpr (MODE == 3) {
            const int gc = interleaved ? (c / 32) * 64 + (c % 32) : c;
            const int up = interleaved ? 32 : K;
            const float4 g = *reinterpret_cast<const float4*>(in + r * 2 * K + gc);
            const float4 u = *reinterpret_cast<const float4*>(in + r * 2 * K + gc + up);
            const float gf[4] = {g.x, g.y, g.z, g.w}, uf[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) y[j] = (gf[j] / (1.0f + expf(-gf[j]))) * uf[j];
        } else {
            long long rin = r;
            if (n_patches > 0) rin = (r / n_patches) * (n_patches + 1) + (r % n_patches) + 1;
            const float4 x = *reinterpret_cast<const float4*>(in + rin * K + c);
            const float xf[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if constexpr (MODE == 1) y[j] = xf[j] / (1.0f + expf(-1.702f * xf[j]));
                else if constexpr (MODE == 2) y[j] = 0.5f * xf[j] * (1.0f + erff(xf[j] * 0.70710678118654752f));
                else y[j] = xf[j];
            }
        }
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + r * K + c) = make_float4(y[0], y[1], y[2], y[3]);
        else store_split4(out3 + r * 3 * K, K, c, y);
    }
}

// ------------------------------------------------------------------------------ row norms (one 128-thread block per row)
// MODE 0: LayerNorm(x)    MODE 1: ViT embeddings ([CLS; patch_out] + pos) then LayerNorm    MODE 2: RMSNorm (HF order:
// x·rsqrt(mean(x²)+eps) then ·w).  Statistics in fp32, two-pass variance.  Writes fp32 (out_f32) and / or split planes.
template <int MODE>
__global__ void __launch_bounds__(128)
x_norm_kernel(const float* __restrict__ in, const bf16* __restrict__ w, const bf16* __restrict__ b, float* __restrict__ out_f32,
              bf16* __restrict__ out3, int d, float eps, const bf16* __restrict__ cls, const bf16* __restrict__ pos, int n_patches) {
    __shared__ float red[32];
    const long long row = blockIdx.x;
    const float* src = in + row * d;
    int tok = 0;
    if constexpr (MODE == 1) {
        tok = static_cast<int>(row % (n_patches + 1));
        const long long frame = row / (n_patches + 1);
        src = in + (frame * n_patches + tok - 1) * d;          // only dereferenced when tok > 0
    }
    auto value = [&](int c) -> float {
        if constexpr (MODE == 1) {
            const float base = tok == 0 ? __bfloat162float(cls[c]) : src[c];
            return base + __bfloat162float(pos[static_cast<long long>(tok) * d + c]);
        } else {
            return src[c];
        }
    };
    float s = 0.f;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const float x = value(c);
        s += (MODE == 2) ? x * x : x;
    }
    s = block_sum(s, red);
    float mean = 0.f, rstd;
    if constexpr (MODE == 2) {
        rstd = rsqrtf(s / static_cast<float>(d) + eps);
    } else {
        mean = s / static_cast<float>(d);
        float q = 0.f;
        for (int c = threadIdx.x; c < d; c += blockDim.x) {
            const float dl = value(c) - mean;
            q += dl * dl;
        }
        q = block_sum(q, red);
        rstd = rsqrtf(q / static_cast<float>(d) + eps);
    }
    for (int c = threadIdx.x * 4; c < d; c += blockDim.x * 4) {
        float y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float x = value(c + j), wv = __bfloat162float(w[c + j]);
            if constexpr (MODE == 2) y[j] = (x * rstd) * wv;
            else y[j] = (x - mean) * rstd * wv + __bfloat162float(b[c + j]);
        }
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * d + c) = make_float4(y[0], y[1], y[2], y[3]);
        if (out3) store_split4(out3 + row * 3 * d, d, c, y);
    }
}

// ------------------------------------------------------------------------------ gathers
// out[row] = src[row] >= 0 ? float(embed[src[row]]) : feats[-(src[row]+1)]     (teo_splice_embed in fp32)
__global__ void x_splice_embed_kernel(const bf16* __restrict__ embed, const float* __restrict__ feats, const int* __restrict__ src,
                                      float* __restrict__ out, int d) {
    const long long row = blockIdx.x;
    const int s = src[row];
    float* to = out + row * d;
    if (s >= 0) {
        const bf16* from = embed + static_cast<long long>(s) * d;
        for (int c = threadIdx.x; c < d; c += blockDim.x) to[c] = __bfloat162float(from[c]);
    } else {
        const float* from = feats + static_cast<long long>(-(s + 1)) * d;
        for (int c = threadIdx.x; c < d; c += blockDim.x) to[c] = from[c];
    }
}
__global__ void x_gather_rows_kernel(const float* __restrict__ in, const int* __restrict__ rows, float* __restrict__ out, int d) {
    const float* from = in + static_cast<long long>(rows[blockIdx.x]) * d;
    float* to = out + static_cast<long long>(blockIdx.x) * d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) to[c] = from[c];
}

// ------------------------------------------------------------------------------ RoPE + fp32 KV pages
// One warp per (token, head); rotate-half on q and k in place in the fp32 qkv buffer, k and v rows into the fp32 page pool
// [page][2][head][slot][dim].  cos/sin tables as in the bf16 path (built on the host like HF's cached tables).
__global__ void x_rope_kv_write_kernel(float* __restrict__ qkv, const int* __restrict__ positions, const int* __restrict__ seq_ids,
                                       float* __restrict__ kv_pages, const int* __restrict__ block_table, int max_pages, int tokens,
                                       int n_heads, int head_dim, int page_size, const float* __restrict__ rope_cos,
                                       const float* __restrict__ rope_sin) {
    const long long gw = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= static_cast<long long>(tokens) * n_heads) return;
    const int tok = static_cast<int>(gw / n_heads), head = static_cast<int>(gw % n_heads);
    const int hidden = n_heads * head_dim, half = head_dim / 2;
    const int pos = positions[tok];
    const int seq = seq_ids ? seq_ids[tok] : tok;
    const int page = block_table[static_cast<size_t>(seq) * max_pages + pos / page_size];
    const int slot = pos % page_size;
    float* q = qkv + static_cast<size_t>(tok) * 3 * hidden + head * head_dim;
    float* k = q + hidden;
    const float* v = k + hidden;
    float* kdst = kv_pages + (((static_cast<size_t>(page) * 2 + 0) * n_heads + head) * page_size + slot) * head_dim;
    float* vdst = kv_pages + (((static_cast<size_t>(page) * 2 + 1) * n_heads + head) * page_size + slot) * head_dim;
    const float* cs = rope_cos + static_cast<size_t>(pos) * half;
    const float* sn = rope_sin + static_cast<size_t>(pos) * half;
    for (int i = lane; i < half; i += 32) {
        const float c = cs[i], s = sn[i];
        const float qa = q[i], qb = q[i + half], ka = k[i], kb = k[i + half];
        q[i] = qa * c - qb * s;
        q[i + half] = qb * c + qa * s;
        const float k0 = ka * c - kb * s, k1 = kb * c + ka * s;
        k[i] = k0;
        k[i + half] = k1;
        kdst[i] = k0;
        kdst[i + half] = k1;
    }
    for (int i = lane; i < head_dim; i += 32) vdst[i] = v[i];
}

// ------------------------------------------------------------------------------ attention in fp32
// One 128-thread block per (query row, head): scores for all visible keys in shared memory, softmax in fp32, then P·V.
//   PAGED = false  keys / values are rows [seq_start, seq_start + len) of the same fp32 qkv buffer (ViT: no mask)
//   PAGED = true   keys / values come from the fp32 page pool of the query's sequence, keys 0 … positions[row]
//                  (LLaMA prefill after x_rope_kv_write, and decode with positions = cached length)
template <bool PAGED>
__global__ void __launch_bounds__(128)
x_attention_kernel(const float* __restrict__ qkv, int ld, float* __restrict__ out, int ldo, const int* __restrict__ cu_seqlens,
                   const int* __restrict__ positions, const int* __restrict__ seq_ids, const float* __restrict__ kv_pages,
                   const int* __restrict__ block_table, int max_pages, int n_heads, int head_dim, int page_size, float scale) {
    extern __shared__ float sm[];
    float* q = sm;                        // [head_dim]
    float* red = sm + head_dim;           // [32]
    float* acc = red + 32;                // [128]
    float* sc = acc + 128;                // [n_keys]
    const int row = blockIdx.x, head = blockIdx.y;
    const int hidden = n_heads * head_dim;
    int seq, n_keys, seq_start = 0;
    if constexpr (PAGED) {
        seq = seq_ids ? seq_ids[row] : row;
        n_keys = positions[row] + 1;
    } else {
        seq = seq_ids[row];
        seq_start = cu_seqlens[seq];
        n_keys = cu_seqlens[seq + 1] - seq_start;
    }
    const float* qrow = qkv + static_cast<size_t>(row) * ld + head * head_dim;
    for (int i = threadIdx.x; i < head_dim; i += blockDim.x) q[i] = qrow[i];
    __syncthreads();
    auto kv_row = [&](int j, int which) -> const float* {
        if constexpr (PAGED) {
            const int page = block_table[static_cast<size_t>(seq) * max_pages + j / page_size];
            return kv_pages + (((static_cast<size_t>(page) * 2 + which) * n_heads + head) * page_size + j % page_size) * head_dim;
        } else {
            return qkv + static_cast<size_t>(seq_start + j) * ld + (which + 1) * hidden + head * head_dim;
        }
    };
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < n_keys; j += blockDim.x) {
        const float4* kr = reinterpret_cast<const float4*>(kv_row(j, 0));
        float dot = 0.f;
        for (int i = 0; i < head_dim / 4; ++i) {
            const float4 kk = kr[i];
            dot += q[4 * i] * kk.x + q[4 * i + 1] * kk.y + q[4 * i + 2] * kk.z + q[4 * i + 3] * kk.w;
        }
        dot *= scale;
        sc[j] = dot;
        mx = fmaxf(mx, dot);
    }
    mx = block_max(mx, red);
    float sum = 0.f;
    for (int j = threadIdx.x; j < n_keys; j += blockDim.x) {
        const float p = expf(sc[j] - mx);
        sc[j] = p;
        sum += p;
    }
    sum = block_sum(sum, red);            // (its barriers also publish sc[])
    // P·V: thread t owns dim t % head_dim and the key slice t / head_dim (128 / head_dim slices)
    const int slices = blockDim.x / head_dim;
    const int dim = threadIdx.x % head_dim, slice = threadIdx.x / head_dim;
    float a = 0.f;
    if (slice < slices)
        for (int j = slice; j < n_keys; j += slices) a += sc[j] * kv_row(j, 1)[dim];
    acc[threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.x < head_dim) {
        float t = 0.f;
        for (int s2 = 0; s2 < slices; ++s2) t += acc[s2 * head_dim + threadIdx.x];
        out[static_cast<size_t>(row) * ldo + head * head_dim + threadIdx.x] = t / sum;
    }
}

__global__ void fill_seq_ids_kernel(int* seq_ids, int rows, int len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) seq_ids[i] = i / len;
}

static inline int grid1d(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    return static_cast<int>(b > 148 * 64 ? 148 * 64 : (b < 1 ? 1 : b));
}

// ---- launch helpers used by model.cu -------------------------------------------------------------------------------------
int x_split(const float* in, bf16* out3, float* out_f32, int rows, int K, int mode, int n_patches, int interleaved, cudaStream_t s) {
    TEO_CHECK_ARG(in && (out3 || out_f32) && rows > 0 && K > 0 && K % 4 == 0, "x_split: bad arguments (rows=%d K=%d)", rows, K);
    TEO_CHECK_ARG(mode != 3 || !interleaved || K % 32 == 0, "x_split: interleaved SwiGLU needs K %% 32 == 0");
    const int g = grid1d(static_cast<long long>(rows) * (K / 4), 256);
    switch (mode) {
        case 0: x_split_kernel<0><<<g, 256, 0, s>>>(in, out3, out_f32, rows, K, n_patches, interleaved); break;
        case 1: x_split_kernel<1><<<g, 256, 0, s>>>(in, out3, out_f32, rows, K, n_patches, interleaved); break;
        case 2: x_split_kernel<2><<<g, 256, 0, s>>>(in, out3, out_f32, rows, K, n_patches, interleaved); break;
        case 3: x_split_kernel<3><<<g, 256, 0, s>>>(in, out3, out_f32, rows, K, n_patches, interleaved); break;
        default: set_error("x_split: unknown mode %d", mode); return TEO_ERR_BAD_ARG;
    }
    TEO_LAUNCH_CHECK("x_split_kernel");
    return TEO_OK;
}

int x_norm(int mode, const float* in, const bf16* w, const bf16* b, float* out_f32, bf16* out3, int rows, int d, float eps,
           const bf16* cls, const bf16* pos, int n_patches, cudaStream_t s) {
    TEO_CHECK_ARG(in && w && (out_f32 || out3) && rows > 0 && d > 0 && d % 4 == 0, "x_norm: bad arguments (rows=%d d=%d)", rows, d);
    TEO_CHECK_ARG(mode == 2 || b != nullptr, "x_norm: LayerNorm needs a bias");
    switch (mode) {
        case 0: x_norm_kernel<0><<<rows, 128, 0, s>>>(in, w, b, out_f32, out3, d, eps, cls, pos, n_patches); break;
        case 1: x_norm_kernel<1><<<rows, 128, 0, s>>>(in, w, b, out_f32, out3, d, eps, cls, pos, n_patches); break;
        case 2: x_norm_kernel<2><<<rows, 128, 0, s>>>(in, w, b, out_f32, out3, d, eps, cls, pos, n_patches); break;
        default: set_error("x_norm: unknown mode %d", mode); return TEO_ERR_BAD_ARG;
    }
    TEO_LAUNCH_CHECK("x_norm_kernel");
    return TEO_OK;
}

int x_splice_embed(const bf16* embed, const float* feats, const int* src, float* out, int tokens, int d, cudaStream_t s) {
    TEO_CHECK_ARG(embed && src && out && tokens > 0 && d > 0, "x_splice_embed: bad arguments");
    x_splice_embed_kernel<<<tokens, 256, 0, s>>>(embed, feats, src, out, d);
    TEO_LAUNCH_CHECK("x_splice_embed_kernel");
    return TEO_OK;
}
int x_gather_rows(const float* in, const int* rows, float* out, int n, int d, cudaStream_t s) {
    x_gather_rows_kernel<<<n, 256, 0, s>>>(in, rows, out, d);
    TEO_LAUNCH_CHECK("x_gather_rows_kernel");
    return TEO_OK;
}
int x_rope_kv_write(float* qkv, const int* positions, const int* seq_ids, float* kv_pages, const int* block_table, int max_pages,
                    int tokens, int n_heads, int head_dim, int page_size, const float* rope_cos, const float* rope_sin, cudaStream_t s) {
    TEO_CHECK_ARG(head_dim % 2 == 0 && page_size > 0, "x_rope_kv_write: bad geometry");
    const long long threads = static_cast<long long>(tokens) * n_heads * 32;
    x_rope_kv_write_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, s>>>(qkv, positions, seq_ids, kv_pages, block_table, max_pages,
                                                                                         tokens, n_heads, head_dim, page_size, rope_cos, rope_sin);
    TEO_LAUNCH_CHECK("x_rope_kv_write_kernel");
    return TEO_OK;
}
// max_keys bounds the score buffer (shared memory): ≤ 11 k keys
int x_attention(bool paged, const float* qkv, int ld, float* out, int ldo, const int* cu_seqlens, const int* positions, const int* seq_ids,
                const float* kv_pages, const int* block_table, int max_pages, int rows, int n_heads, int head_dim, int page_size,
                int max_keys, float scale, cudaStream_t s) {
    TEO_CHECK_ARG(head_dim % 4 == 0 && head_dim <= 128 && 128 % head_dim == 0, "x_attention: head_dim %d must divide 128", head_dim);
    const size_t smem = (static_cast<size_t>(head_dim) + 32 + 128 + max_keys) * sizeof(float);
    TEO_CHECK_ARG(smem <= 48 * 1024, "x_attention: %d keys exceed the 48 KiB score buffer", max_keys);
    dim3 grid(rows, n_heads);
    if (paged)
        x_attention_kernel<true><<<grid, 128, smem, s>>>(qkv, ld, out, ldo, cu_seqlens, positions, seq_ids, kv_pages, block_table, max_pages,
                                                         n_heads, head_dim, page_size, scale);
    else
        x_attention_kernel<false><<<grid, 128, smem, s>>>(qkv, ld, out, ldo, cu_seqlens, positions, seq_ids, kv_pages, block_table, max_pages,
                                                          n_heads, head_dim, page_size, scale);
    TEO_LAUNCH_CHECK("x_attention_kernel");
    return TEO_OK;
}
int x_fill_seq_ids(int* seq_ids, int rows, int len, cudaStream_t s) {
    fill_seq_ids_kernel<<<(rows + 255) / 256, 256, 0, s>>>(seq_ids, rows, len);
    TEO_LAUNCH_CHECK("fill_seq_ids_kernel");
    return TEO_OK;
}

}  // namespace teo
