// Image preprocessing of the reference's processor for frames that are NOT already image_size²
// (processing_image.py:15-25): ToTensor (u8 HWC → f32 / 255) → Resize(short side → S, bicubic, antialias — what
// torchvision 0.17 runs on a tensor) → CenterCrop(S) → Normalize(mean, std), as two HBM-bound passes over one image:
//   resize_rows_kernel   u8 [H,W,3] → f32 [3][rows][S]   horizontal taps, only the cropped columns, only the input
//                                                          rows the cropped output rows read; the row is staged in
//                                                          shared memory with 16-byte loads
//   resize_cols_kernel   f32 [3][rows][S] → f32 [3,S,S]   vertical taps + (x - mean) / std, coalesced along x
// The tap windows and weights restate ATen's separable anti-aliased kernel (area-pixel scale in/out, support
// 2·max(scale,1), cubic with a = -0.5, weights normalised to sum 1); every index decision is made with explicitly
// rounded fp32 operations (no FMA contraction) so the windows are the ones PyTorch picks on the CPU.
#include "common.h"

namespace teo {

struct ResizeGeom {
    int H, W;            // source image
    int nh, nw;          // size after Resize (host applies torchvision's short-side rule)
    int top, left;       // CenterCrop origin in the resized image
    int S;               // output side
    int y_lo, n_rows;    // source rows [y_lo, y_lo + n_rows) held in the intermediate
    float mean[3], std[3];
};

__device__ __forceinline__ float cubic_aa(float x) {          // a = -0.5 (the anti-aliased bicubic of ATen / PIL)
    x = fabsf(x);
    if (x < 1.f) return ((1.5f * x - 2.5f) * x) * x + 1.f;
    if (x < 2.f) return (((x - 5.f) * x + 8.f) * x - 4.f) * -0.5f;
    return 0.f;
}

struct TapWindow {
    int lo, n;
    float center, invscale;
    __device__ __forceinline__ float weight(int j) const {
        return cubic_aa(__fmul_rn(__fadd_rn(__fsub_rn(static_cast<float>(j + lo), center), 0.5f), invscale));
    }
    __device__ __forceinline__ float total() const {
        float t = 0.f;
        for (int j = 0; j < n; ++j) t = __fadd_rn(t, weight(j));
        return t;
    }
};

__device__ __forceinline__ TapWindow tap_window(int i, int in_size, int out_size) {
    const float scale = __fdiv_rn(static_cast<float>(in_size), static_cast<float>(out_size));
    const float support = scale >= 1.f ? __fmul_rn(2.f, scale) : 2.f;
    TapWindow t;
    t.center = __fmul_rn(scale, __fadd_rn(static_cast<float>(i), 0.5f));
    t.invscale = scale >= 1.f ? __fdiv_rn(1.f, scale) : 1.f;
    t.lo = max(static_cast<int>(__fadd_rn(__fsub_rn(t.center, support), 0.5f)), 0);
    t.n = min(static_cast<int>(__fadd_rn(__fadd_rn(t.center, support), 0.5f)), in_size) - t.lo;
    return t;
}

// one CTA per needed source row
__global__ void __launch_bounds__(256) resize_rows_kernel(const uint8_t* __restrict__ src, float* __restrict__ mid, ResizeGeom g) {
    extern __shared__ __align__(16) uint8_t row_raw[];
    const int y = g.y_lo + blockIdx.x;
    const int row_bytes = g.W * 3;
    const uint8_t* p = src + static_cast<size_t>(y) * row_bytes;
    // the shared copy starts at the same offset within a 16-byte line as the global row, so the aligned middle of the
    // row moves with 16-byte loads and stores; the ragged ends go byte by byte
    const int mis = static_cast<int>(reinterpret_cast<uintptr_t>(p) & 15);
    uint8_t* row = row_raw + mis;
    const int head = min((16 - mis) & 15, row_bytes);
    const int nvec = (row_bytes - head) / 16;
    for (int i = threadIdx.x; i < head; i += blockDim.x) row[i] = p[i];
    for (int i = threadIdx.x; i < nvec; i += blockDim.x)
        *reinterpret_cast<uint4*>(row + head + i * 16) = *reinterpret_cast<const uint4*>(p + head + i * 16);
    for (int i = head + nvec * 16 + threadIdx.x; i < row_bytes; i += blockDim.x) row[i] = p[i];
    __syncthreads();
    for (int xo = threadIdx.x; xo < g.S; xo += blockDim.x) {
        const TapWindow t = tap_window(g.left + xo, g.W, g.nw);
        const float tot = t.total();
        const float inv_ok = tot != 0.f ? 1.f : 0.f;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int j = 0; j < t.n; ++j) {
            float w = t.weight(j);
            if (inv_ok != 0.f) w = __fdiv_rn(w, tot);
            const uint8_t* px = row + (t.lo + j) * 3;
            a0 = fmaf(w, __fdiv_rn(static_cast<float>(px[0]), 255.f), a0);
            a1 = fmaf(w, __fdiv_rn(static_cast<float>(px[1]), 255.f), a1);
            a2 = fmaf(w, __fdiv_rn(static_cast<float>(px[2]), 255.f), a2);
        }
        const size_t o = static_cast<size_t>(blockIdx.x) * g.S + xo;
        const size_t plane = static_cast<size_t>(g.n_rows) * g.S;
        mid[o] = a0;
        mid[plane + o] = a1;
        mid[2 * plane + o] = a2;
    }
}

// grid (S output rows, 3 channels)
__global__ void __launch_bounds__(256) resize_cols_kernel(const float* __restrict__ mid, float* __restrict__ dst, ResizeGeom g) {
    const int yo = blockIdx.x, c = blockIdx.y;
    const TapWindow t = tap_window(g.top + yo, g.H, g.nh);
    const float tot = t.total();
    const float* plane = mid + static_cast<size_t>(c) * g.n_rows * g.S;
    const float mean = c == 0 ? g.mean[0] : c == 1 ? g.mean[1] : g.mean[2];      // constant indices: no local copy of g
    const float sd = c == 0 ? g.std[0] : c == 1 ? g.std[1] : g.std[2];
    for (int xo = threadIdx.x; xo < g.S; xo += blockDim.x) {
        float a = 0.f;
        for (int j = 0; j < t.n; ++j) {
            float w = t.weight(j);
            if (tot != 0.f) w = __fdiv_rn(w, tot);
            a = fmaf(w, plane[static_cast<size_t>(t.lo + j - g.y_lo) * g.S + xo], a);
        }
        dst[(static_cast<size_t>(c) * g.S + yo) * g.S + xo] = __fdiv_rn(__fsub_rn(a, mean), sd);
    }
}

// source rows the cropped output needs, with one row of slack either side (the device decides the exact windows)
static void needed_rows(int H, int nh, int top, int S, int* y_lo, int* n_rows) {
    const float scale = static_cast<float>(H) / static_cast<float>(nh);
    const float support = scale >= 1.f ? 2.f * scale : 2.f;
    const float c0 = scale * (static_cast<float>(top) + 0.5f), c1 = scale * (static_cast<float>(top + S - 1) + 0.5f);
    int lo = static_cast<int>(c0 - support + 0.5f) - 1, hi = static_cast<int>(c1 + support + 0.5f) + 1;
    lo = lo < 0 ? 0 : lo;
    hi = hi > H ? H : hi;
    *y_lo = lo;
    *n_rows = hi - lo;
}

}  // namespace teo

using namespace teo;

extern "C" size_t teo_resize_workspace_bytes(int H, int W, int S) {
    (void)W;
    if (H <= 0 || S <= 0) return 0;
    return static_cast<size_t>(3) * H * S * sizeof(float);
}

extern "C" int teo_resize_crop_normalize_u8(const void* src_u8, int H, int W, int nh, int nw, int top, int left, int S,
                                            const float* mean3, const float* std3, void* dst_f32, void* workspace,
                                            size_t workspace_bytes, void* stream) {
    TEO_CHECK_ARG(src_u8 && dst_f32 && mean3 && std3, "resize: null pointer");
    TEO_CHECK_ARG(H > 0 && W > 0 && S > 0 && nh >= S && nw >= S, "resize: bad sizes H=%d W=%d -> %dx%d, crop %d", H, W, nh, nw, S);
    TEO_CHECK_ARG(top >= 0 && left >= 0 && top + S <= nh && left + S <= nw, "resize: crop (%d,%d)+%d outside %dx%d", top, left, S, nh, nw);
    TEO_CHECK_ARG(static_cast<size_t>(W) * 3 + 16 <= 200 * 1024, "resize: image rows wider than %d pixels are not supported", (200 * 1024 - 16) / 3);
    TEO_CHECK_ARG(std3[0] != 0.f && std3[1] != 0.f && std3[2] != 0.f, "resize: zero std");
    ResizeGeom g{};
    g.H = H; g.W = W; g.nh = nh; g.nw = nw; g.top = top; g.left = left; g.S = S;
    needed_rows(H, nh, top, S, &g.y_lo, &g.n_rows);
    for (int c = 0; c < 3; ++c) { g.mean[c] = mean3[c]; g.std[c] = std3[c]; }
    const size_t need = static_cast<size_t>(3) * g.n_rows * S * sizeof(float);
    if (workspace == nullptr || workspace_bytes < need) {
        set_error("resize: workspace of %zu bytes needed, got %zu", need, workspace_bytes);
        return TEO_ERR_WORKSPACE;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t smem = static_cast<size_t>(W) * 3 + 16;
    if (smem > 48 * 1024) TEO_CUDA(cudaFuncSetAttribute(resize_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    resize_rows_kernel<<<g.n_rows, 256, smem, s>>>(static_cast<const uint8_t*>(src_u8), static_cast<float*>(workspace), g);
    TEO_LAUNCH_CHECK("resize_rows_kernel");
    resize_cols_kernel<<<dim3(S, 3), 256, 0, s>>>(static_cast<const float*>(workspace), static_cast<float*>(dst_f32), g);
    TEO_LAUNCH_CHECK("resize_cols_kernel");
    return TEO_OK;
}
