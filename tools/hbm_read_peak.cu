// Development tool: read-only HBM streaming ceilings on this GPU, to put the decode-attention kernel's
// achieved GB/s in context (MEASURED_PEAKS.json's figure is a copy: half reads, half writes).
//   (a) LDG.128 grid-stride reads          (b) cp.async.bulk 16 KiB tiles into a shared-memory ring
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/hbm_read_peak tools/hbm_read_peak.cu
#include <cstdio>
#include <cstdlib>

#include "../teochat_b200/csrc/ptx.cuh"

using namespace teo;

__global__ void __launch_bounds__(256) ldg_kernel(const uint4* __restrict__ src, size_t n16, unsigned* sink) {
    unsigned acc = 0;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        const uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        acc ^= a.x ^ b.y ^ c.z ^ d.w;
    }
    for (; i < n16; i += stride) acc ^= src[i].x;
    if (acc == 0x12345678u) *sink = acc;
}

// each CTA streams a contiguous region in TILE-byte pieces through a STAGES-deep ring
template <int TILE, int STAGES>
__global__ void __launch_bounds__(128) bulk_kernel(const uint8_t* __restrict__ src, size_t bytes_per_cta, unsigned* sink) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 127) & ~uintptr_t(127));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + STAGES * TILE);
    const uint8_t* base = src + static_cast<size_t>(blockIdx.x) * bytes_per_cta;
    const int n = static_cast<int>(bytes_per_cta / TILE);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES && s < n; ++s) {
            mbar_arrive_expect_tx(&bar[s], TILE);
            bulk_load_1d(smem + s * TILE, base + static_cast<size_t>(s) * TILE, TILE, &bar[s]);
        }
    unsigned acc = 0;
    for (int i = 0; i < n; ++i) {
        const int s = i % STAGES;
        mbar_wait(&bar[s], (i / STAGES) & 1);
        acc ^= reinterpret_cast<const unsigned*>(smem + s * TILE)[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0 && i + STAGES < n) {
            mbar_arrive_expect_tx(&bar[s], TILE);
            bulk_load_1d(smem + s * TILE, base + static_cast<size_t>(i + STAGES) * TILE, TILE, &bar[s]);
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// The decode-attention access pattern: CTA = (head, group of pages); per page one 16 KiB K slice and one 16 KiB V slice.
// layout 0: pool [page][K|V][32 heads][16 KiB]  (slices of one head 1 MiB apart, K and V 512 KiB apart)
// layout 1: pool [page][32 heads][K|V][16 KiB]  (K|V adjacent: 32 KiB runs, 1 MiB apart)
// layout 2: pool [32 heads][page][K|V][16 KiB]  (one head's pages contiguous: a linear stream per CTA)
template <int STAGES>
__global__ void __launch_bounds__(128) paged_kernel(const uint8_t* __restrict__ src, int pages_per_cta, int n_pages_total, int layout,
                                                    unsigned* sink) {
    constexpr int TILE = 16384;
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 127) & ~uintptr_t(127));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + STAGES * 2 * TILE);
    const int head = blockIdx.x % 32, grp = blockIdx.x / 32;
    auto addr = [&](int p, int kv) -> const uint8_t* {
        const size_t page = static_cast<size_t>(grp) * pages_per_cta + p;
        size_t slice;
        if (layout == 0) slice = (page * 2 + kv) * 32 + head;
        else if (layout == 1) slice = (page * 32 + head) * 2 + kv;
        else slice = (static_cast<size_t>(head) * n_pages_total + page) * 2 + kv;
        return src + slice * TILE;
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar[s], 1);
        fence_barrier_init();
    }
    __syncthreads();
    auto issue = [&](int p) {
        const int s = p % STAGES;
        mbar_arrive_expect_tx(&bar[s], 2 * TILE);
        bulk_load_1d(smem + s * 2 * TILE, addr(p, 0), TILE, &bar[s]);
        bulk_load_1d(smem + s * 2 * TILE + TILE, addr(p, 1), TILE, &bar[s]);
    };
    if (threadIdx.x == 0)
        for (int p = 0; p < STAGES && p < pages_per_cta; ++p) issue(p);
    unsigned acc = 0;
    for (int p = 0; p < pages_per_cta; ++p) {
        const int s = p % STAGES;
        mbar_wait(&bar[s], (p / STAGES) & 1);
        acc ^= reinterpret_cast<const unsigned*>(smem + s * 2 * TILE)[threadIdx.x];
        __syncthreads();
        if (threadIdx.x == 0 && p + STAGES < pages_per_cta) issue(p + STAGES);
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <typename F>
static float time_ms(F f, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / iters;
}

template <int TILE, int STAGES>
static void run_bulk(const uint8_t* buf, size_t bytes, unsigned* sink, int ctas_per_sm) {
    const int smem = STAGES * TILE + 128 + 64;
    cudaFuncSetAttribute(bulk_kernel<TILE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int grid = 148 * ctas_per_sm;
    const size_t per = (bytes / grid) / TILE * TILE;
    const float ms = time_ms([&] { bulk_kernel<TILE, STAGES><<<grid, 128, smem>>>(buf, per, sink); }, 5);
    printf("bulk tile %5d B x %d stages, %d CTA/SM (%3d KiB in flight/SM): %7.1f GB/s\n", TILE, STAGES, ctas_per_sm,
           ctas_per_sm * STAGES * TILE / 1024, per * grid / ms / 1e6);
}

int main() {
    const size_t bytes = 8ull << 30;
    uint8_t* buf;
    unsigned* sink;
    cudaMalloc(&buf, bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    for (int mult : {4, 8, 16}) {
        const float ms = time_ms([&] { ldg_kernel<<<148 * mult, 256>>>(reinterpret_cast<const uint4*>(buf), bytes / 16, sink); }, 5);
        printf("LDG.128 grid 148x%-2d: %7.1f GB/s\n", mult, bytes / ms / 1e6);
    }
    run_bulk<16384, 2>(buf, bytes, sink, 3);
    run_bulk<16384, 4>(buf, bytes, sink, 3);
    run_bulk<16384, 4>(buf, bytes, sink, 2);
    run_bulk<16384, 6>(buf, bytes, sink, 2);
    run_bulk<32768, 2>(buf, bytes, sink, 3);
    run_bulk<32768, 3>(buf, bytes, sink, 2);
    run_bulk<32768, 6>(buf, bytes, sink, 1);
    run_bulk<8192, 8>(buf, bytes, sink, 3);
    {
        // 3 CTAs/SM × 148 = 444 CTAs = 32 heads × 13 page groups (416 CTAs) — close to the decode launch
        constexpr int ST = 2;
        const int smem = ST * 2 * 16384 + 128 + 64;
        cudaFuncSetAttribute(paged_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        const int groups = 13, grid = 32 * groups;
        const int n_pages_total = static_cast<int>(bytes / (32 * 2 * 16384));      // 8 GiB / 1 MiB per page = 8192 pages
        const int ppc = n_pages_total / groups;
        for (int layout = 0; layout < 3; ++layout) {
            const float ms = time_ms([&] { paged_kernel<ST><<<grid, 128, smem>>>(buf, ppc, n_pages_total, layout, sink); }, 5);
            printf("paged pattern, layout %d, %d CTAs x %d pages x 32 KiB, 2 stages: %7.1f GB/s\n", layout, grid, ppc,
                   static_cast<double>(grid) * ppc * 32768 / ms / 1e6);
        }
        // the same pattern at the size of ONE decode-attention launch (≈ 1.2 GB): launch ramp and tail included
        for (int pages : {87, 174, 348}) {
            const float ms = time_ms([&] { paged_kernel<ST><<<grid, 128, smem>>>(buf, pages, n_pages_total, 0, sink); }, 20);
            printf("paged pattern, layout 0, %d pages per CTA (%.2f GB per launch): %6.1f us, %7.1f GB/s\n", pages,
                   static_cast<double>(grid) * pages * 32768 / 1e9, ms * 1e3, static_cast<double>(grid) * pages * 32768 / ms / 1e6);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
