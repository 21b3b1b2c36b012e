// Hardware probe (development tool, not product): checks the two tcgen05 operand forms the
// flash-attention kernel depends on, against a CPU reference —
//   (1) B operand MN-major under the 128-byte swizzle (V tile stored [key][head_dim]),
//   (2) A operand taken from tensor memory (P written with tcgen05.st, two bf16 per column).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_build/probe_umma tools/probe_umma.cu
// Run:   tools/_build/probe_umma <use_ts 0|1> <N 64|128> <lbo> <sbo> <kstep_bytes> <swap_pack 0|1>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../teochat_b200/csrc/ptx.cuh"

using namespace teo;
typedef __nv_bfloat16 bf16;

constexpr int M = 128, K = 64;

__global__ void __launch_bounds__(128) probe_kernel(const bf16* __restrict__ A, const bf16* __restrict__ V, float* __restrict__ D,
                                                    int N, int use_ts, uint32_t lbo, uint32_t sbo, uint32_t kstep, int swap_pack) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // [128 rows][64 k] K-major, swizzled: 16 KiB
    uint8_t* sV = smem + 16384;         // halves of [64 keys][64 cols], 16 KiB apart (matches the attention kernel)
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<256>(&tslot);
    // A → smem (K-major SW128): row r, 16-byte chunk c at r*128 + ((c ^ (r & 7)) << 4)
    for (int i = tid; i < M * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * K + c * 8);
    }
    // V → smem: half nh (64 columns), key row k: k*128 + ((c ^ (k & 7)) << 4)
    for (int i = tid; i < K * (N / 8); i += 128) {
        const int k = i / (N / 8), cc = i % (N / 8);
        const int nh = cc >> 3, c = cc & 7;
        *reinterpret_cast<uint4*>(sV + nh * 16384 + k * 128 + ((c ^ (k & 7)) << 4)) = *reinterpret_cast<const uint4*>(V + k * N + cc * 8);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tslot;
    const uint32_t t_a = tbase + 128;      // A operand columns (TS form)
    if (use_ts) {
        uint32_t v[32];
        for (int j = 0; j < 32; ++j) {
            const bf16 lo = A[tid * K + 2 * j], hi = A[tid * K + 2 * j + 1];
            const uint32_t l16 = *reinterpret_cast<const uint16_t*>(&lo), h16 = *reinterpret_cast<const uint16_t*>(&hi);
            v[j] = swap_pack ? ((l16 << 16) | h16) : ((h16 << 16) | l16);
        }
        tmem_st_32x32(t_a + (static_cast<uint32_t>(warp * 32) << 16), v);
        tmem_st_wait();
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_bf16(128, N) | UMMA_IDESC_B_MN_MAJOR;
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t bd = umma_desc_mn_sw128(smem_u32(sV) + ks * kstep, lbo, sbo);
            if (use_ts) umma_bf16_ts(tbase, t_a + 8 * ks, bd, idesc, ks > 0);
            else umma_bf16(tbase, umma_desc_k_sw128(smem_u32(sA)) + 2 * ks, bd, idesc, ks > 0);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tbase + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tbase);
}

int main(int argc, char** argv) {
    const int use_ts = argc > 1 ? atoi(argv[1]) : 0;
    const int N = argc > 2 ? atoi(argv[2]) : 64;
    const uint32_t lbo = argc > 3 ? atoi(argv[3]) : 16384;
    const uint32_t sbo = argc > 4 ? atoi(argv[4]) : 1024;
    const uint32_t kstep = argc > 5 ? atoi(argv[5]) : 2048;
    const int swap_pack = argc > 6 ? atoi(argv[6]) : 0;
    std::vector<bf16> hA(M * K), hV(K * N);
    std::vector<float> fA(M * K), fV(K * N), ref(M * N, 0.f), out(M * N);
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 32768.0f - 1.0f; };
    for (int i = 0; i < M * K; ++i) { hA[i] = __float2bfloat16(rnd()); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < K * N; ++i) { hV[i] = __float2bfloat16(rnd()); fV[i] = __bfloat162float(hV[i]); }
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < K; ++k)
            for (int n = 0; n < N; ++n) ref[m * N + n] += fA[m * K + k] * fV[k * N + n];
    bf16 *dA, *dV;
    float* dD;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dV, K * N * 2); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dV, hV.data(), K * N * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, M * N * 4);
    const int smem = 16384 + 32768 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(dA, dV, dD, N, use_ts, lbo, sbo, kstep, swap_pack);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe ts=%d N=%d lbo=%u sbo=%u kstep=%u swap=%d: CUDA error %s\n", use_ts, N, lbo, sbo, kstep, swap_pack, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(out.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int i = 0; i < M * N; ++i) { maxerr = fmax(maxerr, fabs(out[i] - ref[i])); maxref = fmax(maxref, fabs(ref[i])); }
    printf("probe ts=%d N=%d lbo=%u sbo=%u kstep=%u swap=%d: max|err| %.4g (max|ref| %.3g) %s\n", use_ts, N, lbo, sbo, kstep, swap_pack, maxerr,
           maxref, maxerr < 1e-3 * maxref ? "MATCH" : "mismatch");
    return maxerr < 1e-3 * maxref ? 0 : 1;
}
