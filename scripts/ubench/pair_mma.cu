// Experiment (debug tooling): semantics of tcgen05.mma.cta_group::2 -- which CTA supplies which rows of A and B, and where D lands.
// Cluster of 2 CTAs.  CTA r holds A_r (128 x 64 bf16, K-major, 128B swizzle) and B_r (32 x 64).  The leader (rank 0) issues
// M = 256, N = 64, K = 64 (4 instructions); both CTAs then dump their TMEM (128 lanes x 64 columns) to global memory.
// Hypothesis checked on the host: D_r[i][j] = sum_k A_r[i][k] * Bcat[j][k], Bcat = [B_0 ; B_1].
#include <cstdio>
#include <vector>
#include <cmath>
#include "tc_common.cuh"
namespace dfw { void set_error(const char* f, ...) { printf("error: %s\n", f); } std::atomic<long long> g_launches{0}; }
using namespace dfw::tc;

__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

// element (row, k) of a K-major SW128 tile: row * 128 + ((k / 8) ^ (row % 8)) * 16 + (k % 8) * 2
__device__ __forceinline__ uint32_t sw128(int row, int k) { return row * 128 + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2; }

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {   // A: [2][128][64], B: [2][32][64], D: [2][128][64]
    extern __shared__ unsigned char raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = base;              // 16 KB
    unsigned char* sB = base + 16384;      // 4 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
        const int r = i / 64, kk = i % 64;
        *reinterpret_cast<__nv_bfloat16*>(sA + sw128(r, kk)) = A[(rank * 128 + r) * 64 + kk];
    }
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) {
        const int r = i / 64, kk = i % 64;
        *reinterpret_cast<__nv_bfloat16*>(sB + sw128(r, kk)) = B[(rank * 32 + r) * 64 + kk];
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&holder)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tm = holder;
    if (rank == 0 && threadIdx.x < 32) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc(256, 64);
            const uint64_t ad = make_desc_sw128(smem_u32(sA)), bd = make_desc_sw128(smem_u32(sB));
            for (int ks = 0; ks < 4; ++ks) umma_bf16_2cta(tm, ad + 2 * ks, bd + 2 * ks, idesc, ks ? 1u : 0u);
            umma_commit_2cta(&bar, 3);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0, nullptr, 0);
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t v[16];
    for (int c0 = 0; c0 < 64; c0 += 16) {
        tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) D[(rank * 128 + warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64) : "memory");
}

int main() {
    std::vector<__nv_bfloat16> hA(2 * 128 * 64), hB(2 * 32 * 64);
    std::vector<float> fA(hA.size()), fB(hB.size());
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((int)((s >> 16) & 15) - 8) * 0.125f; };
    for (size_t i = 0; i < hA.size(); ++i) { fA[i] = rnd(); hA[i] = __float2bfloat16(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { fB[i] = rnd(); hB[i] = __float2bfloat16(fB[i]); }
    __nv_bfloat16 *dA, *dB; float* dD;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 2 * 128 * 64 * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, 2 * 128 * 64 * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    k<<<2, 128, 32768>>>(dA, dB, dD);
    std::vector<float> hD(2 * 128 * 64);
    cudaError_t e = cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    // hypotheses for column j of CTA r's rows: (a) Bcat[j] = [B0;B1][j]; (b) swapped halves
    for (int hyp = 0; hyp < 2; ++hyp) {
        double maxerr = 0;
        for (int r = 0; r < 2; ++r) for (int i = 0; i < 128; ++i) for (int j = 0; j < 64; ++j) {
            const int jj = hyp == 0 ? j : (j + 32) % 64;
            double acc = 0;
            for (int kk = 0; kk < 64; ++kk) acc += (double)fA[(r * 128 + i) * 64 + kk] * fB[jj * 64 + kk];
            maxerr = fmax(maxerr, fabs(acc - hD[(r * 128 + i) * 64 + j]));
        }
        printf("hypothesis %d (D_r[i][j] = A_r[i] . B%s[j]): max |err| = %g\n", hyp, hyp ? "(halves swapped)" : "cat", maxerr);
    }
    printf("sample D[0][0..3] = %g %g %g %g ; D[128][0..3] = %g %g %g %g\n", hD[0], hD[1], hD[2], hD[3], hD[128 * 64], hD[128 * 64 + 1], hD[128 * 64 + 2], hD[128 * 64 + 3]);
    return 0;
}
