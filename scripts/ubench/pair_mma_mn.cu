// Experiment (debug tooling): MN-major (sample-contiguous) B operand of tcgen05.mma.cta_group::2 -- which shared-memory layout and
// descriptor fields the hardware expects when the activations are stored [k][sample] instead of [sample][k].
// Same set-up as pair_mma.cu (CTA r: A_r 128 x 64 K-major SW128, B_r = 32 samples x 64 k); the B placement and the descriptor
// (layout type, LBO, SBO, per-k-step advance) are run-time parameters so that several hypotheses run in one go.
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
__device__ __forceinline__ uint32_t sw128(int row, int k) { return row * 128 + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2; }

struct Cfg {
    int mode;          // 0: K-major SW128 (reference), 1: MN-major no swizzle, 2: MN-major SW64
    uint32_t lbo, sbo; // descriptor fields, bytes
    uint32_t layout;   // descriptor layout type
    uint32_t kstep;    // bytes the start address advances per K = 16 step
    uint32_t bmajor;   // idesc bit 16
    uint32_t p_lbo, p_sbo;   // strides used to PLACE the data (mode 1: k-group stride, n-group stride; mode 2: k-group stride)
};

__device__ __forceinline__ uint32_t place(const Cfg& c, int n, int k) {     // byte offset of B element (sample n, k)
    if (c.mode == 0) return sw128(n, k);
    if (c.mode == 1) return (n >> 3) * c.p_sbo + (k >> 3) * c.p_lbo + (k & 7) * 16 + (n & 7) * 2;
    return (k >> 3) * c.p_sbo + (k & 7) * 64 + ((((n >> 3) ^ ((k >> 1) & 3)) & 3) << 4) + (n & 7) * 2;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, Cfg c) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = base;              // 16 KB
    unsigned char* sB = base + 16384;      // 8 KB reserved
    __shared__ uint64_t bar;
    __shared__ uint32_t holder;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
        const int r = i / 64, kk = i % 64;
        *reinterpret_cast<__nv_bfloat16*>(sA + sw128(r, kk)) = A[(rank * 128 + r) * 64 + kk];
    }
    for (int i = threadIdx.x; i < 8192 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sB)[i] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) {
        const int r = i / 64, kk = i % 64;
        *reinterpret_cast<__nv_bfloat16*>(sB + place(c, r, kk)) = B[(rank * 32 + r) * 64 + kk];
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
            const uint32_t idesc = make_idesc(256, 64) | (c.bmajor << 16);
            const uint64_t ad = make_desc_sw128(smem_u32(sA));
            uint64_t bd = 0;
            bd |= (uint64_t)((smem_u32(sB) & 0x3FFFF) >> 4);
            bd |= (uint64_t)(c.lbo >> 4) << 16;
            bd |= (uint64_t)(c.sbo >> 4) << 32;
            bd |= (uint64_t)1 << 46;
            bd |= (uint64_t)c.layout << 61;
            for (int ks = 0; ks < 4; ++ks) umma_bf16_2cta(tm, ad + 2 * ks, bd + (uint64_t)((c.kstep >> 4) * ks), idesc, ks ? 1u : 0u);
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
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
    struct Named { const char* name; Cfg c; };
    const Named cfgs[] = {
        {"K-major SW128 (reference)",                          {0, 16, 1024, 2, 32, 0, 0, 0}},
        {"MN no-swizzle, LBO=k-group 128, SBO=n-group 1024",   {1, 128, 1024, 0, 256, 1, 128, 1024}},
        {"MN no-swizzle, fields swapped (LBO=1024, SBO=128)",  {1, 1024, 128, 0, 256, 1, 128, 1024}},
        {"MN SW64, SBO=k-group 512, LBO=4096",                 {2, 4096, 512, 4, 1024, 1, 0, 512}},
        {"MN SW64, fields swapped (LBO=512, SBO=4096)",        {2, 512, 4096, 4, 1024, 1, 0, 512}},
    };
    for (const Named& nc : cfgs) {
        cudaMemset(dD, 0xff, 2 * 128 * 64 * 4);
        k<<<2, 128, 40960>>>(dA, dB, dD, nc.c);
        std::vector<float> hD(2 * 128 * 64);
        cudaError_t e = cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("%s: kernel error %s\n", nc.name, cudaGetErrorString(e)); return 1; }
        double maxerr = 0;
        for (int r = 0; r < 2; ++r) for (int i = 0; i < 128; ++i) for (int j = 0; j < 64; ++j) {
            double acc = 0;
            for (int kk = 0; kk < 64; ++kk) acc += (double)fA[(r * 128 + i) * 64 + kk] * fB[j * 64 + kk];
            maxerr = fmax(maxerr, fabs(acc - hD[(r * 128 + i) * 64 + j]));
        }
        printf("%-55s max |err| = %g\n", nc.name, maxerr);
    }
    return 0;
}
