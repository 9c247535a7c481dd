// Micro-benchmark (debug tooling): tcgen05.mma throughput on static shared-memory operands, lean elect-guarded issue.
//   For M in {128, 64}, N in {32 .. 256}: R MMAs (K = 16, bf16), the A tile advancing through a ring of 6 x 16 KB
//   (4 k-steps per tile, like a streamed weight operand) or fixed; B fixed.  Prints cycles per MMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I xsdeepfwfm_deprecated_b200/csrc scripts/ubench/mma_rate.cu -o scripts/ubench/build/mma_rate
#include <cstdio>
#include "tc_common.cuh"
namespace dfw { void set_error(const char*, ...) {} std::atomic<long long> g_launches{0}; }
using namespace dfw::tc;

template <int N>
__global__ void __launch_bounds__(128, 1) k(int M, int R, int ring, int commit_every, long long* out) {
    const int b_ring = 1;
    extern __shared__ unsigned char raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = base;                 // 6 x 16 KB A tiles
    unsigned char* sB = base + 6 * 16384;     // 3 x 32 KB B tiles (256 rows x 128 B)
    __shared__ uint64_t bar, dummy[8];
    __shared__ uint32_t holder;
    for (int i = threadIdx.x; i < (6 * 16384 + 3 * 32768) / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 8; ++i) mbar_init(&dummy[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) tmem_alloc(&holder, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = holder;
    if (threadIdx.x < 32) {
        const uint32_t idesc = make_idesc(M, N);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        uint32_t phase = 0;
        for (int rep = 0; rep < 3; ++rep) {
            long long t0 = clock64();
            uint32_t st = 0, bs = 0;
            for (int r = 0; r < R; r += 4) {
                const uint64_t ad = make_desc_sw128(a0 + st * 16384u), bd = make_desc_sw128(b0 + bs * 32768u);
                if (elect_one()) {
                    umma_bf16(tm, ad, bd, idesc, 1u);
                    umma_bf16(tm, ad + 2, bd + 2, idesc, 1u);
                    umma_bf16(tm, ad + 4, bd + 4, idesc, 1u);
                    umma_bf16(tm, ad + 6, bd + 6, idesc, 1u);
                    if (commit_every && ((r >> 2) % commit_every) == commit_every - 1) umma_commit(&dummy[(r >> 2) & 7]);
                }
                __syncwarp();
                if (++st >= (uint32_t)ring) st = 0;
                if (++bs >= (uint32_t)b_ring) bs = 0;
            }
            long long t1 = clock64();
            if (elect_one()) umma_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, phase, nullptr, 0);
            phase ^= 1;
            long long t2 = clock64();
            if (threadIdx.x == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N> void run(long long* d) {
    const int SM = 6 * 16384 + 3 * 32768 + 2048, R = 512;
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
    for (int M : {128}) for (int ring : {6}) for (int b_ring : {0, 1, 2, 4}) {
        k<N><<<1, 128, SM>>>(M, R, ring, b_ring, d);
        long long h[8]; cudaError_t e = cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
        printf("%3d %4d   A ring %d  commit every %d stages   issue %7.1f  total %7.1f cyc/mma   (math floor %d)\n", M, N, ring, b_ring, h[4] / (double)R, h[5] / (double)R, N / 2);
    }
}

int main() {
    long long* d; cudaMalloc(&d, 64);
    run<32>(d); run<64>(d); run<128>(d); run<256>(d);
    return 0;
}
