// Micro-benchmark (debug tooling): back-to-back launch cost of an (almost) empty kernel as a function of parameter bytes,
// cluster size, dynamic shared memory and TMEM allocation -- what the fused kernel's launch configuration costs per launch.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
struct Small { int x[16]; };
struct Big { int x[1800]; };
template <class P, bool TMEM>
__global__ void __launch_bounds__(512, 1) k(const __grid_constant__ P p, int* out) {
    extern __shared__ unsigned char smem[];
    __shared__ uint32_t holder;
    if (TMEM) {
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&holder)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(holder), "r"(512) : "memory");
    }
    if (threadIdx.x == 0 && p.x[0] == 12345) out[blockIdx.x] = (int)smem[0];
}
template <class P, bool TMEM>
void run(const char* name, int cl, int smem) {
    P p = {};
    int* d; cudaMalloc(&d, 1024);
    cudaFuncSetAttribute(k<P, TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(128); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 20; ++i) cudaLaunchKernelEx(&cfg, k<P, TMEM>, p, d);
    cudaEventRecord(a);
    for (int i = 0; i < 500; ++i) cudaLaunchKernelEx(&cfg, k<P, TMEM>, p, d);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // same through a CUDA graph
    cudaStream_t st; cudaStreamCreate(&st); cfg.stream = st;
    cudaGraph_t g = nullptr; cudaGraphExec_t ge = nullptr;
    float ms2 = -1.f;
    cudaError_t e1 = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int i = 0; i < 50 && e1 == cudaSuccess; ++i) e1 = cudaLaunchKernelEx(&cfg, k<P, TMEM>, p, d);
    cudaError_t e2 = cudaStreamEndCapture(st, &g);
    if (e1 == cudaSuccess && e2 == cudaSuccess && g && cudaGraphInstantiate(&ge, g, 0) == cudaSuccess && ge) {
        cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
        cudaEventRecord(a, st);
        for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, st);
        cudaEventRecord(b, st); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms2, a, b);
    } else printf("  (graph capture failed: %s / %s)\n", cudaGetErrorString(e1), cudaGetErrorString(e2));
    printf("%-34s cluster %d smem %3d KB : %6.2f us/launch (stream)  %6.2f us/launch (graph)  %s\n", name, cl, smem / 1024, ms * 2.0f, ms2 * 2.0f,
           cudaGetLastError() == cudaSuccess ? "" : "ERROR");
    fflush(stdout);
}
int main() {
    for (int cl : {1, 4}) for (int smem : {0, 220 * 1024}) {
        run<Small, false>("64 B params", cl, smem);
        run<Big, false>("7.2 KB params", cl, smem);
        run<Small, true>("64 B params + TMEM alloc", cl, smem);
        run<Big, true>("7.2 KB params + TMEM alloc", cl, smem);
    }
    return 0;
}
