// Micro-benchmark (debug tooling): how fast can every SM stream the SAME weight matrix from L2 into shared memory with TMA?
// G CTAs (cluster 1 / 2 / 4, multicast like the fused kernel), each pulling `boxes` 16 KB boxes (128 rows x 64 bf16) through a
// ring of `nstage` stages; the consumer only waits and releases.  Prints bytes/clk/SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I include -I xsdeepfwfm_deprecated_b200/csrc scripts/ubench/tma_ingest.cu -o scripts/ubench/build/tma_ingest
#include <cstdio>
#include <vector>
#include "tc_common.cuh"
namespace dfw { void set_error(const char* f, ...) { printf("error: %s\n", f); } std::atomic<long long> g_launches{0}; }
using namespace dfw::tc;

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap map, int boxes, int nstage, int cl, int box_rows, long long* out) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[12], empty[12];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstage; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], cl); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (cl > 1) cluster_sync_all();
    const uint32_t crank = cl > 1 ? cluster_ctarank() : 0;
    const uint16_t mask = (uint16_t)((1u << cl) - 1);
    const int per = box_rows / cl;
    const uint32_t box_bytes = (uint32_t)box_rows * 128u;
    long long t0 = clock64();
    if (warp == 0) {
        uint32_t st = 0, ph = 0;
        int mt = 0, c = 0;                                             // 3 neuron tiles x 7 K chunks of one 400 x 448 layer image
        for (int b = 0; b < boxes; ++b) {
            if (++c == 7) { c = 0; if (++mt == 3) mt = 0; }
            mbar_wait(&empty[st], ph ^ 1, nullptr, 0);
            if (elect_one()) {
                mbar_expect_tx(&full[st], box_bytes);
                unsigned char* dst = base + st * box_bytes + crank * per * 128;
                if (cl > 1) tma_load_2d_mc(dst, &map, &full[st], c * 64, mt * box_rows + (int)crank * per, mask);
                else tma_load_2d(dst, &map, &full[st], c * 64, mt * box_rows);
            }
            __syncwarp();
            if (++st == (uint32_t)nstage) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        uint32_t st = 0, ph = 0;
        for (int b = 0; b < boxes; ++b) {
            mbar_wait(&full[st], ph, nullptr, 0);
            if (cl > 1) {
                if (lane < cl) {                     // release the stage in every CTA of the cluster
                    uint32_t remote;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(&empty[st])), "r"(lane));
                    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
                }
            } else if (lane == 0) mbar_arrive(&empty[st]);
            __syncwarp();
            if (++st == (uint32_t)nstage) { st = 0; ph ^= 1; }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (cl > 1) cluster_sync_all();
}

int main() {
    const int rows = 768, cols = 448;
    void* W; cudaMalloc(&W, rows * cols * 2); cudaMemset(W, 0, rows * cols * 2);
    long long* d; cudaMalloc(&d, 148 * 8);
    const int SM = 12 * 16384 + 2048;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SM);
    const int boxes = 21 * 20;
    for (int cl : {1}) for (int box_rows : {128, 256}) for (int nstage : {2, 3, 4, 5, 6, 8, 12}) for (int grid : {64, 148}) {
        if ((size_t)nstage * box_rows * 128 > 12 * 16384) continue;
        CUtensorMap map;
        if (make_map(&map, W, rows, cols, cols, box_rows / cl)) return 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = SM;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, k, map, boxes, nstage, cl, box_rows, d);
        std::vector<long long> h(grid);
        cudaError_t e = cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        long long mx = 0; for (auto v : h) mx = v > mx ? v : mx;
        printf("cluster %d  box %3d rows  stages %2d  grid %3d : %8lld cycles (%5.0f per box) -> %6.1f B/clk/SM\n", cl, box_rows, nstage, grid, mx, (double)mx / boxes, boxes * box_rows * 128.0 / mx);
    }
    return 0;
}
