// Definitions shared by the fused forward kernels (fused_tc.cu: one CTA per 32-sample tile; fused_pair.cu: a cta_group::2 pair of
// CTAs per 64 samples).
#pragma once
#include <stdlib.h>
#include <string.h>

#include "embed_device.cuh"
#include "tc_common.cuh"

namespace dfw {
namespace fz {

using namespace dfw::tc;

constexpr int TS = 32;                         // samples per tile
constexpr int G_WARPS = 10;
constexpr int G_THREADS = 32 * G_WARPS;
constexpr int EPI_WARPS = 4;
constexpr int EPI_THREADS = 32 * EPI_WARPS;
constexpr int RINGS = 2;                       // weight rings; one producer warp and one MMA warp per ring
constexpr int MMA_WARP0 = RINGS;               // producers: warps 0..1, MMA issuers: warps 2..3
constexpr int EPI_WARP0 = 2 * RINGS;           // epilogue warps 4..7: warp % 4 == TMEM lane quarter
constexpr int G_WARP0 = EPI_WARP0 + EPI_WARPS; // gather warps 8..17
constexpr int NTHREADS = 32 * G_WARP0 + G_THREADS;         // 576 (18 warps: 96 registers per thread)
constexpr int STAGE_BYTES = 128 * 128;         // one weight box: 128 neurons x 64 bf16
constexpr int RING_MAX = 6;                    // stages per ring
constexpr int MAX_L = 4;
constexpr int MAX_MT = 4;                      // 128-neuron tiles per layer (width <= 512)
constexpr int MAX_W = 512;
constexpr int G_ROUNDS = 2;                    // phase-D rounds of the gather group (generic shapes): K <= G_ROUNDS * G_WARPS
constexpr int BAR_GATHER = 1, BAR_EPI = 2, BAR_INIT = 3, BAR_CORE = 4;     // named barriers
constexpr int MAX_KCH = 8;                     // 64-wide K chunks of the widest operand (512)
constexpr size_t SMEM_LIMIT = 227 * 1024;

struct alignas(64) Maps {
    CUtensorMap w[MAX_L][2][2];                // [layer][hi | lo][128-row tile | 64-row last tile]
};

struct Params {
    EmbedParams ep;
    int depth, in_dim;
    int widths[MAX_L];
    const float* bias[MAX_L];
    const float* fc;
    float* logits;
    float* prob;
    long long B;
    int num_tiles, cluster;
    int x_chunks;                               // 64-wide K chunks of the activation buffer
    int nst[2];                                 // stages of ring 0 / ring 1
    uint32_t oE, oRing, oImg, oPart, oIdx, oXv, oMisc;   // shared-memory offsets from the 1024-aligned base (X is at 0)
    int* err;
    long long* clk;                             // optional per-CTA timeline (debug tooling): FZ_NCLK x int64 per CTA
    volatile int* prog;                         // optional progress markers in pinned host memory (debug): 32 ints per CTA
};
#define FZ_PROG(slot, val) do { if (p.prog && lane == 0) p.prog[blockIdx.x * 32 + (slot)] = (val); } while (0)
#define FZ_NCLK 128
#define FZ_CLK(slot) do { if (p.clk) p.clk[blockIdx.x * FZ_NCLK + (slot)] = clock64(); } while (0)

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}

// neuron tiles of a layer of (padded) width npad: 128 rows each; a last tile of <= 16 rows runs as M = 64
__host__ __device__ inline int n_mtiles(int npad) { return (npad + 127) / 128; }
__host__ __device__ inline int mtile_rows(int npad, int mt) {
    const int rem = npad - mt * 128;
    return rem >= 128 ? 128 : (rem <= 16 ? 64 : 128);
}

// The symmetrised field matrix as a kernel parameter (same column layout as the shallow image's U).  valid = 0: absent.
constexpr int MAX_U = 1168;                    // usize(47) + 4 = 1156 floats (usize(39) = 800)
struct alignas(16) UParam { int valid; int pad_[3]; float u[MAX_U]; };

// U = strict upper triangle of (R + R^T) / 2 by columns, in fp32 exactly as pack_shallow_kernel computes it (valid = 1), or ones
// for FM (model/DeepFMs.py:353-355; valid = 2).  valid = 0: no host snapshot of field_cov, or F too large for the parameter.
inline void build_uparam(const dfw_model* m, UParam& up) {
    const int F = m->field_size;
    up.valid = 0;
    if (usize(F) + 4 > MAX_U) return;
    if ((m->flags & DFW_USE_FWFM) && m->field_cov_host) {
        const float* cov = m->field_cov_host;
        for (int j = 1; j < F; ++j)
            for (int i = 0; i < pad4(j); ++i)
                up.u[ucol_off(j) + i] = i < j ? (cov[j * F + i] + cov[i * F + j]) * 0.5f : 0.f;
        up.valid = 1;
    } else if (!(m->flags & DFW_USE_FWFM)) {
        for (int j = 1; j < F; ++j)
            for (int i = 0; i < pad4(j); ++i) up.u[ucol_off(j) + i] = i < j ? 1.f : 0.f;
        up.valid = 2;
    }
}

struct RingPos {
    uint32_t s, ph;
    __device__ __forceinline__ void next(uint32_t n) { if (++s == n) { s = 0; ph ^= 1; } }
};

}  // namespace fz
}  // namespace dfw
