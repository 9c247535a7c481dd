// Row pull for rank-sharded tables: the exchange step of the sharded forward as its own small kernel.
//
// With the tables row-sharded over P GPUs (row i on rank i mod P, local row i div P) the fused kernel can fetch peer rows itself
// (exchange "p2p": one launch, lowest latency), but a CTA that waits ~10 us for NVLink round trips holds its SM -- 213 KB of shared
// memory, the tensor core -- idle meanwhile.  This kernel does the fetching instead, one batch AHEAD of the fused kernel: small
// CTAs (128 threads, <= 32 registers: they fit beside a resident fused CTA) read every sharded field's row of every sample from
// the owning GPU -- direct peer loads over NVLink / NVSwitch through the CUDA-IPC mapped shard pointers, cp.async into shared
// memory and out again with coalesced stores -- into a batch-ordered staging buffer in local HBM, and rewrite the index columns so that the fused kernel
// reads the staging buffer as a B-row table.  Rows are copied, never combined: results are bit-identical to one GPU.
//
//   staged  (n_sharded, B, K) fp32       row of sample b for the j-th sharded field (quotient row for a QR table)
//   xi2     (B, C) same dtype as xi      copy of xi; sharded columns become  b * c + idx mod c   (c = QR collisions, 1 for plain)
#include <stdlib.h>
#include <string.h>

#include "embed_device.cuh"

namespace dfw {
namespace pl {

constexpr int THREADS = 128;

struct PullParams {
    const dfw_field_desc* fields;        // DEVICE array [F] with the shard pointers
    const void* xi; int64_t xi_sb, xi_sc;
    float* staged; void* xi2;
    int32_t* err;
    int32_t B, C, K, num, n_sf, xi32, check;
    int8_t col_sf[DFW_MAX_FIELDS];       // categorical column -> index among the sharded fields, or -1
    int8_t sf_col[DFW_MAX_FIELDS];       // sharded field j -> categorical column
};

__device__ __forceinline__ int64_t load_idx(const PullParams& p, int b, int col) {
    const int64_t e = (int64_t)b * p.xi_sb + (int64_t)col * p.xi_sc;
    return p.xi32 ? (int64_t)__ldg(static_cast<const int32_t*>(p.xi) + e) : __ldg(static_cast<const int64_t*>(p.xi) + e);
}
__device__ __forceinline__ void store_idx(const PullParams& p, int b, int col, int64_t v) {
    const int64_t e = (int64_t)b * p.C + col;
    if (p.xi32) static_cast<int32_t*>(p.xi2)[e] = (int32_t)v; else static_cast<int64_t*>(p.xi2)[e] = v;
}

__device__ __forceinline__ void cp_async_f(float* smem_dst, const float* gsrc, int bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if (bytes == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}

// One CTA = chunks of ROWS rows.  Rows are numbered r = j * B + b (sharded field j, sample b), which is also their position in
// the staging buffer, so a chunk lands in ROWS * K contiguous floats.  Every thread starts the copies of its rows with cp.async
// (global / peer memory -> shared memory: the loads in flight cost no registers, which is what lets 128 threads keep 1280 of them
// outstanding beside a fused CTA that owns 94 % of the register file), then the chunk is written out with coalesced stores.
template <int SEGW, int ROWS>      // SEGW floats per piece: 2 when K is even (all row bases 8-byte aligned), else 1
__global__ void __launch_bounds__(THREADS, 16) pull_rows_kernel(const PullParams p) {
    extern __shared__ __align__(16) float sbuf[];          // ROWS x K
    const uint32_t total_rows = (uint32_t)p.B * (uint32_t)p.n_sf;
    const uint32_t nV = (uint32_t)p.K / SEGW;
    // pass-through index columns
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < (uint32_t)p.B * (uint32_t)p.C; e += gridDim.x * blockDim.x) {
        const uint32_t b = e / (uint32_t)p.C, col = e - b * (uint32_t)p.C;
        if (p.col_sf[col] < 0) store_idx(p, (int)b, (int)col, load_idx(p, (int)b, (int)col));
    }
    for (uint32_t r0 = blockIdx.x * ROWS; r0 < total_rows; r0 += gridDim.x * ROWS) {
#pragma unroll 1
        for (uint32_t t = threadIdx.x; t < ROWS; t += THREADS) {
            const uint32_t r = r0 + t;
            if (r >= total_rows) break;
            const uint32_t j = r / (uint32_t)p.B, b = r - j * (uint32_t)p.B;
            const int col = p.sf_col[j];
            const dfw_field_desc& fd = p.fields[p.num + col];
            int64_t idx = load_idx(p, (int)b, col);
            if (idx < 0 || idx >= fd.rows) {                 // same defined behaviour as the gather: error word, row 0
                if (p.check && p.err) atomicExch(p.err, 1 + p.num + col);
                idx = 0;
            }
            const uint32_t c = (uint32_t)fd.collisions;
            const uint32_t q = div_small((uint32_t)idx, c);      // stored row: the quotient for a QR table
            const uint32_t P = (uint32_t)fd.n_ranks;
            const uint32_t local = div_small(q, P);
            const float* src = fd.w2_shard[q - local * P] + (size_t)local * p.K;
            float* dst = sbuf + t * p.K;
            for (uint32_t v = 0; v < nV; ++v) cp_async_f(dst + v * SEGW, src + v * SEGW, 4 * SEGW);
            store_idx(p, (int)b, col, (int64_t)b * c + ((uint32_t)idx - q * c));
        }
        asm volatile("cp.async.wait_all;\n" ::: "memory");
        __syncthreads();
        const uint32_t nrows = min((uint32_t)ROWS, total_rows - r0);
        const uint32_t nfl = nrows * (uint32_t)p.K;
        float* out = p.staged + (size_t)r0 * p.K;
        if (((reinterpret_cast<uintptr_t>(out) | (uintptr_t)(nfl * 4)) & 15) == 0) {
            for (uint32_t i = threadIdx.x; i < nfl / 4; i += THREADS)
                reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(sbuf)[i];
        } else {
            for (uint32_t i = threadIdx.x; i < nfl; i += THREADS) out[i] = sbuf[i];
        }
        __syncthreads();
    }
}

}  // namespace pl
}  // namespace dfw

using namespace dfw;

extern "C" int dfw_pull_rows(const dfw_model* m, const int32_t* sharded_fields, int32_t n_sharded, const int64_t* xi,
                             int64_t xi_stride_b, int64_t xi_stride_c, int64_t B, float* staged_out, void* xi2_out,
                             int32_t* err_word, void* stream) {
    dfw::NvtxRange nvtx_("sharded row exchange (dfw_pull_rows)");
    if (int rc = check_model(m)) return rc;
    DFW_REQUIRE(B >= 0 && B < (1ll << 24), DFW_E_ARG, "batch %lld outside [0, 2^24)", (long long)B);
    if (B == 0) return 0;
    const int C = m->field_size - m->numerical;
    DFW_REQUIRE(sharded_fields && n_sharded >= 1 && n_sharded <= C, DFW_E_ARG, "n_sharded %d outside [1, %d]", n_sharded, C);
    DFW_REQUIRE(xi && staged_out && xi2_out, DFW_E_ARG, "xi / staged_out / xi2_out is NULL");
    pl::PullParams p;
    memset(&p, 0, sizeof(p));
    for (int c = 0; c < DFW_MAX_FIELDS; ++c) p.col_sf[c] = -1;
    for (int j = 0; j < n_sharded; ++j) {
        const int f = sharded_fields[j];
        DFW_REQUIRE(f >= m->numerical && f < m->field_size && p.col_sf[f - m->numerical] < 0, DFW_E_ARG,
                    "sharded field %d is not a distinct categorical field", f);
        p.col_sf[f - m->numerical] = (int8_t)j;
        p.sf_col[j] = (int8_t)(f - m->numerical);
    }
    p.fields = m->fields;
    p.xi = xi; p.xi_sb = xi_stride_b; p.xi_sc = xi_stride_c;
    p.staged = staged_out; p.xi2 = xi2_out; p.err = err_word;
    p.B = (int32_t)B; p.C = C; p.K = m->embedding_size; p.num = m->numerical; p.n_sf = n_sharded;
    p.xi32 = (m->flags & DFW_XI_INT32) ? 1 : 0;
    p.check = (m->flags & DFW_CHECK_INDEX) ? 1 : 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    constexpr int ROWS = 256;
    DFW_REQUIRE((size_t)ROWS * p.K * sizeof(float) <= 40 * 1024, DFW_E_UNSUPPORTED, "embedding_size %d too wide for the pull kernel", p.K);
    const bool vec2 = p.K % 2 == 0;
    const long long chunks = ((long long)B * n_sharded + ROWS - 1) / ROWS;
    // at most one CTA per SM: two of them on an SM would keep the next fused CTA (94 % of the registers, 214 KB of shared
    // memory) from becoming resident there until they finish
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static const int cap = dbg_getenv("DFW_PULL_ROWS_CTAS") ? atoi(dbg_getenv("DFW_PULL_ROWS_CTAS")) : 0;
    if (cap > 0) sms = cap;
    const int grid = (int)(chunks > sms ? sms : chunks);
    const size_t smem = (size_t)ROWS * p.K * sizeof(float);
    if (vec2) pl::pull_rows_kernel<2, ROWS><<<grid, pl::THREADS, smem, st>>>(p);
    else pl::pull_rows_kernel<1, ROWS><<<grid, pl::THREADS, smem, st>>>(p);
    count_launch();
    return check_launch("pull_rows_kernel");
}
