// Stage 2, tensor-core path: the whole deep MLP of one 128-sample tile in ONE kernel on tcgen05 (5th-gen
// tensor cores), operands fed by TMA, accumulators in TMEM, activations never leaving the SM.
//
// Replaces model/DeepFMs.py:408-436 and :458 (+ the sigmoid of :777) of the reference.
//
//   CTA (448 threads, 1 per SM, persistent over 128-row tiles):
//     warp 0   TMA producer: the tile's bf16 embedding block X (B x F*K) -> A buffer, then every layer's
//              weights W_l (N x K, K-major = nn.Linear layout) as (64-column chunk x <=256-row part) boxes
//              through a 3-stage mbarrier ring; also owns the TMEM allocation
//     warp 1   MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=256, K=16),
//              A and B from 128B-swizzled shared memory, D in TMEM (one column per output neuron, <=512)
//     warps 2-13 epilogue: thread = one sample row x one third of the 32-column blocks (a warp may only touch
//              TMEM lanes 32*(warp%4)..+31); tcgen05.ld 32 columns at a time, + bias, ReLU, -> bf16,
//              written back into the A buffer in the swizzled K-major layout (the next layer's operand);
//              last layer: dot with net_1_fc in registers, + shallow, optional sigmoid -> global
//   Shared memory (227 KiB): A buffer ceil(K/64) x 16 KiB + as many weight stages as fit (4 x 26 KiB for
//   the 390-400-400-400 network).
//   Limits of this fused form: F*K <= 512 and every hidden width <= 512 (TMEM columns).
//
// Numerics: bf16 operands, fp32 accumulate; activations are rounded to bf16 between layers.  This is the
// "stated looser bound" path (5e-4 * max|logit|); the fp32 parity path is mlp_fp32.cu.
#include "tc_common.cuh"

namespace dfw {
namespace tc {

constexpr int TILE_M = 128;
constexpr int A_CHUNK_BYTES = TILE_M * 128;   // 16 KiB
constexpr int MAX_KCHUNKS = 8;                // K <= 512
constexpr int STAGE_ROWS = 256;               // weight rows per stage at most (UMMA N <= 256)
constexpr int MAX_STAGES = 8;
constexpr int CLUSTER = 2;                    // CTAs sharing every weight load by TMA multicast
constexpr int EPI_SPLIT = 3;                  // epilogue warps per TMEM lane quarter (column blocks round-robin)
constexpr int EPI_WARPS = 4 * EPI_SPLIT;
constexpr int NTHREADS = 64 + 32 * EPI_WARPS;
constexpr int MAX_W = 512;
constexpr size_t SMEM_LIMIT = 227 * 1024;
constexpr size_t SMEM_FIXED = 1024 /*align slack*/ + 256 /*barriers*/ + EPI_SPLIT * TILE_M * 4 /*fc-dot partials*/;

struct alignas(64) Maps {
    CUtensorMap x;
    CUtensorMap w[DFW_MAX_DEPTH];
};

struct Params {
    int depth;
    int in_dim;                     // F*K
    int widths[DFW_MAX_DEPTH];      // real widths
    const float* bias[DFW_MAX_DEPTH];
    const float* fc;
    const float* shallow;
    float* logits;
    float* prob;
    long long B;
    int num_tiles;
    int a_chunks;                   // 16 KiB chunks reserved for the activation (A) buffer
    int stage_bytes, nstage;        // weight ring: as many stages as shared memory allows
    int* err;                       // device error word (watchdog), may be NULL
    long long* clk;                 // optional per-CTA timeline (debug tooling): 32 x int64 per CTA
};
#define TC_CLK(slot) do { if (p.clk) p.clk[blockIdx.x * 32 + (slot)] = clock64(); } while (0)

__host__ __device__ inline int nparts(int npad) { return (npad + STAGE_ROWS - 1) / STAGE_ROWS; }
// rows of part p when npad columns are split into nparts(npad) parts of multiples of 16
__host__ __device__ inline int part_rows(int npad, int p) {
    const int np = nparts(npad);
    const int base = pad16((npad + np - 1) / np);
    const int r = npad - p * base;
    return r < base ? r : base;
}
__host__ __device__ inline int part_row0(int npad, int p) { return p * pad16((npad + nparts(npad) - 1) / nparts(npad)); }

// ---------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(NTHREADS, 1)
mlp_tc_kernel(const __grid_constant__ Maps maps, const Params p) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment: the 128B swizzle is a function of shared-memory address bits [4,10)
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sA = base;
    const int NSTAGE = p.nstage, STAGE_BYTES = p.stage_bytes;
    unsigned char* sW = base + (size_t)p.a_chunks * A_CHUNK_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sW + (size_t)NSTAGE * STAGE_BYTES);
    uint64_t* full = bars;                     // [MAX_STAGES] weights landed
    uint64_t* empty = bars + MAX_STAGES;       // [MAX_STAGES] stage consumed by the MMAs
    uint64_t* x_full = bars + 2 * MAX_STAGES;  // X tile landed in the A buffer
    uint64_t* acc_full = x_full + 1;           // [2] accumulators of column part q complete in TMEM
    uint64_t* epi_done = x_full + 3;           // [2] epilogue of column part q finished (A rewritten, TMEM drained)
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(x_full + 5);
    float* sRed = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(bars) + 256);   // [EPI_SPLIT][TILE_M]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = p.depth;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CLUSTER); }   // a stage is free when BOTH CTAs consumed it
        mbar_init(x_full, 1);
        for (int q = 0; q < 2; ++q) { mbar_init(&acc_full[q], 1); mbar_init(&epi_done[q], 32 * EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&maps.x);
        for (int l = 0; l < L; ++l) tma_prefetch_desc(&maps.w[l]);
    }
    if (warp == 0) tmem_alloc(tmem_holder, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // the peer's barriers exist before anything is multicast to it
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    const uint32_t crank = cluster_ctarank();
    // every CTA of a cluster runs the same number of tiles (the weight ring is shared); tiles past the end are dummies
    const int n_iter = (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;

    // per-layer K (input width, padded to 16) and N (output width padded to 16)
    auto layer_k = [&](int l) { return pad16(l == 0 ? p.in_dim : p.widths[l - 1]); };
    auto layer_n = [&](int l) { return pad16(p.widths[l]); };
    // Work order inside a layer is PART-major: all K chunks of output-column part 0, then part 1.  Part 0's
    // accumulators are complete at half time, so its epilogue (TMEM drain + activation write-back) overlaps the MMAs
    // of part 1, and the next layer's first K chunks (which only need part 0's columns) overlap part 1's epilogue.

    if (warp == 0) {
        // ================================================================= TMA producer (one lane)
        if (lane == 0) {
            int stage = 0; uint32_t sphase = 0;
            int parts_per_tile[2] = {0, 0};
            for (int l = 0; l < L; ++l) { parts_per_tile[0] += 1; parts_per_tile[1] += (nparts(layer_n(l)) > 1); }
            int it = 0;
            for (int tile = blockIdx.x; it < n_iter; tile += gridDim.x, ++it) {
                // the A buffer is free once every epilogue of the previous tile has finished
                if (it > 0) {
                    for (int q = 0; q < 2; ++q)
                        if (parts_per_tile[q]) mbar_wait(&epi_done[q], (uint32_t)((it * parts_per_tile[q] - 1) & 1), p.err, 11);
                }
                const int kch0 = (layer_k(0) + KCH - 1) / KCH;
                mbar_expect_tx(x_full, (uint32_t)(kch0 * A_CHUNK_BYTES));
                for (int c = 0; c < kch0; ++c)
                    tma_load_2d(sA + (size_t)c * A_CHUNK_BYTES, &maps.x, x_full, c * KCH, tile * TILE_M);
                for (int l = 0; l < L; ++l) {
                    const int kch = (layer_k(l) + KCH - 1) / KCH, npad = layer_n(l), np = nparts(npad);
                    for (int q = 0; q < np; ++q) {
                        for (int c = 0; c < kch; ++c) {
                            mbar_wait(&empty[stage], sphase ^ 1, p.err, 12);
                            // every part is fetched with the same box (rows of part 0); rows past the tensor end are zero-filled
                            // by TMA and still count towards the transaction bytes
                            // this CTA fetches its half of the box and multicasts it to both CTAs of the cluster
                            const int half = part_rows(npad, 0) / CLUSTER;
                            mbar_expect_tx(&full[stage], (uint32_t)(part_rows(npad, 0) * 128));
                            tma_load_2d_mc(sW + (size_t)stage * STAGE_BYTES + (size_t)crank * half * 128, &maps.w[l], &full[stage],
                                           c * KCH, part_row0(npad, q) + (int)crank * half, (uint16_t)((1u << CLUSTER) - 1));
                            if (++stage == NSTAGE) { stage = 0; sphase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (one lane)
        if (lane == 0) {
            int stage = 0; uint32_t sphase = 0;
            uint32_t epi_cnt[2] = {0, 0};      // epilogue completions of part q that WILL have happened for all layers issued so far
            for (int it = 0; it < n_iter; ++it) {
                int prev_np = 0, prev_npad = 0;
                for (int l = 0; l < L; ++l) {
                    // what the previous layer (or the previous tile's last layer) must have finished, per part
                    const uint32_t target[2] = {epi_cnt[0], epi_cnt[1]};
                    bool waited[2] = {false, false};
                    auto need_prev = [&](int upto) {
                        for (int q2 = 0; q2 <= upto; ++q2)
                            if (!waited[q2] && target[q2] > 0) { mbar_wait(&epi_done[q2], (target[q2] - 1) & 1, p.err, 22 + q2); waited[q2] = true; }
                    };
                    if (l == 0) {
                        if (it == 0) TC_CLK(0);
                        mbar_wait(x_full, (uint32_t)(it & 1), p.err, 21);
                        if (it == 0) TC_CLK(1);
                        need_prev(1);                                   // TMEM fully drained by the previous tile
                    }
                    const int K = layer_k(l), kch = (K + KCH - 1) / KCH, npad = layer_n(l), np = nparts(npad);
                    for (int q = 0; q < np; ++q) {
                        const uint32_t idesc = make_idesc(TILE_M, part_rows(npad, q));
                        const uint32_t dcol = tmem_base + (uint32_t)part_row0(npad, q);
                        for (int c = 0; c < kch; ++c) {
                            if (l > 0) {
                                // A chunk c holds columns [64c, 64c+64) of the previous layer's output; TMEM columns of part q
                                // were last read by the previous layer's epilogue of the parts overlapping them
                                const int last_col = min(c * KCH + KCH, prev_npad) - 1;
                                int need = (prev_np > 1 && last_col >= part_row0(prev_npad, 1)) ? 1 : 0;
                                if (prev_np > 1 && (q > 0 || part_rows(npad, 0) > part_row0(prev_npad, 1))) need = 1;
                                need_prev(need);
                            }
                            const int ksteps = min(4, (K - c * KCH) / 16);
                            const uint64_t adesc0 = make_desc_sw128(smem_u32(sA + (size_t)c * A_CHUNK_BYTES));
                            mbar_wait(&full[stage], sphase, p.err, 24);
                            tc_fence_after();
                            const uint64_t bdesc0 = make_desc_sw128(smem_u32(sW + (size_t)stage * STAGE_BYTES));
                            for (int ks = 0; ks < ksteps; ++ks)      // +32 bytes (16 bf16) along K inside the swizzle atom
                                umma_bf16(dcol, adesc0 + (uint64_t)(ks * 2), bdesc0 + (uint64_t)(ks * 2), idesc, (c | ks) ? 1u : 0u);
                            umma_commit_mc(&empty[stage], (uint16_t)((1u << CLUSTER) - 1));   // this CTA is done with the stage (both CTAs are told)
                            if (++stage == NSTAGE) { stage = 0; sphase ^= 1; }
                        }
                        umma_commit(&acc_full[q]);                     // accumulators of part q of layer l complete
                        ++epi_cnt[q];                                  // its epilogue will complete epi_done[q] once
                        if (it == 0 && l < 4) TC_CLK(2 + 2 * l + q);   // MMAs of (layer l, part q) all issued
                    }
                    prev_np = np; prev_npad = npad;
                }
            }
        }
    } else {
        // ================================================================= epilogue warps (thread = sample row)
        const int q4 = warp & 3;                       // TMEM lane quarter this warp may access
        const int split = (warp - 2) >> 2;             // which third of the 32-column blocks this warp handles
        const int row = q4 * 32 + lane;
        const uint32_t taddr_row = tmem_base + ((uint32_t)(q4 * 32) << 16);
        uint32_t n_acc[2] = {0, 0};
        int eit = 0;
        for (int tile = blockIdx.x; eit < n_iter; tile += gridDim.x, ++eit) {
            const long long b = (long long)tile * TILE_M + row;
            const bool first_tile = tile == (int)blockIdx.x;
            for (int l = 0; l < L; ++l) {
                const int N = p.widths[l], npad = layer_n(l), np = nparts(npad);
                const bool last = (l == L - 1);
                const float* bias = p.bias[l];
                float dot = 0.f;
                auto load16 = [&](const float* src, int n0, float (&o)[16]) {
                    if (n0 + 16 <= N) {
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(src + n0) + h);
                            o[4 * h] = t.x; o[4 * h + 1] = t.y; o[4 * h + 2] = t.z; o[4 * h + 3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) o[i] = (n0 + i < N) ? __ldg(src + n0 + i) : 0.f;
                    }
                };
                auto process = [&](int n0, const uint32_t (&r)[16], const float (&bb)[16], const float (&ff)[16]) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(__uint_as_float(r[i]) + bb[i], 0.f);
                    if (last) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) dot = fmaf(v[i], ff[i], dot);      // ff is 0 past the real width
                    } else {
                        // bf16, swizzled K-major: 16-byte unit u of row `row` lives at unit (u ^ (row & 7))
                        const int c = n0 >> 6, u0 = (n0 & 63) >> 3;
                        unsigned char* rowp = sA + (size_t)c * A_CHUNK_BYTES + (size_t)row * 128;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            uint4 pk;
                            __nv_bfloat162 t0 = __floats2bfloat162_rn(v[8 * h + 0], v[8 * h + 1]);
                            __nv_bfloat162 t1 = __floats2bfloat162_rn(v[8 * h + 2], v[8 * h + 3]);
                            __nv_bfloat162 t2 = __floats2bfloat162_rn(v[8 * h + 4], v[8 * h + 5]);
                            __nv_bfloat162 t3 = __floats2bfloat162_rn(v[8 * h + 6], v[8 * h + 7]);
                            pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
                            pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
                            *reinterpret_cast<uint4*>(rowp + (((u0 + h) ^ (row & 7)) << 4)) = pk;
                        }
                    }
                };
                // The epilogue of a hidden layer rewrites the A buffer in place, so it may only start once EVERY MMA of the
                // layer has finished reading A (the last part's accumulators are complete).  The last layer writes nothing
                // to shared memory: there part 0 is drained while part 1 is still being multiplied.
                if (!last) {
                    for (int q = 0; q < np; ++q) { mbar_wait(&acc_full[q], n_acc[q] & 1, p.err, 31 + q); ++n_acc[q]; }
                }
                for (int q = 0; q < np; ++q) {
                    if (last) { mbar_wait(&acc_full[q], n_acc[q] & 1, p.err, 33 + q); ++n_acc[q]; }
                    if (threadIdx.x == 64 && first_tile && l < 4) TC_CLK(10 + 4 * l + 2 * q);      // accumulators of (l, q) complete
                    tc_fence_after();
                    const int col0 = part_row0(npad, q), col1 = col0 + part_rows(npad, q);
                    uint32_t rA[16], rB[16];
                    float bA[16], bB[16], fA[16], fB[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) { fA[i] = 0.f; fB[i] = 0.f; }
                    for (int n0 = col0 + split * 32; n0 < col1; n0 += 32 * EPI_SPLIT) {
                        const bool hasB = n0 + 16 < col1;
                        tmem_ld16(taddr_row + (uint32_t)n0, rA);
                        if (hasB) tmem_ld16(taddr_row + (uint32_t)(n0 + 16), rB);
                        load16(bias, n0, bA);
                        if (hasB) load16(bias, n0 + 16, bB);
                        if (last) { load16(p.fc, n0, fA); if (hasB) load16(p.fc, n0 + 16, fB); }
                        tmem_ld_wait();
                        process(n0, rA, bA, fA);
                        if (hasB) process(n0 + 16, rB, bB, fB);
                    }
                    if (last && q == np - 1) {
                        // combine the column thirds of each row in a fixed order
                        sRed[split * TILE_M + row] = dot;
                        asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory");
                        if (split == 0 && b < p.B) {
                            float z = 0.f;
#pragma unroll
                            for (int s2 = 0; s2 < EPI_SPLIT; ++s2) z += sRed[s2 * TILE_M + row];
                            z += (p.shallow ? p.shallow[b] : 0.f);
                            if (p.logits) p.logits[b] = z;
                            if (p.prob) p.prob[b] = 1.0f / (1.0f + expf(-z));
                        }
                    }
                    if (!last) fence_async_smem();     // generic-proxy stores -> visible to the tensor-core (async) proxy
                    tc_fence_before();                 // TMEM reads ordered before the arrive
                    if (threadIdx.x == 64 && first_tile && l < 4) TC_CLK(11 + 4 * l + 2 * q);      // this thread's epilogue of (l, q) done
                    mbar_arrive(&epi_done[q]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();            // no CTA exits while its peer may still multicast into it
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// fp32 (out, in) row-major -> bf16 (pad16(out), pad64(in)) row-major, zero padded
__global__ void pack_bf16_kernel(const float* __restrict__ W, int out_dim, int in_dim, int out_pad, int in_pad,
                                 __nv_bfloat16* __restrict__ dst) {
    const long long total = (long long)out_pad * in_pad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / in_pad), c = (int)(i - (long long)r * in_pad);
        dst[i] = __float2bfloat16_rn((r < out_dim && c < in_dim) ? W[(long long)r * in_dim + c] : 0.f);
    }
}

// ---------------------------------------------------------------------------------------- host side
static long long* g_clk = nullptr;

}  // namespace tc
}  // namespace dfw

using namespace dfw;

// Debug tooling (not part of the product ABI): per-CTA clock64() timeline of the next mlp_tc launches.
extern "C" void dfw_debug_set_mlp_clock_buffer(void* dev_buf) { tc::g_clk = static_cast<long long*>(dev_buf); }

extern "C" size_t dfw_pack_mlp_bf16_bytes(int32_t out_dim, int32_t in_dim) {
    return (size_t)tc::pad16(out_dim) * ((in_dim + 63) / 64 * 64) * 2;
}

extern "C" int dfw_pack_mlp_bf16(const float* W, int32_t out_dim, int32_t in_dim, void* dst, void* stream) {
    DFW_REQUIRE(W && dst && out_dim > 0 && in_dim > 0, DFW_E_ARG, "bad pack_mlp_bf16 arguments");
    const int out_pad = tc::pad16(out_dim), in_pad = (in_dim + 63) / 64 * 64;
    tc::pack_bf16_kernel<<<148, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(W, out_dim, in_dim, out_pad, in_pad,
                                                                                 static_cast<__nv_bfloat16*>(dst));
    count_launch();
    return check_launch("pack_bf16_kernel");
}

extern "C" int dfw_mlp_bf16(const dfw_model* m, const void* Xb, int64_t ldXb, int64_t B, const float* shallow,
                            void* workspace, size_t workspace_bytes, float* logits_out, float* prob_out, void* stream) {
    dfw::NvtxRange nvtx_("Deep - Component, tcgen05 (dfw_mlp_bf16)");
    if (int rc = check_model(m)) return rc;
    DFW_REQUIRE(m->flags & DFW_USE_DEEP, DFW_E_ARG, "model has no deep part");
    DFW_REQUIRE(Xb && (logits_out || prob_out), DFW_E_ARG, "X / outputs NULL");
    if (B <= 0) return 0;
    const int in_dim = m->field_size * m->embedding_size;
    DFW_REQUIRE(in_dim <= tc::MAX_W, DFW_E_UNSUPPORTED, "tensor MLP: F*K = %d > %d", in_dim, tc::MAX_W);
    DFW_REQUIRE(ldXb % 8 == 0 && ldXb >= in_dim, DFW_E_ARG, "ldXb must be a multiple of 8 (TMA 16-byte pitch) and >= F*K");
    DFW_REQUIRE((reinterpret_cast<uintptr_t>(Xb) & 15) == 0, DFW_E_ARG, "Xb must be 16-byte aligned");
    tc::Maps maps;
    tc::Params p;
    p.depth = m->depth; p.in_dim = in_dim;
    int k = in_dim;
    for (int l = 0; l < m->depth; ++l) {
        const int n = m->widths[l];
        DFW_REQUIRE(n <= tc::MAX_W, DFW_E_UNSUPPORTED, "tensor MLP: layer width %d > %d", n, tc::MAX_W);
        DFW_REQUIRE(m->Wbf16[l], DFW_E_ARG, "layer %d has no bf16 image (call dfw_pack_mlp_bf16)", l + 1);
        const int npad = tc::pad16(n), kpad = (k + 63) / 64 * 64;
        // all parts of a layer share one map whose box holds the rows of part 0 (the largest)
        if (int rc = tc::make_map(&maps.w[l], m->Wbf16[l], npad, kpad, kpad, tc::part_rows(npad, 0) / tc::CLUSTER)) return rc;
        p.widths[l] = n; p.bias[l] = m->b[l];
        k = n;
    }
    if (int rc = tc::make_map(&maps.x, Xb, B, ldXb, ldXb, tc::TILE_M)) return rc;
    p.fc = m->fc; p.shallow = shallow; p.logits = logits_out; p.prob = prob_out; p.B = B;
    p.num_tiles = (int)((B + tc::TILE_M - 1) / tc::TILE_M);
    // shared-memory plan: activation buffer for the widest operand, then as many weight stages as fit
    {
        int kmax = in_dim, rows_max = 16;
        for (int l = 0; l < m->depth; ++l) {
            if (l + 1 < m->depth && m->widths[l] > kmax) kmax = m->widths[l];
            const int pr = tc::part_rows(tc::pad16(m->widths[l]), 0);
            if (pr > rows_max) rows_max = pr;
        }
        p.a_chunks = (tc::pad16(kmax) + tc::KCH - 1) / tc::KCH;
        p.stage_bytes = (rows_max * 128 + 1023) / 1024 * 1024;
        const size_t avail = tc::SMEM_LIMIT - tc::SMEM_FIXED - (size_t)p.a_chunks * tc::A_CHUNK_BYTES;
        p.nstage = (int)(avail / p.stage_bytes);
        if (p.nstage > tc::MAX_STAGES) p.nstage = tc::MAX_STAGES;
        DFW_REQUIRE(p.nstage >= 2, DFW_E_UNSUPPORTED, "tensor MLP: shapes leave room for %d weight stage(s)", p.nstage);
    }
    const size_t smem_bytes = tc::SMEM_FIXED + (size_t)p.a_chunks * tc::A_CHUNK_BYTES + (size_t)p.nstage * p.stage_bytes;
    p.err = (workspace && workspace_bytes >= 4) ? static_cast<int*>(workspace) : nullptr;
    p.clk = tc::g_clk;
    static thread_local bool configured_dev[16] = {};       // per device: function attributes are per context
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool& configured = configured_dev[cur_dev & 15];
    if (!configured) {
        DFW_CUDA_OK(cudaFuncSetAttribute(tc::mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_LIMIT));
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = p.num_tiles < sms ? p.num_tiles : sms;
    grid = (grid + tc::CLUSTER - 1) / tc::CLUSTER * tc::CLUSTER;
    if (grid > sms) grid -= tc::CLUSTER;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(tc::NTHREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = reinterpret_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = tc::CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    DFW_CUDA_OK(cudaLaunchKernelEx(&cfg, tc::mlp_tc_kernel, maps, p));
    count_launch();
    return check_launch("mlp_tc_kernel");
}
