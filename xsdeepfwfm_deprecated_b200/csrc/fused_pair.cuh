// The fused forward on a CTA PAIR (tcgen05 cta_group::2): two CTAs of a cluster share every MMA.
//
// Why: with one CTA per 32-sample tile (fused_tc.cu) every SM streams ALL MLP weights through its shared memory -- each
// weight byte is written once by TMA and read once by the tensor core -- and at B = 4096 that shared-memory traffic, not the
// tensor pipe, bounds the MLP (measured: ~128 B/clk/SM).  A cta_group::2 MMA computes D (256 neurons x 64 samples) from
// A halves that live in the two CTAs' shared memories (128 weight rows each) and B halves (32 samples each, scripts/ubench/
// pair_mma.cu checks the layout: D_r[i][j] = A_r[i] . [B_0 ; B_1][j]).  So each CTA streams only HALF of the weights for its
// 32 samples, and the pair issues half as many (twice as large) MMAs.
//
//   pair = CTA 0 (leader) + CTA 1; CTA r gathers the samples of tile 2p + r and owns neuron tiles t = r (mod 2) of every layer
//   warps 0-1   TMA producers (both CTAs), one per ring: ring j carries pair-tile j = neuron tiles (2j, 2j+1); each CTA loads
//               its own tile's 128 x 64 box into its own stage s and signals the LEADER's full[j][s] barrier
//               (cp.async.bulk.tensor ... cta_group::2), which expects both boxes
//   warps 2-3   MMA issuers (leader only), one per ring: tcgen05.mma.cta_group::2, M = 256, N = 64, K = 16, K-outer order;
//               tcgen05.commit ... multicast releases the stage in both CTAs and signals both CTAs' acc_full
//   epilogue    four sets of four warps (4-7, 8-11, 12-15, 16-19; warp % 4 = TMEM lane quarter, thread = neuron of this CTA's tile):
//               set q takes sample columns 16 q .. 16 q + 15 (the samples of CTA q / 2) of pair-tile 0, then of pair-tile 1, so
//               the next layer's first operand chunks are ready after half of the epilogue work.  The bf16 (hi | lo) activations
//               go straight into the activation buffer of the CTA that owns the samples as 16-byte stores (sample-contiguous
//               operand, see X_SBO below) -- st.shared::cluster to the peer -- and the leader's act_ready barrier of the tile
//               gets one arrive per warp (remote from CTA 1).  Sets B-D are the gather warps once their interaction is done.
//   warps 8-17  gather group: as in fused_tc.cu (register path); x_ready is the leader's barrier, 10 arrives from each CTA
//   last layer  every warp reduces its 32 neurons x 16 samples to per-sample sums and stores them into the owner CTA's
//               `red` array; the owner adds them to its shallow part
//
//   bf16x3: three N = 64 MMAs per k-step into the same 64 accumulator columns -- W_hi X_hi, W_hi X_lo (B descriptor + X_HB),
//   W_lo X_hi.  One N = 128 MMA on [X_hi | X_lo] would read the weights once for two products (measured: -15 % MMA time), but the
//   third product then lands on different accumulator columns for the two CTAs' samples, i.e. a sample's rounding would depend on
//   its position in the batch; the tests pin batch-order equivariance, so it is not used.
//   One wave only (every tile owned by exactly one CTA of the grid); larger batches use fused_tc.cu.
#pragma once
#include "fused_common.cuh"

namespace dfw {
namespace fz {

struct PairBars {
    uint64_t full[2][RING_MAX];      // leader: both CTAs' boxes of stage s have landed
    uint64_t empty[2][RING_MAX];     // per CTA: stage s consumed (multicast commit of the leader)
    uint64_t x_ready;                // leader: layer-1 operand written in both CTAs (2 x G_WARPS arrives)
    uint64_t shallow_ready;          // per CTA
    uint64_t act_ready[2][MAX_MT];   // leader: [layer parity][neuron tile] 16 arrives (4 sets x 4 warps) from the CTA that owns the tile
    uint64_t acc_full[2][2];         // per CTA: [layer parity][pair-tile] accumulators complete (multicast commit)
    uint64_t fin;                    // per CTA: the 8 partial-sum blocks of its samples are in `red`
    uint32_t tmem_holder, pad_;
    float shallow[TS];
    float red[2][2][EPI_WARPS][TS];  // [source CTA][pair-tile][lane quarter][sample]
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Arrive on a barrier of another CTA of the cluster with CTA-scope release (the form CUTLASS's ClusterBarrier::arrive(cta_id) uses).
// For a hand-off to the ASYNC proxy (tcgen05.mma operands): the writer's fence.proxy.async has already waited for its stores --
// measured: it returns ~130 cycles after local stores and only after 1.3-1.8 k cycles when 8 KB of st.shared::cluster are still
// draining -- so the cluster-scope release of mbar_arrive_cluster (MEMBAR.ALL.GPU + ERRBAR: 0.8-1.9 k cycles per arrive even with
// nothing outstanding) adds latency to the hand-off and no ordering the consumer needs.
__device__ __forceinline__ void mbar_arrive_remote_cta(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cta.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void st_cluster_b16(uint32_t cluster_addr, __nv_bfloat16 v) {
    asm volatile("st.shared::cluster.b16 [%0], %1;" ::"r"(cluster_addr), "h"(*reinterpret_cast<const uint16_t*>(&v)) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 lo, __nv_bfloat16 hi) {
    return (uint32_t)*reinterpret_cast<const uint16_t*>(&lo) | ((uint32_t)*reinterpret_cast<const uint16_t*>(&hi) << 16);
}
__device__ __forceinline__ void st_cluster_f32(uint32_t cluster_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}
// wait with cluster-scope acquire (the producers of these barriers may be threads of the peer CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int* err, int code) {
    const uint32_t addr = smem_u32(bar);
    unsigned long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
        if ((spin & 0x3ffu) == 0x3ffu && watchdog_expired(t0)) {     // only a broken pipeline gets here
            if (err) { atomicExch(err, code); atomicOr(err + 1, 1 << (code & 31)); }      // err[1]: set of waits that timed out (debug word)
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d_2cta(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// The activation operand (B of every MMA) is stored SAMPLE-contiguous ("MN-major", no swizzle; layout and descriptor fields
// checked by scripts/ubench/pair_mma_mn.cu): element (k, sample s) of a 64-wide K chunk lives at
//     (s / 8) * X_SBO + k * 16 + (s % 8) * 2          core matrix = 8 k-rows x 8 samples = 128 contiguous bytes
// so the epilogue thread of neuron k writes its 32 samples as four 16-byte stores (a warp: 512 contiguous bytes) instead of 32
// scattered 2-byte stores into a K-major row.  X_SBO = 1024 + 32 keeps the gather group's 2-byte stores (a warp = 16 samples x
// 2 adjacent k) conflict-free.  Descriptor: LBO = 128 (next 8 k), SBO = X_SBO (next 8 samples), K = 16 step = 256 bytes.
constexpr uint32_t X_SBO = 1056;
constexpr uint32_t X_HB = 4 * X_SBO;           // hi (or lo) part of one chunk: 32 samples x 64 k
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(128 >> 4) << 16;
    d |= (uint64_t)(X_SBO >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// 20 warps: 0-1 producers, 2-3 MMA issuers, 4-7 epilogue set A, 8-17 gather group (8-11 = set B, 12-15 = set C once their tile's
// interaction is done), 16-19 = set D (18-19 do nothing else).  Set (j, h) = (pair-tile, sample half): A (0,0) B (0,1) C (1,0) D (1,1),
// so the four (tile, half) units of a layer's epilogue run side by side.
constexpr int PAIR_WARPS = G_WARP0 + G_WARPS + 2;
constexpr int PAIR_THREADS = 32 * PAIR_WARPS;      // 640 (5 warps per scheduler: 96 registers per thread)
constexpr int PAIR_CORE_THREADS = 32 * (G_WARP0 + 2);

template <bool SPLIT, int FT, int KT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)
fused_pair_kernel(const __grid_constant__ Maps maps, const __grid_constant__ UParam up, const Params p) {
    static_assert(FT > 0 && KT > 0 && KT <= G_WARPS, "the pair kernel is built for the specialised shapes only");
    constexpr int H = SPLIT ? 2 : 1;
    constexpr int CH = H * (int)X_HB;                   // one K chunk of the activation buffer: hi [| lo] of 32 samples
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sX = base;
    unsigned char* sW = base + p.oRing;
    PairBars* bars = reinterpret_cast<PairBars*>(base + p.oMisc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = p.depth;
    const uint32_t NS0 = (uint32_t)p.nst[0], NS1 = (uint32_t)p.nst[1];
    const uint32_t rank = cluster_ctarank();            // 0 = leader
    const bool leader = rank == 0;
    if (p.clk && threadIdx.x == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.clk[blockIdx.x * FZ_NCLK + 28] = clock64(); p.clk[blockIdx.x * FZ_NCLK + 29] = (long long)t;
    }

    if (threadIdx.x == 0) {
        for (int r = 0; r < 2; ++r)
            for (int s = 0; s < RING_MAX; ++s) { mbar_init(&bars->full[r][s], 1); mbar_init(&bars->empty[r][s], 1); }
        mbar_init(&bars->x_ready, 2 * G_WARPS);
        mbar_init(&bars->shallow_ready, G_WARPS);
        mbar_init(&bars->fin, 4 * EPI_WARPS);        // sets (0,h) and (1,h) of both CTAs
        for (int b = 0; b < 2; ++b) {
            for (int m = 0; m < MAX_MT; ++m) mbar_init(&bars->act_ready[b][m], 4 * EPI_WARPS);
            for (int j = 0; j < 2; ++j) mbar_init(&bars->acc_full[b][j], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int l = 0; l < L; ++l)
            for (int h = 0; h < H; ++h) tma_prefetch_desc(&maps.w[l][h][0]);
    }
    const bool gather_warp = warp >= G_WARP0 && warp < G_WARP0 + G_WARPS;
    uint32_t tmem_base = 0;
    if (gather_warp) {
        cluster_arrive();
    } else {
        if (warp == MMA_WARP0) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_holder)), "r"(256) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        asm volatile("bar.sync %0, %1;" ::"n"(BAR_CORE), "n"(PAIR_CORE_THREADS) : "memory");
        cluster_arrive(); cluster_wait();        // both CTAs' barriers and TMEM exist before anything crosses over
        tc_fence_after();
        tmem_base = bars->tmem_holder;
        asm volatile("bar.arrive %0, %1;" ::"n"(BAR_INIT), "n"(PAIR_THREADS) : "memory");
    }
    const int tile = (int)blockIdx.x;                   // one wave: CTA b owns tile b

    auto layer_k = [&](int l) { return pad16(l == 0 ? p.in_dim : p.widths[l - 1]); };
    auto layer_n = [&](int l) { return pad16(p.widths[l]); };

    // ---------------------------------------------------------------- epilogue: set q = sample columns [16 q, 16 q + 16)
    // The four sets (4 warps each, warp % 4 = TMEM lane quarter) split the 64 sample columns of a pair-tile four ways and walk
    // this CTA's neuron tiles IN ORDER: tile j = 0 of every set first, then j = 1.  The operand chunks of the first tiles are
    // therefore complete -- and the next layer's MMAs start -- when half of the layer's epilogue work is done; the second half
    // runs under those MMAs.  Columns 16 q .. belong to the samples of CTA h = q / 2: activations and partial sums go to CTA h.
    auto epilogue_pass = [&](int q, uint32_t tbase) {
        const int q4 = warp & 3;
        const int row = q4 * 32 + lane;
        const int half = q >> 1, sub = q & 1;
        const uint32_t taddr_row = tbase + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(16 * q);
        const uint32_t xdst = mapa_u32(smem_u32(sX), (uint32_t)half) + (uint32_t)(2 * sub) * X_SBO;   // n-groups 2 sub, 2 sub + 1
        uint32_t acc_bits = 0;
        float zsum[2] = {0.f, 0.f};                    // per pair-tile: this warp's neurons' relu(.) * fc, sample 16 sub + lane / 2 of CTA `half`
        for (int l = 0; l < L; ++l) {
            const int buf = l & 1, N = p.widths[l], npad = layer_n(l), MT = n_mtiles(npad), PT = (MT + 1) / 2;
            const bool last = (l == L - 1);
            // bias / fc of this thread's neuron in both tiles: loaded before the wait, off the critical path
            float bb2[2], ff2[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = (2 * j + (int)rank) * 128 + row;
                const bool real = n < npad && n < N;
                bb2[j] = real ? __ldg(p.bias[l] + n) : 0.f;
                ff2[j] = (real && last) ? __ldg(p.fc + n) : 0.f;
            }
            // the tile outputs overwrite the activation buffer the layer still reads: wait for every pair-tile of the layer
            for (int w = 0; w < PT; ++w) {
                const int bit = buf * 2 + w;
                FZ_PROG(8 + 4 * q + q4, (l << 16) | (w << 4) | (6 << 24));
                mbar_wait(&bars->acc_full[buf][w], (acc_bits >> bit) & 1u, p.err, 31);
                acc_bits ^= 1u << bit;
            }
            if (threadIdx.x == 32 * EPI_WARP0 && l < 4) FZ_CLK(8 + 2 * l);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j >= PT) break;
                const int t = 2 * j + (int)rank;                    // this CTA's neuron tile of pair-tile j
                const int n = t * 128 + row;
                const int rows_valid = max(0, min(128, npad - t * 128));
                const float bb = bb2[j], ff = ff2[j];
                FZ_PROG(8 + 4 * q + q4, (l << 16) | (j << 4) | (7 << 24));
                const bool stamp = p.clk && threadIdx.x == 32 * EPI_WARP0 && l == 0;
                if (stamp) FZ_CLK(40 + 4 * j);
                if (q4 * 32 < rows_valid) {
                    uint32_t d[16];
                    tmem_ld16(taddr_row + (uint32_t)(buf * 128 + j * 64), d);
                    tmem_ld_wait();
                    if (stamp) FZ_CLK(41 + 4 * j);
                    if (last) {
                        float v[16];
#pragma unroll
                        for (int s = 0; s < 16; ++s) v[s] = fmaxf(__uint_as_float(d[s]) + bb, 0.f) * ff;
                        // transpose-reduce over the warp's 32 neurons: after the four steps lane pair (2 s, 2 s + 1) holds sample s
#pragma unroll
                        for (int off = 16, nn = 16; off >= 2; off >>= 1, nn >>= 1) {
                            const bool upper = (lane & off) != 0;
#pragma unroll
                            for (int i = 0; i < nn / 2; ++i) {
                                const float send = upper ? v[i] : v[i + nn / 2];
                                const float keep = upper ? v[i + nn / 2] : v[i];
                                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                            }
                        }
                        zsum[j] += v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
                    } else if (row < rows_valid) {
                        // 16 samples of k = n in the next operand of CTA `half`: chunk n / 64, two groups of 8 samples
                        const uint32_t xc = xdst + (uint32_t)((n >> 6) * CH + (n & 63) * 16);
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            uint32_t wh[4], wl[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float a0 = fmaxf(__uint_as_float(d[8 * g + 2 * i]) + bb, 0.f);
                                const float a1 = fmaxf(__uint_as_float(d[8 * g + 2 * i + 1]) + bb, 0.f);
                                const __nv_bfloat162 h2 = __floats2bfloat162_rn(a0, a1);        // .x = a0 (low half)
                                wh[i] = *reinterpret_cast<const uint32_t*>(&h2);
                                if constexpr (SPLIT) {
                                    const __nv_bfloat162 l2 = __floats2bfloat162_rn(a0 - __uint_as_float(wh[i] << 16),
                                                                                      a1 - __uint_as_float(wh[i] & 0xffff0000u));
                                    wl[i] = *reinterpret_cast<const uint32_t*>(&l2);
                                }
                            }
                            st_cluster_v4(xc + g * X_SBO, wh[0], wh[1], wh[2], wh[3]);
                            if constexpr (SPLIT) st_cluster_v4(xc + g * X_SBO + X_HB, wl[0], wl[1], wl[2], wl[3]);
                        }
                    }
                }
                if (stamp) FZ_CLK(42 + 4 * j);
                if (!last && t * 128 < npad) {
                    // generic-proxy stores -> async proxy (the tensor core of the CTA that holds them); the cluster-scope form costs
                    // ~1.2 k cycles, the CTA-scope form ~50, so only the sets that stored into the peer pay for it
                    if (half == (int)rank) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    else asm volatile("fence.proxy.async.shared::cluster;" ::: "memory");
                    if (stamp) FZ_CLK(43 + 4 * j);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->act_ready[(l + 1) & 1][t]), 0));
                }
            }
            if (threadIdx.x == 32 * EPI_WARP0 && l < 4) FZ_CLK(9 + 2 * l);
        }
        // lane pair (2 s, 2 s + 1) holds this warp's partial sums for sample 16 sub + s of CTA `half`
        if ((lane & 1) == 0) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
                st_cluster_f32(mapa_u32(smem_u32(&bars->red[rank][j][q4][16 * sub + (lane >> 1)]), (uint32_t)half), zsum[j]);
        }
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->fin), (uint32_t)half));
    };

    if (warp < RINGS) {
        // ================================================================= TMA producers (both CTAs): ring pw = pair-tile pw
        const int pw = warp;
        const uint32_t NS = pw ? NS1 : NS0;
        unsigned char* ring = sW + (pw ? (size_t)NS0 * STAGE_BYTES : 0);
        RingPos rp{0, 0};
        for (int l = 0; l < L; ++l) {
            const int kch = (layer_k(l) + KCH - 1) / KCH, npad = layer_n(l), MT = n_mtiles(npad), PT = (MT + 1) / 2;
            if (pw >= PT) continue;
            const int row0 = (2 * pw + (int)rank) * 128;            // this CTA's tile (rows past the matrix are zero-filled)
            for (int c = 0; c < kch; ++c) {
                RingPos s0 = rp, s1 = rp;
                if (SPLIT) s1.next(NS);
                const bool k0 = mbar_try(&bars->empty[pw][s0.s], s0.ph ^ 1);
                const bool k1 = SPLIT ? mbar_try(&bars->empty[pw][s1.s], s1.ph ^ 1) : true;
                FZ_PROG(pw, (l << 16) | (c << 8) | (1 << 24));
                if (!k0) mbar_wait(&bars->empty[pw][s0.s], s0.ph ^ 1, p.err, 12);
                if (!k1) mbar_wait(&bars->empty[pw][s1.s], s1.ph ^ 1, p.err, 13);
                FZ_PROG(pw, (l << 16) | (c << 8) | (2 << 24));
                const uint32_t f0 = mapa_u32(smem_u32(&bars->full[pw][s0.s]), 0), f1 = mapa_u32(smem_u32(&bars->full[pw][s1.s]), 0);
                if (elect_one()) {
                    if (leader) mbar_expect_tx(&bars->full[pw][s0.s], 2u * STAGE_BYTES);     // both CTAs' boxes
                    tma_load_2d_2cta(ring + (size_t)s0.s * STAGE_BYTES, &maps.w[l][0][0], f0, c * KCH, row0);
                    if (SPLIT) {
                        if (leader) mbar_expect_tx(&bars->full[pw][s1.s], 2u * STAGE_BYTES);
                        tma_load_2d_2cta(ring + (size_t)s1.s * STAGE_BYTES, &maps.w[l][1][0], f1, c * KCH, row0);
                    }
                }
                __syncwarp();
                rp = s1; rp.next(NS);
            }
        }
    } else if (warp < EPI_WARP0) {
        // ================================================================= MMA issuers (leader only): ring mw = pair-tile mw
        const int mw = warp - MMA_WARP0;
        if (leader) {
            const uint32_t NS = mw ? NS1 : NS0;
            uint64_t* full = bars->full[mw];
            uint64_t* empty = bars->empty[mw];
            const uint32_t sW_u32 = smem_u32(sW) + (mw ? NS0 * (uint32_t)STAGE_BYTES : 0u);
            const uint32_t sX_u32 = smem_u32(sX);
            const uint32_t idesc = make_idesc(256, 64) | (1u << 16);     // B is MN-major
            RingPos rp{0, 0};
            uint32_t act_bits = 0;
            for (int l = 0; l < L; ++l) {
                const int buf = l & 1;
                const int K = layer_k(l), kch = (K + KCH - 1) / KCH, npad = layer_n(l), MT = n_mtiles(npad), PT = (MT + 1) / 2;
                const bool active = mw < PT;
                const uint32_t dcol = tmem_base + (uint32_t)(buf * 128 + mw * 64);
                for (int c = 0; c < kch; ++c) {
                    if (l == 0) {
                        if (c == 0) {
                            if (lane == 0 && mw == 0) FZ_CLK(0);
                            mbar_wait_cluster(&bars->x_ready, 0, p.err, 21);
                            if (lane == 0 && mw == 0) FZ_CLK(1);
                        }
                    } else if ((c & 1) == 0) {
                        const int g = c >> 1, bit = buf * MAX_MT + g;
                        FZ_PROG(2 + mw, (l << 16) | (c << 8) | (5 << 24));
                        mbar_wait_cluster(&bars->act_ready[buf][g], (act_bits >> bit) & 1u, p.err, 22);
                        act_bits ^= 1u << bit;
                    }
                    if (!active) continue;
                    const int ks = min(4, (K - c * KCH) / 16);
                    const uint64_t bhi = make_desc_mn(sX_u32 + (uint32_t)(c * CH));
                    const uint64_t blo = make_desc_mn(sX_u32 + (uint32_t)(c * CH) + X_HB);
                    const uint32_t acc0 = c ? 1u : 0u;
                    const bool last_c = c == kch - 1;
                    RingPos s0 = rp, s1 = rp;
                    if (SPLIT) s1.next(NS);
                    const bool k0 = mbar_try(&full[s0.s], s0.ph), k1 = SPLIT ? mbar_try(&full[s1.s], s1.ph) : true;
                    FZ_PROG(2 + mw, (l << 16) | (c << 8) | (3 << 24));
                    if (!k0) mbar_wait_cluster(&full[s0.s], s0.ph, p.err, 24);
                    if (!k1) mbar_wait_cluster(&full[s1.s], s1.ph, p.err, 25);
                    tc_fence_after();
                    const uint64_t ah = make_desc_sw128(sW_u32 + s0.s * (uint32_t)STAGE_BYTES);
                    const uint64_t al = make_desc_sw128(sW_u32 + s1.s * (uint32_t)STAGE_BYTES);
                    if (elect_one()) {
                        umma_bf16_2cta(dcol, ah, bhi, idesc, acc0);
                        if (ks > 1) umma_bf16_2cta(dcol, ah + 2, bhi + 16, idesc, 1u);
                        if (ks > 2) umma_bf16_2cta(dcol, ah + 4, bhi + 32, idesc, 1u);
                        if (ks > 3) umma_bf16_2cta(dcol, ah + 6, bhi + 48, idesc, 1u);
                        if (SPLIT) {
                            umma_bf16_2cta(dcol, ah, blo, idesc, 1u);                 // W_hi X_lo
                            if (ks > 1) umma_bf16_2cta(dcol, ah + 2, blo + 16, idesc, 1u);
                            if (ks > 2) umma_bf16_2cta(dcol, ah + 4, blo + 32, idesc, 1u);
                            if (ks > 3) umma_bf16_2cta(dcol, ah + 6, blo + 48, idesc, 1u);
                        }
                        umma_commit_2cta(&empty[s0.s]);
                        if (SPLIT) {
                            umma_bf16_2cta(dcol, al, bhi, idesc, 1u);                 // W_lo X_hi
                            if (ks > 1) umma_bf16_2cta(dcol, al + 2, bhi + 16, idesc, 1u);
                            if (ks > 2) umma_bf16_2cta(dcol, al + 4, bhi + 32, idesc, 1u);
                            if (ks > 3) umma_bf16_2cta(dcol, al + 6, bhi + 48, idesc, 1u);
                            umma_commit_2cta(&empty[s1.s]);
                        }
                        if (last_c) umma_commit_2cta(&bars->acc_full[buf][mw]);
                    }
                    __syncwarp();
                    rp = s1; rp.next(NS);
                }
                if (l < 4 && lane == 0 && mw == 0) FZ_CLK(2 + l);
            }
        }
    } else if (warp < G_WARP0) {
        // ================================================================= epilogue warps: sample columns 0..31 (CTA 0's samples)
        epilogue_pass(0, tmem_base);
        if (warp == EPI_WARP0) {
            // this CTA's samples: shallow part + the partial sums of both CTAs' neuron tiles
            mbar_wait_cluster(&bars->fin, 0, p.err, 34);
            mbar_wait(&bars->shallow_ready, 0, p.err, 33);
            const long long b = (long long)tile * TS + lane;
            float z = bars->shallow[lane];
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) z += bars->red[r][j][q][lane];
            if (b < p.B) {
                if (p.logits) p.logits[b] = z;
                if (p.prob) p.prob[b] = 1.0f / (1.0f + expf(-z));
            }
            if (lane == 0) FZ_CLK(16);
        }
    } else if (!gather_warp) {
        // ================================================================= warps 18-19: the second half of epilogue set D
        epilogue_pass(3, tmem_base);
    } else {
        // ================================================================= gather group (register path of fused_tc.cu)
        const int gtid = threadIdx.x - 32 * G_WARP0;
        constexpr int FK = FT * KT, Kp = (FK + 15) & ~15;
        TileSmem sm;
        sm.img = base + p.oImg;
        sm.E = reinterpret_cast<float*>(base + p.oE);
        sm.part = reinterpret_cast<float*>(base + p.oPart);
        sm.idx = reinterpret_cast<int32_t*>(base + p.oIdx);
        sm.xv = reinterpret_cast<float*>(base + p.oXv);
        sm.EP = e_pitch(FK);
        const int64_t b0 = (int64_t)tile * TS;
        int64_t left = p.ep.B - b0;
        const int nrows = (int)(left < 0 ? 0 : (left > TS ? TS : left));
        float first_acc[G_ROUNDS];
        long long* gclk = p.clk ? p.clk + blockIdx.x * FZ_NCLK + 96 : nullptr;
        embed_gather<FT, KT, TS, G_ROUNDS, BAR_GATHER>(p.ep, sm, base + p.oImg, true, gtid, G_THREADS, b0, nrows, first_acc, gclk);
        if (gtid == 0) FZ_CLK(20);
        const ImgLayout IL = img_layout(FT, KT);
        const ImgHeader* hdr = reinterpret_cast<const ImgHeader*>(sm.img + IL.oHdr);
        const PairEnt* sPairs = reinterpret_cast<const PairEnt*>(sm.img + IL.oPairs);
        const float* sWl = reinterpret_cast<const float*>(sm.img + IL.oWl);
        const int smp = owner_sample<TS>(gtid), kk = owner_col<TS>(gtid, G_THREADS, 0);
        const bool owner = kk < KT;
        const bool use_list = (hdr->live * 6 < FT * (FT - 1) / 2) || !up.valid;
        const float* myE = sm.E + smp * sm.EP + (owner ? kk : 0);
        float second = 0.f;
        if (use_list && owner) {
            const int n = hdr->n_list;
            float s0 = 0.f, s1 = 0.f;
            int q = 0;
#pragma unroll 1
            for (; q + 1 < n; q += 2) {
                const PairEnt a = sPairs[q], b = sPairs[q + 1];
                s0 = fmaf(a.u * myE[a.ij & 0xffffu], myE[a.ij >> 16], s0);
                s1 = fmaf(b.u * myE[b.ij & 0xffffu], myE[b.ij >> 16], s1);
            }
            if (q < n) {
                const PairEnt a = sPairs[q];
                s0 = fmaf(a.u * myE[a.ij & 0xffffu], myE[a.ij >> 16], s0);
            }
            second = s0 + s1;
        }
        float e[FT];
#pragma unroll
        for (int f = 0; f < FT; ++f) e[f] = owner ? myE[f * KT] : 0.f;
        if (gclk && gtid == 0) gclk[8] = clock64();
        group_sync<BAR_GATHER>(G_THREADS);          // every thread holds its values: the block may be overwritten
        if (gclk && gtid == 0) gclk[9] = clock64();
        asm volatile("bar.sync %0, %1;" ::"n"(BAR_INIT), "n"(PAIR_THREADS) : "memory");
        cluster_wait();                             // set-up of both CTAs complete: barriers may be used
        if (gclk && gtid == 0) gclk[10] = clock64();
        for (int i = gtid; i < TS * (Kp - FK); i += G_THREADS) {     // K padding columns [F*K, Kp) are zero
            const int s = i / (Kp - FK), col = FK + (i - s * (Kp - FK));
            unsigned char* dst = sX + (size_t)(col >> 6) * CH + (s >> 3) * X_SBO + (col & 63) * 16 + (s & 7) * 2;
            *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(0.f);
            if constexpr (SPLIT) *reinterpret_cast<__nv_bfloat16*>(dst + X_HB) = __float2bfloat16_rn(0.f);
        }
        if (owner) {
            // element (k = f * KT + kk, sample smp): chunk (k / 64) * CH + (smp / 8) * X_SBO + (k % 64) * 16 + (smp % 8) * 2.  With f
            // unrolled everything but "does f * KT + kk cross into the next 64-wide chunk" is a compile-time constant.
            unsigned char* const xbase = sX + (smp >> 3) * X_SBO + (smp & 7) * 2 + kk * 16;
            constexpr int JUMP = CH - 64 * 16;
#pragma unroll
            for (int f = 0; f < FT; ++f) {
                const int c0 = f * KT, lo = c0 & 63;
                int off = c0 * 16 + (c0 >> 6) * JUMP;
                if (lo + KT > 64) off += (kk >= 64 - lo) ? JUMP : 0;
                unsigned char* dst = xbase + off;
                const __nv_bfloat16 hi = __float2bfloat16_rn(e[f]);
                *reinterpret_cast<__nv_bfloat16*>(dst) = hi;
                if constexpr (SPLIT) *reinterpret_cast<__nv_bfloat16*>(dst + X_HB) = __float2bfloat16_rn(e[f] - __bfloat162float(hi));
            }
        }
        if (gclk && gtid == 0) gclk[11] = clock64();
        fence_async_smem();                         // these stores are local
        if (gclk && gtid == 0) gclk[12] = clock64();
        __syncwarp();
        if ((gtid & 31) == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->x_ready), 0));
        if (gtid == 0) FZ_CLK(21);
        if (gclk && gtid == 0) gclk[5] = clock64();
        if (owner) {
            float acc = first_acc[0];
            if (p.ep.flags & DFW_USE_FWLW) {
                float a0 = 0.f, a1 = 0.f;
#pragma unroll
                for (int f = 0; f < FT; ++f) {
                    if (f & 1) a1 = fmaf(e[f], sWl[f * KT + kk], a1); else a0 = fmaf(e[f], sWl[f * KT + kk], a0);
                }
                acc = a0 + a1;
            }
            if (!use_list) {
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int j = 1; j < FT; ++j) {
                    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                    for (int i = 0; i < j; ++i) {
                        const float u = up.u[ucol_off(j) + i];
                        if ((i & 3) == 0) d0 = fmaf(u, e[i], d0);
                        else if ((i & 3) == 1) d1 = fmaf(u, e[i], d1);
                        else if ((i & 3) == 2) d2 = fmaf(u, e[i], d2);
                        else d3 = fmaf(u, e[i], d3);
                    }
                    const float dot = (d0 + d1) + (d2 + d3);
                    if (j & 1) s0 = fmaf(e[j], dot, s0); else s1 = fmaf(e[j], dot, s1);
                }
                second = s0 + s1;
            }
            sm.part[kk * TS + smp] = acc + second;
        }
        group_sync<BAR_GATHER>(G_THREADS);
        if (gclk && gtid == 0) gclk[6] = clock64();
        if (gtid < nrows) {
            float tot = 0.f;
#pragma unroll 1
            for (int k = 0; k < KT; ++k) tot += sm.part[k * TS + gtid];
            bars->shallow[gtid] = tot + __ldg(p.ep.bias);
        }
        __syncwarp();
        if ((gtid & 31) == 0) mbar_arrive(&bars->shallow_ready);
        if (gtid == 0) FZ_CLK(22);
        // ---- the gather warps now serve in the epilogue sets B (8-11), C (12-15) and D (16-17, with warps 18-19)
        {
            const int set = (warp - EPI_WARP0) >> 2;        // 1, 2, 3
            tc_fence_after();
            epilogue_pass(set, bars->tmem_holder);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();        // neither CTA exits (or frees TMEM) while the other may still read its shared memory / TMEM
    if (warp == MMA_WARP0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    if (p.clk && threadIdx.x == 32 * MMA_WARP0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.clk[blockIdx.x * FZ_NCLK + 30] = clock64(); p.clk[blockIdx.x * FZ_NCLK + 31] = (long long)t;
    }
}

}  // namespace fz
}  // namespace dfw
