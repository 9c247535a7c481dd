// The whole DeepFwFM forward of a 32-sample tile in ONE kernel: gather + Xv scale + first order + FwFM second order on
// the CUDA cores, the deep MLP on tcgen05 tensor cores (TMA-fed weights, TMEM accumulators), fused logit / sigmoid.
//
// Replaces the body of DeepFMs.forward (model/DeepFMs.py:285-469) and the sigmoid of :777.
//
// Why "transposed": at the reference's batch sizes (B = 4096) a sample-major GEMM tile (128 samples x N) gives only 32
// CTAs for 148 SMs.  Here the OUTPUT NEURONS are the MMA's M dimension (128 rows of W_l per tile = one TMA box of
// nn.Linear's (out, in) layout, K-major as stored) and the SAMPLES are N = 32, so a CTA owns 32 samples, B = 4096 is
// 128 CTAs, and activations never leave the SM:  D_l^T (neurons x samples, TMEM) = W_l (smem, streamed) x X_l^T (smem).
//
//   CTA = 576 threads, one per SM, cluster of `cluster` CTAs sharing every weight box by TMA multicast:
//     warps 0-1   TMA producers, one per ring: stream (layer, 64-wide K chunk, 128-neuron tile[, hi/lo]) 16 KB weight boxes
//                 into TWO mbarrier rings -- even neuron tiles into ring 0, odd tiles into ring 1; each CTA of the cluster
//                 fetches 1/cluster of a box's rows and multicasts it to all
//     warps 2-3   MMA issuers, one per ring: tcgen05.mma.cta_group::1.kind::f16, M = 128 (64 for a short last tile),
//                 N = 32 / 64, K = 16.  The whole warp walks the loop and one elected lane issues (descriptors stay in
//                 uniform registers).  K-outer order: chunk c of every tile, then chunk c+1, so the two warps' barrier
//                 probes / commits (a few hundred cycles per stage) hide behind each other's MMAs.  One ring per warp keeps
//                 every barrier's phases observed in order (1-bit parity waits are only safe that way).
//     warps 4-7   epilogue (samples 0..15; gather warps 8-11 take samples 16..31 once their tile's interaction is done):
//                 thread = one neuron (TMEM lane); tcgen05.ld its samples, + bias, ReLU, -> bf16 (hi | lo) into
//                 the activation buffer (K-major, 128B swizzle) for the next layer; last layer: x net_1_fc, warp
//                 transpose-reduction, + shallow, optional sigmoid -> global
//     warps 8-17  gather group (embed_device.cuh): indices -> rows (cp.async) -> fix-ups (fp32 block in shared memory);
//                 each thread then takes its (sample, column) values into registers, the block is overwritten IN PLACE by
//                 the bf16 operand of layer 1, and first order + FwFM second order run from the registers while the tensor
//                 cores already work on layer 1
//
//   SPLIT = false ("bf16"):   operands rounded to bf16, fp32 accumulate -- the looser-bound path (5e-4 * max|logit|)
//   SPLIT = true  ("bf16x3"): every fp32 operand is split x = hi + lo (two bf16), and the product is
//                 W_hi X_hi + W_hi X_lo + W_lo X_hi with fp32 accumulation: relative error ~2^-17 per product, which
//                 keeps the logits inside the reference's fp32 bound (1e-5 * max|logit|; measured 2e-7..2e-6 on the
//                 golden cases).  X_hi and X_lo are stacked along N (one N = 64 MMA against W_hi), W_lo X_hi is a second
//                 N = 32 MMA into the same accumulator columns: 2x the tensor time of bf16, not 3x.
//
//   Shared memory: X (ONE activation buffer; K-outer order makes chunk c dead once every tile has consumed it, so layer l+1's
//   input overwrites layer l's; the fp32 gather block aliases it) | ring 0 | ring 1 | shallow image | misc
//   TMEM: 2 x 256 columns (layer parity) x (<= 4 neuron tiles x 32|64 sample columns)
//   Limits of the fused form: depth <= 4, widths <= 512, F*K <= 512, K <= 20.  Other shapes take the staged path.
#include "fused_common.cuh"
#include "fused_pair.cuh"
#include "fused_wide.cuh"

namespace dfw {
namespace fz {

using namespace dfw::tc;

struct Bars {
    uint64_t full[2][RING_MAX], empty[2][RING_MAX];
    uint64_t x_ready[MAX_KCH], shallow_ready, tile_done;
    uint64_t act_ready[2][MAX_MT], acc_full[2][MAX_MT];
    uint32_t tmem_holder, pad_;
    float shallow[TS];
    float red[2 * EPI_WARPS][16];                 // [sample half * 4 + lane quarter][sample % 16]
};

// ---------------------------------------------------------------------------------------- the kernel
// FT/KT > 0: compile-time field count / embedding width (KT <= 10): the fp32 gather block aliases the activation buffer
// and phase D runs from registers.  FT == 0: generic shapes, separate fp32 block, pair-list second order.
template <bool SPLIT, int FT, int KT>
__global__ void __launch_bounds__(NTHREADS, 1)
fused_forward_kernel(const __grid_constant__ Maps maps, const __grid_constant__ UParam up, const Params p) {
    constexpr int NB = SPLIT ? 64 : 32;                 // B-operand rows per chunk == accumulator columns per neuron tile
    constexpr int CH = NB * 128;                        // bytes of one 64-wide K chunk of the activation buffer
    constexpr int H = SPLIT ? 2 : 1;                    // weight boxes per (tile, chunk): hi [, lo]
    static_assert(KT <= G_WARPS, "the register path owns one embedding column per gather warp pair");
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment: the 128B swizzle is a function of shared-memory address bits [4,10)
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sX = base;
    unsigned char* sW = base + p.oRing;                 // ring 0 stages, then ring 1 stages
    Bars* bars = reinterpret_cast<Bars*>(base + p.oMisc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = p.depth, CL = p.cluster;
    const uint32_t NS0 = (uint32_t)p.nst[0], NS1 = (uint32_t)p.nst[1];
    const uint16_t cmask = (uint16_t)((1u << CL) - 1);
    if (p.clk && threadIdx.x == 0) {           // debug timeline: kernel entry in SM clocks and in the global ns timer
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.clk[blockIdx.x * FZ_NCLK + 28] = clock64(); p.clk[blockIdx.x * FZ_NCLK + 29] = (long long)t;
    }

    if (threadIdx.x == 0) {
        for (int r = 0; r < 2; ++r)
            for (int s = 0; s < RING_MAX; ++s) { mbar_init(&bars->full[r][s], 1); mbar_init(&bars->empty[r][s], CL); }
        // producers of these barriers arrive once per WARP (fence, __syncwarp, lane 0): hundreds of arrives on one mbarrier serialise
        for (int c = 0; c < MAX_KCH; ++c) mbar_init(&bars->x_ready[c], G_WARPS);
        mbar_init(&bars->shallow_ready, G_WARPS);
        mbar_init(&bars->tile_done, 2 * EPI_WARPS);
        for (int b = 0; b < 2; ++b)
            for (int m = 0; m < MAX_MT; ++m) { mbar_init(&bars->act_ready[b][m], 2 * EPI_WARPS); mbar_init(&bars->acc_full[b][m], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int l = 0; l < L; ++l)
            for (int h = 0; h < H; ++h) { tma_prefetch_desc(&maps.w[l][h][0]); tma_prefetch_desc(&maps.w[l][h][1]); }
    }
    // The gather warps do not wait for the set-up (barrier init, TMEM allocation, cluster handshake ~3000 cycles): they start on
    // the batch's indices at once and only join before their first mbarrier arrive (BAR_INIT / cluster wait in the gather code).
    const bool gather_warp = warp >= G_WARP0;
    uint32_t tmem_base = 0;
    if (gather_warp) {
        if (CL > 1) cluster_arrive();
    } else {
        if (warp == MMA_WARP0) tmem_alloc(&bars->tmem_holder, 512);
        tc_fence_before();
        asm volatile("bar.sync %0, %1;" ::"n"(BAR_CORE), "n"(32 * G_WARP0) : "memory");
        if (CL > 1) { cluster_arrive(); cluster_wait(); }   // the peers' barriers exist before anything is multicast to them
        tc_fence_after();
        tmem_base = bars->tmem_holder;
        asm volatile("bar.arrive %0, %1;" ::"n"(BAR_INIT), "n"(NTHREADS) : "memory");      // releases the gather warps
    }
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
    // every CTA of a cluster runs the same number of tiles (the weight rings are shared); tiles past the end are dummies
    const int n_iter = (p.num_tiles + (int)gridDim.x - 1) / (int)gridDim.x;

    auto layer_k = [&](int l) { return pad16(l == 0 ? p.in_dim : p.widths[l - 1]); };
    auto layer_n = [&](int l) { return pad16(p.widths[l]); };

    // ---------------------------------------------------------------- one tile's epilogue for 16 of the 32 samples
    // Run by the four epilogue warps (half 0: samples 0..15) and by four gather warps once their tile's interaction is done
    // (half 1: samples 16..31) -- warp % 4 selects the TMEM lane quarter, so two warps share each quarter on different columns.
    auto epilogue_pass = [&](int it, int half, uint32_t tbase, uint32_t& acc_bits) {
        const int q4 = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = q4 * 32 + lane;
        const uint32_t taddr_row = tbase + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(16 * half);
        const int s_base = 16 * half;
        float zsum = 0.f;                              // lanes 2j, 2j+1: sum over this warp's neurons of relu(.) * fc for one sample
        for (int l = 0; l < L; ++l) {
            const int buf = l & 1, N = p.widths[l], npad = layer_n(l), MT = n_mtiles(npad);
            const bool last = (l == L - 1);
            for (int mt = 0; mt < MT; ++mt) {
                // The output of tile mt overwrites chunks 2mt, 2mt+1 of the SAME activation buffer the layer reads, so every MMA
                // reading them must have finished: both issuers interleave their tiles chunk by chunk, hence once the first tile
                // of each ring (0 and 1) is complete all chunks but the last are dead; the last chunk is written by the last
                // tile, by which time every acc_full of the layer has been waited for.
                // this neuron's bias / net_1_fc weight: issued before the wait so their L2 latency hides behind it
                const int n = mt * 128 + row;
                const int rows_valid = min(128, npad - mt * 128);   // neurons [N, npad) are zero rows: they write the K padding
                const bool real = row < rows_valid && n < N;
                const float bb = real ? __ldg(p.bias[l] + n) : 0.f;
                const float ff = (real && last) ? __ldg(p.fc + n) : 0.f;
                const int nwait = (mt == 0 && MT > 1) ? 2 : (mt == 1 ? 0 : 1);
                for (int w = 0; w < nwait; ++w) {
                    const int t = mt + w, bit = buf * MAX_MT + t;
                    FZ_PROG(8 + 4 * half + q4, (l << 16) | (t << 4) | (6 << 24));
                    mbar_wait(&bars->acc_full[buf][t], (acc_bits >> bit) & 1u, p.err, 31);
                    acc_bits ^= 1u << bit;
                }
                FZ_PROG(8 + 4 * half + q4, (l << 16) | (mt << 4) | (7 << 24));
                if (threadIdx.x == 32 * EPI_WARP0 && it == 0 && l < 4 && mt == 0) FZ_CLK(8 + 2 * l);
                tc_fence_after();
                if (q4 * 32 < rows_valid) {                     // warp-uniform: some lane of this quarter holds a neuron
                    uint32_t d[16];
                    tmem_ld16(taddr_row + (uint32_t)(buf * 256 + mt * NB), d);
                    float v[16];
                    if constexpr (SPLIT) {
                        uint32_t d2[16];
                        tmem_ld16(taddr_row + (uint32_t)(buf * 256 + mt * NB + 32), d2);
                        tmem_ld_wait();
#pragma unroll
                        for (int s = 0; s < 16; ++s) v[s] = __uint_as_float(d[s]) + __uint_as_float(d2[s]);
                    } else {
                        tmem_ld_wait();
#pragma unroll
                        for (int s = 0; s < 16; ++s) v[s] = __uint_as_float(d[s]);
                    }
                    if (last) {
                        // net_1_fc dot: this neuron's contribution to each sample, then a transpose-reduction over the warp's 32
                        // neurons: 8 + 4 + 2 + 1 + 1 shuffles leave the sum for sample s16 in lanes 2 s16' and 2 s16' + 1
#pragma unroll
                        for (int s = 0; s < 16; ++s) v[s] = fmaxf(v[s] + bb, 0.f) * ff;
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
                        zsum += v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
                    } else if (row < rows_valid) {
                        // element (sample s, k = n) of the next operand: chunk n/64, row s (hi) / 32+s (lo), 16-byte unit
                        // ((n%64)/8) ^ (row%8), byte (n%8)*2
                        unsigned char* xc = sX + (size_t)(n >> 6) * CH + (n & 7) * 2 + s_base * 128;
                        const int u = (n & 63) >> 3;
#pragma unroll
                        for (int s = 0; s < 16; ++s) {
                            const float a = fmaxf(v[s] + bb, 0.f);
                            const __nv_bfloat16 hi = __float2bfloat16_rn(a);
                            unsigned char* dst = xc + s * 128 + ((u ^ (s & 7)) << 4);     // (s_base + s) % 8 == s % 8
                            *reinterpret_cast<__nv_bfloat16*>(dst) = hi;
                            if constexpr (SPLIT)
                                *reinterpret_cast<__nv_bfloat16*>(dst + 32 * 128) = __float2bfloat16_rn(a - __bfloat162float(hi));
                        }
                    }
                }
                if (!last) {
                    fence_async_smem();            // generic-proxy stores -> visible to the tensor-core (async) proxy
                    tc_fence_before();             // TMEM reads ordered before the arrive
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars->act_ready[(l + 1) & 1][mt]);
                }
                if (threadIdx.x == 32 * EPI_WARP0 && it == 0 && l < 4 && mt == MT - 1) FZ_CLK(9 + 2 * l);
            }
        }
        // lane -> sample of its sum: bits 4..1 of the lane index, most significant first
        const int s16 = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
        if ((lane & 1) == 0) bars->red[half * 4 + q4][s16] = zsum;
        tc_fence_before();
    };

    if (warp < RINGS) {
        // ================================================================= TMA producers: warp 0 feeds ring 0 (even neuron tiles),
        // warp 1 feeds ring 1 (odd tiles).  Issuing a box costs the warp ~250 cycles (probe, expect_tx, TMA issue), so one warp
        // tops out near 60 B/clk; two keep up with the MMA pipe.  Each ring is filled and drained strictly in order.
        const int pw = warp;
        const uint32_t NS = pw ? NS1 : NS0;
        uint64_t* full = bars->full[pw];
        uint64_t* empty = bars->empty[pw];
        unsigned char* ring = sW + (pw ? (size_t)NS0 * STAGE_BYTES : 0);
        RingPos rp{0, 0};
        for (int it = 0; it < n_iter; ++it) {
            for (int l = 0; l < L; ++l) {
                const int kch = (layer_k(l) + KCH - 1) / KCH, npad = layer_n(l), MT = n_mtiles(npad);
                for (int c = 0; c < kch; ++c) {
                    for (int t = pw; t < MT; t += RINGS) {
                        RingPos s0 = rp, s1 = rp;
                        if (SPLIT) s1.next(NS);
                        // both probes in flight before either is resolved (~100 cycles each even when the stage is long free)
                        const bool k0 = mbar_try(&empty[s0.s], s0.ph ^ 1);
                        const bool k1 = SPLIT ? mbar_try(&empty[s1.s], s1.ph ^ 1) : true;
                        FZ_PROG(pw, (l << 16) | (c << 8) | (t << 4) | (1 << 24));
                        if (!k0) mbar_wait(&empty[s0.s], s0.ph ^ 1, p.err, 12);
                        if (!k1) mbar_wait(&empty[s1.s], s1.ph ^ 1, p.err, 13);
                        FZ_PROG(pw, (l << 16) | (c << 8) | (t << 4) | (2 << 24));
                        const int rows = mtile_rows(npad, t), per = rows / CL;
                        unsigned char* d0 = ring + (size_t)s0.s * STAGE_BYTES + (size_t)crank * per * 128;
                        unsigned char* d1 = ring + (size_t)s1.s * STAGE_BYTES + (size_t)crank * per * 128;
                        const int row0 = t * 128 + (int)crank * per, which = rows == 64 ? 1 : 0;
                        const uint32_t bytes = (uint32_t)(rows * 128);
                        if (elect_one()) {
                            mbar_expect_tx(&full[s0.s], bytes);
                            if (CL > 1) tma_load_2d_mc(d0, &maps.w[l][0][which], &full[s0.s], c * KCH, row0, cmask);
                            else tma_load_2d(d0, &maps.w[l][0][which], &full[s0.s], c * KCH, row0);
                            if (SPLIT) {
                                mbar_expect_tx(&full[s1.s], bytes);
                                if (CL > 1) tma_load_2d_mc(d1, &maps.w[l][1][which], &full[s1.s], c * KCH, row0, cmask);
                                else tma_load_2d(d1, &maps.w[l][1][which], &full[s1.s], c * KCH, row0);
                            }
                        }
                        __syncwarp();
                        rp = s1; rp.next(NS);
                    }
                }
            }
        }
    } else if (warp < EPI_WARP0) {
        // ================================================================= MMA issuers: warp 2 = ring 0 (even tiles), warp 3 = ring 1
        const int mw = warp - MMA_WARP0;
        const uint32_t NS = mw ? NS1 : NS0;
        uint64_t* full = bars->full[mw];
        uint64_t* empty = bars->empty[mw];
        const uint32_t sW_u32 = smem_u32(sW) + (mw ? NS0 * (uint32_t)STAGE_BYTES : 0u);
        const uint32_t sX_u32 = smem_u32(sX);
        RingPos rp{0, 0};
        uint32_t act_bits = 0;                       // phase parity of act_ready[buf][g], bit buf * MAX_MT + g
        for (int it = 0; it < n_iter; ++it) {
            if (it > 0) mbar_wait(&bars->tile_done, (uint32_t)((it - 1) & 1), p.err, 20);   // TMEM drained, X free
            for (int l = 0; l < L; ++l) {
                const int buf = l & 1;
                const int K = layer_k(l), kch = (K + KCH - 1) / KCH, npad = layer_n(l), MT = n_mtiles(npad);
                const bool active = mw < MT;          // a one-tile layer puts nothing into ring 1; its barriers are still observed
                const int tA = mw, tB = mw + 2;
                const bool hasB = tB < MT;
                const int rowsA = mtile_rows(npad, tA), rowsB = hasB ? mtile_rows(npad, tB) : 128;
                const uint32_t dA = tmem_base + (uint32_t)(buf * 256 + tA * NB), dB = tmem_base + (uint32_t)(buf * 256 + tB * NB);
                const uint32_t idA = make_idesc(rowsA, NB), idB = make_idesc(rowsB, NB);
                const uint32_t idAlo = make_idesc(rowsA, 32), idBlo = make_idesc(rowsB, 32);
                for (int c = 0; c < kch; ++c) {
                    if (l == 0) {
                        // layer 1 starts on chunk c as soon as the gather group has written it
                        if (c == 0 && it == 0 && lane == 0 && mw == 0) FZ_CLK(0);
                        mbar_wait(&bars->x_ready[c], (uint32_t)(it & 1), p.err, 21);
                        if (c == 0 && it == 0 && lane == 0 && mw == 0) FZ_CLK(1);
                    } else if ((c & 1) == 0) {
                        // chunks c, c+1 hold neurons [64c, 64c+128) of the previous layer = its neuron tile c/2
                        const int g = c >> 1, bit = buf * MAX_MT + g;
                        FZ_PROG(2 + mw, (l << 16) | (c << 8) | (5 << 24));
                        mbar_wait(&bars->act_ready[buf][g], (act_bits >> bit) & 1u, p.err, 22);
                        act_bits ^= 1u << bit;
                    }
                    if (!active) continue;
                    const int ks = min(4, (K - c * KCH) / 16);
                    const uint64_t bdesc = make_desc_sw128(sX_u32 + (uint32_t)(c * CH));
                    const uint32_t acc0 = c ? 1u : 0u;
                    const bool last_c = c == kch - 1;
                    if constexpr (SPLIT) {
                        // one iteration per tile: its hi and lo boxes (two stages, probes overlapped), 2 * ks MMAs
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            if (j == 1 && !hasB) break;
                            const uint32_t dcol = j ? dB : dA, idh = j ? idB : idA, idl = j ? idBlo : idAlo;
                            RingPos s0 = rp, s1 = rp; s1.next(NS);
                            const bool k0 = mbar_try(&full[s0.s], s0.ph), k1 = mbar_try(&full[s1.s], s1.ph);
                            FZ_PROG(2 + mw, (l << 16) | (c << 8) | ((j ? tB : tA) << 4) | (3 << 24));
                            if (!k0) mbar_wait(&full[s0.s], s0.ph, p.err, 24);
                            if (!k1) mbar_wait(&full[s1.s], s1.ph, p.err, 25);
                            tc_fence_after();
                            const uint64_t ah = make_desc_sw128(sW_u32 + s0.s * (uint32_t)STAGE_BYTES);
                            const uint64_t al = make_desc_sw128(sW_u32 + s1.s * (uint32_t)STAGE_BYTES);
                            if (elect_one()) {
                                umma_bf16(dcol, ah, bdesc, idh, acc0);                     // +32 bytes (16 bf16) along K per step
                                if (ks > 1) umma_bf16(dcol, ah + 2, bdesc + 2, idh, 1u);
                                if (ks > 2) umma_bf16(dcol, ah + 4, bdesc + 4, idh, 1u);
                                if (ks > 3) umma_bf16(dcol, ah + 6, bdesc + 6, idh, 1u);
                                if (CL > 1) umma_commit_mc(&empty[s0.s], cmask); else umma_commit(&empty[s0.s]);
                                umma_bf16(dcol, al, bdesc, idl, 1u);
                                if (ks > 1) umma_bf16(dcol, al + 2, bdesc + 2, idl, 1u);
                                if (ks > 2) umma_bf16(dcol, al + 4, bdesc + 4, idl, 1u);
                                if (ks > 3) umma_bf16(dcol, al + 6, bdesc + 6, idl, 1u);
                                if (CL > 1) umma_commit_mc(&empty[s1.s], cmask); else umma_commit(&empty[s1.s]);
                                if (last_c) umma_commit(&bars->acc_full[buf][j ? tB : tA]);   // accumulators of (layer l, tile) complete
                            }
                            __syncwarp();
                            rp = s1; rp.next(NS);
                        }
                    } else {
                        // one iteration per chunk: this ring's (up to) two tiles, one box each
                        RingPos s0 = rp, s1 = rp; s1.next(NS);
                        const bool k0 = mbar_try(&full[s0.s], s0.ph), k1 = hasB ? mbar_try(&full[s1.s], s1.ph) : true;
                        FZ_PROG(2 + mw, (l << 16) | (c << 8) | (tA << 4) | (3 << 24));
                        if (!k0) mbar_wait(&full[s0.s], s0.ph, p.err, 24);
                        if (!k1) mbar_wait(&full[s1.s], s1.ph, p.err, 25);
                        tc_fence_after();
                        const uint64_t a0 = make_desc_sw128(sW_u32 + s0.s * (uint32_t)STAGE_BYTES);
                        const uint64_t a1 = make_desc_sw128(sW_u32 + s1.s * (uint32_t)STAGE_BYTES);
                        if (elect_one()) {
                            umma_bf16(dA, a0, bdesc, idA, acc0);
                            if (ks > 1) umma_bf16(dA, a0 + 2, bdesc + 2, idA, 1u);
                            if (ks > 2) umma_bf16(dA, a0 + 4, bdesc + 4, idA, 1u);
                            if (ks > 3) umma_bf16(dA, a0 + 6, bdesc + 6, idA, 1u);
                            if (CL > 1) umma_commit_mc(&empty[s0.s], cmask); else umma_commit(&empty[s0.s]);
                            if (last_c) umma_commit(&bars->acc_full[buf][tA]);
                            if (hasB) {
                                umma_bf16(dB, a1, bdesc, idB, acc0);
                                if (ks > 1) umma_bf16(dB, a1 + 2, bdesc + 2, idB, 1u);
                                if (ks > 2) umma_bf16(dB, a1 + 4, bdesc + 4, idB, 1u);
                                if (ks > 3) umma_bf16(dB, a1 + 6, bdesc + 6, idB, 1u);
                                if (CL > 1) umma_commit_mc(&empty[s1.s], cmask); else umma_commit(&empty[s1.s]);
                                if (last_c) umma_commit(&bars->acc_full[buf][tB]);
                            }
                        }
                        __syncwarp();
                        rp = s0; rp.next(NS);
                        if (hasB) rp.next(NS);
                    }
                }
                if (it == 0 && l < 4 && lane == 0 && mw == 0) FZ_CLK(2 + l);
            }
        }
    } else if (warp < G_WARP0) {
        // ================================================================= epilogue warps (thread = neuron), samples 0..15
        uint32_t acc_bits = 0;                         // phase parity of acc_full[buf][mt], bit buf * MAX_MT + mt
        for (int it = 0; it < n_iter; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            epilogue_pass(it, 0, tmem_base, acc_bits);
            asm volatile("bar.sync %0, %1;" ::"n"(BAR_EPI), "n"(2 * EPI_THREADS) : "memory");
            if (warp == EPI_WARP0) {
                mbar_wait(&bars->shallow_ready, (uint32_t)(it & 1), p.err, 33);
                const long long b = (long long)tile * TS + lane;
                float z = bars->shallow[lane];
#pragma unroll
                for (int q = 0; q < 4; ++q) z += bars->red[(lane >> 4) * 4 + q][lane & 15];
                if (b < p.B) {
                    if (p.logits) p.logits[b] = z;
                    if (p.prob) p.prob[b] = 1.0f / (1.0f + expf(-z));
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tile_done);
            if (threadIdx.x == 32 * EPI_WARP0 && it == 0) FZ_CLK(16);
        }
    } else {
        // ================================================================= gather group
        const int gtid = threadIdx.x - 32 * G_WARP0;
        const int F = FT > 0 ? FT : p.ep.F, K = KT > 0 ? KT : p.ep.K;
        const int FK = F * K, Kp = pad16(FK);
        TileSmem sm;
        sm.img = base + p.oImg;
        sm.E = reinterpret_cast<float*>(base + p.oE);    // FT > 0: aliases the activation buffer
        sm.part = reinterpret_cast<float*>(base + p.oPart);
        sm.idx = reinterpret_cast<int32_t*>(base + p.oIdx);
        sm.xv = reinterpret_cast<float*>(base + p.oXv);
        sm.EP = e_pitch(FK);
        uint32_t helper_acc_bits = 0;
        bool joined = false;
        auto join_core = [&]() {          // first use of the mbarriers: the set-up by the other warps (and the peers') is complete
            if (!joined) {
                asm volatile("bar.sync %0, %1;" ::"n"(BAR_INIT), "n"(NTHREADS) : "memory");
                if (CL > 1) cluster_wait();
                joined = true;
            }
        };
        for (int it = 0; it < n_iter; ++it) {
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int64_t b0 = (int64_t)tile * TS;
            int64_t left = p.ep.B - b0;
            const int nrows = (int)(left < 0 ? 0 : (left > TS ? TS : left));
            if (it > 0) { join_core(); mbar_wait(&bars->tile_done, (uint32_t)((it - 1) & 1), p.err, 40); }   // the previous tile is done with X
            float first_acc[G_ROUNDS];
            long long* gclk = (p.clk && it == 0) ? p.clk + blockIdx.x * FZ_NCLK + 96 : nullptr;
            embed_gather<FT, KT, TS, G_ROUNDS, BAR_GATHER>(p.ep, sm, base + p.oImg, it == 0, gtid, G_THREADS, b0, nrows, first_acc, gclk);
            if (gtid == 0 && it == 0) FZ_CLK(20);
            if constexpr (FT > 0) {
                // ---- register path: this thread owns (sample smp, column kk); its F values leave the fp32 block for registers,
                // the block is then overwritten in place by the bf16 (hi | lo) operand of layer 1, and the interaction runs from
                // the registers while the tensor cores start.  A pruned field matrix (pair list) is walked before the overwrite.
                constexpr int FTc = FT > 0 ? FT : 1, KTc = KT > 0 ? KT : 1;
                const ImgLayout IL = img_layout(FTc, KTc);
                const ImgHeader* hdr = reinterpret_cast<const ImgHeader*>(sm.img + IL.oHdr);
                const PairEnt* sPairs = reinterpret_cast<const PairEnt*>(sm.img + IL.oPairs);
                const float* sWl = reinterpret_cast<const float*>(sm.img + IL.oWl);
                const int smp = owner_sample<TS>(gtid), kk = owner_col<TS>(gtid, G_THREADS, 0);
                const bool owner = kk < KTc;
                const bool use_list = (hdr->live * 6 < FTc * (FTc - 1) / 2) || !up.valid;
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
                float e[FTc];
#pragma unroll
                for (int f = 0; f < FTc; ++f) e[f] = owner ? myE[f * KTc] : 0.f;
                group_sync<BAR_GATHER>(G_THREADS);          // every thread holds its values: the block may be overwritten
                join_core();
                for (int i = gtid; i < TS * (Kp - FK); i += G_THREADS) {     // K padding columns [F*K, Kp) are zero
                    const int s = i / (Kp - FK), col = FK + (i - s * (Kp - FK));
                    unsigned char* dst = sX + (size_t)(col >> 6) * CH + s * 128 + ((((col & 63) >> 3) ^ (s & 7)) << 4) + (col & 7) * 2;
                    *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(0.f);
                    if constexpr (SPLIT) *reinterpret_cast<__nv_bfloat16*>(dst + 32 * 128) = __float2bfloat16_rn(0.f);
                }
                if (owner) {
#pragma unroll
                    for (int f = 0; f < FTc; ++f) {
                        const int col = f * KTc + kk;       // element (sample smp, k = col): chunk col/64, 16-byte unit, byte (col%8)*2
                        unsigned char* dst = sX + (size_t)(col >> 6) * CH + smp * 128 + ((((col & 63) >> 3) ^ (smp & 7)) << 4) + (col & 7) * 2;
                        const __nv_bfloat16 hi = __float2bfloat16_rn(e[f]);
                        *reinterpret_cast<__nv_bfloat16*>(dst) = hi;
                        if constexpr (SPLIT) *reinterpret_cast<__nv_bfloat16*>(dst + 32 * 128) = __float2bfloat16_rn(e[f] - __bfloat162float(hi));
                    }
                }
                // One release for the whole operand: signalling chunk by chunk lets layer 1 start ~1700 cycles earlier, but its
                // MMA + TMA traffic then saturates shared memory and the remaining column stores (and phase D) crawl -- measured
                // slower end to end.
                fence_async_smem();
                __syncwarp();
                if ((gtid & 31) == 0) for (int c = 0; c < (Kp + KCH - 1) / KCH; ++c) mbar_arrive(&bars->x_ready[c]);
                if (gtid == 0 && it == 0) FZ_CLK(21);
                if (gclk && gtid == 0) gclk[5] = clock64();
                if (owner) {
                    float acc = first_acc[0];
                    if (p.ep.flags & DFW_USE_FWLW) {          // model/DeepFMs.py:344-345
                        float a0 = 0.f, a1 = 0.f;
#pragma unroll
                        for (int f = 0; f < FTc; ++f) {
                            if (f & 1) a1 = fmaf(e[f], sWl[f * KTc + kk], a1); else a0 = fmaf(e[f], sWl[f * KTc + kk], a0);
                        }
                        acc = a0 + a1;
                    }
                    if (!use_list) {
                        // second = sum_j e_j * (sum_{i<j} U_ij e_i): every U_ij is a constant-bank operand, 4 independent chains per j
                        float s0 = 0.f, s1 = 0.f;
#pragma unroll
                        for (int j = 1; j < FTc; ++j) {
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
                if (gtid < nrows) {                             // fixed-order reduction over k
                    float tot = 0.f;
#pragma unroll 1
                    for (int k = 0; k < KTc; ++k) tot += sm.part[k * TS + gtid];
                    bars->shallow[gtid] = tot + __ldg(p.ep.bias);
                }
                __syncwarp();
                if ((gtid & 31) == 0) mbar_arrive(&bars->shallow_ready);
                if (gtid == 0 && it == 0) FZ_CLK(22);
            } else {
                // ---- generic shapes: separate fp32 block -> bf16 (hi | lo) operand of layer 1, then the pair-list interaction
                const int units = Kp >> 3;                  // 16-byte (8 x bf16) units per operand row
                for (int i = gtid; i < TS * units; i += G_THREADS) {
                    const int s = i / units, u8 = i - s * units;
                    const float* src = sm.E + s * sm.EP + 8 * u8;
                    float x[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c0 = 8 * u8 + 2 * j;
                        float2 t = make_float2(0.f, 0.f);
                        if (c0 + 1 < FK) t = *reinterpret_cast<const float2*>(src + 2 * j);
                        else if (c0 < FK) t.x = src[2 * j];
                        x[2 * j] = t.x; x[2 * j + 1] = t.y;
                    }
                    uint4 hi4, lo4;
                    uint32_t* hp = reinterpret_cast<uint32_t*>(&hi4);
                    uint32_t* lp = reinterpret_cast<uint32_t*>(&lo4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(x[2 * j], x[2 * j + 1]);
                        hp[j] = *reinterpret_cast<const uint32_t*>(&h2);
                        if constexpr (SPLIT) {
                            const float2 hf = __bfloat1622float2(h2);
                            const __nv_bfloat162 l2 = __floats2bfloat162_rn(x[2 * j] - hf.x, x[2 * j + 1] - hf.y);
                            lp[j] = *reinterpret_cast<const uint32_t*>(&l2);
                        }
                    }
                    unsigned char* dst = sX + (size_t)(u8 >> 3) * CH + s * 128 + (((u8 & 7) ^ (s & 7)) << 4);
                    *reinterpret_cast<uint4*>(dst) = hi4;
                    if constexpr (SPLIT) *reinterpret_cast<uint4*>(dst + 32 * 128) = lo4;
                }
                join_core();
                fence_async_smem();
                __syncwarp();
                if ((gtid & 31) == 0) for (int c = 0; c < (Kp + KCH - 1) / KCH; ++c) mbar_arrive(&bars->x_ready[c]);
                if (gtid == 0 && it == 0) FZ_CLK(21);
                embed_interact<FT, KT, TS, G_ROUNDS, BAR_GATHER, true>(p.ep, sm, gtid, G_THREADS, nrows, first_acc, bars->shallow, gclk, nullptr);
                __syncwarp();
                if ((gtid & 31) == 0) mbar_arrive(&bars->shallow_ready);
                if (gtid == 0 && it == 0) FZ_CLK(22);
            }
            // ---- the first four gather warps now serve as the second epilogue set (samples 16..31) of this tile
            if (warp < G_WARP0 + EPI_WARPS) {
                tc_fence_after();
                epilogue_pass(it, 1, bars->tmem_holder, helper_acc_bits);
                asm volatile("bar.sync %0, %1;" ::"n"(BAR_EPI), "n"(2 * EPI_THREADS) : "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->tile_done);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();    // no CTA exits while a peer may still multicast into it
    if (warp == MMA_WARP0) tmem_dealloc(tmem_base, 512);
    if (p.clk && threadIdx.x == 32 * MMA_WARP0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.clk[blockIdx.x * FZ_NCLK + 30] = clock64(); p.clk[blockIdx.x * FZ_NCLK + 31] = (long long)t;
    }
}

// fp32 (out, in) row-major -> bf16 hi (and lo = bf16(w - hi)) images (pad16(out), pad64(in)), zero padded
__global__ void pack_split_kernel(const float* __restrict__ W, int out_dim, int in_dim, int out_pad, int in_pad,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
    const long long total = (long long)out_pad * in_pad;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / in_pad), c = (int)(i - (long long)r * in_pad);
        const float w = (r < out_dim && c < in_dim) ? W[(long long)r * in_dim + c] : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(w);
        if (hi) hi[i] = h;
        if (lo) lo[i] = __float2bfloat16_rn(w - __bfloat162float(h));
    }
}

// ---------------------------------------------------------------------------------------- host side
static long long* g_clk = nullptr;
static volatile int* g_prog = nullptr;
static int* g_errword = nullptr;          // debug: where a barrier watchdog writes its code when the caller passes no error word

struct Plan {
    Params p;
    UParam up;
    size_t smem_bytes;
};

// shapes with a compiled register path (the dataset shapes BASELINE.json names)
static bool specialised(int F, int K) { return K == 10 && (F == 39 || F == 47); }

// Shared-memory plan; returns false (with `why`) when the shapes do not fit the fused form.
static bool make_plan(const dfw_model* m, bool split, Plan& pl, const char** why) {
    const int F = m->field_size, K = m->embedding_size, num = m->numerical, FK = F * K;
    *why = nullptr;
    if (!(m->flags & DFW_USE_DEEP)) { *why = "no deep part"; return false; }
    if (m->depth > MAX_L) { *why = "depth > 4"; return false; }
    if (FK > MAX_W) { *why = "F*K > 512"; return false; }
    if (K > G_ROUNDS * G_WARPS) { *why = "K > 20"; return false; }
    int kmax = pad16(FK);
    for (int l = 0; l < m->depth; ++l) {
        if (m->widths[l] > MAX_W) { *why = "layer width > 512"; return false; }
        if (!m->Wbf16[l] || (split && !m->Wbf16_lo[l])) { *why = "bf16 weight image missing"; return false; }
        if (l + 1 < m->depth && pad16(m->widths[l]) > kmax) kmax = pad16(m->widths[l]);
    }
    Params& p = pl.p;
    const bool alias = specialised(F, K);          // the register path: the fp32 gather block lives inside the activation buffer
    // one 64-wide K chunk of the activation buffer: 32 samples x 128 bytes (hi [+ lo]) in this file's K-major form; the pair
    // kernel's sample-contiguous form (fused_pair.cuh) pads it to X_HB
    const int CH = (split ? 2 : 1) * (alias ? (int)X_HB : 32 * 128);
    p.x_chunks = (kmax + KCH - 1) / KCH;
    const TileSizes ts = tile_sizes(F, K, num, TS);
    size_t x = (size_t)p.x_chunks * CH;
    if (alias && ts.bE > x) x = ts.bE;
    x = (x + 1023) / 1024 * 1024;
    size_t o = x;
    p.oE = 0;
    if (!alias) { p.oE = (uint32_t)o; o += (ts.bE + 1023) / 1024 * 1024; }
    p.oRing = (uint32_t)o;
    const size_t bars_bytes = sizeof(Bars) > sizeof(PairBars) ? sizeof(Bars) : sizeof(PairBars);
    const size_t tail = up16(img_layout(F, K).total) + ts.bPart + ts.bIdx + ts.bXv + up16(bars_bytes) + 64;
    // two rings; a SPLIT iteration takes 2 stages of a ring (hi + lo box), so its rings hold an even number of stages
    const int min_ring = 2;
    if (o + tail + 1024 + 2 * (size_t)min_ring * STAGE_BYTES > SMEM_LIMIT) { *why = "shared memory: the two weight rings do not fit"; return false; }
    int total = (int)((SMEM_LIMIT - 1024 - o - tail) / STAGE_BYTES);
    if (total > 2 * RING_MAX) total = 2 * RING_MAX;
    int n0 = (total + 1) / 2, n1 = total / 2;
    if (split) { n0 &= ~1; n1 &= ~1; }
    p.nst[0] = n0; p.nst[1] = n1;
    o += (size_t)(n0 + n1) * STAGE_BYTES; p.oImg = (uint32_t)o;
    o += up16(img_layout(F, K).total);    p.oPart = (uint32_t)o;
    o += ts.bPart;                        p.oIdx = (uint32_t)o;
    o += ts.bIdx;                         p.oXv = (uint32_t)o;
    o += ts.bXv;                          p.oMisc = (uint32_t)((o + 15) & ~size_t(15));
    o = p.oMisc + bars_bytes;
    pl.smem_bytes = o + 1024;
    return true;
}

template <bool SPLIT, int FT, int KT>
static int launch(const Maps& maps, const Plan& pl, int grid, cudaStream_t st) {
    auto kern = fused_forward_kernel<SPLIT, FT, KT>;
    static thread_local bool configured_dev[16] = {};       // per device: function attributes are per context
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool& configured = configured_dev[cur_dev & 15];
    if (!configured) {
        DFW_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
        DFW_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = pl.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)pl.p.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    DFW_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, maps, pl.up, pl.p));
    count_launch();
    return check_launch("fused_forward_kernel");
}

// how many clusters of `cl` CTAs of this kernel the device can hold at once (cached per device and cluster size)
template <bool SPLIT, int FT, int KT>
static int max_clusters(int cl, size_t smem_bytes) {
    static thread_local int cache[8][5] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 8 && cache[dev][cl] > 0) return cache[dev][cl];
    auto kern = fused_forward_kernel<SPLIT, FT, KT>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cl * 64));
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    if (n <= 0) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        n = sms / cl * 7 / 8;
    }
    if (dev < 8) cache[dev][cl] = n;
    return n;
}

// Tensor maps depend only on the weight images and the cluster size: encode once per (model, cluster), not per call.
struct MapKey {
    const void* img[MAX_L][2];
    int widths[MAX_L], in_dim, depth, cluster, split;
    bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapCache {
    static constexpr int N = 16;
    std::mutex mu;
    MapKey keys[N];
    Maps maps[N];
    int used = 0, next = 0;
};
static MapCache g_maps;

static int get_maps(const dfw_model* m, bool split, int cluster, int in_dim, Maps& out) {
    MapKey key;
    memset(&key, 0, sizeof(key));
    for (int l = 0; l < m->depth; ++l) { key.img[l][0] = m->Wbf16[l]; key.img[l][1] = split ? m->Wbf16_lo[l] : nullptr; key.widths[l] = m->widths[l]; }
    key.in_dim = in_dim; key.depth = m->depth; key.cluster = cluster; key.split = split;
    std::lock_guard<std::mutex> lock(g_maps.mu);
    for (int i = 0; i < g_maps.used; ++i)
        if (g_maps.keys[i] == key) { out = g_maps.maps[i]; return 0; }
    Maps maps;
    int k = in_dim;
    for (int l = 0; l < m->depth; ++l) {
        const int npad = pad16(m->widths[l]), kpad = (k + 63) / 64 * 64;
        for (int h = 0; h < (split ? 2 : 1); ++h) {
            const void* img = h ? m->Wbf16_lo[l] : m->Wbf16[l];
            if (int rc = make_map(&maps.w[l][h][0], img, npad, kpad, kpad, 128 / cluster)) return rc;
            if (int rc = make_map(&maps.w[l][h][1], img, npad, kpad, kpad, 64 / cluster)) return rc;
        }
        k = m->widths[l];
    }
    const int slot = g_maps.used < MapCache::N ? g_maps.used++ : (g_maps.next++ % MapCache::N);
    g_maps.keys[slot] = key;
    g_maps.maps[slot] = maps;
    out = maps;
    return 0;
}

static int env_int(const char* name, int dflt) {
    const char* e = dbg_getenv(name);
    return e ? atoi(e) : dflt;
}
static int env_cluster() {
    static const int v = [] { const char* e = dbg_getenv("DFW_FUSED_CLUSTER"); return e ? atoi(e) : 0; }();
    return v;
}

static int env_pair() {
    static const int v = env_int("DFW_FUSED_PAIR", 1);
    return v;
}

// cta_group::2 pair kernel (fused_pair.cuh): CTA b owns tile b, CTAs (2q, 2q+1) form a pair
template <bool SPLIT, int FT, int KT>
static int launch_pair(const dfw_model* m, Plan& pl, Maps& maps, cudaStream_t st) {
    Params& p = pl.p;
    p.cluster = 2;
    if (int rc = get_maps(m, SPLIT, 1, p.in_dim, maps)) return rc;       // whole 128-row boxes: each CTA loads its own tile
    auto kern = fused_pair_kernel<SPLIT, FT, KT>;
    static thread_local bool configured_dev[16] = {};       // per device: function attributes are per context
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool& configured = configured_dev[cur_dev & 15];
    if (!configured) {
        DFW_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
        configured = true;
    }
    const unsigned grid = (unsigned)((p.num_tiles + 1) / 2 * 2);
    kern<<<grid, PAIR_THREADS, pl.smem_bytes, st>>>(maps, pl.up, pl.p);
    count_launch();
    return check_launch("fused_pair_kernel");
}

// Persistent 64-sample pair kernel (fused_wide.cuh).  Shared-memory plan: X (64 samples) | weight ring | shallow image | per-column
// partial sums | numeric rows | barriers.  Returns false when it does not fit (the caller then takes the 32-sample kernels).
static int env_wide() {
    static const int v = env_int("DFW_FUSED_WIDE", 1);
    return v;
}
static bool make_plan_wide(const dfw_model* m, bool split, const Plan& pl, wd::WideParams& wp, size_t& smem_bytes) {
    const int F = m->field_size, K = m->embedding_size, num = m->numerical;
    wp.p = pl.p;
    Params& p = wp.p;
    const int CH = (split ? 2 : 1) * (int)wd::X_HBW;
    size_t o = ((size_t)p.x_chunks * CH + 1023) / 1024 * 1024;
    p.oE = 0;
    p.oRing = (uint32_t)o;
    // compact copy of the shallow image: header | fwlw weights | field descriptors (fused_wide.cuh)
    const size_t img = 16 + up16(sizeof(float) * F * K) + up16(sizeof(dfw_field_desc) * F);
    const size_t part = up16(sizeof(float) * K * wd::TSW), nrow = up16(sizeof(float) * (num > 0 ? num : 1) * K);
    const size_t tail = img + part + nrow + up16(sizeof(wd::WideBars)) + 64;
    if (o + tail + 1024 + 2 * (size_t)STAGE_BYTES > SMEM_LIMIT) return false;
    int ns = (int)((SMEM_LIMIT - 1024 - o - tail) / STAGE_BYTES);
    if (ns > wd::NS_MAX) ns = wd::NS_MAX;
    {
        static const int cap = env_int("DFW_FUSED_STAGES", 0);
        if (cap >= 2 && cap < ns) ns = cap;
    }
    if (ns < (split ? 4 : 2)) return false;
    // grouped ring (one full / empty barrier and ONE tcgen05.commit per two boxes): bf16x3 pairs a step's hi and lo box; bf16
    // pairs two steps, which needs an even number of boxes per tile so that no group is left half filled at the end
    long long boxes = 0;
    {
        int k = p.in_dim;
        for (int l = 0; l < p.depth; ++l) {
            boxes += (long long)((n_mtiles(pad16(p.widths[l])) + 1) / 2) * ((pad16(k) + KCH - 1) / KCH);
            k = p.widths[l];
        }
    }
    static const int grp_env = env_int("DFW_WIDE_GRP", 2);
    wp.grp = (grp_env == 2 && ns >= 4 && (split || boxes % 2 == 0)) ? 2 : 1;
    if (wp.grp == 2) ns &= ~1;
    wp.ns = ns;
    o += (size_t)ns * STAGE_BYTES; p.oImg = (uint32_t)o;
    o += img;                       p.oPart = (uint32_t)o;
    o += part;                      wp.oNum = (uint32_t)o;
    o += nrow;                      p.oMisc = (uint32_t)((o + 15) & ~size_t(15));
    p.oIdx = p.oXv = 0;
    smem_bytes = p.oMisc + sizeof(wd::WideBars) + 1024;
    return smem_bytes <= SMEM_LIMIT;
}

template <bool SPLIT, int FT, int KT, int NUMT>
static int launch_wide(const dfw_model* m, const Plan& pl, Maps& maps, cudaStream_t st, bool& taken) {
    taken = false;
    if (!pl.up.valid || m->numerical != NUMT) return 0;       // needs the field matrix as a kernel parameter
    wd::WideParams wp;
    size_t smem_bytes = 0;
    if (!make_plan_wide(m, SPLIT, pl, wp, smem_bytes)) return 0;
    taken = true;
    wp.p.cluster = 2;
    wp.dbg = env_int("DFW_WIDE_DBG", 0);
    if (int rc = get_maps(m, SPLIT, 1, wp.p.in_dim, maps)) return rc;    // whole 128-row boxes: each CTA loads its own tile
    auto kern = wd::fused_wide_kernel<SPLIT, FT, KT, NUMT>;
    static thread_local bool configured_dev[16] = {};       // per device: function attributes are per context
    static thread_local int sm_count[16] = {};
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    bool& configured = configured_dev[cur_dev & 15];
    if (!configured) {
        DFW_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LIMIT));
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cur_dev);
        sm_count[cur_dev & 15] = sms;
        configured = true;
    }
    const long long tiles = (wp.p.B + wd::TSW - 1) / wd::TSW;
    wp.n_pair_tiles = (int)((tiles + 1) / 2);
    int pairs = sm_count[cur_dev & 15] / 2;
    {
        static const int cap = env_int("DFW_FUSED_PAIRS", 0);
        if (cap > 0 && cap < pairs) pairs = cap;
    }
    if (wp.n_pair_tiles < pairs) pairs = wp.n_pair_tiles;
    {
        // Tiles per pair.  A pair that owns ONE tile pays the whole gather (20 k cycles) in front of its first MMA: 68 k cycles per
        // tile; from the second tile on the gather runs under the previous tile's MLP: (20 + 2 x 43 + 3.5 + 5) / 2 = 57 k per tile.
        // With DFW_HINT_THROUGHPUT (several launches in flight) a launch that has the tiles for it therefore uses half as many pairs
        // with two tiles each: 16 % less SM time per batch (measured 204 -> 238 M samples/s at B = 4096 with 16 launches in
        // flight; three tiles per pair gain nothing more), 67 instead of 41 us for a launch running alone -- hence a hint.
        static const int tpp_env = env_int("DFW_WIDE_TPP", 0);
        const int tpp = tpp_env > 0 ? tpp_env : ((m->flags & DFW_HINT_THROUGHPUT) && wp.n_pair_tiles >= 8) ? 2 : 1;
        const int want = (wp.n_pair_tiles + tpp - 1) / tpp;
        if (want < pairs) pairs = want;
    }
    kern<<<(unsigned)(2 * pairs), wd::THREADS, smem_bytes, st>>>(maps, pl.up, wp);
    count_launch();
    return check_launch("fused_wide_kernel");
}

template <bool SPLIT, int FT, int KT>
static int run(const dfw_model* m, Plan& pl, Maps& maps, cudaStream_t st) {
    Params& p = pl.p;
    if constexpr (FT > 0) {
        if (env_wide()) {      // every batch size: one kernel per model keeps a sample's bits independent of the batch it arrives in
            bool taken = false;
            const int rc = launch_wide<SPLIT, FT, KT, (FT == 39 ? 13 : 11)>(m, pl, maps, st, taken);
            if (taken) return rc;
        }
        if (env_pair() && p.num_tiles >= 2) return launch_pair<SPLIT, FT, KT>(m, pl, maps, st);
    }
    // cluster size: the largest of {4, 2, 1} that still covers all tiles in the fewest waves
    int best_cl = 1, best_grid = 1, best_iter = 1 << 30;
    for (int cl = 4; cl >= 1; cl >>= 1) {
        const int mc = cl == 1 ? 148 : max_clusters<SPLIT, FT, KT>(cl, pl.smem_bytes);
        if (mc <= 0) continue;
        const int need = (p.num_tiles + cl - 1) / cl;
        const int clusters = need < mc ? need : mc;
        const int grid = clusters * cl;
        const int iters = (p.num_tiles + grid - 1) / grid;
        if (iters < best_iter) { best_iter = iters; best_cl = cl; best_grid = grid; }
    }
    const int cl = env_cluster();
    if (cl == 1 || cl == 2 || cl == 4) {
        const int mc = cl == 1 ? 148 : max_clusters<SPLIT, FT, KT>(cl, pl.smem_bytes);
        const int need = (p.num_tiles + cl - 1) / cl;
        best_cl = cl; best_grid = (need < mc ? need : mc) * cl;
    }
    p.cluster = best_cl;
    {   // debug knobs
        static const int cap = env_int("DFW_FUSED_STAGES", 0);      // stages per ring
        if (cap >= 2) { if (cap < p.nst[0]) p.nst[0] = cap; if (cap < p.nst[1]) p.nst[1] = cap; }
    }
    if (int rc = get_maps(m, SPLIT, best_cl, p.in_dim, maps)) return rc;
    return launch<SPLIT, FT, KT>(maps, pl, best_grid, st);
}

}  // namespace fz
}  // namespace dfw

using namespace dfw;

// Debug tooling (not part of the product ABI): per-CTA clock64() timeline of the next fused launches.
extern "C" void dfw_debug_set_fused_clock_buffer(void* dev_buf) { fz::g_clk = static_cast<long long*>(dev_buf); }
// Debug tooling: progress markers (32 ints per CTA) written to pinned host memory, readable after a watchdog trap.
extern "C" void dfw_debug_set_fused_progress_buffer(void* host_buf) { fz::g_prog = static_cast<volatile int*>(host_buf); }
// Debug tooling: a (pinned host) word that receives the code of a timed-out barrier wait when the caller gave no error word.
extern "C" void dfw_debug_set_fused_error_word(void* host_word) { fz::g_errword = static_cast<int*>(host_word); }

extern "C" int dfw_fused_supported(const dfw_model* m, int precision) {
    if (check_model(m)) return 0;
    if (precision != DFW_PREC_BF16 && precision != DFW_PREC_BF16X3) return 0;
    fz::Plan pl;
    const char* why;
    return fz::make_plan(m, precision == DFW_PREC_BF16X3, pl, &why) ? 1 : 0;
}

extern "C" int dfw_pack_mlp_bf16_split(const float* W, int32_t out_dim, int32_t in_dim, void* dst_hi, void* dst_lo, void* stream) {
    DFW_REQUIRE(W && (dst_hi || dst_lo) && out_dim > 0 && in_dim > 0, DFW_E_ARG, "bad pack_mlp_bf16_split arguments");
    const int out_pad = tc::pad16(out_dim), in_pad = (in_dim + 63) / 64 * 64;
    fz::pack_split_kernel<<<148, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        W, out_dim, in_dim, out_pad, in_pad, static_cast<__nv_bfloat16*>(dst_hi), static_cast<__nv_bfloat16*>(dst_lo));
    count_launch();
    return check_launch("pack_split_kernel");
}

extern "C" int dfw_forward_fused(const dfw_model* m, const int64_t* xi, int64_t xi_stride_b, int64_t xi_stride_c,
                                 const float* xv, int64_t xv_stride_b, int64_t xv_stride_c, int64_t B, int precision,
                                 float* logits_out, float* prob_out, int32_t* err_word, void* stream) {
    dfw::NvtxRange nvtx_("FM - Component + Deep - Component, one kernel (dfw_forward_fused)");
    if (int rc = check_model(m)) return rc;
    DFW_REQUIRE(precision == DFW_PREC_BF16 || precision == DFW_PREC_BF16X3, DFW_E_ARG,
                "the fused kernel computes in bf16 or bf16x3, not precision %d", precision);
    DFW_REQUIRE(B >= 0, DFW_E_ARG, "negative batch");
    if (B == 0) return 0;
    DFW_REQUIRE(logits_out || prob_out, DFW_E_ARG, "no output requested");
    DFW_REQUIRE(m->shallow_image, DFW_E_ARG, "model has no shallow image (call dfw_pack_shallow)");
    const int F = m->field_size, K = m->embedding_size, num = m->numerical;
    DFW_REQUIRE(F - num == 0 || xi, DFW_E_ARG, "xi is NULL");
    DFW_REQUIRE(num == 0 || xv, DFW_E_ARG, "xv is NULL");
    const bool split = precision == DFW_PREC_BF16X3;
    fz::Plan pl;
    const char* why = nullptr;
    DFW_REQUIRE(fz::make_plan(m, split, pl, &why), DFW_E_UNSUPPORTED, "fused forward: %s", why ? why : "unsupported shape");
    fz::Params& p = pl.p;
    EmbedParams& e = p.ep;
    e.image = static_cast<const unsigned char*>(m->shallow_image);
    e.xi = xi; e.xi_sb = xi_stride_b; e.xi_sc = xi_stride_c;
    e.xv = xv; e.xv_sb = xv_stride_b; e.xv_sc = xv_stride_c;
    e.fm1 = m->fm_1st; e.bias = m->bias;
    e.E = nullptr; e.ldE = 0; e.Eb = nullptr; e.ldEb = 0; e.shallow = nullptr;
    e.err = (m->flags & DFW_CHECK_INDEX) ? err_word : nullptr;
    e.B = B; e.F = F; e.num = num; e.K = K; e.flags = m->flags; e.clk = nullptr;
    p.depth = m->depth; p.in_dim = F * K;
    for (int l = 0; l < m->depth; ++l) { p.widths[l] = m->widths[l]; p.bias[l] = m->b[l]; }
    p.fc = m->fc; p.logits = logits_out; p.prob = prob_out; p.B = B;
    p.num_tiles = (int)((B + fz::TS - 1) / fz::TS);
    p.err = err_word ? err_word : fz::g_errword; p.clk = fz::g_clk; p.prog = fz::g_prog;
    fz::build_uparam(m, pl.up);
    fz::Maps maps;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // the two dataset shapes BASELINE.json names get the fully unrolled dense second order
    if (split) {
        if (F == 39 && K == 10) return fz::run<true, 39, 10>(m, pl, maps, st);
        if (F == 47 && K == 10) return fz::run<true, 47, 10>(m, pl, maps, st);
        return fz::run<true, 0, 0>(m, pl, maps, st);
    }
    if (F == 39 && K == 10) return fz::run<false, 39, 10>(m, pl, maps, st);
    if (F == 47 && K == 10) return fz::run<false, 47, 10>(m, pl, maps, st);
    return fz::run<false, 0, 0>(m, pl, maps, st);
}
