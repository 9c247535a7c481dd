// The fused forward as a PERSISTENT, software-pipelined CTA pair with 64 samples per CTA (round 2; replaces the one-wave
// 32-sample pair kernel of fused_pair.cuh as the default for the dataset shapes).
//
// What changed against fused_pair.cuh and why (measured there: 49 k cycles per 32-sample tile, of which 12 k were the gather in
// front of the first MMA, 3 x 3 k the hand-offs between layers, and every N = 64 MMA ran at 40 cycles against a 32-cycle floor
// because a 128-row weight operand costs 4 KB of shared-memory reads per instruction whatever N is):
//   * 64 samples per CTA, 128 per pair: every MMA is M = 256 x N = 128 x K = 16 -- the weight bytes streamed from L2 and read
//     from shared memory are the same as for N = 64, so the instruction sits on the 64-cycle math floor and every fixed latency
//     is paid once per 128 samples.  TMEM: 2 (layer parity) x 2 (pair-tiles) x 128 columns = all 512.
//   * persistent: a pair walks pair-tiles q = cluster, cluster + n_clusters, ...; the gather group runs one tile AHEAD of the
//     tensor pipe.  There is no fp32 staging block any more (X for 64 samples + the weight ring leave no room for one): a gather
//     thread owns (sample, embedding column), loads its column of every field's row straight into registers (ten lanes cover
//     one 40-byte row, three rows per warp instruction -- whole sectors, no cp.async pieces, no fix-up pass), holds the first
//     32 samples' values while the previous tile's last layer still reads X, and writes the bf16 (hi | lo) operand the moment
//     the tensor pipe releases X (x_free = tcgen05.commit after the tile's last MMA).  First order + FwFM second order run
//     from the registers: one sample's while waiting for that release, the other's under layer 1's MMAs.
//   * one MMA issuer, one weight ring, and a HYBRID order inside a layer: [pair-tile 0: chunks 0-3][pair-tile 1: chunks 0-3]
//     [pair-tile 0: chunks 4..][pair-tile 1: chunks 4..].  Pair-tile 0 (neurons 0..255 = the next layer's chunks 0-3) completes a
//     quarter of a layer early, its epilogue runs under pair-tile 1's last MMAs, and the next layer starts on chunks 0-3 while
//     the epilogue of pair-tile 1 (chunks 4..) is still running.  X stays ONE buffer: chunks 0-3 are dead after segment 2,
//     chunks 4.. after the layer's last MMA.
//   * the dense field matrix is always the kernel-parameter (constant-bank) form; a pruned R runs through it with its zeros
//     (the pair-list walk needed the fp32 block).
//
// Roles by warp id (a warp's scheduler is warp % 4; a warpgroup -- the setmaxnreg granule -- is four consecutive warps):
//   warp  3          MMA issuer (leader only): tcgen05.mma.cta_group::2, commits multicast to both CTAs.  Scheduler 3 holds
//                    nothing else that is busy: two epilogue warps and the three producers, all mostly asleep.
//   warps 15, 19, 23 TMA producers (both CTAs), boxes round-robin: this CTA's 128-row half of every (pair-tile, chunk[, hi|lo])
//                    box into a ring of NS slots; signals the LEADER's full[slot] (cp.async.bulk.tensor ... cta_group::2)
//   warps 4-11       epilogue; thread = neuron (TMEM lane quarter = warp % 4).  Warp (hh, q4) takes 16-sample blocks 2 hh, 2 hh + 1 of
//                    the REMOTE CTA's 64 sample columns first (st.shared::cluster into the peer's X), then of its own: + bias,
//                    ReLU, split to bf16 hi/lo, 16-byte stores; last layer: x net_1_fc, warp transpose-reduction into the
//                    owning CTA's `red`.  Warp 4 also finishes the tile (shallow + deep sums, sigmoid, store).
//   warps 0-2, 12-14, 16-18, 20-22   gather group (12 warps): 3 x 2 samples x 10 columns per warp
#pragma once
#include "fused_pair.cuh"

namespace dfw {
namespace fz {
namespace wd {

constexpr int TSW = 64;                        // samples per CTA and tile
constexpr int W_MMA = 3, N_PROD = 3;          // producers: warp 3 of warpgroups 3, 4, 5 (15, 19, 23)
constexpr int W_EPI0 = 4, W_EPI = 8;
constexpr int W_G0 = W_EPI0 + W_EPI, W_G = 12;
constexpr int THREADS = 32 * (W_G0 + W_G);     // 768: 80 registers per thread
constexpr int CORE_THREADS = 32 * W_G0;
constexpr int G_THREADS_W = 32 * W_G;
constexpr int NS_MAX = 6;
constexpr int BAR_G = 1, BAR_SETUP = 3, BAR_COREW = 4;
constexpr uint32_t X_HBW = 8 * X_SBO;          // hi (or lo) part of one 64-wide K chunk: 64 samples

struct WideBars {
    uint64_t full[NS_MAX];           // leader: both CTAs' boxes of the slot have landed
    uint64_t empty[NS_MAX];          // per CTA: slot consumed (multicast commit)
    uint64_t x_ready;                // leader: layer-1 operand of the tile written in both CTAs (2 x W_G arrives)
    uint64_t x_ready03;              // leader: chunks 0-3 of it (k < 256) are written: layer 1 starts on them while the gather still
                                     // waits for the previous tile's last MMAs to release chunks 4..
    uint64_t x_free;                 // per CTA: every MMA of the tile has completed, X may be overwritten (multicast commit)
    uint64_t x_free03;               // per CTA: chunks 0-3 of X are dead already (after segment 2 of the tile's last layer)
    uint64_t shallow_ready[2];       // per CTA, [tile parity]: W_G arrives per tile.  Two barriers because its waiter (the finisher) is not
                                     // in the gather's dependency chain: with one, the gather could complete the NEXT tile's phase before a
                                     // delayed finisher tested this tile's, and a 1-bit parity wait then never returns (seen as a hang of
                                     // concurrent multi-tile launches that store to host memory)
    uint64_t act_ready[2][MAX_MT];   // leader: [layer parity][neuron tile], W_EPI arrives from the CTA that owns the tile
    uint64_t acc_full[2][2];         // per CTA: [layer parity][pair-tile] accumulators complete (multicast commit)
    uint64_t fin;                    // per CTA: the partial sums of its samples are in `red` (2 x W_EPI arrives per tile)
    uint32_t tmem_holder, pad_;
    float shallow[2][TSW];           // [tile parity]
    float red[2][4][TSW];            // [source CTA][lane quarter][sample] (one buffer: a warp writes tile t + 1 only after every
                                     // epilogue warp of the pair, the finisher included, has passed tile t -- act_ready couples them)
};

struct WideParams {
    Params p;
    int n_pair_tiles;                // ceil(ceil(B / 64) / 2)
    int ns;                          // ring slots (16 KB each)
    int grp;                         // slots per full / empty barrier: 2 = one barrier pair (and ONE tcgen05.commit) per two boxes
    uint32_t oNum;                   // (num, K) floats: the single row of every numeric field, fetched once per CTA
    int dbg;                         // -DDFW_DEBUG builds only (timeline experiments; 0 otherwise and the code is compiled out)
};

// step i of a layer's MMA sequence -> (pair-tile, chunk); see the header: [0: 0..sp)[1: 0..sp)[0: sp..kch)[1: sp..kch)
__device__ __forceinline__ void wide_step(int i, int PT, int kch, int& pt, int& c) {
    const int sp = kch < 4 ? kch : 4;
    if (PT == 1) { pt = 0; c = i; }
    else if (i < sp) { pt = 0; c = i; }
    else if (i < 2 * sp) { pt = 1; c = i - sp; }
    else {
        const int j = i - 2 * sp, rem = kch - sp;
        if (j < rem) { pt = 0; c = sp + j; } else { pt = 1; c = sp + j - rem; }
    }
}
// after which step pair-tile 0's accumulators may be handed to the epilogue: they are complete AND no MMA of the layer reads
// chunks 0-3 (which that epilogue overwrites) any more
__device__ __forceinline__ int wide_pt0_done(int PT, int kch) {
    const int sp = kch < 4 ? kch : 4;
    if (PT == 1) return kch - 1;
    return kch > sp ? 2 * sp + (kch - sp) - 1 : 2 * sp - 1;
}

// one round of the gather: this thread's column `kk` of every field of global sample b  (model/DeepFMs.py:312-337,
// model/QREmbeddingBag.py:156-174).  Exactly the reference's arithmetic: a row copy; one fp32 (*|+) for a QR table; one fp32
// multiply by Xv for a numeric field.
// low 32 bits of one index: every table has < 2^31 rows, so the low word decides (a negative int64 has a low word >= rows or,
// for -2^32 k + small, is caught by the high-word pass below when DFW_CHECK_INDEX is on)
__device__ __forceinline__ uint32_t load_index_lo(const EmbedParams& p, int64_t elem) {
    if (p.flags & DFW_XI_INT32) return (uint32_t)__ldg(reinterpret_cast<const int32_t*>(p.xi) + elem);
    return (uint32_t)__ldg(reinterpret_cast<const int32_t*>(p.xi) + 2 * elem);          // little endian
}

// the categorical indices of global sample b, raw low words (no model state needed: these loads are issued first).  Contiguous
// rows (xi_sc == 1: what forward() and the host path pass) are read with 16- / 8-byte loads at constant offsets -- the strided
// form costs ~25 instructions of 64-bit address arithmetic per index.
template <int CT>
__device__ __forceinline__ void wide_idx(const EmbedParams& ep, int64_t b, bool live, uint32_t (&idx)[CT]) {
    if (!live) {
#pragma unroll
        for (int c = 0; c < CT; ++c) idx[c] = 0u;
        return;
    }
    const bool i32 = ep.flags & DFW_XI_INT32;
    const char* row = reinterpret_cast<const char*>(ep.xi) + b * ep.xi_sb * (i32 ? 4 : 8);
    if (ep.xi_sc == 1 && !i32 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
        const uint4* r4 = reinterpret_cast<const uint4*>(row);              // two int64 per load: low words .x and .z
#pragma unroll
        for (int c = 0; c + 1 < CT; c += 2) {
            const uint4 v = __ldg(r4 + (c >> 1));
            idx[c] = v.x; idx[c + 1] = v.z;
        }
        if (CT & 1) idx[CT - 1] = __ldg(reinterpret_cast<const uint32_t*>(row) + 2 * (CT - 1));
    } else if (ep.xi_sc == 1 && i32 && (reinterpret_cast<uintptr_t>(row) & 7) == 0) {
        const uint2* r2 = reinterpret_cast<const uint2*>(row);
#pragma unroll
        for (int c = 0; c + 1 < CT; c += 2) {
            const uint2 v = __ldg(r2 + (c >> 1));
            idx[c] = v.x; idx[c + 1] = v.y;
        }
        if (CT & 1) idx[CT - 1] = __ldg(reinterpret_cast<const uint32_t*>(row) + (CT - 1));
    } else {
#pragma unroll
        for (int c = 0; c < CT; ++c) idx[c] = load_index_lo(ep, b * ep.xi_sb + c * ep.xi_sc);
    }
}

// Read-only loads with an L1 policy.  The 227 KB of shared memory leave the SM ~28 KB of L1: table rows stream through it without
// allocating, so that the few hundred bytes every sample re-reads (QR remainder tables: c rows per field) stay resident -- without
// this all 148 SMs fetch the same few L2 lines 26 times per sample and queue on those L2 banks (a QR pass took 8 k cycles
// against 2 k for plain tables).
__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_keep(const float* p) {
    float v;
    asm("ld.global.nc.L1::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

// How the categorical rows of a model are addressed -- decided once per CTA from the field descriptors, so that the per-sample
// path carries no per-field case analysis (the generic form costs ~40 instructions per field and sample: 20 k cycles per pass):
//   GM_PLAIN    every table plain and local: row = w2 + idx * K
//   GM_SHARD    plain tables, some row-sharded over a power-of-two rank count: owner = idx & (P - 1), local row = idx >> log2 P
//   GM_QR       quotient-remainder tables with a power-of-two collision count, nothing sharded: q = idx >> log2 c, r = idx & (c - 1)
//   GM_GENERIC  anything else (QR and sharding together, other divisors)
enum { GM_PLAIN = 0, GM_SHARD = 1, GM_QR = 2, GM_GENERIC = 3 };

// addressing mode of a model with special tables (QR / sharded), and THE quotient-remainder operation of a GM_QR model
template <int FT, int NUMT>
__device__ __forceinline__ int wide_gmode(const dfw_field_desc* sF, int& qop) {
    bool qr = false, sh = false, odd = false;
    qop = DFW_TABLE_PLAIN;
    for (int f = 0; f < FT; ++f) {
        const dfw_field_desc& fd = sF[f];
        const uint32_t cc = (uint32_t)fd.collisions, P = (uint32_t)fd.n_ranks;
        if (fd.qr_op != DFW_TABLE_PLAIN) {
            odd |= f < NUMT || cc == 0 || (cc & (cc - 1)) != 0 || (qr && fd.qr_op != qop);
            qr = true; qop = fd.qr_op;
        }
        if (P > 1) { sh = true; odd |= (P & (P - 1)) != 0; }
    }
    return (odd || (qr && sh)) ? GM_GENERIC : qr ? GM_QR : sh ? GM_SHARD : GM_PLAIN;
}

template <int FT, int KT, int NUMT, int MODE>
__device__ __forceinline__ void wide_rows(const EmbedParams& ep, const dfw_field_desc* sF, const float* sNum, int64_t b,
                                          bool live, int kk, int qop, uint32_t (&idx)[FT - NUMT > 0 ? FT - NUMT : 1], float (&e)[FT]) {
    constexpr int CT = FT - NUMT;
    constexpr bool PLAIN = MODE == GM_PLAIN;
    if (live) {
        if (ep.err) {       // DFW_CHECK_INDEX: the full 64-bit value must be inside the table
#pragma unroll 1
            for (int c = 0; c < CT; ++c) {
                const int64_t v = load_index(ep, b * ep.xi_sb + c * ep.xi_sc);
                if (v < 0 || v >= sF[NUMT + c].rows) atomicExch(ep.err, 1 + NUMT + c);
            }
        }
        if constexpr (MODE == GM_QR) {
            // One operation for the whole model (the reference has a single qr_operation: model/DeepFMs.py:96-108), so the combine
            // is a select, not a branch per field; a plain table of a QR model combines with the neutral element (x * 1, x + 0:
            // exact).  Two halves of 13 fields: quotient + remainder loads of all 26 fields at once are 52 results on top of the other
            // sample's 26 rows -- ptxas spilled the remainders and the operation codes, and the combine then ran as 40 dependent
            // local-memory loads (a pass took 15 k cycles against 2 k for plain tables).
            const float neutral = qop == DFW_TABLE_QR_MULT ? 1.f : 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c0 = h ? CT / 2 : 0, c1 = h ? CT : CT / 2;
                float r[(CT + 1) / 2 > 0 ? (CT + 1) / 2 : 1];
#pragma unroll
                for (int c = c0; c < c1; ++c) {
                    const dfw_field_desc& fd = sF[NUMT + c];
                    uint32_t i = idx[c];
                    if (i >= (uint32_t)fd.rows) i = 0;                // defined behaviour instead of a wild read
                    const bool isqr = fd.qr_op != DFW_TABLE_PLAIN;
                    const uint32_t cc = isqr ? (uint32_t)fd.collisions : 1u;          // a plain table of a QR model: q = i, no remainder
                    const uint32_t q = i >> (31 - __clz((int)cc));
                    e[NUMT + c] = ldg_stream(fd.w2 + (size_t)q * KT + kk);
                    r[c - c0] = neutral;
                    if (isqr) r[c - c0] = ldg_keep(fd.w2_r + (i & (cc - 1)) * KT + kk);
                }
#pragma unroll
                for (int c = c0; c < c1; ++c)      // quotient row (x|+) remainder row (model/QREmbeddingBag.py:169-172)
                    e[NUMT + c] = qop == DFW_TABLE_QR_MULT ? e[NUMT + c] * r[c - c0] : e[NUMT + c] + r[c - c0];
            }
        } else {
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                const dfw_field_desc& fd = sF[NUMT + c];
                if (idx[c] >= (uint32_t)fd.rows) idx[c] = 0;          // defined behaviour instead of a wild read
                const float* src;
                if constexpr (MODE == GM_PLAIN) {
                    src = fd.w2 + (size_t)idx[c] * KT;
                } else if constexpr (MODE == GM_SHARD) {
                    // branch-free: a replicated table is "sharded over one rank" whose pointer is w2
                    const uint32_t P = fd.n_ranks > 1 ? (uint32_t)fd.n_ranks : 1u;
                    const float* const* tbl = fd.n_ranks > 1 ? fd.w2_shard : &fd.w2;
                    src = tbl[idx[c] & (P - 1)] + (size_t)(idx[c] >> (31 - __clz((int)P))) * KT;
                } else {
                    src = locate_row(fd, (int32_t)idx[c], KT);
                }
                e[NUMT + c] = __ldg(src + kk);
            }
            if constexpr (MODE == GM_GENERIC) {
                // quotient row (x|+) remainder row (model/QREmbeddingBag.py:169-172); the c-row remainder tables are L1-resident
#pragma unroll
                for (int c = 0; c < CT; ++c) {
                    const dfw_field_desc& fd = sF[NUMT + c];
                    const int op = fd.qr_op;
                    if (op != DFW_TABLE_PLAIN) {
                        const uint32_t cc = (uint32_t)fd.collisions;
                        const float r = __ldg(fd.w2_r + (idx[c] - div_small(idx[c], cc) * cc) * KT + kk);
                        e[NUMT + c] = op == DFW_TABLE_QR_MULT ? e[NUMT + c] * r : e[NUMT + c] + r;
                    }
                }
            }
        }
        const float* xrow = ep.xv + b * ep.xv_sb;
        const int64_t xsc = ep.xv_sc;
#pragma unroll
        for (int f = 0; f < NUMT; ++f) {
            float v = sNum[f * KT + kk];
            if constexpr (!PLAIN && MODE != GM_SHARD) {
                const int op = sF[f].qr_op;
                if (op == DFW_TABLE_QR_MULT) v *= __ldg(sF[f].w2_r + kk);
                else if (op == DFW_TABLE_QR_ADD) v += __ldg(sF[f].w2_r + kk);
            }
            e[f] = v * (xsc == 1 ? __ldg(xrow + f) : __ldg(xrow + f * xsc));
        }
    } else {
#pragma unroll
        for (int f = 0; f < FT; ++f) e[f] = 0.f;
    }
}

// first-order table terms of this thread's fields f = kk, kk + K, ... (use_fwlw = 0; model/DeepFMs.py:300-309, 445-450)
template <int FT, int KT, int NUMT>
__device__ __forceinline__ float wide_first(const EmbedParams& ep, const dfw_field_desc* sF, int64_t b, bool live, int kk) {
    float acc = 0.f;
    if (!live) return acc;
#pragma unroll 4
    for (int f = kk; f < FT; f += KT) {
        const dfw_field_desc& fd = sF[f];
        int64_t iv = 0;
        if (f >= NUMT) {
            iv = load_index(ep, b * ep.xi_sb + (f - NUMT) * ep.xi_sc);
            if (iv < 0 || iv >= fd.rows) iv = 0;
        }
        const int32_t idx = (int32_t)iv;
        float v;
        if (fd.qr1_op != DFW_TABLE_PLAIN) {
            const uint32_t c = (uint32_t)fd.collisions;
            const uint32_t q = (uint32_t)idx / c, rr = (uint32_t)idx - q * c;
            const float a = __ldg(fd.w1 + q), bq = __ldg(fd.w1_r + rr);
            v = fd.qr1_op == DFW_TABLE_QR_MULT ? a * bq : a + bq;
        } else {
            v = __ldg(fd.w1 + idx);
        }
        if (f < NUMT) v *= ep.xv[b * ep.xv_sb + f * ep.xv_sc];
        if (ep.flags & DFW_USE_LW) v *= __ldg(ep.fm1 + f);
        acc += v;
    }
    return acc;
}

// bf16 (hi | lo) operand columns of fields [F0, F1) of the ADJACENT tile-local samples s (even) and s + 1:  element (k = f * KT + kk,
// s) lives at (k / 64) * CH + (s / 8) * X_SBO + (k % 64) * 16 + (s % 8) * 2  (+ X_HBW for the lo part), so the two samples share a
// 32-bit word: one cvt.rn.bf16x2 and one store per field and part.  With f unrolled everything but "does f * KT + kk cross into
// the next 64-wide chunk" is a compile-time constant.
template <bool SPLIT, int FT, int KT, int F0, int F1>
__device__ __forceinline__ void wide_write2(unsigned char* sX, int s, int kk, const float (&ea)[FT], const float (&eb)[FT]) {
    constexpr int CH = (SPLIT ? 2 : 1) * (int)X_HBW, JUMP = CH - 64 * 16;
    unsigned char* const xbase = sX + (s >> 3) * X_SBO + (s & 7) * 2 + kk * 16;
#pragma unroll
    for (int f = F0; f < F1; ++f) {
        const int c0 = f * KT, lo = c0 & 63;
        int off = c0 * 16 + (c0 >> 6) * JUMP;
        if (lo + KT > 64) off += (kk >= 64 - lo) ? JUMP : 0;
        unsigned char* dst = xbase + off;
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(ea[f], eb[f]);        // .x = sample s (low half)
        const uint32_t wh = *reinterpret_cast<const uint32_t*>(&h2);
        *reinterpret_cast<uint32_t*>(dst) = wh;
        if constexpr (SPLIT) {
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(ea[f] - __uint_as_float(wh << 16), eb[f] - __uint_as_float(wh & 0xffff0000u));
            *reinterpret_cast<uint32_t*>(dst + X_HBW) = *reinterpret_cast<const uint32_t*>(&l2);
        }
    }
}

// first order (fwlw: <E_f, fwfm_linear_f>, model/DeepFMs.py:338-347) + FwFM / FM second order (model/DeepFMs.py:351-367) of one
// (sample, column) from registers; every U_ij is a constant-bank operand of its FFMA, four independent chains per column j
template <int FT, int KT>
__device__ __forceinline__ float wide_interact(const float (&e)[FT], const UParam& up, const float* sWl, int kk, bool fwlw, float first) {
    float acc = first;
    if (fwlw) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int f = 0; f < FT; ++f) {
            if (f & 1) a1 = fmaf(e[f], sWl[f * KT + kk], a1); else a0 = fmaf(e[f], sWl[f * KT + kk], a0);
        }
        acc = a0 + a1;
    }
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
    return acc + (s0 + s1);
}

template <bool SPLIT, int FT, int KT, int NUMT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
fused_wide_kernel(const __grid_constant__ Maps maps, const __grid_constant__ UParam up, const WideParams wp) {
    static_assert(FT > 0 && KT > 0 && 3 * KT <= 32 && FT >= NUMT, "built for the dataset shapes");
    const Params& p = wp.p;
    constexpr int H = SPLIT ? 2 : 1;
    constexpr int CH = H * (int)X_HBW;                  // one K chunk of the activation buffer: hi [| lo] of 64 samples
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment (128-byte swizzle atoms) as an OFFSET into the shared array: rounding the pointer itself through uintptr_t
    // loses the address space, and every access through `base` then compiles to generic LD.E / ST.E instead of LDS / STS
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sX = base;
    unsigned char* sW = base + p.oRing;
    WideBars* bars = reinterpret_cast<WideBars*>(base + p.oMisc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = p.depth;
    const uint32_t NS = (uint32_t)wp.ns;
    const uint32_t rank = cluster_ctarank();            // 0 = leader
    const bool leader = rank == 0;
    const int n_clusters = (int)(gridDim.x >> 1), cluster_id = (int)(blockIdx.x >> 1);
    const int n_iter = wp.n_pair_tiles > cluster_id ? (wp.n_pair_tiles - cluster_id + n_clusters - 1) / n_clusters : 0;
    if (p.clk && threadIdx.x == 0) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.clk[blockIdx.x * FZ_NCLK + 28] = clock64(); p.clk[blockIdx.x * FZ_NCLK + 29] = (long long)t;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < NS_MAX; ++s) { mbar_init(&bars->full[s], (uint32_t)wp.grp); mbar_init(&bars->empty[s], 1); }
        mbar_init(&bars->x_ready, 2 * W_G);
        mbar_init(&bars->x_ready03, 2 * W_G);
        mbar_init(&bars->x_free, 1);
        mbar_init(&bars->x_free03, 1);
        mbar_init(&bars->shallow_ready[0], W_G);
        mbar_init(&bars->shallow_ready[1], W_G);
        mbar_init(&bars->fin, 2 * W_EPI);          // every epilogue warp of the pair holds partial sums of both CTAs' samples
        for (int b = 0; b < 2; ++b) {
            for (int m = 0; m < MAX_MT; ++m) mbar_init(&bars->act_ready[b][m], W_EPI);
            for (int j = 0; j < 2; ++j) mbar_init(&bars->acc_full[b][j], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int l = 0; l < L; ++l)
            for (int h = 0; h < H; ++h) tma_prefetch_desc(&maps.w[l][h][0]);
    }
    // Roles by warp id.  A warp's scheduler is warp % 4, and a warpgroup (setmaxnreg granule) is four consecutive warps, i.e. one
    // per scheduler.  Scheduler 3 is kept for the MMA issuer: besides it only two epilogue warps and the three TMA producers
    // live there, all of which sleep most of the time; the twelve gather warps are the other three members of warpgroups 0, 3, 4, 5.
    //   3              MMA issuer            0-2, 12-14, 16-18, 20-22   gather group (gw = 0..11)
    //   4-11           epilogue              15, 19, 23                 TMA producers
    const bool epi_warp = warp >= W_EPI0 && warp < W_G0;
    const bool prod_warp = warp >= W_G0 && (warp & 3) == 3;
    const bool gather_warp = !epi_warp && !prod_warp && warp != W_MMA;
    uint32_t tmem_base = 0;
    if (gather_warp) {
        cluster_arrive();                      // non-blocking: the gather starts on the batch at once
    } else {
        if (warp == W_MMA) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_holder)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        asm volatile("bar.sync %0, %1;" ::"n"(BAR_COREW), "n"(CORE_THREADS) : "memory");
        cluster_arrive(); cluster_wait();      // both CTAs' barriers and TMEM exist before anything crosses over
        tc_fence_after();
        tmem_base = bars->tmem_holder;
        asm volatile("bar.arrive %0, %1;" ::"n"(BAR_SETUP), "n"(THREADS) : "memory");
    }

    auto layer_k = [&](int l) { return pad16(l == 0 ? p.in_dim : p.widths[l - 1]); };
    auto layer_n = [&](int l) { return pad16(p.widths[l]); };

    auto gather_role = [&]() __attribute__((always_inline)) {
        // ================================================================= gather group: one tile ahead of the tensor pipe
        const int gw = warp < W_MMA ? warp : 3 + (warp - W_G0) - ((warp - W_G0) >> 2);      // 0..11
        const int gtid = 32 * gw + lane;
        constexpr int FK = FT * KT, Kp = (FK + 15) & ~15;
        const ImgLayout IL = img_layout(FT, KT);
        unsigned char* sImg = base + p.oImg;
        float* sPart = reinterpret_cast<float*>(base + p.oPart);
        float* sNum = reinterpret_cast<float*>(base + wp.oNum);
        const EmbedParams& ep = p.ep;
        // model state once per CTA: header, fwlw weights and field descriptors of the shallow image (U and the pair list stay in
        // global memory: the field matrix is a kernel parameter here) and the numeric fields' single rows
        constexpr uint32_t W_WL = 16, W_FIELDS = W_WL + (uint32_t)up16(sizeof(float) * FT * KT);
        if (gtid == 0) cp_async16(sImg, ep.image + IL.oHdr);
        for (uint32_t i = gtid; i < (uint32_t)(up16(sizeof(float) * FT * KT) >> 4); i += G_THREADS_W)
            cp_async16(sImg + W_WL + 16 * i, ep.image + IL.oWl + 16 * i);
        for (uint32_t i = gtid; i < (uint32_t)(up16(sizeof(dfw_field_desc) * FT) >> 4); i += G_THREADS_W)
            cp_async16(sImg + W_FIELDS + 16 * i, ep.image + IL.oFields + 16 * i);
        const ImgHeader* hdr = reinterpret_cast<const ImgHeader*>(sImg);
        const dfw_field_desc* sF = reinterpret_cast<const dfw_field_desc*>(sImg + W_FIELDS);
        const float* sWl = reinterpret_cast<const float*>(sImg + W_WL);

        // A thread owns column kk of TWO ADJACENT samples (they share a 32-bit word of the operand); a warp holds three such pairs:
        // 32 pairs per tile on 12 x 3 places.
        constexpr int CT = FT - NUMT > 0 ? FT - NUMT : 1;
        const int sl = lane / KT, kk = lane - sl * KT;
        const int slot = 3 * gw + sl;
        const bool owner = sl < 3 && slot < 32;
        const int sa = 2 * slot;                                // tile-local samples sa, sa + 1
        const bool fwlw = ep.flags & DFW_USE_FWLW;
        constexpr int FSPLIT = 256 / KT;                        // fields [0, FSPLIT) lie entirely in chunks 0-3 of X (k < 256)
        constexpr int KB = 256 - FSPLIT * KT;                   // columns of field FSPLIT that still do
        static_assert(FSPLIT < FT, "layer 1 has more than four K chunks for the dataset shapes");
        bool joined = false;
        int gmode = GM_PLAIN, qop = DFW_TABLE_PLAIN;           // qop: THE operation of a QR model (GM_QR needs a single one)
        for (int it = 0; it < n_iter; ++it) {
            const int par = it & 1;
            const long long q = cluster_id + (long long)it * n_clusters;
            const int64_t b0 = (2 * q + rank) * TSW;
            const int64_t ba = b0 + sa, bb = ba + 1;
            const bool livea = owner && ba < ep.B, liveb = owner && bb < ep.B;
            // ---- indices, then rows, of this thread's two samples into registers, all of it under the previous tile's MLP (first
            //      tile: the first index loads fly under the model-state copy)
            float e0[FT], e1[FT];
#pragma unroll 1
            for (int r = 1; r >= 0; --r) {          // one copy of the load code; 26 + 26 loads in flight per pass (two passes in flight
                                                    // at once cost more in spilled registers than they gain: 24 k against 12 k cycles)
                uint32_t ix[CT];
                wide_idx<CT>(ep, r ? bb : ba, r ? liveb : livea, ix);
                if (it == 0 && r == 1) {
                    cp_async_wait_all();
                    group_sync<BAR_G>(G_THREADS_W);                 // header, fwlw weights, descriptors are in shared memory
                    gmode = GM_PLAIN;
                    if (hdr->any_special) {
                        gmode = wide_gmode<FT, NUMT>(sF, qop);
#ifdef DFW_DEBUG
                        if (wp.dbg & 32) gmode = GM_GENERIC;          // timing experiments: force the generic addressing form
                        if (gtid == 0 && blockIdx.x == 0 && (wp.dbg & 64)) printf("gmode %d qop %d\n", gmode, qop);
#endif
                    }
                    for (int i = gtid; i < NUMT * KT; i += G_THREADS_W) {
                        const int f = i / KT, k = i - f * KT;
                        sNum[i] = __ldg((gmode == GM_PLAIN ? sF[f].w2 : locate_row(sF[f], 0, KT)) + k);
                    }
                    group_sync<BAR_G>(G_THREADS_W);
                    if (gtid == 0) FZ_CLK(102);
                }
                const int64_t bl = r ? bb : ba;
                const bool lv = r ? liveb : livea;
                if (gmode == GM_PLAIN) wide_rows<FT, KT, NUMT, GM_PLAIN>(ep, sF, sNum, bl, lv, kk, qop, ix, e0);
                else if (gmode == GM_SHARD) wide_rows<FT, KT, NUMT, GM_SHARD>(ep, sF, sNum, bl, lv, kk, qop, ix, e0);
                else if (gmode == GM_QR) wide_rows<FT, KT, NUMT, GM_QR>(ep, sF, sNum, bl, lv, kk, qop, ix, e0);
                else wide_rows<FT, KT, NUMT, GM_GENERIC>(ep, sF, sNum, bl, lv, kk, qop, ix, e0);
                if (r) {
#pragma unroll
                    for (int f = 0; f < FT; ++f) e1[f] = e0[f];
                    if (gtid == 0 && it < 4) FZ_CLK(101 + 8 * it);
                }
            }
            if (gtid == 0 && it < 4) FZ_CLK(96 + 8 * it);
            if (!joined) {
                asm volatile("bar.sync %0, %1;" ::"n"(BAR_SETUP), "n"(THREADS) : "memory");
                cluster_wait();                             // set-up of both CTAs complete: barriers may be used
                joined = true;
            }
            // ---- first order + FwFM second order of both samples from the registers (ONE copy of the code), and in between the
            //      operand write.  From the second tile on sample a's interaction runs FIRST, in the time the gather group would
            //      otherwise wait for the tensor pipe to release X (rows are in registers ~8 k cycles before x_free03): only sample
            //      b's half is then left to run under layer 1, and it ends before layer 1's epilogue starts (next to the
            //      interaction that epilogue took 6.2 + 5.0 k cycles per pair-tile instead of 5.0 + 3.0 k).  The first tile of a
            //      launch has nothing to hide under and writes first.
#pragma unroll 1
            for (int r = 0; r < 2; ++r) {
                if (r == (it == 0 ? 0 : 1)) {
                    // the tensor pipe releases X in two steps: chunks 0-3 after segment 2 of the previous tile's last layer, the
                    // rest after its last MMA
                    if (it > 0) mbar_wait(&bars->x_free03, (uint32_t)((it - 1) & 1), p.err, 40);
                    if (gtid == 0 && it < 4) FZ_CLK(100 + 8 * it);
                    // columns k < 256 (chunks 0-3): fields [0, FSPLIT) and the first KB columns of field FSPLIT, which straddles
                    // the boundary.  Layer 1's first two segments read nothing else, so they get their own barrier: the issuer goes
                    // from the previous tile's last MMA straight into them, and the rest of X is written underneath.
                    if (owner) {
                        wide_write2<SPLIT, FT, KT, 0, FSPLIT>(sX, sa, kk, e0, e1);
                        if constexpr (KB > 0) { if (kk < KB) wide_write2<SPLIT, FT, KT, FSPLIT, FSPLIT + 1>(sX, sa, kk, e0, e1); }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote_cta(mapa_u32(smem_u32(&bars->x_ready03), 0));
                    if (it > 0) mbar_wait(&bars->x_free, (uint32_t)((it - 1) & 1), p.err, 41);
                    if (gtid == 0 && it < 4) FZ_CLK(97 + 8 * it);
                    if (owner) {
                        if constexpr (KB > 0) { if (kk >= KB) wide_write2<SPLIT, FT, KT, FSPLIT, FSPLIT + 1>(sX, sa, kk, e0, e1); }
                        wide_write2<SPLIT, FT, KT, FSPLIT + (KB > 0 ? 1 : 0), FT>(sX, sa, kk, e0, e1);
                    }
                    // K padding columns [F*K, Kp) of all 64 samples are zero
                    for (int i = gtid; i < TSW * (Kp - FK); i += G_THREADS_W) {
                        const int sp = i / (Kp - FK), col = FK + (i - sp * (Kp - FK));
                        unsigned char* dst = sX + (size_t)(col >> 6) * CH + (sp >> 3) * X_SBO + (col & 63) * 16 + (sp & 7) * 2;
                        *reinterpret_cast<__nv_bfloat16*>(dst) = __float2bfloat16_rn(0.f);
                        if constexpr (SPLIT) *reinterpret_cast<__nv_bfloat16*>(dst + X_HBW) = __float2bfloat16_rn(0.f);
                    }
                    fence_async_smem();                         // these stores are local
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote_cta(mapa_u32(smem_u32(&bars->x_ready), 0));
                    if (gtid == 0 && it < 4) FZ_CLK(98 + 8 * it);
                }
                if (r == 1) {
#pragma unroll
                    for (int f = 0; f < FT; ++f) e0[f] = e1[f];
                }
#ifdef DFW_DEBUG
                const bool dbg_skip = (wp.dbg & 2) || ((wp.dbg & 1) && (warp & 3) == 1);       // WRONG logits: timeline experiments only
#else
                constexpr bool dbg_skip = false;
#endif
                if (owner && !dbg_skip) {
                    const float first = fwlw ? 0.f : wide_first<FT, KT, NUMT>(ep, sF, r ? bb : ba, r ? liveb : livea, kk);
                    sPart[kk * TSW + sa + r] = wide_interact<FT, KT>(e0, up, sWl, kk, fwlw, first);
                }
            }
            group_sync<BAR_G>(G_THREADS_W);
            if (gtid < TSW) {
                float tot = 0.f;
#pragma unroll 1
                for (int k = 0; k < KT; ++k) tot += sPart[k * TSW + gtid];
                bars->shallow[par][gtid] = tot + __ldg(ep.bias);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->shallow_ready[par]);
            if (gtid == 0 && it < 4) FZ_CLK(99 + 8 * it);
            group_sync<BAR_G>(G_THREADS_W);             // sPart is free for the next tile
        }
        if (!joined) {      // a pair without any tile still takes part in the set-up handshake
            asm volatile("bar.sync %0, %1;" ::"n"(BAR_SETUP), "n"(THREADS) : "memory");
            cluster_wait();
        }
    };

    // Register budget (64 K per SM, 768 threads launched at 80; the pool is the CTA's own launch allocation 768 x 80 = 61,440 -- a
    // larger total never gets its registers and the kernel hangs): warpgroups 1-2 (epilogue) give up 16 registers per thread,
    // warpgroups 0, 3, 4, 5 (gather + issuer / producer) take 8: 64 x 256 + 88 x 512 = 61,440.  setmaxnreg sits at the head of the
    // two branches so that ptxas allocates each accordingly.
    if (!epi_warp) {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 88;");
      if (prod_warp) {
        // ================================================================= TMA producers (both CTAs): warps 15, 19, 23
        // One warp gets a box out every ~390 cycles whatever its size or the ring depth (scripts/ubench/tma_ingest.cu: the
        // try_wait -> expect_tx -> cp.async.bulk.tensor chain, not L2 latency or bytes), and a step of the issuer consumes two
        // boxes per 768 cycles (bf16: one per 256).  So the boxes, numbered in consumption order, are dealt round-robin to three
        // producer warps; box n lives in slot n mod NS.
        const int pj = (warp - W_G0) >> 2;                        // 0, 1, 2
        const uint32_t GRP = (uint32_t)wp.grp;                    // slots s, s + 1 share full / empty barrier s / GRP
        uint32_t n = 0;                                           // box counter
        RingPos rp{0, 0};
        for (int it = 0; it < n_iter; ++it) {
            for (int l = 0; l < L; ++l) {
                const int kch = (layer_k(l) + KCH - 1) / KCH, npad = layer_n(l), MT = n_mtiles(npad), PT = (MT + 1) / 2;
                const int nstep = PT * kch;
                for (int i = 0; i < nstep; ++i) {
                    int pt, c;
                    wide_step(i, PT, kch, pt, c);
                    const int row0 = (2 * pt + (int)rank) * 128;          // rows past the matrix are zero-filled by TMA
#pragma unroll
                    for (int h = 0; h < H; ++h, ++n, rp.next(NS)) {
                        if ((int)(n % N_PROD) != pj) continue;
                        const uint32_t g = rp.s / GRP;
                        mbar_wait(&bars->empty[g], rp.ph ^ 1, p.err, 12);
                        const uint32_t f0 = mapa_u32(smem_u32(&bars->full[g]), 0);
                        if (elect_one()) {
#ifdef DFW_DEBUG
                            if (wp.dbg & 8) {        // dbg 8: no weight traffic at all (WRONG results): the barrier just completes
                                if (leader) mbar_arrive(&bars->full[g]);
                            } else
#endif
                            {
                                if (leader) mbar_expect_tx(&bars->full[g], 2u * STAGE_BYTES);        // both CTAs' boxes
                                tma_load_2d_2cta(sW + (size_t)rp.s * STAGE_BYTES, &maps.w[l][h][0], f0, c * KCH, row0);
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
      } else if (warp == W_MMA) {
        // ================================================================= MMA issuer (leader only)
        // Measured (scripts/ubench/mma_rate.cu): back-to-back N = 128 MMAs run at the 64-cycle math floor, and every
        // tcgen05.commit in the stream costs the pipe ~100-190 cycles.  Interleaving the next step's barrier probes between the MMA
        // groups made things worse (ptxas then spreads descriptor moves between the UTCHMMAs: 123 cycles per MMA), so the step
        // stays a plain block: probes, then the MMAs back to back inside ONE elect-guarded region, then the commits.
        // It shares its scheduler with two epilogue and three gather warps and gets about a quarter of the issue slots while
        // they run (measured: layer 1, under the FwFM interaction, took 26 k cycles against 14 k; with the gather warps of this
        // scheduler idled, 16 k), so the loop is kept lean: explicit segment loops instead of index arithmetic, descriptors advanced
        // by adds, ONE barrier probe and ONE slot-release commit per step when the ring is grouped (grp = 2).
        if (leader) {
            const uint32_t sW_u32 = smem_u32(sW), sX_u32 = smem_u32(sX);
#ifdef DFW_DEBUG
            const uint32_t idesc = make_idesc(256, (wp.dbg & 4) ? TSW : 2 * TSW) | (1u << 16);      // dbg 4: half-width MMAs (WRONG results)
#else
            const uint32_t idesc = make_idesc(256, 2 * TSW) | (1u << 16);     // B is MN-major
#endif
            const uint64_t a_base = make_desc_sw128(sW_u32), b_base = make_desc_mn(sX_u32);
            constexpr uint64_t A_SLOT = STAGE_BYTES >> 4, B_CH = (uint64_t)CH >> 4, B_LO = X_HBW >> 4;
            const bool grouped = wp.grp == 2;
            RingPos rp{0, 0};
            uint32_t act_bits = 0;
            int gl = 0;                                                      // global layer counter: accumulator parity
            for (int it = 0; it < n_iter; ++it) {
                for (int l = 0; l < L; ++l, ++gl) {
                    const int buf = gl & 1;
                    const int K = layer_k(l), kch = (K + KCH - 1) / KCH, PT = (n_mtiles(layer_n(l)) + 1) / 2;
                    const int sp = kch < 4 ? kch : 4;
                    const bool lastl = l == L - 1;
#pragma unroll 1
                    for (int seg = 0; seg < 4; ++seg) {
                        const int pt = seg & 1;
                        const int cb = seg < 2 ? 0 : sp, ce = seg < 2 ? sp : kch;
                        if (pt >= PT || cb >= ce) continue;
                        const uint32_t dcol = tmem_base + (uint32_t)(buf * 256 + pt * 128);
#pragma unroll 1
                        for (int c = cb; c < ce; ++c) {
                            if (pt == 0) {
                                if (l == 0) {
                                    if (c == 0) {
                                        if (lane == 0 && it < 4) FZ_CLK(32 + 8 * it);
                                        mbar_wait_cluster(&bars->x_ready03, (uint32_t)(it & 1), p.err, 21);
                                        if (kch <= sp) mbar_wait_cluster(&bars->x_ready, (uint32_t)(it & 1), p.err, 23);
                                        if (lane == 0 && it < 4) FZ_CLK(33 + 8 * it);
                                    } else if (c == sp) {
                                        mbar_wait_cluster(&bars->x_ready, (uint32_t)(it & 1), p.err, 23);     // chunks 4..
                                    }
                                } else if ((c & 1) == 0) {
                                    const int g = c >> 1, bit = buf * MAX_MT + g;
                                    mbar_wait_cluster(&bars->act_ready[buf][g], (act_bits >> bit) & 1u, p.err, 22);
                                    act_bits ^= 1u << bit;
                                }
                            }
                            const int ks = min(4, (K - c * KCH) / 16);
                            const bool endc = c == kch - 1;
                            RingPos s0 = rp, s1 = rp;
                            if (SPLIT) s1.next(NS);
                            const uint32_t g0 = grouped ? (s0.s >> 1) : s0.s, g1 = grouped ? (s1.s >> 1) : s1.s;
                            // grouped + SPLIT: hi and lo box of the step share one barrier; grouped bf16: two steps share one
                            const bool probe0 = !grouped || (s0.s & 1) == 0, probe1 = SPLIT && !grouped;
                            const bool k0 = probe0 ? mbar_try(&bars->full[g0], s0.ph) : true;
                            const bool k1 = probe1 ? mbar_try(&bars->full[g1], s1.ph) : true;
                            if (!k0) mbar_wait_cluster(&bars->full[g0], s0.ph, p.err, 24);
                            if (!k1) mbar_wait_cluster(&bars->full[g1], s1.ph, p.err, 25);
                            tc_fence_after();
                            const uint64_t ah = a_base + s0.s * A_SLOT, al = a_base + s1.s * A_SLOT;
                            const uint64_t bhi = b_base + (uint64_t)c * B_CH, blo = bhi + B_LO;
                            // pair-tile 0 may go to the epilogue once it is complete AND nothing reads chunks 0-3 any more
                            const bool acc0_done = endc && (pt == 0 ? (kch > sp || PT == 1) : kch <= sp);
                            if (elect_one()) {
                                umma_bf16_2cta(dcol, ah, bhi, idesc, c ? 1u : 0u);
                                if (ks > 1) umma_bf16_2cta(dcol, ah + 2, bhi + 16, idesc, 1u);
                                if (ks > 2) umma_bf16_2cta(dcol, ah + 4, bhi + 32, idesc, 1u);
                                if (ks > 3) umma_bf16_2cta(dcol, ah + 6, bhi + 48, idesc, 1u);
                                if (SPLIT) {
                                    umma_bf16_2cta(dcol, ah, blo, idesc, 1u);                 // W_hi X_lo
                                    if (ks > 1) umma_bf16_2cta(dcol, ah + 2, blo + 16, idesc, 1u);
                                    if (ks > 2) umma_bf16_2cta(dcol, ah + 4, blo + 32, idesc, 1u);
                                    if (ks > 3) umma_bf16_2cta(dcol, ah + 6, blo + 48, idesc, 1u);
                                    if (!grouped) umma_commit_2cta(&bars->empty[g0]);
                                    umma_bf16_2cta(dcol, al, bhi, idesc, 1u);                 // W_lo X_hi
                                    if (ks > 1) umma_bf16_2cta(dcol, al + 2, bhi + 16, idesc, 1u);
                                    if (ks > 2) umma_bf16_2cta(dcol, al + 4, bhi + 32, idesc, 1u);
                                    if (ks > 3) umma_bf16_2cta(dcol, al + 6, bhi + 48, idesc, 1u);
                                    umma_commit_2cta(&bars->empty[g1]);
                                } else if (!grouped || (s0.s & 1)) {
                                    umma_commit_2cta(&bars->empty[g0]);
                                }
                                if (acc0_done) umma_commit_2cta(&bars->acc_full[buf][0]);
                                if (pt == 1 && endc) umma_commit_2cta(&bars->acc_full[buf][1]);
                                if (lastl && c == sp - 1 && pt == PT - 1) umma_commit_2cta(&bars->x_free03);
                                if (lastl && endc && pt == PT - 1) umma_commit_2cta(&bars->x_free);
                            }
                            __syncwarp();
                            rp = s1; rp.next(NS);
                        }
                    }
                    if (lane == 0 && it < 4 && l < 4) FZ_CLK(34 + 8 * it + l);
                }
            }
        }
      } else {
        gather_role();
      }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        // ================================================================= epilogue
        // A CTA's accumulators are its 128 neurons x the 128 samples of BOTH CTAs; the 64 columns of the other CTA's samples go to
        // that CTA's operand buffer over DSMEM (32 KB per pair-tile and direction, while the tensor pipe reads the same shared
        // memory at ~80 B/cycle).  Every warp takes two 16-sample blocks of the REMOTE set first and two of the local set after
        // them, so that all eight warps feed the link from the start.  Measured per pair-tile: first tcgen05.ld 0.65 k cycles,
        // four blocks 1.7 k, proxy fence 0.1 k after local stores but 1.6 k after remote ones: the drain of the remote stores is what
        // bounds a pair-tile's epilogue at ~4 k cycles (one set per warp: 4.1 k; this order: 2.6-4.5 k).
        const int hh = (warp - W_EPI0) >> 2, q4 = warp & 3;          // hh: blocks 2 hh, 2 hh + 1 of either set
        const int row = q4 * 32 + lane;
        const uint32_t taddr_q = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const uint32_t xdst_loc = mapa_u32(smem_u32(sX), rank), xdst_rem = mapa_u32(smem_u32(sX), rank ^ 1u);
        uint32_t acc_bits = 0;
        int gl = 0;
        for (int it = 0; it < n_iter; ++it) {
            const int par = it & 1;
            float zsum[4] = {0.f, 0.f, 0.f, 0.f};         // unit u; lane pair (2 s, 2 s + 1): sample 16 (2 hh + (u & 1)) + s of its set, summed over this warp's neurons
            for (int l = 0; l < L; ++l, ++gl) {
                const int buf = gl & 1, N = p.widths[l], npad = layer_n(l), MT = n_mtiles(npad), PT = (MT + 1) / 2;
                const bool last = (l == L - 1);
                float bb2[2], ff2[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int n = (2 * j + (int)rank) * 128 + row;
                    const bool real = n < npad && n < N;
                    bb2[j] = real ? __ldg(p.bias[l] + n) : 0.f;
                    ff2[j] = (real && last) ? __ldg(p.fc + n) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (j >= PT) break;
                    const int bit = buf * 2 + j;
                    mbar_wait(&bars->acc_full[buf][j], (acc_bits >> bit) & 1u, p.err, 31);
                    acc_bits ^= 1u << bit;
                    tc_fence_after();
                    if (threadIdx.x == 32 * W_EPI0 && it < 2 && l < 4) FZ_CLK(64 + 16 * it + 4 * l + 2 * j);
                    const int t = 2 * j + (int)rank;                    // this CTA's neuron tile of pair-tile j
                    const int n = t * 128 + row;
                    const int rows_valid = max(0, min(128, npad - t * 128));
                    const float bb = bb2[j], ff = ff2[j];
                    if (q4 * 32 < rows_valid) {
                        const uint32_t xoff = (uint32_t)((n >> 6) * CH + (n & 63) * 16);
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const uint32_t set = u < 2 ? (rank ^ 1u) : rank;             // remote first
                            const int blk = 2 * hh + (u & 1);
                            const uint32_t xc = (u < 2 ? xdst_rem : xdst_loc) + xoff;
                            uint32_t d[16];
                            tmem_ld16(taddr_q + (uint32_t)(buf * 256 + j * 128) + (uint32_t)TSW * set + (uint32_t)(16 * blk), d);
                            tmem_ld_wait();
                            if (last) {
                                float v[16];
#pragma unroll
                                for (int s = 0; s < 16; ++s) v[s] = fmaxf(__uint_as_float(d[s]) + bb, 0.f) * ff;
                                // transpose-reduce over the warp's 32 neurons: lane pair (2 s, 2 s + 1) ends with sample s
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
                                zsum[u] += v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
                            } else if (row < rows_valid) {
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
                                    const uint32_t dst = xc + (uint32_t)(2 * blk + g) * X_SBO;
                                    st_cluster_v4(dst, wh[0], wh[1], wh[2], wh[3]);
                                    if constexpr (SPLIT) st_cluster_v4(dst + X_HBW, wl[0], wl[1], wl[2], wl[3]);
                                }
                            }
                        }
                    }
                    if (!last && t * 128 < npad) {
                        asm volatile("fence.proxy.async.shared::cluster;" ::: "memory");
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote_cta(mapa_u32(smem_u32(&bars->act_ready[(gl + 1) & 1][t]), 0));
                    }
                    if (threadIdx.x == 32 * W_EPI0 && it < 2 && l < 4) FZ_CLK(65 + 16 * it + 4 * l + 2 * j);
                }
            }
            // this warp's partial sums of its blocks of either set -> the CTA that owns the samples
            if ((lane & 1) == 0) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t set = u < 2 ? (rank ^ 1u) : rank;
                    st_cluster_f32(mapa_u32(smem_u32(&bars->red[rank][q4][16 * (2 * hh + (u & 1)) + (lane >> 1)]), set), zsum[u]);
                }
            }
            asm volatile("fence.acq_rel.cluster;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane < 2) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->fin), (uint32_t)lane));
            if (warp == W_EPI0) {
                // this CTA's samples of the tile: shallow part + the partial sums of both CTAs' neuron tiles
                mbar_wait_cluster(&bars->fin, (uint32_t)par, p.err, 34);
                mbar_wait(&bars->shallow_ready[par], (uint32_t)((it >> 1) & 1), p.err, 33);
                const long long ptile = cluster_id + (long long)it * n_clusters;
                const long long b0 = (2 * ptile + rank) * TSW;
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                    const int s = lane + 32 * s2;
                    float z = bars->shallow[par][s];
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int q = 0; q < 4; ++q) z += bars->red[r][q][s];
                    if (b0 + s < p.B) {
                        if (p.logits) p.logits[b0 + s] = z;
                        if (p.prob) p.prob[b0 + s] = 1.0f / (1.0f + expf(-z));
                    }
                }
                if (lane == 0 && it < 4) FZ_CLK(39 + 8 * it);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();        // neither CTA exits (or frees TMEM) while the other may still read its shared memory / TMEM
    if (warp == W_MMA) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    if (p.clk && threadIdx.x == 32 * W_MMA) {
        unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.clk[blockIdx.x * FZ_NCLK + 30] = clock64(); p.clk[blockIdx.x * FZ_NCLK + 31] = (long long)t;
    }
}

}  // namespace wd
}  // namespace fz
}  // namespace dfw
