// Device-side building blocks of stage 1 (embedding gather, Xv scaling, first order, FM / FwFM second order),
// shared by the stand-alone embed_fwfm kernel (embed_fwfm.cu) and the fused forward kernel (fused_tc.cu).
//
// Replaces model/DeepFMs.py:297-367, 445-450 and model/QREmbeddingBag.py:156-174 of the reference.
//
// A "group" of threads (a whole CTA, or the gather warps of the fused kernel) works on a tile of S samples:
//   SMEM = shallow image copy + the S x (F*K) embedding block (pitch = 2 mod 32 floats: the column reads of a
//          warp hit 32 distinct banks, rows stay 8-byte aligned) + the S x C indices + the S x num dense values
//   embed_gather    phase A  indices / dense values (DRAM) into registers, shallow image (L2) by 16-byte cp.async
//                   phase B  every stored row segment of the block goes global -> SMEM with 8-byte cp.async, all in
//                            flight together (one latency for the whole block, no register staging); then the
//                            fix-ups: quotient (*|+) remainder row for QR tables, times Xv for numeric fields -- one
//                            fp32 operation each, exactly the reference's arithmetic
//   embed_interact  phase D  thread (s,k): e[f] = E[s][f][k] in registers; for each column j: t_i += U_ij e_j (i<j)
//                            -- F-1 independent accumulators, U read as broadcast LDS.128 -> 4 FFMA per LDS;
//                            second = sum_i e_i t_i.  A pruned field matrix is walked as the compacted pair list.
//                   phase E  fixed-order reduction over k -> shallow[s] = first + second + bias
#pragma once
#include "dfw_common.cuh"

namespace dfw {

struct EmbedParams {
    const unsigned char* image;
    const int64_t* xi; int64_t xi_sb, xi_sc;
    const float* xv; int64_t xv_sb, xv_sc;
    const float* fm1; const float* bias;
    float* E; int64_t ldE; __nv_bfloat16* Eb; int64_t ldEb; float* shallow; int32_t* err;
    int64_t B; int F, num, K; unsigned flags;
    long long* clk;   // optional per-CTA phase timestamps (debug tooling, see dfw_debug_set_clock_buffer)
};

struct PairEnt { uint32_t ij; float u; };  // ij = (i*K) | (j*K) << 16
struct ImgHeader { int32_t live, n_list, any_special, misaligned; };   // any_special: some table is QR or rank-sharded

__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }
// column j of the strict upper triangle holds U_0j .. U_(j-1)j, padded to a multiple of 4 floats:
// offset = sum_{c=1}^{j-1} pad4(c) = 4 (m+1)(2m + r) with j-1 = 4m + r
__host__ __device__ constexpr int ucol_off(int j) {
    const int n = j - 1, m = n >> 2, r = n & 3;
    return n <= 0 ? 0 : 4 * (m + 1) * (2 * m + r);
}
__host__ __device__ constexpr int usize(int F) { return ucol_off(F); }
__host__ __device__ constexpr int e_pitch(int FK) { return FK + ((34 - (FK & 31)) & 31); }   // >= FK, == 2 mod 32
__host__ __device__ constexpr size_t up16(size_t v) { return (v + 15) & ~size_t(15); }

struct ImgLayout { size_t oHdr, oU, oPairs, oWl, oFields, total; };
__host__ __device__ inline ImgLayout img_layout(int F, int K) {
    ImgLayout L;
    size_t o = 0;
    L.oHdr = o;    o += 16;
    L.oU = o;      o += up16(sizeof(float) * (usize(F) + 4));
    L.oPairs = o;  o += up16(sizeof(PairEnt) * (size_t)(F * (F - 1) / 2) + 8);
    L.oWl = o;     o += up16(sizeof(float) * F * K);
    L.oFields = o; o += up16(sizeof(dfw_field_desc) * F);
    L.total = o;
    return L;
}

// Shared memory of one tile of S samples.  The image and the E block may live apart (fused kernel), so every
// region is addressed through its own pointer.
struct TileSmem {
    const unsigned char* img;   // shallow image copy (img_layout)
    float* E;                   // S x EP
    float* part;                // K x S
    int32_t* idx;               // S x C
    float* xv;                  // S x num
    int EP;
};
struct TileSizes { int EP; size_t bE, bPart, bIdx, bXv; };
__host__ __device__ inline TileSizes tile_sizes(int F, int K, int num, int S) {
    TileSizes t;
    const int C = F - num;
    t.EP = e_pitch(F * K);
    t.bE = up16(sizeof(float) * S * t.EP);
    t.bPart = up16(sizeof(float) * K * S);
    t.bIdx = up16(sizeof(int32_t) * S * (C > 0 ? C : 1));
    t.bXv = up16(sizeof(float) * (S * (num > 0 ? num : 1) + num * K));     // dense values + one copy of the numeric rows
    return t;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// BAR == 0: the group is the whole CTA (__syncthreads); otherwise named barrier BAR over `nthreads` threads.
template <int BAR>
__device__ __forceinline__ void group_sync(int nthreads) {
    if constexpr (BAR == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------ gather helpers
// Where the stored row of category `idx` of a field lives: quotient-remainder split
// (model/QREmbeddingBag.py:157-158) and rank sharding (owner = row mod P, local row = row div P).
// x / d for the divisors this path meets (QR collisions, rank counts): a shift when d is a power of two (c = 4, P = 2 / 4 / 8: the
// common case), the ~20-instruction division otherwise; the branch is uniform per field
__device__ __forceinline__ uint32_t div_small(uint32_t x, uint32_t d) {
    return (d & (d - 1)) == 0 ? x >> (31 - __clz((int)d)) : x / d;
}
__device__ __forceinline__ const float* locate_row(const dfw_field_desc& fd, int32_t idx, int K) {
    uint32_t row = (uint32_t)idx;
    if (fd.qr_op != DFW_TABLE_PLAIN) row = div_small((uint32_t)idx, (uint32_t)fd.collisions);
    const float* base = fd.w2;
    if (fd.n_ranks > 1) {
        const uint32_t P = (uint32_t)fd.n_ranks;
        const uint32_t local = div_small(row, P);
        base = fd.w2_shard[row - local * P];
        row = local;
    }
    return base + (int64_t)row * K;
}

// (q, r) = divmod(start + n * stride, d) advanced without divisions
struct DivStep {
    uint32_t q, r, dq, dr, d;
    __device__ __forceinline__ DivStep(uint32_t start, uint32_t stride, uint32_t d_) : d(d_) {
        q = start / d; r = start - q * d; dq = stride / d; dr = stride - dq * d;
    }
    __device__ __forceinline__ void next() { q += dq; r += dr; if (r >= d) { r -= d; ++q; } }
};

struct GatherCtx {
    const dfw_field_desc* sF; const int32_t* sIdx; const float* sXv; float* sE;
    int F, K, num, C, EP, nrows, tid, nthreads;
    float* sNum;     // (num, K): the single row of every numeric field, fetched once per tile (they follow the dense values)
};

// Adjacent lanes take adjacent 8-byte (SEGW = 2: all row bases 8-byte aligned, K even) or 4-byte pieces of the
// same row, so one warp instruction touches ~7 rows / cache lines instead of 32 (LSU wavefronts are the cost
// of a gather); all pieces of the block are in flight together, global -> SMEM without register staging.
// PLAIN = no table of the model is QR or rank-sharded: branch-free body (row = w2 + idx * K).
template <int SEGW, int FT, int KT, bool PLAIN>
__device__ __forceinline__ void issue_rows(const GatherCtx& g) {
    const int K = KT > 0 ? KT : g.K;
    const uint32_t nV = (uint32_t)(K / SEGW);
    const uint32_t C = (uint32_t)g.C, num = (uint32_t)g.num;
    // categorical fields: one row per (sample, field)
    if (C == 0) {
        // no categorical field at all
    } else if ((uint32_t)g.nthreads % nV == 0) {
        // the group is a whole number of rows wide: a thread keeps its piece and walks rows tid / nV, + nthreads / nV, ...
        // with (sample, column) advanced incrementally -- no division in the loop
        const uint32_t row0 = (uint32_t)g.tid / nV, k = ((uint32_t)g.tid - row0 * nV) * SEGW, rstep = (uint32_t)g.nthreads / nV;
        const uint32_t nrow = (uint32_t)g.nrows * C;
        DivStep st(row0, rstep, C);
#pragma unroll 4
        for (uint32_t row = row0; row < nrow; row += rstep) {
            const uint32_t s = st.q, f = num + st.r;
            const int32_t idx = g.sIdx[row];
            const float* src = (PLAIN ? g.sF[f].w2 + (int64_t)idx * K : locate_row(g.sF[f], idx, K)) + k;
            float* dst = g.sE + s * g.EP + f * K + k;
            if (SEGW == 2) cp_async8(dst, src); else cp_async4(dst, src);
            st.next();
        }
    } else {
        const uint32_t total = (uint32_t)g.nrows * C * nV;
#pragma unroll 4
        for (uint32_t e = g.tid; e < total; e += g.nthreads) {
            const uint32_t row = e / nV, k = (e - row * nV) * SEGW;
            const uint32_t s = row / C, c = row - s * C, f = num + c;
            const int32_t idx = g.sIdx[row];
            const float* src = (PLAIN ? g.sF[f].w2 + (int64_t)idx * K : locate_row(g.sF[f], idx, K)) + k;
            float* dst = g.sE + s * g.EP + f * K + k;
            if (SEGW == 2) cp_async8(dst, src); else cp_async4(dst, src);
        }
    }
    // numeric fields are one-row tables (model/DeepFMs.py:185-196): fetch each row ONCE per tile, not once per sample --
    // they used to be a third of the gather's copies
    for (uint32_t e = g.tid; e < num * nV; e += g.nthreads) {
        const uint32_t f = e / nV, k = (e - f * nV) * SEGW;
        const float* src = (PLAIN ? g.sF[f].w2 : locate_row(g.sF[f], 0, K)) + k;
        if (SEGW == 2) cp_async8(g.sNum + f * K + k, src); else cp_async4(g.sNum + f * K + k, src);
    }
}

// quotient (*|+) remainder row for QR tables (model/QREmbeddingBag.py:169-172), times Xv for numeric fields
// (model/DeepFMs.py:334): one fp32 operation each, applied after the block's copies have landed.
template <int FT, int KT>
__device__ __forceinline__ void fixup_rows(const GatherCtx& g, bool any_qr) {
    const int K = KT > 0 ? KT : g.K;
    // numeric fields: E[s, f, :] = row_f (x|+ remainder row 0 for a QR table) * Xv[s, f]
    const uint32_t num = (uint32_t)g.num;
#pragma unroll 1
    for (uint32_t e = g.tid; e < (uint32_t)g.nrows * num; e += g.nthreads) {
        const uint32_t s = e / num, f = e - s * num;
        const dfw_field_desc& fd = g.sF[f];
        const int op = fd.qr_op;
        const float x = g.sXv[e];
        const float* row = g.sNum + f * K;
        float* dst = g.sE + s * g.EP + f * K;
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            float v = row[k];
            if (op == DFW_TABLE_QR_MULT) v *= __ldg(fd.w2_r + k);
            else if (op == DFW_TABLE_QR_ADD) v += __ldg(fd.w2_r + k);
            dst[k] = v * x;
        }
    }
    if (!any_qr) return;
    // categorical QR tables: quotient row (already in place) (x|+) remainder row
    const uint32_t C = (uint32_t)g.C;
#pragma unroll 1
    for (uint32_t e = g.tid; e < (uint32_t)g.nrows * C; e += g.nthreads) {
        const uint32_t s = e / C, c = e - s * C, f = num + c;
        const dfw_field_desc& fd = g.sF[f];
        const int op = fd.qr_op;
        if (op == DFW_TABLE_PLAIN) continue;
        float* dst = g.sE + s * g.EP + f * K;
        const int32_t idx = g.sIdx[e];
        const uint32_t cc = (uint32_t)fd.collisions;
        const float* rrow = fd.w2_r + ((uint32_t)idx - ((uint32_t)idx / cc) * cc) * K;
#pragma unroll 2
        for (int k = 0; k < K; ++k) {
            float v = dst[k];
            if (op == DFW_TABLE_QR_MULT) v *= __ldg(rrow + k);
            else v += __ldg(rrow + k);
            dst[k] = v;
        }
    }
}

// Phase-D ownership: thread `tid` of the group owns one sample and, in round r, one embedding column.  A warp always
// covers 16 samples x 2 adjacent columns: with the E pitch == 2 (mod 32) its column reads hit 32 distinct banks.
//   S = 16: sample tid % 16, column tid / 16                       (+ r * nthreads / 16)
//   S = 32: sample tid % 16 + 16 * (warp % 2), column (tid / 16) % 2 + 2 * (warp / 2)   (+ r * nthreads / 32)
template <int S> __device__ __forceinline__ int owner_sample(int tid) {
    if constexpr (S == 32) return (tid & 15) + (((tid >> 5) & 1) << 4);
    else return tid % S;
}
template <int S> __device__ __forceinline__ int owner_col(int tid, int nthreads, int r) {
    if constexpr (S == 32) return ((tid >> 4) & 1) + 2 * (tid >> 6) + r * (nthreads / S);
    else return tid / S + r * (nthreads / S);
}

// one index of the batch: int64 as the reference feeds it, or int32 (DFW_XI_INT32)
__device__ __forceinline__ int64_t load_index(const EmbedParams& p, int64_t elem) {
    if (p.flags & DFW_XI_INT32) return (int64_t)__ldg(reinterpret_cast<const int32_t*>(p.xi) + elem);
    return __ldg(p.xi + elem);
}

// ------------------------------------------------------------------------------------------ phases A + B
// Fills sm.E (S x EP, rows >= nrows zeroed), sm.idx, sm.xv for samples [b0, b0 + nrows).  Returns through
// first_acc[r] the first-order table partial (use_fwlw = 0) of this thread's (sample, column) of round r.
// Ends with a group barrier: the E block is complete and visible to the whole group.
template <int FT, int KT, int S, int R, int BAR>
__device__ __forceinline__ void embed_gather(const EmbedParams& p, const TileSmem& sm, unsigned char* sImgDst, bool load_image,
                                              int tid, int nthreads, int64_t b0, int nrows, float (&first_acc)[R],
                                              long long* clk) {
    const int F = FT > 0 ? FT : p.F;
    const int K = KT > 0 ? KT : p.K;
    const int num = p.num;
    const int C = F - num;
    const int FK = F * K;
    const ImgLayout IL = img_layout(F, K);
    const ImgHeader* hdr = reinterpret_cast<const ImgHeader*>(sm.img + IL.oHdr);
    const dfw_field_desc* sF = reinterpret_cast<const dfw_field_desc*>(sm.img + IL.oFields);
    float* sE = sm.E;
    int32_t* sIdx = sm.idx;
    float* sXv = sm.xv;
    const int EP = sm.EP;
    const bool fwlw = p.flags & DFW_USE_FWLW;
#define DFW_CLK(slot) do { if (clk && tid == 0) clk[(slot)] = clock64(); } while (0)
    DFW_CLK(0);
    // ------------------------------------------------------------------ phase A
    // batch inputs first: they come from DRAM; the shallow image is L2-resident model state
    constexpr int kMaxIdx = 4, kMaxXv = 2;
    int64_t myidx[kMaxIdx];
    int mycol[kMaxIdx];
    float myxv[kMaxXv];
    const int nIdx = S * C, nXv = S * num;
    {
        DivStep st(tid, nthreads, C > 0 ? C : 1);
#pragma unroll
        for (int r = 0; r < kMaxIdx; ++r) {
            myidx[r] = 0; mycol[r] = (int)st.r;
            if (tid + r * nthreads < nIdx && (int)st.q < nrows) myidx[r] = load_index(p, (b0 + st.q) * p.xi_sb + st.r * p.xi_sc);
            st.next();
        }
        DivStep sv(tid, nthreads, num > 0 ? num : 1);
#pragma unroll
        for (int r = 0; r < kMaxXv; ++r) {
            myxv[r] = 0.f;
            if (tid + r * nthreads < nXv && (int)sv.q < nrows) myxv[r] = p.xv[(b0 + sv.q) * p.xv_sb + sv.r * p.xv_sc];
            sv.next();
        }
    }
    if (load_image) {
#pragma unroll 2
        for (uint32_t i = tid; i < (uint32_t)(IL.total >> 4); i += nthreads) cp_async16(sImgDst + 16 * i, p.image + 16 * i);
        cp_async_wait_all();
    }
    group_sync<BAR>(nthreads);   // image visible (row counts for the bounds check); previous tile's readers are done
    DFW_CLK(1);

#pragma unroll
    for (int r = 0; r < kMaxIdx; ++r) {
        const int e = tid + r * nthreads;
        if (e < nIdx) {
            const int c = mycol[r];
            int64_t idx = myidx[r];
            if (idx < 0 || idx >= sF[num + c].rows) {   // defined behaviour instead of a wild read
                if (p.err) atomicExch(p.err, 1 + num + c);
                idx = 0;
            }
            sIdx[e] = (int32_t)idx;
        }
    }
#pragma unroll
    for (int r = 0; r < kMaxXv; ++r) {
        const int e = tid + r * nthreads;
        if (e < nXv) sXv[e] = myxv[r];
    }
#pragma unroll 1
    for (uint32_t e = tid + kMaxXv * nthreads; e < (uint32_t)nXv; e += nthreads) {      // small groups only
        const uint32_t s = e / (uint32_t)num, f = e - s * num;
        sXv[e] = (int)s < nrows ? p.xv[(b0 + s) * p.xv_sb + f * p.xv_sc] : 0.f;
    }
#pragma unroll 1
    for (uint32_t e = tid + kMaxIdx * nthreads; e < (uint32_t)nIdx; e += nthreads) {    // small groups only
        const uint32_t s = e / (uint32_t)C, c = e - s * C;
        int64_t idx = (int)s < nrows ? load_index(p, (b0 + s) * p.xi_sb + c * p.xi_sc) : 0;
        if (idx < 0 || idx >= sF[num + c].rows) {
            if (p.err) atomicExch(p.err, 1 + num + c);
            idx = 0;
        }
        sIdx[e] = (int32_t)idx;
    }
    group_sync<BAR>(nthreads);   // indices + dense values visible
    DFW_CLK(2);

    const bool any_qr = hdr->any_special != 0;
    const bool vec2 = hdr->misaligned == 0 && (K % 2 == 0);    // all row bases 8-byte aligned

    // ------------------------------------------------------------------ phase B: gather
    GatherCtx g{sF, sIdx, sXv, sE, F, K, num, C, EP, nrows, tid, nthreads, sXv + S * (num > 0 ? num : 1)};
    if (!any_qr) { if (vec2) issue_rows<2, FT, KT, true>(g); else issue_rows<1, FT, KT, true>(g); }
    else         { if (vec2) issue_rows<2, FT, KT, false>(g); else issue_rows<1, FT, KT, false>(g); }
    // rows of samples past the end of the batch: zeros (never written out, keeps phase D finite)
#pragma unroll 1
    for (uint32_t e = tid; e < (uint32_t)((S - nrows) * FK); e += nthreads) {
        const uint32_t s = e / (uint32_t)FK;
        sE[(nrows + s) * EP + (e - s * FK)] = 0.f;
    }

    // first-order table values of this thread's fields f = kk, kk+K, ...   (model/DeepFMs.py:300-309)
    const int smp = owner_sample<S>(tid);
    const bool live_sample = smp < nrows;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        first_acc[r] = 0.f;
        const int kk = owner_col<S>(tid, nthreads, r);
        if (!fwlw && live_sample && kk < K) {
            float acc = 0.f;
#pragma unroll 4
            for (int f = kk; f < F; f += K) {
                const dfw_field_desc& fd = sF[f];
                const int32_t idx = f < num ? 0 : sIdx[smp * C + (f - num)];
                float v;
                if (fd.qr1_op != DFW_TABLE_PLAIN) {
                    const uint32_t c = (uint32_t)fd.collisions;
                    const uint32_t q = (uint32_t)idx / c, rr = (uint32_t)idx - q * c;
                    const float a = __ldg(fd.w1 + q), b = __ldg(fd.w1_r + rr);
                    v = fd.qr1_op == DFW_TABLE_QR_MULT ? a * b : a + b;
                } else {
                    v = __ldg(fd.w1 + idx);
                }
                if (f < num) v *= sXv[smp * num + f];
                if (p.flags & DFW_USE_LW) v *= __ldg(p.fm1 + f);                              // model/DeepFMs.py:450
                acc += v;
            }
            first_acc[r] = acc;
        }
    }

    DFW_CLK(3);
    cp_async_wait_all();
    if (any_qr || num > 0) {
        group_sync<BAR>(nthreads);   // every thread's pieces have landed: fix rows up in place
        fixup_rows<FT, KT>(g, any_qr);
    }
    group_sync<BAR>(nthreads);   // E block complete
    DFW_CLK(4);
}

// ------------------------------------------------------------------------------------------ phases D + E
// shallow_dst[s] = sum(first) + sum(second) + bias for s < nrows (any address space).
// cU != nullptr (dense unrolled path only): U is read from there with compile-time offsets -- the fused kernel passes its
// __grid_constant__ parameter copy, so every U_ij becomes a constant-bank operand of its FFMA and phase D issues no
// shared-memory loads for the field matrix (a broadcast LDS.128 still costs 4 register-write cycles per warp).
// CONSTU (fused kernel): the dense path takes U only from cU; without it the kernel walks the pair list.
template <int FT, int KT, int S, int R, int BAR, bool CONSTU = false>
__device__ __forceinline__ void embed_interact(const EmbedParams& p, const TileSmem& sm, int tid, int nthreads, int nrows,
                                                const float (&first_acc)[R], float* shallow_dst, long long* clk,
                                                const float* cU = nullptr) {
    const int F = FT > 0 ? FT : p.F;
    const int K = KT > 0 ? KT : p.K;
    const ImgLayout IL = img_layout(F, K);
    const ImgHeader* hdr = reinterpret_cast<const ImgHeader*>(sm.img + IL.oHdr);
    const float* sU = reinterpret_cast<const float*>(sm.img + IL.oU);
    const PairEnt* sPairs = reinterpret_cast<const PairEnt*>(sm.img + IL.oPairs);
    const float* sWl = reinterpret_cast<const float*>(sm.img + IL.oWl);
    const float* sE = sm.E;
    float* sPart = sm.part;
    const int EP = sm.EP;
    const bool fwlw = p.flags & DFW_USE_FWLW;
    const int P = F * (F - 1) / 2;
    const bool use_list = (FT == 0) || (hdr->live * 6 < P) || (CONSTU && cU == nullptr);
    const int smp = owner_sample<S>(tid);
    DFW_CLK(5);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int kk = owner_col<S>(tid, nthreads, r);
        if (kk >= K) continue;
        const float* myE = sE + smp * EP + kk;   // E[s][f][k] at myE[f*K]
        float acc = first_acc[r];
        if (fwlw) {                                // model/DeepFMs.py:344-345
            float a0 = 0.f, a1 = 0.f;
            int f = 0;
#pragma unroll 2
            for (; f + 1 < F; f += 2) {
                a0 = fmaf(myE[f * K], sWl[f * K + kk], a0);
                a1 = fmaf(myE[(f + 1) * K], sWl[(f + 1) * K + kk], a1);
            }
            if (f < F) a0 = fmaf(myE[f * K], sWl[f * K + kk], a0);
            acc = a0 + a1;
        }
        float second = 0.f;
        if (use_list) {
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
        } else if constexpr (FT > 0 && CONSTU) {
            // second = sum_j e_j * (sum_{i<j} U_ij e_i): every U_ij is a constant-bank operand, 4 independent chains per j
            constexpr int FTc = FT > 0 ? FT : 2, KTc = KT > 0 ? KT : 1;
            float e[FTc];
#pragma unroll
            for (int f = 0; f < FTc; ++f) e[f] = myE[f * KTc];
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int j = 1; j < FTc; ++j) {
                float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                for (int i = 0; i < j; ++i) {
                    const float u = cU[ucol_off(j) + i];
                    if ((i & 3) == 0) d0 = fmaf(u, e[i], d0);
                    else if ((i & 3) == 1) d1 = fmaf(u, e[i], d1);
                    else if ((i & 3) == 2) d2 = fmaf(u, e[i], d2);
                    else d3 = fmaf(u, e[i], d3);
                }
                const float dot = (d0 + d1) + (d2 + d3);
                if (j & 1) s0 = fmaf(e[j], dot, s0); else s1 = fmaf(e[j], dot, s1);
            }
            second = s0 + s1;
        } else if constexpr (FT > 0) {
            constexpr int FTc = FT > 0 ? FT : 2, KTc = KT > 0 ? KT : 1;
            float e[FTc], t[FTc];
#pragma unroll
            for (int f = 0; f < FTc; ++f) { e[f] = myE[f * KTc]; t[f] = 0.f; }
            const float4* sU4 = reinterpret_cast<const float4*>(sU);
#pragma unroll
            for (int j = 1; j < FTc; ++j) {
                const int off4 = ucol_off(j) >> 2;
#pragma unroll
                for (int c4 = 0; c4 < pad4(j) / 4; ++c4) {
                    const float4 u = sU4[off4 + c4];
                    const int i = 4 * c4;
                    if (i + 0 < j) t[i + 0] = fmaf(u.x, e[j], t[i + 0]);
                    if (i + 1 < j) t[i + 1 < FTc ? i + 1 : 0] = fmaf(u.y, e[j], t[i + 1 < FTc ? i + 1 : 0]);
                    if (i + 2 < j) t[i + 2 < FTc ? i + 2 : 0] = fmaf(u.z, e[j], t[i + 2 < FTc ? i + 2 : 0]);
                    if (i + 3 < j) t[i + 3 < FTc ? i + 3 : 0] = fmaf(u.w, e[j], t[i + 3 < FTc ? i + 3 : 0]);
                }
            }
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i + 1 < FTc; i += 2) {
                s0 = fmaf(e[i], t[i], s0);
                if (i + 1 < FTc - 1) s1 = fmaf(e[i + 1], t[i + 1], s1);
            }
            second = s0 + s1;
        }
        sPart[kk * S + smp] = acc + second;
    }
    group_sync<BAR>(nthreads);
    DFW_CLK(6);
    // ------------------------------------------------------------------ phase E: reduce over k
    if (tid < nrows) {
        float tot = 0.f;
#pragma unroll 1
        for (int k = 0; k < K; ++k) tot += sPart[k * S + tid];
        shallow_dst[tid] = tot + __ldg(p.bias);
    }
}
#undef DFW_CLK

}  // namespace dfw
