// Stage 1 of the DeepFwFM forward: embedding gather (plain / quotient-remainder / rank-sharded),
// Xv scaling, first-order term (tables or field-weighted linear), FM / FwFM second order.
//
// Replaces model/DeepFMs.py:297-367, 445-450 and model/QREmbeddingBag.py:156-174 of the reference.
//
// Layout of the work
//   CTA  = 32 samples x K "k-lanes":  blockDim = (32, K);  thread (s, k) owns column k of sample s
//   SMEM = the 32 x (F*K) embedding block of the CTA's samples (odd pitch -> conflict-free column
//          reads), the symmetrised upper-triangular field matrix U (rows padded to float4), the
//          fwlw weights, the field descriptors and the 32 x C indices
//   phase A  descriptors, indices, U and fwlw weights -> SMEM                     (1 DRAM latency)
//   phase B  all row segments of the CTA are requested before any is consumed     (1 DRAM latency)
//            (32 samples x C rows x K floats, float2/float4 per request when K allows)
//   phase C  E block streamed out (fp32 and/or bf16) while the FMAs of phase D run
//   phase D  thread (s,k): e[f] = E[s][f][k] in registers;  t_i = sum_{j>i} U_ij e_j ;  acc += e_i t_i
//            U comes from SMEM as broadcast LDS.128, so the inner loop is 4 FFMA per LDS
//            when the field matrix is pruned (few live pairs) a compacted pair list is walked instead
//   phase E  fixed-order reduction over k -> shallow[b] = first + second + bias
//
// Bytes per sample the algorithm needs (Criteo, fwlw): 26*8 + 13*4 + 26*40 + 4 = 1304.
#include "dfw_common.cuh"

namespace dfw {

constexpr int kSamples = 32;  // samples per CTA (one warp lane each)

struct EmbedParams {
    const dfw_field_desc* fields;
    const int64_t* xi; int64_t xi_sb, xi_sc;
    const float* xv; int64_t xv_sb, xv_sc;
    const float* wl; const float* fm1; const float* cov; const float* bias;
    float* E; int64_t ldE; __nv_bfloat16* Eb; int64_t ldEb; float* shallow; int32_t* err;
    int64_t B; int F, num, K; unsigned flags; int my_rank;
};

struct PairEnt { uint32_t ij; float u; };  // ij = (i*K) | (j*K) << 16

__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }
__host__ __device__ constexpr int urow_off(int F, int i) {
    int o = 0;
    for (int r = 0; r < i; ++r) o += pad4(F - 1 - r);
    return o;
}
__host__ __device__ constexpr int usize(int F) { return urow_off(F, F - 1); }

struct SmemLayout {
    int EP;       // pitch of the E block in floats (odd)
    size_t oU, oPairs, oE, oWl, oPart, oFields, oIdx, oMisc, total;
};
__host__ __device__ inline SmemLayout smem_layout(int F, int K, int C) {
    SmemLayout L;
    const int FK = F * K;
    L.EP = FK | 1;
    size_t o = 0;
    L.oU = o;      o += sizeof(float) * pad4(usize(F) + 4);
    L.oPairs = o;  o += sizeof(PairEnt) * (size_t)(F * (F - 1) / 2);
    o = (o + 15) & ~size_t(15);
    L.oFields = o; o += sizeof(dfw_field_desc) * F;
    L.oE = o;      o += sizeof(float) * kSamples * L.EP;
    L.oWl = o;     o += sizeof(float) * FK;
    L.oPart = o;   o += sizeof(float) * K * kSamples;
    L.oIdx = o;    o += sizeof(int32_t) * kSamples * (C > 0 ? C : 1);
    L.oMisc = o;   o += 16;
    L.total = o;
    return L;
}

template <int VEC> struct VecT;
template <> struct VecT<1> { using T = float; };
template <> struct VecT<2> { using T = float2; };
template <> struct VecT<4> { using T = float4; };

template <int VEC>
__device__ __forceinline__ void ld_vec(const float* p, float (&v)[VEC]) {
    if ((reinterpret_cast<uintptr_t>(p) & (VEC * 4 - 1)) == 0) {
        typename VecT<VEC>::T t = __ldg(reinterpret_cast<const typename VecT<VEC>::T*>(p));
        const float* tf = reinterpret_cast<const float*>(&t);
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = tf[i];
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = __ldg(p + i);
    }
}

// Row segment [off, off+VEC) of field `fd` for category `idx` (already bounds-checked, < 2^31).
template <int VEC>
__device__ __forceinline__ void fetch_row(const dfw_field_desc& fd, int32_t idx, int K, int off, float (&v)[VEC]) {
    uint32_t row = (uint32_t)idx, rem = 0;
    if (fd.qr_op != DFW_TABLE_PLAIN) {          // model/QREmbeddingBag.py:157-158
        const uint32_t c = (uint32_t)fd.collisions;
        row = (uint32_t)idx / c;
        rem = (uint32_t)idx - row * c;
    }
    const float* base = fd.w2;
    if (fd.n_ranks > 1) {                       // row-sharded: owner = row mod P, local row = row div P
        const uint32_t P = (uint32_t)fd.n_ranks;
        const uint32_t local = row / P;
        base = fd.w2_shard[row - local * P];
        row = local;
    }
    ld_vec<VEC>(base + (int64_t)row * K + off, v);
    if (fd.qr_op != DFW_TABLE_PLAIN) {
        float r[VEC];
        ld_vec<VEC>(fd.w2_r + rem * K + off, r);
#pragma unroll
        for (int i = 0; i < VEC; ++i) v[i] = (fd.qr_op == DFW_TABLE_QR_MULT) ? v[i] * r[i] : v[i] + r[i];
    }
}

__device__ __forceinline__ float fetch_first(const dfw_field_desc& fd, int32_t idx) {
    if (fd.qr1_op != DFW_TABLE_PLAIN) {
        const uint32_t c = (uint32_t)fd.collisions;
        const uint32_t q = (uint32_t)idx / c, r = (uint32_t)idx - q * c;
        const float a = __ldg(fd.w1 + q), b = __ldg(fd.w1_r + r);
        return fd.qr1_op == DFW_TABLE_QR_MULT ? a * b : a + b;
    }
    return __ldg(fd.w1 + idx);
}

template <int FT, int KT, int VEC>
__global__ void __launch_bounds__(FT > 0 ? kSamples * KT : 1024)
embed_fwfm_kernel(const EmbedParams p) {
    const int F = FT > 0 ? FT : p.F;
    const int K = KT > 0 ? KT : p.K;
    const int num = p.num;
    const int C = F - num;
    const int FK = F * K;
    const int lane = threadIdx.x;            // sample within the CTA
    const int kk = threadIdx.y;              // embedding column owned in phase D
    const int nthreads = kSamples * K;
    const int tid = kk * kSamples + lane;
    const int64_t b0 = (int64_t)blockIdx.x * kSamples;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SmemLayout L = smem_layout(F, K, C);
    float* sU = reinterpret_cast<float*>(smem_raw + L.oU);
    PairEnt* sPairs = reinterpret_cast<PairEnt*>(smem_raw + L.oPairs);
    dfw_field_desc* sF = reinterpret_cast<dfw_field_desc*>(smem_raw + L.oFields);
    float* sE = reinterpret_cast<float*>(smem_raw + L.oE);
    float* sWl = reinterpret_cast<float*>(smem_raw + L.oWl);
    float* sPart = reinterpret_cast<float*>(smem_raw + L.oPart);
    int32_t* sIdx = reinterpret_cast<int32_t*>(smem_raw + L.oIdx);
    int* sMisc = reinterpret_cast<int*>(smem_raw + L.oMisc);   // [0] = number of live pairs
    const int EP = L.EP;
    const bool fwfm = p.flags & DFW_USE_FWFM;
    const bool fwlw = p.flags & DFW_USE_FWLW;

    // ------------------------------------------------------------------ phase A
    {   // field descriptors (F x 120 B), copied as 8-byte words
        const uint64_t* src = reinterpret_cast<const uint64_t*>(p.fields);
        uint64_t* dst = reinterpret_cast<uint64_t*>(sF);
        const int nw = F * (int)(sizeof(dfw_field_desc) / 8);
        for (int i = tid; i < nw; i += nthreads) dst[i] = __ldg(src + i);
    }
    if (fwlw) {   // fwlw weights; with use_lw the per-field projection fm_1st[f] is folded in (model/DeepFMs.py:450)
        const bool lw = p.flags & DFW_USE_LW;
        for (int i = tid; i < FK; i += nthreads) sWl[i] = __ldg(p.wl + i) * (lw ? __ldg(p.fm1 + i / K) : 1.0f);
    }
    if (tid == 0) sMisc[0] = 0;
    __syncthreads();   // descriptors visible (rows needed for the bounds check)

    for (int e = tid; e < kSamples * C; e += nthreads) {
        const int s = e / C, c = e - s * C;
        const int64_t b = b0 + s;
        int64_t idx = 0;
        if (b < p.B) {
            idx = p.xi[b * p.xi_sb + c * p.xi_sc];
            if (idx < 0 || idx >= sF[num + c].rows) {   // defined behaviour instead of a wild read
                if (p.err) atomicExch(p.err, 1 + num + c);
                idx = 0;
            }
        }
        sIdx[e] = (int32_t)idx;
    }

    // U = upper triangle of (W + W^T)/2 (model/DeepFMs.py:364), rows padded to float4; FM: ones.
    // Warp w builds rows w, w+nwarps, ...; the live-pair count decides dense vs pair-list walking.
    if constexpr (FT > 0) {
        const int nwarps = K;
        int live = 0;
        for (int i = kk; i < F - 1; i += nwarps) {
            const int off = urow_off(F, i);
            const int width = pad4(F - 1 - i);
            for (int jj = lane; jj < width; jj += 32) {
                const int j = i + 1 + jj;
                float u = 0.f;
                if (j < F) u = fwfm ? (__ldg(p.cov + j * F + i) + __ldg(p.cov + i * F + j)) * 0.5f : 1.0f;
                sU[off + jj] = u;
                live += (u != 0.f);
            }
        }
        for (int o = 16; o > 0; o >>= 1) live += __shfl_xor_sync(0xffffffffu, live, o);
        if (lane == 0 && live) atomicAdd(&sMisc[0], live);
    }
    __syncthreads();   // indices + U + live count visible

    const int P = F * (F - 1) / 2;
    const bool use_list = (FT == 0) || (sMisc[0] * 6 < P);
    if (use_list && kk == 0) {
        // ordered compaction of the live pairs by one warp -> deterministic summation order
        int n = 0;
        for (int i = 0; i < F - 1; ++i) {
            for (int jb = i + 1; jb < F; jb += 32) {
                const int j = jb + lane;
                float u = 0.f;
                if (j < F) u = fwfm ? (__ldg(p.cov + j * F + i) + __ldg(p.cov + i * F + j)) * 0.5f : 1.0f;
                const unsigned m = __ballot_sync(0xffffffffu, u != 0.f);
                if (u != 0.f) {
                    const int pos = n + __popc(m & ((1u << lane) - 1u));
                    sPairs[pos].ij = (uint32_t)(i * K) | ((uint32_t)(j * K) << 16);
                    sPairs[pos].u = u;
                }
                n += __popc(m);
            }
        }
        if (lane == 0) sMisc[1] = n;
    }

    // first-order table values of this thread's fields f = kk, kk+K, ... (requested early, used in phase D)
    float first_acc = 0.f;
    const int64_t bme = b0 + lane;
    const bool live_sample = bme < p.B;
    if (!fwlw && live_sample) {
        for (int f = kk; f < F; f += K) {
            float v;
            if (f < num) v = fetch_first(sF[f], 0) * p.xv[bme * p.xv_sb + f * p.xv_sc];   // model/DeepFMs.py:302-304
            else v = fetch_first(sF[f], sIdx[lane * C + (f - num)]);
            if (p.flags & DFW_USE_LW) v *= __ldg(p.fm1 + f);                              // model/DeepFMs.py:450
            first_acc += v;
        }
    }

    // ------------------------------------------------------------------ phase B: gather
    {
        const int nV = K / VEC;
        const int per_sample = F * nV;
        const int total = kSamples * per_sample;
        if constexpr (FT > 0) {
            constexpr int FTc = FT > 0 ? FT : 1, KTc = KT > 0 ? KT : 1;
            constexpr int TOTAL = kSamples * FTc * (KTc / VEC);
            constexpr int NT = kSamples * KTc;
            constexpr int ITEMS = (TOTAL + NT - 1) / NT;
            float v[ITEMS][VEC];
#pragma unroll
            for (int it = 0; it < ITEMS; ++it) {
                const int e = tid + it * NT;
#pragma unroll
                for (int i = 0; i < VEC; ++i) v[it][i] = 0.f;
                if (e < TOTAL) {
                    const int s = e / per_sample, r = e - s * per_sample;
                    const int f = r / nV, off = (r - f * nV) * VEC;
                    const int64_t b = b0 + s;
                    if (b < p.B) {
                        if (f < num) {
                            fetch_row<VEC>(sF[f], 0, K, off, v[it]);
                            const float x = p.xv[b * p.xv_sb + f * p.xv_sc];
#pragma unroll
                            for (int i = 0; i < VEC; ++i) v[it][i] *= x;               // model/DeepFMs.py:334
                        } else {
                            fetch_row<VEC>(sF[f], sIdx[s * C + (f - num)], K, off, v[it]);
                        }
                    }
                }
            }
#pragma unroll
            for (int it = 0; it < ITEMS; ++it) {
                const int e = tid + it * NT;
                if (e < TOTAL) {
                    const int s = e / per_sample, r = e - s * per_sample;
                    const int f = r / nV, off = (r - f * nV) * VEC;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) sE[s * EP + f * K + off + i] = v[it][i];
                }
            }
        } else {
            for (int e0 = tid; e0 < total; e0 += 4 * nthreads) {
                float v[4][VEC];
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int e = e0 + it * nthreads;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) v[it][i] = 0.f;
                    if (e < total) {
                        const int s = e / per_sample, r = e - s * per_sample;
                        const int f = r / nV, off = (r - f * nV) * VEC;
                        const int64_t b = b0 + s;
                        if (b < p.B) {
                            if (f < num) {
                                fetch_row<VEC>(sF[f], 0, K, off, v[it]);
                                const float x = p.xv[b * p.xv_sb + f * p.xv_sc];
#pragma unroll
                                for (int i = 0; i < VEC; ++i) v[it][i] *= x;
                            } else {
                                fetch_row<VEC>(sF[f], sIdx[s * C + (f - num)], K, off, v[it]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int e = e0 + it * nthreads;
                    if (e < total) {
                        const int s = e / per_sample, r = e - s * per_sample;
                        const int f = r / nV, off = (r - f * nV) * VEC;
#pragma unroll
                        for (int i = 0; i < VEC; ++i) sE[s * EP + f * K + off + i] = v[it][i];
                    }
                }
            }
        }
    }
    __syncthreads();   // E block + pair list complete

    // ------------------------------------------------------------------ phase C: stream E out
    const int nrows = (int)min((int64_t)kSamples, p.B - b0);
    if (p.E) {
        const int ld = (int)p.ldE;
        float* dst = p.E + b0 * p.ldE;
        int s = tid / ld, c = tid - s * ld;                      // (row, column) advanced without divisions
        const int ds = nthreads / ld, dc = nthreads - ds * ld;
        for (int i = tid; i < nrows * ld; i += nthreads) {
            dst[i] = c < FK ? sE[s * EP + c] : 0.f;
            s += ds; c += dc;
            if (c >= ld) { c -= ld; ++s; }
        }
    }
    if (p.Eb) {
        const int ld2 = (int)(p.ldEb >> 1);
        __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(p.Eb + b0 * p.ldEb);
        int s = tid / ld2, c = tid - s * ld2;
        const int ds = nthreads / ld2, dc = nthreads - ds * ld2;
        for (int i = tid; i < nrows * ld2; i += nthreads) {
            const int c0 = 2 * c;
            const float lo = c0 < FK ? sE[s * EP + c0] : 0.f;
            const float hi = c0 + 1 < FK ? sE[s * EP + c0 + 1] : 0.f;
            dst[i] = __floats2bfloat162_rn(lo, hi);
            s += ds; c += dc;
            if (c >= ld2) { c -= ld2; ++s; }
        }
    }

    // ------------------------------------------------------------------ phase D: first + second order
    const float* myE = sE + lane * EP + kk;   // E[s][f][k] at myE[f*K]
    float acc = first_acc;
    if (fwlw) {                                // model/DeepFMs.py:344-345
        float a1 = 0.f;
        for (int f = 0; f < F; ++f) a1 = fmaf(myE[f * K], sWl[f * K + kk], a1);
        acc = a1;
    }
    float second = 0.f;
    if (use_list) {
        const int n = sMisc[1];
        float s0 = 0.f, s1 = 0.f;
        int q = 0;
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
    } else if constexpr (FT > 0) {
        constexpr int FTc = FT > 0 ? FT : 2, KTc = KT > 0 ? KT : 1;
        float e[FTc];
#pragma unroll
        for (int f = 0; f < FTc; ++f) e[f] = myE[f * KTc];
        const float4* sU4 = reinterpret_cast<const float4*>(sU);
#pragma unroll
        for (int i = 0; i < FTc - 1; ++i) {
            const int off4 = urow_off(FTc, i) >> 2;
            float t = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < pad4(FTc - 1 - i) / 4; ++c4) {
                const float4 u = sU4[off4 + c4];
                const int j = i + 1 + 4 * c4;
                if (j + 0 < FTc) t = fmaf(u.x, e[j + 0 < FTc ? j + 0 : 0], t);
                if (j + 1 < FTc) t = fmaf(u.y, e[j + 1 < FTc ? j + 1 : 0], t);
                if (j + 2 < FTc) t = fmaf(u.z, e[j + 2 < FTc ? j + 2 : 0], t);
                if (j + 3 < FTc) t = fmaf(u.w, e[j + 3 < FTc ? j + 3 : 0], t);
            }
            second = fmaf(e[i], t, second);
        }
    }
    sPart[kk * kSamples + lane] = acc + second;
    __syncthreads();

    // ------------------------------------------------------------------ phase E: reduce over k
    if (kk == 0 && live_sample) {
        float tot = 0.f;
        for (int k = 0; k < K; ++k) tot += sPart[k * kSamples + lane];
        p.shallow[bme] = tot + __ldg(p.bias);
    }
}

template <int FT, int KT, int VEC>
static int launch_embed(const EmbedParams& p, cudaStream_t st) {
    const int F = p.F, K = p.K, C = F - p.num;
    const SmemLayout L = smem_layout(F, K, C);
    auto kern = embed_fwfm_kernel<FT, KT, VEC>;
    static thread_local size_t configured = 0;
    if (L.total > 48 * 1024 && L.total > configured) {
        DFW_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
        configured = L.total;
    }
    const dim3 block(kSamples, K);
    const dim3 grid((unsigned)((p.B + kSamples - 1) / kSamples));
    kern<<<grid, block, L.total, st>>>(p);
    count_launch();
    return check_launch("embed_fwfm_kernel");
}

}  // namespace dfw

extern "C" int dfw_embed_fwfm(const dfw_model* m, const int64_t* xi, int64_t xi_stride_b, int64_t xi_stride_c,
                              const float* xv, int64_t xv_stride_b, int64_t xv_stride_c, int64_t B,
                              float* E_out, int64_t ldE, void* E_bf16_out, int64_t ldEb,
                              float* shallow_out, int32_t* err_word, int32_t my_rank, void* stream) {
    using namespace dfw;
    if (int rc = check_model(m)) return rc;
    DFW_REQUIRE(B >= 0, DFW_E_ARG, "negative batch");
    if (B == 0) return 0;
    const int F = m->field_size, K = m->embedding_size, num = m->numerical;
    DFW_REQUIRE(shallow_out, DFW_E_ARG, "shallow_out is NULL");
    DFW_REQUIRE(F - num == 0 || xi, DFW_E_ARG, "xi is NULL");
    DFW_REQUIRE(num == 0 || xv, DFW_E_ARG, "xv is NULL");
    DFW_REQUIRE(!E_out || ldE >= (int64_t)F * K, DFW_E_ARG, "ldE %lld < F*K", (long long)ldE);
    DFW_REQUIRE(!E_bf16_out || (ldEb >= (int64_t)F * K && ldEb % 2 == 0), DFW_E_ARG, "ldEb must be even and >= F*K");
    DFW_REQUIRE((int64_t)F * K < 65536, DFW_E_UNSUPPORTED, "F*K must be < 65536");
    EmbedParams p;
    p.fields = m->fields;
    p.xi = xi; p.xi_sb = xi_stride_b; p.xi_sc = xi_stride_c;
    p.xv = xv; p.xv_sb = xv_stride_b; p.xv_sc = xv_stride_c;
    p.wl = m->fwfm_linear; p.fm1 = m->fm_1st; p.cov = m->field_cov; p.bias = m->bias;
    p.E = E_out; p.ldE = ldE; p.Eb = reinterpret_cast<__nv_bfloat16*>(E_bf16_out); p.ldEb = ldEb;
    p.shallow = shallow_out; p.err = (m->flags & DFW_CHECK_INDEX) ? err_word : nullptr;
    p.B = B; p.F = F; p.num = num; p.K = K; p.flags = m->flags; p.my_rank = my_rank;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // the two dataset shapes BASELINE.json names get fully unrolled instantiations
    if (F == 39 && K == 10) return launch_embed<39, 10, 2>(p, st);
    if (F == 47 && K == 10) return launch_embed<47, 10, 2>(p, st);
    if (K % 4 == 0) return launch_embed<0, 0, 4>(p, st);
    if (K % 2 == 0) return launch_embed<0, 0, 2>(p, st);
    return launch_embed<0, 0, 1>(p, st);
}
