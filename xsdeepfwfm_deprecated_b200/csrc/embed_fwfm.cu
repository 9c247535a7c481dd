// Stage 1 of the DeepFwFM forward: embedding gather (plain / quotient-remainder / rank-sharded),
// Xv scaling, first-order term (tables or field-weighted linear), FM / FwFM second order.
//
// Replaces model/DeepFMs.py:297-367, 445-450 and model/QREmbeddingBag.py:156-174 of the reference.
//
// Two kernels:
//
// pack_shallow_kernel (one CTA, run when the model is (re)packed, not per batch) builds the "shallow image":
//   header | U = strict upper triangle of (R + R^T)/2 stored by columns, padded to float4 (FM: ones) |
//   compacted list of live pairs (for pruned R) | fwlw weights with fm_1st folded in | field descriptors
//
// embed_fwfm_kernel, per batch (device code shared with the fused forward kernel: embed_device.cuh):
//   CTA  = 16 samples x K columns: thread t owns (sample t & 15, column t >> 4) in phase D
//   SMEM = shallow image copy + the 16 x (F*K) embedding block of the CTA's samples (pitch = 2 mod 32
//          floats: the column reads of a warp -- 16 samples x 2 columns -- hit 32 distinct banks, rows stay
//          8-byte aligned) + the 16 x C indices + the 16 x num dense values
//   phase A  indices / dense values (DRAM) into registers, shallow image (L2) by 16-byte cp.async  (1 latency)
//   phase B  every stored row segment of the block goes global -> SMEM with 8-byte cp.async, all in flight
//            together (one latency for the whole block, no register staging); the issuing thread then fixes
//            up its own segments: quotient (*|+) remainder row for QR tables, times Xv for numeric fields --
//            one fp32 operation each, exactly the reference's arithmetic
//   phase C  E block streamed out (fp32 and/or bf16), coalesced
//   phase D  thread (s,k): e[f] = E[s][f][k] in registers; for each column j: t_i += U_ij e_j (i < j) --
//            F-1 independent accumulators, U read as broadcast LDS.128 -> 4 FFMA per LDS, no dependent
//            chains; second = sum_i e_i t_i.  A pruned field matrix (few live pairs) is walked as the
//            compacted pair list instead.
//   phase E  fixed-order reduction over k -> shallow[b] = first + second + bias
//
// Bytes per sample the algorithm needs (Criteo, fwlw): 26*8 + 13*4 + 26*40 + 4 = 1304.
#include "embed_device.cuh"
#include "fused_wide.cuh"

namespace dfw {

constexpr int kS = 16;  // samples per CTA of the stand-alone kernel

struct SmemLayout { int EP; size_t oImg, oE, oPart, oIdx, oXv, total; };
__host__ __device__ inline SmemLayout smem_layout(int F, int K, int num) {
    SmemLayout L;
    const TileSizes t = tile_sizes(F, K, num, kS);
    L.EP = t.EP;
    size_t o = 0;
    L.oImg = o;  o += img_layout(F, K).total;
    L.oE = o;    o += t.bE;
    L.oPart = o; o += t.bPart;
    L.oIdx = o;  o += t.bIdx;
    L.oXv = o;   o += t.bXv;
    L.total = o;
    return L;
}

// ------------------------------------------------------------------------------------------ pack kernel
struct PackParams {
    const dfw_field_desc* fields; const float* wl; const float* fm1; const float* cov;
    unsigned char* image; int F, K; unsigned flags;
};

__global__ void __launch_bounds__(256) pack_shallow_kernel(const PackParams p) {
    const int F = p.F, K = p.K, tid = threadIdx.x, lane = tid & 31;
    const ImgLayout L = img_layout(F, K);
    ImgHeader* hdr = reinterpret_cast<ImgHeader*>(p.image + L.oHdr);
    float* U = reinterpret_cast<float*>(p.image + L.oU);
    PairEnt* pairs = reinterpret_cast<PairEnt*>(p.image + L.oPairs);
    float* wl = reinterpret_cast<float*>(p.image + L.oWl);
    uint64_t* fd = reinterpret_cast<uint64_t*>(p.image + L.oFields);
    const bool fwfm = p.flags & DFW_USE_FWFM;
    __shared__ int s_live, s_qr, s_mis;
    if (tid == 0) { s_live = 0; s_qr = 0; s_mis = 0; }
    __syncthreads();
    // field descriptors verbatim
    for (int i = tid; i < F * (int)(sizeof(dfw_field_desc) / 8); i += blockDim.x)
        fd[i] = reinterpret_cast<const uint64_t*>(p.fields)[i];
    if (tid < F) {
        const dfw_field_desc d = p.fields[tid];
        if (d.qr_op != DFW_TABLE_PLAIN || d.n_ranks > 1) s_qr = 1;
        bool mis = (reinterpret_cast<uintptr_t>(d.w2) & 7) != 0;
        for (int r = 0; r < d.n_ranks && r < DFW_MAX_RANKS; ++r) mis |= (reinterpret_cast<uintptr_t>(d.w2_shard[r]) & 7) != 0;
        if (mis) s_mis = 1;
    }
    // fwlw weights with the use_lw projection folded in: first = sum_f fm_1st[f] <E_f, wl_f>  (model/DeepFMs.py:344-345, 450)
    for (int i = tid; i < F * K; i += blockDim.x) {
        float w = 0.f;
        if (p.flags & DFW_USE_FWLW) w = p.wl[i] * ((p.flags & DFW_USE_LW) ? p.fm1[i / K] : 1.0f);
        wl[i] = w;
    }
    // U by columns, zero padded; (W.t() + W) * 0.5 in fp32 as the reference does (model/DeepFMs.py:364)
    int live = 0;
    for (int j = 1; j < F; ++j) {
        for (int i = tid; i < pad4(j); i += blockDim.x) {
            float u = 0.f;
            if (i < j) u = fwfm ? (p.cov[j * F + i] + p.cov[i * F + j]) * 0.5f : 1.0f;
            U[ucol_off(j) + i] = u;
            live += (u != 0.f);
        }
    }
    if (tid < 4) U[usize(F) + tid] = 0.f;
    if (live) atomicAdd(&s_live, live);
    // ordered compaction of the live pairs by one warp -> deterministic summation order
    if (tid < 32) {
        int n = 0;
        for (int i = 0; i < F - 1; ++i) {
            for (int jb = i + 1; jb < F; jb += 32) {
                const int j = jb + lane;
                float u = 0.f;
                if (j < F) u = fwfm ? (p.cov[j * F + i] + p.cov[i * F + j]) * 0.5f : 1.0f;
                const unsigned m = __ballot_sync(0xffffffffu, u != 0.f);
                if (u != 0.f) {
                    const int pos = n + __popc(m & ((1u << lane) - 1u));
                    pairs[pos].ij = (uint32_t)(i * K) | ((uint32_t)(j * K) << 16);
                    pairs[pos].u = u;
                }
                n += __popc(m);
            }
        }
        if (lane == 0) hdr->n_list = n;
    }
    __syncthreads();
    if (tid == 0) { hdr->live = s_live; hdr->any_special = s_qr; hdr->misaligned = s_mis; }
}

// ------------------------------------------------------------------------------------------ main kernel
// FT/KT > 0: compile-time field count / embedding width (dense unrolled second order available).
// CTA = 16 samples x K columns: thread t owns (sample t & 15, column t >> 4) in phase D.
template <int FT, int KT>
__global__ void __launch_bounds__(FT > 0 ? ((kS * KT + 31) / 32) * 32 : 512)
embed_fwfm_kernel(const EmbedParams p) {
    const int F = FT > 0 ? FT : p.F;
    const int K = KT > 0 ? KT : p.K;
    const int FK = F * K;
    const int tid = threadIdx.x;
    const int nthreads = blockDim.x;
    const int64_t b0 = (int64_t)blockIdx.x * kS;
    const int nrows = (int)min((int64_t)kS, p.B - b0);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const SmemLayout L = smem_layout(F, K, p.num);
    TileSmem sm;
    sm.img = smem_raw + L.oImg;
    sm.E = reinterpret_cast<float*>(smem_raw + L.oE);
    sm.part = reinterpret_cast<float*>(smem_raw + L.oPart);
    sm.idx = reinterpret_cast<int32_t*>(smem_raw + L.oIdx);
    sm.xv = reinterpret_cast<float*>(smem_raw + L.oXv);
    sm.EP = L.EP;
    const float* sE = sm.E;
    const int EP = L.EP;
    long long* clk = p.clk ? p.clk + blockIdx.x * 8 : nullptr;
    if (clk && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); clk[7] = (long long)t; }

    float first_acc[1];
    embed_gather<FT, KT, kS, 1, 0>(p, sm, smem_raw + L.oImg, true, tid, nthreads, b0, nrows, first_acc, clk);

    // ------------------------------------------------------------------ phase C: stream E out
    if (p.E) {
        if ((p.ldE & 1) == 0 && (reinterpret_cast<uintptr_t>(p.E) & 7) == 0) {      // float2 path (rows 8-byte aligned)
            const int ld2 = (int)(p.ldE >> 1);
            float2* dst = reinterpret_cast<float2*>(p.E + b0 * p.ldE);
            DivStep st(tid, nthreads, ld2);
#pragma unroll 2
            for (int i = tid; i < nrows * ld2; i += nthreads) {
                const int c0 = 2 * (int)st.r;
                float2 v = make_float2(0.f, 0.f);
                if (c0 + 1 < FK) v = *reinterpret_cast<const float2*>(sE + st.q * EP + c0);
                else if (c0 < FK) v.x = sE[st.q * EP + c0];
                dst[i] = v;
                st.next();
            }
        } else {
            const int ld = (int)p.ldE;
            float* dst = p.E + b0 * p.ldE;
            DivStep st(tid, nthreads, ld);
#pragma unroll 2
            for (int i = tid; i < nrows * ld; i += nthreads) {
                dst[i] = (int)st.r < FK ? sE[st.q * EP + st.r] : 0.f;
                st.next();
            }
        }
    }
    if (p.Eb) {
        const int ld2 = (int)(p.ldEb >> 1);
        __nv_bfloat162* dst = reinterpret_cast<__nv_bfloat162*>(p.Eb + b0 * p.ldEb);
        DivStep st(tid, nthreads, ld2);
#pragma unroll 2
        for (int i = tid; i < nrows * ld2; i += nthreads) {
            const int c0 = 2 * (int)st.r;
            float2 v = make_float2(0.f, 0.f);
            if (c0 + 1 < FK) v = *reinterpret_cast<const float2*>(sE + st.q * EP + c0);
            else if (c0 < FK) v.x = sE[st.q * EP + c0];
            dst[i] = __floats2bfloat162_rn(v.x, v.y);
            st.next();
        }
    }

    embed_interact<FT, KT, kS, 1, 0>(p, sm, tid, nthreads, nrows, first_acc, p.shallow + b0, clk);
}

// ------------------------------------------------------------------------------------------ register-gather kernel
// The shallow part only (no E output: FM / FwFM models without the deep part, BASELINE config 1) for the dataset shapes, with the
// gather of the fused kernel (fused_wide.cuh): a thread owns one embedding column of one sample, ten lanes cover a 40-byte row,
// three rows per warp instruction -- every row load of a sample is in flight at once, whole sectors, no shared-memory block, no
// cp.async pieces, no fix-up pass -- and the first + second order run from the registers with the field matrix as constant-bank
// operands.  Persistent grid-stride loop; 2 CTAs x 8 warps per SM keep ~13 k row loads in flight per SM.  HBM-bound: 1304
// algorithmic bytes per sample (two 32-byte sectors per 40-byte row: 0.68 of peak is the ceiling of the algorithmic fraction).
constexpr int RG_WARPS = 8, RG_THREADS = 32 * RG_WARPS;

template <int FT, int KT, int NUMT>
__global__ void __launch_bounds__(RG_THREADS, 2) embed_reg_kernel(const __grid_constant__ fz::UParam up, const EmbedParams p) {
    using namespace fz::wd;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const ImgLayout IL = img_layout(FT, KT);
    constexpr uint32_t W_WL = 16, W_FIELDS = W_WL + (uint32_t)up16(sizeof(float) * FT * KT);
    constexpr uint32_t W_NUM = W_FIELDS + (uint32_t)up16(sizeof(dfw_field_desc) * FT);
    unsigned char* sImg = smem_raw;
    float* sNum = reinterpret_cast<float*>(smem_raw + W_NUM);
    if (tid == 0) cp_async16(sImg, p.image + IL.oHdr);
    for (uint32_t i = tid; i < (uint32_t)(up16(sizeof(float) * FT * KT) >> 4); i += RG_THREADS) cp_async16(sImg + W_WL + 16 * i, p.image + IL.oWl + 16 * i);
    for (uint32_t i = tid; i < (uint32_t)(up16(sizeof(dfw_field_desc) * FT) >> 4); i += RG_THREADS) cp_async16(sImg + W_FIELDS + 16 * i, p.image + IL.oFields + 16 * i);
    cp_async_wait_all();
    __syncthreads();
    const ImgHeader* hdr = reinterpret_cast<const ImgHeader*>(sImg);
    const dfw_field_desc* sF = reinterpret_cast<const dfw_field_desc*>(sImg + W_FIELDS);
    const float* sWl = reinterpret_cast<const float*>(sImg + W_WL);
    int gmode = GM_PLAIN, qop = DFW_TABLE_PLAIN;
    if (hdr->any_special) gmode = wide_gmode<FT, NUMT>(sF, qop);
    for (int i = tid; i < NUMT * KT; i += RG_THREADS) {
        const int f = i / KT, k = i - f * KT;
        sNum[i] = __ldg((gmode == GM_PLAIN ? sF[f].w2 : locate_row(sF[f], 0, KT)) + k);
    }
    __syncthreads();
    constexpr int CT = FT - NUMT > 0 ? FT - NUMT : 1, SPW = 32 / KT, PER_CTA = RG_WARPS * SPW;
    const int sl = lane / KT, kk = lane - sl * KT;
    const bool fwlw = p.flags & DFW_USE_FWLW;
    const float bias = __ldg(p.bias);
    for (int64_t base = (int64_t)blockIdx.x * PER_CTA; base < p.B; base += (int64_t)gridDim.x * PER_CTA) {
        const int64_t b = base + warp * SPW + sl;
        const bool live = sl < SPW && b < p.B;
        float e[FT];
        uint32_t ix[CT];
        wide_idx<CT>(p, b, live, ix);
        if (gmode == GM_PLAIN) wide_rows<FT, KT, NUMT, GM_PLAIN>(p, sF, sNum, b, live, kk, qop, ix, e);
        else if (gmode == GM_SHARD) wide_rows<FT, KT, NUMT, GM_SHARD>(p, sF, sNum, b, live, kk, qop, ix, e);
        else if (gmode == GM_QR) wide_rows<FT, KT, NUMT, GM_QR>(p, sF, sNum, b, live, kk, qop, ix, e);
        else wide_rows<FT, KT, NUMT, GM_GENERIC>(p, sF, sNum, b, live, kk, qop, ix, e);
        float v = 0.f;
        if (live) {
            const float first = fwlw ? 0.f : wide_first<FT, KT, NUMT>(p, sF, b, live, kk);
            v = wide_interact<FT, KT>(e, up, sWl, kk, fwlw, first);
        }
        // fixed-order sum over the sample's K lanes (the order of embed_interact's phase E), + bias
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < KT; ++k) tot += __shfl_sync(0xffffffffu, v, (sl < SPW ? sl : 0) * KT + k);
        if (live && kk == 0) p.shallow[b] = tot + bias;
    }
}

template <int FT, int KT, int NUMT>
static int launch_embed_reg(const fz::UParam& up, const EmbedParams& p, cudaStream_t st) {
    auto kern = embed_reg_kernel<FT, KT, NUMT>;
    const size_t smem = 16 + up16(sizeof(float) * FT * KT) + up16(sizeof(dfw_field_desc) * FT) + up16(sizeof(float) * (NUMT > 0 ? NUMT : 1) * KT);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    constexpr int PER_CTA = RG_WARPS * (32 / KT);
    const long long want = (p.B + PER_CTA - 1) / PER_CTA;
    const unsigned grid = (unsigned)(want < 2LL * sms ? want : 2LL * sms);
    kern<<<grid, RG_THREADS, smem, st>>>(up, p);
    count_launch();
    return check_launch("embed_reg_kernel");
}

template <int FT, int KT>
static int launch_embed(const EmbedParams& p, cudaStream_t st) {
    const SmemLayout L = smem_layout(p.F, p.K, p.num);
    auto kern = embed_fwfm_kernel<FT, KT>;
    DFW_REQUIRE(L.total <= 227 * 1024, DFW_E_UNSUPPORTED, "embed kernel needs %zu B of shared memory", L.total);
    if (L.total > 48 * 1024)
        DFW_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    const int threads = ((kS * p.K + 31) / 32) * 32;
    const unsigned grid = (unsigned)((p.B + kS - 1) / kS);
    kern<<<grid, threads, L.total, st>>>(p);
    count_launch();
    return check_launch("embed_fwfm_kernel");
}

}  // namespace dfw

namespace dfw { static long long* g_clk = nullptr; }
// Debug tooling (not part of the product ABI): per-CTA clock64() stamps at the phase boundaries of the next
// embed_fwfm launches are written to `dev_buf` (8 x int64 per CTA); NULL switches it off.
extern "C" void dfw_debug_set_clock_buffer(void* dev_buf) { dfw::g_clk = static_cast<long long*>(dev_buf); }

extern "C" size_t dfw_shallow_image_bytes(const dfw_model* m) {
    if (!m || m->field_size < 1 || m->embedding_size < 1) return 0;
    return dfw::img_layout(m->field_size, m->embedding_size).total;
}

extern "C" int dfw_pack_shallow(const dfw_model* m, void* image, void* stream) {
    using namespace dfw;
    if (int rc = check_model(m)) return rc;
    DFW_REQUIRE(image && (reinterpret_cast<uintptr_t>(image) & 15) == 0, DFW_E_ARG, "image must be 16-byte aligned");
    DFW_REQUIRE((int64_t)m->field_size * m->embedding_size < 65536, DFW_E_UNSUPPORTED, "F*K must be < 65536");
    PackParams p{m->fields, m->fwfm_linear, m->fm_1st, m->field_cov, static_cast<unsigned char*>(image),
                 m->field_size, m->embedding_size, m->flags};
    pack_shallow_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    count_launch();
    return check_launch("pack_shallow_kernel");
}

extern "C" int dfw_embed_fwfm(const dfw_model* m, const int64_t* xi, int64_t xi_stride_b, int64_t xi_stride_c,
                              const float* xv, int64_t xv_stride_b, int64_t xv_stride_c, int64_t B,
                              float* E_out, int64_t ldE, void* E_bf16_out, int64_t ldEb,
                              float* shallow_out, int32_t* err_word, void* stream) {
    dfw::NvtxRange nvtx_("FM - Component: FM FW LW + FM Outer FwFM + FM Second Order (dfw_embed_fwfm)");
    using namespace dfw;
    if (int rc = check_model(m)) return rc;
    DFW_REQUIRE(B >= 0, DFW_E_ARG, "negative batch");
    if (B == 0) return 0;
    const int F = m->field_size, K = m->embedding_size, num = m->numerical;
    DFW_REQUIRE(m->shallow_image, DFW_E_ARG, "model has no shallow image (call dfw_pack_shallow)");
    DFW_REQUIRE(shallow_out, DFW_E_ARG, "shallow_out is NULL");
    DFW_REQUIRE(F - num == 0 || xi, DFW_E_ARG, "xi is NULL");
    DFW_REQUIRE(num == 0 || xv, DFW_E_ARG, "xv is NULL");
    DFW_REQUIRE(!E_out || ldE >= (int64_t)F * K, DFW_E_ARG, "ldE %lld < F*K", (long long)ldE);
    DFW_REQUIRE(!E_bf16_out || (ldEb >= (int64_t)F * K && ldEb % 2 == 0), DFW_E_ARG, "ldEb must be even and >= F*K");
    EmbedParams p;
    p.image = static_cast<const unsigned char*>(m->shallow_image);
    p.xi = xi; p.xi_sb = xi_stride_b; p.xi_sc = xi_stride_c;
    p.xv = xv; p.xv_sb = xv_stride_b; p.xv_sc = xv_stride_c;
    p.fm1 = m->fm_1st; p.bias = m->bias;
    p.E = E_out; p.ldE = ldE; p.Eb = reinterpret_cast<__nv_bfloat16*>(E_bf16_out); p.ldEb = ldEb;
    p.shallow = shallow_out; p.err = (m->flags & DFW_CHECK_INDEX) ? err_word : nullptr;
    p.B = B; p.F = F; p.num = num; p.K = K; p.flags = m->flags; p.clk = g_clk;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // shallow part only, dataset shapes, field matrix available as a kernel parameter: the register-gather kernel
    if (!E_out && !E_bf16_out && K == 10 && ((F == 39 && num == 13) || (F == 47 && num == 11))) {
        static thread_local fz::UParam up;
        fz::build_uparam(m, up);
        if (up.valid) {
            if (F == 39) return launch_embed_reg<39, 10, 13>(up, p, st);
            return launch_embed_reg<47, 10, 11>(up, p, st);
        }
    }
    // the two dataset shapes BASELINE.json names get the fully unrolled dense second order
    if (F == 39 && K == 10) return launch_embed<39, 10>(p, st);
    if (F == 47 && K == 10) return launch_embed<47, 10>(p, st);
    return launch_embed<0, 0>(p, st);
}
