// Stage 2, fp32 CSR path for magnitude-pruned MLPs (DeepLight: 90 % of net_1_linear_* zeroed in place,
// model/DeepFMs.py:662-666).  The reference's PyTorch inference multiplies the zeros densely; its C++
// side-car walks a CRS matrix per sample (latency/criteo_latency.cpp:144-170).  Here the whole
// Linear/ReLU chain + fc + total runs in ONE kernel per batch tile with activations resident in SMEM:
//
//   CTA   = 64 samples (lane l owns samples l and l+32), 16 warps
//   SMEM  = two ping-pong activation tiles [64][pitch] fp32, pitch odd -> the per-lane column read
//           sIn[lane][col] with a warp-uniform col is conflict-free
//   warp w computes neurons n = w, w+16, ...: walks row n of the CSR image (warp-uniform (col,val)
//           loads served by L1), 2 FFMA per non-zero per lane, writes relu(acc + bias) to the other tile
//   after the last layer each warp dots its neurons with fc, a fixed-order cross-warp sum gives deep[b]
//
// Useful FLOPs at 90 % sparsity: 2 * 47.6 k per sample instead of 2 * 476 k.
#include "dfw_common.cuh"

namespace dfw {

constexpr int CS_WARPS = 16;

struct CsrParams {
    int depth, in_dim;
    int widths[DFW_MAX_DEPTH];
    dfw_csr csr[DFW_MAX_DEPTH];
    const float* bias[DFW_MAX_DEPTH];
    const float* fc;
    const float* X; int64_t ldX;
    const float* shallow; float* logits; float* prob;
    int64_t B; int pitch;
};

// S = samples per lane: 2 (64-sample tile) when two activation tiles fit in shared memory, else 1.
template <int S>
__global__ void __launch_bounds__(CS_WARPS * 32, 1)
csr_mlp_kernel(const CsrParams p) {
    constexpr int CS_SAMPLES = 32 * S;
    extern __shared__ __align__(16) float smem[];
    float* tile[2] = {smem, smem + CS_SAMPLES * p.pitch};
    float* sRed = smem + 2 * CS_SAMPLES * p.pitch;          // [CS_WARPS][64]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t b0 = (int64_t)blockIdx.x * CS_SAMPLES;
    const int nrows = (int)min((int64_t)CS_SAMPLES, p.B - b0);
    const int pitch = p.pitch;

    // input tile, coalesced
    for (int i = threadIdx.x; i < CS_SAMPLES * p.in_dim; i += blockDim.x) {
        const int s = i / p.in_dim, c = i - s * p.in_dim;
        tile[0][s * pitch + c] = s < nrows ? p.X[(b0 + s) * p.ldX + c] : 0.f;
    }
    __syncthreads();

    float d0 = 0.f, d1 = 0.f;    // running fc dot of this warp's neurons (last layer only)
    for (int l = 0; l < p.depth; ++l) {
        const float* in = tile[l & 1];
        float* out = tile[(l & 1) ^ 1];
        const dfw_csr M = p.csr[l];
        const bool last = (l == p.depth - 1);
        const float* in0 = in + lane * pitch;
        const float* in1 = in + (lane + (S > 1 ? 32 : 0)) * pitch;
        for (int n = warp; n < p.widths[l]; n += CS_WARPS) {
            const int beg = __ldg(M.row_ptr + n), end = __ldg(M.row_ptr + n + 1);
            float a0 = 0.f, a1 = 0.f, c0 = 0.f, c1 = 0.f;
            int t = beg;
            for (; t + 1 < end; t += 2) {
                const int ca = __ldg(M.col + t), cb = __ldg(M.col + t + 1);
                const float va = __ldg(M.val + t), vb = __ldg(M.val + t + 1);
                a0 = fmaf(va, in0[ca], a0); a1 = fmaf(va, in1[ca], a1);
                c0 = fmaf(vb, in0[cb], c0); c1 = fmaf(vb, in1[cb], c1);
            }
            if (t < end) {
                const int ca = __ldg(M.col + t);
                const float va = __ldg(M.val + t);
                a0 = fmaf(va, in0[ca], a0); a1 = fmaf(va, in1[ca], a1);
            }
            const float bn = __ldg(p.bias[l] + n);
            const float h0 = fmaxf(a0 + c0 + bn, 0.f), h1 = fmaxf(a1 + c1 + bn, 0.f);
            if (last) {
                const float w = __ldg(p.fc + n);
                d0 = fmaf(h0, w, d0); d1 = fmaf(h1, w, d1);
            } else {
                out[lane * pitch + n] = h0;
                if (S > 1) out[(lane + 32) * pitch + n] = h1;
            }
        }
        __syncthreads();
    }
    sRed[warp * CS_SAMPLES + lane] = d0;
    if (S > 1) sRed[warp * CS_SAMPLES + lane + 32] = d1;
    __syncthreads();
    if (threadIdx.x < CS_SAMPLES && threadIdx.x < nrows) {
        float z = 0.f;
        for (int w = 0; w < CS_WARPS; ++w) z += sRed[w * CS_SAMPLES + threadIdx.x];
        const int64_t b = b0 + threadIdx.x;
        if (p.shallow) z += p.shallow[b];
        if (p.logits) p.logits[b] = z;
        if (p.prob) p.prob[b] = 1.0f / (1.0f + expf(-z));
    }
}

// one warp per output row: count non-zeros; block 0 then scans (out_dim is a few hundred)
__global__ void csr_count_kernel(const float* __restrict__ W, int out_dim, int in_dim, int32_t* __restrict__ row_ptr) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= out_dim) return;
    int c = 0;
    for (int k = lane; k < in_dim; k += 32) c += (W[(int64_t)row * in_dim + k] != 0.f);
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) row_ptr[row + 1] = c;
}
__global__ void csr_scan_kernel(int32_t* row_ptr, int out_dim) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int acc = 0;
        row_ptr[0] = 0;
        for (int r = 1; r <= out_dim; ++r) { acc += row_ptr[r]; row_ptr[r] = acc; }
    }
}
__global__ void csr_fill_kernel(const float* __restrict__ W, int out_dim, int in_dim, const int32_t* __restrict__ row_ptr,
                                int32_t* __restrict__ col, float* __restrict__ val) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= out_dim) return;
    int pos = row_ptr[row];
    for (int k0 = 0; k0 < in_dim; k0 += 32) {
        const int k = k0 + lane;
        const float w = k < in_dim ? W[(int64_t)row * in_dim + k] : 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, w != 0.f);
        if (w != 0.f) {
            const int q = pos + __popc(m & ((1u << lane) - 1u));
            col[q] = k;
            val[q] = w;
        }
        pos += __popc(m);
    }
}

}  // namespace dfw

using namespace dfw;

extern "C" int dfw_csr_count(const float* W, int32_t out_dim, int32_t in_dim, int32_t* row_ptr, void* stream) {
    DFW_REQUIRE(W && row_ptr && out_dim > 0 && in_dim > 0, DFW_E_ARG, "bad csr_count arguments");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    csr_count_kernel<<<(out_dim + 7) / 8, 256, 0, st>>>(W, out_dim, in_dim, row_ptr);
    csr_scan_kernel<<<1, 32, 0, st>>>(row_ptr, out_dim);
    count_launch(2);
    return check_launch("csr_count");
}

extern "C" int dfw_csr_fill(const float* W, int32_t out_dim, int32_t in_dim, const int32_t* row_ptr,
                            int32_t* col, float* val, void* stream) {
    DFW_REQUIRE(W && row_ptr && col && val && out_dim > 0 && in_dim > 0, DFW_E_ARG, "bad csr_fill arguments");
    csr_fill_kernel<<<(out_dim + 7) / 8, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(W, out_dim, in_dim, row_ptr, col, val);
    count_launch();
    return check_launch("csr_fill");
}

extern "C" int dfw_mlp_csr(const dfw_model* m, const float* X, int64_t ldX, int64_t B, const float* shallow,
                           void* workspace, size_t workspace_bytes, float* logits_out, float* prob_out, void* stream) {
    dfw::NvtxRange nvtx_("Deep - Component, CSR (dfw_mlp_csr)");
    (void)workspace; (void)workspace_bytes;
    if (int rc = check_model(m)) return rc;
    DFW_REQUIRE(m->flags & DFW_USE_DEEP, DFW_E_ARG, "model has no deep part");
    DFW_REQUIRE(X && (logits_out || prob_out), DFW_E_ARG, "X / outputs NULL");
    if (B <= 0) return 0;
    CsrParams p;
    p.depth = m->depth;
    p.in_dim = m->field_size * m->embedding_size;
    int wmax = p.in_dim;
    for (int l = 0; l < m->depth; ++l) {
        DFW_REQUIRE(m->csr[l].row_ptr && m->csr[l].col && m->csr[l].val, DFW_E_ARG,
                    "layer %d has no CSR image (call dfw_csr_count/dfw_csr_fill)", l + 1);
        p.widths[l] = m->widths[l];
        p.csr[l] = m->csr[l];
        p.bias[l] = m->b[l];
        wmax = wmax > m->widths[l] ? wmax : m->widths[l];
    }
    p.fc = m->fc; p.X = X; p.ldX = ldX; p.shallow = shallow; p.logits = logits_out; p.prob = prob_out; p.B = B;
    p.pitch = wmax | 1;
    auto smem_for = [&](int samples) { return sizeof(float) * ((size_t)2 * samples * p.pitch + CS_WARPS * samples); };
    const int S = smem_for(64) <= 227 * 1024 ? 2 : 1;
    const size_t smem = smem_for(32 * S);
    DFW_REQUIRE(smem <= 227 * 1024, DFW_E_UNSUPPORTED, "CSR MLP: layer width %d needs %zu B of shared memory (> 227 KiB)", wmax, smem);
    auto kern = S == 2 ? csr_mlp_kernel<2> : csr_mlp_kernel<1>;
    DFW_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int samples = 32 * S;
    kern<<<(unsigned)((B + samples - 1) / samples), CS_WARPS * 32, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    count_launch();
    return check_launch("csr_mlp_kernel");
}
