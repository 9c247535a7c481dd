// Stage 2, fp32 parity path: the deep MLP on CUDA cores (FFMA), one launch per Linear layer with
// bias + ReLU fused, then the fc dot + total + sigmoid epilogue.
//
// Replaces model/DeepFMs.py:408-436 and :458 of the reference (Linear -> ReLU chain, Dropout is the
// identity in eval()).  Weights are read IN PLACE from the module's nn.Linear parameters:
// W (out, in) row-major, so both operands are K-contiguous ("NT" GEMM).
//
// Tile: 64 (samples) x 80 (neurons) x 16 (k) per 256-thread CTA, 4 x 5 outputs per thread,
// operands staged k-major in shared memory (A read as one LDS.128, W as 5 conflict-free LDS.32 per k),
// register double buffering of the global loads.  N = 400 is 5 exact tiles of 80.
#include "dfw_common.cuh"

namespace dfw {

constexpr int BM = 64, BN = 80, BK = 16, NTH = 256;
constexpr int APITCH = BM + 4, WPITCH = BN + 1;

__global__ void __launch_bounds__(NTH)
linear_relu_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, const float* __restrict__ bias,
                   float* __restrict__ C, int64_t ldc, int64_t M, int N, int Kd, int relu) {
    __shared__ __align__(16) float As[2][BK][APITCH];
    __shared__ float Ws[2][BK][WPITCH];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // global -> register staging assignments
    const int a_r = tid >> 2, a_c = (tid & 3) * 4;            // A: row a_r, k a_c..a_c+3
    const bool a_vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool w_vec = ((Kd & 3) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    float a_reg[4], w_reg[2][4];

    auto load_tiles = [&](int k0) {
        {
            const int64_t m = m0 + a_r;
            const int k = k0 + a_c;
            if (m < M && a_vec && k + 3 < Kd) {
                const float4 t = *reinterpret_cast<const float4*>(A + m * lda + k);
                a_reg[0] = t.x; a_reg[1] = t.y; a_reg[2] = t.z; a_reg[3] = t.w;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) a_reg[i] = (m < M && k + i < Kd) ? A[m * lda + k + i] : 0.f;
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int slot = tid + h * NTH;                    // 320 float4 slots: 80 rows x 4
            if (slot < BN * 4) {
                const int n = n0 + (slot >> 2), k = k0 + (slot & 3) * 4;
                if (n < N && w_vec && k + 3 < Kd) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * Kd + k));
                    w_reg[h][0] = t.x; w_reg[h][1] = t.y; w_reg[h][2] = t.z; w_reg[h][3] = t.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) w_reg[h][i] = (n < N && k + i < Kd) ? __ldg(W + (int64_t)n * Kd + k + i) : 0.f;
                }
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 4; ++i) As[buf][a_c + i][a_r] = a_reg[i];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int slot = tid + h * NTH;
            if (slot < BN * 4) {
#pragma unroll
                for (int i = 0; i < 4; ++i) Ws[buf][(slot & 3) * 4 + i][slot >> 2] = w_reg[h][i];
            }
        }
    };

    float acc[4][5];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) acc[i][j] = 0.f;

    const int nk = (Kd + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int t = 0; t < nk; ++t) {
        const int buf = t & 1;
        if (t + 1 < nk) load_tiles((t + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            float w[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) w[j] = Ws[buf][k][tx * 5 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 5; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        if (t + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const int n = n0 + tx * 5 + j;
            if (n < N) {
                float v = acc[i][j] + __ldg(bias + n);
                if (relu) v = fmaxf(v, 0.f);
                C[m * ldc + n] = v;
            }
        }
    }
}

// logits[b] = shallow[b] + <H[b,:], fc>  (model/DeepFMs.py:428, 458); prob = sigmoid (:777).  One warp per sample.
__global__ void __launch_bounds__(256)
fc_total_kernel(const float* __restrict__ H, int64_t ldh, int N, const float* __restrict__ fc,
                const float* __restrict__ shallow, float* __restrict__ logits, float* __restrict__ prob, int64_t B) {
    const int lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B) return;
    float s = 0.f;
    for (int n = lane; n < N; n += 32) s = fmaf(H[b * ldh + n], __ldg(fc + n), s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const float z = s + (shallow ? shallow[b] : 0.f);
        if (logits) logits[b] = z;
        if (prob) prob[b] = 1.0f / (1.0f + expf(-z));
    }
}

__global__ void finish_shallow_kernel(const float* __restrict__ shallow, float* __restrict__ logits,
                                      float* __restrict__ prob, int64_t B) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float z = shallow[b];
    if (logits && logits != shallow) logits[b] = z;
    if (prob) prob[b] = 1.0f / (1.0f + expf(-z));
}

int launch_fc_total(const float* H, int64_t ldh, int N, const float* fc, const float* shallow,
                    float* logits, float* prob, int64_t B, cudaStream_t st) {
    const int wpb = 8;
    fc_total_kernel<<<(unsigned)((B + wpb - 1) / wpb), wpb * 32, 0, st>>>(H, ldh, N, fc, shallow, logits, prob, B);
    count_launch();
    return check_launch("fc_total_kernel");
}

}  // namespace dfw

extern "C" size_t dfw_mlp_workspace_bytes(const dfw_model* m, int64_t B, int precision) {
    if (!m || !(m->flags & DFW_USE_DEEP) || B <= 0) return 256;
    int wmax = 0;
    for (int l = 0; l < m->depth; ++l) wmax = wmax > m->widths[l] ? wmax : m->widths[l];
    const int64_t Bp = (B + 127) / 128 * 128;
    if (precision == DFW_PREC_BF16) return 4096;   // activations never leave the SM
    // two ping-pong activation buffers (B, wmax_pad) fp32
    const size_t one = dfw::align_up((size_t)Bp * ((wmax + 3) / 4 * 4) * sizeof(float), 256);
    return 2 * one + 256;
}

extern "C" int dfw_mlp_fp32(const dfw_model* m, const float* X, int64_t ldX, int64_t B, const float* shallow,
                            void* workspace, size_t workspace_bytes, float* logits_out, float* prob_out,
                            void* stream) {
    dfw::NvtxRange nvtx_("Deep - Component (dfw_mlp_fp32)");
    using namespace dfw;
    if (int rc = check_model(m)) return rc;
    DFW_REQUIRE(m->flags & DFW_USE_DEEP, DFW_E_ARG, "model has no deep part");
    DFW_REQUIRE(X && (logits_out || prob_out), DFW_E_ARG, "X / outputs NULL");
    if (B <= 0) return 0;
    DFW_REQUIRE(workspace && workspace_bytes >= dfw_mlp_workspace_bytes(m, B, DFW_PREC_FP32), DFW_E_WORKSPACE,
                "mlp workspace too small: %zu < %zu", workspace_bytes, dfw_mlp_workspace_bytes(m, B, DFW_PREC_FP32));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int wmax = 0;
    for (int l = 0; l < m->depth; ++l) wmax = wmax > m->widths[l] ? wmax : m->widths[l];
    const int64_t Bp = (B + 127) / 128 * 128;
    const int ldh = (wmax + 3) / 4 * 4;
    const size_t one = align_up((size_t)Bp * ldh * sizeof(float), 256);
    float* buf[2] = {reinterpret_cast<float*>(workspace), reinterpret_cast<float*>(static_cast<char*>(workspace) + one)};
    const float* in = X;
    int64_t ldin = ldX;
    int in_dim = m->field_size * m->embedding_size;
    for (int l = 0; l < m->depth; ++l) {
        const int N = m->widths[l];
        float* out = buf[l & 1];
        dim3 grid((unsigned)((B + BM - 1) / BM), (unsigned)((N + BN - 1) / BN));
        linear_relu_kernel<<<grid, NTH, 0, st>>>(in, ldin, m->W[l], m->b[l], out, ldh, B, N, in_dim, 1);
        count_launch();
        if (int rc = check_launch("linear_relu_kernel")) return rc;
        in = out; ldin = ldh; in_dim = N;
    }
    return launch_fc_total(in, ldin, in_dim, m->fc, shallow, logits_out, prob_out, B, st);
}

extern "C" int dfw_finish_shallow(const float* shallow, int64_t B, float* logits_out, float* prob_out, void* stream) {
    using namespace dfw;
    DFW_REQUIRE(shallow, DFW_E_ARG, "shallow is NULL");
    if (B <= 0) return 0;
    finish_shallow_kernel<<<(unsigned)((B + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        shallow, logits_out, prob_out, B);
    count_launch();
    return check_launch("finish_shallow_kernel");
}
