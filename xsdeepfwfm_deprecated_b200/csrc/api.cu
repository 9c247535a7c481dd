// C-ABI glue of libdeepfwfm_sm100a: error strings, device check, whole-forward entry points, shard helpers.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "dfw_common.cuh"

namespace dfw {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct FwdLayout {
    int ldE, ldEb;
    int64_t Bp;
    size_t oErr, oImg, oE, oEb, oShallow, oMlp, total;
};

// bf16 on a shape neither the fused kernel nor the staged tensor MLP (F*K <= 512, widths <= 512) takes runs the CUDA-core fp32 MLP
// (stricter, slower) -- the same fallback bf16x3 has
static bool bf16_staged_ok(const dfw_model* m) {
    if (m->field_size * m->embedding_size > 512) return false;
    for (int l = 0; l < m->depth; ++l)
        if (m->widths[l] > 512) return false;
    return true;
}

static FwdLayout fwd_layout(const dfw_model* m, int64_t B, int precision) {
    if (precision == DFW_PREC_BF16 && (m->flags & DFW_USE_DEEP) && !bf16_staged_ok(m)) precision = DFW_PREC_FP32;
    FwdLayout L;
    const int FK = m->field_size * m->embedding_size;
    L.ldE = (FK + 3) / 4 * 4;
    L.ldEb = (FK + 7) / 8 * 8;
    L.Bp = (B + 127) / 128 * 128;
    const bool deep = m->flags & DFW_USE_DEEP;
    size_t o = 0;
    L.oErr = o; o += 256;
    L.oImg = o; if (!m->shallow_image) o += align_up(dfw_shallow_image_bytes(m), 256);
    L.oE = o;   if (deep && precision != DFW_PREC_BF16) o += align_up((size_t)L.Bp * L.ldE * sizeof(float), 256);
    L.oEb = o;  if (deep && precision == DFW_PREC_BF16) o += align_up((size_t)L.Bp * L.ldEb * 2, 256);
    L.oShallow = o; o += align_up((size_t)L.Bp * sizeof(float), 256);
    L.oMlp = o; o += deep ? align_up(dfw_mlp_workspace_bytes(m, B, precision), 256) : 0;
    L.total = o;
    return L;
}

__global__ void shard_rows_kernel(const float* __restrict__ src, int64_t rows, int width, int rank, int n_ranks,
                                  float* __restrict__ dst) {
    const int64_t local_rows = (rows - rank + n_ranks - 1) / n_ranks;
    const int64_t total = local_rows * width;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / width;
        const int c = (int)(i - r * width);
        dst[i] = src[(r * n_ranks + rank) * width + c];
    }
}

__global__ void gather_rows_kernel(const float* __restrict__ shard, int64_t shard_rows, int width,
                                   const int64_t* __restrict__ req, int64_t n, int n_ranks, float* __restrict__ out) {
    const int64_t total = n * width;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / width;
        const int c = (int)(i - r * width);
        int64_t local = req[r] / n_ranks;
        if (local < 0 || local >= shard_rows) local = 0;
        out[i] = __ldg(shard + local * width + c);
    }
}

// Mapped host transport, stage 1: a few small CTAs pull one batch's Xi / Xv from pinned host memory over PCIe into a device
// staging slot (16-byte loads, two in flight per thread), ahead of the fused kernel that consumes it.  128 threads x 32
// registers: such a CTA fits beside a resident fused CTA (640 x 96 registers), so the copy needs no SM of its own.
__device__ __forceinline__ void pull_bytes(const void* __restrict__ src, void* __restrict__ dst, size_t bytes, uint32_t tid, uint32_t nth) {
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | bytes) & 15) == 0) {
        const uint4* s4 = static_cast<const uint4*>(src);
        uint4* d4 = static_cast<uint4*>(dst);
        const uint32_t n = (uint32_t)(bytes >> 4);       // a batch is far below 2^32 x 16 bytes
        uint32_t i = tid;
        for (; i + nth < n; i += 2 * nth) {
            const uint4 a = s4[i], b = s4[i + nth];
            d4[i] = a; d4[i + nth] = b;
        }
        if (i < n) d4[i] = s4[i];
    } else {                                            // ragged tails / odd batch sizes: elements are 4 or 8 bytes wide
        const uint32_t* s1 = static_cast<const uint32_t*>(src);
        uint32_t* d1 = static_cast<uint32_t*>(dst);
        for (size_t i = tid; i < (bytes >> 2); i += nth) d1[i] = s1[i];
    }
}
__global__ void __launch_bounds__(128, 16) stage_inputs_kernel(const void* __restrict__ xi_src, void* __restrict__ xi_dst, size_t xi_bytes,
                                                           const void* __restrict__ xv_src, void* __restrict__ xv_dst, size_t xv_bytes) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    if (xi_bytes) pull_bytes(xi_src, xi_dst, xi_bytes, tid, nth);
    if (xv_bytes) pull_bytes(xv_src, xv_dst, xv_bytes, tid, nth);
}

}  // namespace dfw

using namespace dfw;

extern "C" int dfw_version(void) { return DFW_ABI_VERSION; }
extern "C" const char* dfw_last_error_string(void) { return g_err; }
extern "C" size_t dfw_struct_bytes(int which) {
    return which == 0 ? sizeof(dfw_model) : which == 1 ? sizeof(dfw_field_desc) : which == 2 ? sizeof(dfw_csr) : 0;
}
extern "C" int64_t dfw_launch_count(void) { return (int64_t)g_launches.load(); }

extern "C" int dfw_check_device(int ordinal) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        set_error("no CUDA device visible (%s); this library has no CPU fallback", cudaGetErrorString(e));
        return DFW_E_NODEVICE;
    }
    DFW_REQUIRE(ordinal >= 0 && ordinal < n, DFW_E_ARG, "device ordinal %d outside [0,%d)", ordinal, n);
    cudaDeviceProp prop;
    DFW_CUDA_OK(cudaGetDeviceProperties(&prop, ordinal));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libdeepfwfm_sm100a carries sm_100a code only", ordinal, prop.major, prop.minor);
        return DFW_E_NODEVICE;
    }
    return 0;
}

extern "C" size_t dfw_forward_workspace_bytes(const dfw_model* m, int64_t B, int precision) {
    if (!m || B <= 0) return 256;
    return fwd_layout(m, B, precision).total;
}

extern "C" int dfw_forward(const dfw_model* m, const int64_t* xi, int64_t xi_stride_b, int64_t xi_stride_c,
                           const float* xv, int64_t xv_stride_b, int64_t xv_stride_c, int64_t B, int precision,
                           void* workspace, size_t workspace_bytes, float* logits_out, float* prob_out,
                           int32_t* err_word, void* stream) {
    dfw::NvtxRange nvtx_("DeepFMs.forward (dfw_forward)");
    if (int rc = check_model(m)) return rc;
    DFW_REQUIRE(B >= 0, DFW_E_ARG, "negative batch");
    if (B == 0) return 0;
    DFW_REQUIRE(logits_out || prob_out, DFW_E_ARG, "no output requested");
    DFW_REQUIRE(precision == DFW_PREC_FP32 || precision == DFW_PREC_BF16 || precision == DFW_PREC_FP32_CSR ||
                precision == DFW_PREC_BF16X3, DFW_E_ARG, "unknown precision %d", precision);
    const FwdLayout L = fwd_layout(m, B, precision);
    DFW_REQUIRE(workspace && workspace_bytes >= L.total, DFW_E_WORKSPACE, "forward workspace too small: %zu < %zu",
                workspace_bytes, L.total);
    DFW_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, DFW_E_ARG, "workspace must be 256-byte aligned");
    char* ws = static_cast<char*>(workspace);
    const bool deep = m->flags & DFW_USE_DEEP;
    const int requested = precision;          // what the fused kernel is asked for; the staged kernels may fall back to fp32
    if (precision == DFW_PREC_BF16 && deep && !bf16_staged_ok(m)) precision = DFW_PREC_FP32;
    float* E = (deep && precision != DFW_PREC_BF16) ? reinterpret_cast<float*>(ws + L.oE) : nullptr;
    void* Eb = (deep && precision == DFW_PREC_BF16) ? static_cast<void*>(ws + L.oEb) : nullptr;
    float* shallow = reinterpret_cast<float*>(ws + L.oShallow);
    dfw_model local;
    if (!err_word) err_word = reinterpret_cast<int32_t*>(ws + L.oErr);
    if (!m->shallow_image) {   // "always fresh" mode: rebuild the shallow image from the live parameters per call
        local = *m;
        local.shallow_image = ws + L.oImg;
        if (int rc = dfw_pack_shallow(m, ws + L.oImg, stream)) return rc;
        m = &local;
    }
    // one kernel for the whole forward when the tensor-core form fits the model's shapes
    static const bool no_fused = dbg_getenv("DFW_NO_FUSED") != nullptr;
    if (deep && !no_fused && (requested == DFW_PREC_BF16 || requested == DFW_PREC_BF16X3) && dfw_fused_supported(m, requested))
        return dfw_forward_fused(m, xi, xi_stride_b, xi_stride_c, xv, xv_stride_b, xv_stride_c, B, requested, logits_out,
                                 prob_out, err_word, stream);
    if (int rc = dfw_embed_fwfm(m, xi, xi_stride_b, xi_stride_c, xv, xv_stride_b, xv_stride_c, B, E, L.ldE, Eb, L.ldEb,
                                shallow, err_word, stream))
        return rc;
    if (!deep) return dfw_finish_shallow(shallow, B, logits_out, prob_out, stream);
    void* mws = ws + L.oMlp;
    const size_t mbytes = workspace_bytes - L.oMlp;
    if (precision == DFW_PREC_BF16) return dfw_mlp_bf16(m, Eb, L.ldEb, B, shallow, mws, mbytes, logits_out, prob_out, stream);
    if (precision == DFW_PREC_FP32_CSR) return dfw_mlp_csr(m, E, L.ldE, B, shallow, mws, mbytes, logits_out, prob_out, stream);
    // DFW_PREC_FP32, and DFW_PREC_BF16X3 on shapes the fused kernel does not take: CUDA-core fp32 (stricter, slower)
    return dfw_mlp_fp32(m, E, L.ldE, B, shallow, mws, mbytes, logits_out, prob_out, stream);
}

namespace {
struct HostLayout { size_t oXi, oXv, oLogit, oProb, oFwd, total; };
HostLayout host_layout(const dfw_model* m, int64_t B, int precision) {
    HostLayout H;
    const int C = m->field_size - m->numerical;
    size_t o = 0;
    H.oXi = o;    o += align_up((size_t)B * (C > 0 ? C : 1) * sizeof(int64_t), 256);      // sized for int64; int32 uses half
    H.oXv = o;    o += align_up((size_t)B * (m->numerical > 0 ? m->numerical : 1) * sizeof(float), 256);
    H.oLogit = o; o += align_up((size_t)B * sizeof(float), 256);
    H.oProb = o;  o += align_up((size_t)B * sizeof(float), 256);
    H.oFwd = o;   o += fwd_layout(m, B, precision).total;
    H.total = o;
    return H;
}
}  // namespace

extern "C" size_t dfw_forward_host_workspace_bytes(const dfw_model* m, int64_t B, int precision) {
    if (!m || B <= 0) return 256;
    return host_layout(m, B, precision).total;
}

extern "C" int dfw_forward_host(const dfw_model* m, const int64_t* xi_host, const float* xv_host, int64_t B,
                                int precision, void* workspace, size_t workspace_bytes, float* logits_host,
                                float* prob_host, void* stream) {
    if (int rc = check_model(m)) return rc;
    if (B <= 0) return 0;
    DFW_REQUIRE(logits_host || prob_host, DFW_E_ARG, "no output requested");
    const HostLayout H = host_layout(m, B, precision);
    DFW_REQUIRE(workspace && workspace_bytes >= H.total, DFW_E_WORKSPACE, "host-forward workspace too small: %zu < %zu",
                workspace_bytes, H.total);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    const int C = m->field_size - m->numerical, num = m->numerical;
    int64_t* xi = reinterpret_cast<int64_t*>(ws + H.oXi);
    float* xv = reinterpret_cast<float*>(ws + H.oXv);
    float* logit = reinterpret_cast<float*>(ws + H.oLogit);
    float* prob = reinterpret_cast<float*>(ws + H.oProb);
    const size_t ib = (m->flags & DFW_XI_INT32) ? sizeof(int32_t) : sizeof(int64_t);
    if (C > 0) DFW_CUDA_OK(cudaMemcpyAsync(xi, xi_host, (size_t)B * C * ib, cudaMemcpyHostToDevice, st));
    if (num > 0) DFW_CUDA_OK(cudaMemcpyAsync(xv, xv_host, (size_t)B * num * sizeof(float), cudaMemcpyHostToDevice, st));
    if (int rc = dfw_forward(m, xi, C, 1, xv, num, 1, B, precision, ws + H.oFwd, workspace_bytes - H.oFwd,
                             logits_host ? logit : nullptr, prob_host ? prob : nullptr, nullptr, stream))
        return rc;
    if (logits_host) DFW_CUDA_OK(cudaMemcpyAsync(logits_host, logit, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (prob_host) DFW_CUDA_OK(cudaMemcpyAsync(prob_host, prob, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, st));
    DFW_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

// ---- streamed host-buffer inference --------------------------------------------------------------
// N samples in pinned host memory, processed as batches of `batch` round-robin over kSlots internal streams: batch i+1's
// H2D copy overlaps batch i's kernels and batch i-1's D2H, one host synchronisation at the end.  This is what
// eval_by_batch / predict_proba do around forward (model/DeepFMs.py:750-784, 864-873) without their per-batch
// .cuda() / .cpu() round trips.
namespace {
constexpr int kSlots = 3;      // staged transport: slots of the device workspace = streams in rotation
constexpr int kLanes = 6;      // mapped transport: compute streams in rotation
constexpr int kStage = 6;      // mapped transport, pull-kernel form (debug builds): staging slots of one batch each
constexpr int kChunk = 8;      // mapped transport: batches per copy-engine transfer (a chunk); the first chunks of a call are 1, 2, 4
constexpr int kChunkSlots = 3; // mapped transport: chunk-sized staging slots in rotation
struct HostPipe {
    cudaStream_t streams[kLanes] = {};
    cudaEvent_t done[kLanes] = {};
    cudaEvent_t start = nullptr;
    cudaStream_t pull = nullptr;                    // high priority: the PCIe pull of batch i + 1 runs under the kernels of batch i
    cudaEvent_t staged[kStage] = {}, freed[kStage] = {};
    cudaStream_t copy = nullptr;                    // chunked H2D transfers of the mapped transport
    cudaEvent_t chunk_in[kChunkSlots] = {}, chunk_free[kChunkSlots][kLanes] = {};
    int device = -1;
};
thread_local HostPipe g_pipe[8];

int get_pipe(HostPipe** out) {
    int dev = 0;
    DFW_CUDA_OK(cudaGetDevice(&dev));
    DFW_REQUIRE(dev >= 0 && dev < 8, DFW_E_UNSUPPORTED, "device ordinal %d >= 8", dev);
    HostPipe& hp = g_pipe[dev];
    if (hp.device != dev) {
        for (int i = 0; i < kLanes; ++i) {
            DFW_CUDA_OK(cudaStreamCreateWithFlags(&hp.streams[i], cudaStreamNonBlocking));
            DFW_CUDA_OK(cudaEventCreateWithFlags(&hp.done[i], cudaEventDisableTiming));
        }
        DFW_CUDA_OK(cudaEventCreateWithFlags(&hp.start, cudaEventDisableTiming));
        int lo_prio = 0, hi_prio = 0;
        DFW_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
        DFW_CUDA_OK(cudaStreamCreateWithPriority(&hp.pull, cudaStreamNonBlocking, hi_prio));
        for (int i = 0; i < kStage; ++i) {
            DFW_CUDA_OK(cudaEventCreateWithFlags(&hp.staged[i], cudaEventDisableTiming));
            DFW_CUDA_OK(cudaEventCreateWithFlags(&hp.freed[i], cudaEventDisableTiming));
        }
        DFW_CUDA_OK(cudaStreamCreateWithPriority(&hp.copy, cudaStreamNonBlocking, hi_prio));
        for (int i = 0; i < kChunkSlots; ++i) {
            DFW_CUDA_OK(cudaEventCreateWithFlags(&hp.chunk_in[i], cudaEventDisableTiming));
            for (int l = 0; l < kLanes; ++l) DFW_CUDA_OK(cudaEventCreateWithFlags(&hp.chunk_free[i][l], cudaEventDisableTiming));
        }
        hp.device = dev;
    }
    *out = &hp;
    return 0;
}

// Device-visible alias of a host buffer: pinned (cudaHostAlloc / cudaHostRegister) memory under unified addressing can be
// loaded and stored by kernels directly over PCIe.  nullptr for pageable memory (which only a staged copy can move).
void* mapped_alias(const void* host) {
    if (!host) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

// 0 = choose (mapped when every buffer is device-visible and the fused kernel takes the model), 1 = staged copies, 2 = mapped,
// 3 = mapped without the pull stage (debug: the fused kernel's gather warps read Xi / Xv from host memory themselves)
int host_transport_override() {
    static const int v = [] {
        const char* e = dbg_getenv("DFW_HOST_TRANSPORT");
        if (!e) return 0;
        return !strcmp(e, "copy") ? 1 : !strcmp(e, "mapped") ? 2 : !strcmp(e, "mapped_direct") ? 3 : 0;
    }();
    return v;
}
}  // namespace

extern "C" int dfw_host_transport_is_mapped(const dfw_model* m, int precision, const void* xi_host, const void* xv_host,
                                            const void* logits_host, const void* prob_host) {
    if (!m || check_model(m)) return 0;
    if (host_transport_override() == 1 || dbg_getenv("DFW_NO_FUSED")) return 0;
    const int C = m->field_size - m->numerical, num = m->numerical;
    if (!(m->flags & DFW_USE_DEEP) || !(precision == DFW_PREC_BF16 || precision == DFW_PREC_BF16X3) ||
        !dfw_fused_supported(m, precision))
        return 0;
    if ((C > 0 && !mapped_alias(xi_host)) || (num > 0 && !mapped_alias(xv_host))) return 0;
    if ((logits_host && !mapped_alias(logits_host)) || (prob_host && !mapped_alias(prob_host))) return 0;
    return 1;
}

extern "C" size_t dfw_forward_host_stream_workspace_bytes(const dfw_model* m, int64_t batch, int precision) {
    if (!m || batch <= 0) return 256;
    const HostLayout H = host_layout(m, batch, precision);
    const size_t staged = kSlots * align_up(H.total, 256);
    const size_t mapped = (size_t)kChunkSlots * kChunk * H.oLogit;          // mapped transport: 3 staging slots of kChunk batches (Xi + Xv)
    return staged > mapped ? staged : mapped;
}

extern "C" int dfw_forward_host_stream(const dfw_model* m, const int64_t* xi_host, const float* xv_host, int64_t N,
                                       int64_t batch, int precision, void* workspace, size_t workspace_bytes,
                                       float* logits_host, float* prob_host, void* stream) {
    dfw::NvtxRange nvtx_("eval_by_batch / predict_proba (dfw_forward_host_stream)");
    if (int rc = check_model(m)) return rc;
    if (N <= 0) return 0;
    DFW_REQUIRE(batch > 0, DFW_E_ARG, "batch must be positive");
    DFW_REQUIRE(logits_host || prob_host, DFW_E_ARG, "no output requested");
    const HostLayout H = host_layout(m, batch, precision);
    const size_t slot_bytes = align_up(H.total, 256);
    DFW_REQUIRE(workspace && workspace_bytes >= kSlots * slot_bytes, DFW_E_WORKSPACE,
                "streamed host-forward workspace too small: %zu < %zu", workspace_bytes, kSlots * slot_bytes);
    DFW_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, DFW_E_ARG, "workspace must be 256-byte aligned");
    HostPipe* hp = nullptr;
    if (int rc = get_pipe(&hp)) return rc;
    cudaStream_t main_st = reinterpret_cast<cudaStream_t>(stream);
    const int C = m->field_size - m->numerical, num = m->numerical;
    // the internal streams start after everything already queued on the caller's stream (weights, images)
    DFW_CUDA_OK(cudaEventRecord(hp->start, main_st));
    for (int i = 0; i < kLanes; ++i) DFW_CUDA_OK(cudaStreamWaitEvent(hp->streams[i], hp->start, 0));
    const size_t ib = (m->flags & DFW_XI_INT32) ? sizeof(int32_t) : sizeof(int64_t);
    if (dfw_host_transport_is_mapped(m, precision, xi_host, xv_host, logits_host, prob_host)) {
        // Mapped transport: the host buffers are pinned, so kernels move the bytes over PCIe themselves -- no copy-engine round
        // trips.  A small pull kernel on a high-priority stream brings batch i's Xi / Xv into a staging slot while earlier batches
        // compute (a fused CTA that waited on PCIe for its indices would hold its SM idle); the fused kernel then gathers from
        // HBM and its epilogue stores logits / probabilities straight into host memory.  Two launches per batch.
        const char* xi_d = static_cast<const char*>(mapped_alias(xi_host));
        const char* xv_d = static_cast<const char*>(mapped_alias(xv_host));
        float* logit_d = static_cast<float*>(mapped_alias(logits_host));
        float* prob_d = static_cast<float*>(mapped_alias(prob_host));
        const bool direct = host_transport_override() == 3;
        const size_t stage_bytes = H.oLogit;            // the Xi + Xv part of a slot
        if (host_transport_override() == 0 && workspace_bytes >= 2 * stage_bytes) {
            // Default form.  Measured: SM-issued loads from pinned host memory saturate at ~34 GB/s on these boxes whatever the
            // kernel (the round-1 pull kernel, or the fused kernel reading host memory itself), while the copy engine reaches
            // ~55 GB/s -- once a transfer is a few MB, a single batch (1.06 MB) pays ~8 us of set-up per copy.  So Xi / Xv travel
            // in CHUNKS of up to kChunk batches (8.5 MB) on one copy stream into kChunkSlots rotating staging slots; the fused
            // kernels of a chunk's batches (one launch each, kLanes streams) wait on the chunk's event, read HBM and store their
            // probabilities / logits straight into the pinned host buffers.  The first chunks of a call are 1, 2 and 4 batches so
            // that the pipeline fills quickly.
            int chunk_cap = (int)(workspace_bytes / (kChunkSlots * stage_bytes));
            if (chunk_cap > kChunk) chunk_cap = kChunk;
            int nslots = kChunkSlots;
            if (chunk_cap < 1) { chunk_cap = 1; nslots = (int)(workspace_bytes / stage_bytes); if (nslots > kChunkSlots) nslots = kChunkSlots; }
            const size_t slot_bytes = (size_t)chunk_cap * stage_bytes;
            const size_t xi_b = (size_t)batch * C * ib, xv_b = (size_t)batch * num * sizeof(float);
            DFW_CUDA_OK(cudaStreamWaitEvent(hp->copy, hp->start, 0));
            int64_t done = 0, bi = 0;
            int lane_rr = 0;
            bool lane_touched[kChunkSlots][kLanes] = {};
            for (int64_t k = 0; done < N; ++k) {
                const int slot = (int)(k % nslots);
                const int64_t left = N - done;
                const int64_t left_b = (left + batch - 1) / batch;
                // chunk sizes ramp up 1, 2, 4, .. and taper off the same way (at most half of what is left): a short call pays a
                // one-batch copy before its first kernel and a one-batch compute after its last copy, not a chunk's worth of either
                int64_t g = k < 3 ? (1 << k) : chunk_cap;
                if (g > chunk_cap) g = chunk_cap;
                if (g > (left_b + 1) / 2) g = (left_b + 1) / 2;
                int64_t nb = left_b < g ? left_b : g;
                const int64_t rows = nb * batch < left ? nb * batch : left;
                char* ws = static_cast<char*>(workspace) + (size_t)slot * slot_bytes;
                char* xi_dev = ws;                                         // [nb x Xi][nb x Xv]: each batch's Xi / Xv contiguous
                char* xv_dev = ws + (size_t)chunk_cap * align_up(xi_b, 256);
                if (k >= nslots)                                            // the slot's previous chunk has been consumed by every lane that ran it
                    for (int l = 0; l < kLanes; ++l)
                        if (lane_touched[slot][l]) DFW_CUDA_OK(cudaStreamWaitEvent(hp->copy, hp->chunk_free[slot][l], 0));
                if (C > 0) DFW_CUDA_OK(cudaMemcpyAsync(xi_dev, xi_d + (size_t)done * C * ib, (size_t)rows * C * ib, cudaMemcpyHostToDevice, hp->copy));
                if (num > 0) DFW_CUDA_OK(cudaMemcpyAsync(xv_dev, xv_d + (size_t)done * num * sizeof(float), (size_t)rows * num * sizeof(float), cudaMemcpyHostToDevice, hp->copy));
                DFW_CUDA_OK(cudaEventRecord(hp->chunk_in[slot], hp->copy));
                for (int l = 0; l < kLanes; ++l) lane_touched[slot][l] = false;
                for (int64_t j = 0; j < nb; ++j, ++bi) {
                    const int64_t off = j * batch;
                    const int64_t b = rows - off < batch ? rows - off : batch;
                    const int ln = lane_rr;
                    lane_rr = (lane_rr + 1) % kLanes;
                    cudaStream_t lane = hp->streams[ln];
                    if (!lane_touched[slot][ln]) DFW_CUDA_OK(cudaStreamWaitEvent(lane, hp->chunk_in[slot], 0));
                    lane_touched[slot][ln] = true;
                    if (int rc = dfw_forward_fused(m, reinterpret_cast<const int64_t*>(xi_dev + (size_t)off * C * ib), C, 1,
                                                   reinterpret_cast<const float*>(xv_dev + (size_t)off * num * sizeof(float)), num, 1, b,
                                                   precision, logit_d ? logit_d + done + off : nullptr, prob_d ? prob_d + done + off : nullptr,
                                                   nullptr, lane))
                        return rc;
                }
                for (int l = 0; l < kLanes; ++l)
                    if (lane_touched[slot][l]) DFW_CUDA_OK(cudaEventRecord(hp->chunk_free[slot][l], hp->streams[l]));
                done += rows;
            }
            for (int i = 0; i < kLanes; ++i) {
                DFW_CUDA_OK(cudaEventRecord(hp->done[i], hp->streams[i]));
                DFW_CUDA_OK(cudaStreamWaitEvent(main_st, hp->done[i], 0));
            }
            DFW_CUDA_OK(cudaStreamSynchronize(main_st));
            return 0;
        }
        // as many staging slots as the workspace holds (it is sized for the staged transport's 3 full slots: at least 3 fit)
        const int nstage = (int)(workspace_bytes / stage_bytes < (size_t)kStage ? workspace_bytes / stage_bytes : (size_t)kStage);
        DFW_REQUIRE(direct || nstage >= 2, DFW_E_WORKSPACE, "workspace too small for two staging slots");
        if (!direct) DFW_CUDA_OK(cudaStreamWaitEvent(hp->pull, hp->start, 0));
        int64_t done = 0;
        for (int64_t i = 0; done < N; ++i, done += batch) {
            const int64_t b = N - done < batch ? N - done : batch;
            cudaStream_t lane = hp->streams[i % kLanes];
            const char* xi_src = xi_d + (size_t)done * C * ib;
            const char* xv_src = xv_d ? xv_d + (size_t)done * num * sizeof(float) : nullptr;
            const void* xi_in = xi_src;
            const void* xv_in = xv_src;
            const int slot = (int)(i % (direct ? kStage : nstage));
            if (!direct) {
                char* ws = static_cast<char*>(workspace) + (size_t)slot * stage_bytes;
                if (i >= nstage) DFW_CUDA_OK(cudaStreamWaitEvent(hp->pull, hp->freed[slot], 0));   // the slot's previous batch has been consumed
                static const int pull_ctas = dbg_getenv("DFW_PULL_CTAS") ? atoi(dbg_getenv("DFW_PULL_CTAS")) : 64;   // 64 x 128 threads x 32 B in flight
                stage_inputs_kernel<<<pull_ctas, 128, 0, hp->pull>>>(xi_src, ws + H.oXi, C > 0 ? (size_t)b * C * ib : 0,
                                                              xv_src, ws + H.oXv, num > 0 ? (size_t)b * num * sizeof(float) : 0);
                count_launch();
                if (int rc = check_launch("stage_inputs_kernel")) return rc;
                DFW_CUDA_OK(cudaEventRecord(hp->staged[slot], hp->pull));
                DFW_CUDA_OK(cudaStreamWaitEvent(lane, hp->staged[slot], 0));
                xi_in = ws + H.oXi;
                xv_in = ws + H.oXv;
            }
            if (int rc = dfw_forward_fused(m, static_cast<const int64_t*>(xi_in), C, 1, static_cast<const float*>(xv_in), num, 1, b,
                                           precision, logit_d ? logit_d + done : nullptr, prob_d ? prob_d + done : nullptr, nullptr,
                                           lane))
                return rc;
            if (!direct) DFW_CUDA_OK(cudaEventRecord(hp->freed[slot], lane));
        }
        for (int i = 0; i < kLanes; ++i) {
            DFW_CUDA_OK(cudaEventRecord(hp->done[i], hp->streams[i]));
            DFW_CUDA_OK(cudaStreamWaitEvent(main_st, hp->done[i], 0));
        }
        DFW_CUDA_OK(cudaStreamSynchronize(main_st));
        return 0;
    }
    DFW_REQUIRE(host_transport_override() < 2, DFW_E_UNSUPPORTED,
                "the mapped host transport was forced (debug build), but a host buffer is not pinned or the fused kernel does not take this model");
    int64_t done = 0;
    for (int64_t i = 0; done < N; ++i, done += batch) {
        const int64_t b = N - done < batch ? N - done : batch;
        const int slot = (int)(i % kSlots);
        cudaStream_t st = hp->streams[slot];
        char* ws = static_cast<char*>(workspace) + (size_t)slot * slot_bytes;
        int64_t* xi = reinterpret_cast<int64_t*>(ws + H.oXi);
        float* xv = reinterpret_cast<float*>(ws + H.oXv);
        float* logit = reinterpret_cast<float*>(ws + H.oLogit);
        float* prob = reinterpret_cast<float*>(ws + H.oProb);
        if (C > 0) DFW_CUDA_OK(cudaMemcpyAsync(xi, reinterpret_cast<const char*>(xi_host) + (size_t)done * C * ib, (size_t)b * C * ib, cudaMemcpyHostToDevice, st));
        if (num > 0) DFW_CUDA_OK(cudaMemcpyAsync(xv, xv_host + done * num, (size_t)b * num * sizeof(float), cudaMemcpyHostToDevice, st));
        if (int rc = dfw_forward(m, xi, C, 1, xv, num, 1, b, precision, ws + H.oFwd, slot_bytes - H.oFwd,
                                 logits_host ? logit : nullptr, prob_host ? prob : nullptr, nullptr, st))
            return rc;
        if (logits_host) DFW_CUDA_OK(cudaMemcpyAsync(logits_host + done, logit, (size_t)b * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (prob_host) DFW_CUDA_OK(cudaMemcpyAsync(prob_host + done, prob, (size_t)b * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < kSlots; ++i) {
        DFW_CUDA_OK(cudaEventRecord(hp->done[i], hp->streams[i]));
        DFW_CUDA_OK(cudaStreamWaitEvent(main_st, hp->done[i], 0));
    }
    DFW_CUDA_OK(cudaStreamSynchronize(main_st));
    return 0;
}

// ---- multi-GPU helpers -------------------------------------------------------------------------
extern "C" int dfw_shard_alloc(size_t bytes, void** dev_ptr) {
    DFW_REQUIRE(dev_ptr, DFW_E_ARG, "dev_ptr is NULL");
    DFW_CUDA_OK(cudaMalloc(dev_ptr, bytes ? bytes : 256));
    return 0;
}
extern "C" int dfw_shard_free(void* dev_ptr) {
    if (dev_ptr) DFW_CUDA_OK(cudaFree(dev_ptr));
    return 0;
}
extern "C" int dfw_ipc_export(const void* dev_ptr, uint8_t handle_out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    DFW_REQUIRE(dev_ptr && handle_out, DFW_E_ARG, "NULL argument");
    cudaIpcMemHandle_t h;
    DFW_CUDA_OK(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
    memcpy(handle_out, &h, 64);
    return 0;
}
extern "C" int dfw_ipc_import(const uint8_t handle[64], void** peer_ptr) {
    DFW_REQUIRE(handle && peer_ptr, DFW_E_ARG, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    DFW_CUDA_OK(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int dfw_ipc_close(void* peer_ptr) {
    if (peer_ptr) DFW_CUDA_OK(cudaIpcCloseMemHandle(peer_ptr));
    return 0;
}
extern "C" int dfw_shard_rows(const float* src, int64_t rows, int32_t width, int32_t rank, int32_t n_ranks,
                              float* dst, void* stream) {
    DFW_REQUIRE(src && dst && rows >= 0 && width > 0 && n_ranks > 0 && rank >= 0 && rank < n_ranks, DFW_E_ARG,
                "bad shard_rows arguments");
    if (rows == 0) return 0;
    shard_rows_kernel<<<592, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, rows, width, rank, n_ranks, dst);
    count_launch();
    return check_launch("shard_rows_kernel");
}
extern "C" int dfw_gather_rows(const float* shard, int64_t shard_rows, int32_t width, const int64_t* req, int64_t n,
                               int32_t n_ranks, float* out, void* stream) {
    DFW_REQUIRE(shard && out && width > 0 && n_ranks > 0 && n >= 0, DFW_E_ARG, "bad gather_rows arguments");
    if (n == 0) return 0;
    DFW_REQUIRE(req, DFW_E_ARG, "req is NULL");
    gather_rows_kernel<<<592, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(shard, shard_rows, width, req, n, n_ranks, out);
    count_launch();
    return check_launch("gather_rows_kernel");
}
