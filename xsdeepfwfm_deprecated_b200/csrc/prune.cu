// One-shot magnitude pruning on the device: the reference's `binary_search_threshold`
// (model/DeepFMs.py:807-823) and the masking block of `fit` (model/DeepFMs.py:647-673) without its per-probe host round trips.
//
// The reference bisects t in [0, 100]: every probe is a full pass `(abs(param) < mid).sum().item()` -- a kernel, a reduction and a
// device->host synchronisation -- up to 101 times per tensor.  Here the whole bisection of one tensor set is ONE cooperative
// launch: a persistent grid (one CTA per SM) counts |w| < (float)mid over its slice (float4 loads; the 53 MB of Criteo tables
// stay L2-resident between probes), meets at a grid barrier, and every CTA then takes the same bisection step in fp64 registers
// -- the same doubles the Python loop computes, so the threshold is bit-identical to the reference's.  HBM/L2-bound integer
// work: 4 bytes per element per probe.
//
// `sym` mode is the field_cov recipe (model/DeepFMs.py:667-673): values are |0.5 (R + R^T)| of one (F, F) matrix.
#include "dfw_common.cuh"

namespace dfw {
namespace pr {

constexpr int MAX_SPANS = 2 * DFW_MAX_FIELDS;      // every table of a QR model is a (quotient, remainder) pair
constexpr int THREADS = 256;

struct Spans {
    int n;
    int sym_F;                       // > 0: one (F, F) matrix, values |0.5 (R + R^T)|
    dfw_prune_span s[MAX_SPANS];
};

struct State {                       // workspace, zero-initialised before the launch
    unsigned long long count[3];     // rotating per-probe counters
    unsigned int arrived;            // grid barrier
    unsigned int generation;
    int probes;
    int pad_;
    double threshold;
};

__device__ __forceinline__ void grid_barrier(State* st, unsigned int nblocks, int* err) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned int* gen = &st->generation;
        const unsigned int g = *gen;
        __threadfence();
        if (atomicAdd(&st->arrived, 1u) == nblocks - 1) {
            st->arrived = 0;
            __threadfence();
            *gen = g + 1;
        } else {
            for (unsigned long long spin = 0; *gen == g; ++spin)
                if (spin > (1ull << 31)) {            // only a broken launch (grid not co-resident) gets here
                    if (err) atomicExch(err, 77);
                    __trap();
                }
        }
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ unsigned int below(float v, float t) { return fabsf(v) < t ? 1u : 0u; }

// elements of this CTA's share of one span with |w| < t
__device__ __forceinline__ unsigned int count_span(const float* __restrict__ p, long long n, float t) {
    unsigned int c = 0;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        const long long n4 = n >> 2;
        const float4* p4 = reinterpret_cast<const float4*>(p);
        long long i = gtid;
        for (; i + 3 * gsz < n4; i += 4 * gsz) {      // four independent 16-byte loads in flight per thread
            const float4 a = p4[i], b = p4[i + gsz], d = p4[i + 2 * gsz], e = p4[i + 3 * gsz];
            c += below(a.x, t) + below(a.y, t) + below(a.z, t) + below(a.w, t);
            c += below(b.x, t) + below(b.y, t) + below(b.z, t) + below(b.w, t);
            c += below(d.x, t) + below(d.y, t) + below(d.z, t) + below(d.w, t);
            c += below(e.x, t) + below(e.y, t) + below(e.z, t) + below(e.w, t);
        }
        for (; i < n4; i += gsz) {
            const float4 a = p4[i];
            c += below(a.x, t) + below(a.y, t) + below(a.z, t) + below(a.w, t);
        }
        for (long long j = (n4 << 2) + gtid; j < n; j += gsz) c += below(p[j], t);
    } else {
        for (long long j = gtid; j < n; j += gsz) c += below(p[j], t);
    }
    return c;
}

__global__ void __launch_bounds__(THREADS)
bisect_kernel(const Spans sp, const double target, const long long total, State* st, double* threshold_out, int* probes_out,
              int* err) {
    __shared__ unsigned int warp_sums[THREADS / 32];
    __shared__ unsigned long long s_count;
    double lo = 0.0, hi = 1e2, mid = 0.0;
    int probes = 0;
    while (lo < hi) {                                             // model/DeepFMs.py:810
        ++probes;
        mid = (lo + hi) / 2;
        const float t = (float)mid;                               // abs(param) < mid compares in the tensor's dtype
        const int slot = probes % 3;
        if (blockIdx.x == 0 && threadIdx.x == 0) st->count[(probes + 1) % 3] = 0;    // the next probe's counter (last read two barriers ago)
        unsigned int c = 0;
        if (sp.sym_F > 0) {
            const float* R = sp.s[0].ptr;
            const int F = sp.sym_F;
            for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)F * F; e += (long long)gridDim.x * blockDim.x) {
                const int i = (int)(e / F), j = (int)(e - (long long)i * F);
                c += below(0.5f * (R[i * F + j] + R[j * F + i]), t);
            }
        } else {
            for (int s = 0; s < sp.n; ++s) c += count_span(sp.s[s].ptr, sp.s[s].count, t);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long b = 0;
            for (int w = 0; w < THREADS / 32; ++w) b += warp_sums[w];
            if (b) atomicAdd(&st->count[slot], b);
        }
        grid_barrier(st, gridDim.x, err);
        if (threadIdx.x == 0) s_count = *reinterpret_cast<volatile unsigned long long*>(&st->count[slot]);
        __syncthreads();
        const double rate = (double)s_count / (double)total;      // sparse_items * 1.0 / total_no
        if (fabs(rate - target) < 0.0001) break;                  // :816
        else if (rate > target) hi = mid;
        else lo = mid;
        if (probes > 100) break;                                  // :822
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *threshold_out = mid;
        if (probes_out) *probes_out = probes;
    }
}

// w = 0 where |w| < (float)threshold (model/DeepFMs.py:660-666); zeroed_out += newly or already masked elements
__global__ void __launch_bounds__(THREADS)
apply_kernel(const Spans sp, const double* __restrict__ threshold, unsigned long long* zeroed_out) {
    const float t = (float)*threshold;
    unsigned int c = 0;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
    for (int s = 0; s < sp.n; ++s) {
        float* p = sp.s[s].ptr;
        const long long n = sp.s[s].count;
        if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
            float4* p4 = reinterpret_cast<float4*>(p);
            const long long n4 = n >> 2;
            for (long long i = gtid; i < n4; i += gsz) {
                float4 a = p4[i];
                const unsigned int m = below(a.x, t) | (below(a.y, t) << 1) | (below(a.z, t) << 2) | (below(a.w, t) << 3);
                if (m) {
                    if (m & 1) a.x = 0.f;
                    if (m & 2) a.y = 0.f;
                    if (m & 4) a.z = 0.f;
                    if (m & 8) a.w = 0.f;
                    p4[i] = a;
                    c += __popc(m);
                }
            }
            for (long long j = (n4 << 2) + gtid; j < n; j += gsz)
                if (below(p[j], t)) { p[j] = 0.f; ++c; }
        } else {
            for (long long j = gtid; j < n; j += gsz)
                if (below(p[j], t)) { p[j] = 0.f; ++c; }
        }
    }
    if (zeroed_out) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(zeroed_out, (unsigned long long)c);
    }
}

// field_cov: R[i][j] = 0 where |0.5 (R + R^T)[i][j]| < threshold (model/DeepFMs.py:667-673).  One CTA: each thread owns the
// unordered pairs {(i, j), (j, i)} it zeroes, so nobody reads an element another thread may already have cleared.
__global__ void __launch_bounds__(THREADS)
apply_sym_kernel(float* R, int F, const double* __restrict__ threshold, unsigned long long* zeroed_out) {
    const float t = (float)*threshold;
    unsigned int c = 0;
    for (int e = threadIdx.x; e < F * F; e += blockDim.x) {
        const int i = e / F, j = e - i * F;
        if (i > j) continue;
        if (below(0.5f * (R[i * F + j] + R[j * F + i]), t)) {
            R[i * F + j] = 0.f;
            R[j * F + i] = 0.f;
            c += i == j ? 1 : 2;
        }
    }
    if (zeroed_out && c) atomicAdd(zeroed_out, (unsigned long long)c);
}

static int fill_spans(Spans& sp, const dfw_prune_span* spans, int n_spans, int sym_F) {
    DFW_REQUIRE(spans && n_spans >= 1 && n_spans <= MAX_SPANS, DFW_E_ARG, "n_spans %d outside [1,%d]", n_spans, MAX_SPANS);
    sp.n = n_spans;
    sp.sym_F = sym_F;
    for (int i = 0; i < n_spans; ++i) {
        DFW_REQUIRE(spans[i].count >= 0 && (spans[i].ptr || spans[i].count == 0), DFW_E_ARG, "span %d is NULL or negative", i);
        sp.s[i] = spans[i];
    }
    if (sym_F > 0) {
        DFW_REQUIRE(n_spans == 1 && sym_F <= DFW_MAX_FIELDS && spans[0].count == (int64_t)sym_F * sym_F, DFW_E_ARG,
                    "symmetric mode takes one (F, F) matrix with F <= %d", DFW_MAX_FIELDS);
    }
    return 0;
}

}  // namespace pr
}  // namespace dfw

using namespace dfw;

extern "C" size_t dfw_prune_workspace_bytes(void) { return align_up(sizeof(pr::State), 256); }

extern "C" int dfw_prune_threshold(const dfw_prune_span* spans, int n_spans, int sym_F, double target, int64_t total,
                                   void* workspace, size_t workspace_bytes, double* threshold_out, int32_t* probes_out,
                                   void* stream) {
    pr::Spans sp;
    if (int rc = pr::fill_spans(sp, spans, n_spans, sym_F)) return rc;
    DFW_REQUIRE(total > 0 && threshold_out, DFW_E_ARG, "total must be positive and threshold_out non-NULL");
    DFW_REQUIRE(workspace && workspace_bytes >= sizeof(pr::State), DFW_E_WORKSPACE, "prune workspace too small: %zu < %zu",
                workspace_bytes, sizeof(pr::State));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int dev = 0, sms = 0, coop = 0, per_sm = 0;
    DFW_CUDA_OK(cudaGetDevice(&dev));
    DFW_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    DFW_CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    DFW_REQUIRE(coop, DFW_E_NODEVICE, "device %d cannot launch cooperative grids", dev);
    DFW_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pr::bisect_kernel, pr::THREADS, 0));
    DFW_REQUIRE(per_sm >= 1, DFW_E_UNSUPPORTED, "bisect_kernel does not fit an SM");
    int64_t elems = 0;
    for (int i = 0; i < n_spans; ++i) elems += spans[i].count;
    // a persistent grid of up to 4 CTAs per SM (the probes of an L2-resident tensor are latency-bound: bytes in flight matter),
    // fewer for small tensors (a CTA should have at least ~4 K elements to count); all CTAs must be co-resident
    const int cap = sms * (per_sm < 4 ? per_sm : 4);
    long long want = (elems + 4095) / 4096;
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    DFW_CUDA_OK(cudaMemsetAsync(workspace, 0, sizeof(pr::State), st));
    pr::State* state = static_cast<pr::State*>(workspace);
    long long total_ll = total;
    int* err = nullptr;
    void* args[] = {&sp, &target, &total_ll, &state, &threshold_out, &probes_out, &err};
    DFW_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(pr::bisect_kernel), dim3((unsigned)grid), dim3(pr::THREADS),
                                            args, 0, st));
    count_launch();
    return check_launch("bisect_kernel");
}

extern "C" int dfw_prune_apply(const dfw_prune_span* spans, int n_spans, int sym_F, const double* threshold,
                               uint64_t* zeroed_out, void* stream) {
    pr::Spans sp;
    if (int rc = pr::fill_spans(sp, spans, n_spans, sym_F)) return rc;
    DFW_REQUIRE(threshold, DFW_E_ARG, "threshold is NULL");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (sym_F > 0) {
        pr::apply_sym_kernel<<<1, pr::THREADS, 0, st>>>(sp.s[0].ptr, sym_F, threshold, reinterpret_cast<unsigned long long*>(zeroed_out));
        count_launch();
        return check_launch("apply_sym_kernel");
    }
    int64_t elems = 0;
    for (int i = 0; i < n_spans; ++i) elems += spans[i].count;
    if (elems == 0) return 0;
    long long want = (elems + 4 * pr::THREADS - 1) / (4 * pr::THREADS);
    const int grid = (int)(want > 148 * 8 ? 148 * 8 : want);
    pr::apply_kernel<<<grid, pr::THREADS, 0, st>>>(sp, threshold, reinterpret_cast<unsigned long long*>(zeroed_out));
    count_launch();
    return check_launch("apply_kernel");
}
