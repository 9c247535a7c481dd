// Shared helpers of libdeepfwfm_sm100a (sm_100a only; there is no other code path).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>
#include <nvtx3/nvToolsExt.h>

#include "deepfwfm_b200.h"

namespace dfw {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define DFW_CUDA_OK(expr)                                                              \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            dfw::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                           __FILE__, __LINE__);                                        \
            return (int)_e;                                                            \
        }                                                                              \
    } while (0)

#define DFW_REQUIRE(cond, code, ...)                                                   \
    do {                                                                               \
        if (!(cond)) {                                                                 \
            dfw::set_error(__VA_ARGS__);                                               \
            return (code);                                                             \
        }                                                                              \
    } while (0)

inline int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

inline int check_model(const dfw_model* m) {
    if (!m) { set_error("model is NULL"); return DFW_E_ARG; }
    if (m->struct_bytes != sizeof(dfw_model) || m->abi_version != DFW_ABI_VERSION) {
        set_error("dfw_model ABI mismatch: got %u bytes / v%u, library has %zu bytes / v%d",
                  m->struct_bytes, m->abi_version, sizeof(dfw_model), DFW_ABI_VERSION);
        return DFW_E_ABI;
    }
    if (m->field_size < 1 || m->field_size > DFW_MAX_FIELDS) { set_error("field_size %d outside [1,%d]", m->field_size, DFW_MAX_FIELDS); return DFW_E_UNSUPPORTED; }
    if (m->embedding_size < 1 || m->embedding_size > DFW_MAX_K) { set_error("embedding_size %d outside [1,%d]", m->embedding_size, DFW_MAX_K); return DFW_E_UNSUPPORTED; }
    if (m->numerical < 0 || m->numerical > m->field_size) { set_error("numerical %d outside [0,F]", m->numerical); return DFW_E_ARG; }
    if (!m->fields || !m->bias) { set_error("fields/bias pointer is NULL"); return DFW_E_ARG; }
    if ((m->flags & DFW_USE_FWFM) && !m->field_cov) { set_error("USE_FWFM without field_cov"); return DFW_E_ARG; }
    if ((m->flags & DFW_USE_FWLW) && !m->fwfm_linear) { set_error("USE_FWLW without fwfm_linear"); return DFW_E_ARG; }
    if ((m->flags & DFW_USE_LW) && !m->fm_1st) { set_error("USE_LW without fm_1st"); return DFW_E_ARG; }
    if (m->flags & DFW_USE_DEEP) {
        if (m->depth < 1 || m->depth > DFW_MAX_DEPTH) { set_error("depth %d outside [1,%d]", m->depth, DFW_MAX_DEPTH); return DFW_E_UNSUPPORTED; }
        for (int l = 0; l < m->depth; ++l)
            if (!m->W[l] || !m->b[l] || m->widths[l] < 1) { set_error("MLP layer %d incomplete", l + 1); return DFW_E_ARG; }
        if (!m->fc) { set_error("net_1_fc pointer is NULL"); return DFW_E_ARG; }
    }
    return 0;
}

__device__ __forceinline__ float sigmoidf_dev(float x) { return 1.0f / (1.0f + __expf(-x)); }

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// NVTX range for the duration of a C-ABI call, named after the region of the reference's forward it replaces (the
// torch.profiler.record_function names of model/DeepFMs.py:294-395), so that a timeline of this library lines up with one of the
// reference.  nvtx3 is header-only and a no-op (one predictable branch) unless a tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// Tuning / experiment knobs are read from the environment ONLY in a -DDFW_DEBUG build (python -m ...build --debug).  The
// shipped library reads no environment variable at all, so nothing outside the caller's arguments can change what a timed
// call does.
#ifdef DFW_DEBUG
inline const char* dbg_getenv(const char* name) { return getenv(name); }
#else
inline const char* dbg_getenv(const char*) { return nullptr; }
#endif

}  // namespace dfw
