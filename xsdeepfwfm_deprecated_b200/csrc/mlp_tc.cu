// Stage 2, tcgen05 / TMA bf16 path (placeholder until the tensor-core kernel lands).
#include "dfw_common.cuh"
using namespace dfw;
extern "C" size_t dfw_pack_mlp_bf16_bytes(int32_t out_dim, int32_t in_dim) {
    return (size_t)((out_dim + 15) / 16 * 16) * ((in_dim + 63) / 64 * 64) * 2;
}
extern "C" int dfw_pack_mlp_bf16(const float*, int32_t, int32_t, void*, void*) {
    set_error("bf16 tensor path not built yet");
    return DFW_E_UNSUPPORTED;
}
extern "C" int dfw_mlp_bf16(const dfw_model*, const void*, int64_t, int64_t, const float*, void*, size_t, float*, float*, void*) {
    set_error("bf16 tensor path not built yet");
    return DFW_E_UNSUPPORTED;
}
