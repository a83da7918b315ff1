// K1/K2 tensor-core path (tcgen05 + TMA).  Placeholder until the kernel lands: AUTO falls back
// to the simple path, FKV_IMPL_TENSOR reports FKV_EMODE.
#include "common.cuh"

namespace fkv {
bool score_tensor_supported(const fkv_score_params&) { return false; }
int launch_score_tensor(const fkv_score_params&, cudaStream_t) { return FKV_EMODE; }
}  // namespace fkv
