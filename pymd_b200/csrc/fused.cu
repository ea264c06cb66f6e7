// Fused on-chip step kernel (mode 2) -- placeholder until the per-step pipeline is parity-green.
#include "engine.cuh"

namespace pymd {
int fused_supported(const pymd_ctx*) { return 0; }
int fused_step(pymd_ctx*, long long, int, cudaStream_t) {
    pymd_set_error("fused path not built");
    return PYMD_ESTATE;
}
}  // namespace pymd
