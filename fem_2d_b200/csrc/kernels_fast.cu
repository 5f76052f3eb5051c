// Re-ordered ("fast") integrators -- placeholders until implemented; the EXACT path is the product default.
#include <cuda_runtime.h>

#include "device_plan.hpp"

namespace fem2d {
cudaError_t launch_k2_sumfact(Plan&, uint32_t, uint32_t, uint32_t, uint32_t, cudaStream_t, uint32_t*) { return cudaErrorNotSupported; }
cudaError_t launch_k2_dmma(Plan&, uint32_t, uint32_t, uint32_t, uint32_t, cudaStream_t, uint32_t*) { return cudaErrorNotSupported; }
}  // namespace fem2d
