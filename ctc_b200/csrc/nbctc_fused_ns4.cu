// Instantiations of the fused kernel for NS = 4 states per chain lane (Lmax <= 64).
#include "fused_kernel.cuh"
namespace nbctc {
int launch_fused_ns4(const Problem& p, const FusedCfg& cfg, cudaStream_t stream) { return fused::launch_ns<4>(p, cfg, stream); }
}  // namespace nbctc
