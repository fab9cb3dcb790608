// Instantiations of the fused kernel for NS = 2 states per chain lane (Lmax <= 32).
#include "fused_kernel.cuh"
namespace nbctc {
int launch_fused_ns2(const Problem& p, const FusedCfg& cfg, cudaStream_t stream) { return fused::launch_ns<2>(p, cfg, stream); }
}  // namespace nbctc
