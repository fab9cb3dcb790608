// Instantiations of the fused kernel for NS = 8 states per chain lane (Lmax <= 128).
#include "fused_kernel.cuh"
namespace nbctc {
int launch_fused_ns8(const Problem& p, const FusedCfg& cfg, cudaStream_t stream) { return fused::launch_ns<8>(p, cfg, stream); }
}  // namespace nbctc
