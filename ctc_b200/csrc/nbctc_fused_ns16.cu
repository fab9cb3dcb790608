// Instantiations of the fused kernel for NS = 16 states per chain lane (Lmax <= 256).
#include "fused_kernel.cuh"
namespace nbctc {
int launch_fused_ns16(const Problem& p, const FusedCfg& cfg, cudaStream_t stream) { return fused::launch_ns<16>(p, cfg, stream); }
}  // namespace nbctc
