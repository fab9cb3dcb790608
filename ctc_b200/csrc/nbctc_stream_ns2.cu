// Instantiations of the block-streaming kernel for NS = 2 states per chain lane (Lmax <= 32).
#include "stream_kernel.cuh"
namespace nbctc {
int launch_stream_ns2(const Problem& p, const StreamCfg& cfg, cudaStream_t stream) { return stream::launch_ns<2>(p, cfg, stream); }
}  // namespace nbctc
