// Instantiations of the block-streaming kernel for NS = 4 states per chain lane (Lmax <= 64).
#include "stream_kernel.cuh"
namespace nbctc {
int launch_stream_ns4(const Problem& p, const StreamCfg& cfg, cudaStream_t stream) { return stream::launch_ns<4>(p, cfg, stream); }
}  // namespace nbctc
