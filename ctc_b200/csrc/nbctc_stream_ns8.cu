// Instantiations of the block-streaming kernel for NS = 8 states per chain lane (Lmax <= 128).
#include "stream_kernel.cuh"
namespace nbctc {
int launch_stream_ns8(const Problem& p, const StreamCfg& cfg, cudaStream_t stream) { return stream::launch_ns<8>(p, cfg, stream); }
}  // namespace nbctc
