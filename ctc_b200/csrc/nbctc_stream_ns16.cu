// Instantiations of the block-streaming kernel for NS = 16 states per chain lane (Lmax <= 256).
#include "stream_kernel.cuh"
namespace nbctc {
int launch_stream_ns16(const Problem& p, const StreamCfg& cfg, cudaStream_t stream) { return stream::launch_ns<16>(p, cfg, stream); }
}  // namespace nbctc
