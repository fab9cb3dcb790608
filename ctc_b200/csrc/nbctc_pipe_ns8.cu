// Explicit instantiation unit of the pipeline kernel for NS = 8 (Lmax <= 128); see pipe_kernel.cuh.
#include "pipe_kernel.cuh"

namespace nbctc {
int launch_pipe_ns8(const Problem& p, const PipeCfg& cfg, cudaStream_t stream) { return pipe::launch_pipe_ns<8>(p, cfg, stream); }
}  // namespace nbctc
