// Explicit instantiation unit of the pipeline kernel for NS = 4 (Lmax <= 64); see pipe_kernel.cuh.
#include "pipe_kernel.cuh"

namespace nbctc {
int launch_pipe_ns4(const Problem& p, const PipeCfg& cfg, cudaStream_t stream) { return pipe::launch_pipe_ns<4>(p, cfg, stream); }
}  // namespace nbctc
