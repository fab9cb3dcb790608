// Explicit instantiation unit of the pipeline kernel for NS = 2 (Lmax <= 32); see pipe_kernel.cuh.
#include "pipe_kernel.cuh"

namespace nbctc {
int launch_pipe_ns2(const Problem& p, const PipeCfg& cfg, cudaStream_t stream) { return pipe::launch_pipe_ns<2>(p, cfg, stream); }
}  // namespace nbctc
