// Explicit instantiation unit of the pipeline kernel for NS = 16 (Lmax <= 256); see pipe_kernel.cuh.
#include "pipe_kernel.cuh"

namespace nbctc {
int launch_pipe_ns16(const Problem& p, const PipeCfg& cfg, cudaStream_t stream) { return pipe::launch_pipe_ns<16>(p, cfg, stream); }
}  // namespace nbctc
