// explicit instantiation unit of the wide-row sequence-per-warp kernel: rows of up to 1024 classes
#include "seqwide_kernel.cuh"
namespace nbctc {
template int launch_seqwide_nv<8>(const WideParams&, int, int, cudaStream_t);
}
