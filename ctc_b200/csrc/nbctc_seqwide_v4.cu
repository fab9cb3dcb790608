// explicit instantiation unit of the wide-row sequence-per-warp kernel: rows of up to 512 classes
#include "seqwide_kernel.cuh"
namespace nbctc {
template int launch_seqwide_nv<4>(const WideParams&, int, int, cudaStream_t);
}
