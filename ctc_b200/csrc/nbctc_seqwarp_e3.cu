// explicit instantiation unit of the sequence-per-warp kernel: rows of 65..96 classes
#include "seqwarp_kernel.cuh"
namespace nbctc {
template int launch_seqwarp_epl<3>(const SwParams&, int, int, cudaStream_t);
template int seqwarp_occupancy_epl<3>(int);
}
