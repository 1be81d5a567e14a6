// instantiation of the fused kernels for G = 4 lanes per world
#include "abr_kernels.cuh"
namespace abr {
ABR_DEFINE_LAUNCHERS(4)
}
