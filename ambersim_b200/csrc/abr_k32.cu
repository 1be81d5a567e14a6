// instantiation of the fused kernels for G = 32 lanes per world
#include "abr_kernels.cuh"
namespace abr {
ABR_DEFINE_LAUNCHERS(32)
}
