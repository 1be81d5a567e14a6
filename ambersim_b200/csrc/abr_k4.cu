// instantiation of the fused kernels for G = 4 lanes per world
#define ABR_MINB 1
#include "abr_kernels.cuh"
namespace abr {
ABR_DEFINE_LAUNCHERS(4)
}
