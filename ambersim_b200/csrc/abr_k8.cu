// instantiation of the fused kernels for G = 8 lanes per world
#define ABR_MINB 2
#include "abr_kernels.cuh"
namespace abr {
ABR_DEFINE_LAUNCHERS(8)
}
