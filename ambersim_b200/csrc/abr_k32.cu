// instantiation of the fused kernels for G = 32 lanes per world
#define ABR_MINB 4
#include "abr_kernels.cuh"
namespace abr {
ABR_DEFINE_LAUNCHERS(32)
}
