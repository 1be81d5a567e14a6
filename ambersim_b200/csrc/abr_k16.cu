// instantiation of the fused kernels for G = 16 lanes per world
#define ABR_MINB 4
#include "abr_kernels.cuh"
namespace abr {
ABR_DEFINE_LAUNCHERS(16)
}
