// biped-class limb kernels, fast variant: sampler mode
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(6, 4, 86, true, b, 8, s8)
}
