// biped-class limb kernels, fast variant: sampler mode with kept trajectories
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(6, 4, 86, true, b, 12, s12)
}
