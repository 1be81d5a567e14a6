// biped-class limb kernels, fast variant with implicit joint damping: sampler mode with kept trajectories
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(6, 4, 86, true, b, 13, s13)
}
