// Barkour-class limb kernels, fast variants (eulerdamp off, one Newton iteration): sampler mode, with and without kept trajectories
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 2, false, f2, 8, s8)
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 2, false, f2, 12, s12)
}
