// Barkour-class limb kernels, fast variants with implicit joint damping: sampler mode
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 2, false, f2, 9, s9)
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 2, false, f2, 13, s13)
}
