// flat 8-lane pattern, fast variants: sampler mode
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 3, false, f3, 8, s8)
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 3, false, f3, 12, s12)
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 3, false, f3, 9, s9)
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 3, false, f3, 13, s13)
}
