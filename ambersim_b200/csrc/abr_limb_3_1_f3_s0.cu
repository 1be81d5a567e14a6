// flat 8-lane pattern, fast variants: explicit controls; env step
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 3, false, f3, 0, s0)
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 3, false, f3, 4, s4)
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 3, false, f3, 1, s1)
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 3, false, f3, 5, s5)
ABR_DEFINE_LIMB_ENV(3, 1, 3, false, f3, 0, s0)
ABR_DEFINE_LIMB_ENV(3, 1, 3, false, f3, 1, s1)
}
