// Barkour-class limb kernels, fast variants (eulerdamp off, one Newton iteration): explicit controls, with and without outputs; env step
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 2, false, f2, 0, s0)
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 2, false, f2, 4, s4)
ABR_DEFINE_LIMB_ENV(3, 1, 2, false, f2, 0, s0)
}
