// biped-class limb kernels, fast variant with implicit joint damping: explicit controls, no outputs; env step
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(6, 4, 86, true, b, 1, s1)
ABR_DEFINE_LIMB_ENV(6, 4, 86, true, b, 1, s1)
}
