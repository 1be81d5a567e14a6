// Barkour-class limb kernels, fast variants with implicit joint damping (eulerdamp on, MuJoCo's default): explicit controls; env step
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 2, false, f2, 1, s1)
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 2, false, f2, 5, s5)
ABR_DEFINE_LIMB_ENV(3, 1, 2, false, f2, 1, s1)
}
