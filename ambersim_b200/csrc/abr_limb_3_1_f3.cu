// limb-path kernels for the flat 8-lane pattern (hexapods: up to 8 limbs of up to 3 joints on one trunk): general variants
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 3, false, f3, -1, sg)
ABR_DEFINE_LIMB_ENV(3, 1, 3, false, f3, -1, sg)
}
