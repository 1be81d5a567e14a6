// limb-path kernels for chains of up to 3 joints and 1 contact per path (Barkour-class quadrupeds), flat 4-lane sharing
// pattern resolved at compile time: the general variants (every option decided at run time)
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(3, 1, 2, false, f2, -1, sg)
ABR_DEFINE_LIMB_ENV(3, 1, 2, false, f2, -1, sg)
}
