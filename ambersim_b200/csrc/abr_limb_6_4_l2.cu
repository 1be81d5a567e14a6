// limb-path kernels for the flat 2-lane pattern with chains of up to 6 joints and 4 contacts per path in the contact-body form
// (legs only: exoskeleton / biped without arms): general variants; the variants that are not built separately run these
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(6, 4, 1, true, l2, -1, sg)
ABR_DEFINE_LIMB_ENV(6, 4, 1, true, l2, -1, sg)
ABR_ALIAS_LIMB_ROLLOUT(6, 4, l2, s4)
ABR_ALIAS_LIMB_ROLLOUT(6, 4, l2, s1)
ABR_ALIAS_LIMB_ROLLOUT(6, 4, l2, s5)
ABR_ALIAS_LIMB_ROLLOUT(6, 4, l2, s9)
ABR_ALIAS_LIMB_ROLLOUT(6, 4, l2, s13)
ABR_ALIAS_LIMB_ENV(6, 4, l2, s1)
}
