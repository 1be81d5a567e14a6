// limb-path kernels for the biped / exoskeleton class: chains of up to 6 joints, 4 contacts per path, sharing pattern
// "trunk by 4 lanes, three torso positions by 2" resolved at compile time, contact-body form: the general variants
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(6, 4, 86, true, b, -1, sg)
ABR_DEFINE_LIMB_ENV(6, 4, 86, true, b, -1, sg)
}
