// biped-class limb kernels, fast variant with implicit joint damping: sampler mode
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(6, 4, 86, true, b, 9, s9)
}
