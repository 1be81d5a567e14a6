// legs-only class, fast variant: sampler mode
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(6, 4, 1, true, l2, 8, s8)
}
