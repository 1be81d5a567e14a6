// flat 4-lane long-chain class, fast variant: sampler mode
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(6, 4, 2, true, f2, 8, s8)
}
