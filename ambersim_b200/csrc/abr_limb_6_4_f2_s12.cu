// flat 4-lane long-chain class, fast variant: sampler mode with kept trajectories
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(6, 4, 2, true, f2, 12, s12)
}
