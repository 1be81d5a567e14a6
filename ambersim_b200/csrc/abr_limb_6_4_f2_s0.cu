// flat 4-lane long-chain class, fast variant: explicit controls, no outputs; env step
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(6, 4, 2, true, f2, 0, s0)
ABR_DEFINE_LIMB_ENV(6, 4, 2, true, f2, 0, s0)
}
