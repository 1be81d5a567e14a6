// biped-class limb kernels, fast variant: explicit controls, trajectories written out
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(6, 4, 86, true, b, 4, s4)
}
