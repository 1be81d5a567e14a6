// biped-class limb kernels, fast variant with implicit joint damping: explicit controls, trajectories written out
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(6, 4, 86, true, b, 5, s5)
}
