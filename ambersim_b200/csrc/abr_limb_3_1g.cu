// limb-path kernels for chains of up to 3 joints and 1 contact per path, any sharing pattern
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(3, 1, -1, false, g, -1, sg)
ABR_DEFINE_LIMB_ENV(3, 1, -1, false, g, -1, sg)
}
