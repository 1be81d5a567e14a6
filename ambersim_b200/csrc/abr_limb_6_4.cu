// limb-path kernels for chains of up to 6 joints and 4 contacts per path, any sharing pattern
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_ROLLOUT(6, 4, -1, false, g, -1, sg)
ABR_DEFINE_LIMB_ENV(6, 4, -1, false, g, -1, sg)
}
