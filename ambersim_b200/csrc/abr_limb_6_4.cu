// limb-path kernels for chains of up to 6 joints and 4 contacts per path (biped / exoskeleton class),
// any sharing pattern
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_LAUNCHERS(6, 4, -1, false, g)
}
