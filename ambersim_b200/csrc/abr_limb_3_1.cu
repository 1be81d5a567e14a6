// limb-path kernels for chains of up to 3 joints and 1 contact per path (Barkour-class quadrupeds):
// the flat 4-lane sharing pattern resolved at compile time
#include "abr_limb.cuh"
namespace abr {
ABR_DEFINE_LIMB_LAUNCHERS(3, 1, 2, false, f2)
}
