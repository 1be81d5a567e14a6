// fixed-base chain ("hand") kernels, chains of up to 3 joints (Barrett hand class)
#include "abr_hand.cuh"
namespace abr {
int launch_hand_rollout_3(const Layout& L, const RolloutArgs& a, cudaStream_t st) { return hand::launch_hand_rollout_t<3>(L, a, st); }
int launch_hand_env_3(const Layout& L, const EnvArgs& a, cudaStream_t st) { return hand::launch_hand_env_t<3>(L, a, st); }
}
