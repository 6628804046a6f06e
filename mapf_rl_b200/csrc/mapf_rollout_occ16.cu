// mapf_rollout_occ16.cu — the persistent rollout kernel (mapf_rollout_device.cuh) compiled for 16 resident CTAs = 32 warps per SM
// (64 registers per thread).
#include "mapf_rollout_device.cuh"

int mapf_launch_rollout_occ16(mapf_env *env, const StepParams &p, const void *args, const void *tuning, cudaStream_t st)
{
    return launch_rollout_class<16>(env, p, *static_cast<const RolloutArgs *>(args), *static_cast<const RolloutTuning *>(tuning), st);
}
