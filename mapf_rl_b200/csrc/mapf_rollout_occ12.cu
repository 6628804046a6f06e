// mapf_rollout_occ12.cu — the persistent rollout kernel (mapf_rollout_device.cuh) compiled for 12 resident CTAs = 24 warps per SM
// (85 registers per thread).
#include "mapf_rollout_device.cuh"

int mapf_launch_rollout_occ12(mapf_env *env, const StepParams &p, const void *args, const void *tuning, cudaStream_t st)
{
    return launch_rollout_class<12>(env, p, *static_cast<const RolloutArgs *>(args), *static_cast<const RolloutTuning *>(tuning), st);
}
