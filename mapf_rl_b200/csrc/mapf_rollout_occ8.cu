// mapf_rollout_occ8.cu — the persistent rollout kernel (mapf_rollout_device.cuh) compiled for 8 resident CTAs = 16 warps per SM
// (128 registers per thread).
#include "mapf_rollout_device.cuh"

int mapf_launch_rollout_occ8(mapf_env *env, const StepParams &p, const void *args, const void *tuning, cudaStream_t st)
{
    return launch_rollout_class<8>(env, p, *static_cast<const RolloutArgs *>(args), *static_cast<const RolloutTuning *>(tuning), st);
}
