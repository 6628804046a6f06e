// mapf_rollout_occ8.cu — the persistent rollout kernel (mapf_rollout_device.cuh) compiled for 8 resident CTAs = 16 warps per SM
// (128 registers per thread).
#include "mapf_rollout_device.cuh"

int mapf_launch_rollout_occ8(mapf_env *env, const StepParams &p, const void *args, const void *tuning, cudaStream_t st)
{
    return launch_rollout_class<8>(env, p, *static_cast<const RolloutArgs *>(args), *static_cast<const RolloutTuning *>(tuning), st);
}

#ifdef MAPF_ENABLE_DIAG
// diagnosis build only: search-phase cycles of environment 0 inside this translation unit's rollout kernels; reset = 1 zeroes them
extern "C" int mapf_diag_bfs_cycles_occ8(unsigned long long *out8, int reset)
{
    if (reset) {
        unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        return cudaMemcpyToSymbol(g_bfs_cycles, z, sizeof(z)) == cudaSuccess ? 0 : -1;
    }
    return cudaMemcpyFromSymbol(out8, g_bfs_cycles, sizeof(unsigned long long) * 8) == cudaSuccess ? 0 : -1;
}
#endif
