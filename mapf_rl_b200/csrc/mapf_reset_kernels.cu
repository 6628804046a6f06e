// mapf_reset_kernels.cu — device-side instance generation: Environment.reset / __init__ without the
// host (environment.py:100-138 == :156-192, map_partition :21-70), sm_100a.
//
// One warp per environment (mapf_reset_device.cuh), followed by a masked BFS launch.
#include "mapf_reset_device.cuh"

namespace {

template <int RW, int RPL>
__global__ void __launch_bounds__(128)
reset_kernel(EnvDims d, const uint8_t *__restrict__ mask, uint64_t seed, uint64_t env_offset, float density,
             uint32_t *__restrict__ obst, uint8_t *__restrict__ pos, uint8_t *__restrict__ goal,
             int32_t *__restrict__ steps, int32_t *__restrict__ err)
{
    const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (e >= d.B) return;
    if (mask && !mask[e]) return;
    reset_env_warp<RW, RPL>(d, e, seed, env_offset + (uint64_t)e, density, obst, pos, goal, steps, err);
}

template <int RW>
int launch_reset_rw(mapf_env *env, const uint8_t *mask, uint64_t seed, uint64_t off, float density, cudaStream_t st)
{
    const EnvDims &d = env->d;
    const int rpl = (d.L + 31) / 32;
    const int warps = 4;
    const int grid = (d.B + warps - 1) / warps;
#define MAPF_RESET_LAUNCH(RPL)                                                                                      \
    reset_kernel<RW, RPL><<<grid, warps * 32, 0, st>>>(d, mask, seed, off, density, env->obst, env->pos, env->goal, \
                                                        env->steps, env->err)
    switch (rpl) {
        case 1: MAPF_RESET_LAUNCH(1); break;
        case 2: MAPF_RESET_LAUNCH(2); break;
        case 3: MAPF_RESET_LAUNCH(3); break;
        case 4: MAPF_RESET_LAUNCH(4); break;
        default: mapf_set_error("unsupported map size"); return MAPF_EINVAL;
    }
#undef MAPF_RESET_LAUNCH
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

}  // namespace

int mapf_launch_bfs_masked(mapf_env *env, const uint8_t *d_mask, cudaStream_t st);

int mapf_launch_reset(mapf_env *env, const uint8_t *d_mask, uint64_t seed, uint64_t env_offset, float density,
                      cudaStream_t st)
{
    int rc;
    switch (env->d.RW) {
        case 1: rc = launch_reset_rw<1>(env, d_mask, seed, env_offset, density, st); break;
        case 2: rc = launch_reset_rw<2>(env, d_mask, seed, env_offset, density, st); break;
        case 3: rc = launch_reset_rw<3>(env, d_mask, seed, env_offset, density, st); break;
        case 4: rc = launch_reset_rw<4>(env, d_mask, seed, env_offset, density, st); break;
        default: mapf_set_error("unsupported map size"); return MAPF_EINVAL;
    }
    if (rc != MAPF_OK) return rc;
    return mapf_launch_bfs_masked(env, d_mask, st);  // get_navi_map, environment.py:195
}
