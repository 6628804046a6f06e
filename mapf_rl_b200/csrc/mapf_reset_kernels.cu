// mapf_reset_kernels.cu — device-side instance generation (environment.py:146-196).  Placeholder until
// the generator lands: fails loudly instead of silently doing nothing.
#include "mapf_common.cuh"

int mapf_launch_reset(mapf_env *, const uint8_t *, uint64_t, uint64_t, float, cudaStream_t)
{
    mapf_set_error("mapf_env_reset: device-side generator not built in this version");
    return MAPF_EINVAL;
}
