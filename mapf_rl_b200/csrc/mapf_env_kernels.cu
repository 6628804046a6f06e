// mapf_env_kernels.cu — the environment hot path as hand-written CUDA for sm_100a.
//
//   (K1+K2, the fused step + observe kernel, lives in mapf_step_kernels.cu)
//   K3     bfs_navi_kernel       Environment.get_navi_map (environment.py:217-276): bit-parallel
//                                wavefront BFS, one warp per (env, agent), lane = map row.
//   pack / unpack helpers        Environment.load (environment.py:198-215) and attribute reads.
//
// None of this is a dense contraction: no tensor cores.  K1+K2 is bound by the HBM write of the
// observation bytes (486 B per agent-step); everything else stays in shared memory / registers.
#include <type_traits>

#include "mapf_bfs_device.cuh"
#include "mapf_common.cuh"

namespace {

constexpr int kBfsWarps = 4;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---------------------------------------------------------------------------------------------
// pack: u8 maps -> padded obstacle bitmaps; copy coordinates; steps = 0      (environment.py:198-215)
// ---------------------------------------------------------------------------------------------
__global__ void pack_load_kernel(EnvDims d, const int32_t *__restrict__ env_ids, int n, const uint8_t *__restrict__ maps,
                                 const uint8_t *__restrict__ agents, const uint8_t *__restrict__ goals,
                                 uint32_t *__restrict__ obst, uint8_t *__restrict__ pos, uint8_t *__restrict__ goal,
                                 int32_t *__restrict__ steps, int32_t *__restrict__ err)
{
    int i = blockIdx.x;
    if (i >= n) return;
    int e = env_ids ? env_ids[i] : i;
    if (e < 0 || e >= d.B) {  // a slot id outside the batch: nothing is written (the reference would raise IndexError)
        if (threadIdx.x == 0) atomicOr(err, MAPF_ERRBIT_STATE);
        return;
    }
    const uint8_t *m = maps + (size_t)i * d.L * d.L;
    uint32_t *o = obst + (size_t)e * d.obst_stride;
    for (int w = threadIdx.x; w < d.obst_stride; w += blockDim.x) {
        int row = w / d.RWS, ww = w - row * d.RWS;
        uint32_t bits = 0;
        int x = row - 4;
        if (row < d.R && ww < d.RW && x >= 0 && x < d.L) {
            for (int b = 0; b < 32; ++b) {
                int y = ww * 32 + b - 4;
                if (y >= 0 && y < d.L && m[x * d.L + y] != 0) bits |= 1u << b;
            }
        }
        o[w] = bits;
    }
    for (int k = threadIdx.x; k < 2 * d.N; k += blockDim.x) {
        pos[(size_t)e * 2 * d.N + k] = agents[(size_t)i * 2 * d.N + k];
        goal[(size_t)e * 2 * d.N + k] = goals[(size_t)i * 2 * d.N + k];
    }
    if (threadIdx.x == 0) steps[e] = 0;
}

// ---------------------------------------------------------------------------------------------
// state validation after load / set_state: the step kernel indexes shared memory with the coordinates, so a coordinate
// outside the map must never reach it.  Out-of-range coordinates are clamped into the map and latched as
// MAPF_ERRBIT_STATE (the reference raises IndexError on the same input); two agents on one cell are latched as
// MAPF_ERRBIT_UNIQUE (the reference raises RuntimeError('unique') at the next step, environment.py:424-428).
// One warp per environment.
// ---------------------------------------------------------------------------------------------
__global__ void validate_state_kernel(EnvDims d, const int32_t *__restrict__ env_ids, int n, uint8_t *__restrict__ pos,
                                      uint8_t *__restrict__ goal, int32_t *__restrict__ err)
{
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    const int e = env_ids ? env_ids[i] : i;
    if (e < 0 || e >= d.B) return;  // latched by pack_load_kernel
    uchar2 *pp = reinterpret_cast<uchar2 *>(pos) + (size_t)e * d.N;
    uchar2 *gg = reinterpret_cast<uchar2 *>(goal) + (size_t)e * d.N;
    bool bad = false;
    for (int a = lane; a < d.N; a += 32) {
        uchar2 p = pp[a], g = gg[a];
        if (p.x >= d.L || p.y >= d.L) {
            bad = true;
            p.x = min((int)p.x, d.L - 1), p.y = min((int)p.y, d.L - 1);
            pp[a] = p;
        }
        if (g.x >= d.L || g.y >= d.L) {
            bad = true;
            g.x = min((int)g.x, d.L - 1), g.y = min((int)g.y, d.L - 1);
            gg[a] = g;
        }
    }
    __syncwarp();
    bool dup = false;
    for (int a = lane; a < d.N; a += 32) {
        const uchar2 p = pp[a];
        for (int b = 0; b < a; ++b) {
            const uchar2 q = pp[b];
            if (q.x == p.x && q.y == p.y) dup = true;
        }
    }
    bad = __any_sync(MAPF_FULL_MASK, bad);
    dup = __any_sync(MAPF_FULL_MASK, dup);
    if (lane == 0 && (bad || dup)) atomicOr(err, (bad ? MAPF_ERRBIT_STATE : 0) | (dup ? MAPF_ERRBIT_UNIQUE : 0));
}

// ---------------------------------------------------------------------------------------------
// K3: bit-parallel wavefront BFS (environment.py:217-276): one warp per APW agents, see mapf_bfs_device.cuh
// ---------------------------------------------------------------------------------------------
template <int RW, int RPL, int APW>
__global__ void __launch_bounds__(kBfsWarps * 32, RW * RPL <= 9 ? 8 : 1)  // <= 64 registers (32 resident warps per SM)
                                                                          // wherever that does not spill
bfs_navi_kernel(EnvDims d, const int32_t *__restrict__ env_ids, const uint8_t *__restrict__ env_mask, int n,
                const uint32_t *__restrict__ obst, const uint8_t *__restrict__ goal, uint32_t *__restrict__ navi,
                uint32_t *__restrict__ navi_alt, const uint8_t *__restrict__ navi_sel, int32_t *__restrict__ dist_out)
{
    constexpr int LW = 32 / APW;
    const int g = (blockIdx.x * kBfsWarps + (threadIdx.x >> 5)) * APW + lane_id() / LW;
    bool alive = g < n * d.N;
    const int i = alive ? g / d.N : 0, a = alive ? g - i * d.N : 0;
    int e = env_ids ? env_ids[i] : i;
    if (e < 0 || e >= d.B) alive = false, e = 0;           // bad slot id: latched by pack_load_kernel
    if (alive && env_mask && !env_mask[e]) alive = false;  // masked re-computation after a device-side reset
    if (!__any_sync(MAPF_FULL_MASK, alive)) return;
    // the slot's live buffer (mapf_common.cuh)
    uint32_t *nv = (navi_alt && navi_sel[e]) ? navi_alt : navi;
    bfs_navi_warp<RW, RPL, APW>(d, e, a, i, alive, obst, goal, nv, dist_out);
}

// The heuristic maps of the pre-generated NEXT instances of the listed slots (episode handling of mapf_env_rollout): reads
// the staged obstacle bitmaps / goals, writes the buffer the slot's live instance does not use.  The list's length is only
// known on the device, so a fixed grid claims (slot, agent group) items from a counter.
template <int RW, int RPL, int APW>
__global__ void __launch_bounds__(kBfsWarps * 32, RW * RPL <= 9 ? 8 : 4)
pregen_bfs_kernel(EnvDims d, const uint32_t *__restrict__ list, int min_count, unsigned long long *__restrict__ counter,
                  const uint32_t *__restrict__ pg_obst, const uint8_t *__restrict__ pg_goal, uint32_t *__restrict__ navi,
                  uint32_t *__restrict__ navi_alt, const uint8_t *__restrict__ navi_sel, const uint32_t *__restrict__ pg_n,
                  uint32_t *__restrict__ pg_cnt, uint32_t *__restrict__ pg_epi)
{
    constexpr int LW = 32 / APW;
    const int lane = lane_id();
    const unsigned groups = (unsigned)(d.N + APW - 1) / APW;
    if (list[0] < (unsigned)min_count) return;
    const unsigned long long items = (unsigned long long)list[0] * groups;
    for (;;) {
        unsigned long long it = 0;
        if (lane == 0) it = atomicAdd(counter, 1ull);
        it = __shfl_sync(MAPF_FULL_MASK, it, 0);
        if (it >= items) break;
        const unsigned i = (unsigned)(it / groups);
        const int a = (int)(it - (unsigned long long)i * groups) * APW + lane / LW;
        const int e = (int)list[1 + i];
        const bool alive = a < d.N;
        // (the slot's selector only changes when THIS instance is adopted, i.e. after the last group has published it)
        uint32_t *nv = __ldcg(navi_sel + e) ? navi : navi_alt;
        bfs_navi_warp<RW, RPL, APW>(d, e, alive ? a : 0, 0, alive, pg_obst, pg_goal, nv, nullptr);
        __syncwarp();
        if (lane == 0) {
            __threadfence();  // the tiles (and, through the generator's launch, the staged state) before the count
            if (atomicAdd(pg_cnt + e, 1u) + 1u == groups) {
                __threadfence();
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(pg_epi + e), "r"(__ldcg(pg_n + e)) : "memory");
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// communication mask of the actor-side inference glue                    (model.py:196-208)
//
// comm_mask[i][j] = (|dx| <= obs_radius and |dy| <= obs_radius) and j is among the k nearest agents of i
// (Euclidean, i itself included at distance 0).  torch.topk leaves the order of equal distances
// unspecified; here ties go to the lower agent id (key = d^2 << 8 | j).  One warp per environment,
// lane = agent; positions of the env sit in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kCommWarps = 4;

__global__ void __launch_bounds__(kCommWarps * 32)
comm_mask_kernel(EnvDims d, const uint8_t *__restrict__ pos, int k_nearest, uint8_t *__restrict__ out)
{
    __shared__ uint16_t s_pos[kCommWarps][MAPF_MAX_AGENTS];
    __shared__ uint32_t s_bits[kCommWarps][MAPF_MAX_AGENTS][MAPF_MAX_AGENTS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * kCommWarps + warp;
    if (e >= d.B) return;
    const int N = d.N;
    for (int a = lane; a < N; a += 32) s_pos[warp][a] = reinterpret_cast<const uint16_t *>(pos)[(size_t)e * N + a];
    __syncwarp();
    for (int i = lane; i < N; i += 32) {
        const int xi = s_pos[warp][i] & 0xff, yi = s_pos[warp][i] >> 8;
        // three smallest keys, ascending.  Only agents inside the field of view can end up in the mask, and those lie within
        // d^2 <= 2 r^2; an agent farther away can neither be one nor displace one from the three nearest: start at that bound
        // (the insertion below then runs for the few close agents only)
        constexpr uint32_t kFar = (uint32_t)(2 * MAPF_OBS_RADIUS * MAPF_OBS_RADIUS + 1) << 8;
        uint32_t b0 = kFar, b1 = kFar, b2 = kFar;
        for (int j = 0; j < N; ++j) {
            const int dx = xi - (s_pos[warp][j] & 0xff), dy = yi - (s_pos[warp][j] >> 8);
            uint32_t key = ((uint32_t)(dx * dx + dy * dy) << 8) | (uint32_t)j;
            if (key < b2) {
                b2 = key;
                if (b2 < b1) { const uint32_t t = b1; b1 = b2; b2 = t; }
                if (b1 < b0) { const uint32_t t = b0; b0 = b1; b1 = t; }
            }
        }
        uint32_t row[MAPF_MAX_AGENTS / 32] = {0, 0, 0, 0};
        const uint32_t best[3] = {b0, b1, b2};
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            if (q < k_nearest && best[q] < kFar) {
                const int j = best[q] & 0xff;
                const int dx = abs(xi - (s_pos[warp][j] & 0xff)), dy = abs(yi - (s_pos[warp][j] >> 8));
                if (dx <= MAPF_OBS_RADIUS && dy <= MAPF_OBS_RADIUS) {
#pragma unroll
                    for (int w = 0; w < MAPF_MAX_AGENTS / 32; ++w)
                        if ((j >> 5) == w) row[w] |= 1u << (j & 31);
                }
            }
        }
#pragma unroll
        for (int w = 0; w < MAPF_MAX_AGENTS / 32; ++w) s_bits[warp][i][w] = row[w];
    }
    __syncwarp();
    uint8_t *o = out + (size_t)e * N * N;
    if ((N & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        // 16 mask bytes (16 consecutive j of one row) per lane and store: bits -> bool bytes as in the observation path
        const int per_row = N >> 4;
        for (int c = lane; c < N * per_row; c += 32) {
            const int i = c / per_row, j0 = (c - i * per_row) << 4;
            const uint32_t b = (s_bits[warp][i][j0 >> 5] >> (j0 & 31)) & 0xffffu;
            uint4 v;
            v.x = ((b & 0xfu) * 0x00204081u) & 0x01010101u;
            v.y = (((b >> 4) & 0xfu) * 0x00204081u) & 0x01010101u;
            v.z = (((b >> 8) & 0xfu) * 0x00204081u) & 0x01010101u;
            v.w = ((b >> 12) * 0x00204081u) & 0x01010101u;
            reinterpret_cast<uint4 *>(o)[c] = v;
        }
    } else {
        for (int idx = lane; idx < N * N; idx += 32) {
            const int i = idx / N, j = idx - i * N;
            o[idx] = (s_bits[warp][i][j >> 5] >> (j & 31)) & 1u;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// attribute reads: unpack state for parity dumps / drop-in attributes
// ---------------------------------------------------------------------------------------------
__global__ void unpack_map_kernel(EnvDims d, const uint32_t *__restrict__ obst, uint8_t *__restrict__ map_out)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)d.B * d.L * d.L;
    if (idx >= total) return;
    int y = idx % d.L;
    int x = (idx / d.L) % d.L;
    int e = idx / ((size_t)d.L * d.L);
    uint32_t w = obst[(size_t)e * d.obst_stride + (x + 4) * d.RWS + ((y + 4) >> 5)];
    map_out[idx] = (w >> ((y + 4) & 31)) & 1u;
}

__global__ void unpack_navi_kernel(EnvDims d, const uint32_t *__restrict__ navi, const uint32_t *__restrict__ navi_alt,
                                   const uint8_t *__restrict__ navi_sel, uint8_t *__restrict__ navi_out)
{
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)d.B * d.N * 4 * d.L * d.L;
    if (idx >= total) return;
    int y = idx % d.L;
    int x = (idx / d.L) % d.L;
    int k = (idx / ((size_t)d.L * d.L)) % 4;
    size_t ea = idx / ((size_t)4 * d.L * d.L);
    const int pr = x + 4, pc = y + 4;
    const int bx = min(pr >> 3, d.NB - 1), by = min(pc >> 3, d.NB - 1);
    const uint32_t *live = (navi_alt && navi_sel[ea / d.N]) ? navi_alt : navi;
    const uint2 v = reinterpret_cast<const uint2 *>(live + ea * d.navi_agent_stride)[((size_t)(bx * d.NB + by) << 4) + (pr - 8 * bx)];
    const uint32_t half = k < 2 ? v.x : v.y;
    navi_out[idx] = (half >> (16 * (k & 1) + (pc - 8 * by))) & 1u;
}

}  // namespace

// ---- entry points used by mapf_abi.cu -----------------------------------------------------------
int mapf_launch_pack_load(mapf_env *env, const int32_t *d_env_ids, int n, const uint8_t *d_maps,
                          const uint8_t *d_agents, const uint8_t *d_goals, cudaStream_t st)
{
    pack_load_kernel<<<n, 128, 0, st>>>(env->d, d_env_ids, n, d_maps, d_agents, d_goals, env->obst, env->pos, env->goal,
                                        env->steps, env->err);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

// slots env_ids[0..n) (NULL: slots 0..n-1)
int mapf_launch_validate_state(mapf_env *env, const int32_t *d_env_ids, int n, cudaStream_t st)
{
    if (n <= 0) return MAPF_OK;
    validate_state_kernel<<<(n + 3) / 4, 128, 0, st>>>(env->d, d_env_ids, n, env->pos, env->goal, env->err);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

// pregen_min >= 0: pregen_bfs_kernel over the list in env->ro_prio (ids / mask / n / dist unused)
template <int RW>
static int launch_bfs_rw(mapf_env *env, const int32_t *ids, const uint8_t *mask, int n, int32_t *dist, int pregen_min, cudaStream_t st)
{
    const bool pregen = pregen_min >= 0;
    const EnvDims &d = env->d;
    // four agents per warp (8 lanes x up to 5 rows each) for two-word maps of up to 40 cells a side, two (16 lanes x up to 6 rows)
    // for every other map of up to 88 cells, else one
    const int apw = d.bfs_apw;
    const int rpl = (d.L + 32 / apw - 1) / (32 / apw);
    const long warps = ((long)n * d.N + apw - 1) / apw;
    const int grid = pregen ? env->num_sms * 8 : (int)((warps + kBfsWarps - 1) / kBfsWarps);
#define MAPF_BFS_LAUNCH(RPL, APW)                                                                                                   \
    do {                                                                                                                            \
        if (pregen)                                                                                                                 \
            pregen_bfs_kernel<RW, RPL, APW><<<grid, kBfsWarps * 32, 0, st>>>(d, env->ro_prio, pregen_min, env->ro_work + 3, env->pg_obst, \
                                                                             env->pg_goal, env->navi, env->navi_alt, env->navi_sel, \
                                                                             env->pg_n, env->pg_cnt, env->pg_epi);                   \
        else                                                                                                                        \
            bfs_navi_kernel<RW, RPL, APW><<<grid, kBfsWarps * 32, 0, st>>>(d, ids, mask, n, env->obst, env->goal, env->navi,         \
                                                                           env->navi_alt, env->navi_sel, dist);                      \
    } while (0)
    bool ok = true;
    if constexpr (RW == 1) {
        switch (rpl) {
            case 1: MAPF_BFS_LAUNCH(1, 2); break;
            case 2: MAPF_BFS_LAUNCH(2, 2); break;
            default: ok = false;
        }
    } else if constexpr (RW == 2) {
        if (apw == 4) {
            if (rpl <= 4) MAPF_BFS_LAUNCH(4, 4);
            else MAPF_BFS_LAUNCH(5, 4);
        } else {
            switch (rpl) {
                case 2: MAPF_BFS_LAUNCH(2, 2); break;
                case 3: MAPF_BFS_LAUNCH(3, 2); break;
                case 4: MAPF_BFS_LAUNCH(4, 2); break;
                default: ok = false;
            }
        }
    } else if constexpr (RW == 3) {
        switch (rpl) {
            case 4: MAPF_BFS_LAUNCH(4, 2); break;
            case 5: MAPF_BFS_LAUNCH(5, 2); break;
            case 6: MAPF_BFS_LAUNCH(6, 2); break;
            default: ok = false;
        }
    } else {
        switch (rpl) {
            case 3: MAPF_BFS_LAUNCH(3, 1); break;
            case 4: MAPF_BFS_LAUNCH(4, 1); break;
            default: ok = false;
        }
    }
#undef MAPF_BFS_LAUNCH
    if (!ok) {
        mapf_set_error("unsupported map size");
        return MAPF_EINVAL;
    }
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

static int launch_bfs(mapf_env *env, const int32_t *d_env_ids, const uint8_t *d_mask, int n, int32_t *d_dist_out, int pregen_min,
                      cudaStream_t st)
{
    switch (env->d.RW) {
        case 1: return launch_bfs_rw<1>(env, d_env_ids, d_mask, n, d_dist_out, pregen_min, st);
        case 2: return launch_bfs_rw<2>(env, d_env_ids, d_mask, n, d_dist_out, pregen_min, st);
        case 3: return launch_bfs_rw<3>(env, d_env_ids, d_mask, n, d_dist_out, pregen_min, st);
        case 4: return launch_bfs_rw<4>(env, d_env_ids, d_mask, n, d_dist_out, pregen_min, st);
    }
    mapf_set_error("unsupported map size");
    return MAPF_EINVAL;
}

int mapf_launch_bfs(mapf_env *env, const int32_t *d_env_ids, int n, int32_t *d_dist_out, cudaStream_t st)
{
    return launch_bfs(env, d_env_ids, nullptr, n, d_dist_out, -1, st);
}

// all B slots, skipping those whose mask byte is zero (NULL mask = all)
int mapf_launch_bfs_masked(mapf_env *env, const uint8_t *d_mask, cudaStream_t st)
{
    return launch_bfs(env, nullptr, d_mask, env->d.B, nullptr, -1, st);
}

// heuristic maps of the staged next instances listed in env->ro_prio (mapf_launch_pregen, mapf_reset_kernels.cu)
int mapf_launch_pregen_bfs(mapf_env *env, int min_count, cudaStream_t st)
{
    return launch_bfs(env, nullptr, nullptr, 0, nullptr, min_count < 0 ? 0 : min_count, st);
}

int mapf_launch_comm_mask(mapf_env *env, int k_nearest, uint8_t *d_out, cudaStream_t st)
{
    const int grid = (env->d.B + kCommWarps - 1) / kCommWarps;
    comm_mask_kernel<<<grid, kCommWarps * 32, 0, st>>>(env->d, env->pos, k_nearest, d_out);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

int mapf_launch_unpack(mapf_env *env, uint8_t *d_map, uint8_t *d_navi, cudaStream_t st)
{
    const EnvDims &d = env->d;
    if (d_map) {
        size_t total = (size_t)d.B * d.L * d.L;
        unpack_map_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d, env->obst, d_map);
        MAPF_CUDA(cudaGetLastError());
    }
    if (d_navi) {
        size_t total = (size_t)d.B * d.N * 4 * d.L * d.L;
        unpack_navi_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d, env->navi, env->navi_alt, env->navi_sel, d_navi);
        MAPF_CUDA(cudaGetLastError());
    }
    return MAPF_OK;
}
