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
                                 int32_t *__restrict__ steps)
{
    int i = blockIdx.x;
    if (i >= n) return;
    int e = env_ids ? env_ids[i] : i;
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
// K3: bit-parallel wavefront BFS                                   (environment.py:217-276)
//
// Lane l owns the RPL consecutive map rows RPL*l .. RPL*l + RPL - 1 (blocked, so that only the first and
// the last of them need a neighbour lane: two shuffles per word and wave whatever RPL is), each row RW words
// of padded column bits.  One wave:
//   new = (frontier shifted to the four neighbours) & free-and-unvisited
// and the heuristic bit "neighbour in direction d is strictly closer" (environment.py:260-274) is
// exactly "this cell is new in wave t and that neighbour was in the frontier of wave t-1", so the
// four direction planes fall out of the same step and no distance array is needed.
// ---------------------------------------------------------------------------------------------
// APW agents share a warp (LW = 32 / APW lanes each): maps of up to 48 rows fit 16 lanes x 3 rows, so two agents' waves
// run in one instruction stream (at 40x40 one agent per warp leaves 12 of 32 lanes idle); the warp iterates until
// both are done.
template <int RW, int RPL, int APW>
__global__ void __launch_bounds__(kBfsWarps * 32, RW * RPL <= 9 ? 8 : 1)  // <= 64 registers (32 resident warps per SM)
                                                                          // wherever that does not spill
bfs_navi_kernel(EnvDims d, const int32_t *__restrict__ env_ids, const uint8_t *__restrict__ env_mask, int n,
                const uint32_t *__restrict__ obst, const uint8_t *__restrict__ goal, uint32_t *__restrict__ navi,
                int32_t *__restrict__ dist_out)
{
    constexpr int LW = 32 / APW;
    const int lane = lane_id() % LW;             // lane within the agent's group
    const int g = (blockIdx.x * kBfsWarps + (threadIdx.x >> 5)) * APW + lane_id() / LW;
    bool alive = g < n * d.N;
    const int i = alive ? g / d.N : 0, a = alive ? g - i * d.N : 0;
    const int e = env_ids ? env_ids[i] : i;
    if (alive && env_mask && !env_mask[e]) alive = false;  // masked re-computation after a device-side reset
    if (!__any_sync(MAPF_FULL_MASK, alive)) return;
    const uint32_t *ob = obst + (size_t)e * d.obst_stride;

    // unv = free and not yet visited, fro = frontier of the previous wave, pl = the four heuristic planes
    uint32_t unv[RPL][RW], fro[RPL][RW], pl[4][RPL][RW];
    const int gx = goal[((size_t)e * d.N + a) * 2], gy = goal[((size_t)e * d.N + a) * 2 + 1];
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
        const int row = alive ? lane * RPL + q : d.L;  // a group without an agent owns no rows
#pragma unroll
        for (int w = 0; w < RW; ++w) {
            // in-map column mask for this word: padded bits [4, L+4)
            int lo = max(4 - 32 * w, 0), hi = min(d.L + 4 - 32 * w, 32);
            uint32_t cm = 0;
            if (hi > lo) cm = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
            const uint32_t fre = (row < d.L) ? (~__ldg(ob + (row + 4) * d.RWS + w) & cm) : 0u;
            const int p = gy + 4;
            fro[q][w] = (row == gx && (p >> 5) == w) ? ((1u << (p & 31)) & fre) : 0u;
            unv[q][w] = fre & ~fro[q][w];
#pragma unroll
            for (int k = 0; k < 4; ++k) pl[k][q][w] = 0;
        }
    }

    int32_t *dist = (dist_out && alive) ? dist_out + ((size_t)i * d.N + a) * d.L * d.L : nullptr;
    if (dist_out) {
        for (int q = 0; q < RPL; ++q) {
            const int row = lane * RPL + q;
            if (dist && row < d.L)
                for (int y = 0; y < d.L; ++y) dist[row * d.L + y] = MAPF_DIST_UNREACHABLE;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < RPL; ++q)
#pragma unroll
            for (int w = 0; w < RW; ++w)
                if (dist && fro[q][w]) dist[gx * d.L + gy] = 0;
    }

    // Wavefront loop.  Adjacent reachable cells of a 4-connected grid differ by exactly one in distance, so "the
    // neighbour is strictly closer" (environment.py:260-274) is "its distance mod 3 is mine minus one": the loop only
    // records WHICH residue a cell's wave had (m1 / m2; residue 0 = reached and in neither) -- one LOP per word and wave
    // instead of four plane updates -- and the four planes are derived once at the end.
    uint32_t m1[RPL][RW], m2[RPL][RW];
#pragma unroll
    for (int q = 0; q < RPL; ++q)
#pragma unroll
        for (int w = 0; w < RW; ++w) m1[q][w] = m2[q][w] = 0;
    auto wave = [&](auto selc, const int t) -> bool {
        constexpr int sel = decltype(selc)::value;  // t mod 3
        uint32_t nw[RPL][RW];
        uint32_t any = 0;
#pragma unroll
        for (int w = 0; w < RW; ++w) {
            // rows of the neighbouring lanes that touch this lane's block
            uint32_t above = __shfl_up_sync(MAPF_FULL_MASK, fro[RPL - 1][w], 1, LW);
            uint32_t below = __shfl_down_sync(MAPF_FULL_MASK, fro[0][w], 1, LW);
            if (lane == 0) above = 0;
            if (lane == LW - 1) below = 0;
#pragma unroll
            for (int q = 0; q < RPL; ++q) {
                const uint32_t f = fro[q][w];
                const uint32_t fl = w > 0 ? __funnelshift_l(fro[q][w > 0 ? w - 1 : 0], f, 1) : f << 1;
                const uint32_t fr = w < RW - 1 ? __funnelshift_r(f, fro[q][w < RW - 1 ? w + 1 : w], 1) : f >> 1;
                const uint32_t up = q > 0 ? fro[q > 0 ? q - 1 : 0][w] : above;
                const uint32_t dn = q < RPL - 1 ? fro[q < RPL - 1 ? q + 1 : q][w] : below;
                const uint32_t x = (fl | fr | up | dn) & unv[q][w];
                nw[q][w] = x;
                any |= x;
            }
        }
        if (!__any_sync(MAPF_FULL_MASK, any != 0)) return false;
#pragma unroll
        for (int q = 0; q < RPL; ++q)
#pragma unroll
            for (int w = 0; w < RW; ++w) {
                unv[q][w] &= ~nw[q][w];
                fro[q][w] = nw[q][w];
                if constexpr (sel == 1) m1[q][w] |= nw[q][w];
                if constexpr (sel == 2) m2[q][w] |= nw[q][w];
                if (dist_out) {  // a kernel argument: the loop is compiled twice, the common one without any of this
                    uint32_t x = dist ? nw[q][w] : 0u;
                    while (x) {
                        int b = __ffs(x) - 1;
                        x &= x - 1;
                        dist[(lane * RPL + q) * d.L + (32 * w + b - 4)] = t;
                    }
                }
            }
        return true;
    };
    for (int t = 1;; t += 3) {
        if (!wave(std::integral_constant<int, 1>{}, t)) break;
        if (!wave(std::integral_constant<int, 2>{}, t + 1)) break;
        if (!wave(std::integral_constant<int, 0>{}, t + 2)) break;
    }

    // The four heuristic planes from the residues: with z = reached cells of residue 0, direction "neighbour n is
    // closer" holds at a cell c iff (c in m1, n in z) or (c in m2, n in m1) or (c in z, n in m2).
    {
        uint32_t z[RPL][RW];
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
            const int row = alive ? lane * RPL + q : d.L;
#pragma unroll
            for (int w = 0; w < RW; ++w) {
                int lo = max(4 - 32 * w, 0), hi = min(d.L + 4 - 32 * w, 32);
                uint32_t cm = 0;
                if (hi > lo) cm = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
                const uint32_t fre = (row < d.L) ? (~__ldg(ob + (row + 4) * d.RWS + w) & cm) : 0u;
                z[q][w] = fre & ~unv[q][w] & ~m1[q][w] & ~m2[q][w];
            }
        }
        auto closer = [](uint32_t c1, uint32_t c2, uint32_t cz, uint32_t n1, uint32_t n2, uint32_t nz) -> uint32_t {
            return (c1 & nz) | (c2 & n1) | (cz & n2);
        };
#pragma unroll
        for (int w = 0; w < RW; ++w) {
            // residue rows of the neighbouring lanes that touch this lane's block
            uint32_t a1 = __shfl_up_sync(MAPF_FULL_MASK, m1[RPL - 1][w], 1, LW), b1 = __shfl_down_sync(MAPF_FULL_MASK, m1[0][w], 1, LW);
            uint32_t a2 = __shfl_up_sync(MAPF_FULL_MASK, m2[RPL - 1][w], 1, LW), b2 = __shfl_down_sync(MAPF_FULL_MASK, m2[0][w], 1, LW);
            uint32_t az = __shfl_up_sync(MAPF_FULL_MASK, z[RPL - 1][w], 1, LW), bz = __shfl_down_sync(MAPF_FULL_MASK, z[0][w], 1, LW);
            if (lane == 0) a1 = a2 = az = 0;
            if (lane == LW - 1) b1 = b2 = bz = 0;
#pragma unroll
            for (int q = 0; q < RPL; ++q) {
                auto left = [&](const uint32_t(&m)[RPL][RW]) {  // bit y <- bit y-1
                    return w > 0 ? __funnelshift_l(m[q][w > 0 ? w - 1 : 0], m[q][w], 1) : m[q][w] << 1;
                };
                auto right = [&](const uint32_t(&m)[RPL][RW]) {  // bit y <- bit y+1
                    return w < RW - 1 ? __funnelshift_r(m[q][w], m[q][w < RW - 1 ? w + 1 : w], 1) : m[q][w] >> 1;
                };
                const uint32_t c1 = m1[q][w], c2 = m2[q][w], cz = z[q][w];
                const uint32_t u1 = q > 0 ? m1[q > 0 ? q - 1 : 0][w] : a1, u2 = q > 0 ? m2[q > 0 ? q - 1 : 0][w] : a2,
                               uz = q > 0 ? z[q > 0 ? q - 1 : 0][w] : az;
                const uint32_t d1 = q < RPL - 1 ? m1[q < RPL - 1 ? q + 1 : q][w] : b1, d2 = q < RPL - 1 ? m2[q < RPL - 1 ? q + 1 : q][w] : b2,
                               dz = q < RPL - 1 ? z[q < RPL - 1 ? q + 1 : q][w] : bz;
                pl[0][q][w] = closer(c1, c2, cz, u1, u2, uz);                       // neighbour x-1   environment.py:260
                pl[1][q][w] = closer(c1, c2, cz, d1, d2, dz);                       // neighbour x+1   environment.py:264
                pl[2][q][w] = closer(c1, c2, cz, left(m1), left(m2), left(z));      // neighbour y-1   environment.py:268
                pl[3][q][w] = closer(c1, c2, cz, right(m1), right(m2), right(z));   // neighbour y+1   environment.py:272
            }
        }
    }

    // emit the overlapping 16x16 tiles (mapf_common.cuh): a map row is padded row pr = row + 4, which is
    // row pr & 7 of tile row-block pr >> 3 and row (pr & 7) + 8 of the block above
    uint2 *nv = reinterpret_cast<uint2 *>(navi + ((size_t)e * d.N + a) * d.navi_agent_stride);
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
        const int row = lane * RPL + q;
        if (row >= d.L || !alive) continue;
        const int pr = row + 4, bx1 = pr >> 3, r1 = pr & 7;
#pragma unroll
        for (int by = 0; by < 4 * RW; ++by) {
            if (by >= d.NB) continue;
            constexpr int kLast = RW - 1;
            const int w = by >> 2;          // compile-time after unrolling
            const int s = (by & 3) * 8;
            uint32_t f[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                f[k] = __funnelshift_r(pl[k][q][w], w < kLast ? pl[k][q][w < kLast ? w + 1 : kLast] : 0u, s) & 0xffffu;
            const uint2 v = make_uint2(f[0] | (f[1] << 16), f[2] | (f[3] << 16));
            if (bx1 < d.NB) nv[((size_t)(bx1 * d.NB + by) << 4) + r1] = v;
            if (bx1 > 0) nv[((size_t)((bx1 - 1) * d.NB + by) << 4) + r1 + 8] = v;
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
        uint32_t b0 = 0xffffffffu, b1 = 0xffffffffu, b2 = 0xffffffffu;  // three smallest keys, ascending
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
            if (q < k_nearest && best[q] != 0xffffffffu) {
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
    for (int idx = lane; idx < N * N; idx += 32) {
        const int i = idx / N, j = idx - i * N;
        o[idx] = (s_bits[warp][i][j >> 5] >> (j & 31)) & 1u;
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

__global__ void unpack_navi_kernel(EnvDims d, const uint32_t *__restrict__ navi, uint8_t *__restrict__ navi_out)
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
    const uint2 v = reinterpret_cast<const uint2 *>(navi + ea * d.navi_agent_stride)[((size_t)(bx * d.NB + by) << 4) + (pr - 8 * bx)];
    const uint32_t half = k < 2 ? v.x : v.y;
    navi_out[idx] = (half >> (16 * (k & 1) + (pc - 8 * by))) & 1u;
}

}  // namespace

// ---- entry points used by mapf_abi.cu -----------------------------------------------------------
int mapf_launch_pack_load(mapf_env *env, const int32_t *d_env_ids, int n, const uint8_t *d_maps,
                          const uint8_t *d_agents, const uint8_t *d_goals, cudaStream_t st)
{
    pack_load_kernel<<<n, 128, 0, st>>>(env->d, d_env_ids, n, d_maps, d_agents, d_goals, env->obst, env->pos, env->goal,
                                        env->steps);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

template <int RW>
static int launch_bfs_rw(mapf_env *env, const int32_t *ids, const uint8_t *mask, int n, int32_t *dist, cudaStream_t st)
{
    const EnvDims &d = env->d;
    const int apw = d.L <= 48 ? 2 : 1;       // two agents per warp while the map fits 16 lanes x 3 rows
    const int rpl = (d.L + 32 / apw - 1) / (32 / apw);
    const long warps = ((long)n * d.N + apw - 1) / apw;
    const int grid = (int)((warps + kBfsWarps - 1) / kBfsWarps);
#define MAPF_BFS_LAUNCH(RPL, APW) \
    bfs_navi_kernel<RW, RPL, APW><<<grid, kBfsWarps * 32, 0, st>>>(d, ids, mask, n, env->obst, env->goal, env->navi, dist)
    bool ok = true;
    if (apw == 2) {
        if constexpr (RW <= 2) {
            switch (rpl) {
                case 1: MAPF_BFS_LAUNCH(1, 2); break;
                case 2: MAPF_BFS_LAUNCH(2, 2); break;
                case 3: MAPF_BFS_LAUNCH(3, 2); break;
                default: ok = false;
            }
        } else ok = false;
    } else {
        if constexpr (RW >= 2) {
            switch (rpl) {
                case 2: MAPF_BFS_LAUNCH(2, 1); break;
                case 3: MAPF_BFS_LAUNCH(3, 1); break;
                case 4: MAPF_BFS_LAUNCH(4, 1); break;
                default: ok = false;
            }
        } else ok = false;
    }
#undef MAPF_BFS_LAUNCH
    if (!ok) {
        mapf_set_error("unsupported map size");
        return MAPF_EINVAL;
    }
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

static int launch_bfs(mapf_env *env, const int32_t *d_env_ids, const uint8_t *d_mask, int n, int32_t *d_dist_out,
                      cudaStream_t st)
{
    switch (env->d.RW) {
        case 1: return launch_bfs_rw<1>(env, d_env_ids, d_mask, n, d_dist_out, st);
        case 2: return launch_bfs_rw<2>(env, d_env_ids, d_mask, n, d_dist_out, st);
        case 3: return launch_bfs_rw<3>(env, d_env_ids, d_mask, n, d_dist_out, st);
        case 4: return launch_bfs_rw<4>(env, d_env_ids, d_mask, n, d_dist_out, st);
    }
    mapf_set_error("unsupported map size");
    return MAPF_EINVAL;
}

int mapf_launch_bfs(mapf_env *env, const int32_t *d_env_ids, int n, int32_t *d_dist_out, cudaStream_t st)
{
    return launch_bfs(env, d_env_ids, nullptr, n, d_dist_out, st);
}

// all B slots, skipping those whose mask byte is zero (NULL mask = all)
int mapf_launch_bfs_masked(mapf_env *env, const uint8_t *d_mask, cudaStream_t st)
{
    return launch_bfs(env, nullptr, d_mask, env->d.B, nullptr, st);
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
        unpack_navi_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d, env->navi, d_navi);
        MAPF_CUDA(cudaGetLastError());
    }
    return MAPF_OK;
}
