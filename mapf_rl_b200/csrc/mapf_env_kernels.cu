// mapf_env_kernels.cu — the environment hot path as hand-written CUDA for sm_100a.
//
//   K1+K2  step_observe_kernel   Environment.step (environment.py:278-430) fused with the observe() it
//                                ends in (environment.py:430,433-467).  One warp per environment,
//                                lane = agent (K = ceil(N/32) agents per lane).
//   K3     bfs_navi_kernel       Environment.get_navi_map (environment.py:217-276): bit-parallel
//                                wavefront BFS, one warp per (env, agent), lane = map row.
//   pack / unpack helpers        Environment.load (environment.py:198-215) and attribute reads.
//
// None of this is a dense contraction: no tensor cores.  K1+K2 is bound by the HBM write of the
// observation bytes (486 B per agent-step); everything else stays in shared memory / registers.
#include "mapf_common.cuh"

namespace {

constexpr int kStepWarps = 8;  // warps (= environments in flight) per CTA of the step kernel
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
// Lane l owns map rows l, l+32, ... (RPL rows), each row RW words of padded column bits.  One wave:
//   new = (frontier shifted to the four neighbours) & free & ~visited
// and the heuristic bit "neighbour in direction d is strictly closer" (environment.py:260-274) is
// exactly "this cell is new in wave t and that neighbour was in the frontier of wave t-1", so the
// four direction planes fall out of the same step and no distance array is needed.
// ---------------------------------------------------------------------------------------------
template <int RW, int RPL>
__global__ void __launch_bounds__(kBfsWarps * 32)
bfs_navi_kernel(EnvDims d, const int32_t *__restrict__ env_ids, int n, const uint32_t *__restrict__ obst,
                const uint8_t *__restrict__ goal, uint32_t *__restrict__ navi, int32_t *__restrict__ dist_out)
{
    const int lane = lane_id();
    const int g = blockIdx.x * kBfsWarps + (threadIdx.x >> 5);
    if (g >= n * d.N) return;
    const int i = g / d.N, a = g - i * d.N;
    const int e = env_ids ? env_ids[i] : i;
    const uint32_t *ob = obst + (size_t)e * d.obst_stride;

    uint32_t fre[RPL][RW], vis[RPL][RW], fro[RPL][RW], pl[4][RPL][RW];
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
        int row = lane + 32 * q;
#pragma unroll
        for (int w = 0; w < RW; ++w) {
            // in-map column mask for this word: padded bits [4, L+4)
            int lo = max(4 - 32 * w, 0), hi = min(d.L + 4 - 32 * w, 32);
            uint32_t cm = 0;
            if (hi > lo) cm = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
            fre[q][w] = (row < d.L) ? (~ob[(row + 4) * d.RWS + w] & cm) : 0u;
            vis[q][w] = 0;
            fro[q][w] = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) pl[k][q][w] = 0;
        }
    }
    const int gx = goal[((size_t)e * d.N + a) * 2], gy = goal[((size_t)e * d.N + a) * 2 + 1];
#pragma unroll
    for (int q = 0; q < RPL; ++q)
#pragma unroll
        for (int w = 0; w < RW; ++w) {
            int p = gy + 4;
            if (lane + 32 * q == gx && (p >> 5) == w) fro[q][w] = (1u << (p & 31)) & fre[q][w];
            vis[q][w] = fro[q][w];
        }

    int32_t *dist = dist_out ? dist_out + ((size_t)i * d.N + a) * d.L * d.L : nullptr;
    if (dist) {
        for (int q = 0; q < RPL; ++q) {
            int row = lane + 32 * q;
            if (row < d.L)
                for (int y = 0; y < d.L; ++y) dist[row * d.L + y] = MAPF_DIST_UNREACHABLE;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < RPL; ++q)
#pragma unroll
            for (int w = 0; w < RW; ++w)
                if (fro[q][w]) dist[gx * d.L + gy] = 0;
    }

    for (int t = 1;; ++t) {
        uint32_t nw[RPL][RW];
        uint32_t any = 0;
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
#pragma unroll
            for (int w = 0; w < RW; ++w) {
                uint32_t f = fro[q][w];
                uint32_t fl = f << 1, fr = f >> 1;
                if (w > 0) fl |= fro[q][w - 1] >> 31;
                if (w < RW - 1) fr |= fro[q][w + 1] << 31;
                uint32_t up = __shfl_up_sync(MAPF_FULL_MASK, f, 1);
                uint32_t dn = __shfl_down_sync(MAPF_FULL_MASK, f, 1);
                uint32_t wrap_up = 0, wrap_dn = 0;
                if (q > 0) wrap_up = __shfl_sync(MAPF_FULL_MASK, fro[q > 0 ? q - 1 : 0][w], 31);
                if (q < RPL - 1) wrap_dn = __shfl_sync(MAPF_FULL_MASK, fro[q < RPL - 1 ? q + 1 : q][w], 0);
                if (lane == 0) up = wrap_up;
                if (lane == 31) dn = wrap_dn;
                uint32_t x = (fl | fr | up | dn) & fre[q][w] & ~vis[q][w];
                nw[q][w] = x;
                pl[0][q][w] |= x & up;  // neighbour x-1 (row above) is closer   environment.py:260
                pl[1][q][w] |= x & dn;  // neighbour x+1                          environment.py:264
                pl[2][q][w] |= x & fl;  // neighbour y-1                          environment.py:268
                pl[3][q][w] |= x & fr;  // neighbour y+1                          environment.py:272
                any |= x;
            }
        }
        if (!__any_sync(MAPF_FULL_MASK, any != 0)) break;
#pragma unroll
        for (int q = 0; q < RPL; ++q)
#pragma unroll
            for (int w = 0; w < RW; ++w) {
                vis[q][w] |= nw[q][w];
                fro[q][w] = nw[q][w];
                if (dist) {
                    uint32_t x = nw[q][w];
                    while (x) {
                        int b = __ffs(x) - 1;
                        x &= x - 1;
                        dist[(lane + 32 * q) * d.L + (32 * w + b - 4)] = t;
                    }
                }
            }
    }

    // emit [cb][row] words: byte k = 8 column bits of direction k
    uint32_t *nv = navi + ((size_t)e * d.N + a) * d.navi_agent_stride;
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
        int row = lane + 32 * q;
#pragma unroll
        for (int w = 0; w < RW; ++w)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int cb = 4 * w + b;
                if (cb < d.CB && row < d.L) {
                    uint32_t v = ((pl[0][q][w] >> (8 * b)) & 0xffu) | (((pl[1][q][w] >> (8 * b)) & 0xffu) << 8) |
                                 (((pl[2][q][w] >> (8 * b)) & 0xffu) << 16) | (((pl[3][q][w] >> (8 * b)) & 0xffu) << 24);
                    nv[cb * d.R + row + 4] = v;
                }
            }
    }
}

// ---------------------------------------------------------------------------------------------
// K1 + K2: fused step + observe
// ---------------------------------------------------------------------------------------------
struct StepParams {
    EnvDims d;
    const uint32_t *obst;
    uint8_t *pos;
    const uint8_t *goal;
    const uint32_t *navi;
    int32_t *steps;
    int32_t *err;
    const uint8_t *actions;  // [B,N]           (step only)
    uint8_t *obs;            // [B,N,6,9,9]
    float *rewards;          // [B,N]           (step only)
    uint8_t *done;           // [B]             (step only)
    int32_t *steps_out;      // [B] optional
    uint8_t *pos_out;        // [B,N,2] optional (observe only)
    float r_move, r_stay_on, r_stay_off, r_collision, r_finish;
    int warp_smem_words;     // per-warp shared memory, multiple of 4 words
    int obst_words;          // = d.obst_stride
    int bits_words;          // words of the per-env observation bit stream
};

// 4 bits -> 4 bool bytes: bit b lands at bit 8b.  The four shifted copies of x (shifts 0,7,14,21)
// do not overlap for x < 16, so the multiply has no carries.
__device__ __forceinline__ uint32_t expand4(uint32_t x) { return (x * 0x00204081u) & 0x01010101u; }

__device__ __forceinline__ uint32_t window9(const uint32_t *row, int bitoff)
{
    int w = bitoff >> 5;
    return __funnelshift_r(row[w], row[w + 1], bitoff & 31) & 0x1ffu;
}

// OR a 9-bit row into the agent-local 486-bit stream at a compile-time bit position.
template <int POS>
__device__ __forceinline__ void put9(uint32_t (&W)[16], uint32_t v)
{
    constexpr int w = POS >> 5, s = POS & 31;
    W[w] |= v << s;
    if constexpr (s > 23) W[w + 1] |= v >> (32 - s);
}

template <int U>
struct RowBuilder {
    template <typename F>
    __device__ __forceinline__ static void run(F &&f)
    {
        f(std::integral_constant<int, U>{});
        if constexpr (U + 1 < 9) RowBuilder<U + 1>::run(f);
    }
};

template <int RW, int K, bool DO_STEP>
__global__ void __launch_bounds__(kStepWarps * 32)
step_observe_kernel(const StepParams p)
{
    constexpr int RWS = RW + 1;
    extern __shared__ __align__(16) uint32_t smem[];
    const EnvDims &d = p.d;
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int N = d.N, L = d.L;

    uint32_t *s_obst = smem + (size_t)warp * p.warp_smem_words;
    uint32_t *s_agent = s_obst + p.obst_words;
    uint32_t *s_bits = s_agent + p.obst_words;
    uint16_t *s_tgt = reinterpret_cast<uint16_t *>(s_bits + p.bits_words);
    uint16_t *s_cell = s_tgt + 32 * K;
    uint8_t *s_occ = reinterpret_cast<uint8_t *>(s_cell + 32 * K);

    // the agent bitmap must start all-zero; afterwards each env clears the bits it set
    for (int w = lane; w < p.obst_words; w += 32) s_agent[w] = 0;
    __syncwarp();

    for (int e = blockIdx.x * kStepWarps + warp; e < d.B; e += gridDim.x * kStepWarps) {
        // ---- stage the obstacle bitmap (L2-resident, 16-byte vectors) ----
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(p.obst + (size_t)e * d.obst_stride);
            uint4 *dst = reinterpret_cast<uint4 *>(s_obst);
            for (int w = lane; w < (p.obst_words >> 2); w += 32) dst[w] = __ldg(src + w);
        }

        int px[K], py[K];
        bool valid[K];
        float rew[K];
        bool at_goal[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int a = k * 32 + lane;
            valid[k] = a < N;
            px[k] = py[k] = 0;
            rew[k] = 0.f;
            at_goal[k] = true;
            if (valid[k]) {
                uchar2 pp = reinterpret_cast<const uchar2 *>(p.pos)[(size_t)e * N + a];
                px[k] = pp.x;
                py[k] = pp.y;
            }
        }

        if constexpr (DO_STEP) {
            int gx[K], gy[K], tx[K], ty[K], act[K], tcell[K], mycell[K], occ_j[K];
            bool mover[K], occ_ok[K], fail[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int a = k * 32 + lane;
                act[k] = 0;
                gx[k] = gy[k] = 0;
                if (valid[k]) {
                    uchar2 gg = reinterpret_cast<const uchar2 *>(p.goal)[(size_t)e * N + a];
                    gx[k] = gg.x;
                    gy[k] = gg.y;
                    act[k] = p.actions[(size_t)e * N + a];
                    if (act[k] > 4) {  // environment.py:289-290 raises; we latch and treat as stay
                        atomicOr(p.err, MAPF_ERRBIT_ACTION);
                        act[k] = 0;
                    }
                }
                mycell[k] = px[k] * L + py[k];
                s_cell[a] = valid[k] ? (uint16_t)mycell[k] : (uint16_t)0xffff;
                if (valid[k]) s_occ[mycell[k]] = (uint8_t)a;
            }
            __syncwarp();  // staged obstacle bitmap, s_cell and s_occ visible to every lane
#pragma unroll
            for (int k = 0; k < K; ++k) {
                // stay / move pass, environment.py:298-311
                const bool on_goal = px[k] == gx[k] && py[k] == gy[k];
                rew[k] = act[k] == 0 ? (on_goal ? p.r_stay_on : p.r_stay_off) : p.r_move;
                mover[k] = valid[k] && act[k] != 0;
                // action table environment.py:12
                tx[k] = px[k] + (act[k] == 2) - (act[k] == 1);
                ty[k] = py[k] + (act[k] == 4) - (act[k] == 3);
                tcell[k] = tx[k] * L + ty[k];
                if (mover[k]) {
                    // round 1: out of range / obstacle, environment.py:320-332
                    bool bad = tx[k] < 0 || ty[k] < 0 || tx[k] >= L || ty[k] >= L;
                    if (!bad) bad = (s_obst[(tx[k] + 4) * RWS + ((ty[k] + 4) >> 5)] >> ((ty[k] + 4) & 31)) & 1u;
                    if (bad) {
                        rew[k] = p.r_collision;
                        mover[k] = false;
                    }
                }
                s_tgt[k * 32 + lane] = mover[k] ? (uint16_t)tcell[k] : (uint16_t)0xffff;
            }
            __syncwarp();
            // round 2: swap, environment.py:335-365 (order-independent form: both partners revert)
            bool swapped[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                occ_j[k] = 0xff;
                occ_ok[k] = false;
                swapped[k] = false;
                if (mover[k]) {
                    const int j = s_occ[tcell[k]];
                    // the occupancy grid is never cleared: an entry is real iff it round-trips
                    occ_ok[k] = j < N && s_cell[j] == (uint16_t)tcell[k];
                    occ_j[k] = j;
                    swapped[k] = occ_ok[k] && s_tgt[j] == (uint16_t)mycell[k];
                }
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < K; ++k)
                if (swapped[k]) {
                    s_tgt[k * 32 + lane] = 0xffff;
                    mover[k] = false;
                    rew[k] = p.r_collision;
                }
            __syncwarp();
            // round 3: vertex conflicts, environment.py:369-406, as the greatest fixed point:
            //   fail if the target's occupant is not a live mover,
            //   fail if not the lowest id among live movers with the same target (:389-394),
            //   fail if the target's occupant is a live mover that fails (propagates backwards).
#pragma unroll
            for (int k = 0; k < K; ++k) {
                fail[k] = mover[k] && occ_ok[k] && s_tgt[occ_j[k]] == 0xffff;
                bool lower_claim = false;
                if (K > 1 && k > 0 && mover[k]) {
                    const int c = s_occ[tcell[k]];  // claim left by a lower slot (verified, never cleared)
                    lower_claim = c < N && (c >> 5) < k && s_tgt[c] == (uint16_t)tcell[k];
                }
                const unsigned code = mover[k] ? (unsigned)tcell[k] : (0x10000u | lane);
                const unsigned m = __match_any_sync(MAPF_FULL_MASK, code);
                const bool first = (__ffs(m) - 1) == lane;
                if (mover[k] && (!first || lower_claim)) fail[k] = true;
                if (K > 1 && k + 1 < K) {
                    __syncwarp();
                    if (mover[k]) s_occ[tcell[k]] = (uint8_t)(k * 32 + lane);
                    __syncwarp();
                }
            }
            for (;;) {
                unsigned fm[K];
#pragma unroll
                for (int k = 0; k < K; ++k) fm[k] = __ballot_sync(MAPF_FULL_MASK, fail[k]);
                bool changed = false;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    if (mover[k] && !fail[k] && occ_ok[k]) {
                        const int j = occ_j[k];
                        unsigned mj = fm[0];
#pragma unroll
                        for (int q = 1; q < K; ++q)
                            if ((j >> 5) == q) mj = fm[q];
                        if ((mj >> (j & 31)) & 1u) {
                            fail[k] = true;
                            changed = true;
                        }
                    }
                }
                if (!__any_sync(MAPF_FULL_MASK, changed)) break;
            }
            // commit, environment.py:410-421
            bool all_goal = true;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (mover[k]) {
                    if (fail[k]) rew[k] = p.r_collision;
                    else {
                        px[k] = tx[k];
                        py[k] = ty[k];
                    }
                }
                at_goal[k] = !valid[k] || (px[k] == gx[k] && py[k] == gy[k]);
                all_goal = all_goal && at_goal[k];
            }
            const bool done = __all_sync(MAPF_FULL_MASK, all_goal);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int a = k * 32 + lane;
                if (valid[k]) {
                    reinterpret_cast<uchar2 *>(p.pos)[(size_t)e * N + a] = make_uchar2((unsigned char)px[k], (unsigned char)py[k]);
                    p.rewards[(size_t)e * N + a] = done ? p.r_finish : rew[k];
                }
            }
            if (lane == 0) {
                const int st = p.steps[e] + 1;
                p.steps[e] = st;
                if (p.steps_out) p.steps_out[e] = st;
                p.done[e] = done ? 1 : 0;
            }
        } else {
            if (p.pos_out) {
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (valid[k])
                        reinterpret_cast<uchar2 *>(p.pos_out)[(size_t)e * N + k * 32 + lane] =
                            make_uchar2((unsigned char)px[k], (unsigned char)py[k]);
            }
            __syncwarp();  // s_obst visible
        }

        // ---------------- observe, environment.py:433-467 ----------------
        // agent bitmap (environment.py:449-451): one shared-memory atomic per agent
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (valid[k]) atomicOr(&s_agent[(px[k] + 4) * RWS + ((py[k] + 4) >> 5)], 1u << ((py[k] + 4) & 31));
        __syncwarp();

        const size_t env_bytes = (size_t)N * MAPF_OBS_BYTES_PER_AGENT;
        uint8_t *obs_env = p.obs + (size_t)e * env_bytes;
        const int head = (int)(reinterpret_cast<uintptr_t>(obs_env) & 15);  // bytes before the 16-B boundary

#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int a = k * 32 + lane;
            uint32_t W[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) W[m] = 0;
            if (valid[k]) {
                const int x = px[k], y = py[k];
                // window rows x-4..x+4 are padded rows x..x+8; columns y-4..y+4 are padded bits y..y+8
                const uint32_t *nb = p.navi + ((size_t)e * N + a) * d.navi_agent_stride + (size_t)(y >> 3) * d.R + x;
                uint32_t wa[9], wb[9];
#pragma unroll
                for (int u = 0; u < 9; ++u) {
                    wa[u] = __ldg(nb + u);
                    wb[u] = __ldg(nb + d.R + u);
                }
                const int sh = y & 7;
                RowBuilder<0>::run([&](auto uc) {
                    constexpr int u = decltype(uc)::value;
                    uint32_t v = window9(s_agent + (x + u) * RWS, y);
                    if constexpr (u == 4) v &= ~0x10u;  // own centre cleared, environment.py:461
                    put9<0 * 81 + 9 * u>(W, v);
                    put9<1 * 81 + 9 * u>(W, window9(s_obst + (x + u) * RWS, y));
                    put9<2 * 81 + 9 * u>(W, (__byte_perm(wa[u], wb[u], 0x0040) >> sh) & 0x1ffu);
                    put9<3 * 81 + 9 * u>(W, (__byte_perm(wa[u], wb[u], 0x0051) >> sh) & 0x1ffu);
                    put9<4 * 81 + 9 * u>(W, (__byte_perm(wa[u], wb[u], 0x0062) >> sh) & 0x1ffu);
                    put9<5 * 81 + 9 * u>(W, (__byte_perm(wa[u], wb[u], 0x0073) >> sh) & 0x1ffu);
                });
            }
            // place the 486 bits at stream bit (head + 486 a): words f .. f + last
            const int gbit = head + MAPF_OBS_BYTES_PER_AGENT * a;
            const int o = gbit & 31, f = gbit >> 5;
            uint32_t X[17];
            X[0] = W[0] << o;
#pragma unroll
            for (int m = 1; m < 16; ++m) X[m] = __funnelshift_l(W[m - 1], W[m], o);
            X[16] = __funnelshift_l(W[15], 0u, o);
            const int last = (o + 485) >> 5;  // 15 or 16
            uint32_t *S = s_bits + f;
            // neighbouring agents share boundary words: even agents store first, odd agents OR in.
            if (valid[k] && !(lane & 1)) {
                if (k > 0 && lane == 0 && o != 0) S[0] |= X[0];  // agent 32k-1 already stored its last word
                else S[0] = X[0];
#pragma unroll
                for (int m = 1; m < 16; ++m) S[m] = X[m];
                if (last == 16) S[16] = X[16];
            }
            __syncwarp();
            if (valid[k] && (lane & 1)) {
                if (o != 0) S[0] |= X[0];
                else S[0] = X[0];
#pragma unroll
                for (int m = 1; m < 15; ++m) S[m] = X[m];
                const bool share_next = (((o + 486) & 31) != 0) && (a + 1 < N) && lane != 31;
                if (last == 16) {
                    S[15] = X[15];
                    if (share_next) S[16] |= X[16];
                    else S[16] = X[16];
                } else {
                    if (share_next) S[15] |= X[15];
                    else S[15] = X[15];
                }
            }
            __syncwarp();
        }

        // expand 1 bit -> 1 bool byte, 16 bytes per lane per store, fully coalesced streaming stores
        {
            const int total = head + (int)env_bytes;
            const int nchunks = (total + 15) >> 4;
            uint8_t *obase = obs_env - head;  // 16-byte aligned
            const uint16_t *S16 = reinterpret_cast<const uint16_t *>(s_bits);
            for (int c = lane; c < nchunks; c += 32) {
                const uint32_t s = S16[c];
                uint4 v;
                v.x = expand4(s & 0xfu);
                v.y = expand4((s >> 4) & 0xfu);
                v.z = expand4((s >> 8) & 0xfu);
                v.w = expand4(s >> 12);
                const int b0 = c << 4;
                if (b0 >= head && b0 + 16 <= total) {
                    __stcs(reinterpret_cast<uint4 *>(obase + b0), v);
                } else {
                    const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
                    for (int b = 0; b < 16; ++b)
                        if (b0 + b >= head && b0 + b < total) obase[b0 + b] = (uint8_t)((vv[b >> 2] >> (8 * (b & 3))) & 0xffu);
                }
            }
        }
        __syncwarp();
        // clear the agent bits this env set
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (valid[k]) s_agent[(px[k] + 4) * RWS + ((py[k] + 4) >> 5)] = 0;
        __syncwarp();
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
    uint32_t w = navi[ea * d.navi_agent_stride + (size_t)((y + 4) >> 3) * d.R + x + 4];
    navi_out[idx] = (w >> (8 * k + ((y + 4) & 7))) & 1u;
}

template <int RW, int K, bool DO_STEP>
int launch_step_rwk(const mapf_env *env, StepParams &p, cudaStream_t st)
{
    auto kern = step_observe_kernel<RW, K, DO_STEP>;
    const size_t smem = (size_t)p.warp_smem_words * 4 * kStepWarps;
    if (smem > 48 * 1024)  // per-device attribute; cheap enough to set on every large-smem launch
        MAPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (env->d.B + kStepWarps - 1) / kStepWarps;
    kern<<<grid, kStepWarps * 32, smem, st>>>(p);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

template <int RW, bool DO_STEP>
int launch_step_rw(const mapf_env *env, StepParams &p, cudaStream_t st)
{
    switch (env->d.K) {
        case 1: return launch_step_rwk<RW, 1, DO_STEP>(env, p, st);
        case 2: return launch_step_rwk<RW, 2, DO_STEP>(env, p, st);
        case 3: return launch_step_rwk<RW, 3, DO_STEP>(env, p, st);
        case 4: return launch_step_rwk<RW, 4, DO_STEP>(env, p, st);
    }
    mapf_set_error("unsupported agent count");
    return MAPF_EINVAL;
}

template <bool DO_STEP>
int launch_step(const mapf_env *env, StepParams &p, cudaStream_t st)
{
    switch (env->d.RW) {
        case 1: return launch_step_rw<1, DO_STEP>(env, p, st);
        case 2: return launch_step_rw<2, DO_STEP>(env, p, st);
        case 3: return launch_step_rw<3, DO_STEP>(env, p, st);
        case 4: return launch_step_rw<4, DO_STEP>(env, p, st);
    }
    mapf_set_error("unsupported map size");
    return MAPF_EINVAL;
}

StepParams make_params(const mapf_env *env)
{
    StepParams p{};
    const EnvDims &d = env->d;
    p.d = d;
    p.obst = env->obst;
    p.pos = env->pos;
    p.goal = env->goal;
    p.navi = env->navi;
    p.steps = env->steps;
    p.err = env->err;
    p.r_move = env->reward[0];
    p.r_stay_on = env->reward[1];
    p.r_stay_off = env->reward[2];
    p.r_collision = env->reward[3];
    p.r_finish = env->reward[4];
    p.obst_words = d.obst_stride;
    // stream words: 15 head bits max + N*486 bits, +1 word of slack for the u16 tail read
    p.bits_words = ((15 + d.N * MAPF_OBS_BYTES_PER_AGENT + 31) >> 5) + 2;
    int words = 2 * p.obst_words + p.bits_words + (32 * d.K) /* s_tgt+s_cell u16 */ + ((d.L * d.L + 3) >> 2);
    p.warp_smem_words = (words + 3) & ~3;
    return p;
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
static int launch_bfs_rw(mapf_env *env, const int32_t *ids, int n, int32_t *dist, cudaStream_t st)
{
    const EnvDims &d = env->d;
    const int rpl = (d.L + 31) / 32;
    const long warps = (long)n * d.N;
    const int grid = (int)((warps + kBfsWarps - 1) / kBfsWarps);
    switch (rpl) {
        case 1: bfs_navi_kernel<RW, 1><<<grid, kBfsWarps * 32, 0, st>>>(d, ids, n, env->obst, env->goal, env->navi, dist); break;
        case 2: bfs_navi_kernel<RW, 2><<<grid, kBfsWarps * 32, 0, st>>>(d, ids, n, env->obst, env->goal, env->navi, dist); break;
        case 3: bfs_navi_kernel<RW, 3><<<grid, kBfsWarps * 32, 0, st>>>(d, ids, n, env->obst, env->goal, env->navi, dist); break;
        case 4: bfs_navi_kernel<RW, 4><<<grid, kBfsWarps * 32, 0, st>>>(d, ids, n, env->obst, env->goal, env->navi, dist); break;
        default: mapf_set_error("unsupported map size"); return MAPF_EINVAL;
    }
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

int mapf_launch_bfs(mapf_env *env, const int32_t *d_env_ids, int n, int32_t *d_dist_out, cudaStream_t st)
{
    switch (env->d.RW) {
        case 1: return launch_bfs_rw<1>(env, d_env_ids, n, d_dist_out, st);
        case 2: return launch_bfs_rw<2>(env, d_env_ids, n, d_dist_out, st);
        case 3: return launch_bfs_rw<3>(env, d_env_ids, n, d_dist_out, st);
        case 4: return launch_bfs_rw<4>(env, d_env_ids, n, d_dist_out, st);
    }
    mapf_set_error("unsupported map size");
    return MAPF_EINVAL;
}

int mapf_launch_step(mapf_env *env, const uint8_t *d_actions, uint8_t *d_obs, float *d_rewards, uint8_t *d_done,
                     int32_t *d_steps, cudaStream_t st)
{
    StepParams p = make_params(env);
    p.actions = d_actions;
    p.obs = d_obs;
    p.rewards = d_rewards;
    p.done = d_done;
    p.steps_out = d_steps;
    return launch_step<true>(env, p, st);
}

int mapf_launch_observe(mapf_env *env, uint8_t *d_obs, uint8_t *d_pos, cudaStream_t st)
{
    StepParams p = make_params(env);
    p.obs = d_obs;
    p.pos_out = d_pos;
    return launch_step<false>(env, p, st);
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
