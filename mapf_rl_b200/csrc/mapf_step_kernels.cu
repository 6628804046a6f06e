// mapf_step_kernels.cu — K1 + K2: Environment.step (environment.py:278-430) fused with the observe()
// it ends in (environment.py:430, 433-467), hand-written for sm_100a.
//
// One warp per environment, lane = agent (K = ceil(N/32) agents per lane).  Per environment:
//   1. all inputs of the env (positions, goals, actions, step counter, obstacle bitmap) are requested
//      up front so their DRAM/L2 latencies overlap;
//   2. conflict resolution runs in registers / shared memory (order-independent fixed point of the
//      reference's restart-on-change scans);
//   3. every lane gathers its agent's 6 x 9 x 9 window as a 486-BIT stream (obstacle / agent bitmaps in
//      shared memory, heuristic bits straight from the nibble-planar navi array) and streams it, word
//      by word, into the env's dense bit stream in shared memory;
//   4. the warp expands bits to bool bytes and writes the env's N*486-byte block with fully coalesced
//      128-bit stores.
// HBM-bound (486 B written per agent-step); nothing here is a dense contraction, so no tensor cores.
#include <cstdlib>
#include <type_traits>

#include "mapf_common.cuh"

namespace {

struct StepParams {
    EnvDims d;
    const uint32_t *obst;
    uint8_t *pos;
    const uint8_t *goal;
    const uint32_t *navi;
    int32_t *steps;
    int32_t *err;
    const uint8_t *actions;  // [B,N]           (step only)
    uint8_t *obs;            // [B,N,6,9,9], or the base of a replay store when obs_rows is given
    const int64_t *obs_rows; // optional [B]: env e writes its N*486-byte block at row obs_rows[e] of `obs`
    float *rewards;          // [B,N]           (step only)
    uint8_t *done;           // [B]             (step only)
    int32_t *steps_out;      // [B] optional
    uint8_t *pos_out;        // [B,N,2] optional (observe only)
    float r_move, r_stay_on, r_stay_off, r_collision, r_finish;
    int warp_smem_words;     // per-warp shared memory, multiple of 4 words
    int obst_words;          // = d.obst_stride
    int bits_words;          // words of the per-env observation bit stream (also holds the occupancy grid)
    int flags;               // MAPF_STEPF_*
    unsigned long long *trace;  // diagnosis: u64[B][16] globaltimer stamps per env (NULL = off), see profiles/step_timeline.py
    int chunks_per_env;      // split form: N * 486 / 16 16-byte output chunks (= 16-bit stream pieces) per env
    int env_begin, env_end;  // single-role kernel: the launch covers environments [env_begin, env_end) of the batch
};

enum : int {
    MAPF_STEPF_NAVI_KEEP = 1,    // heuristic-map loads carry an L2 evict_last policy
    MAPF_STEPF_OBS_POLICY = 2,   // observation stores carry an L2 evict_first policy (else st.global.cs)
    // diagnosis only (results are WRONG with these set; profiles/step_variants.py uses them to bound the kernel)
    MAPF_STEPF_DIAG_NO_NAVI = 4,   // skip the heuristic-map loads
    MAPF_STEPF_DIAG_NO_STORE = 8,  // skip the observation stores
    MAPF_STEPF_DIAG_NO_STAGE = 128,  // step_only_kernel: every warp loads its own action row (no CTA-wide staging)
};

// 4 bits -> 4 bool bytes: bit b lands at bit 8b.  The four shifted copies of x (shifts 0,7,14,21)
// do not overlap for x < 16, so the multiply has no carries.
__device__ __forceinline__ uint32_t expand4(uint32_t x) { return (x * 0x00204081u) & 0x01010101u; }

__device__ __forceinline__ uint32_t smem_addr(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void trace_stamp(const StepParams &p, int e, int k, int lane)
{
    if (p.trace && lane == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[(size_t)e * 16 + k] = t;
    }
}

__device__ __forceinline__ uint32_t window9(const uint32_t *row, int bitoff)
{
    const int w = bitoff >> 5;
    return __funnelshift_r(row[w], row[w + 1], bitoff & 31) & 0x1ffu;
}

__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint2 ldg_policy(const uint2 *ptr, uint64_t pol)
{
    uint2 v;
    asm("ld.global.nc.L2::cache_hint.v2.b32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(ptr), "l"(pol));
    return v;
}
__device__ __forceinline__ void stg_policy(uint4 *ptr, const uint4 &v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;"
                 :
                 : "l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol)
                 : "memory");
}

// Compile-time walk over the 54 nine-bit fields (channel c = F / 9, window row u = F % 9) of one agent's
// observation bit stream.  A 64-bit accumulator collects fields; every completed 32-bit word is handed to
// `emit` immediately, so at most two stream words are live in registers at any time.
template <int F>
struct FieldWalk {
    template <typename Val, typename Emit>
    __device__ __forceinline__ static void run(uint64_t acc, Val &&val, Emit &&emit)
    {
        constexpr int pos = 9 * F, m = pos >> 5, s = pos & 31;
        acc |= (uint64_t)val(std::integral_constant<int, F>{}) << s;
        if constexpr (s + 9 >= 32) {
            emit(std::integral_constant<int, m>{}, (uint32_t)acc);
            acc >>= 32;
        }
        if constexpr (F + 1 < 54) FieldWalk<F + 1>::run(acc, val, emit);
        else emit(std::integral_constant<int, 15>{}, (uint32_t)acc);  // bits 480..485
    }
};

// Registers of one environment that outlive env_step_gather (final positions, to clear the agent bitmap).
template <int K>
struct EnvRegs {
    int px[K], py[K];
    bool valid[K];
    int gx[K], gy[K];  // goals and step counter: carried from step to step by the persistent rollout kernel only
    int step;
};

// One warp, one environment: Environment.step (DO_STEP) and the observation BIT stream of all its agents.
// On return the env's N*486-bit stream sits in s_bits starting at bit `head` (every lane has passed a
// __syncwarp after its last write), positions / rewards / done / steps are stored, and the agent bitmap
// still holds this env's bits (clear_agent_bits undoes them).
// RESIDENT (persistent rollout kernel): when `carried` is set this is not the warp's first step of the env -- its obstacle
// bitmap is still in s_obst and positions, goals and the step counter arrive in `out` from the previous call instead of
// being loaded, so a step starts with one load (the action) instead of a round of them.
template <int RW, int K, bool DO_STEP, bool TRACE = false, bool DO_OBS = true, bool RESIDENT = false>
__device__ __forceinline__ void env_step_gather(const StepParams &p, const int e, const int lane, uint32_t *s_obst,
                                                uint32_t *s_agent, uint32_t *s_bits, uint16_t *s_tgt, uint16_t *s_cell,
                                                const int head, const uint64_t pol_keep, EnvRegs<K> &out,
                                                const bool carried = false, const uint8_t *s_act = nullptr)
{
    constexpr int RWS = RW + 1;
    const EnvDims &d = p.d;
    const int N = d.N, L = d.L;
    // The cell -> agent grid of the step phase lives in the bit-stream buffer (the two are never live at
    // the same time).  It is never cleared: an entry is trusted only if it round-trips through s_cell.
    uint8_t *s_occ = reinterpret_cast<uint8_t *>(s_bits);
    const bool navi_keep = p.flags & MAPF_STEPF_NAVI_KEEP;
    {
        // ---- request every input of this env up front ----
        if (!(RESIDENT && carried)) {
            // obstacle bitmap: global -> shared without passing through registers (LDGSTS), so nothing below
            // waits for it until the cp.async.wait_all in front of the first __syncwarp
            const uint4 *src = reinterpret_cast<const uint4 *>(p.obst + (size_t)e * d.obst_stride);
            const uint32_t dst = smem_addr(s_obst);
            for (int w = lane; w < (p.obst_words >> 2); w += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * w), "l"(src + w) : "memory");
        }
        int px[K], py[K];
        bool valid[K];
        [[maybe_unused]] int gx[K], gy[K], act[K];
        [[maybe_unused]] int step_now = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int a = k * 32 + lane;
            valid[k] = a < N;
            px[k] = py[k] = 0;
            if constexpr (DO_STEP) gx[k] = gy[k] = act[k] = 0;
            if (RESIDENT && carried) {
                px[k] = out.px[k];
                py[k] = out.py[k];
                if constexpr (DO_STEP) {
                    gx[k] = out.gx[k];
                    gy[k] = out.gy[k];
                    if (valid[k]) act[k] = __ldg(p.actions + (size_t)e * N + a);
                }
            } else if (valid[k]) {
                const uchar2 pp = reinterpret_cast<const uchar2 *>(p.pos)[(size_t)e * N + a];
                px[k] = pp.x;
                py[k] = pp.y;
                if constexpr (DO_STEP) {
                    const uchar2 gg = __ldg(reinterpret_cast<const uchar2 *>(p.goal) + (size_t)e * N + a);
                    gx[k] = gg.x;
                    gy[k] = gg.y;
                    act[k] = s_act ? s_act[a] : __ldg(p.actions + (size_t)e * N + a);
                }
            }
        }
        if constexpr (DO_STEP)
            if (lane == 0) step_now = (RESIDENT && carried) ? out.step : p.steps[e];
        if constexpr (DO_STEP) {
            int tx[K], ty[K], tcell[K], mycell[K], occ_j[K];
            float rew[K];
            bool mover[K], occ_ok[K], fail[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int a = k * 32 + lane;
                if (act[k] > 4) {  // environment.py:289-290 raises; we latch and treat as stay
                    atomicOr(p.err, MAPF_ERRBIT_ACTION);
                    act[k] = 0;
                }
                mycell[k] = px[k] * L + py[k];
                s_cell[a] = valid[k] ? (uint16_t)mycell[k] : (uint16_t)0xffff;
                if (valid[k]) s_occ[mycell[k]] = (uint8_t)a;
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();  // staged obstacle bitmap, s_cell and s_occ visible to every lane
            if constexpr (TRACE) trace_stamp(p, e, 7, lane);
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
            if constexpr (TRACE) trace_stamp(p, e, 8, lane);
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
                all_goal = all_goal && (!valid[k] || (px[k] == gx[k] && py[k] == gy[k]));
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
                const int st = step_now + 1;
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
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();  // s_obst visible
        }

        if constexpr (TRACE) trace_stamp(p, e, 2, lane);
        if constexpr (DO_OBS) {
        // ---------------- observe, environment.py:433-467 ----------------
        // agent bitmap (environment.py:449-451): one shared-memory atomic per agent
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (valid[k]) atomicOr(&s_agent[(px[k] + 4) * RWS + ((py[k] + 4) >> 5)], 1u << ((py[k] + 4) & 31));
        __syncwarp();  // also orders the last s_occ reads before the bit stream overwrites that buffer

#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int a = k * 32 + lane;
            // this agent's 486 bits start at stream bit (head + 486 a) = word f, bit o
            const int gbit = head + MAPF_OBS_BYTES_PER_AGENT * a;
            const int o = gbit & 31;
            uint32_t *S = s_bits + (gbit >> 5);
            uint32_t x0 = 0;
            if (valid[k]) {
                const int x = px[k], y = py[k];
                // window rows x-4..x+4 are padded rows x..x+8; columns y-4..y+4 are padded bits y..y+8: all inside
                // navi tile (x >> 3, y >> 3), rows (x & 7) .. (x & 7) + 8 of one 128-byte line
                const uint2 *nb = reinterpret_cast<const uint2 *>(p.navi + ((size_t)e * N + a) * d.navi_agent_stride) +
                                  ((size_t)((x >> 3) * d.NB + (y >> 3)) << 4) + (x & 7);
                uint2 wr[9];
                if (p.flags & MAPF_STEPF_DIAG_NO_NAVI) {
#pragma unroll
                    for (int u = 0; u < 9; ++u) wr[u] = make_uint2(x + u, y);
                } else if (navi_keep) {
#pragma unroll
                    for (int u = 0; u < 9; ++u) wr[u] = ldg_policy(nb + u, pol_keep);
                } else {
#pragma unroll
                    for (int u = 0; u < 9; ++u) wr[u] = __ldg(nb + u);
                }
                const int sh = y & 7;
                const uint32_t *ag_row = s_agent + x * RWS, *ob_row = s_obst + x * RWS;
                uint32_t prev = 0;
                auto val = [&](auto fc) -> uint32_t {
                    constexpr int f = decltype(fc)::value, c = f / 9, u = f % 9;
                    if constexpr (c == 0) {
                        uint32_t v = window9(ag_row + u * RWS, y);
                        if constexpr (u == 4) v &= ~0x10u;  // own centre cleared, environment.py:461
                        return v;
                    } else if constexpr (c == 1) {
                        return window9(ob_row + u * RWS, y);
                    } else {
                        // direction dd = c - 2: 16 column bits at bit 16 dd of the tile row
                        constexpr int dd = c - 2;
                        const uint32_t half = dd < 2 ? wr[u].x : wr[u].y;
                        return (half >> (sh + 16 * (dd & 1))) & 0x1ffu;
                    }
                };
                auto emit = [&](auto mc, uint32_t w) {
                    constexpr int m = decltype(mc)::value;
                    if constexpr (m == 0) x0 = w << o;
                    else S[m] = __funnelshift_l(prev, w, o);
                    prev = w;
                };
                if constexpr (TRACE) if (k == 0) trace_stamp(p, e, 9, lane);
                FieldWalk<0>::run(0ull, val, emit);
                if constexpr (TRACE) if (k == 0) trace_stamp(p, e, 10, lane);
                if (((o + 485) >> 5) == 16) S[16] = __funnelshift_l(prev, 0u, o);
            }
            __syncwarp();
            // first word: shared with the previous agent's last word unless this agent starts a word
            if (valid[k]) {
                if (o == 0 || a == 0) S[0] = x0;
                else S[0] |= x0;
            }
            __syncwarp();
        }
        }  // DO_OBS

#pragma unroll
        for (int k = 0; k < K; ++k) {
            out.px[k] = px[k];
            out.py[k] = py[k];
            out.valid[k] = valid[k];
            if constexpr (RESIDENT && DO_STEP) {
                out.gx[k] = gx[k];
                out.gy[k] = gy[k];
            }
        }
        if constexpr (RESIDENT && DO_STEP) out.step = step_now + 1;
    }
}

template <int RW, int K>
__device__ __forceinline__ void clear_agent_bits(uint32_t *s_agent, const EnvRegs<K> &r)
{
    constexpr int RWS = RW + 1;
#pragma unroll
    for (int k = 0; k < K; ++k)
        if (r.valid[k]) s_agent[(r.px[k] + 4) * RWS + ((r.py[k] + 4) >> 5)] = 0;
    __syncwarp();
}

// Expand the env's bit stream (1 bit -> 1 bool byte) and write its N*486-byte observation block: 16 bytes per lane per
// store, fully coalesced streaming stores; `head` = bytes between the 16-byte boundary below obs_env and obs_env.
__device__ __forceinline__ void expand_store_block(const StepParams &p, uint8_t *obs_env, const int head, const size_t env_bytes,
                                                   const uint32_t *s_bits, const int lane, const bool obs_policy,
                                                   const uint64_t pol_stream)
{
    const int total = head + (int)env_bytes;
    const int c_lo = (head + 15) >> 4, c_hi = total >> 4;  // chunks [c_lo, c_hi) are whole
    uint8_t *obase = obs_env - head;                       // 16-byte aligned
    const uint16_t *S16 = reinterpret_cast<const uint16_t *>(s_bits);
#pragma unroll 4
    for (int c = c_lo + lane; c < c_hi; c += 32) {
        const uint32_t s = S16[c];
        uint4 v;
        v.x = expand4(s & 0xfu);
        v.y = expand4((s >> 4) & 0xfu);
        v.z = expand4((s >> 8) & 0xfu);
        v.w = expand4(s >> 12);
        uint4 *dst = reinterpret_cast<uint4 *>(obase + (c << 4));
        if (p.flags & MAPF_STEPF_DIAG_NO_STORE) {
            if (v.x == 0xdeadbeefu) __stcs(dst, v);  // never true: keeps the expansion alive
        } else if (obs_policy) stg_policy(dst, v, pol_stream);
        else __stcs(dst, v);
    }
    // ragged first / last chunk of an unaligned observation block
    if ((head != 0 && lane == 0) || ((total & 15) != 0 && lane == 1)) {
        const int c = lane == 0 ? 0 : c_hi;
        const uint32_t s = S16[c];
        for (int b = 0; b < 16; ++b) {
            const int g = (c << 4) + b;
            if (g >= head && g < total) obase[g] = (uint8_t)((s >> b) & 1u);
        }
    }
}

// ---- K1+K2, single-role form: every warp steps an env, then expands and stores its own observation block.
// Kept for observe(), unaligned observation bases and agent counts that are not a multiple of 8.
template <int RW, int K, bool DO_STEP, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
step_observe_kernel(const StepParams p)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const EnvDims &d = p.d;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int N = d.N;

    uint32_t *s_obst = smem + (size_t)warp * p.warp_smem_words;
    uint32_t *s_agent = s_obst + p.obst_words;
    uint32_t *s_bits = s_agent + p.obst_words;
    uint16_t *s_tgt = reinterpret_cast<uint16_t *>(s_bits + p.bits_words);
    uint16_t *s_cell = s_tgt + 32 * K;

    const uint64_t pol_keep = l2_policy_evict_last();
    const uint64_t pol_stream = l2_policy_evict_first();
    const bool obs_policy = p.flags & MAPF_STEPF_OBS_POLICY;

    // the agent bitmap must start all-zero; afterwards each env clears the bits it set
    for (int w = lane; w < p.obst_words; w += 32) s_agent[w] = 0;
    __syncwarp();

    for (int e = p.env_begin + blockIdx.x * WARPS + warp; e < p.env_end; e += gridDim.x * WARPS) {
        const size_t env_bytes = (size_t)N * MAPF_OBS_BYTES_PER_AGENT;
        uint8_t *obs_env = p.obs + (size_t)(p.obs_rows ? p.obs_rows[e] : (int64_t)e) * env_bytes;
        const int head = (int)(reinterpret_cast<uintptr_t>(obs_env) & 15);  // bytes before the 16-B boundary
        EnvRegs<K> r;
        env_step_gather<RW, K, DO_STEP>(p, e, lane, s_obst, s_agent, s_bits, s_tgt, s_cell, head, pol_keep, r);

        expand_store_block(p, obs_env, head, env_bytes, s_bits, lane, obs_policy, pol_stream);
        __syncwarp();
        clear_agent_bits<RW, K>(s_agent, r);  // the agent bits this env set
    }
}

// ---- K1+K2 for a scripted rollout, persistent form: a warp takes an environment through ALL T steps before it moves to its
// next one.  Nothing is launched between steps, the warps drift out of phase on their own, and what an environment re-reads
// every step (its agents' heuristic tile lines -- an agent changes tile once in ~12 steps --, goals, the obstacle bitmap) is
// re-read by the same SM a few microseconds later and comes from L1 / L2 instead of DRAM.
struct RolloutArgs {
    int T, action_slots, obs_slots, out_slots;
    const uint8_t *actions;  // [action_slots, B, N]
    uint8_t *obs;            // [obs_slots, B, N, 6, 9, 9]
    float *rewards;          // [out_slots, B, N]
    uint8_t *done;           // [out_slots, B]
    int32_t *steps_out;      // [out_slots, B] or NULL
};

template <int RW, int K, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
step_rollout_kernel(const StepParams p0, const RolloutArgs r)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const EnvDims &d = p0.d;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int N = d.N;
    uint32_t *s_obst = smem + (size_t)warp * p0.warp_smem_words;
    uint32_t *s_agent = s_obst + p0.obst_words;
    uint32_t *s_bits = s_agent + p0.obst_words;
    uint16_t *s_tgt = reinterpret_cast<uint16_t *>(s_bits + p0.bits_words);
    uint16_t *s_cell = s_tgt + 32 * K;
    const uint64_t pol_keep = l2_policy_evict_last();
    const uint64_t pol_stream = l2_policy_evict_first();
    const bool obs_policy = p0.flags & MAPF_STEPF_OBS_POLICY;
    for (int w = lane; w < p0.obst_words; w += 32) s_agent[w] = 0;
    __syncwarp();
    const size_t BN = (size_t)d.B * N;
    const size_t env_bytes = (size_t)N * MAPF_OBS_BYTES_PER_AGENT;
    for (int e = p0.env_begin + blockIdx.x * WARPS + warp; e < p0.env_end; e += gridDim.x * WARPS) {
        int sa = 0, so = 0, sr = 0;  // t % slots without a division per step
        EnvRegs<K> regs;
        for (int t = 0; t < r.T; ++t) {
            StepParams p = p0;
            p.actions = r.actions + (size_t)sa * BN;
            p.obs = r.obs + (size_t)so * BN * MAPF_OBS_BYTES_PER_AGENT;
            p.rewards = r.rewards + (size_t)sr * BN;
            p.done = r.done + (size_t)sr * d.B;
            p.steps_out = r.steps_out ? r.steps_out + (size_t)sr * d.B : nullptr;
            uint8_t *obs_env = p.obs + (size_t)e * env_bytes;
            const int head = (int)(reinterpret_cast<uintptr_t>(obs_env) & 15);
            env_step_gather<RW, K, true, false, true, true>(p, e, lane, s_obst, s_agent, s_bits, s_tgt, s_cell, head, pol_keep, regs,
                                                            t > 0);
            expand_store_block(p, obs_env, head, env_bytes, s_bits, lane, obs_policy, pol_stream);
            __syncwarp();
            clear_agent_bits<RW, K>(s_agent, regs);
            if (++sa == r.action_slots) sa = 0;
            if (++so == r.obs_slots) so = 0;
            if (++sr == r.out_slots) sr = 0;
        }
    }
}

// ---- K1 alone: Environment.step without the observation.  mapf_env_step_host launches it ahead of the observe kernel
// so that the device-to-host copies of rewards / done / steps run WHILE the observation is being written (their
// 1 MB of PCIe traffic otherwise follows, or stretches, the fused kernel).
template <int RW, int K>
__global__ void __launch_bounds__(128)
step_only_kernel(const StepParams p)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t *s_obst = smem + (size_t)warp * p.warp_smem_words;
    uint32_t *s_agent = s_obst + p.obst_words;
    uint32_t *s_bits = s_agent + p.obst_words;
    uint16_t *s_tgt = reinterpret_cast<uint16_t *>(s_bits + p.bits_words);
    uint16_t *s_cell = s_tgt + 32 * K;
    const int e = blockIdx.x * 4 + warp;
    // The CTA's four action rows are contiguous (4 N bytes): fetched as 16-byte words, they are a few 128-byte requests per
    // CTA instead of one 32-byte request per warp -- mapf_env_step_host points `actions` at page-locked HOST memory and the
    // kernel's time is the PCIe round trips of this load.
    __shared__ __align__(16) uint8_t s_actions[4 * MAPF_MAX_AGENTS];
    const int N = p.d.N;
    const uint8_t *rows = p.actions + (size_t)blockIdx.x * 4 * N;
    const int nbytes = min(4, p.d.B - (int)blockIdx.x * 4) * N;
    const bool staged = ((reinterpret_cast<uintptr_t>(rows) | (uintptr_t)nbytes) & 15) == 0 && !(p.flags & MAPF_STEPF_DIAG_NO_STAGE);
    if (staged) {
        for (int w = threadIdx.x; w < (nbytes >> 4); w += blockDim.x)
            reinterpret_cast<uint4 *>(s_actions)[w] = __ldg(reinterpret_cast<const uint4 *>(rows) + w);
    }
    __syncthreads();
    if (e >= p.d.B) return;
    EnvRegs<K> r;
    env_step_gather<RW, K, true, false, false>(p, e, lane, s_obst, s_agent, s_bits, s_tgt, s_cell, 0, 0ull, r, false,
                                               staged ? s_actions + warp * N : nullptr);
}

// ---- K1+K2, split form (the hot path) -----------------------------------------------------------------
// The single-role kernel keeps every warp in lockstep: all of them compute (DRAM idle), then all of them
// store (issue slots idle), because equal sharing of the DRAM bandwidth keeps their phases aligned.  Here the
// two halves are different warps of one CTA, decoupled by a shared-memory ring:
//   * P producer warps run env_step_gather; each owns two slots and leaves the env's 486*N-BIT stream
//     (1944 B at N = 32) in one of them, then arrives on the slot's `full` mbarrier;
//   * C consumer warps visit the slots in the producers' (static) order; for each they expand the stream to
//     bool bytes, 16 B per lane, and write the env's block together (adjacent 512-byte pieces, streaming
//     stores), then arrive on the slot's `empty` mbarrier.
// Producers never wait on DRAM (only on a free slot), so stepping env i+1 overlaps the stores of env i, and a
// CTA emits one contiguous 15.5-KB block at a time instead of P interleaved ones.
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MAPF_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MAPF_DONE;\n"
        "bra MAPF_WAIT;\n"
        "MAPF_DONE:\n"
        "}" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

template <int RW, int K, int P, int C>
__global__ void __launch_bounds__((P + C) * 32, K == 1 ? 1536 / ((P + C) * 32) : 1)
step_split_kernel(const StepParams p)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const EnvDims &d = p.d;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    // barriers: full[w][s] at bars[2 (2 w + s)], empty[w][s] right behind it
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint32_t *work = smem + ((4 * P * 2 + 3) & ~3);
    if (threadIdx.x < 2 * P) {
        mbar_init(bars + 2 * threadIdx.x, 1);      // full: the producer's lane 0
        mbar_init(bars + 2 * threadIdx.x + 1, C);  // empty: lane 0 of every consumer warp
    }
    __syncthreads();
    const int slot_words = p.bits_words;
    const int estride = gridDim.x * P;

    if (warp < P) {
        // ---------------- producer ----------------
        uint32_t *s_obst = work + (size_t)warp * p.warp_smem_words;
        uint32_t *s_agent = s_obst + p.obst_words;
        uint16_t *s_tgt = reinterpret_cast<uint16_t *>(s_agent + p.obst_words);
        uint16_t *s_cell = s_tgt + 32 * K;
        uint32_t *slots = reinterpret_cast<uint32_t *>(s_cell + 32 * K);
        const uint64_t pol_keep = l2_policy_evict_last();
        for (int w = lane; w < p.obst_words; w += 32) s_agent[w] = 0;
        __syncwarp();
        int it = 0;
        for (int e = blockIdx.x * P + warp; e < d.B; e += estride, ++it) {
            const int s = it & 1;
            trace_stamp(p, e, 0, lane);
            mbar_wait(bars + 2 * (2 * warp + s) + 1, ((it >> 1) & 1) ^ 1);  // slot free (first use: passes)
            trace_stamp(p, e, 1, lane);
            EnvRegs<K> r;
            env_step_gather<RW, K, true, true>(p, e, lane, s_obst, s_agent, slots + s * slot_words, s_tgt, s_cell, 0, pol_keep, r);
            if (lane == 0) mbar_arrive(bars + 2 * (2 * warp + s));  // every lane's writes precede the last __syncwarp
            trace_stamp(p, e, 3, lane);
            clear_agent_bits<RW, K>(s_agent, r);
        }
    } else {
        // ---------------- consumer ----------------
        const int cw = warp - P;
        const int cpe = p.chunks_per_env;
        const size_t env_bytes = (size_t)d.N * MAPF_OBS_BYTES_PER_AGENT;
        for (int r = 0, e0 = blockIdx.x * P; e0 < d.B; ++r, e0 += estride) {
            const int s = r & 1;
#pragma unroll 1
            for (int w = 0; w < P; ++w) {
                const int e = e0 + w;
                if (e >= d.B) break;
                const int64_t row = p.obs_rows ? __ldg(p.obs_rows + e) : (int64_t)e;
                uint4 *dst = reinterpret_cast<uint4 *>(p.obs + (size_t)row * env_bytes);
                const uint16_t *S16 = reinterpret_cast<const uint16_t *>(work + (size_t)w * p.warp_smem_words + 2 * p.obst_words +
                                                                         32 * K + s * slot_words);
                if (cw == 0) trace_stamp(p, e, 4, lane);
                mbar_wait(bars + 2 * (2 * w + s), (r >> 1) & 1);
                if (cw == 0) trace_stamp(p, e, 5, lane);
#pragma unroll 4
                for (int c = cw * 32 + lane; c < cpe; c += C * 32) {
                    const uint32_t x = S16[c];
                    uint4 v;
                    v.x = expand4(x & 0xfu);
                    v.y = expand4((x >> 4) & 0xfu);
                    v.z = expand4((x >> 8) & 0xfu);
                    v.w = expand4(x >> 12);
                    __stcs(dst + c, v);
                }
                __syncwarp();
                if (cw == 0) trace_stamp(p, e, 6, lane);
                if (lane == 0) mbar_arrive(bars + 2 * (2 * w + s) + 1);
            }
        }
    }
}

// ---- launch plumbing ---------------------------------------------------------------------------
struct StepTuning {
    int variant;       // CTA shape / register cap of the (RW = 2, K = 1) instantiation, see launch_step_rwk
    int flags;
    int ctas_per_sm;   // > 0: persistent grid of that many CTAs per SM, each warp strides over environments
    unsigned long long *trace = nullptr;
};

StepTuning &tuning()
{
    static StepTuning t = [] {
        StepTuning r{1, MAPF_STEPF_NAVI_KEEP, 0};
        if (const char *s = std::getenv("MAPF_STEP_VARIANT")) r.variant = std::atoi(s);
        if (const char *s = std::getenv("MAPF_STEP_FLAGS")) r.flags = std::atoi(s);
        if (const char *s = std::getenv("MAPF_STEP_CTAS_PER_SM")) r.ctas_per_sm = std::atoi(s);
        return r;
    }();
    return t;
}

template <int RW, int K, bool DO_STEP, int WARPS, int MINB>
int launch_step_cfg(const mapf_env *env, StepParams &p, cudaStream_t st)
{
    auto kern = step_observe_kernel<RW, K, DO_STEP, WARPS, MINB>;
    const size_t smem = (size_t)p.warp_smem_words * 4 * WARPS;
    if (smem > 227 * 1024) {
        mapf_set_error("map too large for the step kernel's shared memory");
        return MAPF_EINVAL;
    }
    if (smem > 48 * 1024)  // per-device attribute; cheap enough to set on every large-smem launch
        MAPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = (p.env_end - p.env_begin + WARPS - 1) / WARPS;
    if (tuning().ctas_per_sm > 0) {
        const int cap = env->num_sms * tuning().ctas_per_sm;
        if (grid > cap) grid = cap;
    }
    kern<<<grid, WARPS * 32, smem, st>>>(p);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

// Split form: one CTA serves P environments per round; the grid is persistent (a few CTAs per SM) so that
// the slot ring stays warm, but nothing requires the CTAs to be co-resident.
template <int RW, int K, int P, int C>
int launch_split_cfg(mapf_env *env, StepParams p, cudaStream_t st)
{
    auto kern = step_split_kernel<RW, K, P, C>;
    const EnvDims &d = env->d;
    // per producer: obstacle + agent bitmaps, s_tgt + s_cell, two slots (bit stream / occupancy grid)
    const int stream_words = (((d.N * MAPF_OBS_BYTES_PER_AGENT + 31) >> 5) + 2 + 3) & ~3;
    const int occ_words = (((d.L * d.L + 3) >> 2) + 3) & ~3;
    p.bits_words = stream_words > occ_words ? stream_words : occ_words;
    p.warp_smem_words = 2 * p.obst_words + 32 * d.K + 2 * p.bits_words;
    const size_t smem = ((size_t)((4 * P * 2 + 3) & ~3) + (size_t)p.warp_smem_words * P) * 4;
    if (smem > 227 * 1024) return MAPF_EINVAL;  // caller falls back to the single-role kernel
    const int key = 1 + P * 64 + C;
    if (env->split_key != key) {
        int per_sm = 0;
        MAPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MAPF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (P + C) * 32, smem));
        env->split_per_sm = per_sm < 1 ? 1 : per_sm;
        env->split_key = key;
    }
    int ctas = env->split_per_sm;
    if (tuning().ctas_per_sm > 0 && ctas > tuning().ctas_per_sm) ctas = tuning().ctas_per_sm;
    int grid = env->num_sms * ctas;
    const int want = (d.B + P - 1) / P;
    if (grid > want) grid = want;
    kern<<<grid, (P + C) * 32, smem, st>>>(p);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

template <int RW, int K>
int launch_split_rwk(mapf_env *env, const StepParams &p, cudaStream_t st)
{
    if constexpr (RW == 2 && K == 1) {
        switch (tuning().variant) {
            case 11: return launch_split_cfg<RW, K, 4, 4>(env, p, st);
            case 12: return launch_split_cfg<RW, K, 6, 2>(env, p, st);
            case 13: return launch_split_cfg<RW, K, 7, 5>(env, p, st);
            case 14: return launch_split_cfg<RW, K, 3, 1>(env, p, st);
            case 15: return launch_split_cfg<RW, K, 4, 2>(env, p, st);
            case 16: return launch_split_cfg<RW, K, 2, 2>(env, p, st);
            case 17: return launch_split_cfg<RW, K, 3, 3>(env, p, st);
            case 18: return launch_split_cfg<RW, K, 8, 4>(env, p, st);
            case 19: return launch_split_cfg<RW, K, 10, 6>(env, p, st);
            default: break;
        }
    }
    return launch_split_cfg<RW, K, 5, 3>(env, p, st);
}

int launch_split(mapf_env *env, const StepParams &p, cudaStream_t st)
{
    const int key = env->d.RW * 10 + env->d.K;
    switch (key) {
        case 11: return launch_split_rwk<1, 1>(env, p, st);
        case 12: return launch_split_rwk<1, 2>(env, p, st);
        case 21: return launch_split_rwk<2, 1>(env, p, st);
        case 22: return launch_split_rwk<2, 2>(env, p, st);
        case 23: return launch_split_rwk<2, 3>(env, p, st);
        case 24: return launch_split_rwk<2, 4>(env, p, st);
        case 31: return launch_split_rwk<3, 1>(env, p, st);
        case 32: return launch_split_rwk<3, 2>(env, p, st);
        case 33: return launch_split_rwk<3, 3>(env, p, st);
        case 34: return launch_split_rwk<3, 4>(env, p, st);
        case 41: return launch_split_rwk<4, 1>(env, p, st);
        case 42: return launch_split_rwk<4, 2>(env, p, st);
        case 43: return launch_split_rwk<4, 3>(env, p, st);
        case 44: return launch_split_rwk<4, 4>(env, p, st);
    }
    return MAPF_EINVAL;  // caller falls back to the single-role kernel
}

template <int RW, int K, bool DO_STEP>
int launch_step_rwk(const mapf_env *env, StepParams &p, cudaStream_t st)
{
    if constexpr (RW == 2 && K == 1) {
        // CTA shape / register budget of the hot geometry (40x40, <= 32 agents), measured at 8192 x 32 (profiles/
        // r1_rollout_cta_shapes.log; us per step as one whole-batch launch / as 8 rollout chains):
        //   0: 8 warps, 48 regs, 40 warps/SM   33.3 / 24.9        7: 8 warps, 64 regs, 32 warps/SM   34.8 / 24.2
        //   4: 4 warps, 40 regs, 48 warps/SM   34.3 / 25.7        8: 4 warps, 64 regs, 32 warps/SM   34.6 / 24.4
        //   2: 4 warps, 32 regs (spills)       38.1 / 32.9        3: 2 warps, 32 regs                 38.1 / 32.2
        // 1 (default) picks 0 for a launch over the whole batch and 7 for a sub-batch launch of mapf_env_rollout, whose
        // concurrent chains supply the parallelism that the extra resident warps otherwise would.
        int v = tuning().variant;
        if (v == 1) v = (p.env_end - p.env_begin < env->d.B) ? 7 : 0;
        switch (v) {
            case 0: return launch_step_cfg<RW, K, DO_STEP, 8, 5>(env, p, st);
            case 2: return launch_step_cfg<RW, K, DO_STEP, 4, 16>(env, p, st);
            case 3: return launch_step_cfg<RW, K, DO_STEP, 2, 32>(env, p, st);
            case 7: return launch_step_cfg<RW, K, DO_STEP, 8, 4>(env, p, st);
            case 8: return launch_step_cfg<RW, K, DO_STEP, 4, 8>(env, p, st);
            default: break;  // 4: the general shape below
        }
    }
    if constexpr (K == 2 && DO_STEP && RW <= 2) {  // (at 80x80 shared memory holds 20 warps per SM and the cap costs 1.4 %)
        // two agents per lane (33..64 agents): the general shape below runs at 80 registers / 24 warps per SM.  Capped at
        // 64 registers (32 warps per SM, no spills) a whole-batch launch of 8192 x 64 agents takes 69.2 instead of 72.7 us;
        // the rollout's sub-batch launches are no faster (52.8 vs 52.4 us per step; profiles/r1_rollout_cta_shapes.log)
        int v = tuning().variant;
        if (v == 1) v = (p.env_end - p.env_begin < env->d.B) ? 4 : 9;
        if (v == 9) return launch_step_cfg<RW, K, DO_STEP, 4, 8>(env, p, st);
    }
    return launch_step_cfg<RW, K, DO_STEP, 4, (K == 1 ? 12 : 1)>(env, p, st);
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
    // stream words: 15 head bits max + N*486 bits, +2 words of slack for the u16 tail read; the same
    // buffer holds the L*L-byte occupancy grid of the step phase
    const int stream_words = ((15 + d.N * MAPF_OBS_BYTES_PER_AGENT + 31) >> 5) + 2;
    const int occ_words = (d.L * d.L + 3) >> 2;
    p.bits_words = stream_words > occ_words ? stream_words : occ_words;
    const int words = 2 * p.obst_words + p.bits_words + (32 * d.K) /* s_tgt + s_cell, u16 each */;
    p.warp_smem_words = (words + 3) & ~3;
    p.flags = tuning().flags;
    p.trace = tuning().trace;
    p.chunks_per_env = d.N * MAPF_OBS_BYTES_PER_AGENT / 16;
    p.env_begin = 0;
    p.env_end = d.B;
    return p;
}

}  // namespace

static int g_tuning_generation = 0;
int mapf_step_tuning_generation() { return g_tuning_generation; }  // captured launches are stale once this moves

void mapf_set_step_tuning(int variant, int flags, int ctas_per_sm)
{
    ++g_tuning_generation;
    StepTuning &t = tuning();
    if (variant >= 0) t.variant = variant;
    if (flags >= 0) t.flags = flags;
    if (ctas_per_sm >= 0) t.ctas_per_sm = ctas_per_sm;
}

void mapf_set_step_trace(unsigned long long *d_trace)
{
    ++g_tuning_generation;
    tuning().trace = d_trace;
}

int mapf_launch_step(mapf_env *env, const uint8_t *d_actions, uint8_t *d_obs, const int64_t *d_obs_rows, float *d_rewards,
                     uint8_t *d_done, int32_t *d_steps, cudaStream_t st)
{
    StepParams p = make_params(env);
    p.actions = d_actions;
    p.obs = d_obs;
    p.obs_rows = d_obs_rows;
    p.rewards = d_rewards;
    p.done = d_done;
    p.steps_out = d_steps;
    // the split (producer / consumer) kernel is the hot path; variant 0..3 select the single-role kernel
    const bool aligned = (reinterpret_cast<uintptr_t>(d_obs) & 15) == 0;
    const bool diag = p.flags & (MAPF_STEPF_DIAG_NO_NAVI | MAPF_STEPF_DIAG_NO_STORE | MAPF_STEPF_OBS_POLICY);
    if (env->d.N % 8 == 0 && aligned && !diag && tuning().variant >= 10) {
        const int rc = launch_split(env, p, st);
        if (rc != MAPF_EINVAL) return rc;
    }
    return launch_step<true>(env, p, st);
}

template <int RW, int K>
static int launch_step_only_cfg(mapf_env *env, const StepParams &p, cudaStream_t st)
{
    auto kern = step_only_kernel<RW, K>;
    const size_t smem = (size_t)p.warp_smem_words * 4 * 4;
    if (smem > 227 * 1024) {
        mapf_set_error("map too large for the step kernel's shared memory");
        return MAPF_EINVAL;
    }
    if (smem > 48 * 1024) MAPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(env->d.B + 3) / 4, 128, smem, st>>>(p);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

// One step of the sub-batch [e0, e1) only (mapf_env_rollout: independent chains of launches over disjoint env ranges);
// all pointers are those of the whole batch.
int mapf_launch_step_range(mapf_env *env, int e0, int e1, const uint8_t *d_actions, uint8_t *d_obs, float *d_rewards,
                           uint8_t *d_done, int32_t *d_steps, cudaStream_t st)
{
    StepParams p = make_params(env);
    p.actions = d_actions;
    p.obs = d_obs;
    p.rewards = d_rewards;
    p.done = d_done;
    p.steps_out = d_steps;
    p.env_begin = e0;
    p.env_end = e1;
    return launch_step<true>(env, p, st);
}

// Persistent rollout (see step_rollout_kernel): only the hot geometry (maps up to 56 cells, up to 32 agents) is instantiated;
// MAPF_EINVAL tells the caller to use chains of launches instead.
template <int WARPS>
static int launch_rollout_persistent_cfg(mapf_env *env, const StepParams &p, const RolloutArgs &r, int epw, cudaStream_t st)
{
    auto kern = step_rollout_kernel<2, 1, WARPS, 32 / WARPS>;   // 64 registers
    const size_t smem = (size_t)p.warp_smem_words * 4 * WARPS;
    if (smem > 48 * 1024) MAPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int warps = (env->d.B + epw - 1) / epw;
    const int grid = (warps + WARPS - 1) / WARPS;
    kern<<<grid, WARPS * 32, smem, st>>>(p, r);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

int mapf_launch_rollout_persistent(mapf_env *env, int T, const uint8_t *d_actions, int action_slots, uint8_t *d_obs, int obs_slots,
                                   float *d_rewards, uint8_t *d_done, int32_t *d_steps, int out_slots, int envs_per_warp, int cta_warps,
                                   cudaStream_t st)
{
    if (env->d.RW != 2 || env->d.K != 1) return MAPF_EINVAL;
    StepParams p = make_params(env);
    RolloutArgs r{T, action_slots, obs_slots, out_slots, d_actions, d_obs, d_rewards, d_done, d_steps};
    // Every warp of the (fully resident) grid takes the same number of environments through their T steps, one after another.
    // Measured at 8192 x 32 agents (profiles/r1_rollout_persistent.log; us per step; state carried in registers):
    //   2-warp CTAs: 2 environments per warp 23.0, 3: 21.7, 4: 21.3        4-warp CTAs: 2: 22.7, 3: 21.6, 4: 22.8
    // (before the state was carried: 4-warp CTAs 24.5 / 22.3 / 23.7 / 24.1 / 28.5 / 33.7 at 2 / 3 / 4 / 5 / 6 / 8; 1-warp CTAs
    // 25.9 / 28.0 at 3 / 4).  About 14-18 resident warps per SM: few enough for their environments' heuristic lines to stay in
    // L1, enough to keep the DRAM write stream busy.
    int epw = envs_per_warp;
    if (epw <= 0) {
        const int capacity = env->num_sms * 16;
        epw = (env->d.B + capacity - 1) / capacity;
    }
    switch (cta_warps) {
        case 1: return launch_rollout_persistent_cfg<1>(env, p, r, epw, st);
        case 4: return launch_rollout_persistent_cfg<4>(env, p, r, epw, st);
        default: return launch_rollout_persistent_cfg<2>(env, p, r, epw, st);
    }
}

int mapf_launch_step_only(mapf_env *env, const uint8_t *d_actions, float *d_rewards, uint8_t *d_done, int32_t *d_steps,
                          cudaStream_t st)
{
    StepParams p = make_params(env);
    p.actions = d_actions;
    p.rewards = d_rewards;
    p.done = d_done;
    p.steps_out = d_steps;
    switch (env->d.RW * 10 + env->d.K) {
        case 11: return launch_step_only_cfg<1, 1>(env, p, st);
        case 12: return launch_step_only_cfg<1, 2>(env, p, st);
        case 13: return launch_step_only_cfg<1, 3>(env, p, st);
        case 14: return launch_step_only_cfg<1, 4>(env, p, st);
        case 21: return launch_step_only_cfg<2, 1>(env, p, st);
        case 22: return launch_step_only_cfg<2, 2>(env, p, st);
        case 23: return launch_step_only_cfg<2, 3>(env, p, st);
        case 24: return launch_step_only_cfg<2, 4>(env, p, st);
        case 31: return launch_step_only_cfg<3, 1>(env, p, st);
        case 32: return launch_step_only_cfg<3, 2>(env, p, st);
        case 33: return launch_step_only_cfg<3, 3>(env, p, st);
        case 34: return launch_step_only_cfg<3, 4>(env, p, st);
        case 41: return launch_step_only_cfg<4, 1>(env, p, st);
        case 42: return launch_step_only_cfg<4, 2>(env, p, st);
        case 43: return launch_step_only_cfg<4, 3>(env, p, st);
        case 44: return launch_step_only_cfg<4, 4>(env, p, st);
    }
    mapf_set_error("unsupported geometry");
    return MAPF_EINVAL;
}

int mapf_launch_observe(mapf_env *env, uint8_t *d_obs, const int64_t *d_obs_rows, uint8_t *d_pos, cudaStream_t st)
{
    StepParams p = make_params(env);
    p.obs = d_obs;
    p.obs_rows = d_obs_rows;
    p.pos_out = d_pos;
    return launch_step<false>(env, p, st);
}
