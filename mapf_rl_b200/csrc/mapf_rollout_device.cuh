// mapf_rollout_device.cuh — K1+K2 for a scripted rollout, persistent form: ONE launch takes every environment of the batch
// through T steps of Environment.step + observe (the loop `for t: env.step(a[t])` of test.py:120-130 / worker.py:383-395
// when the actions do not depend on the observations), optionally with the actor's episode handling inside the launch
// (worker.py:390,422-428: done or steps >= max_steps -> reset -> first observation of the new episode).
//
// Work distribution.  The rollout is cut into work items (environment e, chunk k of C consecutive steps), handed out
// TIME-MAJOR from one global counter: item i is chunk i / nenv of environment i % nenv.  A warp that takes an item waits
// until the previous chunk of that environment is finished (a per-environment progress word, release / acquire; that
// chunk was handed out nenv items earlier, so the wait is almost always over before it starts), loads the environment's
// state once (positions, goals, step counter with ld.global.cg, the obstacle bitmap with cp.async into shared memory),
// runs the chunk's steps with the state resident in registers / shared memory -- a step starts with ONE global load, the
// actions -- and publishes the chunk.  Warps therefore drift out of phase on their own (some resolve conflicts while
// others store), nothing is launched between steps, a slow item (an environment that is re-generated) delays nobody else,
// and the tail of the launch is one chunk long instead of one environment long.  Everything the scheduler touches is
// reset by the last warp that leaves, so the launch can be captured into a CUDA graph and replayed.
//
// Heuristic rows.  Each agent's current 16x16 navi tile (one 128-byte line, mapf_common.cuh) is cached in a per-agent slot
// of shared memory and refilled with cp.async only when the agent crosses into another tile (once in ~12 steps): the
// nine rows of the 9x9 window are shared-memory reads instead of 288 L1 wavefronts per environment-step, and the only
// global reads left per step are the actions and the occasional tile line.
//
// Stores.  Either the direct form of the step kernel (bits -> bool bytes in registers -> 16-byte streaming stores) or,
// store_mode 1, the Blackwell bulk form: the warp expands the block into a 16-byte-aligned staging buffer in shared
// memory and ONE elected lane hands the whole N*486-byte block to the TMA engine (cp.async.bulk.global.shared::cta);
// the warp goes on to the next step's conflict resolution and only waits (cp.async.bulk.wait_group.read) right before
// it expands the next block into the same buffer.
//
// HBM-bound (486 B written per agent-step); nothing here is a dense contraction, so no tensor cores.
#pragma once
#include "mapf_bfs_device.cuh"
#include "mapf_reset_device.cuh"
#include "mapf_step_device.cuh"

namespace {

struct RolloutArgs {
    int T, action_slots, obs_slots, out_slots;
    const uint8_t *actions;  // [action_slots, B, N]
    uint8_t *obs;            // [obs_slots, B, N, 6, 9, 9]
    float *rewards;          // [out_slots, B, N] or NULL
    uint8_t *codes;          // [out_slots, B, N] or NULL
    uint8_t *done;           // [out_slots, B]
    int32_t *steps_out;      // [out_slots, B] or NULL
    int chunk, nchunk;       // steps per work item, items per environment
    unsigned long long *work;  // [0] next item, [1] warps that have left
    uint32_t *progress;      // [B] chunks of the environment finished in this launch (0 between launches)
    // layout of a warp's shared memory (32-bit words)
    int warp_words, tiles_off, block_off;
    int store_mode;          // 0 direct stores, 1 bulk (TMA) stores from a shared-memory staging block
    int stagger_ns;          // > 0: warps start up to this many ns apart (the launch starts out of phase)
    // episode handling inside the launch (max_steps == 0: off)
    int max_steps;
    unsigned long long seed, env_offset, stride;
    float density;
    uint32_t *episode;       // [B] episodes generated for the slot so far
    // environments that will hit the step cap inside this launch (rollout_prio_kernel): taken first, all T steps in one item
    const uint32_t *prio;    // [0] count, [1 ..] environment ids
    const uint8_t *prio_flag;  // [B]
    // pre-generated next instances (mapf_launch_pregen; NULL: every episode end re-generates in place)
    uint32_t *pg_epi;        // [B] episode number of the slot's staged instance (0: none), see mapf_env
    int prio_last;           // the listed environments are handed out LAST (their instances are being staged beside this launch)
    const uint32_t *pg_obst;
    const uint8_t *pg_pos, *pg_goal;
    uint8_t *navi_sel;       // [B]
    int pregen_min;          // the pre-generation kernels ran iff the list holds at least this many environments
    // BFS task queue of in-launch re-generations (NULL: the re-generating warp searches for every agent itself)
    uint32_t *tq;            // [0] head (hint), [1] tail, [4 ..] ring of tq_cap entries {slot, next group, groups done, -}
    int tq_cap;
#ifdef MAPF_ENABLE_DIAG
    unsigned long long *trace;  // [1 + 8 * capacity]: [0] records written; per item {item, env, claimed, loaded, regen_ns, regens, done, smid}
    int trace_cap;
#endif
};

#ifdef MAPF_ENABLE_DIAG
__device__ __forceinline__ unsigned long long diag_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

// Which environments re-generate inside a launch of T steps is known up front for the step cap (steps + T > cap): those
// are the long items (a re-generation is ~80 steps of work), so they are listed here and handed out FIRST, each as one item
// over all T steps; the rest of the batch follows as time-major chunks and fills in around them.  Without this the last
// re-generations start late and the launch ends with a tail of a few warps.  One CTA (the scan is B loads).
__global__ void rollout_prio_kernel(const int32_t *__restrict__ steps, int e0, int e1, int T, int cap, uint32_t *__restrict__ prio,
                                    uint8_t *__restrict__ prio_flag, uint32_t *__restrict__ pg_epi, uint32_t *__restrict__ pg_cnt,
                                    unsigned long long *__restrict__ work)
{
    __shared__ unsigned count;
    if (threadIdx.x == 0) {
        count = 0;
        work[2] = work[3] = 0ull;  // work counters of the pre-generation kernels that follow
    }
    __syncthreads();
    for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        const bool will = steps[e] + T > cap;
        prio_flag[e] = will ? 1 : 0;
        pg_epi[e] = 0;   // published by pregen_bfs_kernel once the slot's next instance is staged
        pg_cnt[e] = 0;
        if (will) prio[1 + atomicAdd(&count, 1u)] = (uint32_t)e;
    }
    __syncthreads();
    if (threadIdx.x == 0) prio[0] = count;
}

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// The environment's state into the warp: obstacle bitmap -> shared memory, positions / goals / step counter -> registers.
// ld.global.cg / cp.async.cg: the data may have been written by another SM earlier in this launch (previous chunk, or an
// in-launch re-generation), so nothing here may come from a stale L1 line.
template <int RW, int K>
__device__ __forceinline__ void load_env_state(const StepParams &p, const int e, const int lane, uint32_t *s_obst, EnvRegs<K> &r)
{
    const EnvDims &d = p.d;
    const int N = d.N;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.obst + (size_t)e * d.obst_stride);
        const uint32_t dst = smem_addr(s_obst);
        for (int w = lane; w < (p.obst_words >> 2); w += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * w), "l"(src + w) : "memory");
    }
    bool all_goal = true;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int a = k * 32 + lane;
        r.valid[k] = a < N;
        r.px[k] = r.py[k] = r.gx[k] = r.gy[k] = 0;
        r.tile[k] = -1;
        if (r.valid[k]) {
            const uchar2 pp = __ldcg(reinterpret_cast<const uchar2 *>(p.pos) + (size_t)e * N + a);
            const uchar2 gg = __ldcg(reinterpret_cast<const uchar2 *>(p.goal) + (size_t)e * N + a);
            r.px[k] = pp.x, r.py[k] = pp.y, r.gx[k] = gg.x, r.gy[k] = gg.y;
            all_goal = all_goal && pp.x == gg.x && pp.y == gg.y;
        }
    }
    int st = 0, sel = 0;
    if (lane == 0) {
        st = __ldcg(p.steps + e);
        sel = p.navi_alt ? __ldcg(p.navi_sel + e) : 0;
    }
    r.step = __shfl_sync(MAPF_FULL_MASK, st, 0);
    r.sel = __shfl_sync(MAPF_FULL_MASK, sel, 0);
    r.finished = __all_sync(MAPF_FULL_MASK, all_goal);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
}

// The heuristic maps of agent group `grp` of slot e (two agents, one per half warp, for maps of up to 88 cells a side, else
// one) into the slot's live buffer: get_navi_map (environment.py:195) with the warp-level BFS of the load / reset path.
// Agents per search pass inside the rollout kernel: two (16 lanes x up to 6 rows each) for maps of up to 88 cells a side, else
// one.  (The load / reset kernels take FOUR agents per warp at 40x40 -- 8 % faster there; compiled into re-generation here, its
// 40 live words per lane made the whole out-of-line function spill, generator included: 27 -> 31 us per step at C2.)
template <int RW>
__device__ __forceinline__ int bfs_groups(const EnvDims &d)
{
    return RW <= 3 ? (d.N + 1) >> 1 : d.N;
}

template <int RW>
__device__ __forceinline__ void bfs_group_into(const StepParams &p, const int e, const int grp, uint32_t *navi)
{
    const EnvDims &d = p.d;
    const int lane = threadIdx.x & 31;
    const uint32_t *obst = p.obst;
    const uint8_t *goal = p.goal;
    if constexpr (RW <= 3) {
        // 16 lanes x RPL rows hold the map: RPL = ceil(L / 16)
        const int a = 2 * grp + (lane >> 4);
        auto pair = [&](auto rplc) {
            constexpr int RPL = decltype(rplc)::value;
            bfs_navi_warp<RW, RPL, 2>(d, e, a < d.N ? a : 0, 0, a < d.N, obst, goal, navi, nullptr);
        };
        const int rpl = (d.L + 15) >> 4;
        if constexpr (RW == 1) {
            if (rpl <= 1) pair(std::integral_constant<int, 1>{});
            else pair(std::integral_constant<int, 2>{});
        } else if constexpr (RW == 2) {
            if (rpl <= 2) pair(std::integral_constant<int, 2>{});
            else if (rpl == 3) pair(std::integral_constant<int, 3>{});
            else pair(std::integral_constant<int, 4>{});
        } else {
            if (rpl <= 4) pair(std::integral_constant<int, 4>{});
            else if (rpl == 5) pair(std::integral_constant<int, 5>{});
            else pair(std::integral_constant<int, 6>{});
        }
    } else {
        bfs_navi_warp<RW, 4, 1>(d, e, grp, 0, true, obst, goal, navi, nullptr);  // 32 lanes x 4 rows >= L
    }
}

// Every agent group of slot e, one after the other on this warp: the free rows of the map and the goals are loaded ONCE (a
// search used to start with two dependent L2 round trips -- goal, then rows -- and to read the rows again when it emitted its
// tiles: ~2 us of a ~10-us search).
template <int RW>
__device__ __forceinline__ void bfs_all_groups(const StepParams &p, const int e, uint32_t *navi)
{
    const EnvDims &d = p.d;
    const int lane = threadIdx.x & 31;
    const uint32_t *ob = p.obst + (size_t)e * d.obst_stride;
    // goals of agents lane and lane + 32 (the rollout kernel serves up to 64 agents)
    uchar2 g0 = make_uchar2(0, 0), g1 = make_uchar2(0, 0);
    if (lane < d.N) g0 = __ldcg(reinterpret_cast<const uchar2 *>(p.goal) + (size_t)e * d.N + lane);
    if (lane + 32 < d.N) g1 = __ldcg(reinterpret_cast<const uchar2 *>(p.goal) + (size_t)e * d.N + lane + 32);
    const int gp0 = g0.x | (g0.y << 8), gp1 = g1.x | (g1.y << 8);
    if constexpr (RW <= 3) {
        auto all = [&](auto rplc) {
            constexpr int RPL = decltype(rplc)::value;
            BfsFree<RW, RPL> F;
            bfs_load_free<RW, RPL, 2>(d, ob, F);
            for (int grp = 0; 2 * grp < d.N; ++grp) {
                const int a = 2 * grp + (lane >> 4);
                const bool alive = a < d.N;
                const int v0 = __shfl_sync(MAPF_FULL_MASK, gp0, a & 31), v1 = __shfl_sync(MAPF_FULL_MASK, gp1, a & 31);
                const int gp = a < 32 ? v0 : v1;
                bfs_navi_warp_pre<RW, RPL, 2>(d, e, alive ? a : 0, alive, ob, F, gp & 0xff, gp >> 8, navi);
            }
        };
        const int rpl = (d.L + 15) >> 4;
        if constexpr (RW == 1) {
            if (rpl <= 1) all(std::integral_constant<int, 1>{});
            else all(std::integral_constant<int, 2>{});
        } else if constexpr (RW == 2) {
            if (rpl <= 2) all(std::integral_constant<int, 2>{});
            else if (rpl == 3) all(std::integral_constant<int, 3>{});
            else all(std::integral_constant<int, 4>{});
        } else {
            if (rpl <= 4) all(std::integral_constant<int, 4>{});
            else if (rpl == 5) all(std::integral_constant<int, 5>{});
            else all(std::integral_constant<int, 6>{});
        }
    } else {
        BfsFree<RW, 4> F;
        bfs_load_free<RW, 4, 1>(d, ob, F);
        for (int a = 0; a < d.N; ++a) {
            const int v0 = __shfl_sync(MAPF_FULL_MASK, gp0, a & 31), v1 = __shfl_sync(MAPF_FULL_MASK, gp1, a & 31);
            const int gp = a < 32 ? v0 : v1;
            bfs_navi_warp_pre<RW, 4, 1>(d, e, a, true, ob, F, gp & 0xff, gp >> 8, navi);
        }
    }
}

// the slot's live heuristic-map buffer
__device__ __forceinline__ uint32_t *navi_live_of(const StepParams &p, const int e)
{
    int sel = 0;
    if ((threadIdx.x & 31) == 0 && p.navi_alt) sel = __ldcg(p.navi_sel + e);
    sel = __shfl_sync(MAPF_FULL_MASK, sel, 0);
    return const_cast<uint32_t *>(sel ? p.navi_alt : p.navi);
}

// (out of line: the task path calls it from two places)
template <int RW>
__device__ __noinline__ void bfs_group(const StepParams &p, const int e, const int grp)
{
    bfs_group_into<RW>(p, e, grp, navi_live_of(p, e));
}

// ---- BFS task queue -------------------------------------------------------------------------------------------------
// A re-generation is the generator (one warp, serial over the agents) followed by one BFS per agent group -- independent
// searches, ~2/3 of the work.  The re-generating warp announces them in a ring entry {slot, next group, groups done}; every
// warp looks at the open entries before it claims its next work item and takes a group with ONE fetch-and-add on the
// entry's `next` word (no retry loops: thousands of warps arrive at once; the owner takes four at a time, a helper two), so the
// searches of one slot run on many warps and
// a re-generation takes the generator's time plus about one search instead of all of them in sequence.
// A free entry has next = kTqFree (negative as int32, far from wrapping): a claim succeeds iff 0 <= claimed < groups.  The
// owner publishes slot BEFORE it zeroes next and frees the entry (next first) only after every group is done, so a
// successful claim always reads the slot it belongs to, whatever incarnation of the entry a late warp runs into.
constexpr uint32_t kTqFree = 0x80808080u;  // never used: the cudaMemset(0x80) pattern
constexpr uint32_t kTqDone = 0xC0C0C0C0u;  // its re-generation is complete
constexpr int kTqUnpublished = -2;         // tq_claim: reserved by an owner that has not published it yet (or never used)

__device__ __forceinline__ uint32_t *tq_entry(const RolloutArgs &r, const uint32_t idx) { return r.tq + 4 + 4 * (idx % (uint32_t)r.tq_cap); }

// Claims up to `want` consecutive groups of ring entry `ent` with ONE fetch-and-add and runs them: returns how many it ran
// (0: none left), or kTqUnpublished.  One fence + one add of the count publish the whole batch -- a claim / completion pair
// costs three L2 round trips (~3 us), as much as half a search at 40x40.  `own`: the caller published the entry itself.
template <int RW>
__device__ __forceinline__ int tq_take(const StepParams &p, uint32_t *ent, const int lane, const int want, const bool own, const int own_e)
{
    const int groups = bfs_groups<RW>(p.d);
    int first = -1, slot = own_e;
    if (lane == 0) {
        const int32_t c = (int32_t)atomicAdd(ent + 1, (uint32_t)want);
        if (c >= 0 && c < groups) {
            first = c;
            if (!own) {
                __threadfence();  // the claim before the slot (published before next was zeroed)
                slot = (int)ld_acquire_u32(ent);
            }
        } else if (c < (int32_t)0xA0000000u) {
            first = kTqUnpublished;
        }
    }
    first = __shfl_sync(MAPF_FULL_MASK, first, 0);
    if (first < 0) return first == kTqUnpublished ? kTqUnpublished : 0;
    const int e = __shfl_sync(MAPF_FULL_MASK, slot, 0);
    const int last = min(first + want, groups);
    for (int g = first; g < last; ++g) bfs_group<RW>(p, e, g);
    __syncwarp();
    if (lane == 0) {
        __threadfence();  // the tiles before the count
        atomicAdd(ent + 2, (uint32_t)(last - first));
    }
    return last - first;
}

// A warp between two work items lends a hand: it looks at ONE announced entry, picked pseudo-randomly among those past the
// head hint (thousands of warps come by while a few hundred re-generations are open: walking every entry would cost each of
// them one L2 round trip per entry, and all of them would crowd the same one), takes up to two of its groups and goes back
// to its own items.  An entry found exhausted right at the head moves the hint forward.
template <int RW>
__device__ __forceinline__ void tq_help(const StepParams &p, const RolloutArgs &r, const int lane, const uint32_t salt)
{
    uint32_t h = 0, t = 0;
    if (lane == 0) h = __ldcg(r.tq), t = __ldcg(r.tq + 1);
    h = __shfl_sync(MAPF_FULL_MASK, h, 0);
    t = __shfl_sync(MAPF_FULL_MASK, t, 0);
    const int32_t open = (int32_t)(t - h);
    if (open <= 0) return;
    const uint32_t idx = h + (uint32_t)(((unsigned long long)(salt * 2654435761u) * (uint32_t)open) >> 32);
    uint32_t *ent = tq_entry(r, idx);
    const int took = tq_take<RW>(p, ent, lane, 2, false, 0);
    if (took == 0 && idx == h && lane == 0) atomicMax(r.tq, h + 1);
}

// (out of line: one call in the work-item loop instead of the search's code and registers)
template <int RW>
__device__ __noinline__ void tq_help_call(const StepParams &p, const RolloutArgs &r, const uint32_t salt)
{
    tq_help<RW>(p, r, (int)(threadIdx.x & 31), salt);
}

// worker.py:422-428 inside the launch: a new instance for slot e (generator + heuristic maps of all its agents).  Instance
// number n of slot e is global instance env_offset + n * stride + e of the Philox stream, i.e. what
// mapf_env_reset(mask = {e}, seed, env_offset + n * stride) draws.  Kept out of line: it runs once per episode.
// (Two interleaved searches per half warp were measured too: a lone re-generation drops from 272 to 244 us, but ~20 % more
// instructions, and the rollout with episode handling got slower; profiles/r2_reset_cost.jsonl.)
template <int RW>
__device__ __noinline__ void regenerate_env(const StepParams &p, const RolloutArgs &r, const int e
#ifdef MAPF_ENABLE_DIAG
                                            , unsigned long long &gen_ns
#endif
)
{
    const EnvDims &d = p.d;
    const int lane = threadIdx.x & 31;
    uint32_t n = 0;
    if (lane == 0) {
        n = __ldcg(r.episode + e) + 1;
        r.episode[e] = n;
    }
    n = __shfl_sync(MAPF_FULL_MASK, n, 0);
    const unsigned long long g = r.env_offset + (unsigned long long)n * r.stride + (unsigned long long)e;
#ifdef MAPF_ENABLE_DIAG
    const unsigned long long tg0 = diag_now();
#endif
    reset_env_warp<RW, RW>(d, e, r.seed, g, r.density, const_cast<uint32_t *>(p.obst), p.pos, const_cast<uint8_t *>(p.goal), p.steps,
                           p.err);
    __threadfence();
    __syncwarp();
#ifdef MAPF_ENABLE_DIAG
    gen_ns += diag_now() - tg0;
#endif
    const int groups = bfs_groups<RW>(d);
    if (!r.tq) {
        bfs_all_groups<RW>(p, e, navi_live_of(p, e));
    } else {
        // announce the searches ...
        uint32_t idx = 0;
        if (lane == 0) idx = atomicAdd(r.tq + 1, 1u);
        idx = __shfl_sync(MAPF_FULL_MASK, idx, 0);
        uint32_t *ent = tq_entry(r, idx);
        if (lane == 0) {
            ent[2] = 0u;
            ent[0] = (uint32_t)e;
            __threadfence();
            st_release_u32(ent + 1, 0u);
        }
        __syncwarp();
        // ... take part in them, and wait for the ones other warps took
        while (tq_take<RW>(p, ent, lane, 4, true, e) > 0) {}
        unsigned spins = 0;
        for (;;) {
            uint32_t fin = 0;
            if (lane == 0) fin = ld_acquire_u32(ent + 2);
            fin = __shfl_sync(MAPF_FULL_MASK, fin, 0);
            if (fin == (uint32_t)groups) break;
            // (work-conserving: when every warp is re-generating, a warp that only waited for the groups others took off its
            // entry would leave the machine half idle)
            tq_help<RW>(p, r, lane, (uint32_t)e * 31u + spins);
            __nanosleep(200);
            if (++spins > (1u << 22)) {  // a scheduling bug latches an error instead of hanging the GPU
                if (lane == 0) atomicOr(p.err, MAPF_ERRBIT_INTERNAL);
                break;
            }
        }
        if (lane == 0) {
            ent[1] = kTqDone;
            // move the head hint over the complete entries at the front (helpers pick among [head, tail))
            __threadfence();
            uint32_t h = __ldcg(r.tq);
            const uint32_t t = __ldcg(r.tq + 1);
            while ((int32_t)(t - h) > 0) {
                const int32_t c = (int32_t)__ldcg(tq_entry(r, h) + 1);
                if (c >= 0 || c < (int32_t)0xA0000000u) break;  // open, or reserved and not published yet
                ++h;
            }
            atomicMax(r.tq, h);
        }
    }
    __threadfence();
    __syncwarp();
}

// The episode of slot e has ended and its next instance was staged before the launch (mapf_launch_pregen): the staged obstacle
// bitmap, starts and goals become the live ones and navi_sel[e] flips to the buffer that holds the instance's heuristic maps.
__device__ __forceinline__ void adopt_pregenerated(const StepParams &p, const RolloutArgs &r, const int e)
{
    const EnvDims &d = p.d;
    const int lane = threadIdx.x & 31;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(r.pg_obst + (size_t)e * d.obst_stride);
        uint4 *dst = reinterpret_cast<uint4 *>(const_cast<uint32_t *>(p.obst) + (size_t)e * d.obst_stride);
        for (int w = lane; w < (d.obst_stride >> 2); w += 32) dst[w] = __ldcg(src + w);
    }
    {
        const uint16_t *sp = reinterpret_cast<const uint16_t *>(r.pg_pos) + (size_t)e * d.N;
        const uint16_t *sg = reinterpret_cast<const uint16_t *>(r.pg_goal) + (size_t)e * d.N;
        uint16_t *dp = reinterpret_cast<uint16_t *>(p.pos) + (size_t)e * d.N;
        uint16_t *dg = reinterpret_cast<uint16_t *>(const_cast<uint8_t *>(p.goal)) + (size_t)e * d.N;
        for (int a = lane; a < d.N; a += 32) dp[a] = __ldcg(sp + a), dg[a] = __ldcg(sg + a);
    }
    if (lane == 0) {
        p.steps[e] = 0;
        r.episode[e] = __ldcg(r.episode + e) + 1;
        r.navi_sel[e] = __ldcg(r.navi_sel + e) ^ 1;
        r.pg_epi[e] = 0;
    }
    __threadfence();
    __syncwarp();
}

// EP: episode handling is compiled in.  The plain instantiation (r.max_steps == 0 launches) carries none of it -- the
// re-generation / adoption / task code costs the stepping loop registers (128 per thread are all it has: spills) and
// instruction-cache space even when it never runs: 21.5 -> 25.5 us per step at C2 when it was one kernel.
template <int RW, int K, int WARPS, int MINB, bool EP>
__global__ void __launch_bounds__(WARPS * 32, MINB)
rollout_kernel(const __grid_constant__ StepParams p0, const __grid_constant__ RolloutArgs r)
{
    extern __shared__ __align__(16) uint32_t smem[];
    const EnvDims &d = p0.d;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int N = d.N;
    uint32_t *base = smem + (size_t)warp * r.warp_words;
    uint32_t *s_obst = base;
    uint32_t *s_agent = s_obst + p0.obst_words;
    uint32_t *s_bits = s_agent + p0.obst_words;
    uint16_t *s_tgt = reinterpret_cast<uint16_t *>(s_bits + p0.bits_words);
    uint16_t *s_cell = s_tgt + 32 * K;
    uint64_t *s_tiles = reinterpret_cast<uint64_t *>(base + r.tiles_off);
    uint4 *s_block = reinterpret_cast<uint4 *>(base + r.block_off);
    const uint64_t pol_keep = l2_policy_evict_last();
    const uint64_t pol_stream = l2_policy_evict_first();
    const bool obs_policy = p0.flags & MAPF_STEPF_OBS_POLICY;
    for (int w = lane; w < p0.obst_words; w += 32) s_agent[w] = 0;
    __syncwarp();
    if (r.stagger_ns > 0) {
        // every warp of a fresh launch would otherwise be in the same phase (all resolve conflicts, then all store) for the
        // first few steps
        const unsigned h = ((blockIdx.x * WARPS + warp) * 2654435761u) >> 22;  // 0..1023
        __nanosleep((unsigned)(((unsigned long long)h * (unsigned)r.stagger_ns) >> 10));
    }

    const size_t BN = (size_t)d.B * N;
    const size_t env_bytes = (size_t)N * MAPF_OBS_BYTES_PER_AGENT;
    const int nenv = p0.env_end - p0.env_begin;
    // priority items (see rollout_prio_kernel): used while they are a minority -- when most environments re-generate
    // inside the launch (T of the order of the cap or more) plain time-major chunks balance better
    unsigned nprio = 0;
    if (EP && r.prio) {
        nprio = __ldcg(r.prio);
        // (pre-generated instances are adopted, not generated here: no long items to hand out first)
        if (nprio * 4u > (unsigned)nenv || (r.pg_epi && !r.prio_last && nprio >= (unsigned)r.pregen_min)) nprio = 0;
    }
    const unsigned long long items = (unsigned long long)nprio + (unsigned long long)nenv * (unsigned)r.nchunk;
    bool bulk_used = false;
    uint32_t visits = 0;
    for (;;) {
        if constexpr (EP)
            if (r.tq) tq_help_call<RW>(p0, r, (uint32_t)(blockIdx.x * WARPS + warp) + 7919u * (uint32_t)visits++);
        unsigned long long it = 0;
        if (lane == 0) it = atomicAdd(r.work, 1ull);
        it = __shfl_sync(MAPF_FULL_MASK, it, 0);
        if (it >= items) break;
        int k, e, t0, t1;
        bool whole = false;
        const unsigned long long first_regular = r.prio_last ? 0ull : (unsigned long long)nprio;
        const unsigned long long first_prio = r.prio_last ? items - nprio : 0ull;
        if (it >= first_prio && it < first_prio + nprio) {
            e = (int)__ldcg(r.prio + 1 + (it - first_prio));
            k = 0, t0 = 0, t1 = r.T;
            whole = true;
        } else {
            const unsigned long long j = it - first_regular;
            k = (int)(j / (unsigned)nenv);
            e = p0.env_begin + (int)(j - (unsigned long long)k * (unsigned)nenv);
            if (nprio && __ldcg(r.prio_flag + e)) continue;  // taken as a priority item
            t0 = k * r.chunk, t1 = min(r.T, t0 + r.chunk);
        }
        if (k > 0) {
            // chunk k - 1 of this environment was handed out nenv items ago to a warp that is running; bounded spin so that
            // a scheduling bug latches an error instead of hanging the GPU
            if (lane == 0) {
                unsigned spins = 0;
                while (ld_acquire_u32(r.progress + e) != (uint32_t)k) {
                    __nanosleep(64);
                    if (++spins > (1u << 24)) {
                        atomicOr(p0.err, MAPF_ERRBIT_INTERNAL);
                        break;
                    }
                }
            }
            __syncwarp();
        }
#ifdef MAPF_ENABLE_DIAG
        const unsigned long long tr_claimed = diag_now();
        unsigned long long tr_regen = 0, tr_regens = 0, tr_gen = 0;
#endif
        EnvRegs<K> regs;
        load_env_state<RW, K>(p0, e, lane, s_obst, regs);
#ifdef MAPF_ENABLE_DIAG
        const unsigned long long tr_loaded = diag_now();
#endif
        int sa = t0 % r.action_slots, so = t0 % r.obs_slots, sr = t0 % r.out_slots;
#pragma unroll
        for (int k = 0; k < K; ++k) {  // the first step's actions; every later step's are requested one step ahead
            const int a = k * 32 + lane;
            regs.act_next[k] = a < N ? __ldg(r.actions + (size_t)sa * BN + (size_t)e * N + a) : 0;
        }
        for (int t = t0; t < t1; ++t) {
            StepParams p = p0;
            p.actions = r.actions + (size_t)sa * BN;
            const int sa_next = sa + 1 == r.action_slots ? 0 : sa + 1;
            const uint8_t *next_actions = t + 1 < t1 ? r.actions + (size_t)sa_next * BN : nullptr;
            p.rewards = r.rewards ? r.rewards + (size_t)sr * BN : nullptr;
            p.codes = r.codes ? r.codes + (size_t)sr * BN : nullptr;
            p.done = r.done + (size_t)sr * d.B;
            p.steps_out = r.steps_out ? r.steps_out + (size_t)sr * d.B : nullptr;
            uint8_t *obs_env = r.obs + ((size_t)so * d.B + e) * env_bytes;
            const int head = (int)(reinterpret_cast<uintptr_t>(obs_env) & 15);
            bool reset_step = false;
            if (EP && r.max_steps > 0) {
                // worker.py:390: the episode ended with the previous step (done, or the step cap) -> this step re-generates the
                // slot and emits the new episode's first observation
                const int st = __shfl_sync(MAPF_FULL_MASK, regs.step, 0);
                if (regs.finished || st >= r.max_steps) {
                    int staged = 0;
                    if (r.pg_epi && lane == 0) {
                        const uint32_t have = ld_acquire_u32(r.pg_epi + e);
                        staged = have != 0 && have == __ldcg(r.episode + e) + 1u;
                    }
                    staged = __shfl_sync(MAPF_FULL_MASK, staged, 0);
#ifdef MAPF_ENABLE_DIAG
                    const unsigned long long tr_a = diag_now();
#endif
                    if (staged) adopt_pregenerated(p0, r, e);
#ifdef MAPF_ENABLE_DIAG
                    else regenerate_env<RW>(p0, r, e, tr_gen);
                    tr_regen += diag_now() - tr_a;
                    tr_regens += staged ? 0x10000ull : 1ull;
#else
                    else regenerate_env<RW>(p0, r, e);
#endif
                    {
                        int keep[K];   // load_env_state starts a fresh EnvRegs; the prefetched actions stay
#pragma unroll
                        for (int k = 0; k < K; ++k) keep[k] = regs.act_next[k];
                        load_env_state<RW, K>(p0, e, lane, s_obst, regs);
#pragma unroll
                        for (int k = 0; k < K; ++k) regs.act_next[k] = keep[k];
                    }
                    reset_step = true;
                }
            }
            env_step_gather<RW, K, true, true, true>(p, e, lane, s_obst, s_agent, s_bits, s_tgt, s_cell, head, pol_keep, regs, nullptr,
                                                     s_tiles, reset_step, next_actions);
            if (r.store_mode == 1 && head == 0 && (env_bytes & 15) == 0) {
                // bulk form: bits -> bool bytes into the staging block, then one TMA store of the whole block
                if (bulk_used) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the engine has read the last block
                    __syncwarp();
                }
                const uint16_t *S16 = reinterpret_cast<const uint16_t *>(s_bits);
                const int chunks = (int)(env_bytes >> 4);
#pragma unroll 4
                for (int c = lane; c < chunks; c += 32) {
                    const uint32_t s = S16[c];
                    uint4 v;
                    v.x = expand4(s & 0xfu);
                    v.y = expand4((s >> 4) & 0xfu);
                    v.z = expand4((s >> 8) & 0xfu);
                    v.w = expand4(s >> 12);
                    s_block[c] = v;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the async proxy
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(obs_env), "r"(smem_addr(s_block)),
                                 "r"((uint32_t)env_bytes)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                bulk_used = true;
            } else {
                expand_store_block(p, obs_env, head, env_bytes, s_bits, lane, obs_policy, pol_stream);
            }
            __syncwarp();
            clear_agent_bits<RW, K>(s_agent, regs);
            if (++sa == r.action_slots) sa = 0;
            if (++so == r.obs_slots) so = 0;
            if (++sr == r.out_slots) sr = 0;
        }
#ifdef MAPF_ENABLE_DIAG
        if (r.trace && lane == 0) {
            const unsigned long long slot = atomicAdd(r.trace, 1ull);
            if (slot < (unsigned long long)r.trace_cap) {
                unsigned smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                unsigned long long *rec = r.trace + 1 + 8 * slot;
                rec[0] = it, rec[1] = (unsigned long long)e | ((unsigned long long)t0 << 32), rec[2] = tr_claimed, rec[3] = tr_loaded;
                rec[4] = tr_regen | (tr_gen << 32), rec[5] = tr_regens, rec[6] = diag_now(), rec[7] = smid | ((unsigned long long)(blockIdx.x * WARPS + warp) << 32);
            }
        }
#endif
        if (r.nchunk > 1 && !whole) {
            // publish the chunk: positions / step counter (and a re-generated instance) before the progress word
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                st_release_u32(r.progress + e, k + 1 < r.nchunk ? (uint32_t)(k + 1) : 0u);
            }
        }
    }
    if (lane == 0) {
        if (bulk_used) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        // the last warp to leave re-arms the scheduler (every other warp has made its final, failing claim already)
        __threadfence();
        const unsigned long long left = atomicAdd(r.work + 1, 1ull) + 1ull;
        if (left == (unsigned long long)gridDim.x * WARPS) {
            r.work[0] = 0ull;
            r.work[1] = 0ull;
            if (EP && r.tq) r.tq[0] = r.tq[1] = 0u;
            __threadfence();
        }
    }
}

struct RolloutTuning {
    int warps_per_sm;  // resident warps per SM (0 = automatic: 16, or 24 with episode handling), capped by shared memory
    int chunk;         // steps per work item (0 = automatic)
    int store_mode;    // 0 direct, 1 bulk
    int stagger_ns;
    int reserve_ctas;  // resident CTA slots per SM left to kernels that run beside the launch
};

// MINB = resident CTAs (of 2 warps) per SM the kernel is compiled for: 8 -> 128 registers per thread, 12 -> 85, 16 -> 64
template <int RW, int K, int MINB, bool EP>
int launch_rollout_cfg(mapf_env *env, const StepParams &p, RolloutArgs r, const RolloutTuning &tn, cudaStream_t st)
{
    constexpr int WARPS = 2;
    auto kern = rollout_kernel<RW, K, WARPS, MINB, EP>;
    const EnvDims &d = env->d;
    // per-warp shared memory: step buffers | tile slots (16-byte aligned) | staging block of the bulk form
    const int step_words = (p.warp_smem_words + 3) & ~3;
    const int tile_words = 32 * d.K * kTileStride * 2;
    const int block_words = r.store_mode == 1 ? ((d.N * MAPF_OBS_BYTES_PER_AGENT + 15) / 16) * 4 : 0;
    r.tiles_off = step_words;
    r.block_off = step_words + tile_words;
    r.warp_words = step_words + tile_words + block_words;
    const size_t smem = (size_t)r.warp_words * 4 * WARPS;
    if (smem > 227 * 1024) return MAPF_EINVAL;  // the caller falls back to chains of single-step launches
    // resident CTAs per SM: cached per (kernel, shared-memory size)
    const int key = (((RW * 8 + K) * 2 + r.store_mode) * 32 + MINB) * 2 + (EP ? 1 : 0);
    if (env->ro_key != key) {
        int per_sm = 0;
        // the opt-in maximum, the same for every handle (a smaller value set later would break a larger handle's launches)
        MAPF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MAPF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem));
        if (per_sm < 1) return MAPF_EINVAL;
        env->ro_per_sm = per_sm;
        env->ro_key = key;
    }
    int ctas_per_sm = env->ro_per_sm;
    if (tn.reserve_ctas > 0 && ctas_per_sm > tn.reserve_ctas) ctas_per_sm -= tn.reserve_ctas;
    const int want = ((tn.warps_per_sm > 0 ? tn.warps_per_sm : MINB * WARPS) + WARPS - 1) / WARPS;
    if (ctas_per_sm > want) ctas_per_sm = want;
    const int nenv = p.env_end - p.env_begin;
    // steps per item: one item per environment unless episodes end inside the launch (a re-generation is ~20 steps of work:
    // short items keep the tail short) or there are too few environments to balance the warps
    const int cap_warps = env->num_sms * ctas_per_sm * WARPS;
    int chunk = tn.chunk > 0 ? tn.chunk : (r.max_steps > 0 ? (r.prio ? (r.T >= 16 ? (r.T + 1) / 2 : r.T) : 16) : r.T);
    if (tn.chunk <= 0 && r.max_steps == 0 && (long long)nenv < 3ll * cap_warps && r.T > 8) chunk = 8;
    if (chunk > r.T) chunk = r.T;
    r.chunk = chunk;
    r.nchunk = (r.T + chunk - 1) / chunk;
    // warps: as many as are resident, trimmed (by up to 1/8) to the count that leaves the fewest idle warp-rounds: the stream
    // is bandwidth-bound, so a few warps less cost nothing while a last round that is mostly empty does
    const long long items = (long long)nenv * r.nchunk;
    long long warps = cap_warps < items ? cap_warps : items;
    {
        long long best = warps, best_cost = (items + warps - 1) / warps * warps;
        for (long long w = warps - 1; w >= warps - warps / 8 && w >= 1; --w) {
            const long long cost = (items + w - 1) / w * w;
            if (cost < best_cost) best = w, best_cost = cost;
        }
        warps = best;
    }
    const int grid = (int)((warps + WARPS - 1) / WARPS);
    kern<<<grid, WARPS * 32, smem, st>>>(p, r);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

// geometry dispatch of one occupancy class; MAPF_EINVAL = not served
template <int MINB>
int launch_rollout_class(mapf_env *env, const StepParams &p, const RolloutArgs &r, const RolloutTuning &tn, cudaStream_t st)
{
    switch (env->d.RW * 10 + env->d.K) {
        case 11: return r.max_steps > 0 ? launch_rollout_cfg<1, 1, MINB, true>(env, p, r, tn, st) : launch_rollout_cfg<1, 1, MINB, false>(env, p, r, tn, st);
        case 12: return r.max_steps > 0 ? launch_rollout_cfg<1, 2, MINB, true>(env, p, r, tn, st) : launch_rollout_cfg<1, 2, MINB, false>(env, p, r, tn, st);
        case 21: return r.max_steps > 0 ? launch_rollout_cfg<2, 1, MINB, true>(env, p, r, tn, st) : launch_rollout_cfg<2, 1, MINB, false>(env, p, r, tn, st);
        case 22: return r.max_steps > 0 ? launch_rollout_cfg<2, 2, MINB, true>(env, p, r, tn, st) : launch_rollout_cfg<2, 2, MINB, false>(env, p, r, tn, st);
        case 31: return r.max_steps > 0 ? launch_rollout_cfg<3, 1, MINB, true>(env, p, r, tn, st) : launch_rollout_cfg<3, 1, MINB, false>(env, p, r, tn, st);
        case 32: return r.max_steps > 0 ? launch_rollout_cfg<3, 2, MINB, true>(env, p, r, tn, st) : launch_rollout_cfg<3, 2, MINB, false>(env, p, r, tn, st);
        case 41: return r.max_steps > 0 ? launch_rollout_cfg<4, 1, MINB, true>(env, p, r, tn, st) : launch_rollout_cfg<4, 1, MINB, false>(env, p, r, tn, st);
        case 42: return r.max_steps > 0 ? launch_rollout_cfg<4, 2, MINB, true>(env, p, r, tn, st) : launch_rollout_cfg<4, 2, MINB, false>(env, p, r, tn, st);
    }
    return MAPF_EINVAL;
}

}  // namespace
