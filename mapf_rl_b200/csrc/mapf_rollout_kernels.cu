// mapf_rollout_kernels.cu — host side of the persistent rollout kernel (mapf_rollout_device.cuh): tuning knobs and the choice
// of the occupancy class the launch runs in.  The kernel itself is compiled three times (mapf_rollout_occ8/12/16.cu: 16, 24
// and 32 resident warps per SM at 128, 85 and 64 registers per thread): plain stepping is HBM-bound and fastest with few,
// register-rich warps; with episode handling a warp that re-generates its environment (generator + BFS of every agent:
// ~80 steps' worth of instructions) is out of the store stream for hundreds of microseconds, and more resident warps keep
// the stream fed and the idle issue slots busy meanwhile.
#include "mapf_rollout_device.cuh"

int mapf_launch_rollout_occ8(mapf_env *, const StepParams &, const void *, const void *, cudaStream_t);
int mapf_launch_rollout_occ10(mapf_env *, const StepParams &, const void *, const void *, cudaStream_t);
int mapf_launch_rollout_occ12(mapf_env *, const StepParams &, const void *, const void *, cudaStream_t);
int mapf_launch_rollout_occ16(mapf_env *, const StepParams &, const void *, const void *, cudaStream_t);

namespace {

RolloutTuning &rollout_tuning()
{
    static RolloutTuning t = [] {
        RolloutTuning r{0, 0, 0, 1000, 0};
        if (const char *s = std::getenv("MAPF_ROLLOUT_WARPS_PER_SM")) r.warps_per_sm = std::atoi(s);
        if (const char *s = std::getenv("MAPF_ROLLOUT_CHUNK")) r.chunk = std::atoi(s);
        if (const char *s = std::getenv("MAPF_ROLLOUT_STORE_MODE")) r.store_mode = std::atoi(s);
        if (const char *s = std::getenv("MAPF_ROLLOUT_STAGGER_NS")) r.stagger_ns = std::atoi(s);
        return r;
    }();
    return t;
}

}  // namespace

void mapf_set_rollout_tuning(int warps_per_sm, int chunk, int store_mode, int stagger_ns)
{
    RolloutTuning &t = rollout_tuning();
    if (warps_per_sm >= 0) t.warps_per_sm = warps_per_sm;
    if (chunk >= 0) t.chunk = chunk;
    if (store_mode >= 0) t.store_mode = store_mode;
    if (stagger_ns >= 0) t.stagger_ns = stagger_ns;
}

bool mapf_rollout_supported(const mapf_env *env) { return env->d.K <= 2; }

// MAPF_ROLLOUT_PREGEN / mapf_debug_rollout_pregen: 0 = episodes that end inside a rollout launch always re-generate in place,
// 1 = automatic, 2 = pre-generate before the launch, 3 = pre-generate beside the launch (see mapf_launch_rollout)
int &rollout_pregen_ref()
{
    static int v = [] {
        const char *s = std::getenv("MAPF_ROLLOUT_PREGEN");
        return s ? std::atoi(s) : 1;
    }();
    return v;
}
void mapf_set_rollout_pregen(int on) { rollout_pregen_ref() = on; }
// MAPF_ROLLOUT_TASKS / mapf_debug_rollout_tasks: 1 = the searches of an in-launch re-generation are announced as tasks any warp
// takes (mapf_rollout_device.cuh).  Off by default: it shortens a lone re-generation (the tail of a launch with FEW, LONG
// re-generations: large maps, many agents, small batches) but a claim / completion pair costs ~3 us of L2 round trips per
// batch of searches and a helper leaves the owner waiting -- at C2 / C3 (40x40) it is 0-3 % slower, and with every environment
// re-generating at once 2.2x slower (profiles/r2_reset_cost.jsonl).
int &rollout_tasks_ref()
{
    static int v = [] {
        const char *s = std::getenv("MAPF_ROLLOUT_TASKS");
        return s ? std::atoi(s) : 0;
    }();
    return v;
}
void mapf_set_rollout_tasks(int on) { rollout_tasks_ref() = on; }
#ifdef MAPF_ENABLE_DIAG
// diagnosis build only: per-item time stamps of the next rollout launches (profiles/tools/r2_rollout_timeline.py)
static unsigned long long *g_trace = nullptr;
static int g_trace_cap = 0;
// cycles of the search phases of environment 0 inside rollout_kernel (each translation unit has its own copy of the counters:
// this one reads the occ8 kernels' through mapf_diag_bfs_cycles_occ8)
extern "C" int mapf_diag_rollout_trace(unsigned long long *d_buf, int capacity)
{
    g_trace = d_buf, g_trace_cap = capacity;
    return 0;
}
#endif
int mapf_launch_pregen(mapf_env *env, int min_count, cudaStream_t st);

// The persistent rollout over environments [e0, e1); MAPF_EINVAL = geometry not served (the caller uses chains of launches).
int mapf_launch_rollout(mapf_env *env, int e0, int e1, int T, const uint8_t *d_actions, int action_slots, uint8_t *d_obs, int obs_slots,
                        const StepOut &out, int out_slots, cudaStream_t st)
{
    if (!mapf_rollout_supported(env)) return MAPF_EINVAL;
    StepParams p = mapf_make_step_params(env);
    p.env_begin = e0;
    p.env_end = e1;
    RolloutArgs r{};
    r.T = T, r.action_slots = action_slots, r.obs_slots = obs_slots, r.out_slots = out_slots;
    r.actions = d_actions, r.obs = d_obs, r.rewards = out.rewards, r.codes = out.codes, r.done = out.done, r.steps_out = out.steps;
    r.work = env->ro_work, r.progress = env->ro_progress;
    r.store_mode = rollout_tuning().store_mode;
    r.stagger_ns = rollout_tuning().stagger_ns;
    r.max_steps = env->ar_max_steps;
    r.seed = env->ar_seed, r.env_offset = env->ar_offset, r.stride = env->ar_stride, r.density = env->ar_density;
    r.episode = env->ro_episode;
#ifdef MAPF_ENABLE_DIAG
    r.trace = g_trace, r.trace_cap = g_trace_cap;
#endif
    static const int use_prio = [] { const char *v = std::getenv("MAPF_ROLLOUT_PRIO"); return v ? std::atoi(v) : 1; }();
    r.prio = nullptr, r.prio_flag = nullptr, r.pg_epi = nullptr, r.prio_last = 0;
    r.tq = nullptr;
    int reserve_ctas = 0;
    if (r.max_steps > 0) {
        if (rollout_tasks_ref()) r.tq = env->ro_tq, r.tq_cap = env->ro_tq_cap;
        // The environments that hit the cap inside this launch are known up front (rollout_prio_kernel), and their next
        // instances can be generated ahead by the dedicated generator / BFS kernels into the staging arrays and the second
        // heuristic-map buffer; the rollout kernel adopts them at the episode's end.
        //  * MANY are due (the estimate B * T / cap reaches 16 x SMs): the dedicated kernels run at full occupancy BEFORE the
        //    rollout kernel, on the same stream (C4, 80x80 / 64 agents / cap 32: 279 -> 155 us per step).
        //  * FEW: the dedicated kernels would be latency-bound (a lone generator takes ~130 us: +190 us in front of a 430-us
        //    launch at C2), so the episodes re-generate inside the rollout kernel -- generator by the slot's warp, searches as
        //    tasks -- and the listed environments, its long items, are handed out first.
        //  * (mode 3, not automatic) BESIDE the rollout kernel on a high-priority stream: the rollout kernel leaves CTA slots
        //    free, hands the listed environments out LAST and adopts an instance only if it has been published by then.  It
        //    moves the same work to other warps and takes slots from the store stream while it runs: C3 57.4 -> 51.7 us per step,
        //    but C2 28.9 -> 30 us (two slots per SM reserved) or 41 us (one slot: the instances are late, every episode end
        //    re-generates at the end of the launch); profiles/r2_rollout_timeline_beside.jsonl.
        // The true count is only known on the device; the estimate picks the form, never the result.
        const int mode = rollout_pregen_ref();   // 0 off, 1 automatic, 2 always before, 3 always beside
        const bool have_mem = env->navi_alt != nullptr;
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        MAPF_CUDA(cudaStreamIsCapturing(st, &cap));
        const long long due = (long long)(e1 - e0) * (T < r.max_steps ? T : r.max_steps) / r.max_steps;
        bool before = false, beside = false;
        if (have_mem && mode != 0) {
            if (mode == 2) before = true;
            else if (mode == 3) beside = true;
            else if (due >= (long long)env->num_sms * 16) before = true;
            if (beside && cap != cudaStreamCaptureStatusNone) beside = false;  // (a captured launch re-generates in place)
        }
        const bool prio = use_prio && T <= r.max_steps / 2;
        if (before || beside || prio) {
            rollout_prio_kernel<<<1, 1024, 0, st>>>(env->steps, e0, e1, T, r.max_steps, env->ro_prio, env->ro_prio_flag, env->pg_epi,
                                                    env->pg_cnt, env->ro_work);
            MAPF_CUDA(cudaGetLastError());
        }
        r.pregen_min = 0;
        if (before) {
            const int rc = mapf_launch_pregen(env, 0, st);
            if (rc != MAPF_OK) return rc;
        } else if (beside) {
            if (!env->pg_stream) {
                int lo = 0, hi = 0;
                MAPF_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                MAPF_CUDA(cudaStreamCreateWithPriority(&env->pg_stream, cudaStreamNonBlocking, hi));
                MAPF_CUDA(cudaEventCreateWithFlags(&env->pg_fork, cudaEventDisableTiming));
                MAPF_CUDA(cudaEventCreateWithFlags(&env->pg_join, cudaEventDisableTiming));
            }
            MAPF_CUDA(cudaEventRecord(env->pg_fork, st));
            MAPF_CUDA(cudaStreamWaitEvent(env->pg_stream, env->pg_fork, 0));
            const int rc = mapf_launch_pregen(env, 0, env->pg_stream);
            if (rc != MAPF_OK) return rc;
            MAPF_CUDA(cudaEventRecord(env->pg_join, env->pg_stream));
            r.prio_last = 1;
            static const int reserve = [] { const char *v = std::getenv("MAPF_ROLLOUT_RESERVE"); return v ? std::atoi(v) : 1; }();
            reserve_ctas = reserve;
        }
        if (before || beside) {
            r.pg_epi = env->pg_epi, r.pg_obst = env->pg_obst, r.pg_pos = env->pg_pos, r.pg_goal = env->pg_goal;
            r.navi_sel = env->navi_sel;
        }
        if (prio || beside) r.prio = env->ro_prio, r.prio_flag = env->ro_prio_flag;
    }
    // occupancy class: the smallest one that holds the requested warps
    const RolloutTuning &tn = rollout_tuning();
    const int want = tn.warps_per_sm > 0 ? tn.warps_per_sm : 16;  // measured: 16 register-rich warps beat 24 / 32 leaner ones, with and without episode handling (profiles/r2_rollout_sweep.jsonl)
    RolloutTuning use = tn;
    use.warps_per_sm = want;
    use.reserve_ctas = reserve_ctas;
    int rc;
    if (want <= 16) rc = mapf_launch_rollout_occ8(env, p, &r, &use, st);
    else if (want <= 20) rc = mapf_launch_rollout_occ10(env, p, &r, &use, st);
    else if (want <= 24) rc = mapf_launch_rollout_occ12(env, p, &r, &use, st);
    else rc = mapf_launch_rollout_occ16(env, p, &r, &use, st);
    // the pre-generation beside the launch joins the caller's stream (it also has to be over before the next launch's list)
    if (r.prio_last) MAPF_CUDA(cudaStreamWaitEvent(st, env->pg_join, 0));
    return rc;
}
