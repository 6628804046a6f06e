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
        RolloutTuning r{0, 0, 0, 1000};
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
// 1 = pre-generate when at least 16 x SMs environments end in the launch, n > 1 = when at least n do
int &rollout_pregen_ref()
{
    static int v = [] {
        const char *s = std::getenv("MAPF_ROLLOUT_PREGEN");
        return s ? std::atoi(s) : 1;
    }();
    return v;
}
void mapf_set_rollout_pregen(int on) { rollout_pregen_ref() = on; }
#ifdef MAPF_ENABLE_DIAG
// diagnosis build only: per-item time stamps of the next rollout launches (profiles/tools/r2_rollout_timeline.py)
static unsigned long long *g_trace = nullptr;
static int g_trace_cap = 0;
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
    r.prio = nullptr, r.prio_flag = nullptr, r.pg_flag = nullptr;
    r.tq = nullptr;
    if (r.max_steps > 0) {
        static const int use_tq = [] { const char *v = std::getenv("MAPF_ROLLOUT_TASKS"); return v ? std::atoi(v) : 1; }();
        if (use_tq) r.tq = env->ro_tq, r.tq_cap = env->ro_tq_cap;
        // The environments that hit the cap inside this launch are known up front (rollout_prio_kernel).  MANY of them: their
        // next instances are generated now, by the dedicated generator / BFS kernels at full occupancy, into the staging arrays
        // and the second heuristic-map buffer, and the rollout kernel adopts them at the episode's end.  FEW (the dedicated
        // kernels would be latency-bound: a lone generator takes ~130 us): they re-generate inside the rollout kernel, are
        // its long items and are handed out first.  The count is only known on the device: both the pre-generation kernels
        // and the rollout kernel compare it with pregen_min.
        const bool pregen = rollout_pregen_ref() != 0 && env->navi_alt != nullptr;
        const bool prio = use_prio && T <= r.max_steps / 2;
        if (pregen || prio) {
            rollout_prio_kernel<<<1, 1024, 0, st>>>(env->steps, e0, e1, T, r.max_steps, env->ro_prio, env->ro_prio_flag, env->pg_flag,
                                                    env->ro_work);
            MAPF_CUDA(cudaGetLastError());
        }
        r.pregen_min = rollout_pregen_ref() > 1 ? rollout_pregen_ref() : env->num_sms * 16;
        if (pregen) {
            const int rc = mapf_launch_pregen(env, r.pregen_min, st);
            if (rc != MAPF_OK) return rc;
            r.pg_flag = env->pg_flag, r.pg_obst = env->pg_obst, r.pg_pos = env->pg_pos, r.pg_goal = env->pg_goal;
            r.navi_sel = env->navi_sel;
        }
        if (prio) r.prio = env->ro_prio, r.prio_flag = env->ro_prio_flag;
    }
    // occupancy class: the smallest one that holds the requested warps
    const RolloutTuning &tn = rollout_tuning();
    const int want = tn.warps_per_sm > 0 ? tn.warps_per_sm : 16;  // measured: 16 register-rich warps beat 24 / 32 leaner ones, with and without episode handling (profiles/r2_rollout_sweep.jsonl)
    RolloutTuning use = tn;
    use.warps_per_sm = want;
    if (want <= 16) return mapf_launch_rollout_occ8(env, p, &r, &use, st);
    if (want <= 20) return mapf_launch_rollout_occ10(env, p, &r, &use, st);
    if (want <= 24) return mapf_launch_rollout_occ12(env, p, &r, &use, st);
    return mapf_launch_rollout_occ16(env, p, &r, &use, st);
}
