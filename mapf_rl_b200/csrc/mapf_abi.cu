// mapf_abi.cu — extern "C" entry points of libmapf_b200.so (declared in include/mapf_b200.h).
// Argument validation, arena ownership and stream plumbing only; the kernels live in
// mapf_env_kernels.cu / mapf_reset_kernels.cu / mapf_per_kernels.cu.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "mapf_common.cuh"

// launchers (other translation units)
int mapf_launch_pack_load(mapf_env *, const int32_t *, int, const uint8_t *, const uint8_t *, const uint8_t *, cudaStream_t);
int mapf_launch_bfs(mapf_env *, const int32_t *, int, int32_t *, cudaStream_t);
int mapf_launch_step(mapf_env *, const uint8_t *, uint8_t *, const int64_t *, float *, uint8_t *, int32_t *, cudaStream_t);
int mapf_launch_observe(mapf_env *, uint8_t *, const int64_t *, uint8_t *, cudaStream_t);
int mapf_launch_step_only(mapf_env *, const uint8_t *, float *, uint8_t *, int32_t *, cudaStream_t);
int mapf_launch_step_range(mapf_env *, int, int, const uint8_t *, uint8_t *, float *, uint8_t *, int32_t *, cudaStream_t);
int mapf_launch_rollout_persistent(mapf_env *, int, const uint8_t *, int, uint8_t *, int, float *, uint8_t *, int32_t *, int, int,
                                   int, cudaStream_t);
void mapf_set_step_tuning(int, int, int);
int mapf_step_tuning_generation();
void mapf_set_step_trace(unsigned long long *);
int mapf_launch_unpack(mapf_env *, uint8_t *, uint8_t *, cudaStream_t);
int mapf_launch_comm_mask(mapf_env *, int, uint8_t *, cudaStream_t);
int mapf_launch_reset(mapf_env *, const uint8_t *, uint64_t, uint64_t, float, cudaStream_t);
int mapf_launch_per_update(mapf_per *, PerScratch *, const int64_t *, const double *, int64_t, cudaStream_t);
int mapf_launch_per_sample(mapf_per *, const double *, int64_t, int64_t *, double *, float *, double, cudaStream_t);
int mapf_launch_per_td_update(mapf_per *, PerScratch *, const float *, const float *, const float *, const int64_t *,
                              const float *, const float *, const float *, const int64_t *, int64_t, float, double, int64_t,
                              int64_t, int64_t, float *, float *, cudaStream_t);
int mapf_launch_actor_td(const float *, const float *, const uint8_t *, const int32_t *, int, int, double *, cudaStream_t);
int mapf_launch_replay_gather(const mapf_replay_view *, const int64_t *, int64_t, const mapf_replay_batch *, int32_t *, cudaStream_t);

static thread_local std::string g_last_error;

void mapf_set_error(const std::string &msg) { g_last_error = msg; }

int mapf_cuda_fail(cudaError_t e, const char *what)
{
    g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
    return MAPF_ECUDA;
}

namespace {
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) return;
        ok = (prev == dev) || cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (ok && prev >= 0) cudaSetDevice(prev);
    }
};

bool host_is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();  // clear: an unregistered pointer is not an error for us
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// Device-side alias of a page-locked, mapped host buffer (cudaHostAlloc / torch pin_memory under UVA), or
// nullptr when the buffer cannot be addressed by a kernel.
template <typename T>
T *host_device_alias(T *p)
{
    void *dptr = nullptr;
    if (!p || cudaHostGetDevicePointer(&dptr, const_cast<void *>(static_cast<const void *>(p)), 0) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return static_cast<T *>(dptr);
}

// MAPF_STEP_HOST_MODE (measured on 8192 x 32 agents, profiles/e2e_modes.py):
//   0  DMA copies either side of the fused kernel                                              109 us per call
//   1  the fused kernel stores rewards / done / steps straight into the page-locked buffers    ~97 us
//   2  ... and reads the actions in place (the 1 MB of PCIe writes stretches the kernel 33 -> 58 us)   87 us
//   3  split: step kernel, then the observe kernel while the results are copied on a side stream     88 us
//   4  the sequence of 3 captured once per buffer set and replayed with one cudaGraphLaunch (default) 78 us
// Measured and dropped: the batch as 2 / 4 / 8 sub-batch chains, each {actions -> fused kernel -> results} on its own stream
// inside the graph (105 / 124 / 160 us: every extra DMA node costs more than the overlap returns; profiles/r1_e2e_chains.log).
int &step_host_mode_ref()
{
    static int m = [] {
        const char *s = std::getenv("MAPF_STEP_HOST_MODE");
        return s ? std::atoi(s) : 4;
    }();
    return m;
}
int step_host_mode() { return step_host_mode_ref(); }
template <typename T>
int dev_alloc(T **p, size_t count, int64_t *total)
{
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(p), bytes);
    if (e != cudaSuccess) {
        mapf_cuda_fail(e, "cudaMalloc");
        return MAPF_ENOMEM;
    }
    *total += (int64_t)bytes;
    return MAPF_OK;
}
}  // namespace

extern "C" {

int mapf_abi_version(void) { return MAPF_ABI_VERSION; }

const char *mapf_last_error(void) { return g_last_error.c_str(); }

int mapf_env_create(const mapf_env_config *cfg, mapf_env **out)
{
    if (!cfg || !out) {
        mapf_set_error("mapf_env_create: NULL argument");
        return MAPF_EINVAL;
    }
    *out = nullptr;
    if (cfg->num_envs < 1 || cfg->num_agents < 1 || cfg->num_agents > MAPF_MAX_AGENTS || cfg->map_length < 2 ||
        cfg->map_length > MAPF_MAX_MAP_SIDE) {
        mapf_set_error("mapf_env_create: num_envs >= 1, 1 <= num_agents <= 128, 2 <= map_length <= 120 required");
        return MAPF_EINVAL;
    }
    if (cfg->obs_radius != MAPF_OBS_RADIUS) {
        mapf_set_error("mapf_env_create: only obs_radius = 4 is supported (config.py:14)");
        return MAPF_EINVAL;
    }
    if ((int64_t)cfg->num_agents > (int64_t)cfg->map_length * cfg->map_length) {
        mapf_set_error("mapf_env_create: more agents than cells");
        return MAPF_EINVAL;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        mapf_set_error("mapf_env_create: no CUDA device (this library has no CPU fallback)");
        return MAPF_ECUDA;
    }
    if (cfg->device < 0 || cfg->device >= ndev) {
        mapf_set_error("mapf_env_create: bad device ordinal");
        return MAPF_EINVAL;
    }
    DeviceGuard guard(cfg->device);
    if (!guard.ok) {
        mapf_set_error("mapf_env_create: cudaSetDevice failed");
        return MAPF_ECUDA;
    }
    mapf_env *env = new (std::nothrow) mapf_env();
    if (!env) return MAPF_ENOMEM;
    std::memset(env, 0, sizeof(*env));
    EnvDims &d = env->d;
    d.B = cfg->num_envs;
    d.N = cfg->num_agents;
    d.L = cfg->map_length;
    d.R = d.L + 8;
    d.RW = (d.L + 8 + 31) / 32;
    d.RWS = d.RW + 1;
    d.NB = (d.L + 7) / 8;
    d.K = (d.N + 31) / 32;
    d.obst_stride = (d.R * d.RWS + 3) & ~3;
    d.navi_agent_stride = d.NB * d.NB * 32;
    env->device = cfg->device;
    env->num_sms = 148;
    cudaDeviceGetAttribute(&env->num_sms, cudaDevAttrMultiProcessorCount, cfg->device);
    for (int i = 0; i < 5; ++i) env->reward[i] = cfg->reward_fn[i];

    int rc = MAPF_OK;
    int64_t total = 0;
    const size_t BN = (size_t)d.B * d.N;
    if (rc == MAPF_OK) rc = dev_alloc(&env->obst, (size_t)d.B * d.obst_stride, &total);
    if (rc == MAPF_OK) rc = dev_alloc(&env->pos, BN * 2, &total);
    if (rc == MAPF_OK) rc = dev_alloc(&env->goal, BN * 2, &total);
    if (rc == MAPF_OK) rc = dev_alloc(&env->navi, BN * d.navi_agent_stride, &total);
    if (rc == MAPF_OK) rc = dev_alloc(&env->steps, (size_t)d.B, &total);
    if (rc == MAPF_OK) rc = dev_alloc(&env->err, 1, &total);
    if (rc == MAPF_OK) {
        cudaError_t e2 = cudaMemset(env->obst, 0, (size_t)d.B * d.obst_stride * 4);
        if (e2 == cudaSuccess) e2 = cudaMemset(env->pos, 0, BN * 2);
        if (e2 == cudaSuccess) e2 = cudaMemset(env->goal, 0, BN * 2);
        if (e2 == cudaSuccess) e2 = cudaMemset(env->navi, 0, BN * d.navi_agent_stride * 4);
        if (e2 == cudaSuccess) e2 = cudaMemset(env->steps, 0, (size_t)d.B * 4);
        if (e2 == cudaSuccess) e2 = cudaMemset(env->err, 0, 4);
        if (e2 == cudaSuccess) e2 = cudaDeviceSynchronize();
        if (e2 != cudaSuccess) rc = mapf_cuda_fail(e2, "cudaMemset(arena)");
    }
    env->arena_bytes = total;
    if (rc != MAPF_OK) {
        mapf_env_destroy(env);
        return rc;
    }
    *out = env;
    return MAPF_OK;
}

int mapf_env_destroy(mapf_env *env)
{
    if (!env) return MAPF_OK;
    DeviceGuard guard(env->device);
    cudaFree(env->obst);
    cudaFree(env->pos);
    cudaFree(env->goal);
    cudaFree(env->navi);
    cudaFree(env->steps);
    cudaFree(env->err);
    cudaFree(env->d_actions);
    cudaFree(env->d_obs);
    cudaFree(env->d_rewards);
    cudaFree(env->d_done);
    cudaFree(env->d_steps_out);
    if (env->h_pinned) cudaFreeHost(env->h_pinned);
    for (auto &c : env->hg)
        if (c.exec) cudaGraphExecDestroy(c.exec);
    if (env->cap_stream) cudaStreamDestroy(env->cap_stream);
    for (int j = 0; j < MAPF_MAX_CHAINS; ++j) {
        if (env->chain_stream[j]) cudaStreamDestroy(env->chain_stream[j]);
        if (env->chain_done[j]) cudaEventDestroy(env->chain_done[j]);
    }
    if (env->chain_fork) cudaEventDestroy(env->chain_fork);
    for (auto &g : env->rg)
        for (auto &x : g.exec)
            if (x) cudaGraphExecDestroy(x);
    if (env->side_stream) cudaStreamDestroy(env->side_stream);
    if (env->ev_stepped) cudaEventDestroy(env->ev_stepped);
    if (env->ev_copied) cudaEventDestroy(env->ev_copied);
    delete env;
    return MAPF_OK;
}

int64_t mapf_env_arena_bytes(const mapf_env *env) { return env ? env->arena_bytes : 0; }

#define REQUIRE_ENV(env)                          \
    if (!(env)) {                                 \
        mapf_set_error("NULL environment handle"); \
        return MAPF_EINVAL;                       \
    }                                             \
    DeviceGuard guard((env)->device);             \
    if (!guard.ok) {                              \
        mapf_set_error("cudaSetDevice failed");   \
        return MAPF_ECUDA;                        \
    }

int mapf_env_load(mapf_env *env, const int32_t *d_env_ids, int32_t n, const uint8_t *d_maps, const uint8_t *d_agents,
                  const uint8_t *d_goals, void *stream)
{
    REQUIRE_ENV(env);
    if (n < 0 || n > env->d.B || !d_maps || !d_agents || !d_goals) {
        mapf_set_error("mapf_env_load: bad arguments");
        return MAPF_EINVAL;
    }
    if (n == 0) return MAPF_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc = mapf_launch_pack_load(env, d_env_ids, n, d_maps, d_agents, d_goals, st);
    if (rc != MAPF_OK) return rc;
    return mapf_launch_bfs(env, d_env_ids, n, nullptr, st);
}

int mapf_env_bfs_navi(mapf_env *env, const int32_t *d_env_ids, int32_t n, int32_t *d_dist_out, void *stream)
{
    REQUIRE_ENV(env);
    if (!d_env_ids) n = env->d.B;
    if (n < 0 || n > env->d.B) {
        mapf_set_error("mapf_env_bfs_navi: bad n");
        return MAPF_EINVAL;
    }
    if (n == 0) return MAPF_OK;
    return mapf_launch_bfs(env, d_env_ids, n, d_dist_out, static_cast<cudaStream_t>(stream));
}

int mapf_env_step_observe(mapf_env *env, const uint8_t *d_actions, uint8_t *d_obs, float *d_rewards, uint8_t *d_done,
                          int32_t *d_steps, void *stream)
{
    REQUIRE_ENV(env);
    if (!d_actions || !d_obs || !d_rewards || !d_done) {
        mapf_set_error("mapf_env_step_observe: NULL buffer");
        return MAPF_EINVAL;
    }
    return mapf_launch_step(env, d_actions, d_obs, nullptr, d_rewards, d_done, d_steps, static_cast<cudaStream_t>(stream));
}

int mapf_env_step_observe_rows(mapf_env *env, const uint8_t *d_actions, uint8_t *d_obs_base, const int64_t *d_obs_rows,
                               float *d_rewards, uint8_t *d_done, int32_t *d_steps, void *stream)
{
    REQUIRE_ENV(env);
    if (!d_actions || !d_obs_base || !d_obs_rows || !d_rewards || !d_done) {
        mapf_set_error("mapf_env_step_observe_rows: NULL buffer");
        return MAPF_EINVAL;
    }
    return mapf_launch_step(env, d_actions, d_obs_base, d_obs_rows, d_rewards, d_done, d_steps,
                            static_cast<cudaStream_t>(stream));
}

int mapf_env_observe_rows(mapf_env *env, uint8_t *d_obs_base, const int64_t *d_obs_rows, uint8_t *d_pos, void *stream)
{
    REQUIRE_ENV(env);
    if (!d_obs_base || !d_obs_rows) {
        mapf_set_error("mapf_env_observe_rows: NULL buffer");
        return MAPF_EINVAL;
    }
    return mapf_launch_observe(env, d_obs_base, d_obs_rows, d_pos, static_cast<cudaStream_t>(stream));
}

int mapf_env_observe(mapf_env *env, uint8_t *d_obs, uint8_t *d_pos, void *stream)
{
    REQUIRE_ENV(env);
    if (!d_obs) {
        mapf_set_error("mapf_env_observe: NULL buffer");
        return MAPF_EINVAL;
    }
    return mapf_launch_observe(env, d_obs, nullptr, d_pos, static_cast<cudaStream_t>(stream));
}

namespace {
// How mapf_env_rollout runs T steps: S chains of `per` environments each; P > 0 when whole slot periods of P steps are
// replayed from per-chain captured graphs.
struct RolloutPlan {
    int S, per, P;
    bool graphs;
};
// Process-wide knobs of the persistent rollout kernel (mapf_debug_rollout_tuning; read once from the environment):
//   MAPF_ROLLOUT_PERSISTENT=0      never use it (chains of launches instead)
//   MAPF_ROLLOUT_ENVS_PER_WARP=n   environments each resident warp takes through their T steps (0 = B / (16 warps per SM))
//   MAPF_ROLLOUT_CTA_WARPS=1|2|4   warps per CTA (default 2)
struct RolloutTuning {
    int persistent, envs_per_warp, cta_warps;
};
RolloutTuning &rollout_tuning()
{
    static RolloutTuning t = [] {
        RolloutTuning r{1, 0, 2};
        if (const char *s = std::getenv("MAPF_ROLLOUT_PERSISTENT")) r.persistent = std::atoi(s);
        if (const char *s = std::getenv("MAPF_ROLLOUT_ENVS_PER_WARP")) r.envs_per_warp = std::atoi(s);
        if (const char *s = std::getenv("MAPF_ROLLOUT_CTA_WARPS")) r.cta_warps = std::atoi(s);
        return r;
    }();
    return t;
}
bool rollout_persistent_enabled() { return rollout_tuning().persistent != 0; }
// the persistent kernel serves the default request (chains = 0) on the hot geometry when the rollout is long enough for its
// environment-major order to fill the GPU
bool rollout_uses_persistent(const EnvDims &d, int T, int chains)
{
    return rollout_persistent_enabled() && chains == 0 && d.RW == 2 && d.K == 1 && d.B >= 2048 && T >= 16;
}

RolloutPlan rollout_plan(const EnvDims &d, int T, int action_slots, int obs_slots, int out_slots, int chains, bool capturing)
{
    auto gcd = [](int a, int b) { while (b) { int r = a % b; a = b; b = r; } return a; };
    auto lcm_cap = [&](int a, int b) { const long long l = (long long)a / gcd(a, b) * b; return l > 64 ? 65 : (int)l; };
    RolloutPlan pl;
    pl.P = lcm_cap(lcm_cap(action_slots, obs_slots), out_slots);
    pl.graphs = pl.P <= 64 && T >= 4 * pl.P && d.B >= 2048 && chains != 1 && !capturing;
    // sub-batches are multiples of 4 environments (one CTA of the step kernel serves 4)
    pl.S = chains ? chains : (d.B >= 2048 ? (pl.graphs ? 8 : 4) : 1);
    pl.per = (((d.B + pl.S - 1) / pl.S) + 3) & ~3;
    pl.S = (d.B + pl.per - 1) / pl.per;
    if (pl.S == 1) pl.graphs = false;
    return pl;
}
}  // namespace

int mapf_env_rollout_plan(mapf_env *env, int32_t T, int32_t action_slots, int32_t obs_slots, int32_t out_slots, int32_t chains,
                          int32_t *chains_out, int32_t *envs_per_chain_out, int32_t *graph_period_out)
{
    REQUIRE_ENV(env);
    if (T < 0 || action_slots < 1 || obs_slots < 1 || out_slots < 1 || chains < 0 || chains > MAPF_MAX_CHAINS) {
        mapf_set_error("mapf_env_rollout_plan: T >= 0, slot counts >= 1 and 0 <= chains <= 16 required");
        return MAPF_EINVAL;
    }
    if (rollout_uses_persistent(env->d, T, chains)) {  // one launch for the whole rollout
        if (chains_out) *chains_out = 0;
        if (envs_per_chain_out) *envs_per_chain_out = env->d.B;
        if (graph_period_out) *graph_period_out = 0;
        return MAPF_OK;
    }
    const RolloutPlan pl = rollout_plan(env->d, T, action_slots, obs_slots, out_slots, chains, false);
    if (chains_out) *chains_out = pl.S;
    if (envs_per_chain_out) *envs_per_chain_out = pl.per;
    if (graph_period_out) *graph_period_out = pl.graphs ? pl.P : 0;
    return MAPF_OK;
}

int mapf_env_rollout(mapf_env *env, int32_t T, const uint8_t *d_actions, int32_t action_slots, uint8_t *d_obs, int32_t obs_slots,
                     float *d_rewards, uint8_t *d_done, int32_t *d_steps, int32_t out_slots, int32_t chains, void *stream)
{
    REQUIRE_ENV(env);
    if (!d_actions || !d_obs || !d_rewards || !d_done) {
        mapf_set_error("mapf_env_rollout: NULL buffer");
        return MAPF_EINVAL;
    }
    if (T < 0 || action_slots < 1 || obs_slots < 1 || out_slots < 1 || chains < 0 || chains > MAPF_MAX_CHAINS) {
        mapf_set_error("mapf_env_rollout: T >= 0, slot counts >= 1 and 0 <= chains <= 16 required");
        return MAPF_EINVAL;
    }
    if (T == 0) return MAPF_OK;
    const EnvDims &d = env->d;
    const size_t BN = (size_t)d.B * d.N;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (rollout_uses_persistent(d, T, chains)) {
        const int rc = mapf_launch_rollout_persistent(env, T, d_actions, action_slots, d_obs, obs_slots, d_rewards, d_done, d_steps,
                                                      out_slots, rollout_tuning().envs_per_warp, rollout_tuning().cta_warps, st);
        if (rc != MAPF_EINVAL) return rc;
    }
    // The launches of a long rollout repeat with period P = lcm(slot counts): those are captured once per chain into a
    // graph of P kernel nodes and replayed (one cudaGraphLaunch per chain and period instead of P launches of ~4 us of
    // host time each, which bound 8 chains at 32 us per step); short rollouts and the tail are launched directly.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    MAPF_CUDA(cudaStreamIsCapturing(st, &cap));
    const RolloutPlan pl = rollout_plan(d, T, action_slots, obs_slots, out_slots, chains, cap != cudaStreamCaptureStatusNone);
    const int S = pl.S, per = pl.per, P = pl.P;
    const bool graphs = pl.graphs;
    auto step_t = [&](int t, int e0, int e1, cudaStream_t q) -> int {
        const size_t sa = (size_t)(t % action_slots), so = (size_t)(t % obs_slots), sr = (size_t)(t % out_slots);
        return mapf_launch_step_range(env, e0, e1, d_actions + sa * BN, d_obs + so * BN * MAPF_OBS_BYTES_PER_AGENT, d_rewards + sr * BN,
                                      d_done + sr * d.B, d_steps ? d_steps + sr * d.B : nullptr, q);
    };
    if (S == 1) {
        for (int t = 0; t < T; ++t) {
            const int rc = step_t(t, 0, d.B, st);
            if (rc != MAPF_OK) return rc;
        }
        return MAPF_OK;
    }
    if (!env->chain_fork) MAPF_CUDA(cudaEventCreateWithFlags(&env->chain_fork, cudaEventDisableTiming));
    for (int j = 0; j < S; ++j) {
        if (!env->chain_stream[j]) {
            MAPF_CUDA(cudaStreamCreateWithFlags(&env->chain_stream[j], cudaStreamNonBlocking));
            MAPF_CUDA(cudaEventCreateWithFlags(&env->chain_done[j], cudaEventDisableTiming));
        }
    }
    // fork: every chain starts after what the caller queued on `stream` (actions ready, earlier readers of the buffers done)
    MAPF_CUDA(cudaEventRecord(env->chain_fork, st));
    for (int j = 0; j < S; ++j) MAPF_CUDA(cudaStreamWaitEvent(env->chain_stream[j], env->chain_fork, 0));
    int rc = MAPF_OK;
    int t_begin = 0;
    if (graphs) {
        mapf_env::RolloutGraph *g = nullptr;
        const int gen = mapf_step_tuning_generation();
        for (auto &c : env->rg)
            if (c.exec[0] && c.act == d_actions && c.obs == d_obs && c.rew == d_rewards && c.done == d_done && c.steps == d_steps &&
                c.action_slots == action_slots && c.obs_slots == obs_slots && c.out_slots == out_slots && c.S == S && c.P == P &&
                c.tuning_gen == gen)
                g = &c;
        if (!g) {
            g = &env->rg[env->rg_next];
            env->rg_next = (env->rg_next + 1) % 2;
            for (auto &x : g->exec)
                if (x) {
                    cudaGraphExecDestroy(x);
                    x = nullptr;
                }
            if (!env->cap_stream) MAPF_CUDA(cudaStreamCreateWithFlags(&env->cap_stream, cudaStreamNonBlocking));
            for (int j = 0; j < S && rc == MAPF_OK; ++j) {
                const int e0 = j * per, e1 = e0 + per < d.B ? e0 + per : d.B;
                cudaGraph_t graph = nullptr;
                MAPF_CUDA(cudaStreamBeginCapture(env->cap_stream, cudaStreamCaptureModeThreadLocal));
                for (int t = 0; t < P && rc == MAPF_OK; ++t) rc = step_t(t, e0, e1, env->cap_stream);
                cudaError_t ce = cudaStreamEndCapture(env->cap_stream, &graph);
                if (rc == MAPF_OK && ce != cudaSuccess) rc = mapf_cuda_fail(ce, "cudaStreamEndCapture");
                if (rc == MAPF_OK) {
                    ce = cudaGraphInstantiate(&g->exec[j], graph, 0);
                    if (ce != cudaSuccess) rc = mapf_cuda_fail(ce, "cudaGraphInstantiate");
                }
                if (graph) cudaGraphDestroy(graph);
            }
            if (rc != MAPF_OK) {
                for (auto &x : g->exec)
                    if (x) {
                        cudaGraphExecDestroy(x);
                        x = nullptr;
                    }
            } else {
                g->act = d_actions, g->obs = d_obs, g->rew = d_rewards, g->done = d_done, g->steps = d_steps;
                g->action_slots = action_slots, g->obs_slots = obs_slots, g->out_slots = out_slots, g->S = S, g->P = P, g->tuning_gen = gen;
            }
        }
        // period-major issue order: every chain has a period of work queued before any gets its second one
        for (; rc == MAPF_OK && t_begin + P <= T; t_begin += P)
            for (int j = 0; j < S && rc == MAPF_OK; ++j) {
                const cudaError_t ce = cudaGraphLaunch(g->exec[j], env->chain_stream[j]);
                if (ce != cudaSuccess) rc = mapf_cuda_fail(ce, "cudaGraphLaunch");
            }
    }
    // step-major issue order, so that every chain has work queued from the start; chain j's launch t+1 follows its
    // launch t by stream order and nothing else
    for (int t = t_begin; t < T && rc == MAPF_OK; ++t)
        for (int j = 0; j < S && rc == MAPF_OK; ++j) {
            const int e0 = j * per, e1 = e0 + per < d.B ? e0 + per : d.B;
            rc = step_t(t, e0, e1, env->chain_stream[j]);
        }
    // join (also after a failed launch: the caller's stream must not run ahead of what was queued)
    for (int j = 0; j < S; ++j) {
        cudaError_t e = cudaEventRecord(env->chain_done[j], env->chain_stream[j]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, env->chain_done[j], 0);
        if (e != cudaSuccess && rc == MAPF_OK) rc = mapf_cuda_fail(e, "mapf_env_rollout join");
    }
    return rc;
}

int mapf_env_step_host(mapf_env *env, const uint8_t *h_actions, uint8_t *h_obs, float *h_rewards, uint8_t *h_done,
                       int32_t *h_steps, uint8_t *d_obs_opt, void *stream)
{
    REQUIRE_ENV(env);
    if (!h_actions || !h_rewards || !h_done) {
        mapf_set_error("mapf_env_step_host: NULL buffer");
        return MAPF_EINVAL;
    }
    const EnvDims &d = env->d;
    const size_t BN = (size_t)d.B * d.N;
    int64_t total = env->arena_bytes;
    int rc = MAPF_OK;
    if (!env->d_actions) {
        if (rc == MAPF_OK) rc = dev_alloc(&env->d_actions, BN, &total);
        if (rc == MAPF_OK) rc = dev_alloc(&env->d_rewards, BN, &total);
        if (rc == MAPF_OK) rc = dev_alloc(&env->d_done, (size_t)d.B, &total);
        if (rc == MAPF_OK) rc = dev_alloc(&env->d_steps_out, (size_t)d.B, &total);
        if (rc == MAPF_OK) {
            cudaError_t e = cudaMallocHost(reinterpret_cast<void **>(&env->h_pinned), BN + BN * 4 + (size_t)d.B * 5 + 64);
            if (e != cudaSuccess) rc = mapf_cuda_fail(e, "cudaMallocHost");
        }
    }
    if (rc == MAPF_OK && !d_obs_opt && !env->d_obs) rc = dev_alloc(&env->d_obs, BN * MAPF_OBS_BYTES_PER_AGENT, &total);
    env->arena_bytes = total;
    if (rc != MAPF_OK) return rc;

    cudaStream_t st = static_cast<cudaStream_t>(stream);
    uint8_t *obs_dev = d_obs_opt ? d_obs_opt : env->d_obs;
    // pinned layout: actions u8[BN] | pad to 16 | rewards f32[BN] | steps i32[B] | done u8[B]
    uint8_t *pin_act = env->h_pinned;
    float *pin_rew = reinterpret_cast<float *>(env->h_pinned + ((BN + 15) & ~(size_t)15));
    int32_t *pin_steps = reinterpret_cast<int32_t *>(pin_rew + BN);
    uint8_t *pin_done = reinterpret_cast<uint8_t *>(pin_steps + d.B);
    // page-locked caller buffers are used as DMA endpoints directly; pageable ones go through the handle's
    // pinned staging area (one extra host memcpy each way)
    const void *keys[4] = {h_actions, h_rewards, h_done, h_steps};
    for (int i = 0; i < 4; ++i) {
        if (keys[i] != env->hc_key[i]) {
            env->hc_key[i] = keys[i];
            // an actor that rotates over a few page-locked action buffers pays the two driver queries once per buffer
            mapf_env::PtrInfo *hit = nullptr;
            for (auto &c : env->ptr_cache)
                if (c.key == keys[i] && c.key) hit = &c;
            if (!hit) {
                hit = &env->ptr_cache[env->ptr_cache_next];
                env->ptr_cache_next = (env->ptr_cache_next + 1) % 32;
                hit->key = keys[i];
                hit->pinned = keys[i] && host_is_pinned(keys[i]);
                hit->alias = hit->pinned ? host_device_alias(const_cast<void *>(keys[i])) : nullptr;
            }
            env->hc_pinned[i] = hit->pinned;
            env->hc_alias[i] = hit->alias;
        }
    }
    const bool act_direct = env->hc_pinned[0];
    const bool out_direct = env->hc_pinned[1] && env->hc_pinned[2] && (!h_steps || env->hc_pinned[3]);
    const uint8_t *src_act = h_actions;
    if (!act_direct) {
        std::memcpy(pin_act, h_actions, BN);
        src_act = pin_act;
    }
    float *dst_rew = out_direct ? h_rewards : pin_rew;
    uint8_t *dst_done = out_direct ? h_done : pin_done;
    int32_t *dst_steps = out_direct ? h_steps : pin_steps;
    const int mode = step_host_mode();
    if (mode >= 3) {
        // split: step kernel -> {observe kernel on `st`  ||  result copies on the side stream} -> join
        if (!env->side_stream) {
            MAPF_CUDA(cudaStreamCreateWithFlags(&env->side_stream, cudaStreamNonBlocking));
            MAPF_CUDA(cudaEventCreateWithFlags(&env->ev_stepped, cudaEventDisableTiming));
            MAPF_CUDA(cudaEventCreateWithFlags(&env->ev_copied, cudaEventDisableTiming));
        }
        static const bool dma_actions = std::getenv("MAPF_STEP_HOST_DMA_ACTIONS") != nullptr;  // probe: H2D copy node instead
        const uint8_t *act_dev = act_direct && !dma_actions ? static_cast<const uint8_t *>(env->hc_alias[0]) : nullptr;  // read in place over PCIe
        // issues the whole sequence on `q` (forking to the side stream and joining back)
        // MAPF_STEP_HOST_TRACE=1 (mode 3 only): timing events between the phases, printed to stderr after the sync
        static const bool trace = std::getenv("MAPF_STEP_HOST_TRACE") != nullptr;
        static cudaEvent_t tev[5] = {};
        const bool tracing = trace && mode == 3;
        if (tracing && !tev[0])
            for (auto &e : tev) MAPF_CUDA(cudaEventCreate(&e));
        auto enqueue = [&](cudaStream_t q) -> int {
            if (tracing) MAPF_CUDA(cudaEventRecord(tev[0], q));
            const uint8_t *a = act_dev;
            if (!a) {
                MAPF_CUDA(cudaMemcpyAsync(env->d_actions, src_act, BN, cudaMemcpyHostToDevice, q));
                a = env->d_actions;
            }
            int r = mapf_launch_step_only(env, a, env->d_rewards, env->d_done, env->d_steps_out, q);
            if (r != MAPF_OK) return r;
            if (tracing) MAPF_CUDA(cudaEventRecord(tev[1], q));
            MAPF_CUDA(cudaEventRecord(env->ev_stepped, q));
            MAPF_CUDA(cudaStreamWaitEvent(env->side_stream, env->ev_stepped, 0));
            MAPF_CUDA(cudaMemcpyAsync(dst_rew, env->d_rewards, BN * 4, cudaMemcpyDeviceToHost, env->side_stream));
            if (dst_steps)
                MAPF_CUDA(cudaMemcpyAsync(dst_steps, env->d_steps_out, (size_t)d.B * 4, cudaMemcpyDeviceToHost, env->side_stream));
            MAPF_CUDA(cudaMemcpyAsync(dst_done, env->d_done, (size_t)d.B, cudaMemcpyDeviceToHost, env->side_stream));
            if (tracing) MAPF_CUDA(cudaEventRecord(tev[2], env->side_stream));
            MAPF_CUDA(cudaEventRecord(env->ev_copied, env->side_stream));
            r = mapf_launch_observe(env, obs_dev, nullptr, nullptr, q);
            if (r != MAPF_OK) return r;
            if (tracing) MAPF_CUDA(cudaEventRecord(tev[3], q));
            if (h_obs) MAPF_CUDA(cudaMemcpyAsync(h_obs, obs_dev, BN * MAPF_OBS_BYTES_PER_AGENT, cudaMemcpyDeviceToHost, q));
            MAPF_CUDA(cudaStreamWaitEvent(q, env->ev_copied, 0));
            if (tracing) MAPF_CUDA(cudaEventRecord(tev[4], q));
            return MAPF_OK;
        };
        if (mode >= 4) {
            // one cudaGraphLaunch instead of nine stream calls: the sequence is captured per buffer set
            mapf_env::HostGraph *g = nullptr;
            for (auto &c : env->hg)
                if (c.exec && c.act == src_act && c.rew == dst_rew && c.done == dst_done && c.steps == dst_steps && c.hobs == h_obs &&
                    c.obs_dev == obs_dev && c.mode == mode)
                    g = &c;
            if (!g) {
                if (!env->cap_stream) MAPF_CUDA(cudaStreamCreateWithFlags(&env->cap_stream, cudaStreamNonBlocking));
                g = &env->hg[env->hg_next];
                env->hg_next = (env->hg_next + 1) % 32;
                if (g->exec) {
                    cudaGraphExecDestroy(g->exec);
                    g->exec = nullptr;
                }
                cudaGraph_t graph = nullptr;
                MAPF_CUDA(cudaStreamBeginCapture(env->cap_stream, cudaStreamCaptureModeThreadLocal));
                rc = enqueue(env->cap_stream);
                cudaError_t ce = cudaStreamEndCapture(env->cap_stream, &graph);
                if (rc != MAPF_OK) {
                    if (graph) cudaGraphDestroy(graph);
                    return rc;
                }
                if (ce != cudaSuccess) return mapf_cuda_fail(ce, "cudaStreamEndCapture");
                ce = cudaGraphInstantiate(&g->exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ce != cudaSuccess) return mapf_cuda_fail(ce, "cudaGraphInstantiate");
                g->act = src_act, g->rew = dst_rew, g->done = dst_done, g->steps = dst_steps, g->hobs = h_obs, g->obs_dev = obs_dev, g->mode = mode;
            }
            MAPF_CUDA(cudaGraphLaunch(g->exec, st));
        } else {
            rc = enqueue(st);
            if (rc != MAPF_OK) return rc;
        }
        MAPF_CUDA(cudaStreamSynchronize(st));
        if (tracing) {
            float t[4] = {};
            for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], tev[0], tev[i + 1]);
            std::fprintf(stderr, "step_host trace (us since start): step kernel done %.1f, copies done %.1f, observe done %.1f, joined %.1f\n",
                         t[0] * 1e3f, t[1] * 1e3f, t[2] * 1e3f, t[3] * 1e3f);
        }
        if (!out_direct) {
            std::memcpy(h_rewards, pin_rew, BN * 4);
            std::memcpy(h_done, pin_done, (size_t)d.B);
            if (h_steps) std::memcpy(h_steps, pin_steps, (size_t)d.B * 4);
        }
        return MAPF_OK;
    }
    // zero-copy: the kernel's own loads / stores reach the host buffers (no DMA launches around the kernel)
    float *zc_rew = nullptr;
    uint8_t *zc_done = nullptr;
    int32_t *zc_steps = nullptr;
    const uint8_t *zc_act = nullptr;
    if (mode >= 1) {
        if (out_direct) {
            zc_rew = static_cast<float *>(env->hc_alias[1]);
            zc_done = static_cast<uint8_t *>(env->hc_alias[2]);
            zc_steps = static_cast<int32_t *>(env->hc_alias[3]);
        } else {
            if (!env->pin_alias) env->pin_alias = host_device_alias(env->h_pinned);
            if (env->pin_alias) {
                zc_rew = reinterpret_cast<float *>(env->pin_alias + (reinterpret_cast<uint8_t *>(pin_rew) - env->h_pinned));
                zc_done = env->pin_alias + (pin_done - env->h_pinned);
                zc_steps = reinterpret_cast<int32_t *>(env->pin_alias + (reinterpret_cast<uint8_t *>(pin_steps) - env->h_pinned));
            }
        }
        if (mode >= 2) {
            if (act_direct) zc_act = static_cast<const uint8_t *>(env->hc_alias[0]);
            else if (env->pin_alias) zc_act = env->pin_alias;
        }
    }
    if (zc_rew && zc_done && (zc_steps || !dst_steps)) {
        if (!zc_act) MAPF_CUDA(cudaMemcpyAsync(env->d_actions, src_act, BN, cudaMemcpyHostToDevice, st));
        rc = mapf_launch_step(env, zc_act ? zc_act : env->d_actions, obs_dev, nullptr, zc_rew, zc_done, dst_steps ? zc_steps : nullptr, st);
        if (rc != MAPF_OK) return rc;
        if (h_obs) MAPF_CUDA(cudaMemcpyAsync(h_obs, obs_dev, BN * MAPF_OBS_BYTES_PER_AGENT, cudaMemcpyDeviceToHost, st));
        MAPF_CUDA(cudaStreamSynchronize(st));
        if (!out_direct) {
            std::memcpy(h_rewards, pin_rew, BN * 4);
            std::memcpy(h_done, pin_done, (size_t)d.B);
            if (h_steps) std::memcpy(h_steps, pin_steps, (size_t)d.B * 4);
        }
        return MAPF_OK;
    }
    MAPF_CUDA(cudaMemcpyAsync(env->d_actions, src_act, BN, cudaMemcpyHostToDevice, st));
    rc = mapf_launch_step(env, env->d_actions, obs_dev, nullptr, env->d_rewards, env->d_done, env->d_steps_out, st);
    if (rc != MAPF_OK) return rc;
    MAPF_CUDA(cudaMemcpyAsync(dst_rew, env->d_rewards, BN * 4, cudaMemcpyDeviceToHost, st));
    if (dst_steps) MAPF_CUDA(cudaMemcpyAsync(dst_steps, env->d_steps_out, (size_t)d.B * 4, cudaMemcpyDeviceToHost, st));
    MAPF_CUDA(cudaMemcpyAsync(dst_done, env->d_done, (size_t)d.B, cudaMemcpyDeviceToHost, st));
    if (h_obs) MAPF_CUDA(cudaMemcpyAsync(h_obs, obs_dev, BN * MAPF_OBS_BYTES_PER_AGENT, cudaMemcpyDeviceToHost, st));
    MAPF_CUDA(cudaStreamSynchronize(st));
    if (!out_direct) {
        std::memcpy(h_rewards, pin_rew, BN * 4);
        std::memcpy(h_done, pin_done, (size_t)d.B);
        if (h_steps) std::memcpy(h_steps, pin_steps, (size_t)d.B * 4);
    }
    return MAPF_OK;
}

int mapf_debug_step_tuning(int32_t variant, int32_t flags, int32_t ctas_per_sm)
{
    mapf_set_step_tuning(variant, flags, ctas_per_sm);
    return MAPF_OK;
}

int mapf_debug_rollout_tuning(int32_t persistent, int32_t envs_per_warp, int32_t cta_warps)
{
    RolloutTuning &t = rollout_tuning();
    if (persistent >= 0) t.persistent = persistent;
    if (envs_per_warp >= 0) t.envs_per_warp = envs_per_warp;
    if (cta_warps >= 0) t.cta_warps = cta_warps;
    return MAPF_OK;
}

int mapf_debug_step_host_mode(int32_t mode)
{
    if (mode >= 0) step_host_mode_ref() = mode;
    return step_host_mode_ref();
}

int mapf_debug_step_trace(uint64_t *d_trace)
{
    mapf_set_step_trace(reinterpret_cast<unsigned long long *>(d_trace));
    return MAPF_OK;
}

int mapf_env_comm_mask(mapf_env *env, int32_t max_comm_agents, uint8_t *d_mask_out, void *stream)
{
    REQUIRE_ENV(env);
    if (!d_mask_out || max_comm_agents < 1 || max_comm_agents > 3) {
        mapf_set_error("mapf_env_comm_mask: NULL buffer or max_comm_agents outside 1..3 (config.py:58)");
        return MAPF_EINVAL;
    }
    return mapf_launch_comm_mask(env, max_comm_agents, d_mask_out, static_cast<cudaStream_t>(stream));
}

int mapf_env_get_state(mapf_env *env, uint8_t *d_map, uint8_t *d_pos, uint8_t *d_goals, int32_t *d_steps, uint8_t *d_navi,
                       void *stream)
{
    REQUIRE_ENV(env);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t BN2 = (size_t)env->d.B * env->d.N * 2;
    if (d_pos) MAPF_CUDA(cudaMemcpyAsync(d_pos, env->pos, BN2, cudaMemcpyDeviceToDevice, st));
    if (d_goals) MAPF_CUDA(cudaMemcpyAsync(d_goals, env->goal, BN2, cudaMemcpyDeviceToDevice, st));
    if (d_steps) MAPF_CUDA(cudaMemcpyAsync(d_steps, env->steps, (size_t)env->d.B * 4, cudaMemcpyDeviceToDevice, st));
    return mapf_launch_unpack(env, d_map, d_navi, st);
}

int mapf_env_set_state(mapf_env *env, const uint8_t *d_pos, const int32_t *d_steps, void *stream)
{
    REQUIRE_ENV(env);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d_pos) MAPF_CUDA(cudaMemcpyAsync(env->pos, d_pos, (size_t)env->d.B * env->d.N * 2, cudaMemcpyDeviceToDevice, st));
    if (d_steps) MAPF_CUDA(cudaMemcpyAsync(env->steps, d_steps, (size_t)env->d.B * 4, cudaMemcpyDeviceToDevice, st));
    return MAPF_OK;
}

int mapf_env_status(mapf_env *env, void *stream)
{
    REQUIRE_ENV(env);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int32_t bits = 0;
    MAPF_CUDA(cudaMemcpyAsync(&bits, env->err, 4, cudaMemcpyDeviceToHost, st));
    MAPF_CUDA(cudaStreamSynchronize(st));
    if (bits) {
        MAPF_CUDA(cudaMemsetAsync(env->err, 0, 4, st));
        MAPF_CUDA(cudaStreamSynchronize(st));
    }
    if (bits & MAPF_ERRBIT_ACTION) {
        mapf_set_error("action index out of range");
        return MAPF_EACTION;
    }
    if (bits & MAPF_ERRBIT_UNIQUE) {
        mapf_set_error("unique");
        return MAPF_EUNIQUE;
    }
    if (bits & MAPF_ERRBIT_RESET) {
        mapf_set_error("no empty position");  // environment.py:31
        return MAPF_ENOSPACE;
    }
    return MAPF_OK;
}

int mapf_env_reset(mapf_env *env, const uint8_t *d_mask, uint64_t seed, uint64_t env_offset, float density, void *stream)
{
    REQUIRE_ENV(env);
    if (density >= 1.0f) {
        mapf_set_error("mapf_env_reset: density must be < 1");
        return MAPF_EINVAL;
    }
    return mapf_launch_reset(env, d_mask, seed, env_offset, density, static_cast<cudaStream_t>(stream));
}

/* ---------------------------------------------------------------------------------------------- */

int mapf_per_create(int64_t capacity, int32_t device, mapf_per **out)
{
    if (!out || capacity < 1 || (capacity & (capacity - 1)) != 0) {
        mapf_set_error("buffer size only support power of 2 size");  // buffer.py:23
        return MAPF_EINVAL;
    }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        mapf_set_error("mapf_per_create: no CUDA device (this library has no CPU fallback)");
        return MAPF_ECUDA;
    }
    if (device < 0 || device >= ndev) {
        mapf_set_error("mapf_per_create: bad device ordinal");
        return MAPF_EINVAL;
    }
    DeviceGuard guard(device);
    mapf_per *t = new (std::nothrow) mapf_per();
    if (!t) return MAPF_ENOMEM;
    std::memset(t, 0, sizeof(*t));
    t->capacity = capacity;
    t->device = device;
    int layer = 1;
    while ((int64_t(1) << (layer - 1)) < capacity) ++layer;  // buffer.py:20-22
    t->layer = layer;
    int64_t total = 0;
    const int64_t nodes = 2 * capacity - 1;
    t->scratch.cap_n = 1 << 16;
    int rc = dev_alloc(&t->tree, (size_t)nodes, &total);
    if (rc == MAPF_OK) rc = dev_alloc(&t->scratch.stamps, (size_t)capacity, &total);
    if (rc == MAPF_OK) rc = dev_alloc(&t->scratch.prio32, (size_t)t->scratch.cap_n, &total);
    if (rc == MAPF_OK) rc = dev_alloc(&t->scratch.active, (size_t)t->scratch.cap_n, &total);
    if (rc == MAPF_OK) {
        cudaError_t e = cudaMemset(t->tree, 0, (size_t)nodes * 8);
        if (e == cudaSuccess) e = cudaMemset(t->scratch.stamps, 0, (size_t)capacity * 8);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = mapf_cuda_fail(e, "cudaMemset(tree)");
    }
    if (rc != MAPF_OK) {
        mapf_per_destroy(t);
        return rc;
    }
    *out = t;
    return MAPF_OK;
}

int mapf_per_destroy(mapf_per *t)
{
    if (!t) return MAPF_OK;
    DeviceGuard guard(t->device);
    cudaFree(t->tree);
    cudaFree(t->scratch.stamps);
    cudaFree(t->scratch.prio32);
    cudaFree(t->scratch.active);
    delete t;
    return MAPF_OK;
}

double *mapf_per_tree_ptr(mapf_per *t) { return t ? t->tree : nullptr; }

#define REQUIRE_PER(t)                          \
    if (!(t)) {                                 \
        mapf_set_error("NULL sum-tree handle"); \
        return MAPF_EINVAL;                     \
    }                                           \
    DeviceGuard guard((t)->device);             \
    if (!guard.ok) {                            \
        mapf_set_error("cudaSetDevice failed"); \
        return MAPF_ECUDA;                      \
    }

int mapf_per_update(mapf_per *t, const int64_t *d_idx, const double *d_prio, int64_t n, void *stream)
{
    REQUIRE_PER(t);
    if (n < 0 || (n > 0 && (!d_idx || !d_prio))) {
        mapf_set_error("mapf_per_update: bad arguments");
        return MAPF_EINVAL;
    }
    return mapf_launch_per_update(t, &t->scratch, d_idx, d_prio, n, static_cast<cudaStream_t>(stream));
}

int mapf_per_sample(mapf_per *t, const double *d_uniforms, int64_t batch, int64_t *d_idx_out, double *d_prio_out,
                    float *d_weight_out, double beta, void *stream)
{
    REQUIRE_PER(t);
    if (batch < 1 || !d_uniforms || !d_idx_out || !d_prio_out) {
        mapf_set_error("mapf_per_sample: bad arguments");
        return MAPF_EINVAL;
    }
    return mapf_launch_per_sample(t, d_uniforms, batch, d_idx_out, d_prio_out, d_weight_out, beta,
                                  static_cast<cudaStream_t>(stream));
}

int mapf_per_td_update(mapf_per *t, const float *d_q_online, const float *d_q_target_next, const float *d_q_online_next,
                       const int64_t *d_action, const float *d_reward, const float *d_done, const float *d_steps,
                       const int64_t *d_idx, int64_t n, float gamma, double alpha, int64_t old_ptr, int64_t ptr,
                       int64_t slot_steps, float *d_td_out, float *d_prio_out, void *stream)
{
    REQUIRE_PER(t);
    if (n < 0 || n > t->scratch.cap_n || (n > 0 && (!d_q_online || !d_q_target_next || !d_action || !d_reward || !d_done ||
                                                    !d_steps || !d_idx))) {
        mapf_set_error("mapf_per_td_update: bad arguments (n <= 65536)");
        return MAPF_EINVAL;
    }
    return mapf_launch_per_td_update(t, &t->scratch, d_q_online, d_q_target_next, d_q_online_next, d_action, d_reward, d_done,
                                     d_steps, d_idx, n, gamma, alpha, old_ptr, ptr, slot_steps, d_td_out, d_prio_out,
                                     static_cast<cudaStream_t>(stream));
}

int mapf_actor_td(const float *d_rew, const float *d_q, const uint8_t *d_act, const int32_t *d_size, int32_t episodes,
                  int32_t capacity, double *d_td_out, void *stream)
{
    if (episodes < 0 || capacity < 1 || (episodes > 0 && (!d_rew || !d_q || !d_act || !d_size || !d_td_out))) {
        mapf_set_error("mapf_actor_td: bad arguments");
        return MAPF_EINVAL;
    }
    return mapf_launch_actor_td(d_rew, d_q, d_act, d_size, episodes, capacity, d_td_out, static_cast<cudaStream_t>(stream));
}

int mapf_replay_gather(const mapf_replay_view *view, const int64_t *d_idx, int64_t batch, const mapf_replay_batch *out,
                       int32_t *d_err, void *stream)
{
    if (!view || !out || batch < 0 || (batch > 0 && !d_idx)) {
        mapf_set_error("mapf_replay_gather: NULL argument");
        return MAPF_EINVAL;
    }
    if (view->num_agents < 1 || view->num_agents > MAPF_MAX_AGENTS || view->max_steps < 1 || view->bt_steps < 1 ||
        view->forward_steps < 1 || view->latent_dim < 1 || batch > 65535) {
        mapf_set_error("mapf_replay_gather: bad dimensions (batch <= 65535)");
        return MAPF_EINVAL;
    }
    if (!view->obs_buf || !view->comm_buf || !view->hid_buf || !view->act_buf || !view->rew_buf || !view->done_buf ||
        !view->size_buf || !out->obs || !out->comm_mask || !out->hidden || !out->action || !out->reward || !out->done ||
        !out->steps || !out->bt_steps) {
        mapf_set_error("mapf_replay_gather: NULL buffer");
        return MAPF_EINVAL;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        mapf_set_error("mapf_replay_gather: no CUDA device (this library has no CPU fallback)");
        return MAPF_ECUDA;
    }
    if (batch == 0) return MAPF_OK;
    return mapf_launch_replay_gather(view, d_idx, batch, out, d_err, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
