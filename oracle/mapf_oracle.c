/*
 * mapf_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's hot path (ZiyuanMa/MAPF_RL), written to be
 * the checker for the CUDA kernels in mapf_rl_b200/csrc.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.  The product
 * path never links, imports or calls it.
 *
 * It deliberately follows the reference's *sequential* formulation (checking list, restart-on-
 * change scans) rather than the order-independent fixed point the CUDA kernel implements, so
 * that agreement between the two is evidence and not a tautology.
 *
 * Parity pinned: tests/test_oracle_vs_reference.py runs this file against the live reference
 * (oracle/ref_loader.py) in the dev container, and tests/golden/ holds vectors generated from the
 * live reference by tests/golden/make_golden.py.
 *
 * Reference citations are file:line into /root/reference.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define MO_INF 2147483647

/* environment.py:12  action id -> (dx, dy); x = row (axis 0), y = col */
static const int MO_DX[5] = {0, -1, 1, 0, 0};
static const int MO_DY[5] = {0, 0, 0, -1, 1};

/* reward_fn order used everywhere in this repo: move, stay_on_goal, stay_off_goal, collision, finish
 * (config.py:8-12) */
enum { R_MOVE = 0, R_STAY_ON = 1, R_STAY_OFF = 2, R_COLLISION = 3, R_FINISH = 4 };

/* ------------------------------------------------------------------------------------------ */
/* Environment.step — environment.py:278-430                                                   */
/* map: u8[L*L], 1 = obstacle.  pos/goals: int32[N*2] (x,y).  actions: u8[N].                   */
/* rewards: double[N] (the reference returns python numbers; callers cast to fp32).             */
/* returns done (0/1), or -1 on invalid action (AssertionError at :289-290),                    */
/*         -2 if the post-step uniqueness check (:424-428) fails.                               */
/* ------------------------------------------------------------------------------------------ */
static int list_index(const int *list, int n, int v)
{
    for (int i = 0; i < n; ++i)
        if (list[i] == v) return i;
    return -1;
}
static void list_remove(int *list, int *n, int v)
{
    int k = list_index(list, *n, v);
    if (k < 0) return;
    memmove(list + k, list + k + 1, (size_t)(*n - k - 1) * sizeof(int));
    --*n;
}

int mo_step(int L, int N, const uint8_t *map, int32_t *pos, const int32_t *goals,
            const uint8_t *actions, const double *reward_fn, double *rewards)
{
    for (int i = 0; i < N; ++i)
        if (actions[i] > 4) return -1; /* :290 */

    int *checking = (int *)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    int32_t *next = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(N > 0 ? N : 1));
    int *group = (int *)malloc(sizeof(int) * (size_t)(N > 0 ? N : 1));
    int nc = 0;
    memcpy(next, pos, sizeof(int32_t) * 2 * (size_t)N); /* :295 */

    /* :298-311 stay / move pass */
    for (int i = 0; i < N; ++i) {
        if (actions[i] == 0) {
            int on_goal = pos[2 * i] == goals[2 * i] && pos[2 * i + 1] == goals[2 * i + 1];
            rewards[i] = on_goal ? reward_fn[R_STAY_ON] : reward_fn[R_STAY_OFF];
        } else {
            next[2 * i] += MO_DX[actions[i]];
            next[2 * i + 1] += MO_DY[actions[i]];
            rewards[i] = reward_fn[R_MOVE];
            checking[nc++] = i;
        }
    }

    /* :320-332 first round: out of range (side 0 used for both axes, :322) / obstacle */
    {
        int snapshot_n = nc;
        int *snapshot = (int *)malloc(sizeof(int) * (size_t)(nc > 0 ? nc : 1));
        memcpy(snapshot, checking, sizeof(int) * (size_t)nc);
        for (int k = 0; k < snapshot_n; ++k) {
            int a = snapshot[k];
            int x = next[2 * a], y = next[2 * a + 1];
            if (x < 0 || y < 0 || x >= L || y >= L || map[x * L + y] == 1) {
                rewards[a] = reward_fn[R_COLLISION];
                next[2 * a] = pos[2 * a];
                next[2 * a + 1] = pos[2 * a + 1];
                list_remove(checking, &nc, a);
            }
        }
        free(snapshot);
    }

    /* :335-365 second round: swapping conflict, restart scan after every hit */
    {
        int all_good = 0;
        while (!all_good) {
            all_good = 1;
            for (int k = 0; k < nc; ++k) {
                int a = checking[k];
                int t = -1; /* :341 agent currently standing on a's target */
                for (int j = 0; j < N; ++j)
                    if (pos[2 * j] == next[2 * a] && pos[2 * j + 1] == next[2 * a + 1]) { t = j; break; }
                /* :343 `if target_agent_id:` — ndarray truthiness: empty -> False, [0] -> False */
                if (t > 0) {
                    if (next[2 * t] == pos[2 * a] && next[2 * t + 1] == pos[2 * a + 1]) { /* :353 */
                        next[2 * a] = pos[2 * a];
                        next[2 * a + 1] = pos[2 * a + 1];
                        rewards[a] = reward_fn[R_COLLISION];
                        next[2 * t] = pos[2 * t];
                        next[2 * t + 1] = pos[2 * t + 1];
                        rewards[t] = reward_fn[R_COLLISION];
                        list_remove(checking, &nc, a);
                        list_remove(checking, &nc, t);
                        all_good = 0;
                        break;
                    }
                }
            }
        }
    }

    /* :369-406 third round: vertex conflict, restart scan after every hit */
    {
        int all_good = 0;
        while (!all_good) {
            all_good = 1;
            for (int k = 0; k < nc; ++k) {
                int a = checking[k];
                int ng = 0; /* :375 everybody whose next_pos equals a's, ascending id */
                for (int j = 0; j < N; ++j)
                    if (next[2 * j] == next[2 * a] && next[2 * j + 1] == next[2 * a + 1]) group[ng++] = j;
                if (ng > 1) {
                    int all_in_checking = 1; /* :380-384 */
                    int m = 0;
                    for (int q = 0; q < ng; ++q) {
                        if (list_index(checking, nc, group[q]) < 0) all_in_checking = 0;
                        else group[m++] = group[q];
                    }
                    ng = m;
                    if (all_in_checking) {
                        /* :389-394 stable sort on a key that is identical for every member
                         * (they all share next_pos) => first == lowest id keeps its move */
                        memmove(group, group + 1, sizeof(int) * (size_t)(ng - 1));
                        --ng;
                    }
                    for (int q = 0; q < ng; ++q) { /* :398-403 */
                        int g = group[q];
                        next[2 * g] = pos[2 * g];
                        next[2 * g + 1] = pos[2 * g + 1];
                        rewards[g] = reward_fn[R_COLLISION];
                    }
                    for (int q = 0; q < ng; ++q) list_remove(checking, &nc, group[q]);
                    all_good = 0;
                    break;
                }
            }
        }
    }

    memcpy(pos, next, sizeof(int32_t) * 2 * (size_t)N); /* :410 */

    int done = 1; /* :415-419 */
    for (int i = 0; i < N; ++i)
        if (pos[2 * i] != goals[2 * i] || pos[2 * i + 1] != goals[2 * i + 1]) { done = 0; break; }
    if (done)
        for (int i = 0; i < N; ++i) rewards[i] = reward_fn[R_FINISH];

    int unique_ok = 1; /* :424-428 */
    for (int i = 0; i < N && unique_ok; ++i)
        for (int j = i + 1; j < N; ++j)
            if (pos[2 * i] == pos[2 * j] && pos[2 * i + 1] == pos[2 * j + 1]) { unique_ok = 0; break; }

    free(checking);
    free(next);
    free(group);
    return unique_ok ? done : -2;
}

/* ------------------------------------------------------------------------------------------ */
/* Environment.get_navi_map — environment.py:217-276 (distances identical to                  */
/* search.compute_heuristics, search.py:24-55, where finite).                                  */
/* dist: int32[N*L*L] (INT32_MAX unreachable / obstacle), navi: u8[N*4*L*L] *unpadded*.        */
/* ------------------------------------------------------------------------------------------ */
void mo_navi(int L, int N, const uint8_t *map, const int32_t *goals, int32_t *dist, uint8_t *navi)
{
    int cells = L * L;
    int32_t *queue = (int32_t *)malloc(sizeof(int32_t) * (size_t)cells * 4 + 16);
    for (int i = 0; i < N; ++i) {
        int32_t *d = dist + (size_t)i * cells;
        for (int c = 0; c < cells; ++c) d[c] = MO_INF; /* :218 */
        int head = 0, tail = 0;
        int gx = goals[2 * i], gy = goals[2 * i + 1];
        d[gx * L + gy] = 0; /* :223 */
        queue[tail++] = gx * L + gy;
        /* :225-251 FIFO relaxation; with unit costs the first label is final, so the
         * `not in open_list` test never changes the result. */
        while (head < tail) {
            int c = queue[head++];
            int x = c / L, y = c % L, dd = d[c];
            if (x - 1 >= 0 && map[(x - 1) * L + y] == 0 && d[(x - 1) * L + y] > dd + 1) { d[(x - 1) * L + y] = dd + 1; queue[tail++] = (x - 1) * L + y; }
            if (x + 1 < L && map[(x + 1) * L + y] == 0 && d[(x + 1) * L + y] > dd + 1) { d[(x + 1) * L + y] = dd + 1; queue[tail++] = (x + 1) * L + y; }
            if (y - 1 >= 0 && map[x * L + y - 1] == 0 && d[x * L + y - 1] > dd + 1) { d[x * L + y - 1] = dd + 1; queue[tail++] = x * L + y - 1; }
            if (y + 1 < L && map[x * L + y + 1] == 0 && d[x * L + y + 1] > dd + 1) { d[x * L + y + 1] = dd + 1; queue[tail++] = x * L + y + 1; }
        }
        if (navi) {
            uint8_t *nv = navi + (size_t)i * 4 * cells;
            memset(nv, 0, (size_t)4 * cells);
            for (int x = 0; x < L; ++x)
                for (int y = 0; y < L; ++y) {
                    if (map[x * L + y] != 0) continue; /* :257 */
                    int32_t here = d[x * L + y];
                    if (x > 0 && d[(x - 1) * L + y] < here) nv[0 * cells + x * L + y] = 1;     /* :260 */
                    if (x < L - 1 && d[(x + 1) * L + y] < here) nv[1 * cells + x * L + y] = 1; /* :264 */
                    if (y > 0 && d[x * L + y - 1] < here) nv[2 * cells + x * L + y] = 1;       /* :268 */
                    if (y < L - 1 && d[x * L + y + 1] < here) nv[3 * cells + x * L + y] = 1;   /* :272 */
                }
        }
    }
    free(queue);
}

/* ------------------------------------------------------------------------------------------ */
/* Environment.observe — environment.py:433-467.  obs: u8[N*6*F*F], F = 2r+1.                  */
/* ch0 other agents (own centre cleared :461), ch1 obstacle (1 = obstacle, 0 outside, :447),   */
/* ch2..5 navi (0 outside, :276,:465).                                                         */
/* ------------------------------------------------------------------------------------------ */
void mo_observe(int L, int N, int r, const uint8_t *map, const int32_t *pos, const uint8_t *navi, uint8_t *obs)
{
    int F = 2 * r + 1, cells = L * L;
    uint8_t *agent_map = (uint8_t *)calloc((size_t)cells, 1);
    for (int i = 0; i < N; ++i) agent_map[pos[2 * i] * L + pos[2 * i + 1]] = 1; /* :449-450 */
    memset(obs, 0, (size_t)N * 6 * F * F);
    for (int i = 0; i < N; ++i) {
        uint8_t *o = obs + (size_t)i * 6 * F * F;
        int x0 = pos[2 * i] - r, y0 = pos[2 * i + 1] - r;
        for (int u = 0; u < F; ++u)
            for (int v = 0; v < F; ++v) {
                int x = x0 + u, y = y0 + v;
                if (x < 0 || y < 0 || x >= L || y >= L) continue;
                int c = x * L + y;
                o[0 * F * F + u * F + v] = agent_map[c];
                o[1 * F * F + u * F + v] = map[c] == 1;
                for (int d = 0; d < 4; ++d) o[(2 + d) * F * F + u * F + v] = navi[((size_t)i * 4 + d) * cells + c];
            }
        o[0 * F * F + r * F + r] = 0; /* :461 */
    }
    free(agent_map);
}

/* Batched lockstep driver used by tests and by bench.py's cpu_baseline leg:               */
/* B independent envs (same L, N), T steps, actions u8[T*B*N]; obs of the last step kept.  */
/* Returns the number of agent-steps executed.  Parallel over envs when built with OpenMP  */
/* (each env is independent, exactly like one reference Environment per actor process).    */
#ifdef _OPENMP
#include <omp.h>
#endif
/* mo_navi over a batch (set-up of the batched CPU baseline): maps u8[B,L,L], goals i32[B,N,2] -> navi u8[B,N,4,L,L] */
void mo_navi_batch(int B, int L, int N, const uint8_t *maps, const int32_t *goals, uint8_t *navi)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8)
#endif
    for (int b = 0; b < B; ++b) {
        int32_t *dist = (int32_t *)malloc(sizeof(int32_t) * (size_t)N * L * L);
        mo_navi(L, N, maps + (size_t)b * L * L, goals + (size_t)b * N * 2, dist, navi + (size_t)b * N * 4 * L * L);
        free(dist);
    }
}

/* host threads used by mo_rollout (the batched CPU baseline); returns the count in effect */
int mo_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

long mo_rollout(int B, int L, int N, int r, int T, const uint8_t *maps, int32_t *pos, const int32_t *goals,
                const uint8_t *navi, const uint8_t *actions, const double *reward_fn,
                float *rewards_out, uint8_t *done_out, uint8_t *obs_out)
{
    int F = 2 * r + 1;
    size_t obs_sz = (size_t)N * 6 * F * F;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int b = 0; b < B; ++b) {
        double *rw = (double *)malloc(sizeof(double) * (size_t)N);
        for (int t = 0; t < T; ++t) {
            int d = mo_step(L, N, maps + (size_t)b * L * L, pos + (size_t)b * N * 2, goals + (size_t)b * N * 2,
                            actions + ((size_t)t * B + b) * N, reward_fn, rw);
            mo_observe(L, N, r, maps + (size_t)b * L * L, pos + (size_t)b * N * 2,
                       navi + (size_t)b * N * 4 * L * L, obs_out + (size_t)b * obs_sz);
            if (rewards_out)
                for (int i = 0; i < N; ++i) rewards_out[((size_t)t * B + b) * N + i] = (float)rw[i];
            if (done_out) done_out[(size_t)t * B + b] = (uint8_t)(d == 1);
        }
        free(rw);
    }
    return (long)B * N * T;
}

/* ------------------------------------------------------------------------------------------ */
/* SumTree — buffer.py:16-105.  tree: double[2*capacity-1], leaves at [capacity-1, ...).       */
/* ------------------------------------------------------------------------------------------ */
static int cmp_i64(const void *a, const void *b)
{
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* buffer.py:95-105.  idx is mutated in place exactly like the reference (:96); on return it holds
 * the leaf *node* indices (idx + capacity - 1), as numpy leaves them. */
void mo_tree_batch_update(double *tree, int64_t capacity, int layer, int64_t *idx, const double *prio, int64_t n)
{
    for (int64_t k = 0; k < n; ++k) idx[k] += capacity - 1;    /* :96 */
    for (int64_t k = 0; k < n; ++k) tree[idx[k]] = prio[k];      /* :97 fancy assignment: last duplicate wins */
    int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    memcpy(cur, idx, sizeof(int64_t) * (size_t)n);
    int64_t m = n;
    for (int l = 0; l < layer - 1; ++l) {                      /* :99-102 */
        for (int64_t k = 0; k < m; ++k) cur[k] = (cur[k] - 1) / 2; /* floor div; cur[k] >= 1 here */
        qsort(cur, (size_t)m, sizeof(int64_t), cmp_i64);        /* np.unique */
        int64_t u = 0;
        for (int64_t k = 0; k < m; ++k)
            if (k == 0 || cur[k] != cur[k - 1]) cur[u++] = cur[k];
        m = u;
        for (int64_t k = 0; k < m; ++k) tree[cur[k]] = tree[2 * cur[k] + 1] + tree[2 * cur[k] + 2];
    }
    free(cur);
}

/* buffer.py:56-78 with the uniforms supplied by the caller: prefix_i = i*interval + u_i*interval
 * (np.arange(0,sum,interval)[i] == i*interval and np.random.uniform(0,interval) == u*interval,
 * SURVEY A.8). */
void mo_tree_batch_sample(const double *tree, int64_t capacity, int layer, const double *uniforms, int64_t B,
                          int64_t *idx_out, double *prio_out)
{
    double sum = tree[0];
    double interval = sum / (double)B; /* :58 */
    for (int64_t i = 0; i < B; ++i) {
        double p = (double)i * interval + uniforms[i] * interval; /* :60 */
        if (i == 0 && p == 0.0) p = 1e-5;                        /* :61-62 */
        int64_t node = 0;
        for (int l = 0; l < layer - 1; ++l) {                   /* :66-70 */
            double left = tree[2 * node + 1];
            if (p <= left) node = 2 * node + 1;
            else { node = 2 * node + 2; p = p - tree[node - 1]; }
            if (p == 0.0) p = 1e-5;
        }
        prio_out[i] = tree[node];             /* :72 */
        idx_out[i] = node - (capacity - 1);   /* :73 */
    }
}

/* ------------------------------------------------------------------------------------------ */
/* LocalBuffer.finish TD — buffer.py:170-177.  rew16: rewards already rounded to fp16 and       */
/* given back as double (`rew_buf.tolist()`), q: float32[(size)*5], act: u8[size].             */
/* td: double[capacity], zeros beyond size.                                                    */
/* ------------------------------------------------------------------------------------------ */
void mo_actor_td(int size, int capacity, const double *rew16, const float *q, const uint8_t *act, double *td)
{
    for (int t = 0; t < capacity; ++t) td[t] = 0.0;
    for (int t = 0; t < size; ++t) {
        float qmax = q[t * 5];
        for (int a = 1; a < 5; ++a)
            if (q[t * 5 + a] > qmax) qmax = q[t * 5 + a];
        double r1 = (t + 1 < size) ? rew16[t + 1] : 0.0;
        /* np.convolve(ret, [0.99, 1.0], 'valid')[t] = ret[t+1]*0.99 + ret[t]*1.0 (forward_steps = 2, :174-175);
         * numpy evaluates the dot as ret[t]*k[1] + ret[t+1]*k[0] in that accumulation order */
        double conv = rew16[t] * 1.0 + r1 * 0.99;
        double target = conv + (double)qmax;            /* float64 + float32 -> float64 */
        td[t] = fabs(target - (double)q[t * 5 + act[t]]); /* :176-177 */
    }
}

/* The same for config.forward_steps = n (buffer.py:174-175): ret = rew + [0]*(n-1); kernel [g^(n-1), ..., g, 1];
 * np.convolve(ret, kernel, 'valid')[t] = sum_j ret[t+j] * g^j, accumulated from j = 0 upwards; g^j as Python's
 * 0.99**j (pow in fp64).  n = 2, g = 0.99 is mo_actor_td. */
void mo_actor_td_n(int size, int capacity, int n, double gamma, const double *rew16, const float *q, const uint8_t *act, double *td)
{
    for (int t = 0; t < capacity; ++t) td[t] = 0.0;
    for (int t = 0; t < size; ++t) {
        float qmax = q[t * 5];
        for (int a = 1; a < 5; ++a)
            if (q[t * 5 + a] > qmax) qmax = q[t * 5 + a];
        double conv = rew16[t] * 1.0;
        for (int j = 1; j < n; ++j) {
            double rj = (t + j < size) ? rew16[t + j] : 0.0;
            double gj = (j == 1) ? gamma : pow(gamma, (double)j);
            conv = conv + rj * gj;
        }
        double target = conv + (double)qmax;
        td[t] = fabs(target - (double)q[t * 5 + act[t]]);
    }
}

/* Learner TD / priority — worker.py:300-308 restated in fp32 (the reference runs it under CUDA
 * autocast fp16, which is not reproducible; parity for this row is "unpinned by the reference",
 * DESIGN.md).  q_online: f32[n*5], q_target_next: f32[n*5]. */
void mo_learner_td(int64_t n, const float *q_online, const float *q_target_next, const int64_t *action,
                   const float *reward, const float *done, const float *steps, float *td, float *prio)
{
    for (int64_t i = 0; i < n; ++i) {
        float m = q_target_next[i * 5];
        for (int a = 1; a < 5; ++a)
            if (q_target_next[i * 5 + a] > m) m = q_target_next[i * 5 + a];
        float q_ = (1.0f - done[i]) * m;                              /* :302 */
        float disc = powf(0.99f, steps[i]);                           /* :306 */
        float e = q_online[i * 5 + action[i]] - (reward[i] + disc * q_); /* :306 */
        td[i] = e;
        float a = fabsf(e);
        prio[i] = a < 1e-6f ? 1e-6f : a;                              /* :308 */
    }
}

int mo_abi_version(void) { return 1; }

/* ------------------------------------------------------------------------------------------ */
/* Communication mask of Network.step — model.py:196-208                                       */
/* pos: int32[N*2]; k = min(config.max_comm_agents, N); out: u8[N*N].                           */
/* Ranking by Euclidean distance (d^2 here: sqrt is monotonic); torch.topk leaves ties          */
/* unspecified — this restatement (like the CUDA kernel) gives them to the lower agent id.      */
/* ------------------------------------------------------------------------------------------ */
void mo_comm_mask(int N, const int32_t *pos, int k, int radius, uint8_t *out)
{
    int *used = (int *)malloc(sizeof(int) * (size_t)N);
    for (int i = 0; i < N; ++i) {
        memset(used, 0, sizeof(int) * (size_t)N);
        for (int j = 0; j < N; ++j) out[i * N + j] = 0;
        for (int q = 0; q < k && q < N; ++q) { /* selection of the q-th nearest, stable in j */
            int best = -1;
            long bd = 0;
            for (int j = 0; j < N; ++j) {
                if (used[j]) continue;
                long dx = pos[2 * i] - pos[2 * j], dy = pos[2 * i + 1] - pos[2 * j + 1];
                long dd = dx * dx + dy * dy;
                if (best < 0 || dd < bd) {
                    best = j;
                    bd = dd;
                }
            }
            used[best] = 1;
            int adx = abs(pos[2 * i] - pos[2 * best]), ady = abs(pos[2 * i + 1] - pos[2 * best + 1]);
            if (adx <= radius && ady <= radius) out[i * N + best] = 1; /* :201 in_obs_mask AND :206 */
        }
    }
    free(used);
}
