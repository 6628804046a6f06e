// mapf_per_kernels.cu — K4: prioritized-replay sum tree (buffer.py:16-105) and the TD-error /
// priority arithmetic around it (buffer.py:170-177, worker.py:186-203, 300-308), sm_100a.
//
// The tree is an fp64 array heap of 2*capacity-1 nodes exactly like SumTree.tree (buffer.py:25), so
// results are bit-comparable with numpy: every parent is recomputed as left + right (order-free),
// duplicate leaf indices resolve "last in batch order wins" (numpy fancy assignment, buffer.py:97),
// and the sampling descent uses the reference's comparisons (buffer.py:66-70) with FMA contraction
// disabled.  The whole tree (8 MiB at the reference's 2^19 leaves) lives in L2; the kernels are
// latency-bound (layer-1 dependent levels), not bandwidth-bound, so one CTA handles a batch and the
// fused learner tail (TD -> priority -> leaf write -> ancestor refresh) is a single launch.
#include <cmath>

#include "mapf_common.cuh"

namespace {

constexpr int kPerThreads = 1024;

__device__ __forceinline__ unsigned long long stamp_of(unsigned long long epoch, int64_t k)
{
    return (epoch << 32) | (unsigned long long)(k + 1);
}

// Leaf writes with numpy's duplicate rule, then the level-by-level ancestor refresh.  Single CTA.
// `active[k]` (optional) masks entries out (stale indices, worker.py:192-201).  An index outside [0, capacity) -- numpy
// raises IndexError -- is skipped and latched in *err (mapf_per_status).
__device__ void tree_update_cta(double *tree, unsigned long long *stamps, unsigned long long epoch, int64_t capacity,
                                int layer, const int64_t *idx, const double *prio_in, const float *prio32, double alpha,
                                const uint8_t *active, int64_t n, int32_t *err)
{
    auto live = [&](int64_t k) -> bool {
        if (active && !active[k]) return false;
        const int64_t i = idx[k];
        return i >= 0 && i < capacity;
    };
    // phase 1: the highest batch position claims each leaf
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
        if (live(k)) atomicMax(&stamps[idx[k]], stamp_of(epoch, k));
        else if (!active || active[k]) atomicOr(err, MAPF_ERRBIT_INDEX);
    }
    __syncthreads();
    // phase 2: winners store the leaf
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
        if (!live(k)) continue;
        if (stamps[idx[k]] == stamp_of(epoch, k)) {
            double v = prio_in ? prio_in[k] : pow((double)prio32[k], alpha);
            tree[capacity - 1 + idx[k]] = v;
        }
    }
    __syncthreads();
    // phase 3: ancestors, buffer.py:99-102 (np.unique only removes redundant work; recomputing a parent
    // twice from final children gives the same bits)
    for (int l = 1; l < layer; ++l) {
        for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
            if (!live(k)) continue;
            int64_t node = ((capacity - 1 + idx[k] + 1) >> l) - 1;  // l-th ancestor in the array heap
            tree[node] = __dadd_rn(tree[2 * node + 1], tree[2 * node + 2]);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kPerThreads)
per_update_kernel(double *tree, unsigned long long *stamps, unsigned long long epoch, int64_t capacity, int layer,
                  const int64_t *idx, const double *prio, int64_t n, int32_t *err)
{
    tree_update_cta(tree, stamps, epoch, capacity, layer, idx, prio, nullptr, 0.0, nullptr, n, err);
}

// multi-CTA variants for batches larger than one CTA handles comfortably
__global__ void per_claim_kernel(unsigned long long *stamps, unsigned long long epoch, int64_t capacity, const int64_t *idx, int64_t n,
                                 int32_t *err)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (idx[k] >= 0 && idx[k] < capacity) atomicMax(&stamps[idx[k]], stamp_of(epoch, k));
    else atomicOr(err, MAPF_ERRBIT_INDEX);
}
__global__ void per_leaf_kernel(double *tree, const unsigned long long *stamps, unsigned long long epoch, int64_t capacity,
                                const int64_t *idx, const double *prio, int64_t n)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && idx[k] >= 0 && idx[k] < capacity && stamps[idx[k]] == stamp_of(epoch, k)) tree[capacity - 1 + idx[k]] = prio[k];
}
__global__ void per_level_kernel(double *tree, int64_t capacity, int l, const int64_t *idx, int64_t n)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && idx[k] >= 0 && idx[k] < capacity) {
        int64_t node = ((capacity - 1 + idx[k] + 1) >> l) - 1;
        tree[node] = __dadd_rn(tree[2 * node + 1], tree[2 * node + 2]);
    }
}

// SumTree.batch_sample, buffer.py:56-78.  One thread per sample, layer-1 dependent L2 reads.
__device__ void sample_cta(const double *__restrict__ tree, int64_t capacity, int layer, const double *__restrict__ uniforms,
                           int64_t batch, int64_t *__restrict__ idx_out, double *__restrict__ prio_out,
                           float *__restrict__ weight_out, double beta)
{
    __shared__ double s_min[32];
    const double sum = tree[0];
    const double interval = __ddiv_rn(sum, (double)batch);  // :58
    double local_min = 1.0e300;
    for (int64_t i = threadIdx.x; i < batch; i += blockDim.x) {
        // :60  np.arange(0,sum,interval)[i] + np.random.uniform(0,interval)[i]; no FMA contraction
        double p = __dadd_rn(__dmul_rn((double)i, interval), __dmul_rn(uniforms[i], interval));
        if (i == 0 && p == 0.0) p = 1e-5;  // :61-62
        int64_t node = 0;
        for (int l = 0; l < layer - 1; ++l) {  // :66-70
            const double left = tree[2 * node + 1];
            if (p <= left) node = 2 * node + 1;
            else {
                node = 2 * node + 2;
                p = __dsub_rn(p, left);
            }
            if (p == 0.0) p = 1e-5;
        }
        const double pr = tree[node];
        prio_out[i] = pr;                     // :72
        idx_out[i] = node - (capacity - 1);   // :73
        local_min = fmin(local_min, pr);
    }
    if (weight_out) {  // worker.py:165-166
        for (int o = 16; o > 0; o >>= 1) local_min = fmin(local_min, __shfl_xor_sync(MAPF_FULL_MASK, local_min, o));
        if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = local_min;
        __syncthreads();
        double m = s_min[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmin(m, s_min[w]);
        for (int64_t i = threadIdx.x; i < batch; i += blockDim.x)
            weight_out[i] = (float)pow(prio_out[i] / m, -beta);
    }
}

__global__ void __launch_bounds__(kPerThreads)
per_sample_kernel(const double *__restrict__ tree, int64_t capacity, int layer, const double *__restrict__ uniforms,
                  int64_t batch, int64_t *__restrict__ idx_out, double *__restrict__ prio_out,
                  float *__restrict__ weight_out, double beta)
{
    sample_cta(tree, capacity, layer, uniforms, batch, idx_out, prio_out, weight_out, beta);
}

// Learner TD error -> priority -> stale mask (worker.py:300-308, 192-201), one thread per transition
__device__ void td_cta(const float *__restrict__ q_online, const float *__restrict__ q_target_next,
                       const float *__restrict__ q_online_next, const int64_t *__restrict__ action,
                       const float *__restrict__ reward, const float *__restrict__ done, const float *__restrict__ steps,
                       const int64_t *__restrict__ idx, int64_t n, float gamma, int64_t old_ptr, int64_t ptr,
                       int64_t slot_steps, float *__restrict__ td_out, float *__restrict__ prio_out, float *prio_scratch,
                       uint8_t *active)
{
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
        const float *qt = q_target_next + k * 5;
        float boot;
        if (q_online_next) {  // double-Q option: online argmax, target value
            const float *qn = q_online_next + k * 5;
            int best = 0;
            for (int a = 1; a < 5; ++a)
                if (qn[a] > qn[best]) best = a;
            boot = qt[best];
        } else {  // worker.py:302 target-network max
            boot = qt[0];
            for (int a = 1; a < 5; ++a) boot = fmaxf(boot, qt[a]);
        }
        const float q_ = __fmul_rn(1.0f - done[k], boot);                                      // :302
        const float disc = powf(gamma, steps[k]);                                              // :306
        const float target = __fadd_rn(reward[k], __fmul_rn(disc, q_));
        const float td = __fsub_rn(q_online[k * 5 + action[k]], target);                       // :306
        const float pr = fmaxf(fabsf(td), 1e-6f);                                              // :308
        if (td_out) td_out[k] = td;
        if (prio_out) prio_out[k] = pr;
        prio_scratch[k] = pr;
        // worker.py:192-201 discard samples whose episode slot was rewritten since sampling
        bool keep = true;
        const int64_t lo = old_ptr * slot_steps, hi = ptr * slot_steps;
        if (ptr > old_ptr) keep = (idx[k] < lo) || (idx[k] >= hi);
        else if (ptr < old_ptr) keep = (idx[k] < lo) && (idx[k] >= hi);
        active[k] = keep ? 1 : 0;
    }
    __syncthreads();
}

// Fused learner tail: TD error -> priority -> stale mask -> leaf = prio^alpha -> ancestor refresh.
__global__ void __launch_bounds__(kPerThreads)
per_td_update_kernel(double *tree, unsigned long long *stamps, unsigned long long epoch, int64_t capacity, int layer,
                     const float *__restrict__ q_online, const float *__restrict__ q_target_next,
                     const float *__restrict__ q_online_next, const int64_t *__restrict__ action,
                     const float *__restrict__ reward, const float *__restrict__ done, const float *__restrict__ steps,
                     const int64_t *__restrict__ idx, int64_t n, float gamma, double alpha, int64_t old_ptr, int64_t ptr,
                     int64_t slot_steps, float *__restrict__ td_out, float *__restrict__ prio_out, float *prio_scratch,
                     uint8_t *active, int32_t *err)
{
    td_cta(q_online, q_target_next, q_online_next, action, reward, done, steps, idx, n, gamma, old_ptr, ptr, slot_steps, td_out,
           prio_out, prio_scratch, active);
    tree_update_cta(tree, stamps, epoch, capacity, layer, idx, nullptr, prio_scratch, alpha, active, n, err);
}

// One learner cycle in ONE launch (north star (4)): the priorities of the batch that has just been through the two Q
// forwards go into the tree (TD -> priority -> stale mask -> leaf -> ancestors), then the NEXT batch is drawn from the
// refreshed tree with its importance-sampling weights (worker.py:300-308, 186-203, 106-116, 165-166).  Either half may be
// empty (n == 0).
__global__ void __launch_bounds__(kPerThreads)
per_cycle_kernel(double *tree, unsigned long long *stamps, unsigned long long epoch, int64_t capacity, int layer,
                 const mapf_per_cycle_args a, float *prio_scratch, uint8_t *active, int32_t *err)
{
    if (a.n_update > 0) {
        td_cta(a.d_q_online, a.d_q_target_next, a.d_q_online_next, a.d_action, a.d_reward, a.d_done, a.d_steps, a.d_idx, a.n_update,
               a.gamma, a.old_ptr, a.ptr, a.slot_steps, a.d_td_out, a.d_prio_out, prio_scratch, active);
        tree_update_cta(tree, stamps, epoch, capacity, layer, a.d_idx, nullptr, prio_scratch, a.alpha, active, a.n_update, err);
    }
    if (a.n_sample > 0)
        sample_cta(tree, capacity, layer, a.d_uniforms, a.n_sample, a.d_sample_idx_out, a.d_sample_prio_out, a.d_sample_weight_out,
                   a.beta);
}

// LocalBuffer.finish, buffer.py:170-177: |sum_j gamma^j r[t+j] + max_a q[t,a] - q[t,a_t]| over j < forward_steps (rewards past
// the episode end are 0); `gpow[j]` = gamma^j as Python evaluates 0.99**j (fp64, computed on the host)
struct ActorTdPowers {
    double g[8];
};
__global__ void actor_td_kernel(const float *__restrict__ rew, const float *__restrict__ q, const uint8_t *__restrict__ act,
                                const int32_t *__restrict__ size, int episodes, int capacity, int forward_steps,
                                const ActorTdPowers gpow, double *__restrict__ td)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= (int64_t)episodes * capacity) return;
    const int ep = (int)(g / capacity), t = (int)(g - (int64_t)ep * capacity);
    const int sz = size[ep];
    double out = 0.0;
    if (t < sz) {
        const float *qq = q + g * 5;
        float qmax = qq[0];
        for (int a = 1; a < 5; ++a) qmax = fmaxf(qmax, qq[a]);
        // np.convolve(ret, [gamma^(n-1), ..., gamma, 1.0], 'valid')[t]: accumulated from j = 0 upwards
        double conv = __dmul_rn((double)rew[g], 1.0);
        for (int j = 1; j < forward_steps; ++j) {
            const double rj = (t + j < sz) ? (double)rew[g + j] : 0.0;
            conv = __dadd_rn(conv, __dmul_rn(rj, gpow.g[j]));
        }
        const double target = __dadd_rn(conv, (double)qmax);
        out = fabs(__dsub_rn(target, (double)qq[act[g]]));
    }
    td[g] = out;
}

}  // namespace

int mapf_launch_per_update(mapf_per *t, PerScratch *s, const int64_t *d_idx, const double *d_prio, int64_t n, cudaStream_t st)
{
    if (n <= 0) return MAPF_OK;
    const unsigned long long epoch = ++s->epoch;
    if (n <= 4096) {
        per_update_kernel<<<1, kPerThreads, 0, st>>>(t->tree, s->stamps, epoch, t->capacity, t->layer, d_idx, d_prio, n, s->err);
    } else {
        const int tb = 256;
        const unsigned grid = (unsigned)((n + tb - 1) / tb);
        per_claim_kernel<<<grid, tb, 0, st>>>(s->stamps, epoch, t->capacity, d_idx, n, s->err);
        per_leaf_kernel<<<grid, tb, 0, st>>>(t->tree, s->stamps, epoch, t->capacity, d_idx, d_prio, n);
        for (int l = 1; l < t->layer; ++l) per_level_kernel<<<grid, tb, 0, st>>>(t->tree, t->capacity, l, d_idx, n);
    }
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

int mapf_launch_per_sample(mapf_per *t, const double *d_uniforms, int64_t batch, int64_t *d_idx_out, double *d_prio_out,
                           float *d_weight_out, double beta, cudaStream_t st)
{
    if (batch <= 0) return MAPF_OK;
    per_sample_kernel<<<1, kPerThreads, 0, st>>>(t->tree, t->capacity, t->layer, d_uniforms, batch, d_idx_out, d_prio_out,
                                                 d_weight_out, beta);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

int mapf_launch_per_td_update(mapf_per *t, PerScratch *s, const float *q_online, const float *q_target_next,
                              const float *q_online_next, const int64_t *action, const float *reward, const float *done,
                              const float *steps, const int64_t *idx, int64_t n, float gamma, double alpha, int64_t old_ptr,
                              int64_t ptr, int64_t slot_steps, float *td_out, float *prio_out, cudaStream_t st)
{
    if (n <= 0) return MAPF_OK;
    const unsigned long long epoch = ++s->epoch;
    per_td_update_kernel<<<1, kPerThreads, 0, st>>>(t->tree, s->stamps, epoch, t->capacity, t->layer, q_online, q_target_next,
                                                    q_online_next, action, reward, done, steps, idx, n, gamma, alpha, old_ptr,
                                                    ptr, slot_steps, td_out, prio_out, s->prio32, s->active, s->err);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

int mapf_launch_per_cycle(mapf_per *t, PerScratch *s, const mapf_per_cycle_args *a, cudaStream_t st)
{
    if (a->n_update <= 0 && a->n_sample <= 0) return MAPF_OK;
    const unsigned long long epoch = ++s->epoch;
    per_cycle_kernel<<<1, kPerThreads, 0, st>>>(t->tree, s->stamps, epoch, t->capacity, t->layer, *a, s->prio32, s->active, s->err);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

int mapf_launch_actor_td(const float *d_rew, const float *d_q, const uint8_t *d_act, const int32_t *d_size, int episodes,
                         int capacity, int forward_steps, double gamma, double *d_td_out, cudaStream_t st)
{
    const int64_t total = (int64_t)episodes * capacity;
    if (total <= 0) return MAPF_OK;
    ActorTdPowers gp{};
    for (int j = 0; j < 8; ++j) gp.g[j] = j == 0 ? 1.0 : (j == 1 ? gamma : std::pow(gamma, (double)j));
    actor_td_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_rew, d_q, d_act, d_size, episodes, capacity, forward_steps, gp,
                                                                     d_td_out);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}
