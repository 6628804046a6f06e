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
#include "mapf_common.cuh"

namespace {

constexpr int kPerThreads = 1024;

__device__ __forceinline__ unsigned long long stamp_of(unsigned long long epoch, int64_t k)
{
    return (epoch << 32) | (unsigned long long)(k + 1);
}

// Leaf writes with numpy's duplicate rule, then the level-by-level ancestor refresh.  Single CTA.
// `active[k]` (optional) masks entries out (stale indices, worker.py:192-201).
__device__ void tree_update_cta(double *tree, unsigned long long *stamps, unsigned long long epoch, int64_t capacity,
                                int layer, const int64_t *idx, const double *prio_in, const float *prio32, double alpha,
                                const uint8_t *active, int64_t n)
{
    // phase 1: the highest batch position claims each leaf
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x)
        if (!active || active[k]) atomicMax(&stamps[idx[k]], stamp_of(epoch, k));
    __syncthreads();
    // phase 2: winners store the leaf
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) {
        if (active && !active[k]) continue;
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
            if (active && !active[k]) continue;
            int64_t node = ((capacity - 1 + idx[k] + 1) >> l) - 1;  // l-th ancestor in the array heap
            tree[node] = __dadd_rn(tree[2 * node + 1], tree[2 * node + 2]);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kPerThreads)
per_update_kernel(double *tree, unsigned long long *stamps, unsigned long long epoch, int64_t capacity, int layer,
                  const int64_t *idx, const double *prio, int64_t n)
{
    tree_update_cta(tree, stamps, epoch, capacity, layer, idx, prio, nullptr, 0.0, nullptr, n);
}

// multi-CTA variants for batches larger than one CTA handles comfortably
__global__ void per_claim_kernel(unsigned long long *stamps, unsigned long long epoch, const int64_t *idx, int64_t n)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) atomicMax(&stamps[idx[k]], stamp_of(epoch, k));
}
__global__ void per_leaf_kernel(double *tree, const unsigned long long *stamps, unsigned long long epoch, int64_t capacity,
                                const int64_t *idx, const double *prio, int64_t n)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && stamps[idx[k]] == stamp_of(epoch, k)) tree[capacity - 1 + idx[k]] = prio[k];
}
__global__ void per_level_kernel(double *tree, int64_t capacity, int l, const int64_t *idx, int64_t n)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        int64_t node = ((capacity - 1 + idx[k] + 1) >> l) - 1;
        tree[node] = __dadd_rn(tree[2 * node + 1], tree[2 * node + 2]);
    }
}

// SumTree.batch_sample, buffer.py:56-78.  One thread per sample, layer-1 dependent L2 reads.
__global__ void __launch_bounds__(kPerThreads)
per_sample_kernel(const double *__restrict__ tree, int64_t capacity, int layer, const double *__restrict__ uniforms,
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

// Fused learner tail: TD error -> priority -> stale mask -> leaf = prio^alpha -> ancestor refresh.
__global__ void __launch_bounds__(kPerThreads)
per_td_update_kernel(double *tree, unsigned long long *stamps, unsigned long long epoch, int64_t capacity, int layer,
                     const float *__restrict__ q_online, const float *__restrict__ q_target_next,
                     const float *__restrict__ q_online_next, const int64_t *__restrict__ action,
                     const float *__restrict__ reward, const float *__restrict__ done, const float *__restrict__ steps,
                     const int64_t *__restrict__ idx, int64_t n, float gamma, double alpha, int64_t old_ptr, int64_t ptr,
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
    tree_update_cta(tree, stamps, epoch, capacity, layer, idx, nullptr, prio_scratch, alpha, active, n);
}

// LocalBuffer.finish, buffer.py:170-177
__global__ void actor_td_kernel(const float *__restrict__ rew, const float *__restrict__ q, const uint8_t *__restrict__ act,
                                const int32_t *__restrict__ size, int episodes, int capacity, double *__restrict__ td)
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
        const double r0 = (double)rew[g];
        const double r1 = (t + 1 < sz) ? (double)rew[g + 1] : 0.0;
        const double conv = __dadd_rn(r0, __dmul_rn(r1, 0.99));  // np.convolve(ret, [0.99, 1.0], 'valid')
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
        per_update_kernel<<<1, kPerThreads, 0, st>>>(t->tree, s->stamps, epoch, t->capacity, t->layer, d_idx, d_prio, n);
    } else {
        const int tb = 256;
        const unsigned grid = (unsigned)((n + tb - 1) / tb);
        per_claim_kernel<<<grid, tb, 0, st>>>(s->stamps, epoch, d_idx, n);
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
                                                    ptr, slot_steps, td_out, prio_out, s->prio32, s->active);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}

int mapf_launch_actor_td(const float *d_rew, const float *d_q, const uint8_t *d_act, const int32_t *d_size, int episodes,
                         int capacity, double *d_td_out, cudaStream_t st)
{
    const int64_t total = (int64_t)episodes * capacity;
    if (total <= 0) return MAPF_OK;
    actor_td_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_rew, d_q, d_act, d_size, episodes, capacity, d_td_out);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}
