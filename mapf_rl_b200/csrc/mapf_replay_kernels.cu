// mapf_replay_kernels.cu — K5: the sampled-window gather of the replay store, the data format on the
// learner side of the hot path (GlobalBuffer.sample_batch, worker.py:106-184), sm_100a.
//
// The store keeps the reference's logical layout on the device (worker.py:36-42): episode slot g holds
// max_steps + 1 observation / comm-mask frames at rows g * (max_steps + 1) + f and max_steps action /
// reward / hidden rows at g * max_steps + t; tree leaf idx = g * max_steps + t.  Observations stay bool
// bytes exactly as the step kernel writes them.  For every sampled leaf the kernel gathers the
// bt_steps + forward_steps frame window (zero padded), converts bool -> fp16 on the fly
// (np.stack(b_obs).astype(np.float16), worker.py:169), gathers the comm masks and the stored hidden
// state, and emits the per-sample scalars — one launch, HBM-bound: per sample it reads at most
// W * N * 486 bytes and writes W * N * 972 (W = bt_steps + forward_steps).
#include <cuda_fp16.h>

#include "mapf_common.cuh"

namespace {

struct GatherParams {
    mapf_replay_view v;
    mapf_replay_batch o;
    const int64_t *idx;
    int64_t batch;
    int32_t *err;
};

// two bool bytes (b0 at bits 0-7, b1 at bits 16-23) -> two fp16 (1.0 = 0x3C00)
__device__ __forceinline__ uint32_t bools_to_half2(uint32_t spread) { return spread * 0x3C00u; }

__global__ void __launch_bounds__(256)
replay_gather_kernel(const GatherParams p)
{
    const mapf_replay_view &v = p.v;
    const int W = v.bt_steps + v.forward_steps;
    const int j = blockIdx.x;  // frame of the window, or W = the hidden-state / scalar block
    const int64_t b = blockIdx.y;
    const int64_t idx = p.idx[b];
    const int64_t g = idx / v.max_steps;             // worker.py:115
    const int t = (int)(idx - g * v.max_steps);      // worker.py:116
    const int size = v.size_buf[g];
    const int steps = min(v.forward_steps, size - t);  // worker.py:122
    const int f0 = max(0, t + 1 - v.bt_steps);         // worker.py:124-137: the window starts at the episode start
    const int len = t + 1 + steps - f0;                //                   until bt_steps frames of history exist
    const int N = v.num_agents;

    if (j == W) {
        // stored hidden state of frame t - bt_steps (worker.py:137), zeros while t < bt_steps (:127,:132)
        const int HB = N * v.latent_dim;  // halfs
        uint16_t *dst = p.o.hidden + (size_t)b * HB;
        const uint16_t *src = t >= v.bt_steps ? v.hid_buf + (size_t)(idx - v.bt_steps) * HB : nullptr;
        if ((HB & 7) == 0) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            uint4 *d4 = reinterpret_cast<uint4 *>(dst);
            for (int k = threadIdx.x; k < (HB >> 3); k += blockDim.x) d4[k] = src ? __ldg(s4 + k) : make_uint4(0, 0, 0, 0);
        } else {
            for (int k = threadIdx.x; k < HB; k += blockDim.x) dst[k] = src ? src[k] : (uint16_t)0;
        }
        if (threadIdx.x == 0) {
            if (t >= size && p.err) atomicOr(p.err, 1);  // the reference asserts local_idx < size (worker.py:120)
            p.o.action[b] = v.act_buf[idx];                                               // :143
            p.o.reward[b] = v.rew_buf[idx];                                               // :144
            const bool done = (t == size - 1) && v.done_buf[g] != 0;                      // :145-148
            p.o.done[b] = done ? 0x3C00 : 0;
            p.o.steps[b] = __half_as_ushort(__int2half_rn(steps));                        // :171 HalfTensor(b_steps)
            p.o.bt_steps[b] = min(t + 1, v.bt_steps);                                     // :150
        }
        return;
    }

    const bool live = j < len;  // frames past the window are zero padding (worker.py:139-142)
    const int64_t row = g * (v.max_steps + 1) + f0 + j;
    // ---- observation frame: N * 486 bool bytes -> fp16 ----
    {
        const int FB = N * MAPF_OBS_BYTES_PER_AGENT;
        const uint8_t *src = v.obs_buf + (size_t)row * FB;
        uint16_t *dst = p.o.obs + ((size_t)b * W + j) * FB;
        if ((FB & 15) == 0) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            uint4 *d4 = reinterpret_cast<uint4 *>(dst);
            for (int k = threadIdx.x; k < (FB >> 4); k += blockDim.x) {
                uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
                if (live) {
                    const uint4 x = __ldg(s4 + k);
                    lo.x = bools_to_half2(__byte_perm(x.x, 0, 0x4140));
                    lo.y = bools_to_half2(__byte_perm(x.x, 0, 0x4342));
                    lo.z = bools_to_half2(__byte_perm(x.y, 0, 0x4140));
                    lo.w = bools_to_half2(__byte_perm(x.y, 0, 0x4342));
                    hi.x = bools_to_half2(__byte_perm(x.z, 0, 0x4140));
                    hi.y = bools_to_half2(__byte_perm(x.z, 0, 0x4342));
                    hi.z = bools_to_half2(__byte_perm(x.w, 0, 0x4140));
                    hi.w = bools_to_half2(__byte_perm(x.w, 0, 0x4342));
                }
                __stcs(d4 + 2 * k, lo);
                __stcs(d4 + 2 * k + 1, hi);
            }
        } else {  // N not a multiple of 8: 2 bools -> one 32-bit store (N * 486 is always even)
            const uint16_t *s2 = reinterpret_cast<const uint16_t *>(src);
            uint32_t *d2 = reinterpret_cast<uint32_t *>(dst);
            for (int k = threadIdx.x; k < (FB >> 1); k += blockDim.x) {
                uint32_t out = 0;
                if (live) {
                    const uint32_t x = s2[k];
                    out = bools_to_half2(__byte_perm(x, 0, 0x4140));
                }
                d2[k] = out;
            }
        }
    }
    // ---- communication mask frame: N * N bool bytes ----
    {
        const int CBytes = N * N;
        const uint8_t *src = v.comm_buf + (size_t)row * CBytes;
        uint8_t *dst = p.o.comm_mask + ((size_t)b * W + j) * CBytes;
        for (int k = threadIdx.x; k < CBytes; k += blockDim.x) dst[k] = live ? src[k] : (uint8_t)0;
    }
}

}  // namespace

int mapf_launch_replay_gather(const mapf_replay_view *view, const int64_t *d_idx, int64_t batch, const mapf_replay_batch *out,
                              int32_t *d_err, cudaStream_t st)
{
    GatherParams p;
    p.v = *view;
    p.o = *out;
    p.idx = d_idx;
    p.batch = batch;
    p.err = d_err;
    const int W = view->bt_steps + view->forward_steps;
    dim3 grid((unsigned)(W + 1), (unsigned)batch);
    replay_gather_kernel<<<grid, 256, 0, st>>>(p);
    MAPF_CUDA(cudaGetLastError());
    return MAPF_OK;
}
