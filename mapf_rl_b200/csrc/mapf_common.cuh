// mapf_common.cuh — shared declarations of libmapf_b200.so (sm_100a only).
//
// Device data layout of one environment batch (all owned by the mapf_env handle):
//
//   obst   u32[B][obst_stride]      obstacle bitmap.  Row r of the PADDED grid (R = L + 8 rows, the
//                                   obs_radius = 4 border included) occupies RWS = RW + 1 words, bit p of
//                                   the row = padded column p (cell column y is bit y + 4).  Border rows /
//                                   columns and the extra word are zero, so a 9-wide window is always a
//                                   two-word funnel shift with no bounds checks (outside the map = 0,
//                                   environment.py:447).
//   pos    u8[B][N][2], goal u8[B][N][2]   (x, y) = (row, col)
//   navi   u64[B][N][NB][NB][16]    heuristic bits (environment.py:253-276) as OVERLAPPING 16x16-cell tiles of
//                                   the padded grid, one 128-byte line each: tile (bx, by) covers padded rows
//                                   8bx .. 8bx+15 and padded columns 8by .. 8by+15; its row r is one u64 whose
//                                   bit 16 d + c is direction d (0 up, 1 down, 2 left, 3 right) at column
//                                   8by + c.  The 9x9 window of an agent at (x, y) (padded top-left corner
//                                   (x, y)) lies entirely inside tile (x >> 3, y >> 3): 9 consecutive u64 of
//                                   ONE aligned 128-byte line, whatever the position (every cell is stored in
//                                   up to 4 tiles; NB = ceil(L / 8) tiles per side).
//   steps  i32[B]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/mapf_b200.h"

#define MAPF_FULL_MASK 0xffffffffu

struct EnvDims {
    int B, N, L;
    int R;            // L + 8 padded rows
    int RW;           // words holding the L + 8 padded column bits
    int RWS;          // RW + 1: row stride in words
    int NB;           // ceil(L / 8): navi tiles per side
    int K;            // ceil(N / 32) agent slots per lane
    int obst_stride;  // words per env in `obst` (R * RWS rounded up to 4)
    int navi_agent_stride;  // NB * NB * 32 words (NB * NB tiles of 128 bytes) per agent
};

struct mapf_env {
    EnvDims d;
    int device;
    int num_sms;
    float reward[5];
    uint32_t *obst;
    uint8_t *pos;
    uint8_t *goal;
    uint32_t *navi;
    int32_t *steps;
    int32_t *err;      // latched device error bits
    // mapf_env_rollout: one internal stream per chain of sub-batch launches, fork / join events
    cudaStream_t chain_stream[MAPF_MAX_CHAINS];
    cudaEvent_t chain_done[MAPF_MAX_CHAINS];
    cudaEvent_t chain_fork;
    // long rollouts replay, per chain, a captured graph of one slot period (P = lcm of the slot counts) of launches
    struct RolloutGraph {
        const void *act, *obs, *rew, *done, *steps;
        int action_slots, obs_slots, out_slots, S, P, tuning_gen;
        cudaGraphExec_t exec[MAPF_MAX_CHAINS];
    } rg[2];
    int rg_next;
    int split_key, split_per_sm;  // resident CTAs per SM of the split step kernel, cached per (variant) key
    // staging for the host-buffer entry point
    uint8_t *d_actions;
    uint8_t *d_obs;
    float *d_rewards;
    uint8_t *d_done;
    int32_t *d_steps_out;
    uint8_t *h_pinned;  // pinned staging: actions | rewards | done | steps
    // last caller buffers of mapf_env_step_host and what they resolved to (a per-step actor reuses its buffers, so
    // the four cudaPointerGetAttributes / cudaHostGetDevicePointer queries are paid once)
    uint8_t *pin_alias;  // device alias of h_pinned
    cudaStream_t side_stream;          // result copies of mapf_env_step_host run here, next to the observe kernel
    cudaEvent_t ev_stepped, ev_copied;
    // the same sequence captured once per (buffer set, observation target) and replayed with one launch
    struct HostGraph {
        const void *act, *rew, *done, *steps, *hobs, *obs_dev;
        int mode;
        cudaGraphExec_t exec;
    } hg[32];
    int hg_next;
    cudaStream_t cap_stream;
    struct PtrInfo {
        const void *key;
        bool pinned;
        void *alias;
    } ptr_cache[32];
    int ptr_cache_next;
    const void *hc_key[4];
    void *hc_alias[4];   // device alias of the page-locked buffer, or NULL
    bool hc_pinned[4];
    int64_t arena_bytes;
};

// scratch owned by a tree handle
struct PerScratch {
    unsigned long long *stamps;  // u64[capacity]: (epoch << 32 | batch position + 1) of the last claimant
    float *prio32;               // f32[cap_n]
    uint8_t *active;             // u8[cap_n]
    int64_t cap_n;
    unsigned long long epoch;
};

struct mapf_per {
    int64_t capacity;
    int layer;
    int device;
    double *tree;
    PerScratch scratch;
};

void mapf_set_error(const std::string &msg);
int mapf_cuda_fail(cudaError_t e, const char *what);

#define MAPF_CUDA(expr)                                        \
    do {                                                       \
        cudaError_t _e = (expr);                               \
        if (_e != cudaSuccess) return mapf_cuda_fail(_e, #expr); \
    } while (0)

#define MAPF_ERRBIT_ACTION 1
#define MAPF_ERRBIT_UNIQUE 2
#define MAPF_ERRBIT_RESET 4
