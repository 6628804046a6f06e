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
        const void *act, *obs, *rew, *codes, *done, *steps;
        int action_slots, obs_slots, out_slots, S, P, tuning_gen;
        cudaGraphExec_t exec[MAPF_MAX_CHAINS];
    } rg[2];
    int rg_next;
    // persistent rollout kernel (mapf_rollout_kernels.cu): scheduler words, resident CTAs per SM (cached per store mode)
    unsigned long long *ro_work;  // [0] next work item, [1] warps that have left; both 0 between launches
    uint32_t *ro_progress;        // [B] chunks of an environment finished inside the running launch
    uint32_t *ro_episode;         // [B] instances generated for the slot by in-launch episode handling
    uint32_t *ro_prio;            // [1 + B] count + ids of the environments that hit the step cap inside the next launch
    uint8_t *ro_prio_flag;        // [B]
    int ro_key, ro_per_sm;
    // episode handling inside mapf_env_rollout (mapf_env_set_autoreset); ar_max_steps == 0: off
    int ar_max_steps;
    unsigned long long ar_seed, ar_offset, ar_stride;
    float ar_density;
    int check_unique;             // post-step uniqueness check (environment.py:424-428) in every step launch
    // staging for the host-buffer entry points
    uint8_t *d_actions;
    uint8_t *d_obs;
    float *d_rewards;
    uint8_t *d_done;
    int32_t *d_steps_out;
    uint8_t *h_pinned;  // pinned staging: actions | rewards or codes | steps | done
    cudaStream_t side_stream;          // result copies of mapf_env_step_host run here, next to the observe kernel
    cudaEvent_t ev_stepped, ev_copied;
    // the same sequence captured once per (buffer set, observation target) and replayed with one launch
    struct HostGraph {
        const void *act, *rew, *done, *steps, *hobs, *obs_dev;
        int gen;
        cudaGraphExec_t exec;
    } hg[32];
    int hg_next;
    int checks_gen;
    cudaStream_t cap_stream;
    // mapf_env_step_host_codes: a two-stage pipeline -- step kernel + result copy on an internal stream, observe kernel on
    // the caller's stream reading a SNAPSHOT of the positions, so that step t+1 overlaps the observation stores of step t
    cudaStream_t hp_stream;        // step kernel + result copies
    cudaEvent_t hp_step[2];        // step t (kernel + copies) done
    cudaEvent_t hp_obs[2];         // observe kernel of step t done (its position snapshot may be overwritten)
    cudaEvent_t hp_in;             // what the caller queued before the first call of a run
    cudaEvent_t hp_stepped;        // the step kernel of the current call is done (the observe kernel may start)
    uint8_t *hp_pos[2];            // u8[B,N,2] position snapshots
    uint8_t *hp_results;           // device staging: codes u8[BN] | pad 16 | steps i32[B] | done u8[B]
    uint64_t hp_t;                 // calls since the pipeline was (re)started
    int hp_active;                 // work of the pipeline may still be in flight (other entry points drain it first)
    int64_t arena_bytes;
};

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
    float *rewards;          // [B,N] optional  (step only)
    uint8_t *codes;          // [B,N] optional  (step only): index of the reward in reward_fn order (MAPF_RCODE_*)
    uint8_t *done;           // [B]             (step only)
    int32_t *steps_out;      // [B] optional
    uint8_t *pos_out;        // [B,N,2] optional (observe only)
    float r[5];              // reward_fn in MAPF_RCODE order: move, stay_on_goal, stay_off_goal, collision, finish
    int warp_smem_words;     // per-warp shared memory, multiple of 4 words
    int obst_words;          // = d.obst_stride
    int bits_words;          // words of the per-env observation bit stream (also holds the occupancy grid)
    int flags;               // MAPF_STEPF_*
    int env_begin, env_end;  // the launch covers environments [env_begin, env_end) of the batch
};

// outputs of a step launch; rewards and codes are each optional
struct StepOut {
    float *rewards = nullptr;   // f32[B,N]
    uint8_t *codes = nullptr;   // u8[B,N] MAPF_RCODE_*
    uint8_t *done = nullptr;    // u8[B]
    int32_t *steps = nullptr;   // i32[B] optional
};

struct mapf_env;
StepParams mapf_make_step_params(const mapf_env *env);

// scratch owned by a tree handle
struct PerScratch {
    unsigned long long *stamps;  // u64[capacity]: (epoch << 32 | batch position + 1) of the last claimant
    float *prio32;               // f32[cap_n]
    uint8_t *active;             // u8[cap_n]
    int64_t cap_n;
    unsigned long long epoch;
    int32_t *err;                // latched device error bits (MAPF_ERRBIT_INDEX)
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
#define MAPF_ERRBIT_INTERNAL 8   /* rollout scheduler: a chunk hand-over never arrived */
#define MAPF_ERRBIT_STATE 16     /* load / set_state: coordinate outside the map, start on an obstacle, duplicate start, bad slot id */
#define MAPF_ERRBIT_INDEX 32     /* sum tree: leaf index outside [0, capacity) */
