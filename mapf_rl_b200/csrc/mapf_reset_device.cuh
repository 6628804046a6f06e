// mapf_reset_device.cuh — device-side instance generation for ONE environment by ONE warp: Environment.reset / __init__
// without the host (environment.py:100-138 == :156-192, map_partition :21-70).  Called by reset_kernel
// (mapf_reset_kernels.cu) and by the persistent rollout kernel when an episode ends inside the launch
// (mapf_rollout_kernels.cu, worker.py:390,422-428).
//
// Lane = map row (rows lane + 32 q, RPL rows per lane), each row RW words of padded column bits -- the same bit-parallel
// representation the BFS uses, so nothing but registers and shuffles is needed:
//   * obstacle map: iid Bernoulli(density) per cell; density < 0 draws triangular(0, 0.33, 0.5) once per
//     environment (environment.py:100,156)
//   * components with >= 2 cells (map_partition keeps only those, :65-67) = free cells that have a free
//     4-neighbour; the map is redrawn while there are none (:107-110)
//   * agent i: start = uniform over all cells of components that still have >= 2 unused cells (the
//     reference picks a component with probability proportional to its remaining size and then a uniform
//     cell in it, :120-131 -- the same distribution), goal = uniform over the unused cells of the same
//     component (:133-135); a component left with < 2 unused cells is dropped (:137).  The component of
//     a start cell is found by a bit-parallel flood fill, cached between agents (almost every start
//     falls in the giant component).
// Randomness is counter-based (Philox4x32-10): the instance with global index g uses key (seed) and counters
// (g, purpose, index), so any sharding of a batch over GPUs draws the instances a single GPU would.  The RNG stream of the
// reference (numpy / random globals) is NOT reproduced: parity with the reference goes through Environment.load; this
// generator is tested distributionally.
#pragma once
#include "mapf_common.cuh"

namespace {


__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

enum : uint32_t { PURPOSE_DENSITY = 0, PURPOSE_MAP = 1, PURPOSE_AGENT = 2 };

template <int RW, int RPL>
struct Bits {
    uint32_t v[RPL][RW];
};

// 4-neighbour dilation of a row-per-lane bitmap (the wave step of the BFS kernel)
template <int RW, int RPL>
__device__ __forceinline__ void dilate(const Bits<RW, RPL> &f, Bits<RW, RPL> &out, int lane)
{
#pragma unroll
    for (int q = 0; q < RPL; ++q)
#pragma unroll
        for (int w = 0; w < RW; ++w) {
            const uint32_t x = f.v[q][w];
            uint32_t fl = x << 1, fr = x >> 1;
            if (w > 0) fl |= f.v[q][w > 0 ? w - 1 : 0] >> 31;
            if (w < RW - 1) fr |= f.v[q][w < RW - 1 ? w + 1 : w] << 31;
            uint32_t up = __shfl_up_sync(MAPF_FULL_MASK, x, 1);
            uint32_t dn = __shfl_down_sync(MAPF_FULL_MASK, x, 1);
            uint32_t wrap_up = 0, wrap_dn = 0;
            if (q > 0) wrap_up = __shfl_sync(MAPF_FULL_MASK, f.v[q > 0 ? q - 1 : 0][w], 31);
            if (q < RPL - 1) wrap_dn = __shfl_sync(MAPF_FULL_MASK, f.v[q < RPL - 1 ? q + 1 : q][w], 0);
            if (lane == 0) up = wrap_up;
            if (lane == 31) dn = wrap_dn;
            out.v[q][w] = fl | fr | up | dn;
        }
}

template <int RW, int RPL>
__device__ __forceinline__ int count_bits(const Bits<RW, RPL> &m)
{
    int c = 0;
#pragma unroll
    for (int q = 0; q < RPL; ++q)
#pragma unroll
        for (int w = 0; w < RW; ++w) c += __popc(m.v[q][w]);
    return c;
}

// Uniform pick among the set bits of a warp-wide bitmap: returns (row << 8) | padded column, or -1 if the
// bitmap is empty.  `total_out` receives the number of set bits.
template <int RW, int RPL>
__device__ __forceinline__ int pick_bit(const Bits<RW, RPL> &m, uint32_t rnd, int lane, int &total_out)
{
    const int c = count_bits(m);
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(MAPF_FULL_MASK, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(MAPF_FULL_MASK, incl, 31);
    total_out = total;
    if (total == 0) return -1;
    const int r = (int)__umulhi(rnd, (uint32_t)total);
    const int excl = incl - c;
    const bool mine = r >= excl && r < incl;
    int code = 0;
    if (mine) {
        int local = r - excl;
        bool found = false;
#pragma unroll
        for (int q = 0; q < RPL; ++q)
#pragma unroll
            for (int w = 0; w < RW; ++w) {
                const int pc = __popc(m.v[q][w]);
                if (!found && local < pc) {
                    const int bit = __fns(m.v[q][w], 0, local + 1);
                    code = ((lane + 32 * q) << 8) | (32 * w + bit);
                    found = true;
                }
                if (!found) local -= pc;
            }
    }
    const int owner = __ffs(__ballot_sync(MAPF_FULL_MASK, mine)) - 1;
    return __shfl_sync(MAPF_FULL_MASK, code, owner);
}

template <int RW, int RPL>
__device__ __forceinline__ void clear_cell(Bits<RW, RPL> &m, int code, int lane)
{
    const int row = code >> 8, col = code & 0xff;
#pragma unroll
    for (int q = 0; q < RPL; ++q)
#pragma unroll
        for (int w = 0; w < RW; ++w)
            if (lane + 32 * q == row && (col >> 5) == w) m.v[q][w] &= ~(1u << (col & 31));
}

template <int RW, int RPL>
__device__ __forceinline__ bool test_cell(const Bits<RW, RPL> &m, int code, int lane)
{
    const int row = code >> 8, col = code & 0xff;
    bool hit = false;
#pragma unroll
    for (int q = 0; q < RPL; ++q)
#pragma unroll
        for (int w = 0; w < RW; ++w)
            if (lane + 32 * q == row && (col >> 5) == w) hit = (m.v[q][w] >> (col & 31)) & 1u;
    return __any_sync(MAPF_FULL_MASK, hit);
}

// All 32 lanes of the warp call together.  Writes obst / pos / goal of environment e and steps[e] = 0.
template <int RW, int RPL>
__device__ __forceinline__ void reset_env_warp(const EnvDims &d, const int e, const uint64_t seed, const uint64_t g,
                                               const float density, uint32_t *__restrict__ obst, uint8_t *__restrict__ pos,
                                               uint8_t *__restrict__ goal, int32_t *__restrict__ steps, int32_t *__restrict__ err)
{
    const int lane = threadIdx.x & 31;
    const int L = d.L, N = d.N;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t g_lo = (uint32_t)g, g_hi = (uint32_t)(g >> 32);

    // environment.py:100 / :156 — one density per environment
    float dens = density;
    if (density < 0.f) {
        const uint4 r = philox4x32_10(make_uint4(g_lo, g_hi, PURPOSE_DENSITY, 0u), key);
        const float u = (float)(r.x >> 8) * (1.0f / 16777216.0f);  // [0, 1)
        // inverse CDF of triangular(a = 0, c = 0.33, b = 0.5): F(c) = (c - a) / (b - a) = 0.66
        dens = u < 0.66f ? sqrtf(u * 0.5f * 0.33f) : 0.5f - sqrtf((1.0f - u) * 0.5f * 0.17f);
    }
    const uint32_t thresh = dens <= 0.f ? 0u : (dens >= 1.f ? 0xffffffffu : (uint32_t)((double)dens * 4294967296.0));

    // `comp` caches the component of the last start cell, `big` the largest component filled so far: almost every start falls
    // into the giant component, but every start that does not used to evict it from the (single) cache and the next agent
    // paid for another ~70-iteration flood fill of the giant component -- half of the generator's instructions at density 0.3
    Bits<RW, RPL> fre, elig, comp, big;
    int big_n = 0;
    for (uint32_t attempt = 0;; ++attempt) {
        if (attempt >= 256) {  // practically unreachable: e.g. density ~1 or more agents than usable cells
            if (lane == 0) atomicOr(err, MAPF_ERRBIT_RESET);
            break;
        }
        // ---- obstacle map: one Philox call per 4 cells ----
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
            const int row = lane + 32 * q;
#pragma unroll
            for (int w = 0; w < RW; ++w) fre.v[q][w] = 0;
            if (row < L) {
                for (int y4 = 0; y4 < L; y4 += 4) {
                    const uint4 r = philox4x32_10(make_uint4(g_lo, g_hi, PURPOSE_MAP | (attempt << 8), (uint32_t)(row * 64 + (y4 >> 2))), key);
                    // the four cells as one nibble at padded column y4 + 4 (a multiple of 4: it never straddles two words)
                    uint32_t nib = (r.x >= thresh ? 1u : 0u) | (r.y >= thresh ? 2u : 0u) | (r.z >= thresh ? 4u : 0u) | (r.w >= thresh ? 8u : 0u);
                    if (y4 + 4 > L) nib &= (1u << (L - y4)) - 1u;
                    const int p = y4 + 4;
#pragma unroll
                    for (int w = 0; w < RW; ++w)
                        if ((p >> 5) == w) fre.v[q][w] |= nib << (p & 31);
                }
            }
        }
        // ---- cells of components with >= 2 cells: free with a free neighbour ----
        dilate(fre, elig, lane);
#pragma unroll
        for (int q = 0; q < RPL; ++q)
#pragma unroll
            for (int w = 0; w < RW; ++w) {
                elig.v[q][w] &= fre.v[q][w];
                comp.v[q][w] = 0;
                big.v[q][w] = 0;
            }
        big_n = 0;
        // ---- agents ----
        bool ok = true;
        for (int i = 0; i < N; ++i) {
            const uint4 r = philox4x32_10(make_uint4(g_lo, g_hi, PURPOSE_AGENT | (attempt << 8), (uint32_t)i), key);
            int total;
            const int s = pick_bit(elig, r.x, lane, total);
            if (s < 0) {  // the reference regenerates the map (:107-110) / would raise mid-way; we redraw
                ok = false;
                break;
            }
            if (!test_cell(comp, s, lane) && big_n > 0 && test_cell(big, s, lane)) {
#pragma unroll
                for (int q = 0; q < RPL; ++q)
#pragma unroll
                    for (int w = 0; w < RW; ++w) comp.v[q][w] = big.v[q][w];
            } else if (!test_cell(comp, s, lane)) {
                // flood fill the component of s over the full free map
#pragma unroll
                for (int q = 0; q < RPL; ++q)
#pragma unroll
                    for (int w = 0; w < RW; ++w) comp.v[q][w] = 0;
                {
                    const int row = s >> 8, col = s & 0xff;
#pragma unroll
                    for (int q = 0; q < RPL; ++q)
#pragma unroll
                        for (int w = 0; w < RW; ++w)
                            if (lane + 32 * q == row && (col >> 5) == w) comp.v[q][w] = 1u << (col & 31);
                }
                for (;;) {
                    Bits<RW, RPL> nb;
                    dilate(comp, nb, lane);
                    uint32_t grew = 0;
#pragma unroll
                    for (int q = 0; q < RPL; ++q)
#pragma unroll
                        for (int w = 0; w < RW; ++w) {
                            const uint32_t add = nb.v[q][w] & fre.v[q][w] & ~comp.v[q][w];
                            comp.v[q][w] |= add;
                            grew |= add;
                        }
                    if (!__any_sync(MAPF_FULL_MASK, grew != 0)) break;
                }
                const int n = __reduce_add_sync(MAPF_FULL_MASK, count_bits(comp));
                if (n > big_n) {
                    big_n = n;
#pragma unroll
                    for (int q = 0; q < RPL; ++q)
#pragma unroll
                        for (int w = 0; w < RW; ++w) big.v[q][w] = comp.v[q][w];
                }
            }
            clear_cell(elig, s, lane);
            Bits<RW, RPL> cand;
#pragma unroll
            for (int q = 0; q < RPL; ++q)
#pragma unroll
                for (int w = 0; w < RW; ++w) cand.v[q][w] = elig.v[q][w] & comp.v[q][w];
            int cnt;
            const int t = pick_bit(cand, r.y, lane, cnt);  // cnt >= 1: the component was eligible
            clear_cell(elig, t, lane);
            if (cnt - 1 < 2) {  // environment.py:137
#pragma unroll
                for (int q = 0; q < RPL; ++q)
#pragma unroll
                    for (int w = 0; w < RW; ++w) elig.v[q][w] &= ~comp.v[q][w];
            }
            if (lane == 0) {
                const size_t o = ((size_t)e * N + i) * 2;
                pos[o] = (uint8_t)(s >> 8);
                pos[o + 1] = (uint8_t)((s & 0xff) - 4);
                goal[o] = (uint8_t)(t >> 8);
                goal[o + 1] = (uint8_t)((t & 0xff) - 4);
            }
        }
        if (ok) break;
    }

    // ---- store the padded obstacle bitmap (border rows / spare word zero), steps = 0 ----
    uint32_t *ob = obst + (size_t)e * d.obst_stride;
    for (int w = lane; w < d.obst_stride; w += 32) {
        const int prow = w / d.RWS;
        if (prow < 4 || prow >= L + 4 || w - prow * d.RWS >= RW) ob[w] = 0;
    }
#pragma unroll
    for (int q = 0; q < RPL; ++q) {
        const int row = lane + 32 * q;
        if (row < L) {
#pragma unroll
            for (int w = 0; w < RW; ++w) {
                const int lo = max(4 - 32 * w, 0), hi = min(L + 4 - 32 * w, 32);
                uint32_t cm = 0;
                if (hi > lo) cm = (hi - lo == 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
                ob[(row + 4) * d.RWS + w] = ~fre.v[q][w] & cm;
            }
        }
    }
    if (lane == 0) steps[e] = 0;
}

}  // namespace
