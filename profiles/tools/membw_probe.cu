// membw_probe.cu — what DRAM rate can a WRITE-dominated stream reach on this B200?
// The step kernel writes 486 B and reads ~128 B per agent-step, so the copy peak (half reads) is not its
// ceiling.  This probe times plain streams with the store flavours available on sm_100a over a ring that is
// much larger than the 126 MB L2, with CUDA events, and prints GB/s per variant (one JSON line each).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o membw_probe membw_probe.cu
//   ./membw_probe [ring_MB=1024] [iters=20]
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                            \
    do {                                                                                 \
        cudaError_t e_ = (x);                                                            \
        if (e_ != cudaSuccess) {                                                         \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                     \
        }                                                                                \
    } while (0)

enum StoreKind { ST_DEFAULT, ST_CS, ST_WT, ST_NOALLOC_V8, ST_EVICT_FIRST };

template <int KIND>
__device__ __forceinline__ void store16(uint4 *p, uint4 v, uint64_t pol)
{
    if constexpr (KIND == ST_DEFAULT) *p = v;
    else if constexpr (KIND == ST_CS) __stcs(p, v);
    else if constexpr (KIND == ST_WT) __stwt(p, v);
    else if constexpr (KIND == ST_EVICT_FIRST)
        asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                     "r"(v.w), "l"(pol)
                     : "memory");
}

// one uint4 per thread per iteration, warp-contiguous 512 B, CTA-contiguous; `per_thread` stores per thread
template <int KIND>
__global__ void write16(uint4 *dst, size_t n16, int per_thread)
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    size_t base = ((size_t)blockIdx.x * per_thread) * blockDim.x + threadIdx.x;
    const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 1, 2);
    for (int i = 0; i < per_thread; ++i) {
        size_t k = base + (size_t)i * blockDim.x;
        if (k < n16) store16<KIND>(dst + k, v, pol);
    }
}

// 32 B per thread: 256-bit store (sm_100+)
__global__ void write32(uint4 *dst, size_t n32, int per_thread)
{
    size_t base = ((size_t)blockIdx.x * per_thread) * blockDim.x + threadIdx.x;
    const uint32_t a = threadIdx.x, b = blockIdx.x;
    for (int i = 0; i < per_thread; ++i) {
        size_t k = base + (size_t)i * blockDim.x;
        if (k < n32)
            asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 2 * k), "r"(a), "r"(b), "r"(a),
                         "r"(b), "r"(a), "r"(b), "r"(a), "r"(b)
                         : "memory");
    }
}

// like the step kernel: every WARP owns one contiguous block of `blk` bytes (15552 there) and writes it with
// 16-byte stores, 512 B per warp instruction
template <int KIND>
__global__ void write_warp_blocks(uint8_t *dst, size_t nblocks, int blk)
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const int lane = threadIdx.x & 31;
    const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= nblocks) return;
    uint4 *p = reinterpret_cast<uint4 *>(dst + w * (size_t)blk);
    const uint4 v = make_uint4(lane, (uint32_t)w, 1, 2);
    const int n = blk >> 4;
#pragma unroll 4
    for (int c = lane; c < n; c += 32) store16<KIND>(p + c, v, pol);
}

// TMA bulk store: every warp fills a `chunk`-byte staging buffer in shared memory and hands it to the copy
// engine (cp.async.bulk shared::cta -> global), double buffered
__global__ void write_bulk(uint8_t *dst, size_t nblocks, int blk, int chunk)
{
    extern __shared__ __align__(128) uint8_t sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (w >= nblocks) return;
    uint8_t *stage = sm + (size_t)warp * 2 * chunk;
    uint8_t *out = dst + w * (size_t)blk;
    int buf = 0;
    for (int off = 0; off < blk; off += chunk) {
        const int len = min(chunk, blk - off);
        uint8_t *s = stage + buf * chunk;
        // wait until the bulk store issued two rounds ago (same buffer) has finished READING shared memory
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        for (int c = lane * 16; c < len; c += 512) *reinterpret_cast<uint4 *>(s + c) = make_uint4(lane, off, 1, 2);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + off),
                         "r"((uint32_t)__cvta_generic_to_shared(s)), "r"(len)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        buf ^= 1;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
}

__global__ void read16(const uint4 *src, size_t n16, int per_thread, uint32_t *sink)
{
    size_t base = ((size_t)blockIdx.x * per_thread) * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (int i = 0; i < per_thread; ++i) {
        size_t k = base + (size_t)i * blockDim.x;
        if (k < n16) {
            uint4 v = __ldcs(src + k);
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

__global__ void copy16(const uint4 *src, uint4 *dst, size_t n16, int per_thread)
{
    size_t base = ((size_t)blockIdx.x * per_thread) * blockDim.x + threadIdx.x;
    for (int i = 0; i < per_thread; ++i) {
        size_t k = base + (size_t)i * blockDim.x;
        if (k < n16) __stcs(dst + k, __ldcs(src + k));
    }
}

// the step kernel's mix: per warp one 15552-byte block written + 32 scattered 128-byte-line reads of `rd` bytes
__global__ void mix_warp_blocks(uint8_t *dst, const uint8_t *src, size_t nblocks, int blk, int rd, size_t src_lines, uint32_t *sink)
{
    const int lane = threadIdx.x & 31;
    const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= nblocks) return;
    // pseudo-random line per (warp, lane)
    size_t h = (w * 32 + lane) * 0x9E3779B97F4A7C15ull;
    const size_t line = (h >> 20) % src_lines;
    const uint2 *q = reinterpret_cast<const uint2 *>(src + line * 128);
    uint32_t acc = 0;
    for (int i = 0; i < rd / 8; ++i) {
        uint2 v = __ldg(q + i);
        acc ^= v.x ^ v.y;
    }
    uint4 *p = reinterpret_cast<uint4 *>(dst + w * (size_t)blk);
    const uint4 v = make_uint4(lane, acc, 1, 2);
    const int n = blk >> 4;
#pragma unroll 4
    for (int c = lane; c < n; c += 32) __stcs(p + c, v);
    if (acc == 0x12345678u) *sink = acc;
}

// L2 retention probe: every warp writes one 15552-byte block (streaming) and every lane reads 72 B of one
// 128-byte line out of a 33.5 MB set (8192 x 32 lines).  REUSE: the same set every launch (what the step kernel's
// current heuristic-map tiles are) or a fresh set per launch.  If L2 keeps the set across the 127 MB write
// stream, the reuse case is served from L2 and gets faster.
template <int STKIND, bool LDKEEP>
__global__ void l2_retention(uint8_t *dst, const uint8_t *src, size_t nblocks, int blk, size_t src_lines, uint64_t salt, uint32_t *sink)
{
    uint64_t pol_first, pol_last;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
    const int lane = threadIdx.x & 31;
    const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= nblocks) return;
    size_t h = ((w * 32 + lane) ^ salt) * 0x9E3779B97F4A7C15ull;
    // salt == ~0: lane a of warp w reads line 32 w + a (one contiguous 4 KB run per warp: a per-env tile cache)
    const size_t line = salt == ~0ull ? w * 32 + lane : (h >> 20) % src_lines;
    const uint2 *q = reinterpret_cast<const uint2 *>(src + line * 128);
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        uint2 v;
        if (LDKEEP) asm volatile("ld.global.nc.L2::cache_hint.v2.b32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(q + i), "l"(pol_last));
        else v = __ldg(q + i);
        acc ^= v.x ^ v.y;
    }
    uint4 *p = reinterpret_cast<uint4 *>(dst + w * (size_t)blk);
    const uint4 v = make_uint4(lane, acc, 1, 2);
    const int n = blk >> 4;
#pragma unroll 4
    for (int c = lane; c < n; c += 32) store16<STKIND>(p + c, v, pol_first);
    if (acc == 0x12345678u) *sink = acc;
}

// The split step kernel's consumer side in isolation: a persistent grid, CTA b writes the blocks
// b*P + w + r*grid*P (w < P, round r) one after another with CW warps.  BULK: the warps fill a shared-memory
// staging buffer (NBUF of them) and one thread issues a single cp.async.bulk per block; else direct streaming stores.
// SWEEP: block of (CTA b, slot w, round r) = r*grid*P + w*grid + b, so that all CTAs write neighbouring blocks at the
// same time (one compact front) instead of blocks P apart.
template <bool BULK, bool SWEEP = false>
__global__ void cta_blocks(uint8_t *dst, int nblocks, int blk, int P, int nbuf)
{
    extern __shared__ __align__(128) uint8_t sm[];
    const int n16 = blk >> 4;
    int use = 0;
    for (int e0 = SWEEP ? 0 : blockIdx.x * P; e0 < nblocks; e0 += gridDim.x * P) {
        for (int w = 0; w < P; ++w, ++use) {
            const int e = SWEEP ? e0 + w * gridDim.x + blockIdx.x : e0 + w;
            if (e >= nblocks) continue;
            uint8_t *out = dst + (size_t)e * blk;
            const uint4 v = make_uint4(threadIdx.x, e0, w, 2);
            if (BULK) {
                uint8_t *s = sm + (size_t)(use % nbuf) * blk;
                if (threadIdx.x == 0) {
                    if (nbuf == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    else if (nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
                }
                __syncthreads();
                for (int c = threadIdx.x; c < n16; c += blockDim.x) reinterpret_cast<uint4 *>(s)[c] = v;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncthreads();
                if (threadIdx.x == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out),
                                 "r"((uint32_t)__cvta_generic_to_shared(s)), "r"(blk)
                                 : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else {
                for (int c = threadIdx.x; c < n16; c += blockDim.x) __stcs(reinterpret_cast<uint4 *>(out) + c, v);
            }
        }
    }
    if (BULK && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

template <typename F>
static float time_ms(F &&launch, int iters)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) launch(i);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) launch(i);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / iters;
}

static void report(const char *name, double bytes, float ms)
{
    printf("{\"probe\": \"%s\", \"MB\": %.1f, \"us\": %.2f, \"GBs\": %.1f}\n", name, bytes / 1e6, ms * 1e3, bytes / (ms * 1e-3) / 1e9);
    fflush(stdout);
}

int main(int argc, char **argv)
{
    const size_t ring_mb = argc > 1 ? atol(argv[1]) : 1024;
    const int iters = argc > 2 ? atoi(argv[2]) : 20;
    const size_t slot = 128ull << 20;  // one "step" = 128 MiB, like the 127.4 MB observation block
    const size_t nslots = (ring_mb << 20) / slot;
    uint8_t *ring, *src;
    uint32_t *sink;
    CK(cudaMalloc(&ring, nslots * slot));
    CK(cudaMalloc(&src, 1ull << 30));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(ring, 1, nslots * slot));
    CK(cudaMemset(src, 2, 1ull << 30));
    const size_t n16 = slot / 16;
    auto slotp = [&](int i) { return ring + (size_t)(i % nslots) * slot; };

    for (int threads : {128, 256, 512}) {
        for (int per : {1, 4, 16}) {
            const unsigned grid = (unsigned)((n16 + (size_t)threads * per - 1) / ((size_t)threads * per));
            char nm[96];
            snprintf(nm, sizeof nm, "write16_default_t%d_x%d", threads, per);
            report(nm, slot, time_ms([&](int i) { write16<ST_DEFAULT><<<grid, threads>>>((uint4 *)slotp(i), n16, per); }, iters));
            snprintf(nm, sizeof nm, "write16_cs_t%d_x%d", threads, per);
            report(nm, slot, time_ms([&](int i) { write16<ST_CS><<<grid, threads>>>((uint4 *)slotp(i), n16, per); }, iters));
        }
    }
    {
        const int threads = 256, per = 4;
        const unsigned grid = (unsigned)((n16 + (size_t)threads * per - 1) / ((size_t)threads * per));
        report("write16_wt_t256_x4", slot, time_ms([&](int i) { write16<ST_WT><<<grid, threads>>>((uint4 *)slotp(i), n16, per); }, iters));
        report("write16_evict_first_t256_x4", slot,
               time_ms([&](int i) { write16<ST_EVICT_FIRST><<<grid, threads>>>((uint4 *)slotp(i), n16, per); }, iters));
        const unsigned grid32 = (unsigned)((n16 / 2 + (size_t)threads * per - 1) / ((size_t)threads * per));
        report("write32_v8_t256_x4", slot, time_ms([&](int i) { write32<<<grid32, threads>>>((uint4 *)slotp(i), n16 / 2, per); }, iters));
        report("memset", slot, time_ms([&](int i) { cudaMemsetAsync(slotp(i), i, slot); }, iters));
        report("read16_cs_t256_x4", slot,
               time_ms([&](int i) { read16<<<grid, threads>>>((const uint4 *)slotp(i), n16, per, sink); }, iters));
        report("copy16_t256_x4(read+write bytes)", 2.0 * slot,
               time_ms([&](int i) { copy16<<<grid, threads>>>((const uint4 *)slotp(i), (uint4 *)slotp(i + nslots / 2), n16, per); }, iters));
        report("memcpyD2D(read+write bytes)", 2.0 * slot,
               time_ms([&](int i) { cudaMemcpyAsync(slotp(i + nslots / 2), slotp(i), slot, cudaMemcpyDeviceToDevice); }, iters));
    }
    // the step kernel's shape: 8192 warps x 15552 B = 127.4 MB
    {
        const int blk = 15552;
        const size_t nblocks = 8192;
        for (int warps : {2, 4, 8}) {
            const unsigned grid = (unsigned)((nblocks + warps - 1) / warps);
            char nm[96];
            snprintf(nm, sizeof nm, "warp_blocks_cs_w%d", warps);
            report(nm, (double)nblocks * blk,
                   time_ms([&](int i) { write_warp_blocks<ST_CS><<<grid, warps * 32>>>(slotp(i), nblocks, blk); }, iters));
            snprintf(nm, sizeof nm, "warp_blocks_default_w%d", warps);
            report(nm, (double)nblocks * blk,
                   time_ms([&](int i) { write_warp_blocks<ST_DEFAULT><<<grid, warps * 32>>>(slotp(i), nblocks, blk); }, iters));
        }
        for (int chunk : {1296, 2592, 5184, 15552}) {
            for (int warps : {2, 4}) {
                const size_t smem = (size_t)warps * 2 * chunk;
                if (smem > 200 * 1024) continue;
                const unsigned grid = (unsigned)((nblocks + warps - 1) / warps);
                CK(cudaFuncSetAttribute(write_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                char nm[96];
                snprintf(nm, sizeof nm, "bulk_store_chunk%d_w%d", chunk, warps);
                report(nm, (double)nblocks * blk,
                       time_ms([&](int i) { write_bulk<<<grid, warps * 32, smem>>>(slotp(i), nblocks, blk, chunk); }, iters));
            }
        }
        const size_t src_lines = (1ull << 30) / 128;
        for (int rd : {0, 72, 128}) {
            const int warps = 4;
            const unsigned grid = (unsigned)((nblocks + warps - 1) / warps);
            char nm[96];
            snprintf(nm, sizeof nm, "mix_blocks_read%dB_per_lane", rd);
            // bytes: block written + per lane one DRAM line touched (64 B if rd <= 64 .. else 128 B)
            report(nm, (double)nblocks * (blk + 32.0 * rd),
                   time_ms([&](int i) { mix_warp_blocks<<<grid, warps * 32>>>(slotp(i), src, nblocks, blk, rd, src_lines, sink); }, iters));
        }
    }
    // consumer side of the split step kernel in isolation
    {
        const int blk = 15552, nblocks = 8192;
        CK(cudaFuncSetAttribute(cta_blocks<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * blk));
        for (int per_sm : {2, 3, 4, 6}) {
            for (int cw : {1, 3}) {
                const int P = 5, grid = 148 * per_sm;
                char nm[128];
                snprintf(nm, sizeof nm, "cta_blocks_stores_%dctas_per_sm_%dwarps", per_sm, cw);
                report(nm, (double)nblocks * blk, time_ms([&](int i) { cta_blocks<false><<<grid, cw * 32>>>(slotp(i), nblocks, blk, P, 1); }, iters));
                snprintf(nm, sizeof nm, "cta_blocks_SWEEP_stores_%dctas_per_sm_%dwarps", per_sm, cw);
                report(nm, (double)nblocks * blk, time_ms([&](int i) { cta_blocks<false, true><<<grid, cw * 32>>>(slotp(i), nblocks, blk, P, 1); }, iters));
                for (int nbuf : {3}) {
                    if ((size_t)per_sm * nbuf * blk > 220 * 1024) continue;
                    snprintf(nm, sizeof nm, "cta_blocks_bulk_%dctas_per_sm_%dwarps_%dbuf", per_sm, cw, nbuf);
                    report(nm, (double)nblocks * blk,
                           time_ms([&](int i) { cta_blocks<true><<<grid, cw * 32, (size_t)nbuf * blk>>>(slotp(i), nblocks, blk, P, nbuf); }, iters));
                }
            }
        }
    }
    // L2 retention of a 33.5 MB re-read set under the 127 MB write stream
    {
        const int blk = 15552, warps = 4;
        const size_t nblocks = 8192, src_lines = (1ull << 30) / 128;
        const unsigned grid = (unsigned)((nblocks + warps - 1) / warps);
        for (int persist_mb : {0, 48}) {
            CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)persist_mb << 20));
            for (int reuse : {0, 1}) {
                char nm[128];
                auto salt = [&](int i) { return reuse ? 0ull : (uint64_t)(i + 1) * 0x51ED2701ull; };
                snprintf(nm, sizeof nm, "l2ret_persist%dMB_reuse%d_st.cs_ld.plain", persist_mb, reuse);
                report(nm, (double)nblocks * (blk + 32.0 * 128),
                       time_ms([&](int i) { l2_retention<ST_CS, false><<<grid, warps * 32>>>(slotp(i), src, nblocks, blk, src_lines, salt(i), sink); }, iters));
                snprintf(nm, sizeof nm, "l2ret_persist%dMB_reuse%d_st.cs_ld.evict_last", persist_mb, reuse);
                report(nm, (double)nblocks * (blk + 32.0 * 128),
                       time_ms([&](int i) { l2_retention<ST_CS, true><<<grid, warps * 32>>>(slotp(i), src, nblocks, blk, src_lines, salt(i), sink); }, iters));
                snprintf(nm, sizeof nm, "l2ret_persist%dMB_reuse%d_st.evict_first_ld.evict_last", persist_mb, reuse);
                report(nm, (double)nblocks * (blk + 32.0 * 128),
                       time_ms([&](int i) { l2_retention<ST_EVICT_FIRST, true><<<grid, warps * 32>>>(slotp(i), src, nblocks, blk, src_lines, salt(i), sink); }, iters));
                snprintf(nm, sizeof nm, "l2ret_persist%dMB_reuse%d_st.default_ld.evict_last", persist_mb, reuse);
                report(nm, (double)nblocks * (blk + 32.0 * 128),
                       time_ms([&](int i) { l2_retention<ST_DEFAULT, true><<<grid, warps * 32>>>(slotp(i), src, nblocks, blk, src_lines, salt(i), sink); }, iters));
            }
        }
        CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0));
        report("l2ret_SEQUENTIAL_4KB_per_warp_st.cs_ld.plain", (double)nblocks * (blk + 32.0 * 128),
               time_ms([&](int i) { l2_retention<ST_CS, false><<<grid, warps * 32>>>(slotp(i), src + (size_t)(i % 8) * (64u << 20), nblocks, blk, src_lines, ~0ull, sink); }, iters));
    }
    return 0;
}
