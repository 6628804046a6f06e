// Phase timing of the sum-tree kernels (GPU box): compiles the product's kernel file with MAPF_PER_TIMING, so thread 0 of the
// single CTA leaves clock64 stamps at the phase boundaries, and prints the deltas of one warm launch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/per_phase_probe profiles/tools/per_phase_probe.cu && /tmp/per_phase_probe
#define MAPF_PER_TIMING
#include "../../mapf_rl_b200/csrc/mapf_per_kernels.cu"

#include <cstdio>
#include <random>
#include <vector>

void mapf_set_error(const std::string &msg) { fprintf(stderr, "error: %s\n", msg.c_str()); }
int mapf_cuda_fail(cudaError_t e, const char *what)
{
    fprintf(stderr, "CUDA error %s in %s\n", cudaGetErrorString(e), what);
    return -2;
}

static void dump(const char *name, int first, int last)
{
    unsigned long long h[96];
    cudaMemcpyFromSymbol(h, g_per_stamps, sizeof(h));
    printf("{\"kernel\": \"%s\", \"cycles\": [", name);
    for (int i = first + 1; i <= last; ++i) printf("%s%lld", i > first + 1 ? ", " : "", (long long)(h[i] - h[i - 1]));
    printf("], \"total\": %lld}\n", (long long)(h[last] - h[first]));
}

int main()
{
    const int64_t cap = 1 << 19;
    const int layer = 20, n = 192;
    std::mt19937_64 rng(1);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::vector<double> tree(2 * cap - 1);
    for (int64_t i = 0; i < cap; ++i) tree[cap - 1 + i] = U(rng);
    for (int64_t i = cap - 2; i >= 0; --i) tree[i] = tree[2 * i + 1] + tree[2 * i + 2];
    mapf_per t{};
    t.capacity = cap, t.layer = layer, t.device = 0;
    cudaMalloc(&t.tree, tree.size() * 8);
    cudaMemcpy(t.tree, tree.data(), tree.size() * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&t.scratch.stamps, cap * 8);
    cudaMemset(t.scratch.stamps, 0, cap * 8);
    cudaMalloc(&t.scratch.prio32, 4096 * 4);
    cudaMalloc(&t.scratch.active, 4096);
    cudaMalloc(&t.scratch.err, 4);
    cudaMemset(t.scratch.err, 0, 4);
    t.scratch.cap_n = 4096;
    std::vector<int64_t> idx(256);
    std::vector<double> pr(256), u(256);
    for (int i = 0; i < 256; ++i) idx[i] = (int64_t)((i + U(rng)) * (cap / 256)), pr[i] = U(rng), u[i] = U(rng);  // stratified = sorted
    int64_t *d_idx, *d_sidx;
    double *d_pr, *d_u, *d_spr;
    float *d_w;
    cudaMalloc(&d_idx, 256 * 8), cudaMalloc(&d_sidx, 256 * 8), cudaMalloc(&d_pr, 256 * 8), cudaMalloc(&d_u, 256 * 8);
    cudaMalloc(&d_spr, 256 * 8), cudaMalloc(&d_w, 256 * 4);
    cudaMemcpy(d_idx, idx.data(), 256 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_pr, pr.data(), 256 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d_u, u.data(), 256 * 8, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        for (int i = 0; i < 20; ++i) mapf_launch_per_update(&t, &t.scratch, d_idx, d_pr, n, 0);
        cudaEventRecord(e0);
        for (int i = 0; i < 100; ++i) mapf_launch_per_update(&t, &t.scratch, d_idx, d_pr, n, 0);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1);
        printf("{\"update_192_us_back_to_back\": %.2f}\n", ms * 10);
        dump("update sorted 192: start | idx | top copy + barrier, adoption, sibling loads issued, sortedness | groups, level 0, publish | 19 levels", 0, 4 + layer - 2);
        for (int i = 0; i < 20; ++i) mapf_launch_per_sample(&t, d_u, n, d_sidx, d_spr, d_w, 0.4, 0);
        cudaEventRecord(e0);
        for (int i = 0; i < 100; ++i) mapf_launch_per_sample(&t, d_u, n, d_sidx, d_spr, d_w, 0.4, 0);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&ms, e0, e1);
        printf("{\"sample_192_us_back_to_back\": %.2f}\n", ms * 10);
        dump("sample 192: start | top copy + barrier | 10 cached levels | 3 groups", 40, 45);
        dump("sample 192: stores | weights", 45, 47);
    }
    // an empty kernel for the launch floor
    return 0;
}
