// micro-benchmark: cost of grid.sync() and of the engine's grid_reduce at the engine's launch shapes
#include "../../fries_b200/csrc/compress.cuh"
#include <cstdio>
void fries_set_error(const char *, ...) {}
__global__ void __launch_bounds__(512) k_sync(int iters) {
    cg::grid_group grid = cg::this_grid();
    for (int i = 0; i < iters; i++) grid.sync();
}
__global__ void __launch_bounds__(512) k_reduce(int iters, double *pd, unsigned long long *pc, double *out) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh_d[34];
    __shared__ unsigned long long sh_c[34];
    GridRed red{pd, pc, 0, (int)gridDim.x, sh_d, sh_c};
    double acc = 0;
    for (int i = 0; i < iters; i++) {
        double d = threadIdx.x * 1e-3 + i;
        unsigned long long c = 1;
        grid_reduce(grid, red, d, c);
        acc += d;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *out = acc;
}
__global__ void __launch_bounds__(512) k_pass(int iters, const double *v, size_t n, double *pd, unsigned long long *pc, double *out) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh_d[34];
    __shared__ unsigned long long sh_c[34];
    GridRed red{pd, pc, 0, (int)gridDim.x, sh_d, sh_c};
    size_t chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + 31) & ~(size_t)31;
    size_t lo = (size_t)blockIdx.x * chunk < n ? (size_t)blockIdx.x * chunk : n, hi = lo + chunk < n ? lo + chunk : n;
    double thr = 1e300, acc = 0;
    for (int it = 0; it < iters; it++) {
        double d = 0;
        unsigned long long c = 0;
        for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            double m = fabs(v[i]);
            if (m >= thr) { d += m; c++; }
        }
        grid_reduce(grid, red, d, c);
        thr = thr * 0.5 + d;
        acc += d;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *out = acc;
}
int main() {
    double *pd, *out, *v;
    unsigned long long *pc;
    size_t n = 260000;
    cudaMalloc(&pd, 1 << 16); cudaMalloc(&pc, 1 << 16); cudaMalloc(&out, 8); cudaMalloc(&v, n * 8);
    cudaMemset(v, 0, n * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid : {148, 296}) {
        int iters = 1000;
        float ms;
        void *a1[] = {&iters};
        cudaLaunchCooperativeKernel((void *)k_sync, dim3(grid), dim3(512), a1, 0, 0);
        cudaEventRecord(e0); cudaLaunchCooperativeKernel((void *)k_sync, dim3(grid), dim3(512), a1, 0, 0); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("grid %d: grid.sync %.2f us\n", grid, ms * 1000 / iters);
        void *a2[] = {&iters, &pd, &pc, &out};
        cudaLaunchCooperativeKernel((void *)k_reduce, dim3(grid), dim3(512), a2, 0, 0);
        cudaEventRecord(e0); cudaLaunchCooperativeKernel((void *)k_reduce, dim3(grid), dim3(512), a2, 0, 0); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("grid %d: grid_reduce %.2f us\n", grid, ms * 1000 / iters);
        void *a3[] = {&iters, &v, &n, &pd, &pc, &out};
        cudaLaunchCooperativeKernel((void *)k_pass, dim3(grid), dim3(512), a3, 0, 0);
        cudaEventRecord(e0); cudaLaunchCooperativeKernel((void *)k_pass, dim3(grid), dim3(512), a3, 0, 0); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("grid %d: pass(260k)+grid_reduce %.2f us  (%s)\n", grid, ms * 1000 / iters, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
