// micro-benchmark: bracket_solve (compress.cuh) in isolation, clock64 per phase
#define FR_BRACKET_TIMING
#include "../../fries_b200/csrc/compress.cuh"
#include <cstdio>
#include <vector>
void fries_set_error(const char *, ...) {}
__global__ void __launch_bounds__(512, 2) k_solve(CandList cl, unsigned long long *gacc, double R0, long long nrem0, double t_lo, double t_hi,
                                                   long long *cyc, double *out) {
    __shared__ double shd[6 * 33];
    __shared__ unsigned long long shc[6 * 33];
    cg::grid_group grid = cg::this_grid();
    grid.sync();
    long long t0 = clock64();
    CommView cm{};
    cm.n_ranks = 1;
    BracketResult br = bracket_solve(grid, cl, gacc, R0, nrem0, t_lo, t_hi, shd, shc, cm, nullptr, true);
    long long t1 = clock64();
    if (threadIdx.x == 0) {
        cyc[blockIdx.x] = t1 - t0;
        if (blockIdx.x == 0) {
            out[0] = br.valid; out[1] = br.x_cut; out[2] = br.R; out[3] = br.nrem; out[4] = (double)br.kept_cand; out[5] = br.rounds;
        }
    }
}
int run(int nc) {
    std::vector<double> x(FR_CAND_GCAP);
    std::vector<uint32_t> m(FR_CAND_GCAP, 1);
    double T = 1.0, h = 0.01;
    for (int i = 0; i < nc; i++) x[i] = T * (1 - h) + 2 * h * T * ((i * 7919) % nc) / nc;
    // state: 100000 budget left, R0 chosen so that the fixed point is ~T
    long long nrem0 = 100000 + 50ll * nc;
    double R0 = T * nrem0 * 1.002;
    CandList cl;
    unsigned long long cnt = nc;
    cudaMalloc(&cl.x, FR_CAND_GCAP * 8); cudaMalloc(&cl.mult, FR_CAND_GCAP * 4); cudaMalloc(&cl.count, 8);
    unsigned long long *gacc; cudaMalloc(&gacc, 40 * 8);
    cudaMemcpy(cl.x, x.data(), FR_CAND_GCAP * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(cl.mult, m.data(), FR_CAND_GCAP * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(cl.count, &cnt, 8, cudaMemcpyHostToDevice);
    long long *cyc; double *out;
    cudaMalloc(&cyc, 296 * 8); cudaMalloc(&out, 64);
    for (int rep = 0; rep < 3; rep++) {
        cudaMemset(gacc, 0, 40 * 8);
        double tl = T * (1 - h), th = T * (1 + h);
        void *args[] = {&cl, &gacc, &R0, &nrem0, &tl, &th, &cyc, &out};
        cudaLaunchCooperativeKernel((void *)k_solve, dim3(296), dim3(512), args, 0, 0);
        cudaDeviceSynchronize();
        long long hc[296]; double ho[6];
        cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
        cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost);
        long long mx = 0, mn = 1ll << 60;
        for (int i = 0; i < 296; i++) { mx = hc[i] > mx ? hc[i] : mx; mn = hc[i] < mn ? hc[i] : mn; }
        printf("bracket_solve: cycles min %lld max %lld (%.2f us at 1.965 GHz)  valid %g x_cut %.6f R %.3f nrem %g kept %g rounds %g  err %s\n",
               mn, mx, mx / 1965.0, ho[0], ho[1], ho[2], ho[3], ho[4], ho[5], cudaGetErrorString(cudaGetLastError()));
        long long bt[16];
        cudaMemcpyFromSymbol(bt, fr_bt, sizeof(bt));
        printf("  phases (cycles):");
        for (int k = 1; k < 8; k++) printf(" %lld", bt[k] - bt[k - 1]);
        printf("  | round 0: work %lld atomics %lld sync %lld loads %lld", bt[8] - bt[1], bt[9] - bt[8], bt[10] - bt[9], bt[11] - bt[10]);
        printf("\n");
    }
    return 0;
}
int main() {
    for (int nc : {2000, 20000, 200000}) { printf("== %d candidates\n", nc); run(nc); }
    return 0;
}
