// Host-side handle of the molecular Hamiltonian tables (fries_mol of include/fries_b200.h).
#pragma once
#include "mol.cuh"

struct fries_mol {
    fries_ctx *ctx = nullptr;
    MolView view;  // device pointers
    DevBuf<double> eris, hcore, blob;
    size_t n_packed = 0;
};

// Stage the small-table blob of `g` in shared memory and return a view bound to the copy.
// sh must hold g.d.blob_doubles doubles.  All threads of the CTA must call this.
#ifdef __CUDACC__
__device__ __forceinline__ MolView mol_stage_shared(const MolView &g, double *sh) {
    const double *src = g.d_diff - g.d.off_d_diff;
    for (unsigned i = threadIdx.x; i < g.d.blob_doubles; i += blockDim.x) sh[i] = src[i];
    __syncthreads();
    MolView m = g;
    mol_bind_blob(m, sh);
    return m;
}
#endif
