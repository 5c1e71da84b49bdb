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

// The same with the bulk asynchronous copy engine of sm_90+/sm_100 (TMA without a tensor map: cp.async.bulk global ->
// shared, completion counted in bytes on an mbarrier): one elected thread issues ONE copy of the whole blob, nobody
// spends load / store issue slots on it, and the CTA's warps wait on the barrier's phase.  `sh` must be 16-byte aligned
// and the blob a multiple of 16 bytes (fries_mol_create pads it); otherwise the plain loop above is used.
__device__ __forceinline__ MolView mol_stage_shared_bulk(const MolView &g, double *sh) {
    const double *src = g.d_diff - g.d.off_d_diff;
    const unsigned bytes = g.d.blob_doubles * 8u;
    __shared__ __align__(8) unsigned long long mbar;
    const unsigned mbar_a = (unsigned)__cvta_generic_to_shared(&mbar), dst_a = (unsigned)__cvta_generic_to_shared(sh);
    if ((bytes & 15u) || (dst_a & 15u) || ((size_t)src & 15u)) return mol_stage_shared(g, sh);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_a));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_a), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_a),
                     "l"(src), "r"(bytes), "r"(mbar_a)
                     : "memory");
    }
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(mbar_a)
            : "memory");
    }
    MolView m = g;
    mol_bind_blob(m, sh);
    return m;
}
#endif
