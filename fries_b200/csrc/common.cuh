// Shared host/device plumbing for the fries_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <type_traits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <map>
#include <string>
#include <vector>
#include "../../include/fries_b200.h"

namespace cg = cooperative_groups;

#define FRIES_HASH_PRIME 1099511628211ull  // FRIES/det_hash.hpp:164
#define FRIES_EMPTY_KEY (~0ull)
#define FRIES_NO_POS 0xffffffffu
#define FRIES_OVF_POS 0xfffffffeu  // index entry of a determinant the full store could not place (merge_fused_kernel)

void fries_set_error(const char *fmt, ...);

#define CUDA_TRY(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            fries_set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                \
                            cudaGetErrorString(e__));                                           \
            return FRIES_ERR_CUDA;                                                              \
        }                                                                                       \
    } while (0)
#define FRIES_TRY(expr)                                                                         \
    do {                                                                                        \
        int r__ = (expr);                                                                       \
        if (r__ != FRIES_OK) return r__;                                                        \
    } while (0)
#define FRIES_REQUIRE(cond, ...)                                                                \
    do {                                                                                        \
        if (!(cond)) {                                                                          \
            fries_set_error(__VA_ARGS__);                                                       \
            return FRIES_ERR_ARG;                                                               \
        }                                                                                       \
    } while (0)

struct KernelStat {
    double ms = 0;
    uint64_t launches = 0;
};

struct fries_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    uint64_t launch_count = 0;
    bool profile = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::map<std::string, KernelStat> stats;
    // small pinned staging area for scalar read-backs
    double *h_pinned = nullptr;  // 64 doubles
    // generic device scratch, grown on demand
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;
    int ensure_scratch(size_t bytes);
    // cooperative grid size for a kernel (blocks) given block size and dynamic smem
    int coop_grid(const void *kernel, int block, size_t smem);
    // multi-rank: peer-mapped inboxes used by the *_dev compression entry points (fries_ctx_set_comm); not owned
    struct fries_comm *comm = nullptr;
};

// RAII-less profiling helpers: wrap a launch region
struct ProfScope {
    fries_ctx *ctx;
    const char *name;
    ProfScope(fries_ctx *c, const char *n) : ctx(c), name(n) {
        if (ctx->profile) cudaEventRecord(ctx->ev0, ctx->stream);
    }
    ~ProfScope() {
        if (ctx->profile) {
            cudaEventRecord(ctx->ev1, ctx->stream);
            cudaEventSynchronize(ctx->ev1);
            float ms = 0;
            cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
            KernelStat &s = ctx->stats[name];
            s.ms += ms;
            s.launches++;
        }
    }
};

template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    int alloc(size_t count) {
        free();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e != cudaSuccess) {
            fries_set_error("cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
            p = nullptr;
            return FRIES_ERR_CUDA;
        }
        n = count;
        return FRIES_OK;
    }
    int ensure(size_t count) { return count <= n ? FRIES_OK : alloc(count); }
    void free() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    ~DevBuf() { free(); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
};

// ---------------------------------------------------------------------------------------------------
// device bit utilities (restating FRIES/math_utils.c, FRIES/fci_utils.c, FRIES/det_store.h on u64 keys)
// ---------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int fr_popc(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
__host__ __device__ __forceinline__ int fr_ctz(uint64_t x) {
#ifdef __CUDA_ARCH__
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
// 32-bit overload: one FLO/BREV pair instead of the 64-bit emulation
__host__ __device__ __forceinline__ int fr_ctz(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
__host__ __device__ __forceinline__ bool fr_read_bit(uint64_t key, int b) { return (key >> b) & 1ull; }

// find_bits (math_utils.c:62-98): ascending list of set bits
__host__ __device__ __forceinline__ int fr_occ_list(uint64_t key, uint8_t *occ) {
    int n = 0;
    while (key) {
        occ[n++] = (uint8_t)fr_ctz(key);
        key &= key - 1;
    }
    return n;
}

// HashTable::hash_fxn det_hash.hpp:160-170: (i + 1) * scrambler wraps at 32 bits, the sum at 64
__host__ __device__ __forceinline__ uint64_t fr_det_hash(uint64_t key, const uint32_t *scr) {
    uint64_t h = 0;
    uint32_t i = 0;
    while (key) {
        int orb = fr_ctz(key);
        key &= key - 1;
        i++;
        h = FRIES_HASH_PRIME * h + (uint32_t)(i * scr[orb]);
    }
    return h;
}

// bits_between (math_utils.c:9-58): occupied bits strictly between a and b
__host__ __device__ __forceinline__ int fr_bits_between(uint64_t key, int a, int b) {
    int lo = a < b ? a : b, hi = a < b ? b : a;
    uint64_t mask = ((1ull << hi) - 1ull) & ~((2ull << lo) - 1ull);
    return fr_popc(key & mask);
}
// excite_sign fci_utils.c:128-135
__host__ __device__ __forceinline__ int fr_excite_sign(int cre, int des, uint64_t key) {
    return (fr_bits_between(key, cre, des) & 1) ? -1 : 1;
}
// sing_det_parity fci_utils.c:46-51
__host__ __device__ __forceinline__ int fr_sing_det_parity(uint64_t &key, int o, int v) {
    key &= ~(1ull << o);
    int s = fr_excite_sign(o, v, key);
    key |= 1ull << v;
    return s;
}
// doub_det_parity fci_utils.c:67-75
__host__ __device__ __forceinline__ int fr_doub_det_parity(uint64_t &key, int o0, int o1, int v2, int v3) {
    key &= ~((1ull << o0) | (1ull << o1));
    int s = fr_excite_sign(v2, o0, key) * fr_excite_sign(v3, o1, key);
    key |= (1ull << v2) | (1ull << v3);
    return s;
}
// sing_parity :54-57, doub_parity :86-94 (determinant unchanged)
__host__ __device__ __forceinline__ int fr_sing_parity(uint64_t key, int o, int v) { return fr_excite_sign(o, v, key); }
__host__ __device__ __forceinline__ int fr_doub_parity(uint64_t key, int o0, int o1, int v2, int v3) {
    key &= ~((1ull << o0) | (1ull << o1));
    return fr_excite_sign(v2, o0, key) * fr_excite_sign(v3, o1, key);
}

// ---------------------------------------------------------------------------------------------------
// block-level primitives (deterministic: fixed shuffle trees, no atomics on doubles)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// sum over the block; result valid in ALL threads.  sh must hold 33 doubles.
__device__ __forceinline__ double block_sum(double v, double *sh) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? sh[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long *sh) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum_u64(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        unsigned long long t = lane < nw ? sh[lane] : 0ull;
        t = warp_sum_u64(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}

// fused (double, u64) block sum: 3 barriers instead of 6.  shd / shc hold 33 entries each.
__device__ __forceinline__ void block_sum_pair(double &d, unsigned long long &c, double *shd, unsigned long long *shc) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        d += __shfl_down_sync(0xffffffffu, d, o);
        c += __shfl_down_sync(0xffffffffu, c, o);
    }
    __syncthreads();
    if (lane == 0) {
        shd[w] = d;
        shc[w] = c;
    }
    __syncthreads();
    if (w == 0) {
        double t = lane < nw ? shd[lane] : 0.0;
        unsigned long long u = lane < nw ? shc[lane] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            t += __shfl_down_sync(0xffffffffu, t, o);
            u += __shfl_down_sync(0xffffffffu, u, o);
        }
        if (lane == 0) {
            shd[32] = t;
            shc[32] = u;
        }
    }
    __syncthreads();
    d = shd[32];
    c = shc[32];
}
// every lane of the warp obtains the same sum (fixed butterfly: identical in every warp of every CTA)
__device__ __forceinline__ double warp_allsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_allsum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// fixed-order sum of `n` per-block partials, executed redundantly by every block after a grid sync;
// every block obtains the bit-identical result.
__device__ __forceinline__ double grid_partial_sum(const double *part, int n, double *sh) {
    double v = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += part[i];
    return block_sum(v, sh);
}
__device__ __forceinline__ unsigned long long grid_partial_sum_u64(const unsigned long long *part, int n,
                                                                   unsigned long long *sh) {
    unsigned long long v = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += part[i];
    return block_sum_u64(v, sh);
}

// Block-wide exclusive scan of a (double, u64) pair.  Returns the exclusive prefix for this thread and
// the block total through tot_*.  sh_d / sh_c must each hold 33 entries.
__device__ __forceinline__ void block_excl_scan(double a, unsigned long long c, double &ex_a, unsigned long long &ex_c,
                                                double &tot_a, unsigned long long &tot_c, double *sh_d,
                                                unsigned long long *sh_c) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    double ia = a;
    unsigned long long ic = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double ta = __shfl_up_sync(0xffffffffu, ia, o);
        unsigned long long tc = __shfl_up_sync(0xffffffffu, ic, o);
        if (lane >= o) {
            ia += ta;
            ic += tc;
        }
    }
    __syncthreads();
    if (lane == 31) {
        sh_d[w] = ia;
        sh_c[w] = ic;
    }
    __syncthreads();
    if (w == 0) {
        double wa = lane < nw ? sh_d[lane] : 0.0;
        unsigned long long wc = lane < nw ? sh_c[lane] : 0ull;
        double sa = wa;
        unsigned long long sc = wc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double ta = __shfl_up_sync(0xffffffffu, sa, o);
            unsigned long long tc = __shfl_up_sync(0xffffffffu, sc, o);
            if (lane >= o) {
                sa += ta;
                sc += tc;
            }
        }
        if (lane < nw) {
            sh_d[lane] = sa - wa;  // exclusive prefix of warp totals
            sh_c[lane] = sc - wc;
        }
        if (lane == 31) {
            sh_d[32] = sa;
            sh_c[32] = sc;
        }
    }
    __syncthreads();
    ex_a = sh_d[w] + (ia - a);
    ex_c = sh_c[w] + (ic - c);
    tot_a = sh_d[32];
    tot_c = sh_c[32];
    __syncthreads();
}

// Row generators call fr_emit(f, j, w) for every entry; a visitor may return void (visit everything) or bool
// (false = the visitor has all it needs: stop generating the row).
template <class F>
__host__ __device__ __forceinline__ bool fr_emit(F &&f, unsigned j, double w) {
    if constexpr (std::is_void<decltype(f(j, w))>::value) {
        f(j, w);
        return true;
    } else {
        return f(j, w);
    }
}

// The systematic-sampling grid g_k = rn0 + k * unit, k = 0 .. n - 1 (n = sampling budget).  The reference
// walks it with a running `rn_sys += unit` and the test `rn_sys < lbound` (compress_utils.cpp:313-318,
// 745-755); here an element asks how many grid points lie strictly below a bound.  Indices >= n do not
// exist (in exact arithmetic g_n >= the total weight), which keeps a prefix sum that is one ulp above the
// total from drawing an (n+1)-th sample.
struct SysGrid {
    double rn0, unit;
    double inv;  // 1 / unit: first guess of an index (the guess is then made consistent with the FP grid, see below)
    long long n;
    __device__ __forceinline__ double point(long long k) const {
        return k < n ? fma((double)k, unit, rn0) : INFINITY;
    }
    __device__ __forceinline__ long long count_below(double x) const {
        if (n <= 0 || !(x > rn0)) return 0;
        double q = (x - rn0) * inv;
        long long k = q >= (double)n ? n : (long long)ceil(q);
        if (k < 0) k = 0;
        if (k > n) k = n;
        // make the count consistent with the FP grid fl(rn0 + k*unit)
        while (k > 0 && !(fma((double)(k - 1), unit, rn0) < x)) k--;
        while (k < n && fma((double)k, unit, rn0) < x) k++;
        return k;
    }
    // The same two with the index kept as a double (exact below 2^53): no 64-bit integer <-> double conversions, which are
    // multi-instruction sequences on the device and sat in every input's path of the count / emit passes.
    __device__ __forceinline__ double point_d(double k) const { return k < (double)n ? fma(k, unit, rn0) : INFINITY; }
    __device__ __forceinline__ double count_below_d(double x) const {
        const double nd = (double)n;
        if (n <= 0 || !(x > rn0)) return 0.0;
        double q = (x - rn0) * inv;
        double k = q >= nd ? nd : ceil(q);
        if (k < 0) k = 0;
        if (k > nd) k = nd;
        while (k > 0 && !(fma(k - 1.0, unit, rn0) < x)) k -= 1.0;
        while (k < nd && fma(k, unit, rn0) < x) k += 1.0;
        return k;
    }
};
