// Context, error reporting, hash/owner kernel (a1) and batch bit utilities (a15).
#include "common.cuh"
#include <cstdarg>

static thread_local char g_err[1024] = "";

void fries_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *fries_last_error(void) { return g_err; }
extern "C" int fries_version(void) { return 100; }

int fries_ctx::ensure_scratch(size_t bytes) {
    if (bytes <= scratch_bytes) return FRIES_OK;
    if (d_scratch) cudaFree(d_scratch);
    d_scratch = nullptr;
    scratch_bytes = 0;
    size_t want = bytes + (bytes >> 2) + 4096;
    CUDA_TRY(cudaMalloc(&d_scratch, want));
    scratch_bytes = want;
    return FRIES_OK;
}

int fries_ctx::coop_grid(const void *kernel, int block, size_t smem) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1) {
        per_sm = 1;
    }
    if (per_sm > 2) per_sm = 2;
    return per_sm * sm_count;
}

extern "C" int fries_ctx_create(int device, fries_ctx **out) {
    FRIES_REQUIRE(out != nullptr, "fries_ctx_create: out is NULL");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        fries_set_error("fries_ctx_create: no CUDA device (%s); fries_b200 has no CPU fallback",
                        e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return FRIES_ERR_CUDA;
    }
    FRIES_REQUIRE(device >= 0 && device < n_dev, "fries_ctx_create: device %d out of range (%d devices)", device, n_dev);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        fries_set_error("fries_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                        prop.major, prop.minor);
        return FRIES_ERR_CUDA;
    }
    fries_ctx *c = new fries_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
    CUDA_TRY(cudaEventCreate(&c->ev0));
    CUDA_TRY(cudaEventCreate(&c->ev1));
    CUDA_TRY(cudaMallocHost(&c->h_pinned, 64 * sizeof(double)));
    *out = c;
    return FRIES_OK;
}

extern "C" int fries_ctx_destroy(fries_ctx *c) {
    if (!c) return FRIES_OK;
    cudaSetDevice(c->device);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->h_pinned) cudaFreeHost(c->h_pinned);
    if (c->d_scratch) cudaFree(c->d_scratch);
    delete c;
    return FRIES_OK;
}

extern "C" int fries_ctx_set_stream(fries_ctx *c, void *s) {
    FRIES_REQUIRE(c, "ctx is NULL");
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    if (s) {
        c->stream = (cudaStream_t)s;
        c->own_stream = false;
    } else {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    return FRIES_OK;
}

extern "C" int fries_ctx_sync(fries_ctx *c) {
    FRIES_REQUIRE(c, "ctx is NULL");
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}
extern "C" int fries_ctx_sm_count(fries_ctx *c) { return c ? c->sm_count : 0; }
extern "C" uint64_t fries_ctx_launch_count(fries_ctx *c) { return c ? c->launch_count : 0; }
extern "C" int fries_ctx_set_profile(fries_ctx *c, int on) {
    FRIES_REQUIRE(c, "ctx is NULL");
    c->profile = on != 0;
    if (on == 2) c->stats.clear();
    return FRIES_OK;
}
extern "C" int fries_ctx_kernel_ms(fries_ctx *c, const char *name, double *total_ms, uint64_t *launches) {
    FRIES_REQUIRE(c && name, "bad argument");
    auto it = c->stats.find(name);
    if (it == c->stats.end()) {
        if (total_ms) *total_ms = 0;
        if (launches) *launches = 0;
        return FRIES_OK;
    }
    if (total_ms) *total_ms = it->second.ms;
    if (launches) *launches = it->second.launches;
    return FRIES_OK;
}

// ---------------------------------------------------------------------------------------------------
// a1: hash + owner.  One thread per key, scrambler staged in shared memory (<= 64 x u32).
// HBM traffic: 8 B key read + 8 B hash + 4 B owner written per key.
// ---------------------------------------------------------------------------------------------------
__global__ void hash_owner_kernel(const uint64_t *__restrict__ keys, size_t n, const uint32_t *__restrict__ scr,
                                  int n_bits, unsigned n_ranks, uint64_t *__restrict__ hash_out,
                                  int32_t *__restrict__ owner_out) {
    __shared__ uint32_t s_scr[64];
    if (threadIdx.x < 64) s_scr[threadIdx.x] = threadIdx.x < n_bits ? scr[threadIdx.x] : 0u;
    __syncthreads();
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t h = fr_det_hash(keys[i] & ~FRIES_INI_FLAG, s_scr);
        if (hash_out) hash_out[i] = h;
        if (owner_out) owner_out[i] = (int32_t)(h % n_ranks);
    }
}

extern "C" int fries_hash_owner_dev(fries_ctx *c, const uint64_t *d_keys, size_t n, const uint32_t *h_scr, int n_bits,
                                    int n_ranks, uint64_t *d_hash, int32_t *d_owner) {
    FRIES_REQUIRE(c && (d_keys || n == 0) && h_scr, "fries_hash_owner_dev: NULL argument");
    FRIES_REQUIRE(n_bits > 0 && n_bits <= 63, "fries_hash_owner_dev: n_bits %d not in 1..63", n_bits);
    FRIES_REQUIRE(n_ranks >= 1, "fries_hash_owner_dev: n_ranks must be >= 1");
    if (n == 0) return FRIES_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    FRIES_TRY(c->ensure_scratch(64 * sizeof(uint32_t)));
    uint32_t *d_scr = (uint32_t *)c->d_scratch;
    CUDA_TRY(cudaMemcpyAsync(d_scr, h_scr, n_bits * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    int block = 256;
    size_t want = (n + block - 1) / block;
    int grid = (int)(want < (size_t)c->sm_count * 8 ? want : (size_t)c->sm_count * 8);
    {
        ProfScope ps(c, "hash_owner");
        hash_owner_kernel<<<grid, block, 0, c->stream>>>(d_keys, n, d_scr, n_bits, (unsigned)n_ranks, d_hash, d_owner);
        c->launch_count++;
    }
    CUDA_TRY(cudaGetLastError());
    // d_scr lives in the shared scratch: make sure the kernel is done with it before anyone reuses it
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}

extern "C" int fries_hash_owner(fries_ctx *c, const uint64_t *h_keys, size_t n, const uint32_t *h_scr, int n_bits,
                                int n_ranks, uint64_t *h_hash, int32_t *h_owner) {
    FRIES_REQUIRE(c && (h_keys || n == 0), "fries_hash_owner: NULL argument");
    if (n == 0) return FRIES_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    DevBuf<uint64_t> keys, hash;
    DevBuf<int32_t> owner;
    FRIES_TRY(keys.alloc(n));
    FRIES_TRY(hash.alloc(n));
    FRIES_TRY(owner.alloc(n));
    CUDA_TRY(cudaMemcpyAsync(keys.p, h_keys, n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    FRIES_TRY(fries_hash_owner_dev(c, keys.p, n, h_scr, n_bits, n_ranks, hash.p, owner.p));
    if (h_hash) CUDA_TRY(cudaMemcpyAsync(h_hash, hash.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    if (h_owner) CUDA_TRY(cudaMemcpyAsync(h_owner, owner.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}

// ---------------------------------------------------------------------------------------------------
// a15: batch bit-string utilities
// ---------------------------------------------------------------------------------------------------
__global__ void bit_op_kernel(int op, uint64_t *keys, const uint8_t *orbs, size_t n, int32_t *sign) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = keys[i];
    int s = 0;
    switch (op) {
        case 0: s = fr_sing_det_parity(k, orbs[2 * i], orbs[2 * i + 1]); keys[i] = k; break;
        case 1: s = fr_doub_det_parity(k, orbs[4 * i], orbs[4 * i + 1], orbs[4 * i + 2], orbs[4 * i + 3]); keys[i] = k; break;
        case 2: s = fr_sing_parity(k, orbs[2 * i], orbs[2 * i + 1]); break;
        case 3: s = fr_doub_parity(k, orbs[4 * i], orbs[4 * i + 1], orbs[4 * i + 2], orbs[4 * i + 3]); break;
        case 4: s = fr_bits_between(k, orbs[2 * i], orbs[2 * i + 1]); break;
    }
    sign[i] = s;
}

extern "C" int fries_bit_op(fries_ctx *c, int op, uint64_t *h_keys, const uint8_t *h_orbs, size_t n, int32_t *h_sign) {
    FRIES_REQUIRE(c && h_keys && h_orbs && h_sign, "fries_bit_op: NULL argument");
    FRIES_REQUIRE(op >= 0 && op <= 4, "fries_bit_op: unknown op %d", op);
    if (n == 0) return FRIES_OK;
    CUDA_TRY(cudaSetDevice(c->device));
    int w = (op == 1 || op == 3) ? 4 : 2;
    DevBuf<uint64_t> keys;
    DevBuf<uint8_t> orbs;
    DevBuf<int32_t> sign;
    FRIES_TRY(keys.alloc(n));
    FRIES_TRY(orbs.alloc(n * w));
    FRIES_TRY(sign.alloc(n));
    CUDA_TRY(cudaMemcpyAsync(keys.p, h_keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(orbs.p, h_orbs, n * w, cudaMemcpyHostToDevice, c->stream));
    bit_op_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(op, keys.p, orbs.p, n, sign.p);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_keys, keys.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(h_sign, sign.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}
