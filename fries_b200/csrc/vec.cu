// a2/a3: determinant store -- merge (DistVec::add_elements), compaction/delete, dot, norms.
#include "vec.cuh"
#include "compress.cuh"


// ---------------------------------------------------------------------------------------------------
// merge phase A: find-or-insert.  DistVec::add_elements vec_utils.hpp:606-631 + HashTable::read
// det_hash.hpp:60-94.  Initiator elements (bit 63) may create an entry; others only look up.
// Writes the table slot of every element (FRIES_NO_POS = dropped) for phase B.
// Random HBM/L2 traffic per element: one 8 B probe (+ CAS and 8*(2+n_vecs) B of initialisation on a
// first insertion); streaming: 16 B in, 4 B out.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FR_VEC_BLOCK)
merge_insert_kernel(VecView v, MergeSrc src, uint32_t *__restrict__ slot_out) {
    __shared__ uint32_t s_scr[64];
    load_scr(s_scr, v.scr_vec);
    size_t n = src.count();
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t k = FRIES_EMPTY_KEY;
        double val = 0;
        bool have = src.get(i, k, val);
        uint32_t result = FRIES_NO_POS;
        if (have && k != FRIES_EMPTY_KEY && val != 0) {  // DistVec::add ignores zero values (:418-423)
            bool ini = (k >> 63) != 0;
            uint64_t key = k & ~FRIES_INI_FLAG;
            uint64_t slot = vec_hash(v, key, s_scr) & v.tmask;
            while (true) {
                uint64_t cur = *((volatile uint64_t *)&v.tkeys[slot]);
                if (cur == key) {
                    result = (uint32_t)slot;
                    break;
                }
                if (cur == FRIES_EMPTY_KEY) {
                    if (!ini) break;
                    unsigned long long old = atomicCAS((unsigned long long *)&v.tkeys[slot], FRIES_EMPTY_KEY, key);
                    if (old == FRIES_EMPTY_KEY) {
                        // this thread owns the new entry: allocate a storage position (append)
                        unsigned long long pos = atomicAdd(&v.cnt->n, 1ull);
                        if (pos < v.cap) {
                            v.keys[pos] = key;
                            for (unsigned r = 0; r < v.n_vecs; r++) v.vals[(size_t)r * v.cap + pos] = 0.0;
                            v.diag[pos] = __longlong_as_double(0x7ff8000000000000ll);
                            v.tpos[slot] = (uint32_t)pos;
                        } else {
                            v.tpos[slot] = FRIES_OVF_POS;
                            atomicAdd(&v.cnt->overflow, 1ull);
                        }
                        result = (uint32_t)slot;
                        break;
                    }
                    if (old == key) {
                        result = (uint32_t)slot;
                        break;
                    }
                }
                slot = (slot + 1) & v.tmask;
            }
        }
        slot_out[i] = result;
    }
}

// merge phase B: accumulate with the initiator rule (vec_utils.hpp:632-637)
__global__ void __launch_bounds__(FR_VEC_BLOCK)
merge_accum_kernel(VecView v, MergeSrc src, const uint32_t *__restrict__ slot_in, unsigned origin, unsigned dest) {
    size_t n = src.count();
    size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long nonini = 0, valid = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t slot = slot_in[i];
        if (slot == FRIES_NO_POS) continue;
        uint32_t pos = v.tpos[slot];
        if (pos >= FRIES_OVF_POS) continue;
        valid++;
        uint64_t k;
        double val;
        src.get(i, k, val);
        bool ini = (k >> 63) != 0;
        bool nonz = v.vals[(size_t)origin * v.cap + pos] != 0;
        if (ini || nonz) atomicAdd(&v.vals[(size_t)dest * v.cap + pos], val);
        if (!ini && nonz) nonini++;
    }
    nonini = warp_sum_u64(nonini);
    valid = warp_sum_u64(valid);
    if ((threadIdx.x & 31) == 0) {
        if (nonini) atomicAdd(&v.cnt->nonini_occ_add, nonini);
        if (valid) atomicAdd(&v.cnt->n_spawn_valid, valid);
    }
}

// ---------------------------------------------------------------------------------------------------
// merge, both phases in one pass (origin != dest: the row the initiator rule reads is not the row being added to, which is
// every merge of an FRI iteration).  An element that finds its determinant in the index adds right away; if the entry was
// made by another element of this batch a moment ago, its position may still be on its way: the inserting thread
// initialises the storage, fences, then publishes the position, and the finder spins on the entry until it is there (the
// inserter never waits for anybody, so the spin ends; a full store publishes FRIES_OVF_POS instead).
// Saves the second pass over the batch and its 4 B per element of slot scratch (0.037 of 0.065 ms at 1e6 spawned elements).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FR_VEC_BLOCK)
merge_fused_kernel(VecView v, MergeSrc src, unsigned origin, unsigned dest) {
    __shared__ uint32_t s_scr[64];
    load_scr(s_scr, v.scr_vec);
    const size_t n = src.count();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long nonini = 0, valid = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t k = FRIES_EMPTY_KEY;
        double val = 0;
        const bool have = src.get(i, k, val);
        if (!(have && k != FRIES_EMPTY_KEY && val != 0)) continue;  // DistVec::add ignores zero values (:418-423)
        const bool ini = (k >> 63) != 0;
        const uint64_t key = k & ~FRIES_INI_FLAG;
        uint64_t slot = vec_hash(v, key, s_scr) & v.tmask;
        uint32_t pos = FRIES_NO_POS;
        bool found = false;
        while (true) {
            const uint64_t cur = *((volatile uint64_t *)&v.tkeys[slot]);
            if (cur == key) {
                found = true;
                break;
            }
            if (cur == FRIES_EMPTY_KEY) {
                if (!ini) break;
                const unsigned long long old = atomicCAS((unsigned long long *)&v.tkeys[slot], FRIES_EMPTY_KEY, key);
                if (old == FRIES_EMPTY_KEY) {
                    // this thread owns the new entry: allocate a storage position (append), initialise, publish
                    const unsigned long long p = atomicAdd(&v.cnt->n, 1ull);
                    if (p < v.cap) {
                        v.keys[p] = key;
                        for (unsigned r = 0; r < v.n_vecs; r++) v.vals[(size_t)r * v.cap + p] = 0.0;
                        v.diag[p] = __longlong_as_double(0x7ff8000000000000ll);
                        __threadfence();
                        *((volatile uint32_t *)&v.tpos[slot]) = (uint32_t)p;
                        pos = (uint32_t)p;
                    } else {
                        *((volatile uint32_t *)&v.tpos[slot]) = FRIES_OVF_POS;
                        atomicAdd(&v.cnt->overflow, 1ull);
                        pos = FRIES_OVF_POS;
                    }
                    break;
                }
                if (old == key) {
                    found = true;
                    break;
                }
            }
            slot = (slot + 1) & v.tmask;
        }
        if (found) {
            do pos = *((volatile uint32_t *)&v.tpos[slot]);
            while (pos == FRIES_NO_POS);
        }
        if (pos >= FRIES_OVF_POS) continue;
        valid++;
        const bool nonz = v.vals[(size_t)origin * v.cap + pos] != 0;
        if (ini || nonz) atomicAdd(&v.vals[(size_t)dest * v.cap + pos], val);
        if (!ini && nonz) nonini++;
    }
    nonini = warp_sum_u64(nonini);
    valid = warp_sum_u64(valid);
    if ((threadIdx.x & 31) == 0) {
        if (nonini) atomicAdd(&v.cnt->nonini_occ_add, nonini);
        if (valid) atomicAdd(&v.cnt->n_spawn_valid, valid);
    }
}

// clamp curr_size after an overflowing merge
__global__ void clamp_count_kernel(VecCounters *cnt, unsigned long long cap) {
    if (cnt->n > cap) cnt->n = cap;
}

// ---------------------------------------------------------------------------------------------------
// delete + compaction + index rebuild (DistVec::del_at_pos vec_utils.hpp:458-476).
// One cooperative kernel: clear the table; stable compaction of the survivors into the spare buffers
// (chunked scan); rebuild the table from the compacted keys.
// Traffic: 8*(2+n_vecs) B read + written per stored element, 12 B * T table clear, one random 12 B
// store per survivor.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FR_COMP_BLOCK)
compact_kernel(VecView v, uint64_t *__restrict__ keys_b, double *__restrict__ vals_b, double *__restrict__ diag_b,
               const uint8_t *__restrict__ del_flags, size_t min_del_idx, double *part_d, unsigned long long *part_c) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh_d[34];
    __shared__ unsigned long long sh_c[34];
    __shared__ double sh_sd[68];
    __shared__ unsigned long long sh_sc[68];
    __shared__ uint32_t s_scr[64];
    load_scr(s_scr, v.scr_vec);
    GridRed red{part_d, part_c, 0, (int)gridDim.x, sh_d, sh_c};
    unsigned long long n64 = *((volatile unsigned long long *)&v.cnt->n);
    const size_t n = n64 < v.cap ? (size_t)n64 : v.cap;
    // phase 0: clear the index
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s <= v.tmask; s += gstride) {
        v.tkeys[s] = FRIES_EMPTY_KEY;
        v.tpos[s] = FRIES_NO_POS;
    }
    size_t chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + 31) & ~(size_t)31;
    const size_t lo = (size_t)blockIdx.x * chunk < n ? (size_t)blockIdx.x * chunk : n;
    const size_t hi = lo + chunk < n ? lo + chunk : n;
    auto survives = [&](size_t i) -> bool {
        if (i < min_del_idx) return true;
        if (del_flags && !del_flags[i]) return true;
        for (unsigned r = 0; r < v.n_vecs; r++)
            if (v.vals[(size_t)r * v.cap + i] != 0) return true;
        return false;
    };
    unsigned long long c = 0;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) c += survives(i) ? 1 : 0;
    c = block_sum_u64(c, sh_c);
    double d0, d1;
    unsigned long long blk_off, total;
    grid_excl_scan(grid, red, 0.0, c, d0, blk_off, d1, total, sh_sd, sh_sc);
    unsigned long long carry = blk_off;
    for (size_t base = lo; base < hi; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        bool sv = i < hi && survives(i);
        double ex, tot;
        unsigned long long ec, tc;
        block_excl_scan(0.0, sv ? 1ull : 0ull, ex, ec, tot, tc, sh_sd, sh_sc);
        if (sv) {
            size_t o = (size_t)(carry + ec);
            keys_b[o] = v.keys[i];
            diag_b[o] = v.diag[i];
            for (unsigned r = 0; r < v.n_vecs; r++) vals_b[(size_t)r * v.cap + o] = v.vals[(size_t)r * v.cap + i];
        }
        carry += tc;
    }
    grid.sync();
    // phase 3: rebuild the index from the compacted keys (all distinct)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gstride) {
        uint64_t key = __ldcg(&keys_b[i]);
        uint64_t slot = vec_hash(v, key, s_scr) & v.tmask;
        while (true) {
            unsigned long long old = atomicCAS((unsigned long long *)&v.tkeys[slot], FRIES_EMPTY_KEY, key);
            if (old == FRIES_EMPTY_KEY) {
                v.tpos[slot] = (uint32_t)i;
                break;
            }
            slot = (slot + 1) & v.tmask;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) v.cnt->n = total;
}

// ---------------------------------------------------------------------------------------------------
// DistVec::dot with a replicated trial vector (vec_utils.hpp:228-253): single CTA, fixed-order sum
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
trial_dot_kernel(VecView v, const uint64_t *__restrict__ t_keys, const double *__restrict__ t_vals, size_t n_trial,
                 unsigned row, double *out) {
    __shared__ uint32_t s_scr[64];
    __shared__ double sh[34];
    load_scr(s_scr, v.scr_vec);
    double acc = 0;
    for (size_t i = threadIdx.x; i < n_trial; i += blockDim.x) {
        uint32_t pos = vec_lookup(v, t_keys[i], s_scr);
        if (pos != FRIES_NO_POS) acc += t_vals[i] * v.vals[(size_t)row * v.cap + pos];
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) *out = acc;
}

// DistVec::local_norm vec_utils.hpp:683-689
__global__ void __launch_bounds__(FR_COMP_BLOCK)
local_norm_kernel(VecView v, unsigned row, int squares, double *part_d, unsigned long long *part_c, double *out) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh_d[34];
    __shared__ unsigned long long sh_c[34];
    GridRed red{part_d, part_c, 0, (int)gridDim.x, sh_d, sh_c};
    unsigned long long n64 = v.cnt->n;
    size_t n = n64 < v.cap ? (size_t)n64 : v.cap;
    double s = 0;
    size_t gstride = (size_t)gridDim.x * blockDim.x;
    const double *x = v.vals + (size_t)row * v.cap;
    // four independent loads in flight per thread
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4 * gstride) {
        double a[4];
#pragma unroll
        for (int k = 0; k < 4; k++) a[k] = i + k * gstride < n ? x[i + k * gstride] : 0.0;
#pragma unroll
        for (int k = 0; k < 4; k++) s += squares ? a[k] * a[k] : fabs(a[k]);
    }
    unsigned long long dummy = 0;
    grid_reduce(grid, red, s, dummy);
    if (blockIdx.x == 0 && threadIdx.x == 0) *out = s;  // two_norm is the SUM of squares (vec_utils.hpp:695-701: no root)
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int fries_vec::read_counters(VecCounters *out) {
    CUDA_TRY(cudaMemcpyAsync(out, cnt.p, sizeof(VecCounters), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return FRIES_OK;
}

extern "C" int fries_vec_create(fries_ctx *c, size_t capacity, unsigned n_bits, unsigned n_elec, unsigned n_vecs,
                                const uint32_t *h_proc_scr, const uint32_t *h_vec_scr, int n_ranks, int rank,
                                fries_vec **out) {
    FRIES_REQUIRE(c && out && h_proc_scr && h_vec_scr, "fries_vec_create: NULL argument");
    FRIES_REQUIRE(n_bits >= 1 && n_bits <= 63, "fries_vec_create: n_bits %u not in 1..63 (one-word keys)", n_bits);
    FRIES_REQUIRE(n_elec >= 1 && n_elec <= FRIES_MAX_ELEC, "fries_vec_create: n_elec %u not in 1..%d", n_elec,
                  FRIES_MAX_ELEC);
    FRIES_REQUIRE(n_vecs >= 1 && n_vecs <= 8, "fries_vec_create: n_vecs %u not in 1..8", n_vecs);
    FRIES_REQUIRE(capacity >= 1 && capacity < 0x7fffffffull, "fries_vec_create: capacity %zu out of range", capacity);
    FRIES_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "fries_vec_create: bad rank %d of %d", rank, n_ranks);
    CUDA_TRY(cudaSetDevice(c->device));
    fries_vec *v = new fries_vec();
    v->ctx = c;
    {
        const char *e = getenv("FRIES_DETERMINISTIC");
        v->deterministic = e && e[0] == '1' && e[1] == 0;
    }
    v->cap = capacity;
    v->n_bits = n_bits;
    v->n_elec = n_elec;
    v->n_vecs = n_vecs;
    v->n_ranks = n_ranks;
    v->rank = rank;
    size_t t = 1024;
    while (t < 2 * capacity) t <<= 1;
    v->tsize = t;
    int rc = FRIES_OK;
    for (int b = 0; b < 2 && rc == FRIES_OK; b++) {
        rc = v->keys[b].alloc(capacity);
        if (rc == FRIES_OK) rc = v->vals[b].alloc(capacity * n_vecs);
        if (rc == FRIES_OK) rc = v->diag[b].alloc(capacity);
    }
    if (rc == FRIES_OK) rc = v->tkeys.alloc(t);
    if (rc == FRIES_OK) rc = v->tpos.alloc(t);
    if (rc == FRIES_OK) rc = v->scr.alloc(128);
    if (rc == FRIES_OK) rc = v->cnt.alloc(1);
    if (rc == FRIES_OK) rc = v->red_d.alloc(FR_RED_PART_LEN);
    if (rc == FRIES_OK) rc = v->red_c.alloc(FR_RED_PART_LEN);
    if (rc != FRIES_OK) {
        delete v;
        return rc;
    }
    uint32_t h_scr[128] = {0};
    memcpy(h_scr, h_vec_scr, n_bits * 4);
    memcpy(h_scr + 64, h_proc_scr, n_bits * 4);
    v->h_scr_vec.assign(h_vec_scr, h_vec_scr + n_bits);
    v->h_scr_proc.assign(h_proc_scr, h_proc_scr + n_bits);
    CUDA_TRY(cudaMemcpyAsync(v->scr.p, h_scr, sizeof(h_scr), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemsetAsync(v->tkeys.p, 0xff, t * 8, c->stream));
    CUDA_TRY(cudaMemsetAsync(v->tpos.p, 0xff, t * 4, c->stream));
    CUDA_TRY(cudaMemsetAsync(v->cnt.p, 0, sizeof(VecCounters), c->stream));
    for (int b = 0; b < 2; b++) CUDA_TRY(cudaMemsetAsync(v->vals[b].p, 0, capacity * n_vecs * 8, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out = v;
    return FRIES_OK;
}

// HubHolVec ctor hh_vec.hpp:27-29: n_bits = 2 n_sites + n_sites * ph_bits; hashes include the phonon numbers
extern "C" int fries_vec_create_hh(fries_ctx *c, size_t capacity, unsigned n_sites, unsigned ph_bits, unsigned n_elec,
                                   unsigned n_vecs, const uint32_t *h_proc_scr, const uint32_t *h_vec_scr, int n_ranks,
                                   int rank, fries_vec **out) {
    FRIES_REQUIRE(n_sites >= 2 && ph_bits >= 1 && n_sites * (2 + ph_bits) <= 63 && (1u << ph_bits) <= 2 * n_sites,
                  "fries_vec_create_hh: %u sites x (2 + %u) bits do not fit one 63-bit key (or phonon numbers exceed the scrambler)",
                  n_sites, ph_bits);
    // the scramblers have 2 n_sites entries (frisys_hh.cpp:72,100); pad to the key width for the generic ctor
    std::vector<uint32_t> ps(n_sites * (2 + ph_bits), 0), vs(n_sites * (2 + ph_bits), 0);
    memcpy(ps.data(), h_proc_scr, 2 * n_sites * 4);
    memcpy(vs.data(), h_vec_scr, 2 * n_sites * 4);
    FRIES_TRY(fries_vec_create(c, capacity, n_sites * (2 + ph_bits), n_elec, n_vecs, ps.data(), vs.data(), n_ranks, rank, out));
    (*out)->hh_sites = n_sites;
    (*out)->hh_ph_bits = ph_bits;
    return FRIES_OK;
}
// Semi-stochastic calculations (DistVec::init_dense vec_utils.hpp:858-897): the first n_dense stored determinants are
// the deterministic subspace -- never deleted, never compressed; their columns of H are applied exactly
extern "C" int fries_vec_set_dense(fries_vec *vec, size_t n_dense) {
    FRIES_REQUIRE(vec, "fries_vec_set_dense: NULL vector");
    VecCounters cnt;
    FRIES_TRY(vec->read_counters(&cnt));
    FRIES_REQUIRE(n_dense <= cnt.n, "fries_vec_set_dense: %zu exceeds the number of stored determinants", n_dense);
    vec->n_dense = n_dense;
    vec->min_del_idx = n_dense;
    if (vec->n_ranks == 1) vec->n_dense_total = n_dense;
    return FRIES_OK;
}
// several ranks: the size of the dense subspace over all ranks (the dense multiplication is collective: a rank without
// dense determinants still receives its share of the others' connections)
extern "C" int fries_vec_set_dense_total(fries_vec *vec, size_t n_dense_total) {
    FRIES_REQUIRE(vec, "fries_vec_set_dense_total: NULL vector");
    FRIES_REQUIRE(n_dense_total >= vec->n_dense, "fries_vec_set_dense_total: %zu is below this rank's %zu", n_dense_total,
                  vec->n_dense);
    vec->n_dense_total = n_dense_total;
    return FRIES_OK;
}

// Reproducible merges: new determinants are appended and values added in the order of the batch, as the reference's
// sequential add_elements does (vec_utils.hpp:606-641), instead of in the order the SMs get to them.  One rank only.
extern "C" int fries_vec_set_deterministic(fries_vec *vec, int on) {
    FRIES_REQUIRE(vec, "fries_vec_set_deterministic: NULL vector");
    vec->deterministic = on != 0;
    return FRIES_OK;
}

// DistVec::set_min_del_idx vec_utils.hpp: positions below idx are never deleted
extern "C" int fries_vec_set_min_del_idx(fries_vec *vec, size_t idx) {
    FRIES_REQUIRE(vec, "NULL argument");
    vec->min_del_idx = idx;
    return FRIES_OK;
}

extern "C" int fries_hbpp_destroy(struct fries_hbpp *hb);
extern "C" int fries_vec_destroy(fries_vec *v) {
    if (v) {
        cudaSetDevice(v->ctx->device);
        if (v->hv_scratch) fries_hbpp_destroy(v->hv_scratch);
        delete v;
    }
    return FRIES_OK;
}

int fries_vec_merge_dev(fries_vec *vec, const uint64_t *d_keys, const double *d_vals, size_t n_max,
                        const unsigned long long *d_n, unsigned origin, unsigned dest) {
    MergeSrc src{d_keys, d_vals, n_max, d_n, nullptr, 0};
    return fries_vec_merge_src_dev(vec, src, origin, dest);
}

int fries_vec_merge_src_dev(fries_vec *vec, const MergeSrc &src, unsigned origin, unsigned dest) {
    fries_ctx *c = vec->ctx;
    FRIES_REQUIRE(origin < vec->n_vecs && dest < vec->n_vecs, "merge: row index out of range");
    const size_t n_max = src.n_max;
    if (n_max == 0) return FRIES_OK;
    if (vec->deterministic && vec->n_ranks == 1) {  // reproducible variant: batch-order append and additions (vec_det.cu)
        FRIES_TRY(fries_vec_merge_det_dev(vec, src, origin, dest));
        CUDA_TRY(cudaGetLastError());
        return FRIES_OK;
    }
    VecView v = vec->view();
    size_t want = (n_max + FR_VEC_BLOCK - 1) / FR_VEC_BLOCK;
    int grid = (int)(want < (size_t)c->sm_count * 8 ? want : (size_t)c->sm_count * 8);
    static const bool two_pass = [] {  // FRIES_MERGE_TWO_PASS=1: the separate insert / accumulate kernels (A/B, regression)
        const char *e = getenv("FRIES_MERGE_TWO_PASS");
        return e && e[0] == '1' && e[1] == 0;
    }();
    // One pass where it pays (same-box A/B, round 2): an index beyond the L2 (1.25e7-determinant store: 2.25 -> 1.54 ms) and
    // few insertions per element; with the index in L2 the second pass is cheap (H2O-sized: 0.063 against 0.085 ms in one
    // pass), and the full H.v inserts a determinant for every few elements, each insertion with its fence (5.1 -> 6.3 ms).
    const bool index_in_l2 = (size_t)vec->tsize * 12 <= ((size_t)96 << 20);
    if (origin != dest && !two_pass && !index_in_l2 && !vec->merge_many_new) {
        ProfScope ps(c, "merge_insert");
        merge_fused_kernel<<<grid, FR_VEC_BLOCK, 0, c->stream>>>(v, src, origin, dest);
        c->launch_count++;
        clamp_count_kernel<<<1, 1, 0, c->stream>>>(vec->cnt.p, (unsigned long long)vec->cap);
        c->launch_count++;
        CUDA_TRY(cudaGetLastError());
        return FRIES_OK;
    }
    FRIES_TRY(vec->slot_scratch.ensure(n_max));
    {
        ProfScope ps(c, "merge_insert");
        merge_insert_kernel<<<grid, FR_VEC_BLOCK, 0, c->stream>>>(v, src, vec->slot_scratch.p);
        c->launch_count++;
    }
    {
        ProfScope ps(c, "merge_accum");
        merge_accum_kernel<<<grid, FR_VEC_BLOCK, 0, c->stream>>>(v, src, vec->slot_scratch.p, origin, dest);
        c->launch_count++;
    }
    clamp_count_kernel<<<1, 1, 0, c->stream>>>(vec->cnt.p, (unsigned long long)vec->cap);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    return FRIES_OK;
}

extern "C" int fries_vec_add_dev(fries_vec *vec, const uint64_t *d_keys, const double *d_vals, size_t n,
                                 const uint32_t *d_n, unsigned origin, unsigned dest) {
    FRIES_REQUIRE(vec && (n == 0 || (d_keys && d_vals)), "fries_vec_add_dev: NULL argument");
    FRIES_REQUIRE(d_n == nullptr, "fries_vec_add_dev: device-side count must be passed through the pipeline API");
    CUDA_TRY(cudaSetDevice(vec->ctx->device));
    return fries_vec_merge_dev(vec, d_keys, d_vals, n, nullptr, origin, dest);
}

extern "C" int fries_vec_add(fries_vec *vec, const uint64_t *h_keys, const double *h_vals, const uint8_t *h_ini,
                             size_t n, unsigned origin, unsigned dest) {
    FRIES_REQUIRE(vec && (n == 0 || (h_keys && h_vals && h_ini)), "fries_vec_add: NULL argument");
    if (n == 0) return FRIES_OK;
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    uint64_t all = vec->n_bits >= 63 ? ~0ull >> 1 : (1ull << vec->n_bits) - 1;
    std::vector<uint64_t> k(n);
    for (size_t i = 0; i < n; i++) {
        if (h_keys[i] & ~all) {
            fries_set_error("fries_vec_add: key %zu has bits above n_bits=%u", i, vec->n_bits);
            return FRIES_ERR_ARG;
        }
        uint64_t ek = vec->hh_sites ? (h_keys[i] & ((1ull << (2 * vec->hh_sites)) - 1)) : h_keys[i];
        if ((unsigned)__builtin_popcountll(ek) != vec->n_elec) {
            // DistVec::idx_to_hash throws for a wrong electron count (vec_utils.hpp:389-399)
            fries_set_error("Determinant %016llx created with an incorrect number of electrons",
                            (unsigned long long)h_keys[i]);
            return FRIES_ERR_ARG;
        }
        k[i] = h_keys[i] | (h_ini[i] ? FRIES_INI_FLAG : 0ull);
    }
    DevBuf<uint64_t> dk;
    DevBuf<double> dv;
    FRIES_TRY(dk.alloc(n));
    FRIES_TRY(dv.alloc(n));
    CUDA_TRY(cudaMemcpyAsync(dk.p, k.data(), n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(dv.p, h_vals, n * 8, cudaMemcpyHostToDevice, c->stream));
    FRIES_TRY(fries_vec_merge_dev(vec, dk.p, dv.p, n, nullptr, origin, dest));
    VecCounters cnt;
    FRIES_TRY(vec->read_counters(&cnt));
    if (cnt.overflow) {
        fries_set_error("fries_vec_add: store is full (capacity %zu); %llu insertions dropped", vec->cap, cnt.overflow);
        return FRIES_ERR_CAPACITY;
    }
    return FRIES_OK;
}

// DistVec::load vec_utils.hpp:761-844 from host arrays: storage <- (keys, value rows), diag cache reset, index rebuilt
__global__ void upload_fixup_kernel(VecView v, size_t n, unsigned n_elec, unsigned n_bits) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t k = v.keys[i];
        uint64_t ek = v.hh_sites ? (k & ((1ull << (2 * v.hh_sites)) - 1)) : k;
        if ((unsigned)__popcll(ek) != n_elec || (k >> n_bits) != 0) bad++;
        v.diag[i] = __longlong_as_double(0x7ff8000000000000ll);
    }
    bad = warp_sum_u64(bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(&v.cnt->bad_keys, bad);
    if (blockIdx.x == 0 && threadIdx.x == 0) v.cnt->n = n;
}

int fries_vec_compact_flags_dev(fries_vec *vec, const uint8_t *d_flags);

extern "C" int fries_vec_upload(fries_vec *vec, const uint64_t *h_keys, const double *h_vals, size_t n) {
    FRIES_REQUIRE(vec && (n == 0 || (h_keys && h_vals)), "fries_vec_upload: NULL argument");
    FRIES_REQUIRE(n <= vec->cap, "fries_vec_upload: %zu elements exceed the capacity %zu", n, vec->cap);
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    VecView v = vec->view();
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(v.keys, h_keys, n * 8, cudaMemcpyHostToDevice, c->stream));
        for (unsigned r = 0; r < vec->n_vecs; r++)
            CUDA_TRY(cudaMemcpyAsync(v.vals + (size_t)r * v.cap, h_vals + (size_t)r * n, n * 8, cudaMemcpyHostToDevice,
                                     c->stream));
    }
    CUDA_TRY(cudaMemsetAsync(vec->cnt.p, 0, sizeof(VecCounters), c->stream));
    upload_fixup_kernel<<<c->sm_count * 4, 256, 0, c->stream>>>(v, n, vec->n_elec, vec->n_bits);
    c->launch_count++;
    FRIES_TRY(fries_vec_compact_flags_dev(vec, nullptr));  // drops all-zero elements, rebuilds the index
    VecCounters cnt;
    FRIES_TRY(vec->read_counters(&cnt));
    if (cnt.bad_keys) {
        fries_set_error("fries_vec_upload: %llu determinants with an incorrect number of electrons", cnt.bad_keys);
        return FRIES_ERR_ARG;
    }
    return FRIES_OK;
}

extern "C" int fries_vec_curr_size(fries_vec *vec, size_t *curr_size) {
    FRIES_REQUIRE(vec && curr_size, "NULL argument");
    VecCounters cnt;
    FRIES_TRY(vec->read_counters(&cnt));
    *curr_size = (size_t)cnt.n;
    return FRIES_OK;
}
extern "C" int fries_vec_n_nonz(fries_vec *vec, size_t *n_nonz) { return fries_vec_curr_size(vec, n_nonz); }
extern "C" int fries_vec_nonini_occ_add(fries_vec *vec, uint64_t *count) {
    FRIES_REQUIRE(vec && count, "NULL argument");
    VecCounters cnt;
    FRIES_TRY(vec->read_counters(&cnt));
    *count = cnt.nonini_occ_add;
    return FRIES_OK;
}

extern "C" int fries_vec_download(fries_vec *vec, uint64_t *h_keys, double *h_vals, size_t cap, size_t *n_out) {
    FRIES_REQUIRE(vec && n_out, "NULL argument");
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    VecCounters cnt;
    FRIES_TRY(vec->read_counters(&cnt));
    size_t n = (size_t)cnt.n;
    *n_out = n;
    if (n > cap) {
        fries_set_error("fries_vec_download: %zu elements stored, buffer holds %zu", n, cap);
        return FRIES_ERR_CAPACITY;
    }
    if (n == 0) return FRIES_OK;
    if (h_keys) CUDA_TRY(cudaMemcpyAsync(h_keys, vec->keys[vec->cur].p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (h_vals)
        for (unsigned r = 0; r < vec->n_vecs; r++)
            CUDA_TRY(cudaMemcpyAsync(h_vals + (size_t)r * n, vec->vals[vec->cur].p + (size_t)r * vec->cap, n * 8,
                                     cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}

int fries_vec_compact_flags_dev(fries_vec *vec, const uint8_t *d_flags) {
    fries_ctx *c = vec->ctx;
    int grid = c->coop_grid((const void *)compact_kernel, FR_COMP_BLOCK, 0);
    FRIES_REQUIRE(grid * 2 <= 4 * 1024, "compact: grid too large");
    VecView v = vec->view();
    int nb = vec->cur ^ 1;
    uint64_t *kb = vec->keys[nb].p;
    double *vb = vec->vals[nb].p, *db = vec->diag[nb].p;
    size_t mdi = vec->min_del_idx;
    double *pd = vec->red_d.p;
    unsigned long long *pc = vec->red_c.p;
    void *args[] = {(void *)&v, (void *)&kb, (void *)&vb, (void *)&db, (void *)&d_flags, (void *)&mdi, (void *)&pd,
                    (void *)&pc};
    {
        ProfScope ps(c, "compact");
        CUDA_TRY(cudaLaunchCooperativeKernel((const void *)compact_kernel, dim3(grid), dim3(FR_COMP_BLOCK), args, 0,
                                             c->stream));
        c->launch_count++;
    }
    vec->cur = nb;
    return FRIES_OK;
}
int fries_vec_compact_dev(fries_vec *vec) { return fries_vec_compact_flags_dev(vec, nullptr); }

extern "C" int fries_vec_del(fries_vec *vec, const uint8_t *h_flags, size_t n) {
    FRIES_REQUIRE(vec && (n == 0 || h_flags), "fries_vec_del: NULL argument");
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    DevBuf<uint8_t> fl;
    FRIES_TRY(fl.alloc(vec->cap));
    CUDA_TRY(cudaMemsetAsync(fl.p, 0, vec->cap, c->stream));
    if (n) CUDA_TRY(cudaMemcpyAsync(fl.p, h_flags, n < vec->cap ? n : vec->cap, cudaMemcpyHostToDevice, c->stream));
    FRIES_TRY(fries_vec_compact_flags_dev(vec, fl.p));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}

extern "C" int fries_vec_dot(fries_vec *vec, const uint64_t *h_keys, const double *h_vals, size_t n, unsigned row,
                             double *out) {
    FRIES_REQUIRE(vec && out && (n == 0 || (h_keys && h_vals)), "fries_vec_dot: NULL argument");
    FRIES_REQUIRE(row < vec->n_vecs, "fries_vec_dot: row out of range");
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    DevBuf<uint64_t> dk;
    DevBuf<double> dv, dout;
    FRIES_TRY(dk.alloc(n));
    FRIES_TRY(dv.alloc(n));
    FRIES_TRY(dout.alloc(1));
    if (n) {
        CUDA_TRY(cudaMemcpyAsync(dk.p, h_keys, n * 8, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(dv.p, h_vals, n * 8, cudaMemcpyHostToDevice, c->stream));
    }
    trial_dot_kernel<<<1, 1024, 0, c->stream>>>(vec->view(), dk.p, dv.p, n, row, dout.p);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out, dout.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}

// DistVec::add_vecs / copy_vec / weight_vec / zero_vec (vec_utils.hpp:547-579) on the stored elements of two rows
__global__ void __launch_bounds__(256)
row_op_kernel(VecView v, int op, unsigned dst, unsigned src, double c) {
    unsigned long long n64 = v.cnt->n;
    const size_t n = n64 < v.cap ? (size_t)n64 : v.cap;
    double *d = v.vals + (size_t)dst * v.cap;
    const double *s = v.vals + (size_t)src * v.cap;
    const size_t gstride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4 * gstride) {
        double a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            size_t j = i + k * gstride;
            a[k] = (j < n && op != 1 && op != 3) ? d[j] : 0.0;
            b[k] = (j < n && op != 3) ? s[j] : 0.0;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            size_t j = i + k * gstride;
            if (j >= n) continue;
            double r;
            switch (op) {
                case 0: r = a[k] + b[k] * c; break;                   // add_vecs(dst, src, c)
                case 1: r = b[k]; break;                              // copy_vec(src, dst)
                case 2: r = a[k] * pow(1 + fabs(b[k]), c); break;     // weight_vec(dst, src, expo)
                default: r = 0.0; break;                              // zero_vec
            }
            d[j] = r;
        }
    }
}

extern "C" int fries_vec_row_op(fries_vec *vec, int op, unsigned dst, unsigned src, double c) {
    FRIES_REQUIRE(vec, "fries_vec_row_op: NULL vector");
    FRIES_REQUIRE(op >= 0 && op <= 3, "fries_vec_row_op: op %d not in 0..3", op);
    FRIES_REQUIRE(dst < vec->n_vecs && src < vec->n_vecs, "fries_vec_row_op: row out of range");
    fries_ctx *ctx = vec->ctx;
    CUDA_TRY(cudaSetDevice(ctx->device));
    row_op_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(vec->view(), op, dst, src, c);
    ctx->launch_count++;
    CUDA_TRY(cudaGetLastError());
    return FRIES_OK;
}

static int vec_norm(fries_vec *vec, unsigned row, int squares, double *out);
extern "C" int fries_vec_two_norm(fries_vec *vec, unsigned row, double *out) {
    FRIES_REQUIRE(vec && out, "NULL argument");
    FRIES_REQUIRE(row < vec->n_vecs, "fries_vec_two_norm: row out of range");
    return vec_norm(vec, row, 1, out);
}
extern "C" int fries_vec_local_norm(fries_vec *vec, unsigned row, double *out) {
    FRIES_REQUIRE(vec && out, "NULL argument");
    FRIES_REQUIRE(row < vec->n_vecs, "fries_vec_local_norm: row out of range");
    return vec_norm(vec, row, 0, out);
}
static int vec_norm(fries_vec *vec, unsigned row, int squares, double *out) {
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    int grid = c->coop_grid((const void *)local_norm_kernel, FR_COMP_BLOCK, 0);
    DevBuf<double> dout;
    FRIES_TRY(dout.alloc(1));
    VecView v = vec->view();
    double *pd = vec->red_d.p;
    unsigned long long *pc = vec->red_c.p;
    double *po = dout.p;
    void *args[] = {(void *)&v, (void *)&row, (void *)&squares, (void *)&pd, (void *)&pc, (void *)&po};
    CUDA_TRY(cudaLaunchCooperativeKernel((const void *)local_norm_kernel, dim3(grid), dim3(FR_COMP_BLOCK), args, 0,
                                         c->stream));
    c->launch_count++;
    CUDA_TRY(cudaMemcpyAsync(out, dout.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}
