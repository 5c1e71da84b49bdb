// Reproducible merge (optional: fries_vec_set_deterministic / FRIES_DETERMINISTIC=1).
//
// The default merge (vec.cu) gives a new determinant the storage position atomicAdd hands its inserting thread, and adds the
// values with atomicAdd: positions and the last bits of the sums depend on the order in which the SMs got there.  Storage
// order feeds sys_comp, so two runs with one seed drift apart -- the reference (DistVec::add_elements
// FRIES/vec_utils.hpp:606-641, one rank) is sequential and reproducible.  This variant restores that:
//   1. find-or-insert as before, but nobody allocates: every element stamps its slot with atomicMin(0x80000000 | batch
//      index), which only lands on slots without a position -- the stamp that survives is the FIRST batch element of each
//      new determinant;
//   2. those first elements are ranked by an exclusive scan over the batch and appended in batch order, which is the order
//      the reference's sequential loop appends them in;
//   3. the batch is sorted by slot (stable radix sort, so batch order survives inside a slot) and one thread per slot adds
//      its elements one after the other with the initiator rule evaluated on the running value -- the reference's order of
//      additions, not just a fixed one.
// The scan and the sort are CUB's (library plumbing of an opt-in mode; the default path has none).  Single rank: with
// several ranks the arrival order inside the receive windows is itself unordered.
#include <cub/cub.cuh>

#include "vec.cuh"

#define FR_DET_PENDING 0x80000000u

__global__ void __launch_bounds__(FR_VEC_BLOCK)
merge_det_insert_kernel(VecView v, MergeSrc src, uint32_t *__restrict__ slot_out, uint32_t *__restrict__ iota) {
    __shared__ uint32_t s_scr[64];
    load_scr(s_scr, v.scr_vec);
    const size_t n = src.count(), n_max = src.n_max;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_max; i += stride) {
        uint64_t k = FRIES_EMPTY_KEY;
        double val = 0;
        const bool have = i < n && src.get(i, k, val);
        uint32_t result = FRIES_NO_POS;
        if (have && k != FRIES_EMPTY_KEY && val != 0) {
            const bool ini = (k >> 63) != 0;
            const uint64_t key = k & ~FRIES_INI_FLAG;
            uint64_t slot = vec_hash(v, key, s_scr) & v.tmask;
            while (true) {
                const uint64_t cur = *((volatile uint64_t *)&v.tkeys[slot]);
                if (cur == key) {
                    result = (uint32_t)slot;
                    break;
                }
                if (cur == FRIES_EMPTY_KEY) {
                    if (!ini) break;
                    const unsigned long long old = atomicCAS((unsigned long long *)&v.tkeys[slot], FRIES_EMPTY_KEY, key);
                    if (old == FRIES_EMPTY_KEY || old == key) {
                        result = (uint32_t)slot;
                        break;
                    }
                }
                slot = (slot + 1) & v.tmask;
            }
            // a slot without a position (empty until this batch) remembers the smallest batch index of an initiator element
            // that asked for it: the element at which the reference's sequential loop creates the entry
            if (result != FRIES_NO_POS && v.tpos[result] >= FR_DET_PENDING && ini)
                atomicMin(&v.tpos[result], FR_DET_PENDING | (uint32_t)i);
        }
        slot_out[i] = result;
        iota[i] = (uint32_t)i;
    }
}

// After all insertions: which elements are the first initiator of a new determinant; and a non-initiator that missed the
// key only because its initiator sibling had not inserted it yet is looked up again, so that the slot of every element is a
// function of the batch alone (found <=> the determinant is in the index once the whole batch is in)
__global__ void __launch_bounds__(FR_VEC_BLOCK)
merge_det_flag_kernel(VecView v, MergeSrc src, uint32_t *__restrict__ slot, uint32_t *__restrict__ flag) {
    __shared__ uint32_t s_scr[64];
    load_scr(s_scr, v.scr_vec);
    const size_t n = src.count(), n_max = src.n_max;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_max; i += stride) {
        uint32_t s = slot[i];
        if (s == FRIES_NO_POS && i < n) {
            uint64_t k = FRIES_EMPTY_KEY;
            double val = 0;
            if (src.get(i, k, val) && k != FRIES_EMPTY_KEY && val != 0 && (k >> 63) == 0) {
                uint64_t sl = vec_hash(v, k, s_scr) & v.tmask;
                while (true) {
                    const uint64_t cur = v.tkeys[sl];
                    if (cur == k) {
                        s = (uint32_t)sl;
                        break;
                    }
                    if (cur == FRIES_EMPTY_KEY) break;
                    sl = (sl + 1) & v.tmask;
                }
                slot[i] = s;
            }
        }
        flag[i] = (s != FRIES_NO_POS && v.tpos[s] == (FR_DET_PENDING | (uint32_t)i)) ? 1u : 0u;
    }
}

// flag / rank of the first element of every new determinant -> its storage position (append in batch order)
__global__ void merge_det_assign_kernel(VecView v, const uint32_t *__restrict__ slot, const uint32_t *__restrict__ flag,
                                        const uint32_t *__restrict__ rank, size_t n_max, const uint64_t *__restrict__ tkeys,
                                        unsigned long long n_old) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_max; i += stride) {
        if (!flag[i]) continue;
        const uint32_t s = slot[i];
        const unsigned long long pos = n_old + rank[i];
        if (pos < v.cap) {
            v.keys[pos] = tkeys[s];
            for (unsigned r = 0; r < v.n_vecs; r++) v.vals[(size_t)r * v.cap + pos] = 0.0;
            v.diag[pos] = __longlong_as_double(0x7ff8000000000000ll);
            v.tpos[s] = (uint32_t)pos;
        } else {
            v.tpos[s] = FRIES_OVF_POS;
            atomicAdd(&v.cnt->overflow, 1ull);
        }
    }
}
__global__ void merge_det_bump_kernel(VecCounters *cnt, const uint32_t *flag, const uint32_t *rank, size_t n_max,
                                      unsigned long long cap) {
    unsigned long long n = cnt->n + (n_max ? rank[n_max - 1] + flag[n_max - 1] : 0u);
    cnt->n = n < cap ? n : cap;
}

// sorted by slot, batch order inside a slot: the head of every run adds the run's elements one after the other
__global__ void merge_det_accum_kernel(VecView v, MergeSrc src, const uint32_t *__restrict__ slot_sorted,
                                       const uint32_t *__restrict__ idx_sorted, size_t n_max, unsigned origin, unsigned dest,
                                       unsigned long long n_old) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long nonini = 0, valid = 0;
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_max; j += stride) {
        const uint32_t s = slot_sorted[j];
        if (s == FRIES_NO_POS || (j > 0 && slot_sorted[j - 1] == s)) continue;
        const uint32_t pos = v.tpos[s];
        if (pos >= FR_DET_PENDING) continue;  // never got a position: only non-initiators asked, or the store is full
        double *dst = &v.vals[(size_t)dest * v.cap + pos];
        const double *org = &v.vals[(size_t)origin * v.cap + pos];
        double acc = *dst;
        bool exists = pos < n_old;  // a determinant made by this batch exists from its first initiator element on
        for (size_t t = j; t < n_max && slot_sorted[t] == s; t++) {
            uint64_t k;
            double val;
            src.get(idx_sorted[t], k, val);
            const bool ini = (k >> 63) != 0;
            if (!exists && !ini) continue;  // HashTable::read without the initiator flag finds nothing yet (det_hash.hpp:60-94)
            exists = true;
            const bool nonz = (origin == dest ? acc : *org) != 0;  // vec_utils.hpp:632-637 on the running value
            valid++;
            if (ini || nonz) acc += val;
            if (!ini && nonz) nonini++;
        }
        *dst = acc;
    }
    nonini = warp_sum_u64(nonini);
    valid = warp_sum_u64(valid);
    if ((threadIdx.x & 31) == 0) {
        if (nonini) atomicAdd(&v.cnt->nonini_occ_add, nonini);
        if (valid) atomicAdd(&v.cnt->n_spawn_valid, valid);
    }
}

int fries_vec_merge_det_dev(fries_vec *vec, const MergeSrc &src, unsigned origin, unsigned dest) {
    fries_ctx *c = vec->ctx;
    const size_t n_max = src.n_max;
    FRIES_REQUIRE(n_max < FR_DET_PENDING - 1 && vec->cap < FR_DET_PENDING, "deterministic merge: batch or store beyond 2^31");
    FRIES_TRY(vec->slot_scratch.ensure(n_max));
    FRIES_TRY(vec->det_u32.ensure(5 * n_max));
    uint32_t *iota = vec->det_u32.p, *flag = iota + n_max, *rank = flag + n_max, *slot_sorted = rank + n_max,
             *idx_sorted = slot_sorted + n_max;
    size_t tmp_scan = 0, tmp_sort = 0;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, flag, rank, (int)n_max, c->stream));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, vec->slot_scratch.p, slot_sorted, iota, idx_sorted, (int)n_max, 0, 32,
                                             c->stream));
    FRIES_TRY(vec->det_tmp.ensure(tmp_scan > tmp_sort ? tmp_scan : tmp_sort));
    VecView v = vec->view();
    VecCounters cnt;
    FRIES_TRY(vec->read_counters(&cnt));
    const size_t want = (n_max + FR_VEC_BLOCK - 1) / FR_VEC_BLOCK;
    const int grid = (int)(want < (size_t)c->sm_count * 8 ? want : (size_t)c->sm_count * 8);
    {
        ProfScope ps(c, "merge_insert");
        merge_det_insert_kernel<<<grid, FR_VEC_BLOCK, 0, c->stream>>>(v, src, vec->slot_scratch.p, iota);
        merge_det_flag_kernel<<<grid, FR_VEC_BLOCK, 0, c->stream>>>(v, src, vec->slot_scratch.p, flag);
        size_t tb = vec->det_tmp.n;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(vec->det_tmp.p, tb, flag, rank, (int)n_max, c->stream));
        merge_det_assign_kernel<<<grid, FR_VEC_BLOCK, 0, c->stream>>>(v, vec->slot_scratch.p, flag, rank, n_max, v.tkeys,
                                                                      (unsigned long long)cnt.n);
        merge_det_bump_kernel<<<1, 1, 0, c->stream>>>(vec->cnt.p, flag, rank, n_max, (unsigned long long)vec->cap);
        c->launch_count += 4;
    }
    {
        ProfScope ps(c, "merge_accum");
        size_t tb = vec->det_tmp.n;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(vec->det_tmp.p, tb, vec->slot_scratch.p, slot_sorted, iota, idx_sorted, (int)n_max, 0,
                                                 32, c->stream));
        merge_det_accum_kernel<<<grid, FR_VEC_BLOCK, 0, c->stream>>>(v, src, slot_sorted, idx_sorted, n_max, origin, dest,
                                                                     (unsigned long long)cnt.n);
        c->launch_count++;
    }
    CUDA_TRY(cudaGetLastError());
    return FRIES_OK;
}
