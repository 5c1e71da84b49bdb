// Device-resident determinant store (a2/a3): dense SoA storage + open-addressing index keyed by the
// reference's det_hash.
//
// Layout in HBM (capacity C, n_vecs value rows, table size T = 2^k >= 2C):
//   keys  u64[C]            determinant bit strings in storage order (DistVec::indices_)
//   vals  f64[n_vecs][C]    value rows (DistVec::values_)
//   diag  f64[C]            cached diagonal matrix elements, NaN = not yet computed (DistVec::matr_el_)
//   tkeys u64[T], tpos u32[T]   linear-probing index: slot = hash_fxn(occ; vec_scrambler) & (T-1)
// keys/vals/diag are double-buffered: deletion is a stable stream compaction into the other buffer
// followed by a rebuild of the index (no tombstones).
#pragma once
#include "common.cuh"

struct VecCounters {
    unsigned long long n;               // curr_size
    unsigned long long nonini_occ_add;  // DistVec::nonini_occ_add
    unsigned long long overflow;        // insertions refused because the store is full
    unsigned long long n_spawn_valid;   // elements seen by the last merge
    unsigned long long bad_keys;        // uploaded keys with a wrong electron count
};

struct VecView {
    uint64_t *keys;
    double *vals;  // row r at vals + r * cap
    double *diag;
    uint64_t *tkeys;
    uint32_t *tpos;
    uint64_t tmask;
    size_t cap;
    unsigned n_vecs;
    const uint32_t *scr_vec;   // device, 64 entries
    const uint32_t *scr_proc;  // device, 64 entries
    VecCounters *cnt;
    unsigned hh_sites, hh_ph_bits;  // Hubbard-Holstein keys (hh_vec.hpp): 2 n_sites electron bits + n_sites phonon fields; 0 = molecular
};

struct fries_vec {
    fries_ctx *ctx = nullptr;
    size_t cap = 0, tsize = 0;
    unsigned n_bits = 0, n_elec = 0, n_vecs = 0;
    unsigned hh_sites = 0, hh_ph_bits = 0;
    int n_ranks = 1, rank = 0;
    DevBuf<uint64_t> keys[2];
    DevBuf<double> vals[2];
    DevBuf<double> diag[2];
    int cur = 0;
    DevBuf<uint64_t> tkeys;
    DevBuf<uint32_t> tpos;
    DevBuf<uint32_t> scr;  // [0..63] vec scrambler, [64..127] proc scrambler
    DevBuf<VecCounters> cnt;
    DevBuf<uint32_t> slot_scratch;
    bool merge_many_new = false;         // hint of the caller (full H.v): most elements of the next merges create determinants
    bool deterministic = false;          // reproducible merge (vec_det.cu): append and add in batch order
    DevBuf<uint32_t> det_u32;            // its scratch
    DevBuf<uint8_t> det_tmp;
    DevBuf<double> red_d;               // reduction partials
    DevBuf<unsigned long long> red_c;
    size_t min_del_idx = 0;
    size_t n_dense = 0;                  // semi-stochastic: the first n_dense stored determinants form the dense subspace
    size_t n_dense_total = 0;            // ... summed over the ranks (fries_vec_set_dense_total)
    fries_mol *diag_mol = nullptr;
    double hf_en = 0;
    uint64_t last_spawned = 0;
    unsigned cur_row = 0;                  // frifull_mol: row holding the current iterate (vec_idx)
    struct fries_hbpp *hv_scratch = nullptr;  // spawn buffers of the stand-alone fries_h_apply
    std::vector<uint32_t> h_scr_vec, h_scr_proc;

    VecView view() {
        VecView v;
        v.keys = keys[cur].p;
        v.vals = vals[cur].p;
        v.diag = diag[cur].p;
        v.tkeys = tkeys.p;
        v.tpos = tpos.p;
        v.tmask = tsize - 1;
        v.cap = cap;
        v.n_vecs = n_vecs;
        v.scr_vec = scr.p;
        v.scr_proc = scr.p + 64;
        v.cnt = cnt.p;
        v.hh_sites = hh_sites;
        v.hh_ph_bits = hh_ph_bits;
        return v;
    }
    int read_counters(VecCounters *out);
};

// HashTable::hash_fxn with phonon numbers (det_hash.hpp:160-170; HubHolVec::idx_to_hash hh_vec.hpp:72-88):
// occupied electron orbitals first, then the phonon number of every site, both through the same scrambler
__host__ __device__ __forceinline__ uint64_t fr_det_hash_hh(uint64_t key, const uint32_t *scr, unsigned n_sites,
                                                            unsigned ph_bits) {
    uint64_t elec = key & ((1ull << (2 * n_sites)) - 1);
    uint64_t h = fr_det_hash(elec, scr);
    uint64_t ph = key >> (2 * n_sites);
    for (uint32_t i = 0; i < n_sites; i++) {
        uint32_t num = (uint32_t)(ph & ((1u << ph_bits) - 1));
        ph >>= ph_bits;
        h = FRIES_HASH_PRIME * h + (uint32_t)((i + 1) * scr[num]);
    }
    return h;
}
__host__ __device__ __forceinline__ uint64_t vec_hash(const VecView &v, uint64_t key, const uint32_t *s_scr) {
    return v.hh_sites ? fr_det_hash_hh(key, s_scr, v.hh_sites, v.hh_ph_bits) : fr_det_hash(key, s_scr);
}

// find the storage position of `key` (flag bit already stripped); FRIES_NO_POS if absent.
// HashTable::read(create = false) det_hash.hpp:60-94
__device__ __forceinline__ uint32_t vec_lookup(const VecView &v, uint64_t key, const uint32_t *s_scr) {
    uint64_t slot = vec_hash(v, key, s_scr) & v.tmask;
    while (true) {
        uint64_t cur = v.tkeys[slot];
        if (cur == key) {
            const uint32_t p = v.tpos[slot];
            return p >= FRIES_OVF_POS ? FRIES_NO_POS : p;  // (an entry the full store could not place)
        }
        if (cur == FRIES_EMPTY_KEY) return FRIES_NO_POS;
        slot = (slot + 1) & v.tmask;
    }
}

// Where a merge reads its (key | INI, value) pairs from: one flat list with an optional device-side count, or
// the receive buffer of the all-to-all: n_seg segments [keys[seg_cap] | vals[seg_cap]] with per-segment counts.
#define FR_VEC_BLOCK 256
__device__ __forceinline__ void load_scr(uint32_t *s_scr, const uint32_t *g_scr) {
    if (threadIdx.x < 64) s_scr[threadIdx.x] = g_scr[threadIdx.x];
    __syncthreads();
}

struct MergeSrc {
    const uint64_t *keys;
    const double *vals;
    size_t n_max;                          // flat: capacity; segmented: n_seg * seg_cap
    const unsigned long long *d_n;         // flat: optional device count
    const unsigned long long *seg_counts;  // segmented: elements per segment (nullptr = flat)
    size_t seg_cap;
    __device__ __forceinline__ size_t count() const {
        if (seg_counts || !d_n) return n_max;
        unsigned long long dn = *d_n;
        return dn < n_max ? (size_t)dn : n_max;
    }
    __device__ __forceinline__ bool get(size_t i, uint64_t &k, double &v) const {
        if (seg_counts) {
            size_t seg = i / seg_cap, j = i - seg * seg_cap;
            if (j >= seg_counts[seg]) return false;
            const uint64_t *base = keys + seg * 2 * seg_cap;
            k = base[j];
            v = __longlong_as_double((long long)base[seg_cap + j]);
            return true;
        }
        k = keys[i];
        v = vals[i];
        return true;
    }
};
int fries_vec_merge_dev(fries_vec *vec, const uint64_t *d_keys, const double *d_vals, size_t n_max,
                        const unsigned long long *d_n, unsigned origin, unsigned dest);
int fries_vec_merge_src_dev(fries_vec *vec, const MergeSrc &src, unsigned origin, unsigned dest);
int fries_vec_merge_det_dev(fries_vec *vec, const MergeSrc &src, unsigned origin, unsigned dest);  // vec_det.cu
int fries_vec_compact_dev(fries_vec *vec);
