// a16/a17/a18: deterministic H.v, spawning, death/cloning and the resident FRI iterations of
// frisys_mol (FRIES_bin/frisys_mol.cpp:405-552) and frifull_mol (FRIES_bin/frifull_mol.cpp:256-320).
#include "hbpp.cuh"
#include "hv_prov.cuh"
#include "vec.cuh"

extern __shared__ double fr_dyn_smem[];

int fries_vec_compact_flags_dev(fries_vec *vec, const uint8_t *d_flags);
int fries_vec_phase_dev(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, double eps, double shift, unsigned target_nonz,
                        double uniform, bool do_death);  // vecphase.cu

#define FR_NAN __longlong_as_double(0x7ff8000000000000ll)

extern "C" int fries_vec_set_diag_mol(fries_vec *vec, fries_mol *mol, double hf_en) {
    FRIES_REQUIRE(vec && mol, "fries_vec_set_diag_mol: NULL argument");
    FRIES_REQUIRE(vec->n_bits == 2 * mol->view.d.n_orb && vec->n_elec == mol->view.d.n_elec,
                  "fries_vec_set_diag_mol: vector (%u bits, %u electrons) does not match the molecule (%u, %u)",
                  vec->n_bits, vec->n_elec, 2 * mol->view.d.n_orb, mol->view.d.n_elec);
    vec->diag_mol = mol;
    vec->hf_en = hf_en;
    return FRIES_OK;
}

// DistVec::matr_el_at_pos vec_utils.hpp:672-677 with diag_calc_ = diag_matrel - hf_en
__device__ __forceinline__ double diag_at(const MolView &m, const VecView &v, size_t i, double hf_en) {
    double d = v.diag[i];
    if (isnan(d)) {
        uint8_t occ[FRIES_MAX_ELEC + 1];
        mol_occ_list(v.keys[i], occ);
        d = mol_diag(m, occ) - hf_en;
        v.diag[i] = d;
    }
    return d;
}

// ---------------------------------------------------------------------------------------------------
// death/cloning + add_vecs(0,1) + zero_vec(1): frisys_mol.cpp:488-499, one streaming pass.
// 32 B/element: read v0, diag, v1; write v0 (v1 is rewritten only where it was nonzero).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
death_axpy_kernel(MolView gm, VecView v, double hf_en, double eps, double shift) {
    MolView m = mol_stage_shared(gm, fr_dyn_smem);
    unsigned long long n64 = v.cnt->n;
    size_t n = n64 < v.cap ? (size_t)n64 : v.cap;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    double *v0 = v.vals, *v1 = v.vals + v.cap;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double a = v0[i], b = v1[i];
        if (a != 0) {
            double d = diag_at(m, v, i, hf_en);
            a *= 1 - eps * (d - shift);
        }
        a += b;
        v0[i] = a;
        if (b != 0) v1[i] = 0;
    }
}

// h_op_diag molecule.cpp:205-219
__global__ void __launch_bounds__(256)
h_diag_kernel(MolView gm, VecView v, double hf_en, unsigned src, unsigned dst, double id_fac, double h_fac) {
    MolView m = mol_stage_shared(gm, fr_dyn_smem);
    unsigned long long n64 = v.cnt->n;
    size_t n = n64 < v.cap ? (size_t)n64 : v.cap;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    const double *vs = v.vals + (size_t)src * v.cap;
    double *vd = v.vals + (size_t)dst * v.cap;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double a = vs[i];
        double r = 0;
        if (a != 0) r = a * (id_fac + h_fac * diag_at(m, v, i, hf_en));
        vd[i] = r;
    }
}

// ---------------------------------------------------------------------------------------------------
// h_op_offdiag molecule.cpp:448-665 -- one warp per parent determinant.
//
// The reference materialises each parent's excitation list (sing_ex_symm / doub_ex_symm) and walks
// it.  Here the 32 lanes of a warp split the (occupied pair, first virtual) triples of the parent;
// for a triple the allowed second virtuals are one AND of the parent's virtual mask with the irrep
// mask, so counting is a popcount and enumeration a ctz loop.  Pass A counts per lane, a warp scan
// turns the counts into write offsets inside the parent's slot range, pass B writes
// (new determinant | INI, sign * element * v * h_fac).  The order inside a parent differs from the
// reference's list order; the merged vector does not depend on it.
// ---------------------------------------------------------------------------------------------------

// pass 1: number of connections per parent (0 for zero-valued parents)
__global__ void __launch_bounds__(256)
hv_count_kernel(MolView gm, VecView v, unsigned src, size_t n_parents, uint32_t *__restrict__ counts) {
    MolView m = mol_stage_shared(gm, fr_dyn_smem);
    const unsigned lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t p = warp; p < n_parents; p += nwarps) {
        double val = v.vals[(size_t)src * v.cap + p];
        unsigned c = 0;
        if (val != 0) {
            ParentCtx pc;
            parent_ctx(m, v.keys[p], pc);
            c = lane_excitations(m, pc, lane, [](bool, unsigned, unsigned, unsigned, unsigned) {});
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
        if (lane == 0) counts[p] = c;
    }
}

// exclusive scan of u32 counts into u64 offsets (cooperative, deterministic)
__global__ void __launch_bounds__(FR_COMP_BLOCK)
scan_counts_kernel(const uint32_t *__restrict__ counts, size_t n, unsigned long long *__restrict__ offs,
                   unsigned long long *total_out, double *part_d, unsigned long long *part_c) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double sh_d[34];
    __shared__ unsigned long long sh_c[34];
    __shared__ double sh_sd[68];
    __shared__ unsigned long long sh_sc[68];
    GridRed red{part_d, part_c, 0, (int)gridDim.x, sh_d, sh_c};
    size_t chunk = (n + gridDim.x - 1) / gridDim.x;
    chunk = (chunk + 31) & ~(size_t)31;
    const size_t lo = (size_t)blockIdx.x * chunk < n ? (size_t)blockIdx.x * chunk : n;
    const size_t hi = lo + chunk < n ? lo + chunk : n;
    unsigned long long c = 0;
    for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) c += counts[i];
    c = block_sum_u64(c, sh_c);
    double d0, d1;
    unsigned long long blk_off, total;
    grid_excl_scan(grid, red, 0.0, c, d0, blk_off, d1, total, sh_sd, sh_sc);
    unsigned long long carry = blk_off;
    for (size_t base = lo; base < hi; base += blockDim.x) {
        size_t i = base + threadIdx.x;
        unsigned long long mine = i < hi ? counts[i] : 0ull;
        double ex, tot;
        unsigned long long ec, tc;
        block_excl_scan(0.0, mine, ex, ec, tot, tc, sh_sd, sh_sc);
        if (i < hi) offs[i] = carry + ec;
        carry += tc;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *total_out = total;
}

// pass 2: write the connections whose global index falls in the window [win_lo, win_lo + win_len)
__global__ void __launch_bounds__(256)
hv_fill_kernel(MolView gm, VecView v, unsigned src, size_t n_parents, const uint32_t *__restrict__ counts,
               const unsigned long long *__restrict__ offs, unsigned long long win_lo, unsigned long long win_len,
               double h_fac, uint64_t *__restrict__ out_keys, double *__restrict__ out_vals, HbSpawnArgs sp) {
    MolView m = mol_stage_shared(gm, fr_dyn_smem);
    __shared__ uint32_t s_pscr[64];
    if (sp.n_ranks > 1) {
        if (threadIdx.x < 64) s_pscr[threadIdx.x] = sp.proc_scr[threadIdx.x];
        __syncthreads();
    }
    const unsigned lane = threadIdx.x & 31;
    size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t p = warp; p < n_parents; p += nwarps) {
        unsigned cn = counts[p];
        if (cn == 0) continue;
        unsigned long long off = offs[p];
        if (off + cn <= win_lo || off >= win_lo + win_len) continue;
        double val = v.vals[(size_t)src * v.cap + p];
        ParentCtx pc;
        parent_ctx(m, v.keys[p], pc);
        uint8_t occ[FRIES_MAX_ELEC + 1];
        mol_occ_list(pc.key, occ);
        unsigned mine = lane_excitations(m, pc, lane, [](bool, unsigned, unsigned, unsigned, unsigned) {});
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        unsigned long long pos = off + (incl - mine);
        lane_excitations(m, pc, lane, [&](bool dbl, unsigned o0, unsigned o1, unsigned v0, unsigned v1) {
            if (pos >= win_lo && pos < win_lo + win_len) {
                uint64_t nk;
                double el = hv_connection(m, pc.key, occ, dbl, o0, o1, v0, v1, val, h_fac, nk);
                if (sp.n_ranks <= 1) {
                    out_keys[pos - win_lo] = nk | FRIES_INI_FLAG;
                    out_vals[pos - win_lo] = el;
                } else {
                    // Adder::add (vec_utils.hpp:957-971): straight into segment `rank` of the owner's receive window;
                    // the lanes that are active here and share a destination take their slots with one atomic
                    int owner = (int)(fr_det_hash(nk, s_pscr) % (unsigned)sp.n_ranks);
                    unsigned am = __activemask();
                    unsigned peers = __match_any_sync(am, owner);
                    int leader = __ffs(peers) - 1;
                    unsigned rank_in = __popc(peers & ((1u << lane) - 1));
                    unsigned long long base = 0;
                    if ((int)lane == leader) base = atomicAdd(&sp.send_counts[owner], (unsigned long long)__popc(peers));
                    base = __shfl_sync(peers, base, leader);
                    unsigned long long slot = base + rank_in;
                    if (slot < sp.seg_cap) {
                        uint64_t *seg = sp.peer_win[owner] + (size_t)sp.rank * 2 * sp.seg_cap;
                        seg[slot] = nk | FRIES_INI_FLAG;
                        seg[sp.seg_cap + slot] = (uint64_t)__double_as_longlong(el);
                    } else {
                        atomicAdd(&sp.send_counts[sp.n_ranks], 1ull);
                    }
                }
            }
            pos++;
        });
    }
}

__global__ void xrank_stats_kernel(CommView cm, double *scal, const VecCounters *cnt, const CompState *st, double *out);
// tiny collective helpers over the inboxes: one CTA per rank
int fries_comm_route_publish(fries_comm *cm, const unsigned long long *d_send_counts);
int fries_comm_route_wait(fries_comm *cm, unsigned long long *d_recv_counts);
int fries_vec_merge_src_dev(fries_vec *vec, const MergeSrc &src, unsigned origin, unsigned dest);
__global__ void xrank_max_kernel(CommView cm, double mine, double *out) {
    __shared__ double sh_x[1][FR_MAX_RANKS];
    CommCursor cur = comm_begin(cm);
    comm_allgather_v(cm, cur, &mine, 1, sh_x);
    if (threadIdx.x == 0) {
        double t = 0;
        for (int p = 0; p < cm.n_ranks; p++) t = fmax(t, sh_x[0][p]);
        *out = t;
    }
    __syncthreads();
    comm_end(cm, cur);
}
static int xrank_max_u64(fries_hbpp *hb, unsigned long long mine, unsigned long long *out) {
    fries_ctx *c = hb->ctx;
    xrank_max_kernel<<<1, 64, 0, c->stream>>>(fries_comm_view(hb->comm), (double)mine, hb->scal.p + 40);
    c->launch_count++;
    double r = 0;
    CUDA_TRY(cudaMemcpyAsync(&r, hb->scal.p + 40, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out = (unsigned long long)r;
    return FRIES_OK;
}
static int xrank_barrier(fries_hbpp *hb) {
    fries_ctx *c = hb->ctx;
    xrank_max_kernel<<<1, 64, 0, c->stream>>>(fries_comm_view(hb->comm), 0.0, hb->scal.p + 41);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    return FRIES_OK;
}

struct HvScratch {
    DevBuf<uint32_t> counts;
    DevBuf<unsigned long long> offs, total;
};

// dest <- id_fac * src + h_fac * H * src.  Spawn buffers: hb->spawn_keys / spawn_vals (cap entries).
static int h_apply_dev(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, unsigned src, unsigned dest, double id_fac,
                       double h_fac, bool do_diag, uint64_t *n_spawned, size_t only_first = (size_t)-1) {
    fries_ctx *c = vec->ctx;
    FRIES_REQUIRE(src < vec->n_vecs && dest < vec->n_vecs && src != dest, "h_apply: need two different rows");
    VecCounters cnt;
    FRIES_TRY(vec->read_counters(&cnt));
    size_t n_parents = (size_t)cnt.n;
    if (only_first < n_parents) n_parents = only_first;  // the dense subspace of a semi-stochastic run (may be empty on a rank)
    if (n_spawned) *n_spawned = 0;
    // A rank that stores nothing (every non-owner rank when frifull_mol starts from the Hartree-Fock determinant) has no
    // connection of its own but must still take part in every collective below: the round count, the publish / wait /
    // merge of every window (it receives its share of the others' connections) and the barrier after each merge.
    if (n_parents == 0 && vec->n_ranks <= 1) return FRIES_OK;
    size_t smem = (size_t)mol->view.d.blob_doubles * 8;
    VecView v = vec->view();
    int grid = c->sm_count * 8;
    HvScratch sc;
    unsigned long long total = 0;
    if (n_parents > 0) {
        if (do_diag) {
            ProfScope ps(c, "h_diag");
            h_diag_kernel<<<grid, 256, smem, c->stream>>>(mol->view, v, vec->hf_en, src, dest, id_fac, h_fac);
            c->launch_count++;
        }
        FRIES_TRY(sc.counts.alloc(n_parents));
        FRIES_TRY(sc.offs.alloc(n_parents));
        FRIES_TRY(sc.total.alloc(1));
        {
            ProfScope ps(c, "hv_count");
            hv_count_kernel<<<grid, 256, smem, c->stream>>>(mol->view, v, src, n_parents, sc.counts.p);
            c->launch_count++;
        }
        {
            int sgrid = c->coop_grid((const void *)scan_counts_kernel, FR_COMP_BLOCK, 0);
            const uint32_t *cp = sc.counts.p;
            unsigned long long *op = sc.offs.p, *tp = sc.total.p;
            double *pd = hb->part_d.p;
            unsigned long long *pc = hb->part_c.p;
            void *args[] = {(void *)&cp, (void *)&n_parents, (void *)&op, (void *)&tp, (void *)&pd, (void *)&pc};
            ProfScope ps(c, "hv_scan");
            CUDA_TRY(cudaLaunchCooperativeKernel((const void *)scan_counts_kernel, dim3(sgrid), dim3(FR_COMP_BLOCK), args, 0,
                                                 c->stream));
            c->launch_count++;
        }
        CUDA_TRY(cudaMemcpyAsync(&total, sc.total.p, 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    if (n_spawned) *n_spawned = total;
    HbSpawnArgs sp{nullptr, 0, 0, nullptr, nullptr, 1, nullptr, nullptr, nullptr, 0, {nullptr}, 0};
    vec->merge_many_new = only_first == (size_t)-1;  // the full H.v (not the dense subspace of a semi-stochastic run)
    if (vec->n_ranks > 1) {
        // Several ranks: every window of <= seg_cap connections is routed into the owners' receive windows
        // (direct route, comm.cuh) and merged there; all ranks run the same number of windows (the maximum).
        FRIES_REQUIRE(hb->p2p && hb->comm, "h_apply on a partitioned vector needs fries_hbpp_set_route_p2p");
        const unsigned long long win = hb->seg_cap;
        unsigned long long rounds = (total + win - 1) / win, g_rounds = 0;
        FRIES_TRY(xrank_max_u64(hb, rounds, &g_rounds));
        sp = HbSpawnArgs{nullptr, 0, 0, nullptr, nullptr, vec->n_ranks, v.scr_proc, nullptr, hb->send_counts_ext,
                         (unsigned long long)hb->seg_cap, {nullptr}, vec->rank};
        for (int q = 0; q < vec->n_ranks; q++) sp.peer_win[q] = hb->comm->route.win[q];
        for (unsigned long long r = 0; r < g_rounds; r++) {
            unsigned long long lo = r * win, len = lo < total ? (total - lo < win ? total - lo : win) : 0;
            CUDA_TRY(cudaMemsetAsync(hb->send_counts_ext, 0, (vec->n_ranks + 1) * 8, c->stream));
            if (len) {
                ProfScope ps(c, "hv_fill");
                hv_fill_kernel<<<grid, 256, smem, c->stream>>>(mol->view, v, src, n_parents, sc.counts.p, sc.offs.p, lo, len,
                                                               h_fac, nullptr, nullptr, sp);
                c->launch_count++;
                CUDA_TRY(cudaGetLastError());
            }
            FRIES_TRY(fries_comm_route_publish(hb->comm, hb->send_counts_ext));
            FRIES_TRY(fries_comm_route_wait(hb->comm, hb->p2p_recv_counts.p));
            MergeSrc msrc{(const uint64_t *)hb->recv_buf, nullptr, (size_t)vec->n_ranks * hb->seg_cap, nullptr,
                          hb->p2p_recv_counts.p, hb->seg_cap};
            FRIES_TRY(fries_vec_merge_src_dev(vec, msrc, 0, dest));
            // the windows are single-buffered: nobody may store round r + 1 before every rank has merged round r
            FRIES_TRY(xrank_barrier(hb));
            v = vec->view();
        }
        unsigned long long ov = 0;
        CUDA_TRY(cudaMemcpyAsync(&ov, hb->send_counts_ext + vec->n_ranks, 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        FRIES_REQUIRE(ov == 0, "h_apply: %llu connections did not fit a route segment", ov);
        unsigned long long cerr = 0;  // set by a poll that timed out (comm.cuh): a peer died or left the protocol
        CUDA_TRY(cudaMemcpyAsync(&cerr, fries_comm_view(hb->comm).error, 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (cerr) {
            fries_set_error("h_apply: a cross-rank exchange timed out at epoch %llu (a peer left the protocol)", cerr);
            return FRIES_ERR_STATE;
        }
    } else
    for (unsigned long long lo = 0; lo < total; lo += hb->cap) {
        unsigned long long len = total - lo < hb->cap ? total - lo : hb->cap;
        {
            ProfScope ps(c, "hv_fill");
            hv_fill_kernel<<<grid, 256, smem, c->stream>>>(mol->view, v, src, n_parents, sc.counts.p, sc.offs.p, lo, len,
                                                           h_fac, hb->spawn_keys.p, hb->spawn_vals.p, sp);
            c->launch_count++;
        }
        CUDA_TRY(cudaGetLastError());
        // h_op_offdiag adds with ini_flag = 1 and perform_add(0) (molecule.cpp:602,608)
        FRIES_TRY(fries_vec_merge_dev(vec, hb->spawn_keys.p, hb->spawn_vals.p, (size_t)len, nullptr, 0, dest));
        v = vec->view();
    }
    vec->merge_many_new = false;
    CUDA_TRY(cudaStreamSynchronize(c->stream));  // scratch is freed on return
    FRIES_TRY(vec->read_counters(&cnt));
    if (cnt.overflow) {
        fries_set_error("h_apply: determinant store is full (capacity %zu); %llu insertions dropped", vec->cap,
                        cnt.overflow);
        return FRIES_ERR_CAPACITY;
    }
    return FRIES_OK;
}

static int ensure_spawn(fries_hbpp *hb) {
    FRIES_TRY(hb->spawn_keys.ensure(hb->cap));
    FRIES_TRY(hb->spawn_vals.ensure(hb->cap));
    return FRIES_OK;
}

extern "C" int fries_h_apply(fries_vec *vec, fries_mol *mol, unsigned src, unsigned dest, double id_fac, double h_fac) {
    FRIES_REQUIRE(vec && mol, "fries_h_apply: NULL argument");
    FRIES_REQUIRE(vec->n_bits == 2 * mol->view.d.n_orb && vec->n_elec == mol->view.d.n_elec,
                  "fries_h_apply: vector does not match the molecule");
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    if (!vec->hv_scratch) {
        size_t cap = vec->cap < (1u << 22) ? (1u << 22) : vec->cap;
        FRIES_TRY(fries_hbpp_alloc(c, cap, &vec->hv_scratch, false));
    }
    FRIES_TRY(ensure_spawn(vec->hv_scratch));
    uint64_t ns = 0;
    FRIES_TRY(h_apply_dev(vec, mol, vec->hv_scratch, src, dest, id_fac, h_fac, true, &ns));
    vec->last_spawned = ns;
    return FRIES_OK;
}
// H.v on a partitioned vector: `hb` carries the direct route (fries_hbpp_set_route_p2p); collective over the ranks
extern "C" int fries_h_apply_routed(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, unsigned src, unsigned dest,
                                    double id_fac, double h_fac, uint64_t *n_spawned) {
    FRIES_REQUIRE(vec && mol && hb, "fries_h_apply_routed: NULL argument");
    FRIES_REQUIRE(vec->n_bits == 2 * mol->view.d.n_orb && vec->n_elec == mol->view.d.n_elec,
                  "fries_h_apply_routed: vector does not match the molecule");
    CUDA_TRY(cudaSetDevice(vec->ctx->device));
    FRIES_TRY(ensure_spawn(hb));
    uint64_t ns = 0;
    FRIES_TRY(h_apply_dev(vec, mol, hb, src, dest, id_fac, h_fac, true, &ns));
    if (n_spawned) *n_spawned = ns;
    return FRIES_OK;
}
extern "C" int fries_h_apply_last_spawned(fries_vec *vec, uint64_t *n) {
    FRIES_REQUIRE(vec && n, "NULL argument");
    *n = vec->last_spawned;
    return FRIES_OK;
}

// ---------------------------------------------------------------------------------------------------
// frisys_mol
// ---------------------------------------------------------------------------------------------------
extern "C" int fries_frisys_mol_setup(fries_vec *vec, fries_mol *mol, size_t spawn_cap, const uint64_t *h_trial_keys,
                                      const double *h_trial_vals, size_t n_trial, const uint64_t *h_htrial_keys,
                                      const double *h_htrial_vals, size_t n_htrial, fries_hbpp **out) {
    FRIES_REQUIRE(vec && mol && out, "fries_frisys_mol_setup: NULL argument");
    FRIES_REQUIRE(n_trial == 0 || (h_trial_keys && h_trial_vals), "fries_frisys_mol_setup: NULL trial vector");
    FRIES_REQUIRE(n_htrial == 0 || (h_htrial_keys && h_htrial_vals), "fries_frisys_mol_setup: NULL H*trial vector");
    FRIES_REQUIRE(vec->n_vecs >= 2, "fries_frisys_mol_setup: the vector needs two value rows");
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    fries_hbpp *hb = nullptr;
    FRIES_TRY(fries_hbpp_alloc(c, spawn_cap, &hb));
    int rc = ensure_spawn(hb);
    if (rc == FRIES_OK) rc = hb->trial_keys.alloc(n_trial);
    if (rc == FRIES_OK) rc = hb->trial_vals.alloc(n_trial);
    if (rc == FRIES_OK) rc = hb->htrial_keys.alloc(n_htrial);
    if (rc == FRIES_OK) rc = hb->htrial_vals.alloc(n_htrial);
    if (rc == FRIES_OK) rc = hb->keep_flags.alloc(vec->cap);
    if (rc != FRIES_OK) {
        fries_hbpp_destroy(hb);
        return rc;
    }
    hb->n_trial = n_trial;
    hb->n_htrial = n_htrial;
    if (n_trial) {
        CUDA_TRY(cudaMemcpyAsync(hb->trial_keys.p, h_trial_keys, n_trial * 8, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(hb->trial_vals.p, h_trial_vals, n_trial * 8, cudaMemcpyHostToDevice, c->stream));
    }
    if (n_htrial) {
        CUDA_TRY(cudaMemcpyAsync(hb->htrial_keys.p, h_htrial_keys, n_htrial * 8, cudaMemcpyHostToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(hb->htrial_vals.p, h_htrial_vals, n_htrial * 8, cudaMemcpyHostToDevice, c->stream));
    }
    CUDA_TRY(cudaMemsetAsync(hb->keep_flags.p, 0, vec->cap, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    vec->diag_mol = mol;
    *out = hb;
    return FRIES_OK;
}

__global__ void trial_dot_kernel(VecView v, const uint64_t *__restrict__ t_keys, const double *__restrict__ t_vals,
                                 size_t n_trial, unsigned row, double *out);

struct IterScalars {  // hb->scal layout (doubles)
    enum { R4 = 0, NUMER = 4, DENOM = 5, NEW_NORM = 6, STATS = 8, DENSE_NORM = 43 };
};

__global__ void iter_stats_kernel(const CompState *st, const VecCounters *cnt, const double *scal, double *out,
                                  const unsigned long long *comm_err) {
    // out[0..]: glob_norm, numer, denom, n_kept, n_matrix_samples, curr_size, overflow flags, anomalies
    out[0] = st[6].glob_norm + scal[IterScalars::DENSE_NORM];  // frisys_mol.cpp:504: glob_norm += dense_norm()
    out[1] = scal[IterScalars::NUMER];
    out[2] = scal[IterScalars::DENOM];
    out[3] = (double)st[6].n_kept;
    out[4] = (double)st[5].n_out;
    out[5] = (double)cnt->n;
    unsigned long long ov = 0, an = 0;
    for (int s = 0; s < 5; s++) {
        ov += st[s].overflow;
        an += st[s].anomalies;
    }
    out[6] = (double)ov;
    out[7] = (double)cnt->overflow;
    out[8] = (double)an;
    out[9] = (double)st[6].n_samp_left;
    out[10] = st[6].loc_norm;
    out[11] = comm_err ? (double)*comm_err : 0.0;  // epoch of a cross-rank poll that timed out (comm.cuh)
}

__global__ void state_to_r4_kernel(const CompState *st, double *r4) {
    r4[0] = st->loc_norm;
    r4[1] = st->glob_norm;
    r4[2] = (double)st->n_samp_left;
    r4[3] = (double)st->n_kept;
}

// number of compressible elements: stored minus the dense subspace (frisys_mol.cpp:419,503)
__global__ void dense_count_kernel(const VecCounters *cnt, unsigned long long n_dense, unsigned long long *out) {
    unsigned long long n = cnt->n;
    *out = n > n_dense ? n - n_dense : 0ull;
}
// sum |v| over the dense subspace (DistVec::dense_norm vec_utils.hpp:903-917); one CTA, fixed order
__global__ void dense_norm_kernel(const double *vals, size_t n_dense, double *out) {
    __shared__ double sh[33];
    double t = 0;
    for (size_t i = threadIdx.x; i < n_dense; i += blockDim.x) t += fabs(vals[i]);
    t = block_sum(t, sh);
    if (threadIdx.x == 0) *out = t;
}
static const unsigned long long *stochastic_count(fries_vec *vec, fries_hbpp *hb) {
    if (vec->n_dense == 0) return &vec->cnt.p->n;
    dense_count_kernel<<<1, 1, 0, vec->ctx->stream>>>(vec->cnt.p, (unsigned long long)vec->n_dense, hb->n_scalar.p + 1);
    vec->ctx->launch_count++;
    return hb->n_scalar.p + 1;
}

static int compress_vector_dev(fries_vec *vec, fries_hbpp *hb, unsigned row, unsigned target_nonz, double uniform) {
    // find_preserve -> sys_comp -> del_at_pos for the zeroed elements (frisys_mol.cpp:501-539); the dense subspace of
    // a semi-stochastic run (the first n_dense positions) is left alone
    fries_ctx *c = vec->ctx;
    VecView v = vec->view();
    const size_t nd = vec->n_dense;
    double *vals = v.vals + (size_t)row * v.cap + nd;
    FRIES_TRY(fries_find_preserve_launch(c, vals, vec->cap - nd, stochastic_count(vec, hb), target_nonz,
                                         hb->keep_flags.p + nd, hb->st.p + 6,
                                         hb->part_d.p, hb->part_c.p, 0, hb->comm, hb->pred.p + 5, hb->cand_x.p,
                                         hb->cand_m.p));
    state_to_r4_kernel<<<1, 1, 0, c->stream>>>(hb->st.p + 6, hb->scal.p + IterScalars::R4);
    c->launch_count++;
    if (nd) {  // the slot is zero otherwise (fries_hbpp_alloc)
        dense_norm_kernel<<<1, 256, 0, c->stream>>>(v.vals + (size_t)row * v.cap, nd, hb->scal.p + IterScalars::DENSE_NORM);
        c->launch_count++;
        hb->dense_norm_set = true;
    } else if (hb->dense_norm_set || vec->n_ranks > 1) {  // (several ranks: the slot holds the global sum after every iteration)
        CUDA_TRY(cudaMemsetAsync(hb->scal.p + IterScalars::DENSE_NORM, 0, sizeof(double), c->stream));
        hb->dense_norm_set = false;
    }
    (void)uniform;
    return FRIES_OK;
}

static int resample_vector_dev(fries_vec *vec, fries_hbpp *hb, unsigned row, double uniform) {
    fries_ctx *c = vec->ctx;
    VecView v = vec->view();
    const size_t nd = vec->n_dense;
    double *vals = v.vals + (size_t)row * v.cap + nd;
    FRIES_TRY(fries_sys_comp_launch(c, vals, vec->cap - nd, stochastic_count(vec, hb), hb->keep_flags.p + nd,
                                    hb->scal.p + IterScalars::R4, 0.0, 0.0, -1LL, uniform, hb->st.p + 7, hb->part_d.p,
                                    hb->part_c.p, 0, hb->comm));
    FRIES_TRY(fries_vec_compact_flags_dev(vec, hb->keep_flags.p));
    return FRIES_OK;
}

static int dot_dev(fries_vec *vec, const uint64_t *tk, const double *tv, size_t n, unsigned row, double *d_out) {
    fries_ctx *c = vec->ctx;
    trial_dot_kernel<<<1, 1024, 0, c->stream>>>(vec->view(), tk, tv, n, row, d_out);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    return FRIES_OK;
}

static int read_stats(fries_vec *vec, fries_hbpp *hb, fries_iter_stats *stats, const char *who) {
    fries_ctx *c = vec->ctx;
    iter_stats_kernel<<<1, 1, 0, c->stream>>>(hb->st.p, vec->cnt.p, hb->scal.p, hb->scal.p + IterScalars::STATS,
                                              hb->comm ? fries_comm_view(hb->comm).error : nullptr);
    c->launch_count++;
    CUDA_TRY(cudaMemcpyAsync(c->h_pinned, hb->scal.p + IterScalars::STATS, 16 * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const double *h = c->h_pinned;
    if (stats) {
        stats->glob_norm = h[0];
        stats->numer = h[1];
        stats->denom = h[2];
        stats->n_kept = (uint64_t)h[3];
        stats->n_matrix_samples = (uint64_t)h[4];
        stats->n_spawned = (uint64_t)h[4];
        stats->curr_size = (uint64_t)h[5];
    }
    if (h[11] != 0) {
        fries_set_error("%s: a cross-rank exchange timed out at epoch %g (a peer died or left the protocol)", who, h[11]);
        return FRIES_ERR_STATE;
    }
    if (h[6] != 0) {
        fries_set_error("%s: insufficient memory allocated for matrix compression (%g samples dropped; raise spawn_cap)",
                        who, h[6]);
        return FRIES_ERR_CAPACITY;
    }
    if (h[7] != 0) {
        fries_set_error("%s: determinant store is full (capacity %zu); %g insertions dropped", who, vec->cap, h[7]);
        return FRIES_ERR_CAPACITY;
    }
    return FRIES_OK;
}

extern "C" int fries_frisys_mol_iterate(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, const fries_frisys_params *p,
                                        const double *u6, fries_iter_stats *stats) {
    FRIES_REQUIRE(vec && mol && hb && p && u6, "fries_frisys_mol_iterate: NULL argument");
    FRIES_REQUIRE(vec->n_ranks == 1, "fries_frisys_mol_iterate: multi-rank vectors go through fries_frisys_mol_spawn + _finish");
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    VecView v = vec->view();
    size_t smem = (size_t)mol->view.d.blob_doubles * 8;
    // steps 2-3: hierarchical compression of H's columns (semi-stochastic: only of the columns outside the dense
    // subspace, frisys_mol.cpp:414-419)
    const size_t nd = vec->n_dense;
    FRIES_TRY(fries_hbpp_stages_dev(hb, mol, v.keys + nd, v.vals + nd, stochastic_count(vec, hb), p->p_doub, p->new_hb, u6,
                                    p->matr_samp));
    // step 5: spawn (fused into finalize) and merge into row 1
    HbSpawnArgs sp{v.vals + nd, p->eps, p->init_thresh, hb->spawn_keys.p, hb->spawn_vals.p, 1, nullptr, nullptr, nullptr, 0,
                   {nullptr}, 0};
    FRIES_TRY(fries_hbpp_finalize_dev(hb, mol, v.keys + nd, p->p_doub, p->new_hb, &sp));
    FRIES_TRY(fries_vec_merge_dev(vec, hb->spawn_keys.p, hb->spawn_vals.p, hb->cap, &hb->st.p[4].n_out, 0, 1));
    if (nd) {
        // deterministic subspace multiplication (frisys_mol.cpp:479-485): every connection of the dense determinants,
        // -eps * <D'|H|D> * v_D with the values from before the spawn (row 0 is untouched so far), into row 1
        FRIES_TRY(h_apply_dev(vec, mol, hb, 0, 1, 0.0, -p->eps, false, nullptr, nd));
        v = vec->view();
    }
    // steps 7, 8, 10, 11 as one cooperative launch (vecphase.cu); FRIES_FUSED_VEC=0 in the environment keeps the seven
    // separate launches below (regression / measurement variant, same arithmetic)
    static const bool fused = [] {
        const char *e = getenv("FRIES_FUSED_VEC");
        return !(e && e[0] == '0' && e[1] == 0);
    }();
    // (at every size: at 1.25e7 stored determinants per GPU the fused kernel takes 2.1 ms against 3.1 ms for the separate
    // ones since its passes were balanced, end of round 2)
    if (fused && vec->n_vecs == 2 && vec->hh_sites == 0) {
        FRIES_TRY(fries_vec_phase_dev(vec, mol, hb, p->eps, p->en_shift, p->target_nonz, u6[5], true));
        return read_stats(vec, hb, stats, "fries_frisys_mol_iterate");
    }
    // step 7: death/cloning, add_vecs(0, 1), zero row 1
    {
        ProfScope ps(c, "death_axpy");
        death_axpy_kernel<<<c->sm_count * 8, 256, smem, c->stream>>>(mol->view, v, vec->hf_en, p->eps, p->en_shift);
        c->launch_count++;
    }
    // step 8: exact preservation
    FRIES_TRY(compress_vector_dev(vec, hb, 0, p->target_nonz, u6[5]));
    // step 10: projected energy (before resampling, frisys_mol.cpp:517-520)
    FRIES_TRY(dot_dev(vec, hb->htrial_keys.p, hb->htrial_vals.p, hb->n_htrial, 0, hb->scal.p + IterScalars::NUMER));
    FRIES_TRY(dot_dev(vec, hb->trial_keys.p, hb->trial_vals.p, hb->n_trial, 0, hb->scal.p + IterScalars::DENOM));
    // step 11: systematic resampling + deletion of the zeroed elements
    FRIES_TRY(resample_vector_dev(vec, hb, 0, u6[5]));
    return read_stats(vec, hb, stats, "fries_frisys_mol_iterate");
}

// ---------------------------------------------------------------------------------------------------
// multi-rank frisys_mol: the iteration is split around the all-to-all that the host issues
// (torch.distributed / NCCL, on the same stream) -- spawn [stages + finalize + pack by owner] -> exchange
// counts + payload -> finish [merge, death/cloning, compression].  Global reductions inside the
// kernels go through the peer-mapped inboxes (comm.cuh).
// ---------------------------------------------------------------------------------------------------
int fries_comm_route_publish(fries_comm *cm, const unsigned long long *d_send_counts);
int fries_comm_route_wait(fries_comm *cm, unsigned long long *d_recv_counts);
extern "C" int fries_hbpp_set_route(fries_hbpp *hb, fries_comm *comm, void *d_send_buf, void *d_recv_buf,
                                    void *d_send_counts, size_t seg_cap) {
    FRIES_REQUIRE(hb && comm && d_send_buf && d_recv_buf && d_send_counts && seg_cap > 0, "fries_hbpp_set_route: bad argument");
    hb->comm = comm;
    hb->send_buf = (int64_t *)d_send_buf;
    hb->recv_buf = (int64_t *)d_recv_buf;
    hb->send_counts_ext = (unsigned long long *)d_send_counts;
    hb->seg_cap = seg_cap;
    return FRIES_OK;
}

// sum (numer, denom) and the per-rank counters over the ranks, in rank order; one CTA
__global__ void xrank_stats_kernel(CommView cm, double *scal, const VecCounters *cnt, const CompState *st, double *out) {
    __shared__ double sh_x0[FR_MAX_RANKS], sh_x1[FR_MAX_RANKS];
    __shared__ unsigned long long sh_xc[FR_MAX_RANKS];
    CommCursor cur = comm_begin(cm);
    comm_allgather(cm, cur, scal[IterScalars::NUMER], scal[IterScalars::DENOM], cnt->n, sh_x0, sh_x1, sh_xc);
    double numer, denom, b;
    comm_sum(cm, sh_x0, numer, b);
    comm_sum(cm, sh_x1, denom, b);
    unsigned long long n_glob = comm_sum_u64(cm, sh_xc);
    comm_allgather(cm, cur, (double)st[5].n_out, scal[IterScalars::DENSE_NORM], cnt->overflow, sh_x0, sh_x1, sh_xc);
    double spawned, dense_norm;
    comm_sum(cm, sh_x0, spawned, b);
    comm_sum(cm, sh_x1, dense_norm, b);  // DistVec::dense_norm vec_utils.hpp:903-917 is a sum over the ranks
    if (threadIdx.x == 0) {
        scal[IterScalars::NUMER] = numer;
        scal[IterScalars::DENOM] = denom;
        scal[IterScalars::DENSE_NORM] = dense_norm;
        out[0] = (double)n_glob;
        out[1] = spawned;
    }
    __syncthreads();
    comm_end(cm, cur);
}

extern "C" int fries_frisys_mol_spawn(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, const fries_frisys_params *p,
                                      const double *u6) {
    FRIES_REQUIRE(vec && mol && hb && p && u6, "fries_frisys_mol_spawn: NULL argument");
    FRIES_REQUIRE(hb->comm && (hb->send_buf || hb->p2p),
                  "fries_frisys_mol_spawn: call fries_hbpp_set_route or fries_hbpp_set_route_p2p first");
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    VecView v = vec->view();
    CUDA_TRY(cudaMemsetAsync(hb->send_counts_ext, 0, (vec->n_ranks + 1) * 8, c->stream));
    // semi-stochastic: only the columns outside this rank's share of the dense subspace are compressed (frisys_mol.cpp:414-419)
    const size_t nd = vec->n_dense;
    FRIES_TRY(fries_hbpp_stages_dev(hb, mol, v.keys + nd, v.vals + nd, stochastic_count(vec, hb), p->p_doub, p->new_hb, u6,
                                    p->matr_samp));
    HbSpawnArgs sp{v.vals + nd, p->eps, p->init_thresh, nullptr, nullptr, vec->n_ranks, v.scr_proc, (uint64_t *)hb->send_buf,
                   hb->send_counts_ext, (unsigned long long)hb->seg_cap, {nullptr}, vec->rank};
    if (hb->p2p)
        for (int q = 0; q < vec->n_ranks; q++) sp.peer_win[q] = hb->comm->route.win[q];
    FRIES_TRY(fries_hbpp_finalize_dev(hb, mol, v.keys + nd, p->p_doub, p->new_hb, &sp));
    // direct route: the elements are already in their owners' windows; publish counts + epoch flag to the peers
    if (hb->p2p) FRIES_TRY(fries_comm_route_publish(hb->comm, hb->send_counts_ext));
    return FRIES_OK;
}

// Direct spawn route over peer-mapped windows (comm.cuh RouteView): no collective call, no host round trip between
// _spawn and _finish.  The window (fries_comm_route_create / _connect) fixes the segment capacity.
extern "C" int fries_hbpp_set_route_p2p(fries_hbpp *hb, fries_comm *comm) {
    FRIES_REQUIRE(hb && comm && comm->win_local, "fries_hbpp_set_route_p2p: create and connect the route window first");
    CUDA_TRY(cudaSetDevice(hb->ctx->device));
    hb->comm = comm;
    hb->p2p = true;
    hb->seg_cap = comm->seg_cap;
    FRIES_TRY(hb->p2p_send_counts.alloc(FR_MAX_RANKS + 1));
    FRIES_TRY(hb->p2p_recv_counts.alloc(FR_MAX_RANKS));
    hb->send_counts_ext = hb->p2p_send_counts.p;
    hb->send_buf = nullptr;
    hb->recv_buf = (int64_t *)comm->win_local;
    return FRIES_OK;
}

extern "C" int fries_frisys_mol_finish(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, const fries_frisys_params *p,
                                       const double *u6, const void *d_recv_counts, fries_iter_stats *stats) {
    FRIES_REQUIRE(vec && mol && hb && p && u6 && (d_recv_counts || hb->p2p), "fries_frisys_mol_finish: NULL argument");
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    VecView v = vec->view();
    size_t smem = (size_t)mol->view.d.blob_doubles * 8;
    if (hb->p2p) {  // wait for every source's epoch flag; the counts arrive with it
        FRIES_TRY(fries_comm_route_wait(hb->comm, hb->p2p_recv_counts.p));
        d_recv_counts = hb->p2p_recv_counts.p;
    }
    MergeSrc src{(const uint64_t *)hb->recv_buf, nullptr, (size_t)vec->n_ranks * hb->seg_cap, nullptr,
                 (const unsigned long long *)d_recv_counts, hb->seg_cap};
    FRIES_TRY(fries_vec_merge_src_dev(vec, src, 0, 1));
    // elements that did not fit a send segment (read now: the dense multiplication below reuses the counters)
    unsigned long long ov = 0;
    CUDA_TRY(cudaMemcpyAsync(&ov, hb->send_counts_ext + vec->n_ranks, 8, cudaMemcpyDeviceToHost, c->stream));
    if (vec->n_dense_total) {
        // deterministic subspace multiplication (frisys_mol.cpp:479-485), collective: every rank routes the connections of
        // its dense determinants to their owners and merges what it receives (row 0 still holds the values from before)
        FRIES_REQUIRE(hb->p2p, "the dense subspace on several ranks needs the direct spawn route (fries_hbpp_set_route_p2p)");
        FRIES_TRY(xrank_barrier(hb));  // the windows are single-buffered: every rank has merged the stochastic spawns
        FRIES_TRY(h_apply_dev(vec, mol, hb, 0, 1, 0.0, -p->eps, false, nullptr, vec->n_dense));
        v = vec->view();
    }
    {
        ProfScope ps(c, "death_axpy");
        death_axpy_kernel<<<c->sm_count * 8, 256, smem, c->stream>>>(mol->view, v, vec->hf_en, p->eps, p->en_shift);
        c->launch_count++;
    }
    FRIES_TRY(compress_vector_dev(vec, hb, 0, p->target_nonz, u6[5]));
    FRIES_TRY(dot_dev(vec, hb->htrial_keys.p, hb->htrial_vals.p, hb->n_htrial, 0, hb->scal.p + IterScalars::NUMER));
    FRIES_TRY(dot_dev(vec, hb->trial_keys.p, hb->trial_vals.p, hb->n_trial, 0, hb->scal.p + IterScalars::DENOM));
    FRIES_TRY(resample_vector_dev(vec, hb, 0, u6[5]));
    xrank_stats_kernel<<<1, 32, 0, c->stream>>>(fries_comm_view(hb->comm), hb->scal.p, vec->cnt.p, hb->st.p,
                                                hb->scal.p + 32);
    c->launch_count++;
    int rc = read_stats(vec, hb, stats, "fries_frisys_mol_finish");
    if (stats) {
        double g[2];
        CUDA_TRY(cudaMemcpyAsync(g, hb->scal.p + 32, 16, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        stats->curr_size = (uint64_t)g[0];          // global number of stored determinants
        stats->n_spawned = stats->n_matrix_samples = (uint64_t)g[1];
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (rc == FRIES_OK && ov) {
        fries_set_error("fries_frisys_mol_finish: %llu spawned elements did not fit the send segments (seg_cap %zu)", ov,
                        hb->seg_cap);
        return FRIES_ERR_CAPACITY;
    }
    return rc;
}

// frifull_mol.cpp:256-320.  The vector alternates between rows: `src` is the row holding the current
// iterate (vec_idx), the result lands in the other row.
extern "C" int fries_frifull_mol_iterate(fries_vec *vec, fries_mol *mol, fries_hbpp *hb, fries_frifull_params *p,
                                         double uniform, fries_iter_stats *stats) {
    FRIES_REQUIRE(vec && mol && hb && p, "fries_frifull_mol_iterate: NULL argument");
    FRIES_REQUIRE(vec->n_ranks == 1 || (hb->p2p && hb->comm),
                  "fries_frifull_mol_iterate: a partitioned vector needs fries_hbpp_set_route_p2p");
    fries_ctx *c = vec->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    unsigned src = vec->cur_row, dst = src ^ 1;
    CUDA_TRY(cudaMemsetAsync(hb->st.p, 0, 8 * sizeof(CompState), c->stream));
    FRIES_TRY(dot_dev(vec, hb->trial_keys.p, hb->trial_vals.p, hb->n_trial, src, hb->scal.p + IterScalars::DENOM));
    FRIES_TRY(compress_vector_dev(vec, hb, src, p->target_nonz, uniform));
    if (p->adjust_shift) {  // needs the (global) one-norm before compression on the host
        double r4[4];
        CUDA_TRY(cudaMemcpyAsync(r4, hb->scal.p + IterScalars::R4, sizeof(r4), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        double one_norm = r4[1];
        if (p->last_one_norm) {
            p->en_shift -= p->damp_factor * log(one_norm / p->last_one_norm);
            p->last_one_norm = one_norm;
        }
        if (p->last_one_norm == 0 && one_norm > p->target_norm) p->last_one_norm = one_norm;
    }
    FRIES_TRY(resample_vector_dev(vec, hb, src, uniform));
    uint64_t ns = 0;
    FRIES_TRY(h_apply_dev(vec, mol, hb, src, dst, 1 + p->eps * p->en_shift, -p->eps, true, &ns));
    vec->cur_row = dst;
    FRIES_TRY(dot_dev(vec, hb->trial_keys.p, hb->trial_vals.p, hb->n_trial, dst, hb->scal.p + IterScalars::NUMER));
    if (vec->n_ranks > 1) {  // <trial|v>, <trial|v'>, sizes and spawn counts summed over the ranks, in rank order
        unsigned long long ns64 = ns;
        CUDA_TRY(cudaMemcpyAsync(&hb->st.p[5].n_out, &ns64, 8, cudaMemcpyHostToDevice, c->stream));
        xrank_stats_kernel<<<1, 32, 0, c->stream>>>(fries_comm_view(hb->comm), hb->scal.p, vec->cnt.p, hb->st.p,
                                                    hb->scal.p + 32);
        c->launch_count++;
    }
    int rc = read_stats(vec, hb, stats, "fries_frifull_mol_iterate");
    if (stats) {
        // numer = ((1 + eps S) denom - <trial|v'>) / eps  (frifull_mol.cpp:294-296)
        stats->numer = ((1 + p->eps * p->en_shift) * stats->denom - stats->numer) / p->eps;
        stats->n_spawned = ns;
        stats->n_matrix_samples = ns;
        if (vec->n_ranks > 1) {
            double g[2];
            CUDA_TRY(cudaMemcpyAsync(g, hb->scal.p + 32, 16, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaStreamSynchronize(c->stream));
            stats->curr_size = (uint64_t)g[0];
            stats->n_spawned = stats->n_matrix_samples = (uint64_t)g[1];
        }
    }
    return rc;
}
