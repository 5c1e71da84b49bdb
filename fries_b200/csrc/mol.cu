// fries_mol: integrals, symmetry and HB-PP tables resident in HBM + the batch parity entry points
// (a7, a8, a9, a11-a14).  The tables of set_up (heat_bathPP.cpp:99-179) are built on the device.
#include "mol.cuh"
#include "molhost.cuh"

// ---------------------------------------------------------------------------------------------------
// set_up heat_bathPP.cpp:99-179.  Inner loops run in the reference's order inside one thread, so
// every table entry is the same FP64 sum as the reference's.
// ---------------------------------------------------------------------------------------------------
__global__ void hb_setup_pairs_kernel(MolView m, double *d_diff, double *d_same) {
    const unsigned M = m.d.n_orb, hf = m.d.tot_orb - M, T = m.d.tot_orb;
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < M * M) {
        unsigned i = t / M, j = t % M;
        double s = 0;
        for (unsigned a = hf; a < T; a++)
            for (unsigned b = hf; b < T; b++)
                if (i != (a - hf) && j != (b - hf)) s += fabs(eri_phys(m, i + hf, j + hf, a, b));
        d_diff[i * M + j] = s;
    }
    if (t < M * (M - 1) / 2) {
        // invert tri index: t = j(j-1)/2 + i, i < j
        unsigned j = 1;
        while (FR_TRI_N(j) <= t) j++;
        unsigned i = t - FR_TRI_N(j - 1);
        double s = 0;
        for (unsigned a = hf; a < T; a++)
            for (unsigned b = hf; b < a; b++)
                if ((a - hf) != j && (a - hf) != i && (b - hf) != j && (b - hf) != i)
                    s += 2 * fabs(eri_phys(m, i + hf, j + hf, a, b) - eri_phys(m, i + hf, j + hf, b, a));
        d_same[t] = s;
    }
}
__global__ void hb_setup_rest_kernel(MolView m, const double *d_diff, const double *d_same, double *s_tens,
                                     double *exch_sqrt, double *diag_sqrt, double *exch_norms, double *s_norm_out) {
    const unsigned M = m.d.n_orb, hf = m.d.tot_orb - M;
    unsigned t = threadIdx.x;
    if (t < M * (M - 1) / 2) {
        unsigned j = 1;
        while (FR_TRI_N(j) <= t) j++;
        unsigned i = t - FR_TRI_N(j - 1);
        exch_sqrt[t] = sqrt(fabs(eri_phys(m, i + hf, j + hf, j + hf, i + hf)));
    }
    if (t < M) diag_sqrt[t] = sqrt(fabs(eri_phys(m, t + hf, t + hf, t + hf, t + hf)));
    __syncthreads();
    if (t < M) {
        unsigned i = t;
        double s = 0;
        for (unsigned j = 0; j < i; j++) s += d_same[FR_TRI_NODIAG(j, i)];
        for (unsigned j = i + 1; j < M; j++) s += d_same[FR_TRI_NODIAG(i, j)];
        for (unsigned j = 0; j < M; j++) s += d_diff[i * M + j];
        s_tens[i] = s;
        double e = 0;
        for (unsigned j = 0; j < i; j++) e += exch_sqrt[FR_TRI_NODIAG(j, i)];
        e += diag_sqrt[i];
        for (unsigned j = i + 1; j < M; j++) e += exch_sqrt[FR_TRI_NODIAG(i, j)];
        exch_norms[i] = e;
    }
    __syncthreads();
    if (t == 0) {
        double s = 0;
        for (unsigned i = 0; i < M; i++) s += s_tens[i];
        *s_norm_out = s;
    }
}

// square copies of exch_sqrt / d_same for the row generators (mol.cuh)
__global__ void hb_square_kernel(MolDims d, double *blob) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x, M = d.n_orb;
    if (t < M * M) mol_square_entry(d, blob, t / M, t % M);
}

extern "C" int fries_mol_create(fries_ctx *c, unsigned n_orb, unsigned n_elec_total, unsigned n_frz,
                                const double *h_hcore, const double *h_eris_packed, const uint8_t *h_symm,
                                fries_mol **out) {
    FRIES_REQUIRE(c && out && h_hcore && h_eris_packed && h_symm, "fries_mol_create: NULL argument");
    FRIES_REQUIRE(n_orb >= 2 && 2 * n_orb <= 63, "fries_mol_create: n_orb %u out of range (2..31)", n_orb);
    FRIES_REQUIRE(n_frz % 2 == 0 && n_elec_total > n_frz && (n_elec_total - n_frz) % 2 == 0,
                  "fries_mol_create: need an even number of frozen and of unfrozen electrons");
    unsigned ne = n_elec_total - n_frz;
    FRIES_REQUIRE(ne <= FRIES_MAX_ELEC && ne / 2 < n_orb, "fries_mol_create: n_elec %u does not fit %u orbitals", ne,
                  n_orb);
    FRIES_REQUIRE(n_orb <= 31 + 1 && n_orb - ne / 2 <= FRIES_MAX_SUB && ne <= FRIES_MAX_SUB,
                  "fries_mol_create: row length exceeds FRIES_MAX_SUB");
    for (unsigned i = 0; i < n_orb; i++)
        FRIES_REQUIRE(h_symm[i] < FR_N_IRREPS, "fries_mol_create: irrep %u of orbital %u not in 0..7", h_symm[i], i);
    CUDA_TRY(cudaSetDevice(c->device));
    const unsigned M = n_orb, T = n_orb + n_frz / 2, TT = M * (M - 1) / 2;
    fries_mol *mol = new fries_mol();
    mol->ctx = c;
    MolDims &d = mol->view.d;
    d.n_orb = M;
    d.n_elec = ne;
    d.n_frz = n_frz;
    d.tot_orb = T;
    const unsigned off = mol_blob_layout(d);
    // symmetry tables (integer bookkeeping; gen_symm_lookup molecule.cpp:1050-1065, SymmInfo molecule.hpp:265-280)
    std::vector<double> h_blob(off, 0.0);
    uint8_t *symm = (uint8_t *)(h_blob.data() + d.off_symm);
    uint8_t *lookup = (uint8_t *)(h_blob.data() + d.off_lookup);
    memcpy(symm, h_symm, M);
    for (unsigned i = 0; i < M; i++) {
        uint8_t s = h_symm[i];
        uint8_t cnt = lookup[s * (M + 1)];
        lookup[s * (M + 1) + 1 + cnt] = (uint8_t)i;
        lookup[s * (M + 1)] = cnt + 1;
    }
    uint32_t *irr_mask = (uint32_t *)(h_blob.data() + d.off_irr);
    for (unsigned i = 0; i < M; i++) irr_mask[h_symm[i] % FR_N_IRREPS] |= 1u << i;
    d.max_n_symm = 0;
    for (unsigned s = 0; s < FR_N_IRREPS; s++)
        if (lookup[s * (M + 1)] > d.max_n_symm) d.max_n_symm = lookup[s * (M + 1)];
    FRIES_REQUIRE(d.max_n_symm <= FRIES_MAX_SUB, "fries_mol_create: max_n_symm exceeds FRIES_MAX_SUB");
    size_t n_pair = (size_t)T * (T + 1) / 2;
    mol->n_packed = n_pair * (n_pair + 1) / 2;
    int rc = mol->eris.alloc(mol->n_packed);
    if (rc == FRIES_OK) rc = mol->hcore.alloc((size_t)T * T);
    if (rc == FRIES_OK) rc = mol->blob.alloc(off + 1);
    if (rc != FRIES_OK) {
        delete mol;
        return rc;
    }
    CUDA_TRY(cudaMemcpyAsync(mol->eris.p, h_eris_packed, mol->n_packed * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(mol->hcore.p, h_hcore, (size_t)T * T * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(mol->blob.p, h_blob.data(), off * 8, cudaMemcpyHostToDevice, c->stream));
    mol->view.eris = mol->eris.p;
    mol->view.hcore = mol->hcore.p;
    mol_bind_blob(mol->view, mol->blob.p);
    double *b = mol->blob.p;
    {
        ProfScope ps(c, "hb_setup");
        hb_setup_pairs_kernel<<<(M * M + 127) / 128, 128, 0, c->stream>>>(mol->view, b + d.off_d_diff, b + d.off_d_same);
        hb_setup_rest_kernel<<<1, 1024, 0, c->stream>>>(mol->view, b + d.off_d_diff, b + d.off_d_same, b + d.off_s_tens,
                                                        b + d.off_exch_sqrt, b + d.off_diag_sqrt, b + d.off_exch_norms,
                                                        b + off);
        hb_square_kernel<<<(M * M + 127) / 128, 128, 0, c->stream>>>(d, b);
        c->launch_count += 3;
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(&d.s_norm, b + off, 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    *out = mol;
    return FRIES_OK;
}

extern "C" int fries_mol_destroy(fries_mol *mol) {
    if (mol) {
        cudaSetDevice(mol->ctx->device);
        delete mol;
    }
    return FRIES_OK;
}

extern "C" int fries_mol_hb_tables(fries_mol *mol, double *h_d_diff, double *h_d_same, double *h_s_tens,
                                   double *h_s_norm, double *h_exch_sqrt, double *h_diag_sqrt, double *h_exch_norms) {
    FRIES_REQUIRE(mol, "fries_mol_hb_tables: NULL handle");
    fries_ctx *c = mol->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    const MolDims &d = mol->view.d;
    unsigned M = d.n_orb, TT = M * (M - 1) / 2;
    const double *b = mol->blob.p;
    if (h_d_diff) CUDA_TRY(cudaMemcpyAsync(h_d_diff, b + d.off_d_diff, M * M * 8, cudaMemcpyDeviceToHost, c->stream));
    if (h_d_same) CUDA_TRY(cudaMemcpyAsync(h_d_same, b + d.off_d_same, TT * 8, cudaMemcpyDeviceToHost, c->stream));
    if (h_s_tens) CUDA_TRY(cudaMemcpyAsync(h_s_tens, b + d.off_s_tens, M * 8, cudaMemcpyDeviceToHost, c->stream));
    if (h_exch_sqrt)
        CUDA_TRY(cudaMemcpyAsync(h_exch_sqrt, b + d.off_exch_sqrt, TT * 8, cudaMemcpyDeviceToHost, c->stream));
    if (h_diag_sqrt)
        CUDA_TRY(cudaMemcpyAsync(h_diag_sqrt, b + d.off_diag_sqrt, M * 8, cudaMemcpyDeviceToHost, c->stream));
    if (h_exch_norms)
        CUDA_TRY(cudaMemcpyAsync(h_exch_norms, b + d.off_exch_norms, M * 8, cudaMemcpyDeviceToHost, c->stream));
    if (h_s_norm) *h_s_norm = d.s_norm;
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}

// ---------------------------------------------------------------------------------------------------
// batch parity kernels: one thread per item, tables read through L2 (these are not the hot path;
// the hot kernels stage the blob in shared memory, see hbpp.cu)
// ---------------------------------------------------------------------------------------------------
__global__ void mol_diag_kernel(MolView m, const uint64_t *keys, size_t n, double *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t occ[FRIES_MAX_ELEC + 1];
    mol_occ_list(keys[i], occ);
    out[i] = mol_diag(m, occ);
}
__global__ void mol_sing_el_kernel(MolView m, const uint64_t *keys, const uint8_t *orbs, size_t n, double *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t occ[FRIES_MAX_ELEC + 1];
    mol_occ_list(keys[i], occ);
    out[i] = mol_sing_el(m, orbs[2 * i], orbs[2 * i + 1], occ);
}
__global__ void mol_doub_el_kernel(MolView m, const uint8_t *orbs, size_t n, double *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = mol_doub_el(m, orbs + 4 * i);
}
// mode 0: count; mode 1: fill at offsets
__global__ void mol_ex_kernel(MolView m, int doubles, int fill, const uint64_t *keys, size_t n, uint64_t *offsets,
                              uint8_t *orbs, size_t cap) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t occ[FRIES_MAX_ELEC + 1];
    uint64_t det = keys[i];
    mol_occ_list(det, occ);
    if (!fill) {
        unsigned cnt;
        if (doubles)
            cnt = mol_for_each_doub(m, det, occ, [](unsigned, unsigned, unsigned, unsigned) {});
        else
            cnt = mol_for_each_sing(m, det, occ, [](unsigned, unsigned) {});
        offsets[i] = cnt;
    } else {
        size_t o = offsets[i];
        if (doubles) {
            mol_for_each_doub(m, det, occ, [&](unsigned a, unsigned b, unsigned k, unsigned l) {
                if (o < cap) {
                    orbs[4 * o] = (uint8_t)a;
                    orbs[4 * o + 1] = (uint8_t)b;
                    orbs[4 * o + 2] = (uint8_t)k;
                    orbs[4 * o + 3] = (uint8_t)l;
                }
                o++;
            });
        } else {
            mol_for_each_sing(m, det, occ, [&](unsigned a, unsigned v) {
                if (o < cap) {
                    orbs[2 * o] = (uint8_t)a;
                    orbs[2 * o + 1] = (uint8_t)v;
                }
                o++;
            });
        }
    }
}
__global__ void mol_hb_rows_kernel(MolView m, int which, const uint64_t *keys, const int32_t *args, size_t n,
                                   double *rows, int32_t *len, double *norm) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t occ[FRIES_MAX_ELEC + 1];
    uint64_t det = keys[i];
    mol_occ_list(det, occ);
    double p[FRIES_MAX_SUB];
    for (int j = 0; j < FRIES_MAX_SUB; j++) p[j] = 0;
    int a0 = args[4 * i], a1 = args[4 * i + 1], a2 = args[4 * i + 2];
    unsigned L = 0;
    double r = 0;
    switch (which) {
        case 0: r = hb_o1_probs(m, p, occ, a0); L = m.d.n_elec - (a0 > 0); break;
        case 1: r = hb_o2_probs(m, p, occ, a0); L = m.d.n_elec; break;
        case 2: r = hb_o2_probs_half(m, p, occ, a0); L = a0; break;
        case 3: r = hb_u1_probs(m, p, a0, occ, a1); L = m.d.n_orb - m.d.n_elec / 2; break;
        case 4: r = hb_u2_probs(m, p, a0, a1, a2, &L); break;
        case 5: r = hb_u2_probs_half(m, p, a0, a1, a2, det, &L); break;
    }
    for (int j = 0; j < FRIES_MAX_SUB; j++) rows[i * FRIES_MAX_SUB + j] = p[j];
    len[i] = (int32_t)L;
    norm[i] = r;
}
__global__ void mol_hb_wt_kernel(MolView m, int normalized, const uint64_t *keys, const uint8_t *orbs, size_t n,
                                 double *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t occ[FRIES_MAX_ELEC + 1];
    uint64_t det = keys[i];
    mol_occ_list(det, occ);
    out[i] = normalized ? hb_norm_wt(m, orbs + 4 * i, occ, det) : hb_unnorm_wt(m, orbs + 4 * i);
}

static int check_keys(const fries_mol *mol, const uint64_t *h_keys, size_t n, const char *who) {
    const MolDims &d = mol->view.d;
    uint64_t half = (1ull << d.n_orb) - 1;
    for (size_t i = 0; i < n; i++) {
        uint64_t k = h_keys[i];
        if ((k >> (2 * d.n_orb)) != 0 || (unsigned)__builtin_popcountll(k & half) != d.n_elec / 2 ||
            (unsigned)__builtin_popcountll(k >> d.n_orb) != d.n_elec / 2) {
            fries_set_error("%s: determinant %zu (%016llx) does not have %u alpha + %u beta electrons in %u orbitals", who,
                            i, (unsigned long long)k, d.n_elec / 2, d.n_elec / 2, d.n_orb);
            return FRIES_ERR_ARG;
        }
    }
    return FRIES_OK;
}

#define GRID1(n) (unsigned)(((n) + 127) / 128), 128, 0, c->stream

extern "C" int fries_mol_diag(fries_mol *mol, const uint64_t *h_keys, size_t n, double *h_out) {
    FRIES_REQUIRE(mol && (n == 0 || (h_keys && h_out)), "fries_mol_diag: NULL argument");
    if (n == 0) return FRIES_OK;
    FRIES_TRY(check_keys(mol, h_keys, n, "fries_mol_diag"));
    fries_ctx *c = mol->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    DevBuf<uint64_t> k;
    DevBuf<double> o;
    FRIES_TRY(k.alloc(n));
    FRIES_TRY(o.alloc(n));
    CUDA_TRY(cudaMemcpyAsync(k.p, h_keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    mol_diag_kernel<<<GRID1(n)>>>(mol->view, k.p, n, o.p);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_out, o.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}

extern "C" int fries_mol_sing_el(fries_mol *mol, const uint64_t *h_keys, const uint8_t *h_orbs, size_t n,
                                 double *h_out) {
    FRIES_REQUIRE(mol && (n == 0 || (h_keys && h_orbs && h_out)), "fries_mol_sing_el: NULL argument");
    if (n == 0) return FRIES_OK;
    FRIES_TRY(check_keys(mol, h_keys, n, "fries_mol_sing_el"));
    for (size_t i = 0; i < 2 * n; i++)
        FRIES_REQUIRE(h_orbs[i] < 2 * mol->view.d.n_orb, "fries_mol_sing_el: orbital out of range");
    fries_ctx *c = mol->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    DevBuf<uint64_t> k;
    DevBuf<uint8_t> ob;
    DevBuf<double> o;
    FRIES_TRY(k.alloc(n));
    FRIES_TRY(ob.alloc(2 * n));
    FRIES_TRY(o.alloc(n));
    CUDA_TRY(cudaMemcpyAsync(k.p, h_keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(ob.p, h_orbs, 2 * n, cudaMemcpyHostToDevice, c->stream));
    mol_sing_el_kernel<<<GRID1(n)>>>(mol->view, k.p, ob.p, n, o.p);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_out, o.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}

extern "C" int fries_mol_doub_el(fries_mol *mol, const uint8_t *h_orbs, size_t n, double *h_out) {
    FRIES_REQUIRE(mol && (n == 0 || (h_orbs && h_out)), "fries_mol_doub_el: NULL argument");
    if (n == 0) return FRIES_OK;
    for (size_t i = 0; i < 4 * n; i++)
        FRIES_REQUIRE(h_orbs[i] < 2 * mol->view.d.n_orb, "fries_mol_doub_el: orbital out of range");
    fries_ctx *c = mol->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    DevBuf<uint8_t> ob;
    DevBuf<double> o;
    FRIES_TRY(ob.alloc(4 * n));
    FRIES_TRY(o.alloc(n));
    CUDA_TRY(cudaMemcpyAsync(ob.p, h_orbs, 4 * n, cudaMemcpyHostToDevice, c->stream));
    mol_doub_el_kernel<<<GRID1(n)>>>(mol->view, ob.p, n, o.p);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_out, o.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}

static int mol_ex(fries_mol *mol, int doubles, const uint64_t *h_keys, size_t n, uint64_t *h_offsets, uint8_t *h_orbs,
                  size_t cap) {
    FRIES_REQUIRE(mol && h_offsets && (n == 0 || h_keys), "fries_mol_*_ex: NULL argument");
    h_offsets[0] = 0;
    if (n == 0) return FRIES_OK;
    FRIES_TRY(check_keys(mol, h_keys, n, "fries_mol_*_ex"));
    fries_ctx *c = mol->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    DevBuf<uint64_t> k, off;
    FRIES_TRY(k.alloc(n));
    FRIES_TRY(off.alloc(n + 1));
    CUDA_TRY(cudaMemcpyAsync(k.p, h_keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    mol_ex_kernel<<<GRID1(n)>>>(mol->view, doubles, 0, k.p, n, off.p, nullptr, 0);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    std::vector<uint64_t> cnt(n);
    CUDA_TRY(cudaMemcpyAsync(cnt.data(), off.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    uint64_t run = 0;
    for (size_t i = 0; i < n; i++) {
        h_offsets[i] = run;
        run += cnt[i];
    }
    h_offsets[n] = run;
    if (!h_orbs) return FRIES_OK;
    if (run > cap) {
        fries_set_error("fries_mol_*_ex: %llu excitations, buffer holds %zu", (unsigned long long)run, cap);
        return FRIES_ERR_CAPACITY;
    }
    if (run == 0) return FRIES_OK;
    int w = doubles ? 4 : 2;
    DevBuf<uint8_t> ob;
    FRIES_TRY(ob.alloc(run * w));
    CUDA_TRY(cudaMemcpyAsync(off.p, h_offsets, (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    mol_ex_kernel<<<GRID1(n)>>>(mol->view, doubles, 1, k.p, n, off.p, ob.p, run);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_orbs, ob.p, run * w, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}
extern "C" int fries_mol_sing_ex(fries_mol *mol, const uint64_t *h_keys, size_t n, uint64_t *h_offsets, uint8_t *h_orbs,
                                 size_t cap) {
    return mol_ex(mol, 0, h_keys, n, h_offsets, h_orbs, cap);
}
extern "C" int fries_mol_doub_ex(fries_mol *mol, const uint64_t *h_keys, size_t n, uint64_t *h_offsets, uint8_t *h_orbs,
                                 size_t cap) {
    return mol_ex(mol, 1, h_keys, n, h_offsets, h_orbs, cap);
}

extern "C" int fries_mol_hb_rows(fries_mol *mol, int which, const uint64_t *h_keys, const int32_t *h_args4, size_t n,
                                 double *h_rows, int32_t *h_len, double *h_norm) {
    FRIES_REQUIRE(mol && (n == 0 || (h_keys && h_args4 && h_rows && h_len && h_norm)), "fries_mol_hb_rows: NULL argument");
    FRIES_REQUIRE(which >= 0 && which <= 5, "fries_mol_hb_rows: unknown row kind %d", which);
    if (n == 0) return FRIES_OK;
    FRIES_TRY(check_keys(mol, h_keys, n, "fries_mol_hb_rows"));
    fries_ctx *c = mol->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    DevBuf<uint64_t> k;
    DevBuf<int32_t> a, l;
    DevBuf<double> r, nm;
    FRIES_TRY(k.alloc(n));
    FRIES_TRY(a.alloc(4 * n));
    FRIES_TRY(l.alloc(n));
    FRIES_TRY(r.alloc(n * FRIES_MAX_SUB));
    FRIES_TRY(nm.alloc(n));
    CUDA_TRY(cudaMemcpyAsync(k.p, h_keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(a.p, h_args4, n * 16, cudaMemcpyHostToDevice, c->stream));
    mol_hb_rows_kernel<<<GRID1(n)>>>(mol->view, which, k.p, a.p, n, r.p, l.p, nm.p);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_rows, r.p, n * FRIES_MAX_SUB * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(h_len, l.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(h_norm, nm.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}

extern "C" int fries_mol_hb_wt(fries_mol *mol, int normalized, const uint64_t *h_keys, const uint8_t *h_orbs, size_t n,
                               double *h_out) {
    FRIES_REQUIRE(mol && (n == 0 || (h_keys && h_orbs && h_out)), "fries_mol_hb_wt: NULL argument");
    if (n == 0) return FRIES_OK;
    FRIES_TRY(check_keys(mol, h_keys, n, "fries_mol_hb_wt"));
    fries_ctx *c = mol->ctx;
    CUDA_TRY(cudaSetDevice(c->device));
    DevBuf<uint64_t> k;
    DevBuf<uint8_t> ob;
    DevBuf<double> o;
    FRIES_TRY(k.alloc(n));
    FRIES_TRY(ob.alloc(4 * n));
    FRIES_TRY(o.alloc(n));
    CUDA_TRY(cudaMemcpyAsync(k.p, h_keys, n * 8, cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaMemcpyAsync(ob.p, h_orbs, 4 * n, cudaMemcpyHostToDevice, c->stream));
    mol_hb_wt_kernel<<<GRID1(n)>>>(mol->view, normalized, k.p, ob.p, n, o.p);
    c->launch_count++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h_out, o.p, n * 8, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return FRIES_OK;
}
