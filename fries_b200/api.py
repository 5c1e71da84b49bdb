"""Host-side mirror of the reference interfaces on top of the C-ABI (numpy host buffers in, numpy out).

Names follow the reference: DistVec -> Vec (FRIES/vec_utils.hpp), find_preserve / sys_comp / comp_sub
(FRIES/compress_utils.hpp), molecule + heat_bathPP routines -> Mol methods.  Every call goes to the CUDA
library; nothing is computed in Python."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._capi import (FrifullParams, FrisysHhParams, FrisysParams, IterStats, MAX_SUB, arr, check, lib, ptr)


class Context:
    def __init__(self, device: int = 0):
        h = C.c_void_p()
        check(lib.fries_ctx_create(device, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib.fries_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(lib.fries_ctx_sync(self.h))

    @property
    def launch_count(self) -> int:
        return int(lib.fries_ctx_launch_count(self.h))

    @property
    def sm_count(self) -> int:
        return int(lib.fries_ctx_sm_count(self.h))

    def set_profile(self, on: int):
        check(lib.fries_ctx_set_profile(self.h, on))

    def kernel_ms(self, name: str):
        ms, n = C.c_double(0), C.c_uint64(0)
        check(lib.fries_ctx_kernel_ms(self.h, name.encode(), C.byref(ms), C.byref(n)))
        return ms.value, n.value


# ---- a1 / a15 --------------------------------------------------------------------------------------------
def hash_owner(ctx, keys, scrambler, n_ranks=1):
    """HashTable::hash_fxn + DistVec::idx_to_proc (det_hash.hpp:160-170, vec_utils.hpp:360-379)."""
    keys = arr(keys, np.uint64)
    scr = arr(scrambler, np.uint32)
    h = np.zeros(keys.size, np.uint64)
    o = np.zeros(keys.size, np.int32)
    check(lib.fries_hash_owner(ctx.h, ptr(keys), keys.size, ptr(scr), scr.size, n_ranks, ptr(h), ptr(o)))
    return h, o


def bit_op(ctx, op, keys, orbs):
    keys = np.array(keys, np.uint64)
    orbs = arr(orbs, np.uint8)
    sign = np.zeros(keys.size, np.int32)
    check(lib.fries_bit_op(ctx.h, op, ptr(keys), ptr(orbs), keys.size, ptr(sign)))
    return keys, sign


# ---- a4 / a5 / a6 ------------------------------------------------------------------------------------------
def find_preserve(ctx, values, n_samp):
    """compress_utils.cpp:29-105 -> (loc_norm, glob_norm, n_samp_left, keep[0/1])"""
    v = arr(values, np.float64)
    keep = np.zeros(v.size, np.uint8)
    ns, gn, ln = C.c_uint(n_samp), C.c_double(0), C.c_double(0)
    check(lib.fries_find_preserve(ctx.h, ptr(v), v.size, C.byref(ns), C.byref(gn), ptr(keep), C.byref(ln)))
    return ln.value, gn.value, ns.value, keep


def sys_comp(ctx, values, loc_norms, n_samp, keep, rand_num, n_ranks=1, rank=0):
    """compress_utils.cpp:278-327 -> (values, delete flags, loc_norms)"""
    v = np.array(values, np.float64)
    k = np.array(keep, np.uint8)
    ln = np.array(np.atleast_1d(loc_norms), np.float64)
    check(lib.fries_sys_comp(ctx.h, ptr(v), v.size, ptr(ln), n_ranks, rank, n_samp, ptr(k), rand_num))
    return v, k, ln


def comp_sub(ctx, values, n_div, sub_weights, sub_sizes, n_samp, rand_num, out_cap):
    """compress_utils.cpp:797-820 -> (new_vals, new_idx[n][2], n_samp_left, loc_norm)"""
    v = arr(values, np.float64)
    nd = arr(n_div, np.uint32)
    sw = arr(sub_weights, np.float64)
    ss = None if sub_sizes is None else arr(sub_sizes, np.uint16)
    nv = np.zeros(out_cap)
    ni = np.zeros((out_cap, 2), np.uint64)
    n_out, left, ln = C.c_size_t(0), C.c_uint(0), C.c_double(0)
    check(lib.fries_comp_sub(ctx.h, ptr(v), v.size, ptr(nd), ptr(sw), sw.shape[1], ptr(ss), n_samp, rand_num, ptr(nv),
                             ptr(ni), out_cap, C.byref(n_out), C.byref(left), C.byref(ln)))
    return nv[:n_out.value].copy(), ni[:n_out.value].copy(), left.value, ln.value


# ---- pivotal family (compress_utils.cpp:354-681); draws = raw outputs of the caller's mt19937 ----------------------
def piv_samp_serial(ctx, values, seg_norm, n_samp, keep, draws):
    """compress_utils.cpp:390-530 -> (values, delete flags, draws consumed)"""
    v = np.array(values, np.float64)
    k = np.array(keep, np.uint8)
    dr = arr(draws, np.uint32)
    if dr.size < 2 * n_samp:
        raise ValueError("piv_samp_serial needs 2 draws per sample")
    used = C.c_size_t(0)
    check(lib.fries_piv_samp_serial(ctx.h, ptr(v), v.size, seg_norm, n_samp, ptr(k), ptr(dr), C.byref(used)))
    return v, k, used.value


def adjust_probs(ctx, values, n_samp_loc, exp_nsamp_loc, n_samp_tot, tot_norm, keep):
    """compress_utils.cpp:617-681 -> (values, keep, n_samp_loc, norm for the sampler)"""
    v = np.array(values, np.float64)
    k = np.array(keep, np.uint8)
    nl, nn = C.c_uint32(n_samp_loc), C.c_double(0)
    check(lib.fries_adjust_probs(ctx.h, ptr(v), v.size, C.byref(nl), exp_nsamp_loc, n_samp_tot, tot_norm, ptr(k),
                                 C.byref(nn)))
    return v, k, nl.value, nn.value


def piv_budget(loc_norms, n_samp, draws):
    """compress_utils.cpp:560-608 (host arithmetic, as on the reference's rank 0) -> (budgets of all ranks, draws used)"""
    ln = arr(loc_norms, np.float64)
    dr = arr(draws, np.uint32)
    if dr.size < 2 * ln.size:
        raise ValueError("piv_budget needs up to 2 draws per rank")
    b = np.zeros(ln.size, np.uint32)
    used = C.c_size_t(0)
    check(lib.fries_piv_budget(ptr(ln), ln.size, n_samp, ptr(dr), C.byref(used), ptr(b)))
    return b, used.value


def piv_comp(ctx, values, compress_size, draws, loc_norms=None, rank=0, keep=None, n_samp_left=0):
    """piv_comp_parallel compress_utils.cpp:354-387 -> (values, delete flags, draws consumed[, loc_norms]).
    Single rank: only values / compress_size / draws.  As one rank of several: loc_norms (all-gathered residual norms),
    keep and n_samp_left from the collective find_preserve."""
    v = np.array(values, np.float64)
    dr = arr(draws, np.uint32)
    n_ranks = 1 if loc_norms is None else len(loc_norms)
    if dr.size < 2 * (compress_size + n_ranks):
        raise ValueError("piv_comp needs 2 * (compress_size + n_ranks) draws")
    k = np.zeros(v.size, np.uint8) if keep is None else np.array(keep, np.uint8)
    ln = None if loc_norms is None else np.array(loc_norms, np.float64)
    used = C.c_size_t(0)
    check(lib.fries_piv_comp(ctx.h, ptr(v), v.size, compress_size, ptr(k), ptr(dr), C.byref(used), ptr(ln), n_ranks, rank,
                             0 if keep is None else 1, n_samp_left))
    return (v, k, used.value) if ln is None else (v, k, used.value, ln)


# ---- molecular Hamiltonian ----------------------------------------------------------------------------------
class Mol:
    """Integrals + SymmInfo + hb_info resident on the GPU (molecule.hpp, heat_bathPP.hpp)."""

    def __init__(self, ctx, n_orb, n_elec_total, n_frz, hcore, eris_packed, symm):
        self.ctx = ctx
        self.n_orb, self.n_elec_total, self.n_frz = n_orb, n_elec_total, n_frz
        self.n_elec = n_elec_total - n_frz
        hc, er, sy = arr(hcore, np.float64), arr(eris_packed, np.float64), arr(symm, np.uint8)
        h = C.c_void_p()
        check(lib.fries_mol_create(ctx.h, n_orb, n_elec_total, n_frz, ptr(hc), ptr(er), ptr(sy), C.byref(h)))
        self.h = h

    @classmethod
    def from_synth(cls, ctx, sm):
        return cls(ctx, sm.n_orb, sm.n_elec_total, sm.n_frz, sm.hcore, sm.eris_packed, sm.symm)

    def close(self):
        if self.h:
            lib.fries_mol_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def hb_tables(self):
        M = self.n_orb
        T = M * (M - 1) // 2
        out = dict(d_diff=np.zeros(M * M), d_same=np.zeros(T), s_tens=np.zeros(M), s_norm=np.zeros(1),
                   exch_sqrt=np.zeros(T), diag_sqrt=np.zeros(M), exch_norms=np.zeros(M))
        check(lib.fries_mol_hb_tables(self.h, *[ptr(out[k]) for k in ("d_diff", "d_same", "s_tens", "s_norm", "exch_sqrt",
                                                                       "diag_sqrt", "exch_norms")]))
        return out

    def diag(self, keys):
        k = arr(keys, np.uint64)
        out = np.zeros(k.size)
        check(lib.fries_mol_diag(self.h, ptr(k), k.size, ptr(out)))
        return out

    def sing_el(self, keys, orbs):
        k, o = arr(keys, np.uint64), arr(orbs, np.uint8)
        out = np.zeros(k.size)
        check(lib.fries_mol_sing_el(self.h, ptr(k), ptr(o), k.size, ptr(out)))
        return out

    def doub_el(self, orbs):
        o = arr(orbs, np.uint8)
        out = np.zeros(len(o))
        check(lib.fries_mol_doub_el(self.h, ptr(o), len(o), ptr(out)))
        return out

    def _ex(self, fn, width, keys):
        k = arr(keys, np.uint64)
        off = np.zeros(k.size + 1, np.uint64)
        check(fn(self.h, ptr(k), k.size, ptr(off), None, 0))
        tot = int(off[-1])
        orbs = np.zeros((max(tot, 1), width), np.uint8)
        check(fn(self.h, ptr(k), k.size, ptr(off), ptr(orbs), tot))
        return off, orbs[:tot]

    def sing_ex(self, keys):
        return self._ex(lib.fries_mol_sing_ex, 2, keys)

    def doub_ex(self, keys):
        return self._ex(lib.fries_mol_doub_ex, 4, keys)

    def hb_rows(self, which, keys, args4):
        k, a = arr(keys, np.uint64), arr(args4, np.int32)
        rows = np.zeros((k.size, MAX_SUB))
        ln = np.zeros(k.size, np.int32)
        nm = np.zeros(k.size)
        check(lib.fries_mol_hb_rows(self.h, which, ptr(k), ptr(a), k.size, ptr(rows), ptr(ln), ptr(nm)))
        return rows, ln, nm

    def hb_wt(self, normalized, keys, orbs):
        k, o = arr(keys, np.uint64), arr(orbs, np.uint8)
        out = np.zeros(k.size)
        check(lib.fries_mol_hb_wt(self.h, int(normalized), ptr(k), ptr(o), k.size, ptr(out)))
        return out

    def apply_hbpp_piv(self, keys, vals, p_doub, new_hb, draws, n_samp, spawn_cap):
        """heat_bathPP.cpp:1014-1419 -> (values, det indices, orbitals[n][4], draws consumed)"""
        k, v, dr = arr(keys, np.uint64), arr(vals, np.float64), arr(draws, np.uint32)
        ov = np.zeros(spawn_cap)
        od = np.zeros(spawn_cap, np.uint64)
        oo = np.zeros((spawn_cap, 4), np.uint8)
        n_out, used = C.c_size_t(0), C.c_size_t(0)
        check(lib.fries_apply_hbpp_piv(self.h, ptr(k), ptr(v), k.size, p_doub, int(new_hb), ptr(dr), dr.size, C.byref(used),
                                       n_samp, spawn_cap, ptr(ov), ptr(od), ptr(oo), spawn_cap, C.byref(n_out)))
        n = n_out.value
        return ov[:n].copy(), od[:n].copy(), oo[:n].copy(), used.value

    def debug_hbpp_stage(self, keys, vals, p_doub, new_hb, uniforms5, n_samp, spawn_cap, stage):
        """diagnostics: the list that leaves the comp_sub call of `stage` (0..4) of apply_HBPP_sys ->
        (values, parent index, 4-byte path of the parent item, chosen sub-index)"""
        k, v, u = arr(keys, np.uint64), arr(vals, np.float64), arr(uniforms5, np.float64)
        ov = np.zeros(spawn_cap)
        od = np.zeros(spawn_cap, np.uint32)
        op = np.zeros(spawn_cap, np.uint32)
        os_ = np.zeros(spawn_cap, np.uint32)
        n = C.c_size_t(0)
        check(lib.fries_debug_hbpp_stage(self.h, ptr(k), ptr(v), k.size, p_doub, int(new_hb), ptr(u), n_samp, spawn_cap,
                                         int(stage), ptr(ov), ptr(od), ptr(op), ptr(os_), C.byref(n)))
        m = n.value
        return ov[:m].copy(), od[:m].copy(), op[:m].copy(), os_[:m].copy()

    def apply_hbpp_sys(self, keys, vals, p_doub, new_hb, uniforms5, n_samp, spawn_cap):
        """apply_HBPP_sys heat_bathPP.cpp:686-992 -> (vals, parent index, orbs[n][4])"""
        k, v, u = arr(keys, np.uint64), arr(vals, np.float64), arr(uniforms5, np.float64)
        ov = np.zeros(spawn_cap)
        od = np.zeros(spawn_cap, np.uint64)
        oo = np.zeros((spawn_cap, 4), np.uint8)
        n = C.c_size_t(0)
        check(lib.fries_apply_hbpp_sys(self.h, ptr(k), ptr(v), k.size, p_doub, int(new_hb), ptr(u), n_samp, spawn_cap,
                                       ptr(ov), ptr(od), ptr(oo), spawn_cap, C.byref(n)))
        return ov[:n.value].copy(), od[:n.value].copy(), oo[:n.value].copy()


# ---- determinant store ----------------------------------------------------------------------------------------
class Vec:
    """DistVec<double> (FRIES/vec_utils.hpp:121-953) resident on one GPU."""

    def __init__(self, ctx, capacity, n_bits, n_elec, n_vecs, proc_scrambler, vec_scrambler, n_ranks=1, rank=0, hh=None):
        """hh = (n_sites, ph_bits) makes a HubHolVec (FRIES/hh_vec.hpp); n_bits is then 2 n_sites + n_sites ph_bits"""
        self.ctx, self.capacity, self.n_bits, self.n_elec, self.n_vecs = ctx, capacity, n_bits, n_elec, n_vecs
        ps, vs = arr(proc_scrambler, np.uint32), arr(vec_scrambler, np.uint32)
        h = C.c_void_p()
        if hh is None:
            check(lib.fries_vec_create(ctx.h, capacity, n_bits, n_elec, n_vecs, ptr(ps), ptr(vs), n_ranks, rank, C.byref(h)))
        else:
            check(lib.fries_vec_create_hh(ctx.h, capacity, hh[0], hh[1], n_elec, n_vecs, ptr(ps), ptr(vs), n_ranks, rank,
                                          C.byref(h)))
        self.h = h
        self.hb = None

    def close(self):
        if getattr(self, "hb", None):
            lib.fries_hbpp_destroy(self.hb)
            self.hb = None
        if self.h:
            lib.fries_vec_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add(self, keys, vals, ini, origin=0, dest=0):
        """DistVec::add x n + perform_add(origin) with curr_vec_idx = dest"""
        k, v, f = arr(keys, np.uint64), arr(vals, np.float64), arr(ini, np.uint8)
        check(lib.fries_vec_add(self.h, ptr(k), ptr(v), ptr(f), k.size, origin, dest))

    def curr_size(self) -> int:
        n = C.c_size_t(0)
        check(lib.fries_vec_curr_size(self.h, C.byref(n)))
        return n.value

    def nonini_occ_add(self) -> int:
        n = C.c_uint64(0)
        check(lib.fries_vec_nonini_occ_add(self.h, C.byref(n)))
        return n.value

    def download(self):
        n = self.curr_size()
        keys = np.zeros(max(n, 1), np.uint64)
        vals = np.zeros((self.n_vecs, max(n, 1)))
        m = C.c_size_t(0)
        check(lib.fries_vec_download(self.h, ptr(keys), ptr(vals), max(n, 1), C.byref(m)))
        return keys[:n], vals[:, :n]

    def upload(self, keys, vals):
        """DistVec::load from host arrays: vals is [n_vecs][n]"""
        k, v = arr(keys, np.uint64), arr(vals, np.float64)
        check(lib.fries_vec_upload(self.h, ptr(k), ptr(v), k.size))

    def download_into(self, keys, vals) -> int:
        """download into caller-owned (e.g. pinned) buffers; returns the element count"""
        m = C.c_size_t(0)
        check(lib.fries_vec_download(self.h, ptr(keys), ptr(vals), keys.size, C.byref(m)))
        return m.value

    def delete(self, flags):
        f = arr(flags, np.uint8)
        check(lib.fries_vec_del(self.h, ptr(f), f.size))

    def set_deterministic(self, on=True):
        """reproducible merges: append and add in batch order (csrc/vec_det.cu); single rank"""
        check(lib.fries_vec_set_deterministic(self.h, 1 if on else 0))

    def compress(self, start_row, end_row, compress_size, draws, method="piv") -> int:
        """compress_vecs / compress_vecs_sys / compress_vecs_multi vec_utils.cpp:10-127 -> draws consumed"""
        dr = arr(draws, np.uint32)
        used = C.c_size_t(0)
        check(lib.fries_vec_compress(self.h, start_row, end_row, compress_size, {"piv": 0, "sys": 1, "multi": 2}[method], ptr(dr),
                                     dr.size, C.byref(used)))
        return used.value

    def dot(self, keys, vals, row=0) -> float:
        k, v = arr(keys, np.uint64), arr(vals, np.float64)
        out = C.c_double(0)
        check(lib.fries_vec_dot(self.h, ptr(k), ptr(v), k.size, row, C.byref(out)))
        return out.value

    def local_norm(self, row=0) -> float:
        out = C.c_double(0)
        check(lib.fries_vec_local_norm(self.h, row, C.byref(out)))
        return out.value

    def two_norm(self, row=0) -> float:
        out = C.c_double(0)
        check(lib.fries_vec_two_norm(self.h, row, C.byref(out)))
        return out.value

    def add_vecs(self, dst, src, c=1.0):
        check(lib.fries_vec_row_op(self.h, 0, dst, src, c))

    def copy_vec(self, src, dst):
        check(lib.fries_vec_row_op(self.h, 1, dst, src, 0.0))

    def weight_vec(self, dst, src, expo):
        check(lib.fries_vec_row_op(self.h, 2, dst, src, expo))

    def zero_vec(self, row):
        check(lib.fries_vec_row_op(self.h, 3, row, row, 0.0))

    def set_diag_mol(self, mol, hf_en):
        check(lib.fries_vec_set_diag_mol(self.h, mol.h, hf_en))

    def h_apply(self, mol, src, dest, id_fac, h_fac) -> int:
        """h_op_diag + h_op_offdiag (molecule.cpp:205-219, 448-665); returns the number of spawned elements"""
        check(lib.fries_h_apply(self.h, mol.h, src, dest, id_fac, h_fac))
        n = C.c_uint64(0)
        check(lib.fries_h_apply_last_spawned(self.h, C.byref(n)))
        return n.value

    # ---- drivers' loop bodies ----
    def frisys_setup(self, mol, spawn_cap, trial_keys, trial_vals, htrial_keys, htrial_vals):
        tk, tv = arr(trial_keys, np.uint64), arr(trial_vals, np.float64)
        hk, hv = arr(htrial_keys, np.uint64), arr(htrial_vals, np.float64)
        h = C.c_void_p()
        check(lib.fries_frisys_mol_setup(self.h, mol.h, spawn_cap, ptr(tk), ptr(tv), tk.size, ptr(hk), ptr(hv), hk.size,
                                         C.byref(h)))
        self.hb = h
        self.mol = mol

    def frisys_iterate(self, params: FrisysParams, uniforms6) -> IterStats:
        u = arr(uniforms6, np.float64)
        st = IterStats()
        check(lib.fries_frisys_mol_iterate(self.h, self.mol.h, self.hb, C.byref(params), ptr(u), C.byref(st)))
        return st

    def frisys_hh_setup(self, spawn_cap):
        h = C.c_void_p()
        check(lib.fries_frisys_hh_setup(self.h, spawn_cap, C.byref(h)))
        self.hb = h

    def frisys_hh_iterate(self, params: FrisysHhParams, uniforms3) -> IterStats:
        u = arr(uniforms3, np.float64)
        st = IterStats()
        check(lib.fries_frisys_hh_iterate(self.h, self.hb, C.byref(params), ptr(u), C.byref(st)))
        return st

    def frifull_hh_iterate(self, params: FrisysHhParams, uniform: float) -> IterStats:
        """loop body of FRIES_bin/frifull_hh.cpp:186-330"""
        st = IterStats()
        check(lib.fries_frifull_hh_iterate(self.h, self.hb, C.byref(params), float(uniform), C.byref(st)))
        return st

    def states(self):
        """CompState records of the last iteration [8 states][20]: loc_norm, glob_norm, new_norm, n_samp_left,
        rounds, n_kept, n_out, n_in, anomalies, n_cand, fast, overflow, 8 phase time stamps (ns)"""
        out = np.zeros((8, 20))
        check(lib.fries_hbpp_states(self.hb, ptr(out)))
        return out

    def debug_vec_phase(self, target_nonz, uniform):
        """diagnostics / parity: find_preserve -> sys_comp -> deletion + compaction of the fused vector kernel
        (csrc/vecphase.cu) on the stored vector -> (loc_norm, glob_norm, n_samp_left, n_kept); needs frisys_setup"""
        out = np.zeros(4)
        check(lib.fries_debug_vec_phase(self.h, self.mol.h, self.hb, int(target_nonz), float(uniform), ptr(out)))
        return float(out[0]), float(out[1]), int(out[2]), int(out[3])

    def timeline(self, s):
        """clock64 timeline (SM cycles from mark 0) of thread 0 of CTA 0 through the last compression of state s"""
        out = np.zeros(48)
        check(lib.fries_hbpp_timeline(self.hb, int(s), ptr(out)))
        return out

    def cta_marks(self, s):
        """per-CTA phase-end times [8][grid] (ns) of stage s (needs FRIES_CTA_MARKS=1)"""
        out = np.zeros((8, 1024))
        g = C.c_int(0)
        check(lib.fries_hbpp_cta_marks(self.hb, int(s), ptr(out), C.byref(g)))
        return out.reshape(-1)[:8 * g.value].reshape(8, g.value)

    def frifull_iterate(self, params: FrifullParams, uniform: float) -> IterStats:
        st = IterStats()
        check(lib.fries_frifull_mol_iterate(self.h, self.mol.h, self.hb, C.byref(params), uniform, C.byref(st)))
        return st


def hh_batch(ctx, what, keys, vals, n_sites, n_elec, ph_bits, ref_key=0, g_over_t=0.0):
    """what 0: hub_diag; 1: neighbour masks [n][2]; 2: calc_ref_ovlp terms (hub_holstein.cpp/.hpp, hh_vec.hpp)"""
    k = arr(keys, np.uint64)
    v = None if vals is None else arr(vals, np.float64)
    out = np.zeros(k.size * (2 if what == 1 else 1))
    check(lib.fries_hh_batch(ctx.h, what, ptr(k), ptr(v), k.size, n_sites, n_elec, ph_bits, int(ref_key), g_over_t, ptr(out)))
    return out.reshape(-1, 2) if what == 1 else out
