"""ctypes view of oracle/_ref/libfries_ref.so -- the REFERENCE's own code compiled from /root/reference by
oracle/Makefile (`make ref`) plus the marshalling shim oracle/ref_capi.cpp.  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libfries_ref.so")

u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C")

_lib = None


def available() -> bool:
    return os.path.exists(REF_SO)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(REF_SO)
        L.ref_gen_hf_bitstring.restype = C.c_uint64
        L.ref_gen_hf_bitstring.argtypes = [C.c_uint, C.c_uint]
        L.ref_bits_between.restype = C.c_uint
        L.ref_bits_between.argtypes = [C.c_uint64, C.c_int, C.c_int]
        for nm in ("ref_sing_det_parity", "ref_doub_det_parity"):
            getattr(L, nm).restype = C.c_int
            getattr(L, nm).argtypes = [C.POINTER(C.c_uint64), u8p]
        for nm in ("ref_sing_parity", "ref_doub_parity"):
            getattr(L, nm).restype = C.c_int
            getattr(L, nm).argtypes = [C.c_uint64, u8p]
        L.ref_find_nth_virt.restype = C.c_int
        L.ref_find_nth_virt.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.ref_hash_keys.restype = None
        L.ref_hash_keys.argtypes = [u64p, C.c_size_t, C.c_int, u32p, C.c_int, u64p, i32p]
        L.ref_find_preserve.restype = C.c_double
        L.ref_find_preserve.argtypes = [f64p, C.c_size_t, C.POINTER(C.c_uint), C.POINTER(C.c_double), u8p]
        L.ref_sys_comp.restype = None
        L.ref_sys_comp.argtypes = [f64p, C.c_size_t, C.POINTER(C.c_double), C.c_uint, u8p, C.c_double]
        L.ref_comp_sub.restype = C.c_size_t
        L.ref_comp_sub.argtypes = [f64p, C.c_size_t, u32p, f64p, C.c_size_t, C.c_void_p, C.c_uint, C.c_double, f64p,
                                   u64p, C.c_size_t]
        L.ref_find_keep_sub.restype = C.c_double
        L.ref_find_keep_sub.argtypes = [f64p, C.c_size_t, u32p, f64p, C.c_size_t, C.c_void_p, C.POINTER(C.c_uint), f64p,
                                        u8p]
        L.ref_mpi_init.restype = C.c_int
        L.ref_mpi_rank.restype = C.c_int
        L.ref_allgather_doubles.restype = None
        L.ref_allgather_doubles.argtypes = [f64p, C.c_int]
        L.ref_mt19937_fill.restype = None
        L.ref_mt19937_fill.argtypes = [C.c_uint32, C.c_size_t, u32p]
        L.ref_piv_samp_serial.restype = C.c_size_t
        L.ref_piv_samp_serial.argtypes = [f64p, C.c_size_t, C.c_double, C.c_uint32, u8p, C.c_uint32, C.c_uint64]
        L.ref_piv_budget.restype = C.c_size_t
        L.ref_piv_budget.argtypes = [f64p, C.c_int, C.c_uint32, C.c_uint32, u32p]
        L.ref_adjust_probs.restype = C.c_double
        L.ref_adjust_probs.argtypes = [f64p, C.c_size_t, C.POINTER(C.c_uint32), C.c_double, C.c_uint32, C.c_double, u8p]
        L.ref_piv_comp_parallel.restype = C.c_size_t
        L.ref_piv_comp_parallel.argtypes = [f64p, C.c_size_t, C.c_uint32, u8p, C.c_uint32]
        L.ref_adjust_shift.restype = None
        L.ref_adjust_shift.argtypes = [C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double), C.c_double, C.c_double]
        L.ref_mol_create.restype = C.c_void_p
        L.ref_mol_create.argtypes = [C.c_uint, C.c_uint, C.c_uint, f64p, f64p, u8p]
        L.ref_mol_destroy.argtypes = [C.c_void_p]
        L.ref_mol_hb_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.ref_mol_diag.argtypes = [C.c_void_p, u64p, C.c_size_t, f64p]
        L.ref_mol_sing_el.argtypes = [C.c_void_p, u64p, u8p, C.c_size_t, f64p]
        L.ref_mol_doub_el.argtypes = [C.c_void_p, u8p, C.c_size_t, f64p]
        L.ref_mol_sing_ex.restype = C.c_size_t
        L.ref_mol_sing_ex.argtypes = [C.c_void_p, C.c_uint64, u8p, C.c_size_t]
        L.ref_mol_doub_ex.restype = C.c_size_t
        L.ref_mol_doub_ex.argtypes = [C.c_void_p, C.c_uint64, u8p, C.c_size_t]
        L.ref_mol_count_singex.restype = C.c_size_t
        L.ref_mol_count_singex.argtypes = [C.c_void_p, C.c_uint64]
        L.ref_mol_hb_row.restype = C.c_double
        L.ref_mol_hb_row.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, f64p,
                                     C.POINTER(C.c_int)]
        L.ref_mol_hb_wt.restype = C.c_double
        L.ref_mol_hb_wt.argtypes = [C.c_void_p, C.c_int, C.c_uint64, u8p]
        L.ref_mol_apply_hbpp_piv.restype = C.c_size_t
        L.ref_mol_apply_hbpp_piv.argtypes = [C.c_void_p, u64p, f64p, C.c_size_t, C.c_double, C.c_int, C.c_uint, C.c_uint,
                                             C.c_size_t, f64p, u64p, u8p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ref_mol_apply_hbpp_sys.restype = C.c_size_t
        L.ref_mol_apply_hbpp_sys.argtypes = [C.c_void_p, u64p, f64p, C.c_size_t, C.c_double, C.c_int, C.c_uint, C.c_uint,
                                             C.c_size_t, f64p, f64p, u64p, u8p, C.c_size_t]
        L.ref_vec_create.restype = C.c_void_p
        L.ref_vec_create.argtypes = [C.c_size_t, C.c_size_t, C.c_uint, C.c_uint, C.c_uint, u32p, u32p]
        L.ref_vec_destroy.argtypes = [C.c_void_p]
        L.ref_vec_add.argtypes = [C.c_void_p, u64p, f64p, u8p, C.c_size_t, C.c_uint, C.c_uint]
        L.ref_vec_curr_size.restype = C.c_size_t
        L.ref_vec_curr_size.argtypes = [C.c_void_p]
        L.ref_vec_nonini_occ_add.restype = C.c_uint64
        L.ref_vec_nonini_occ_add.argtypes = [C.c_void_p]
        L.ref_vec_dump.argtypes = [C.c_void_p, u64p, f64p, C.c_uint]
        L.ref_vec_del.argtypes = [C.c_void_p, u8p]
        L.ref_vec_compress_multi.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint32]
        L.ref_setup_alias.argtypes = [f64p, u32p, f64p, C.c_size_t]
        L.ref_sample_alias.argtypes = [u32p, f64p, C.c_size_t, np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS"), C.c_uint32,
                                       C.c_uint32]
        L.ref_mol_h_apply.restype = C.c_size_t
        L.ref_mol_h_apply.argtypes = [C.c_void_p, u64p, f64p, C.c_size_t, C.c_double, C.c_double, C.c_size_t, u32p, u32p,
                                      u64p, f64p, C.c_size_t]
        _lib = L
    return _lib


class RefMol:
    """Reference molecular Hamiltonian (SymmERIs + hcore + SymmInfo + hb_info) on dense chemist eris."""

    def __init__(self, sm):
        self.sm = sm
        self.h = lib().ref_mol_create(sm.n_orb, sm.n_elec_total, sm.n_frz, sm.hcore, sm.eris_chem, sm.symm)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_mol_destroy(self.h)
            self.h = None

    def hb_tables(self):
        M = self.sm.n_orb
        T = M * (M - 1) // 2
        out = dict(d_diff=np.zeros(M * M), d_same=np.zeros(T), s_tens=np.zeros(M), s_norm=np.zeros(1),
                   exch_sqrt=np.zeros(T), diag_sqrt=np.zeros(M), exch_norms=np.zeros(M))
        lib().ref_mol_hb_tables(self.h, *[out[k].ctypes.data for k in
                                          ("d_diff", "d_same", "s_tens", "s_norm", "exch_sqrt", "diag_sqrt", "exch_norms")])
        return out

    def diag(self, keys):
        out = np.zeros(len(keys))
        lib().ref_mol_diag(self.h, np.ascontiguousarray(keys, np.uint64), len(keys), out)
        return out

    def sing_el(self, keys, orbs):
        out = np.zeros(len(keys))
        lib().ref_mol_sing_el(self.h, np.ascontiguousarray(keys, np.uint64), np.ascontiguousarray(orbs, np.uint8),
                              len(keys), out)
        return out

    def doub_el(self, orbs):
        out = np.zeros(len(orbs))
        lib().ref_mol_doub_el(self.h, np.ascontiguousarray(orbs, np.uint8), len(orbs), out)
        return out

    def sing_ex(self, key):
        buf = np.zeros((4096, 2), np.uint8)
        n = lib().ref_mol_sing_ex(self.h, int(key), buf.reshape(-1), 4096)
        return buf[:n].copy()

    def doub_ex(self, key):
        cap = 1 << 16
        buf = np.zeros((cap, 4), np.uint8)
        n = lib().ref_mol_doub_ex(self.h, int(key), buf.reshape(-1), cap)
        return buf[:n].copy()

    def hb_row(self, which, key, a0=0, a1=0, a2=0):
        row = np.zeros(64)
        ln = C.c_int(0)
        r = lib().ref_mol_hb_row(self.h, which, int(key), a0, a1, a2, 0, row, C.byref(ln))
        return r, row[:ln.value].copy()

    def hb_wt(self, normalized, key, orbs):
        return lib().ref_mol_hb_wt(self.h, normalized, int(key), np.ascontiguousarray(orbs, np.uint8))

    def apply_hbpp_sys(self, keys, vals, p_doub, new_hb, seed, n_samp, spawn_length):
        cap = spawn_length
        uni = np.zeros(5)
        ov = np.zeros(cap)
        od = np.zeros(cap, np.uint64)
        oo = np.zeros((cap, 4), np.uint8)
        n = lib().ref_mol_apply_hbpp_sys(self.h, np.ascontiguousarray(keys, np.uint64),
                                         np.ascontiguousarray(vals, np.float64), len(keys), p_doub, int(new_hb), seed,
                                         n_samp, spawn_length, uni, ov, od, oo.reshape(-1), cap)
        return uni, ov[:n].copy(), od[:n].copy(), oo[:n].copy()

    def apply_hbpp_piv(self, keys, vals, p_doub, new_hb, seed, n_samp, spawn_length):
        """apply_HBPP_piv with mt19937(seed) -> (values, det indices, orbitals, draws consumed)"""
        cap = spawn_length
        ov = np.zeros(cap)
        od = np.zeros(cap, np.uint64)
        oo = np.zeros((cap, 4), np.uint8)
        used = C.c_size_t(0)
        n = lib().ref_mol_apply_hbpp_piv(self.h, np.ascontiguousarray(keys, np.uint64),
                                         np.ascontiguousarray(vals, np.float64), len(keys), p_doub, int(new_hb), seed,
                                         n_samp, spawn_length, ov, od, oo.reshape(-1), cap, C.byref(used))
        return ov[:n].copy(), od[:n].copy(), oo[:n].copy(), used.value

    def h_apply(self, keys, vals, id_fac, h_fac, max_dets, proc_scr, vec_scr):
        ok = np.zeros(max_dets, np.uint64)
        ov = np.zeros(max_dets)
        n = lib().ref_mol_h_apply(self.h, np.ascontiguousarray(keys, np.uint64), np.ascontiguousarray(vals, np.float64),
                                  len(keys), id_fac, h_fac, max_dets, proc_scr, vec_scr, ok, ov, max_dets)
        return ok[:n].copy(), ov[:n].copy()


def find_preserve(values, n_samp):
    v = np.ascontiguousarray(values, np.float64)
    keep = np.zeros(len(v), np.uint8)
    ns = C.c_uint(n_samp)
    gn = C.c_double(0)
    loc = lib().ref_find_preserve(v, len(v), C.byref(ns), C.byref(gn), keep)
    return loc, gn.value, ns.value, keep


def sys_comp(values, loc_norm, n_samp, keep, rn):
    v = np.array(values, np.float64)
    k = np.array(keep, np.uint8)
    ln = C.c_double(loc_norm)
    lib().ref_sys_comp(v, len(v), C.byref(ln), n_samp, k, rn)
    return v, k, ln.value


def comp_sub(values, n_div, sub_weights, sub_sizes, n_samp, rn, cap):
    v = np.ascontiguousarray(values, np.float64)
    nd = np.ascontiguousarray(n_div, np.uint32)
    sw = np.ascontiguousarray(sub_weights, np.float64)
    ss = None if sub_sizes is None else np.ascontiguousarray(sub_sizes, np.uint16)
    nv = np.zeros(cap)
    ni = np.zeros((cap, 2), np.uint64)
    n = lib().ref_comp_sub(v, len(v), nd, sw.reshape(-1), sw.shape[1], None if ss is None else ss.ctypes.data, n_samp,
                           rn, nv, ni.reshape(-1), cap)
    return nv[:n].copy(), ni[:n].copy()


def mt19937(seed, n):
    """first n outputs of std::mt19937(seed)"""
    out = np.zeros(n, np.uint32)
    lib().ref_mt19937_fill(seed, n, out)
    return out


def piv_samp_serial(values, seg_norm, n_samp, keep, seed, skip=0):
    """-> values, keep (1 = zeroed), draws consumed"""
    v = np.array(values, np.float64)
    k = np.array(keep, np.uint8)
    used = lib().ref_piv_samp_serial(v, len(v), seg_norm, n_samp, k, seed, skip)
    return v, k, used


def piv_budget(loc_norms, n_samp, seed):
    ln = np.ascontiguousarray(loc_norms, np.float64)
    b = np.zeros(len(ln), np.uint32)
    used = lib().ref_piv_budget(ln, len(ln), n_samp, seed, b)
    return b, used


def adjust_probs(values, n_loc, exp_loc, n_tot, tot_norm, keep):
    v = np.array(values, np.float64)
    k = np.array(keep, np.uint8)
    nl = C.c_uint32(n_loc)
    r = lib().ref_adjust_probs(v, len(v), C.byref(nl), exp_loc, n_tot, tot_norm, k)
    return v, k, nl.value, r


def piv_comp_parallel(values, compress_size, seed):
    v = np.array(values, np.float64)
    k = np.zeros(len(v), np.uint8)
    used = lib().ref_piv_comp_parallel(v, len(v), compress_size, k, seed)
    return v, k, used
