"""ctypes view of the plain-C oracle (oracle/fries_oracle.c -> oracle/_build/libfries_oracle.so).  TEST ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_build", "libfries_oracle.so")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C")
_lib = None


def build():
    src = [os.path.join(ROOT, "oracle", f) for f in ("fries_oracle.c", "fries_oracle.h")]
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in src):
        return
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        vp, sz, u, i, d = C.c_void_p, C.c_size_t, C.c_uint, C.c_int, C.c_double
        P = C.POINTER
        sig = {
            "fo_bits_between": (u, [C.c_uint64, i, i]),
            "fo_gen_hf_bitstring": (C.c_uint64, [u, u]),
            "fo_sing_det_parity": (i, [P(C.c_uint64), u8p]),
            "fo_doub_det_parity": (i, [P(C.c_uint64), u8p]),
            "fo_sing_parity": (i, [C.c_uint64, u8p]),
            "fo_doub_parity": (i, [C.c_uint64, u8p]),
            "fo_find_nth_virt": (i, [u8p, i, i, i, i]),
            "fo_hash": (C.c_uint64, [C.c_uint64, u32p]),
            "fo_hash_keys": (None, [u64p, sz, u32p, i, vp, vp]),
            "fo_find_preserve": (d, [f64p, sz, P(u), P(d), u8p]),
            "fo_sys_comp": (None, [f64p, sz, f64p, i, i, u, u8p, d]),
            "fo_comp_sub": (sz, [f64p, sz, u32p, f64p, sz, vp, u, d, f64p, u64p, P(u), P(d)]),
            "fo_adjust_shift": (None, [P(d), d, P(d), d, d]),
            "fo_mt19937_fill": (None, [C.c_uint32, sz, u32p]),
            "fo_piv_samp_serial": (None, [f64p, sz, d, C.c_uint32, u8p, u32p, P(sz)]),
            "fo_piv_budget": (None, [f64p, i, C.c_uint32, u32p, P(sz), u32p]),
            "fo_adjust_probs": (d, [f64p, sz, P(C.c_uint32), d, C.c_uint32, d, u8p]),
            "fo_piv_comp": (None, [f64p, sz, C.c_uint32, u8p, u32p, P(sz)]),
            "fo_mol_apply_hbpp_piv": (sz, [vp, u64p, f64p, sz, d, i, u32p, P(sz), u, sz, f64p, u64p, u8p]),
            "fo_mol_create": (vp, [u, u, u, f64p, f64p, u8p]),
            "fo_mol_destroy": (None, [vp]),
            "fo_mol_packed_len": (sz, [vp]),
            "fo_mol_packed_eris": (P(d), [vp]),
            "fo_mol_hb_tables": (None, [vp] + [vp] * 7),
            "fo_mol_diag": (d, [vp, C.c_uint64]),
            "fo_mol_sing_el": (d, [vp, C.c_uint64, u8p]),
            "fo_mol_doub_el": (d, [vp, u8p]),
            "fo_mol_sing_ex": (sz, [vp, C.c_uint64, vp]),
            "fo_mol_doub_ex": (sz, [vp, C.c_uint64, vp]),
            "fo_mol_count_singex": (sz, [vp, C.c_uint64]),
            "fo_mol_hb_row": (d, [vp, i, C.c_uint64, i, i, i, f64p, P(i)]),
            "fo_mol_hb_wt": (d, [vp, i, C.c_uint64, u8p]),
            "fo_mol_apply_hbpp_sys": (sz, [vp, u64p, f64p, sz, d, i, f64p, u, sz, f64p, u64p, u8p]),
            "fo_mol_h_apply_list": (sz, [vp, u64p, f64p, sz, d, d, u64p, f64p, sz]),
        }
        u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
        sig["fo_setup_alias"] = (None, [f64p, u32p, f64p, sz])
        sig["fo_sample_alias"] = (None, [u32p, f64p, sz, u16p, C.c_uint32, u32p])
        sig["fo_compress_multi_row"] = (sz, [f64p, sz, C.c_uint32, u32p])
        sig["fo_set_keep_chunk"] = (None, [sz])
        sig["fo_debug_hbpp_stage"] = (sz, [vp, u64p, f64p, sz, d, i, f64p, u, sz, i, f64p, u64p, u8p, u32p])
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


class OracleMol:
    def __init__(self, sm):
        self.sm = sm
        self.h = lib().fo_mol_create(sm.n_orb, sm.n_elec_total, sm.n_frz, sm.hcore.reshape(-1), sm.eris_chem.reshape(-1),
                                     sm.symm)

    def __del__(self):
        if getattr(self, "h", None):
            lib().fo_mol_destroy(self.h)
            self.h = None

    def packed_eris(self):
        n = lib().fo_mol_packed_len(self.h)
        return np.ctypeslib.as_array(lib().fo_mol_packed_eris(self.h), (n,)).copy()

    def hb_tables(self):
        M = self.sm.n_orb
        T = M * (M - 1) // 2
        out = dict(d_diff=np.zeros(M * M), d_same=np.zeros(T), s_tens=np.zeros(M), s_norm=np.zeros(1),
                   exch_sqrt=np.zeros(T), diag_sqrt=np.zeros(M), exch_norms=np.zeros(M))
        lib().fo_mol_hb_tables(self.h, *[out[k].ctypes.data for k in
                                         ("d_diff", "d_same", "s_tens", "s_norm", "exch_sqrt", "diag_sqrt", "exch_norms")])
        return out

    def diag(self, keys):
        return np.array([lib().fo_mol_diag(self.h, int(k)) for k in keys])

    def sing_el(self, keys, orbs):
        return np.array([lib().fo_mol_sing_el(self.h, int(k), np.ascontiguousarray(o, np.uint8)) for k, o in zip(keys, orbs)])

    def doub_el(self, orbs):
        return np.array([lib().fo_mol_doub_el(self.h, np.ascontiguousarray(o, np.uint8)) for o in orbs])

    def sing_ex(self, key):
        buf = np.zeros((4096, 2), np.uint8)
        n = lib().fo_mol_sing_ex(self.h, int(key), buf.ctypes.data)
        return buf[:n].copy()

    def doub_ex(self, key):
        buf = np.zeros((1 << 16, 4), np.uint8)
        n = lib().fo_mol_doub_ex(self.h, int(key), buf.ctypes.data)
        return buf[:n].copy()

    def hb_row(self, which, key, a0=0, a1=0, a2=0):
        row = np.zeros(64)
        ln = C.c_int(0)
        r = lib().fo_mol_hb_row(self.h, which, int(key), a0, a1, a2, row, C.byref(ln))
        return r, row[:ln.value].copy()

    def hb_wt(self, normalized, key, orbs):
        return lib().fo_mol_hb_wt(self.h, normalized, int(key), np.ascontiguousarray(orbs, np.uint8))

    def apply_hbpp_sys(self, keys, vals, p_doub, new_hb, uniforms5, n_samp, spawn_length):
        ov = np.zeros(spawn_length)
        od = np.zeros(spawn_length, np.uint64)
        oo = np.zeros((spawn_length, 4), np.uint8)
        n = lib().fo_mol_apply_hbpp_sys(self.h, np.ascontiguousarray(keys, np.uint64),
                                        np.ascontiguousarray(vals, np.float64), len(keys), p_doub, int(new_hb),
                                        np.ascontiguousarray(uniforms5, np.float64), n_samp, spawn_length, ov, od,
                                        oo.reshape(-1))
        return ov[:n].copy(), od[:n].copy(), oo[:n].copy()

    def debug_hbpp_stage(self, keys, vals, p_doub, new_hb, uniforms5, n_samp, spawn_length, stage):
        """the list that leaves the comp_sub call of `stage` (0..4): (values, parent index, sub-index)"""
        ov = np.zeros(spawn_length)
        od = np.zeros(spawn_length, np.uint64)
        oo = np.zeros((spawn_length, 4), np.uint8)
        os_ = np.zeros(spawn_length, np.uint32)
        n = lib().fo_debug_hbpp_stage(self.h, np.ascontiguousarray(keys, np.uint64), np.ascontiguousarray(vals, np.float64),
                                      len(keys), p_doub, int(new_hb), np.ascontiguousarray(uniforms5, np.float64), n_samp,
                                      spawn_length, int(stage), ov, od, oo.reshape(-1), os_)
        return ov[:n].copy(), od[:n].copy(), os_[:n].copy()

    def apply_hbpp_piv(self, keys, vals, p_doub, new_hb, draws, n_samp, spawn_length):
        """heat_bathPP.cpp:1014-1419 -> (values, det indices, orbitals, draws consumed)"""
        ov = np.zeros(spawn_length)
        od = np.zeros(spawn_length, np.uint64)
        oo = np.zeros((spawn_length, 4), np.uint8)
        used = C.c_size_t(0)
        n = lib().fo_mol_apply_hbpp_piv(self.h, np.ascontiguousarray(keys, np.uint64),
                                        np.ascontiguousarray(vals, np.float64), len(keys), p_doub, int(new_hb),
                                        np.ascontiguousarray(draws, np.uint32), C.byref(used), n_samp, spawn_length, ov, od,
                                        oo.reshape(-1))
        return ov[:n].copy(), od[:n].copy(), oo[:n].copy(), used.value

    def h_apply(self, keys, vals, id_fac, h_fac):
        """merged result of id_fac*v + h_fac*H*v as sorted (keys, vals)"""
        cap = 1 << 20
        while True:
            ok = np.zeros(cap, np.uint64)
            ov = np.zeros(cap)
            n = lib().fo_mol_h_apply_list(self.h, np.ascontiguousarray(keys, np.uint64),
                                          np.ascontiguousarray(vals, np.float64), len(keys), id_fac, h_fac, ok, ov, cap)
            if n <= cap:
                break
            cap = n
        uk, inv = np.unique(ok[:n], return_inverse=True)
        return uk, np.bincount(inv, weights=ov[:n], minlength=uk.size)


def find_preserve(values, n_samp):
    v = np.ascontiguousarray(values, np.float64)
    keep = np.zeros(len(v), np.uint8)
    ns, gn = C.c_uint(n_samp), C.c_double(0)
    loc = lib().fo_find_preserve(v, len(v), C.byref(ns), C.byref(gn), keep)
    return loc, gn.value, ns.value, keep


def sys_comp(values, loc_norms, n_samp, keep, rn, n_procs=1, rank=0):
    v = np.array(values, np.float64)
    k = np.array(keep, np.uint8)
    ln = np.array(np.atleast_1d(loc_norms), np.float64)
    lib().fo_sys_comp(v, len(v), ln, n_procs, rank, n_samp, k, rn)
    return v, k, ln


def comp_sub(values, n_div, sub_weights, sub_sizes, n_samp, rn, cap):
    v = np.ascontiguousarray(values, np.float64)
    nd = np.ascontiguousarray(n_div, np.uint32)
    sw = np.ascontiguousarray(sub_weights, np.float64)
    ss = None if sub_sizes is None else np.ascontiguousarray(sub_sizes, np.uint16)
    nv = np.zeros(cap)
    ni = np.zeros((cap, 2), np.uint64)
    left, loc = C.c_uint(0), C.c_double(0)
    n = lib().fo_comp_sub(v, len(v), nd, sw.reshape(-1), sw.shape[1], None if ss is None else ss.ctypes.data, n_samp, rn,
                          nv, ni.reshape(-1), C.byref(left), C.byref(loc))
    return nv[:n].copy(), ni[:n].copy(), left.value, loc.value


def hash_keys(keys, scrambler, n_procs):
    k = np.ascontiguousarray(keys, np.uint64)
    h = np.zeros(k.size, np.uint64)
    o = np.zeros(k.size, np.int32)
    lib().fo_hash_keys(k, k.size, np.ascontiguousarray(scrambler, np.uint32), n_procs, h.ctypes.data, o.ctypes.data)
    return h, o


class keep_chunk:
    """with oraclelib.keep_chunk(1): ... -- run the oracle's find_keep_sub with the given chunk size"""

    def __init__(self, chunk):
        self.chunk = chunk

    def __enter__(self):
        lib().fo_set_keep_chunk(self.chunk)

    def __exit__(self, *a):
        lib().fo_set_keep_chunk(8)


def mt19937(seed, n):
    """first n outputs of std::mt19937(seed)"""
    out = np.zeros(n, np.uint32)
    lib().fo_mt19937_fill(seed, n, out)
    return out


def setup_alias(probs):
    """setup_alias compress_utils.cpp:823-857 -> aliases, alias_probs"""
    p = np.ascontiguousarray(probs, np.float64)
    al, ap = np.zeros(p.size, np.uint32), np.zeros(p.size)
    lib().fo_setup_alias(p, al, ap, p.size)
    return al, ap


def sample_alias(aliases, alias_probs, n_samp, draws):
    """sample_alias (counts) compress_utils.cpp:882-897 with two draws per sample -> counts"""
    counts = np.zeros(len(aliases), np.uint16)
    lib().fo_sample_alias(np.ascontiguousarray(aliases, np.uint32), np.ascontiguousarray(alias_probs, np.float64), len(aliases),
                          counts, n_samp, np.ascontiguousarray(draws, np.uint32))
    return counts


def compress_multi_row(values, compress_size, draws):
    """one row of compress_vecs_multi vec_utils.cpp:73-127 (single rank) -> new values, draws consumed"""
    v = np.array(values, np.float64)
    used = lib().fo_compress_multi_row(v, v.size, compress_size, np.ascontiguousarray(draws, np.uint32))
    return v, used


def piv_samp_serial(values, seg_norm, n_samp, keep, draws):
    """-> values, keep (1 = zeroed), draws consumed"""
    v = np.array(values, np.float64)
    k = np.array(keep, np.uint8)
    used = C.c_size_t(0)
    lib().fo_piv_samp_serial(v, len(v), seg_norm, n_samp, k, np.ascontiguousarray(draws, np.uint32), C.byref(used))
    return v, k, used.value


def piv_budget(loc_norms, n_samp, draws):
    ln = np.ascontiguousarray(loc_norms, np.float64)
    b = np.zeros(len(ln), np.uint32)
    used = C.c_size_t(0)
    lib().fo_piv_budget(ln, len(ln), n_samp, np.ascontiguousarray(draws, np.uint32), C.byref(used), b)
    return b, used.value


def adjust_probs(values, n_loc, exp_loc, n_tot, tot_norm, keep):
    v = np.array(values, np.float64)
    k = np.array(keep, np.uint8)
    nl = C.c_uint32(n_loc)
    r = lib().fo_adjust_probs(v, len(v), C.byref(nl), exp_loc, n_tot, tot_norm, k)
    return v, k, nl.value, r


def piv_comp(values, compress_size, draws):
    v = np.array(values, np.float64)
    k = np.zeros(len(v), np.uint8)
    used = C.c_size_t(0)
    lib().fo_piv_comp(v, len(v), compress_size, k, np.ascontiguousarray(draws, np.uint32), C.byref(used))
    return v, k, used.value
