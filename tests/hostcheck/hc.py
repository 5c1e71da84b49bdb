"""Host build of the product's __host__ __device__ arithmetic (tests/hostcheck/hostcheck.cu).  TEST ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libhostcheck.so")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C")
_lib = None


def build():
    src = os.path.join(HERE, "hostcheck.cu")
    deps = [src] + [os.path.join(HERE, "..", "..", "fries_b200", "csrc", f) for f in ("mol.cuh", "common.cuh", "piv.cuh", "hbpp_prov.cuh", "hh_prov.cuh", "hv_prov.cuh")]
    if os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(d) for d in deps):
        return
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                           "--extended-lambda", "-Wno-deprecated-gpu-targets", "-o", SO, src])


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(SO)
        L.hc_mol_create.restype = C.c_void_p
        L.hc_mol_create.argtypes = [C.c_uint, C.c_uint, C.c_uint, f64p, f64p, C.c_size_t, u8p, f64p, f64p, f64p,
                                    C.c_double, f64p, f64p, f64p]
        L.hc_mol_destroy.argtypes = [C.c_void_p]
        L.hc_hash.restype = C.c_uint64
        L.hc_hash.argtypes = [C.c_uint64, u32p]
        L.hc_bit_op.restype = C.c_int
        L.hc_bit_op.argtypes = [C.c_int, C.POINTER(C.c_uint64), u8p]
        L.hc_diag.restype = C.c_double
        L.hc_diag.argtypes = [C.c_void_p, C.c_uint64]
        L.hc_sing_el.restype = C.c_double
        L.hc_sing_el.argtypes = [C.c_void_p, C.c_uint64, u8p]
        L.hc_doub_el.restype = C.c_double
        L.hc_doub_el.argtypes = [C.c_void_p, u8p]
        for nm in ("hc_sing_ex", "hc_doub_ex"):
            getattr(L, nm).restype = C.c_size_t
            getattr(L, nm).argtypes = [C.c_void_p, C.c_uint64, u8p]
        L.hc_count_singex.restype = C.c_size_t
        L.hc_count_singex.argtypes = [C.c_void_p, C.c_uint64]
        L.hc_find_nth_virt.restype = C.c_int
        L.hc_find_nth_virt.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int]
        L.hc_hb_row.restype = C.c_double
        L.hc_hb_row.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, f64p, C.POINTER(C.c_int)]
        L.hc_hb_wt.restype = C.c_double
        L.hc_hb_wt.argtypes = [C.c_void_p, C.c_int, C.c_uint64, u8p]
        L.hc_sing_counts.restype = C.c_uint
        L.hc_sing_counts.argtypes = [C.c_void_p, C.c_uint64, C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint)]
        L.hc_piv_samp.restype = C.c_size_t
        L.hc_piv_samp.argtypes = [f64p, C.c_size_t, u8p, C.c_double, C.c_uint32, u32p, C.POINTER(C.c_uint64)]
        L.hc_adjust_probs.restype = C.c_double
        L.hc_adjust_probs.argtypes = [f64p, C.c_size_t, u8p, C.POINTER(C.c_uint32), C.c_double, C.c_uint32, C.c_double]
        u64p = np.ctypeslib.ndpointer(np.uint64, flags="C")
        L.hc_hbpiv_begin.restype = C.c_void_p
        L.hc_hbpiv_begin.argtypes = [C.c_void_p, u64p, f64p, C.c_size_t, C.c_double, C.c_int, C.c_size_t]
        L.hc_hbpiv_end.argtypes = [C.c_void_p]
        L.hc_hbpiv_expand.restype = C.c_size_t
        L.hc_hbpiv_expand.argtypes = [C.c_void_p, C.c_int, f64p, C.c_size_t]
        L.hc_hbpiv_collapse.restype = C.c_size_t
        L.hc_hbpiv_collapse.argtypes = [C.c_void_p, C.c_int, f64p, u8p]
        L.hc_hbpiv_finalize.restype = C.c_size_t
        L.hc_hbpiv_finalize.argtypes = [C.c_void_p, C.c_double, f64p, u64p, u8p]
        u16p = np.ctypeslib.ndpointer(np.uint16, flags="C")
        L.hc_hbsys_rows.restype = C.c_size_t
        L.hc_hbsys_rows.argtypes = [C.c_void_p, C.c_int, f64p, u32p, f64p, C.c_size_t, u16p]
        L.hc_hbsys_accept.argtypes = [C.c_void_p, C.c_int, f64p, u64p, C.c_size_t]
        L.hc_hh_hub_diag.restype = C.c_uint
        L.hc_hh_hub_diag.argtypes = [C.c_uint64, C.c_uint]
        L.hc_hh_neighbors.argtypes = [C.c_uint64, C.c_uint, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.hc_hh_ref_ovlp.restype = C.c_double
        L.hc_hh_ref_ovlp.argtypes = [u64p, f64p, C.c_size_t, C.c_uint64, C.c_uint, C.c_uint, C.c_uint, C.c_double]
        L.hc_hh_total_ph.restype = C.c_uint
        L.hc_hh_total_ph.argtypes = [C.c_uint64, C.c_uint, C.c_uint, C.c_uint]
        L.hc_spawn_element.restype = C.c_uint64
        L.hc_spawn_element.argtypes = [C.c_uint64, u8p, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                       C.POINTER(C.c_double)]
        L.hc_hv_parent.restype = C.c_size_t
        L.hc_hv_parent.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_double, u64p, f64p, C.c_size_t]
        _lib = L
    return _lib
