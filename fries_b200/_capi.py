"""ctypes binding of libfries_b200.so (include/fries_b200.h).  The library is the product; there is no
Python or CPU fallback: importing this module without the built library raises, and every compute call
fails with FRIES_ERR_CUDA on a machine without an sm_100 GPU."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FRIES_B200_LIB") or os.path.join(HERE, "libfries_b200.so")  # (the override: A/B of two builds)

OK, ERR_ARG, ERR_CUDA, ERR_CAPACITY, ERR_STATE = 0, -1, -2, -3, -4
INI_FLAG = 1 << 63
MAX_SUB = 32


class FriesError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[fries_b200 {code}] {msg}")
        self.code = code


class FrisysParams(C.Structure):
    _fields_ = [("eps", C.c_double), ("init_thresh", C.c_double), ("p_doub", C.c_double), ("new_hb", C.c_int),
                ("matr_samp", C.c_uint), ("target_nonz", C.c_uint), ("en_shift", C.c_double)]


class FrifullParams(C.Structure):
    _fields_ = [("eps", C.c_double), ("target_nonz", C.c_uint), ("en_shift", C.c_double), ("adjust_shift", C.c_int),
                ("damp_factor", C.c_double), ("target_norm", C.c_double), ("last_one_norm", C.c_double)]


class FrisysHhParams(C.Structure):
    _fields_ = [("eps", C.c_double), ("init_thresh", C.c_double), ("hub_u", C.c_double), ("ph_freq", C.c_double),
                ("elec_ph", C.c_double), ("hf_en", C.c_double), ("target_nonz", C.c_uint), ("en_shift", C.c_double),
                ("ref_key", C.c_uint64)]


class IterStats(C.Structure):
    _fields_ = [("glob_norm", C.c_double), ("numer", C.c_double), ("denom", C.c_double), ("n_kept", C.c_uint64),
                ("n_matrix_samples", C.c_uint64), ("n_spawned", C.c_uint64), ("curr_size", C.c_uint64)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C fries_b200/csrc` (or __graft_entry__.build()); "
                          "fries_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, sz, u, i, d = C.c_void_p, C.c_size_t, C.c_uint, C.c_int, C.c_double
    P = C.POINTER
    sig = {
        "fries_last_error": (C.c_char_p, []),
        "fries_version": (i, []),
        "fries_ctx_create": (i, [i, P(vp)]),
        "fries_ctx_destroy": (i, [vp]),
        "fries_ctx_set_stream": (i, [vp, vp]),
        "fries_ctx_sync": (i, [vp]),
        "fries_ctx_sm_count": (i, [vp]),
        "fries_ctx_launch_count": (C.c_uint64, [vp]),
        "fries_ctx_set_profile": (i, [vp, i]),
        "fries_ctx_kernel_ms": (i, [vp, C.c_char_p, P(d), P(C.c_uint64)]),
        "fries_hash_owner": (i, [vp, vp, sz, vp, i, i, vp, vp]),
        "fries_hash_owner_dev": (i, [vp, vp, sz, vp, i, i, vp, vp]),
        "fries_bit_op": (i, [vp, i, vp, vp, sz, vp]),
        "fries_find_preserve": (i, [vp, vp, sz, P(u), P(d), vp, P(d)]),
        "fries_sys_comp": (i, [vp, vp, sz, vp, i, i, u, vp, d]),
        "fries_find_preserve_dev": (i, [vp, vp, sz, u, vp, vp]),
        "fries_sys_comp_dev": (i, [vp, vp, sz, vp, vp, d, vp]),
        "fries_piv_samp_serial": (i, [vp, vp, sz, d, C.c_uint32, vp, vp, P(sz)]),
        "fries_piv_samp_dev": (i, [vp, vp, sz, d, C.c_uint32, vp, vp, vp, sz, vp]),
        "fries_adjust_probs": (i, [vp, vp, sz, P(C.c_uint32), d, C.c_uint32, d, vp, P(d)]),
        "fries_piv_budget": (i, [vp, i, C.c_uint32, vp, P(sz), vp]),
        "fries_piv_comp": (i, [vp, vp, sz, C.c_uint32, vp, vp, P(sz), vp, i, i, i, C.c_uint32]),
        "fries_vec_compress": (i, [vp, u, u, C.c_uint32, i, vp, sz, P(sz)]),
        "fries_comp_sub": (i, [vp, vp, sz, vp, vp, sz, vp, u, d, vp, vp, sz, P(sz), P(u), P(d)]),
        "fries_mol_create": (i, [vp, u, u, u, vp, vp, vp, P(vp)]),
        "fries_mol_destroy": (i, [vp]),
        "fries_mol_hb_tables": (i, [vp] + [vp] * 7),
        "fries_mol_diag": (i, [vp, vp, sz, vp]),
        "fries_mol_sing_el": (i, [vp, vp, vp, sz, vp]),
        "fries_mol_doub_el": (i, [vp, vp, sz, vp]),
        "fries_mol_sing_ex": (i, [vp, vp, sz, vp, vp, sz]),
        "fries_mol_doub_ex": (i, [vp, vp, sz, vp, vp, sz]),
        "fries_mol_hb_rows": (i, [vp, i, vp, vp, sz, vp, vp, vp]),
        "fries_mol_hb_wt": (i, [vp, i, vp, vp, sz, vp]),
        "fries_apply_hbpp_sys": (i, [vp, vp, vp, sz, d, i, vp, u, sz, vp, vp, vp, sz, P(sz)]),
        "fries_apply_hbpp_piv": (i, [vp, vp, vp, sz, d, i, vp, sz, P(sz), u, sz, vp, vp, vp, sz, P(sz)]),
        "fries_debug_hbpp_stage": (i, [vp, vp, vp, sz, d, i, vp, u, sz, i, vp, vp, vp, vp, P(sz)]),
        "fries_hbpp_states": (i, [vp, vp]),
        "fries_hbpp_timeline": (i, [vp, i, vp]),
        "fries_hbpp_cta_marks": (i, [vp, i, vp, P(i)]),
        "fries_hbpp_round_stamps": (i, [vp, i, vp]),
        "fries_debug_set_repeat": (i, [i]),
        "fries_debug_set_bracket": (i, [i]),
        "fries_debug_set_perturb": (i, [d]),
        "fries_debug_last_fast": (i, [P(i)]),
        "fries_debug_stage_ctas": (i, [P(i)]),
        "fries_debug_stage_engine": (i, [P(i)]),
        "fries_debug_vec_phase": (i, [vp, vp, vp, u, d, vp]),
        "fries_comm_create": (i, [vp, i, i, P(vp), vp]),
        "fries_comm_connect": (i, [vp, vp]),
        "fries_comm_destroy": (i, [vp]),
        "fries_comm_error": (i, [vp, P(C.c_uint64)]),
        "fries_comm_pingpong": (i, [vp, i, i, i, P(d)]),
        "fries_ctx_set_comm": (i, [vp, vp]),
        "fries_hbpp_set_route": (i, [vp, vp, vp, vp, vp, sz]),
        "fries_comm_route_create": (i, [vp, sz, vp]),
        "fries_comm_route_connect": (i, [vp, vp]),
        "fries_hbpp_set_route_p2p": (i, [vp, vp]),
        "fries_frisys_mol_spawn": (i, [vp, vp, vp, P(FrisysParams), vp]),
        "fries_frisys_mol_finish": (i, [vp, vp, vp, P(FrisysParams), vp, vp, P(IterStats)]),
        "fries_vec_create_hh": (i, [vp, sz, u, u, u, u, vp, vp, i, i, P(vp)]),
        "fries_vec_set_min_del_idx": (i, [vp, sz]),
        "fries_vec_set_dense": (i, [vp, sz]),
        "fries_vec_set_dense_total": (i, [vp, sz]),
        "fries_vec_set_deterministic": (i, [vp, i]),
        "fries_hh_batch": (i, [vp, i, vp, vp, sz, u, u, u, C.c_uint64, d, vp]),
        "fries_frisys_hh_setup": (i, [vp, sz, P(vp)]),
        "fries_frisys_hh_iterate": (i, [vp, vp, P(FrisysHhParams), vp, P(IterStats)]),
        "fries_frifull_hh_iterate": (i, [vp, vp, P(FrisysHhParams), C.c_double, P(IterStats)]),
        "fries_vec_create": (i, [vp, sz, u, u, u, vp, vp, i, i, P(vp)]),
        "fries_vec_destroy": (i, [vp]),
        "fries_vec_add": (i, [vp, vp, vp, vp, sz, u, u]),
        "fries_vec_add_dev": (i, [vp, vp, vp, sz, vp, u, u]),
        "fries_vec_curr_size": (i, [vp, P(sz)]),
        "fries_vec_n_nonz": (i, [vp, P(sz)]),
        "fries_vec_nonini_occ_add": (i, [vp, P(C.c_uint64)]),
        "fries_vec_download": (i, [vp, vp, vp, sz, P(sz)]),
        "fries_vec_upload": (i, [vp, vp, vp, sz]),
        "fries_vec_del": (i, [vp, vp, sz]),
        "fries_vec_dot": (i, [vp, vp, vp, sz, u, P(d)]),
        "fries_vec_local_norm": (i, [vp, u, P(d)]),
        "fries_vec_two_norm": (i, [vp, u, P(d)]),
        "fries_vec_row_op": (i, [vp, i, u, u, d]),
        "fries_vec_set_diag_mol": (i, [vp, vp, d]),
        "fries_h_apply": (i, [vp, vp, u, u, d, d]),
        "fries_h_apply_last_spawned": (i, [vp, P(C.c_uint64)]),
        "fries_h_apply_routed": (i, [vp, vp, vp, u, u, d, d, P(C.c_uint64)]),
        "fries_frisys_mol_setup": (i, [vp, vp, sz, vp, vp, sz, vp, vp, sz, P(vp)]),
        "fries_hbpp_destroy": (i, [vp]),
        "fries_frisys_mol_iterate": (i, [vp, vp, vp, P(FrisysParams), vp, P(IterStats)]),
        "fries_frifull_mol_iterate": (i, [vp, vp, vp, P(FrifullParams), d, P(IterStats)]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    return L, sorted(sig)


lib, EXPORTS = _load()


def check(rc: int):
    if rc != OK:
        raise FriesError(rc, lib.fries_last_error().decode(errors="replace"))


def ptr(a):
    """raw pointer of a numpy array (None -> NULL) or pass through an int device pointer"""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


def arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)
