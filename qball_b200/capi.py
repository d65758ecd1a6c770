"""ctypes binding of libqball_b200.so (include/qball_b200.h).

Arrays may be numpy arrays (host pointers) or CUDA torch tensors (device pointers, used in place); the C ABI tells
them apart itself.  The library is required: if it cannot be loaded this module raises -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

_lib = None


class QB200Error(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise QB200Error(f"{path} is missing: run `python -m qball_b200.build` (or __graft_entry__.build()); "
                         "the CUDA library is the product, there is no fallback")
    L = C.CDLL(path)
    vp, ip, dp, i, ll, d = C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_longlong, C.c_double
    L.qb200_last_error.restype = C.c_char_p
    L.qb200_version.restype = C.c_char_p
    L.qb200_device_count.restype = i
    L.qb200_plan_create.argtypes = [C.POINTER(vp), i, i, i, i, i, ip, ip, ip, ip, i, i, i]
    L.qb200_plan_destroy.argtypes = [vp]
    L.qb200_plan_set_stream.argtypes = [vp, vp]
    L.qb200_plan_set_workspace.argtypes = [vp, ll]
    L.qb200_plan_set_coefficient_tag.argtypes = [vp, ll]
    L.qb200_plan_query.argtypes = [vp, i]
    L.qb200_plan_query.restype = ll
    L.qb200_fft_backward.argtypes = [vp, dp, dp]
    L.qb200_fft_forward.argtypes = [vp, dp, dp]
    L.qb200_fft_backward_pair.argtypes = [vp, dp, dp, dp]
    L.qb200_fft_forward_pair.argtypes = [vp, dp, dp, dp]
    L.qb200_rs_mul_add.argtypes = [vp, i, i, dp, dp, dp, dp]
    L.qb200_compute_density.argtypes = [vp, i, i, dp, dp, dp]
    L.qb200_compute_current.argtypes = [vp, i, i, dp, dp, dp, dp]
    L.qb200_density_finish.argtypes = [vp, dp, d, dp, C.POINTER(d)]
    L.qb200_nl_create.argtypes = [C.POINTER(vp), i, i, i, d, dp]
    L.qb200_nl_add_species.argtypes = [vp, i, i, ip, dp, dp, dp]
    L.qb200_nl_set_positions.argtypes = [vp, i, dp]
    L.qb200_nl_set_lattice.argtypes = [vp, ip, dp, dp]
    L.qb200_nl_set_stream.argtypes = [vp, vp]
    L.qb200_nl_set_workspace.argtypes = [vp, ll]
    L.qb200_nl_destroy.argtypes = [vp]
    L.qb200_nl_energy.argtypes = [vp, i, i, dp, dp, i, dp, C.POINTER(d)]
    L.qb200_nl_betapsi.argtypes = [vp, i, i, dp, dp]
    L.qb200_nl_add_beta.argtypes = [vp, i, i, dp, dp]
    L.qb200_nl_spsi.argtypes = [vp, i, i, dp, dp, dp, dp]
    L.qb200_nl_last_enl.argtypes = [vp, dp]
    L.qb200_nl_update_twnl.argtypes = [vp, i, ip, ip, i, i, dp, d, dp, dp]
    L.qb200_nl_get_twnl.argtypes = [vp, i, dp]
    L.qb200_nl_update_twnl_semilocal.argtypes = [vp, i, ip, dp]
    L.qb200_nl_us_set_density_basis.argtypes = [vp, i, dp]
    L.qb200_nl_us_set_species.argtypes = [vp, i, i, ip, ip, dp, dp]
    L.qb200_nl_us_energy.argtypes = [vp, i, i, dp, dp, dp, i, dp, dp]
    L.qb200_nl_us_augment_density.argtypes = [vp, vp, i, i, dp, dp, dp, dp]
    L.qb200_nl_query.argtypes = [vp, i]
    L.qb200_nl_query.restype = ll
    L.qb200_hpsi.argtypes = [vp, vp, i, i, dp, dp, dp, dp, dp, C.POINTER(d)]
    L.qb200_exponential.argtypes = [vp, vp, i, i, dp, dp, dp, dp, i, d, d, dp]
    L.qb200_la_create.argtypes = [C.POINTER(vp), i, i, i]
    L.qb200_la_set_stream.argtypes = [vp, vp]
    L.qb200_la_set_workspace.argtypes = [vp, ll]
    L.qb200_la_destroy.argtypes = [vp]
    L.qb200_la_query.argtypes = [vp, i]
    L.qb200_la_query.restype = ll
    L.qb200_residual.argtypes = [vp, i, i, dp, i, dp, dp]
    L.qb200_gram.argtypes = [vp, i, i, dp, ip]
    L.qb200_gram_overlap.argtypes = [vp, i, i, dp, i, i, dp]
    L.qb200_gram_apply.argtypes = [vp, i, i, dp, dp, i, i, dp, ip]
    L.qb200_gram_sharded.argtypes = [vp, vp, i, i, dp, i, i, dp, ip]
    L.qb200_ekin_sums.argtypes = [vp, i, i, dp, dp, dp, dp, dp, dp, dp, dp]
    L.qb200_ekin_sums.restype = i
    L.qb200_comm_get_unique_id.argtypes = [vp]
    L.qb200_comm_init.argtypes = [C.POINTER(vp), i, vp, i, i]
    L.qb200_comm_destroy.argtypes = [vp]
    L.qb200_comm_query.argtypes = [vp, i]
    L.qb200_comm_query.restype = ll
    L.qb200_allreduce_rho.argtypes = [vp, dp, ll, vp]
    L.qb200_allreduce_scalars.argtypes = [vp, dp, i]
    for name in ("qb200_comm_get_unique_id", "qb200_comm_init", "qb200_comm_destroy", "qb200_allreduce_rho", "qb200_allreduce_scalars"):
        getattr(L, name).restype = i
    L.qb200_update_vhxc.argtypes = [vp, i, dp, dp, dp, dp, dp, dp, d, dp, dp, dp]
    L.qb200_update_vhxc.restype = i
    L.qb200_diag.argtypes = [vp, i, i, dp, dp, i, dp, ip]
    L.qb200_diag.restype = i
    L.qb200_psda_update.argtypes = [vp, vp, i, i, dp, dp, dp, dp, dp, dp, i, C.POINTER(d)]
    L.qb200_psda_update.restype = i
    L.qb200_measure_fp64_peak.argtypes = [i, C.POINTER(d)]
    L.qb200_measure_fp64_peak.restype = i
    L.qb200_profile_enable.argtypes = [i]
    L.qb200_profile_read.argtypes = [C.POINTER(d), C.POINTER(ll), i]
    for name in ("qb200_profile_enable", "qb200_profile_read", "qb200_plan_create", "qb200_plan_destroy", "qb200_plan_set_stream", "qb200_plan_set_workspace", "qb200_plan_set_coefficient_tag",
                 "qb200_fft_backward", "qb200_fft_forward", "qb200_fft_backward_pair", "qb200_fft_forward_pair",
                 "qb200_rs_mul_add", "qb200_compute_density", "qb200_density_finish", "qb200_nl_create", "qb200_nl_add_species",
                 "qb200_nl_set_positions", "qb200_nl_set_lattice", "qb200_nl_set_stream", "qb200_nl_set_workspace", "qb200_nl_destroy", "qb200_nl_energy", "qb200_nl_betapsi", "qb200_nl_add_beta", "qb200_nl_spsi", "qb200_nl_last_enl", "qb200_nl_update_twnl", "qb200_nl_update_twnl_semilocal", "qb200_nl_get_twnl", "qb200_nl_us_set_density_basis", "qb200_nl_us_set_species", "qb200_nl_us_energy", "qb200_nl_us_augment_density", "qb200_hpsi", "qb200_exponential",
                 "qb200_compute_current", "qb200_la_create", "qb200_la_set_stream", "qb200_la_set_workspace", "qb200_la_destroy", "qb200_residual", "qb200_gram", "qb200_gram_overlap", "qb200_gram_apply", "qb200_gram_sharded"):
        getattr(L, name).restype = i
    _lib = L
    return L


def _check(rc: int, what: str):
    if rc != 0:
        raise QB200Error(f"{what} failed ({rc}): {load().qb200_last_error().decode()}")


def ptr(a):
    """raw address of a numpy array (host) or torch tensor (host or device); None -> NULL"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags.c_contiguous, "array must be C-contiguous"
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        assert a.is_contiguous(), "tensor must be contiguous"
        return a.data_ptr()
    raise TypeError(type(a))


def _iarr(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def device_count() -> int:
    return load().qb200_device_count()


PROFILE_CATEGORIES = ("k_zcol_bwd", "xy_stage", "k_zcol_fwd", "k_fnl", "k_fnl_finish", "k_back", "k_rho_reduce", "k_anl_gen", "xy_density")


def measure_fp64_peak(device: int = 0):
    """(DMMA TFLOP/s, DFMA TFLOP/s) measured on `device` by the library's issue-rate loops"""
    out = (C.c_double * 2)()
    _check(load().qb200_measure_fp64_peak(int(device), out), "qb200_measure_fp64_peak")
    return float(out[0]), float(out[1])


def profile_enable(on: bool):
    load().qb200_profile_enable(int(on))


def profile_read():
    """{category: (total_ms, launches)} for the launches recorded since the last read"""
    n = len(PROFILE_CATEGORIES)
    ms = (C.c_double * n)()
    cnt = (C.c_longlong * n)()
    load().qb200_profile_read(ms, cnt, n)
    return {k: (ms[i], int(cnt[i])) for i, k in enumerate(PROFILE_CATEGORIES)}
