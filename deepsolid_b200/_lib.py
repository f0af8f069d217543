"""ctypes binding of libdeepsolid_b200.so (the C ABI in include/deepsolid_b200.h).

There is no fallback: if the shared library has not been built, or no CUDA device
is present, the product path raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libdeepsolid_b200.so"

c_double_p = C.POINTER(C.c_double)


class SystemDesc(C.Structure):
    _fields_ = [
        ("n_up", C.c_int32), ("n_dn", C.c_int32), ("n_atoms_prim", C.c_int32), ("n_atoms_sim", C.c_int32),
        ("prim_latvec", c_double_p), ("sim_latvec", c_double_p),
        ("prim_AV", c_double_p), ("prim_BV", c_double_p), ("sim_AV", c_double_p), ("sim_BV", c_double_p),
        ("prim_atoms", c_double_p), ("sim_atoms", c_double_p), ("sim_charges", c_double_p),
        ("klist_up", c_double_p), ("klist_dn", c_double_p),
        ("dist_kind", C.c_int32), ("mi_shifts", c_double_p), ("lattice_displacements", c_double_p),
        ("alpha", C.c_double), ("n_g", C.c_int32),
        ("gpoints", c_double_p), ("gweight", c_double_p), ("ion_exp_re", c_double_p), ("ion_exp_im", c_double_p),
        ("ee_const", C.c_double), ("ei_const", C.c_double), ("ii_total", C.c_double),
    ]


class NetDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("hidden_one", C.c_int32), ("hidden_two", C.c_int32), ("n_det", C.c_int32),
                ("distance_type", C.c_int32), ("envelope_type", C.c_int32),
                ("bias_orbitals", C.c_int32), ("full_det", C.c_int32), ("use_last_layer", C.c_int32)]


#: name -> (restype, argtypes); must list every DS_API symbol of the header
SIGNATURES = {
    "ds_last_error": (C.c_char_p, []),
    "ds_version": (C.c_int, []),
    "ds_ctx_create": (C.c_int, [C.POINTER(SystemDesc), C.POINTER(NetDesc), C.c_int, C.POINTER(C.c_void_p)]),
    "ds_ctx_destroy": (C.c_int, [C.c_void_p]),
    "ds_set_workspace_limit": (C.c_int, [C.c_void_p, C.c_size_t]),
    "ds_set_params": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int]),
    "ds_logpsi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ds_logpsi_vjp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p),
                                C.POINTER(C.c_int64), C.c_int, C.c_void_p]),
    "ds_logpsi_grad_x": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_void_p]),
    "ds_mcmc_step_one_electron": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_uint64,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ds_orbitals_vjp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_int64), C.c_int, C.c_void_p]),
    "ds_rho_q": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ds_kfac_factors": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_void_p),
                                  C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int, C.c_void_p]),
    "ds_orbitals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ds_orbitals_size": (C.c_int64, [C.c_void_p]),
    "ds_local_energy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "ds_ewald": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ds_ewald_ii": (C.c_double, [C.c_void_p]),
    "ds_mcmc_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_uint64, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ds_energy_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ds_stats_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_void_p,
                                     C.c_void_p]),
    "ds_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "ds_nccl_comm_init": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_char_p, C.c_int, C.c_int]),
    "ds_nccl_comm_destroy": (C.c_int, [C.c_void_p]),
    "ds_logpsi_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ds_local_energy_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "ds_mcmc_step_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_uint64, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "ds_launch_count": (C.c_int64, [C.c_void_p]),
    "ds_profile_reset": (C.c_int, [C.c_void_p]),
    "ds_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "ds_profile_get": (C.c_int, [C.c_void_p, c_double_p, C.POINTER(C.c_int64), c_double_p, c_double_p]),
    "ds_workspace_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "ds_debug_buffer": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "ds_debug_set_int": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "ds_ozaki_dgemm_probe": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                       c_double_p, c_double_p, C.c_void_p]),
    "ds_dgemm_probe": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m deepsolid_b200.build` "
            "(there is no CPU or pure-Python fallback for the hot path)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    """Map ds_status to the reference's error conventions (construction errors are
    ValueError; everything else RuntimeError)."""
    if rc == 0:
        return
    msg = load().ds_last_error().decode("utf-8", "replace")
    if rc in (-1, -3):
        raise ValueError(msg)
    if rc == -4:
        raise MemoryError(msg)
    raise RuntimeError(msg)
