"""ctypes binding of liboqb_b200.so (the C ABI declared in include/oqb.h).

There is no CPU fallback: importing this module without the built library, or creating a
context without a B200, raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboqb_b200.so")

c_void_p, c_int, c_i64, c_size_t = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
c_dp = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)

# name -> argtypes (restype is int unless listed in _RESTYPE); mirrors include/oqb.h 1:1
SIGNATURES = {
    "oqb_version": [],
    "oqb_create": [c_int, C.POINTER(c_void_p)],
    "oqb_destroy": [c_void_p],
    "oqb_last_error": [c_void_p],
    "oqb_sync": [c_void_p],
    "oqb_set_stream": [c_void_p, c_void_p],
    "oqb_get_stream": [c_void_p, C.POINTER(c_void_p)],
    "oqb_launch_count": [c_void_p, C.POINTER(C.c_uint64)],
    "oqb_set_zgemm_mode": [c_void_p, c_int],
    "oqb_get_zgemm_mode": [c_void_p, C.POINTER(c_int)],
    "oqb_malloc": [c_void_p, c_size_t, C.POINTER(c_void_p)],
    "oqb_free": [c_void_p, c_void_p],
    "oqb_malloc_host": [c_void_p, c_size_t, C.POINTER(c_void_p)],
    "oqb_free_host": [c_void_p, c_void_p],
    "oqb_upload": [c_void_p, c_void_p, c_void_p, c_size_t],
    "oqb_download": [c_void_p, c_void_p, c_void_p, c_size_t],
    "oqb_axpy_batched": [c_void_p, c_i64, c_int, c_dp, c_void_p, c_i64, c_void_p],
    "oqb_zgemm_batched": [c_void_p, c_int, c_int, c_int, c_int, c_int, c_dp, c_void_p, c_i64, c_i64, c_void_p,
                          c_i64, c_i64, c_dp, c_void_p, c_i64, c_i64, c_int],
    "oqb_dense_create": [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, C.POINTER(c_void_p)],
    "oqb_dense_destroy": [c_void_p],
    "oqb_dense_hamiltonian": [c_void_p, c_int, c_dp, c_int, c_void_p],
    "oqb_dense_update_cache": [c_void_p, c_int, c_dp, c_int, c_dp, c_int, c_void_p],
    "oqb_dense_update_vectorized_cache": [c_void_p, c_int, c_dp, c_int, c_void_p],
    "oqb_dense_rhs": [c_void_p, c_int, c_dp, c_int, c_dp, c_int, c_void_p, c_void_p],
    "oqb_dense_rhs_host": [c_void_p, c_int, c_dp, c_int, c_dp, c_int, c_void_p, c_void_p],
    "oqb_lindblad_acc": [c_void_p, c_int, c_dp, c_int, c_void_p, c_void_p],
    "oqb_redfield_apply_acc": [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p],
    "oqb_redfield_apply_diag_acc": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p],
    "oqb_redfield_apply_mono_acc": [c_void_p, C.c_int, C.c_int, c_void_p, c_void_p, c_void_p, c_void_p, C.c_int, c_void_p, c_void_p],
    "oqb_redfield_vectorized_acc": [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p],
    "oqb_lambda_create": [c_void_p, c_int, c_int, c_void_p, C.POINTER(c_void_p)],
    "oqb_lambda_destroy": [c_void_p],
    "oqb_lambda_set_unitary": [c_void_p, c_int, c_int, c_void_p],
    "oqb_lambda_reserve_slots": [c_void_p, c_i64],
    "oqb_lambda_eval": [c_void_p, c_int, c_i32p, c_i32p, c_i32p, c_dp, c_dp, c_dp, c_i64p, c_dp],
    "oqb_lambda_integrate": [c_void_p, c_void_p, c_int, c_i32p, c_i32p, c_dp, c_dp, c_dp, c_dp, c_int, c_void_p, c_void_p,
                             C.c_double, C.c_double, c_i64, c_void_p, c_i64p, c_i64p, c_i64p],
    "oqb_lambda_integrate_ohmic": [c_void_p, c_void_p, c_int, c_i32p, c_i32p, c_dp, c_dp, c_dp, c_dp, c_int, C.c_double, C.c_double,
                                   C.c_double, C.c_double, C.c_double, c_i64, c_void_p, c_i64p, c_i64p, c_i64p],
    "oqb_lambda_dims": [c_void_p, C.POINTER(c_int), C.POINTER(c_int)],
    "oqb_lambda_total": [c_void_p, c_int, c_i64p, c_i64p, c_void_p, c_dp],
    "oqb_ame_create": [c_void_p, c_int, c_int, c_int, c_void_p, C.POINTER(c_void_p)],
    "oqb_ame_destroy": [c_void_p],
    "oqb_ame_set_eigen": [c_void_p, c_dp, c_void_p],
    "oqb_ame_set_eigen_device": [c_void_p, c_void_p, c_void_p],
    "oqb_ame_set_davies": [c_void_p, c_dp, c_dp, c_i64, c_i32p, c_dp, c_i64, c_i32p, c_dp],
    "oqb_set_real_operand_mode": [c_void_p, c_int],
    "oqb_get_real_operand_mode": [c_void_p, C.POINTER(c_int)],
    "oqb_ame_eigenvectors_real": [c_void_p, C.POINTER(c_int)],
    "oqb_ame_rebind": [c_void_p, c_void_p],
    "oqb_ame_set_onesided_ohmic": [c_void_p, c_int, c_i32p, c_i32p, C.c_double, C.c_double, C.c_double],
    "oqb_ame_set_lambshift_spline": [c_void_p, c_int, C.c_double, C.c_double, c_void_p],
    "oqb_ohmic_lambshift": [c_void_p, C.c_double, C.c_double, C.c_double, c_int, c_void_p, C.c_double, C.c_double, c_int,
                            c_void_p, c_void_p, c_void_p],
    "oqb_ohmic_correlation": [c_void_p, C.c_double, C.c_double, C.c_double, c_int, c_void_p, c_void_p],
    "oqb_ohmic_timescales": [c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, c_int,
                             c_dp, c_dp, c_dp, c_dp],
    "oqb_bspline_eval": [c_void_p, c_int, C.c_double, C.c_double, c_void_p, c_int, c_void_p, c_void_p],
    "oqb_ame_clear_davies": [c_void_p],
    "oqb_set_option": [c_void_p, C.c_char_p, c_i64],
    "oqb_get_option": [c_void_p, C.c_char_p, c_i64p],
    "oqb_ame_gaps_build": [c_void_p, c_int, c_int, c_i64p, c_i64p],
    "oqb_ame_gaps_keys": [c_void_p, c_dp],
    "oqb_ame_gaps_fetch": [c_void_p, c_i64p, c_i32p, c_i32p, c_i32p, c_i32p],
    "oqb_ame_add_davies": [c_void_p, c_int, c_int, c_dp, c_dp, c_dp, c_dp, C.c_double, C.c_double],
    "oqb_ame_add_davies_ohmic": [c_void_p, c_int, c_int, C.c_double, C.c_double, C.c_double, c_int, C.c_double, C.c_double,
                                 c_void_p],
    "oqb_ame_add_correlated": [c_void_p, c_int, c_int, c_dp, c_dp, c_dp, c_dp, C.c_double, C.c_double],
    "oqb_ame_davies_stats": [c_void_p, c_i64p, c_i64p, c_i64p, c_i64p, c_i64p],
    "oqb_ame_get_rotated_couplings": [c_void_p, c_void_p],
    "oqb_ame_set_onesided": [c_void_p, c_int, c_i32p, c_i32p, c_dp, c_dp, c_int],
    "oqb_ame_clear_onesided": [c_void_p],
    "oqb_ame_rhs": [c_void_p, c_int, c_void_p, c_void_p],
    "oqb_ame_rhs_host": [c_void_p, c_int, c_void_p, c_void_p],
    "oqb_ame_eig_acc": [c_void_p, c_int, c_void_p, c_void_p, c_int],
    "oqb_ame_update_cache": [c_void_p, c_void_p],
    "oqb_ame_set_jump": [c_void_p, c_int, c_i64p, c_i32p, c_dp],
    "oqb_ame_clear_jump": [c_void_p],
    "oqb_ame_jump": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int],
    "oqb_sparse_create": [c_void_p, c_int, c_int, c_i64p, c_i64p, c_void_p, c_i64p, c_int, c_i64p, c_i64p, c_void_p,
                          c_i64p, C.POINTER(c_void_p)],
    "oqb_sparse_destroy": [c_void_p],
    "oqb_sparse_pattern_nnz": [c_void_p, c_i64p],
    "oqb_sparse_pattern": [c_void_p, c_i64p, c_i64p],
    "oqb_sparse_update_cache": [c_void_p, c_int, c_dp, c_int, c_dp, c_int, c_void_p],
    "oqb_sparse_update_vectorized_cache": [c_void_p, c_int, c_dp, c_int, c_void_p],
    "oqb_sparse_commutator": [c_void_p, c_int, c_dp, c_int, c_void_p, c_void_p],
    "oqb_traj_matvec": [c_void_p, c_int, c_dp, c_int, c_dp, c_int, c_void_p, c_void_p],
    "oqb_traj_matvec_host": [c_void_p, c_int, c_dp, c_int, c_dp, c_int, c_void_p, c_void_p],
    "oqb_traj_sell_stats": [c_void_p, C.POINTER(C.c_int), c_dp],
    "oqb_dense_structure_probe": [C.c_int, C.c_int, c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), c_void_p],
    "oqb_sell_plan_probe": [C.c_int, C.c_int, C.POINTER(C.POINTER(C.c_int64)), C.POINTER(C.POINTER(C.c_int32)), C.c_int, c_dp,
                            C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "oqb_traj_norm2": [c_void_p, c_int, c_void_p, c_void_p],
    "oqb_traj_jump": [c_void_p, c_int, c_dp, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int],
    "oqb_traj_apply_choice": [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int],
    "oqb_resample": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "oqb_traj_rng_init": [c_void_p, c_int, C.c_uint64, c_void_p],
    "oqb_traj_rng_uniform": [c_void_p, c_int, c_void_p, c_void_p],
    "oqb_traj_evolve": [c_void_p, c_int, c_int, C.c_double, c_dp, c_int, c_dp, c_int, c_void_p, c_void_p, c_void_p,
                        c_void_p, c_void_p, c_int, c_int],
    "oqb_ensemble_population": [c_void_p, c_int, c_int, c_void_p, c_void_p],
    "oqb_expect_diag": [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p],
    "oqb_expect": [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p],
    "oqb_ame_get_levels": [c_void_p, c_dp],
    "oqb_ame_get_eigenvectors": [c_void_p, c_void_p],
    "oqb_liouv_create": [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, C.POINTER(c_void_p)],
    "oqb_liouv_destroy": [c_void_p],
    "oqb_liouv_add_davies_ohmic": [c_void_p, c_int, c_int, C.c_double, C.c_double, C.c_double, c_int, C.c_double, C.c_double, c_void_p],
    "oqb_liouv_add_davies_fn": [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p],
    "oqb_liouv_add_onesided_ohmic": [c_void_p, c_int, c_i32p, c_i32p, C.c_double, C.c_double, C.c_double, c_int, C.c_double,
                                     C.c_double, c_void_p],
    "oqb_liouv_add_onesided_fn": [c_void_p, c_int, c_i32p, c_i32p, c_void_p, c_void_p, c_void_p, c_void_p],
    "oqb_liouv_set_plan_cache": [c_void_p, c_int],
    "oqb_liouv_rhs": [c_void_p, c_int, c_dp, c_int, c_void_p, c_void_p],
    "oqb_liouv_rhs_host": [c_void_p, c_int, c_dp, c_int, c_void_p, c_void_p],
    "oqb_liouv_update_cache": [c_void_p, c_dp, c_void_p],
    "oqb_liouv_prepare": [c_void_p, c_dp, C.POINTER(c_void_p)],
    "oqb_liouv_stats": [c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), c_dp, c_dp],
    "oqb_dense_eigen": [c_void_p, c_dp, c_void_p, c_void_p],
    "oqb_dense_info": [c_void_p, C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int)],
    "oqb_expect_sum": [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p],
    "oqb_nccl_version": [C.POINTER(c_int)],
    "oqb_comm_unique_id": [c_void_p],
    "oqb_comm_init_rank": [c_void_p, c_int, c_int, c_void_p, C.POINTER(c_void_p)],
    "oqb_comm_init_all": [c_int, C.POINTER(c_void_p), C.POINTER(c_void_p)],
    "oqb_comm_destroy": [c_void_p],
    "oqb_comm_rank": [c_void_p, C.POINTER(c_int), C.POINTER(c_int)],
    "oqb_comm_group_start": [],
    "oqb_comm_group_end": [],
    "oqb_allreduce_obs": [c_void_p, c_void_p, c_i64],
    "oqb_allgather_obs": [c_void_p, c_void_p, c_void_p, c_i64],
    "oqb_probe_fp64": [c_void_p, c_dp, c_dp],
    "oqb_probe_smem_gather": [c_void_p, c_int, c_dp],
    "oqb_profile_begin": [c_void_p],
    "oqb_profile_end": [c_void_p, c_dp, C.POINTER(C.c_uint64), c_dp],
}
_RESTYPE = {"oqb_last_error": C.c_char_p}

_lib = None


class OQBError(RuntimeError):
    pass


def load():
    """Load the shared library (raises if it has not been built — no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OQBError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()) first; "
                       "this package has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if include/oqb.h and the .so disagree
        fn.argtypes = argtypes
        fn.restype = _RESTYPE.get(name, c_int)
    _lib = lib
    return lib


def dptr(a):
    """double* view of a C-contiguous float64 numpy array (or None)."""
    return None if a is None else a.ctypes.data_as(c_dp)
