"""ctypes binding of libtnb200.so -- one prototype per export of include/tn_c_api.h.

There is deliberately no fallback: if the CUDA library is missing or no B200 is
visible, importing succeeds (so CPU-only tooling can inspect symbols) but any
compute call raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libtnb200.so")


class tn_cplx(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double)]


class tn_trunc_t(C.Structure):
    _fields_ = [("cutoff", C.c_double), ("maxdim", C.c_int64), ("mindim", C.c_int64)]


class tn_lanczos_t(C.Structure):
    _fields_ = [("krylovdim", C.c_int32), ("maxiter", C.c_int32), ("tol", C.c_double)]


class tn_idx2_t(C.Structure):
    _fields_ = [("n0", C.c_int64), ("s0", C.c_int64), ("s1", C.c_int64)]


P = C.c_void_p
PP = C.POINTER(C.c_void_p)
I32, I64, F64, U64 = C.c_int32, C.c_int64, C.c_double, C.c_uint64
pI32, pI64, pF64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)

APPLY_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p)     # tn_apply_fn

# name -> argtypes ; all return int32 status except the two noted
PROTOTYPES = {
    "tn_ctx_create": [I32, PP],
    "tn_ctx_destroy": [P],
    "tn_sync": [P],
    "tn_ctx_stream": [P, PP],
    "tn_counters": [P, pI64, pI64, pI64],
    "tn_mps_upload": [P, I32, I32, I32, pI64, PP, I32, PP],
    "tn_mps_free": [P],
    "tn_mps_info": [P, pI32, pI32, pI32, pI32],
    "tn_mps_dims": [P, pI64],
    "tn_mps_download_site": [P, I32, P],
    "tn_mps_upload_site": [P, I32, pI64, P],
    "tn_mps_set_center": [P, I32],
    "tn_mps_maxbonddim": [P, pI64],
    "tn_mps_norm": [P, C.POINTER(tn_cplx)],
    "tn_mps_normalize": [P],
    "tn_mps_movecenter": [P, I32, tn_trunc_t],
    "tn_mps_replacesites": [P, P, I32, I32, I32, tn_trunc_t],
    "tn_mps_applyop": [P, I32, P],
    "tn_mps_bond_spectrum": [P, I32, pF64, I64, pI64],
    "tn_mpo_compress": [P, tn_trunc_t],
    "tn_mpo_apply": [P, P, tn_trunc_t, PP],
    "tn_mps_copy": [P, PP],
    "tn_mps_scale": [P, tn_cplx],
    "tn_expect_local": [P, I32, pI32, P, P],
    "tn_svd_trunc": [P, P, I64, I64, tn_trunc_t, P, pF64, P, pI64, pI32],
    "tn_heff_sharded_create": [I32, pI32, I64, I64, I32, I64, I64, I64, P, P, P, P, tn_cplx, PP],
    "tn_heff_sharded_apply": [P, P, P],
    "tn_heff_sharded_free": [P],
    "tn_svd_trunc_split": [P, P, I64, I64, tn_trunc_t, I32, P, pF64, P, pI64, pI32, I32, pF64],
    "tn_jacobi_qr_update_pass": [P, P, I64, I64, I32, P, P],
    "tn_svd_split_schedule": [I32, I32, P, I64, P],
    "tn_philox4x32_10": [P, P, P],
    "tn_qjmc_uniform": [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(C.c_double)],
    "tn_jacobi_pair_pass": [P, P, I64, I64, P, I32, P, P, P, P],
    "tn_svd_trunc_batched": [P, I32, P, I64, I64, tn_trunc_t, P, pF64, P, pI64, pI32],
    "tn_svd_set_precond": [I32],
    "tn_contract_strided": [P, I64, I64, I64, P, I64, tn_idx2_t, tn_idx2_t, I32, P, I64, tn_idx2_t, tn_idx2_t, I32,
                            P, I64, tn_idx2_t, tn_idx2_t, tn_cplx],
    "tn_contract_strided_dev": [P, I64, I64, I64, P, tn_idx2_t, tn_idx2_t, I32, P, tn_idx2_t, tn_idx2_t, I32, P, tn_idx2_t, tn_idx2_t, tn_cplx, tn_cplx],
    "tn_env_create": [P, P, P, P, tn_cplx, I32, PP],
    "tn_env_free": [P],
    "tn_env_buildleft": [P, I32],
    "tn_env_buildright": [P, I32],
    "tn_env_movecenter": [P, I32],
    "tn_env_center": [P, pI32],
    "tn_env_block_dims": [P, I32, pI64],
    "tn_env_block_download": [P, I32, P],
    "tn_env_block_upload": [P, I32, pI64, P],
    "tn_env_set_center": [P, I32],
    "tn_env_product_profile": [P, P, I32, P, I32, pF64],
    "tn_env_product": [P, P, I32, P],
    "tn_env_product_dev": [P, P, I32, P, I32],
    "tn_env_calculate": [P, C.POINTER(tn_cplx)],
    "tn_dmrg_sweep": [P, P, I32, tn_lanczos_t, tn_trunc_t, pF64, pI64],
    "tn_env_create_squared": [P, P, P, tn_cplx, I32, PP],
    "tn_env_product_n": [P, P, I32, I32, P],
    "tn_env_project": [P, I32, I32, P],
    "tn_envsum_create": [P, I32, PP, I32, PP],
    "tn_envsum_free": [P],
    "tn_envsum_movecenter": [P, I32],
    "tn_envsum_calculate": [P, C.POINTER(tn_cplx)],
    "tn_envsum_product": [P, P, I32, I32, P],
    "tn_envsum_project": [P, I32, I32, P],
    "tn_dmrg_sweep_sum": [P, P, I32, I32, tn_lanczos_t, tn_trunc_t, pF64, pI64],
    "tn_vmps_sweep": [P, P, I32, I32, tn_trunc_t, pI64],
    "tn_eigsolve": [P, P, I32, tn_lanczos_t, pF64, P, pI32],
    "tn_mps_site_ptr": [P, I32, PP],
    "tn_mps_replacesites_dev": [P, P, I32, I32, I32, tn_trunc_t],
    "tn_mps_upload_site_dev": [P, I32, pI64, P],
    "tn_memcpy_dev": [P, P, P, I64],
    "tn_svd_dist_begin": [P, P, I64, I64, PP, pI64, pI64, pI32, pF64],
    "tn_svd_dist_step": [P, pI32, I32, pF64],
    "tn_svd_dist_finish": [P, tn_trunc_t, I32, pI64],
    "tn_svd_dist_factors": [P, P, P, P],
    "tn_mps_replacesites_factored": [P, I32, I32, I32],
    "tn_eigsolve_fn": [P, I64, P, P, tn_lanczos_t, APPLY_FN, P, pF64, pI32],
    "tn_imps_create": [P, I32, I32, pI64, PP, PP, PP],
    "tn_imps_free": [P],
    "tn_imps_dims": [P, pI64],
    "tn_imps_download": [P, I32, P, pF64, pF64],
    "tn_itebd_apply_gate": [P, P, I32, tn_trunc_t],
    "tn_gates_upload": [P, I32, I32, pI32, pI32, pI32, PP, PP],
    "tn_gates_free": [P],
    "tn_apply_gates": [P, P, tn_trunc_t],
    "tn_apply_gates_fidelity": [P, P, tn_trunc_t, pF64],
    "tn_qjmc_run": [P, P, I32, pI32, P, pF64, I32, F64, tn_trunc_t, pF64, U64, U64, P, I32, P, pI32, pF64, I32, pI32, I32],
    "tn_inner_oplist": [P, P, I32, pI32, pI32, P, P, P],
    "tn_qjmc_ensemble": [I32, I32, I32, C.POINTER(C.c_uint64), I32, I32, pI64, PP, I32, I32, pI32, pI32, pI32, PP, I32, pI32, P, pF64,
                         I32, F64, tn_trunc_t, U64, P, I32, P, pI32, pI32, pF64, I32, I32],
}

_lib = None


class TNError(RuntimeError):
    pass


def load():
    """Load libtnb200.so (built by ``__graft_entry__.build()`` / ``make -C csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TNError(f"{LIB_PATH} not found: build it with `make -C tensornetworks.jl_b200/csrc` "
                      "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, args in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = I32
    lib.tn_last_error.argtypes = []
    lib.tn_last_error.restype = C.c_char_p
    lib.tn_version.argtypes = []
    lib.tn_version.restype = I32
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise TNError(load().tn_last_error().decode() + f" (status {status})")
