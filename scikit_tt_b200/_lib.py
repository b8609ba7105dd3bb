"""ctypes binding of libsktt_b200.so (the C-ABI declared in include/sktt_b200.h).

There is deliberately no CPU fallback: if the shared library is missing, or no sm_100 device is
visible when a context is requested, the import / call fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsktt_b200.so")

F64, C128 = 0, 1
CONJ_ROW, CONJ_COL = 0, 1

i64, i32, dbl, vp = C.c_int64, C.c_int, C.c_double, C.c_void_p
pdbl, pint = C.POINTER(C.c_double), C.POINTER(C.c_int)


class Idx2(C.Structure):
    _fields_ = [("d", i64), ("s_hi", i64), ("s_lo", i64)]


class LocalOp(C.Structure):
    _fields_ = [("sites", i32), ("r", i64), ("R", i64), ("m", i64), ("n", i64), ("R2", i64),
                ("m2", i64), ("n2", i64), ("R3", i64), ("r3", i64),
                ("Lst", vp), ("A1", vp), ("A2", vp), ("Rst", vp), ("image", vp)]


# name -> (restype, argtypes); kept in one table so tests can check the exported symbol set
SIGNATURES = {
    "sktt_version": (i32, []),
    "sktt_ctx_create": (i32, [i32, vp, C.POINTER(vp)]),
    "sktt_ctx_destroy": (i32, [vp]),
    "sktt_ctx_set_stream": (i32, [vp, vp]),
    "sktt_last_error": (C.c_char_p, [vp]),
    "sktt_launch_count": (i64, [vp]),
    "sktt_ctx_set_gemm_mode": (i32, [vp, i32]),
    "sktt_ctx_set_debug": (i32, [vp, i32]),
    "sktt_ctx_set_qr_deferred": (i32, [vp, i32]),
    "sktt_qr_deferred_failures": (i32, [vp, vp]),
    "sktt_scratch_peek": (i32, [vp, i64, i64, vp]),
    "sktt_gemm2": (i32, [vp, i32, i64, i64, i64, pdbl, vp, Idx2, Idx2, i32, vp, Idx2, Idx2, i32, pdbl, vp, Idx2, Idx2]),
    "sktt_stack_op_work": (i64, [i64] * 6),
    "sktt_stack_left_op": (i32, [vp, i32] + [i64] * 6 + [vp] * 5 + [i32]),
    "sktt_stack_right_op": (i32, [vp, i32] + [i64] * 6 + [vp] * 5),
    "sktt_stack_left_rhs": (i32, [vp, i32] + [i64] * 5 + [vp] * 5),
    "sktt_stack_right_rhs": (i32, [vp, i32] + [i64] * 5 + [vp] * 5),
    "sktt_micro_matrix_als": (i32, [vp, i32] + [i64] * 6 + [vp] * 5),
    "sktt_micro_matvec_als": (i32, [vp, i32] + [i64] * 6 + [vp] * 6),
    "sktt_micro_matrix_mals_work": (i64, [i64] * 9),
    "sktt_micro_matrix_mals": (i32, [vp, i32] + [i64] * 9 + [vp] * 6),
    "sktt_micro_matvec_mals_work": (i64, [i64] * 9),
    "sktt_micro_matvec_mals": (i32, [vp, i32] + [i64] * 9 + [vp] * 7),
    "sktt_micro_rhs_als": (i32, [vp, i32] + [i64] * 5 + [vp] * 5),
    "sktt_micro_rhs_mals": (i32, [vp, i32] + [i64] * 7 + [vp] * 6),
    "sktt_rank1_update": (i32, [vp, i32, i64, dbl, vp, vp]),
    "sktt_lu_factor": (i32, [vp, i32, i64, vp, vp, pint]),
    "sktt_lu_solve": (i32, [vp, i32, i64, i64, vp, vp, vp]),
    "sktt_lu_fused_max_n": (i64, []),
    "sktt_lu_solve_fused": (i32, [vp, i64, vp, vp, vp, vp]),
    "sktt_chol_factor": (i32, [vp, i32, i64, vp, pint]),
    "sktt_chol_solve": (i32, [vp, i32, i64, i64, vp, vp]),
    "sktt_chol_trsm": (i32, [vp, i32, i64, i64, vp, vp, i32]),
    "sktt_local_op_image_size": (i64, [vp, i32, C.POINTER(LocalOp)]),
    "sktt_local_op_prepare": (i32, [vp, i32, C.POINTER(LocalOp), vp]),
    "sktt_local_matvec_work": (i64, [C.POINTER(LocalOp)]),
    "sktt_local_matvec": (i32, [vp, i32, C.POINTER(LocalOp), vp, vp, vp]),
    "sktt_local_op_tiled_len": (i64, [vp, i32, C.POINTER(LocalOp)]),
    "sktt_local_matvec_tiled": (i32, [vp, i32, C.POINTER(LocalOp), vp, vp, vp]),
    "sktt_local_matvec_tiled_repeat": (i32, [vp, i32, C.POINTER(LocalOp), vp, vp, vp, i32]),
    "sktt_krylov_work": (i64, [C.POINTER(LocalOp), i32, i32]),
    "sktt_krylov_solve": (i32, [vp, i32, C.POINTER(LocalOp), i32, i32, vp, vp, dbl, i32, vp, pint, pdbl]),
    "sktt_krylov_solve_refined": (i32, [vp, i32, C.POINTER(LocalOp), vp, vp, dbl, i32, i32, vp, pint, pdbl, pint]),
    "sktt_gauge_factor": (i32, [vp, i32, i64, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64, i64]),
    "sktt_gauge_push": (i32, [vp, i32, i64, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64, i64]),
    "sktt_krylov_solve_refined_async": (i32, [vp, i32, C.POINTER(LocalOp), vp, vp, dbl, i32, vp, vp]),
    "sktt_qr_work": (i64, [i64, i64]),
    "sktt_qr_left": (i32, [vp, i32, i64, i64, vp, vp, vp, vp]),
    "sktt_rq_right": (i32, [vp, i32, i64, i64, vp, vp, vp, vp]),
    "sktt_svd_work": (i64, [i64, i64]),
    "sktt_svd_truncate": (i32, [vp, i32, i64, i64, vp, vp, vp, vp, dbl, i64, vp, pint, pint]),
    "sktt_eigh_jacobi": (i32, [vp, i32, i64, vp, vp, vp, pint]),
    "sktt_eig_si_work": (i64, [i64, i64, i64]),
    "sktt_eig_shift_invert": (i32, [vp, i32, i64, vp, vp, dbl, i64, i64, dbl, i32, vp, vp, vp, pint]),
    "sktt_batch_stack_left_op": (i32, [vp, i32] + [i64] * 7 + [vp] * 5 + [i32]),
    "sktt_batch_stack_right_op": (i32, [vp, i32] + [i64] * 7 + [vp] * 5),
    "sktt_batch_micro_matrix_als": (i32, [vp, i32] + [i64] * 7 + [vp] * 5),
    "sktt_batch_eig_work": (i64, [i64] * 4),
    "sktt_batch_eig_shift_invert": (i32, [vp, i32, i64, i64, vp, dbl, i64, i64, dbl, i32, vp, vp, vp, vp, vp]),
    "sktt_batch_svd_left": (i32, [vp, i64, i64, i64, i64, vp, i64, Idx2, Idx2, i32, vp, i64, i64, i64, i32]),
    "sktt_peer_alloc": (i32, [vp, i64, C.POINTER(vp)]),
    "sktt_peer_free": (i32, [vp, vp]),
    "sktt_peer_export": (i32, [vp, vp, C.c_char_p]),
    "sktt_peer_open": (i32, [vp, C.c_char_p, C.POINTER(vp)]),
    "sktt_peer_close": (i32, [vp, vp]),
    "sktt_peer_barrier": (i32, [vp, i32, i32, C.c_uint64, C.POINTER(vp), vp]),
    "sktt_sharded_matvec_work": (i64, [i64] * 7),
    "sktt_sharded_matvec": (i32, [vp, i32] + [i64] * 6 + [vp] * 4 + [i64, i64, vp, i32, C.POINTER(vp), vp]),
    "sktt_expm_small": (i32, [vp, i64, vp, dbl, dbl, vp]),
    "sktt_tt_matmul_core": (i32, [vp, i32] + [i64] * 7 + [vp] * 3),
    "sktt_arr_stack": (i32, [vp, i32, i64, i64, i64, i64, vp, vp, vp, vp]),
    "sktt_arr_micro_matrix": (i32, [vp, i64, i64, i64, i64, vp, vp, vp, vp]),
    "sktt_pinv_scale": (i32, [vp, i64, vp, dbl, vp]),
    "sktt_axpby": (i32, [vp, i32, i64, pdbl, vp, pdbl, vp, vp]),
    "sktt_nrm2": (i32, [vp, i32, i64, vp, pdbl]),
    "sktt_dotc": (i32, [vp, i32, i64, vp, vp, pdbl]),
    "sktt_widen": (i32, [vp, i64, vp, vp]),
}

_lib = None


def load():
    """Load the shared library (idempotent).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: run `python __graft_entry__.py` (nvcc, sm_100a) first. "
            "scikit_tt_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class SkttError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"sktt_b200 status {status}: {msg}")
        self.status = status
