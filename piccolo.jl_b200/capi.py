"""ctypes binding of include/piccolo_b200.h (the same symbols a Julia ``ccall`` shim binds)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PB2_LIB selects an alternative build of the same library (e.g. the -DPB2_TRACE debug build)
_SO = os.path.join(_HERE, os.environ.get("PB2_LIB", "libpiccolo_b200.so"))

PB2_KET, PB2_UNITARY, PB2_DENSITY = 0, 1, 2
PB2_HOST, PB2_DEVICE = 0, 1
PB2_ALG_AUTO, PB2_ALG_GENERIC, PB2_ALG_DMMA = 0, 1, 2
KIND = {"ket": PB2_KET, "unitary": PB2_UNITARY, "density": PB2_DENSITY}
ALG = {"auto": PB2_ALG_AUTO, "generic": PB2_ALG_GENERIC, "dmma": PB2_ALG_DMMA}
OPT = {"early_z": 1, "pipelined": 2, "hessian_ctas": 3}

# every symbol include/piccolo_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "pb2_version", "pb2_last_error", "pb2_device_count", "pb2_create", "pb2_destroy",
    "pb2_dim", "pb2_nnz_jac", "pb2_nnz_hess", "pb2_algorithm", "pb2_hessian_algorithm", "pb2_structure_jac",
    "pb2_structure_hess", "pb2_residual", "pb2_jacobian", "pb2_residual_jacobian",
    "pb2_hess_lagrangian", "pb2_residual_jacobian_async", "pb2_hess_lagrangian_async",
    "pb2_compact_stride", "pb2_residual_jacobian_compact_async", "pb2_expand_compact_async",
    "pb2_residual_jacobian_exchange_async", "pb2_residual_jacobian_exchange_sync_async", "pb2_enable_peer_access",
    "pb2_aux_create", "pb2_aux_destroy", "pb2_aux_dim", "pb2_aux_nnz_jac", "pb2_aux_nnz_hess",
    "pb2_aux_structure_jac", "pb2_aux_structure_hess", "pb2_aux_residual_jacobian", "pb2_aux_hess_lagrangian",
    "pb2_aux_residual_jacobian_async",
    "pb2_obj_create", "pb2_obj_destroy", "pb2_obj_value_gradient", "pb2_obj_value_gradient_async",
    "pb2_obj_nnz_hess", "pb2_obj_structure_hess", "pb2_obj_hessian", "pb2_obj_hessian_async",
    "pb2_stream", "pb2_sync", "pb2_set_option", "pb2_rollout", "pb2_rollout_async",
    "pb2_set_time_coefficients", "pb2_host_alloc", "pb2_host_free", "pb2_launch_count",
    "pb2_device_alloc", "pb2_device_free", "pb2_device_copy", "pb2_ipc_export", "pb2_ipc_open", "pb2_ipc_close",
    "pb2_batch_create", "pb2_batch_destroy", "pb2_batch_size", "pb2_batch_fused", "pb2_batch_dim",
    "pb2_batch_nnz_jac", "pb2_batch_nnz_hess", "pb2_batch_structure_jac", "pb2_batch_structure_hess",
    "pb2_batch_residual_jacobian", "pb2_batch_hess_lagrangian", "pb2_batch_residual_jacobian_async",
    "pb2_batch_hess_lagrangian_async",
]


class PB2Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libpiccolo_b200 error {code}: {msg}")
        self.code = code


class pb2_desc(ctypes.Structure):
    _fields_ = [
        ("kind", ctypes.c_int32), ("b", ctypes.c_int32), ("n_b", ctypes.c_int32),
        ("m", ctypes.c_int32), ("K", ctypes.c_int32), ("D", ctypes.c_int32),
        ("x_off", ctypes.c_int32), ("dt_off", ctypes.c_int32), ("u_off", ctypes.c_int32),
        ("global_dim", ctypes.c_int32), ("knot0", ctypes.c_int64), ("device", ctypes.c_int32),
        ("algorithm", ctypes.c_int32), ("G0", ctypes.POINTER(ctypes.c_double)),
        ("Gj", ctypes.POINTER(ctypes.c_double)), ("t_off", ctypes.c_int32), ("time_dependent", ctypes.c_int32),
        ("dense_blocks", ctypes.c_int32),
    ]


PB2_AUX_MAX_PAIRS = 8


class pb2_aux_desc(ctypes.Structure):
    _fields_ = [
        ("K", ctypes.c_int32), ("D", ctypes.c_int32), ("dt_off", ctypes.c_int32), ("t_off", ctypes.c_int32),
        ("global_dim", ctypes.c_int32), ("n_pairs", ctypes.c_int32),
        ("x_off", ctypes.c_int32 * PB2_AUX_MAX_PAIRS), ("xdot_off", ctypes.c_int32 * PB2_AUX_MAX_PAIRS),
        ("dim", ctypes.c_int32 * PB2_AUX_MAX_PAIRS), ("device", ctypes.c_int32),
        ("timesteps_all_equal", ctypes.c_int32),
    ]


class pb2_obj_term(ctypes.Structure):
    _fields_ = [
        ("flags", ctypes.c_int32), ("n_rows", ctypes.c_int32), ("rows", ctypes.POINTER(ctypes.c_int32)),
        ("a_re", ctypes.POINTER(ctypes.c_double)), ("a_im", ctypes.POINTER(ctypes.c_double)),
        ("a_sq", ctypes.POINTER(ctypes.c_double)), ("a_lin", ctypes.POINTER(ctypes.c_double)),
        ("scale", ctypes.c_double), ("n_times", ctypes.c_int32), ("times", ctypes.POINTER(ctypes.c_int32)),
        ("Q", ctypes.POINTER(ctypes.c_double)),
    ]


class pb2_obj_reg(ctypes.Structure):
    _fields_ = [
        ("n_rows", ctypes.c_int32), ("rows", ctypes.POINTER(ctypes.c_int32)), ("R", ctypes.POINTER(ctypes.c_double)),
        ("baseline", ctypes.POINTER(ctypes.c_double)), ("dt_power", ctypes.c_int32), ("n_times", ctypes.c_int32),
        ("times", ctypes.POINTER(ctypes.c_int32)),
    ]


class pb2_obj_desc(ctypes.Structure):
    _fields_ = [
        ("K", ctypes.c_int32), ("D", ctypes.c_int32), ("dt_off", ctypes.c_int32), ("n_terms", ctypes.c_int32),
        ("n_regs", ctypes.c_int32), ("terms", ctypes.POINTER(pb2_obj_term)), ("regs", ctypes.POINTER(pb2_obj_reg)),
        ("device", ctypes.c_int32),
    ]


_lib = None


def lib_path():
    return _SO


def load_library():
    """Load libpiccolo_b200.so; fail loudly (no fallback) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_SO):
        raise PB2Error(-1, f"{_SO} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C piccolo.jl_b200` (there is no CPU fallback)")
    L = ctypes.CDLL(_SO)
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)
    vp, H = ctypes.c_void_p, ctypes.c_void_p
    L.pb2_version.restype = ctypes.c_int
    L.pb2_last_error.restype = ctypes.c_char_p
    L.pb2_device_count.restype = ctypes.c_int
    L.pb2_create.argtypes = [ctypes.POINTER(pb2_desc), ctypes.POINTER(H)]
    L.pb2_destroy.argtypes = [H]
    L.pb2_destroy.restype = None
    for f in ("pb2_dim", "pb2_nnz_jac", "pb2_nnz_hess", "pb2_launch_count"):
        getattr(L, f).argtypes = [H]
        getattr(L, f).restype = ctypes.c_int64
    L.pb2_algorithm.argtypes = [H]
    L.pb2_algorithm.restype = ctypes.c_int32
    L.pb2_hessian_algorithm.argtypes = [H]
    L.pb2_hessian_algorithm.restype = ctypes.c_int32
    L.pb2_structure_jac.argtypes = [H, ip, ip]
    L.pb2_structure_hess.argtypes = [H, ip, ip]
    L.pb2_residual.argtypes = [H, vp, vp, ctypes.c_int]
    L.pb2_jacobian.argtypes = [H, vp, vp, ctypes.c_int]
    L.pb2_residual_jacobian.argtypes = [H, vp, vp, vp, ctypes.c_int]
    L.pb2_hess_lagrangian.argtypes = [H, vp, vp, vp, ctypes.c_int]
    L.pb2_residual_jacobian_async.argtypes = [H, vp, vp, vp, vp]
    L.pb2_hess_lagrangian_async.argtypes = [H, vp, vp, vp, vp]
    L.pb2_sync.argtypes = [H]
    L.pb2_compact_stride.argtypes = [H]
    L.pb2_compact_stride.restype = ctypes.c_int64
    L.pb2_residual_jacobian_compact_async.argtypes = [H, vp, vp, vp]
    L.pb2_expand_compact_async.argtypes = [H, vp, ctypes.c_int64, vp, vp, vp]
    L.pb2_enable_peer_access.argtypes = [ctypes.c_int32, ctypes.c_int32]
    L.pb2_residual_jacobian_exchange_async.argtypes = [H, vp, ctypes.c_int32, ctypes.c_int32,
                                                       ctypes.POINTER(vp), ctypes.c_int64, vp]
    L.pb2_set_option.argtypes = [H, ctypes.c_int32, ctypes.c_int64]
    L.pb2_device_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_int64, ctypes.c_int32]
    L.pb2_device_free.argtypes = [vp]
    L.pb2_device_copy.argtypes = [vp, vp, ctypes.c_int64]
    L.pb2_ipc_export.argtypes = [vp, vp]
    L.pb2_ipc_open.argtypes = [vp, ctypes.c_int32, ctypes.POINTER(vp)]
    L.pb2_ipc_close.argtypes = [vp]
    L.pb2_batch_create.argtypes = [ctypes.POINTER(pb2_desc), ctypes.c_int32, ctypes.POINTER(H)]
    L.pb2_batch_destroy.argtypes = [H]
    L.pb2_batch_destroy.restype = None
    for f in ("pb2_batch_size", "pb2_batch_fused"):
        getattr(L, f).argtypes = [H]
        getattr(L, f).restype = ctypes.c_int32
    for f in ("pb2_batch_dim", "pb2_batch_nnz_jac", "pb2_batch_nnz_hess"):
        getattr(L, f).argtypes = [H]
        getattr(L, f).restype = ctypes.c_int64
    L.pb2_batch_structure_jac.argtypes = [H, ctypes.c_int32, ip, ip]
    L.pb2_batch_structure_hess.argtypes = [H, ctypes.c_int32, ip, ip]
    L.pb2_batch_residual_jacobian.argtypes = [H, vp, vp, vp, ctypes.c_int]
    L.pb2_batch_hess_lagrangian.argtypes = [H, vp, vp, vp, ctypes.c_int]
    L.pb2_batch_residual_jacobian_async.argtypes = [H, vp, vp, vp, vp]
    L.pb2_batch_hess_lagrangian_async.argtypes = [H, vp, vp, vp, vp]
    L.pb2_set_time_coefficients.argtypes = [H, vp, vp, ctypes.c_int]
    L.pb2_rollout.argtypes = [H, vp, vp, vp, vp, ctypes.c_int]
    L.pb2_rollout_async.argtypes = [H, vp, vp, vp, vp, vp]
    L.pb2_residual_jacobian_exchange_sync_async.argtypes = [H, vp, ctypes.c_int32, ctypes.c_int32,
                                                            ctypes.POINTER(vp), ctypes.c_int64, ctypes.c_int64, vp]
    L.pb2_stream.argtypes = [H]
    L.pb2_stream.restype = ctypes.c_void_p
    L.pb2_aux_create.argtypes = [ctypes.POINTER(pb2_aux_desc), ctypes.POINTER(H)]
    L.pb2_aux_destroy.argtypes = [H]
    L.pb2_aux_destroy.restype = None
    for f in ("pb2_aux_dim", "pb2_aux_nnz_jac", "pb2_aux_nnz_hess"):
        getattr(L, f).argtypes = [H]
        getattr(L, f).restype = ctypes.c_int64
    L.pb2_aux_structure_jac.argtypes = [H, ip, ip]
    L.pb2_aux_structure_hess.argtypes = [H, ip, ip]
    L.pb2_aux_residual_jacobian.argtypes = [H, vp, vp, vp, ctypes.c_int]
    L.pb2_aux_hess_lagrangian.argtypes = [H, vp, vp, ctypes.c_int]
    L.pb2_aux_residual_jacobian_async.argtypes = [H, vp, vp, vp, vp]
    L.pb2_obj_create.argtypes = [ctypes.POINTER(pb2_obj_desc), ctypes.POINTER(H)]
    L.pb2_obj_destroy.argtypes = [H]
    L.pb2_obj_destroy.restype = None
    L.pb2_obj_value_gradient.argtypes = [H, vp, vp, vp, ctypes.c_int]
    L.pb2_obj_value_gradient_async.argtypes = [H, vp, vp, vp, vp]
    L.pb2_obj_nnz_hess.argtypes = [H]
    L.pb2_obj_nnz_hess.restype = ctypes.c_int64
    L.pb2_obj_structure_hess.argtypes = [H, ip, ip]
    L.pb2_obj_hessian.argtypes = [H, vp, ctypes.c_double, vp, ctypes.c_int]
    L.pb2_obj_hessian_async.argtypes = [H, vp, ctypes.c_double, vp, vp]
    L.pb2_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_int64]
    L.pb2_host_free.argtypes = [vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise PB2Error(rc, load_library().pb2_last_error().decode())
