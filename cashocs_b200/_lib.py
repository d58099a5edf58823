"""ctypes binding of the C-ABI declared in ``include/cashocs_b200.h``.

The CUDA library is the product: if ``libcashocs_b200.so`` is missing this module raises
loudly -- there is no CPU or PyTorch fallback for any hot-path operation.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

from cashocs_b200 import _exceptions

_HERE = os.path.dirname(os.path.abspath(__file__))
# CASHOCS_B200_LIB selects a tuning variant built by cashocs_b200.build.build(tag=...) (kernel experiments)
LIB_PATH = os.environ.get("CASHOCS_B200_LIB") or os.path.join(_HERE, "libcashocs_b200.so")

POLY_MAXT = 24
QUAD_MAXQ = 64


class cb_poly(C.Structure):
    _fields_ = [
        ("nterms", C.c_int32),
        ("pad", C.c_int32),
        ("coef", C.c_double * POLY_MAXT),
        ("ex", C.c_int8 * POLY_MAXT),
        ("ey", C.c_int8 * POLY_MAXT),
        ("ez", C.c_int8 * POLY_MAXT),
    ]


class cb_quad(C.Structure):
    _fields_ = [
        ("nq", C.c_int32),
        ("pad", C.c_int32),
        ("lam", (C.c_double * 4) * QUAD_MAXQ),
        ("w", C.c_double * QUAD_MAXQ),
    ]


class cb_mesh_view(C.Structure):
    _fields_ = [
        ("dim", C.c_int32),
        ("ncolours", C.c_int32),
        ("colour_off", C.c_void_p),
        ("cells", C.c_void_p),
        ("cell2nnz", C.c_void_p),
        ("coords", C.c_void_p),
        ("nv", C.c_int64),
        ("nrows_owned", C.c_int64),
        ("nnz", C.c_int64),
        ("cells_orig", C.c_void_p),
        ("n2c_ptr", C.c_void_p),
        ("n2c_idx", C.c_void_p),
        ("v2c_ptr", C.c_void_p),
        ("v2c_idx", C.c_void_p),
        ("cell_scratch", C.c_void_p),
        ("cell_geo", C.c_void_p),
        ("rowptr", C.c_void_p),
        ("col", C.c_void_p),
    ]


class cb_p2_view(C.Structure):
    _fields_ = [
        ("dim", C.c_int32),
        ("ndof_e", C.c_int32),
        ("nc", C.c_int64),
        ("nv", C.c_int64),
        ("ndofs", C.c_int64),
        ("nnz", C.c_int64),
        ("nrows_owned", C.c_int64),
        ("nv_owned", C.c_int64),
        ("nc_owned", C.c_int64),
        ("cells", C.c_void_p),
        ("cell_dofs", C.c_void_p),
        ("coords", C.c_void_p),
        ("rowptr", C.c_void_p),
        ("col", C.c_void_p),
        ("d2c_ptr", C.c_void_p),
        ("d2c_idx", C.c_void_p),
        ("ref_R", C.c_void_p),
        ("ref_M", C.c_void_p),
        ("ref_I", C.c_void_p),
        ("cell_scratch", C.c_void_p),
    ]


class cb_mat(C.Structure):
    _fields_ = [
        ("nrows", C.c_int64),
        ("ncols", C.c_int64),
        ("bs", C.c_int32),
        ("row_split", C.c_int32),
        ("rowptr", C.c_void_p),
        ("col", C.c_void_p),
        ("val", C.c_void_p),
        ("val32", C.c_void_p),
        ("cta_warps", C.c_int32),
        ("reserved", C.c_int32),
    ]


class cb_amg_level(C.Structure):
    _fields_ = [
        ("A", cb_mat),
        ("dinv", C.c_void_p),
        ("lmax", C.c_double),
        ("nc", C.c_int64),
        ("p_rowptr", C.c_void_p),
        ("p_col", C.c_void_p),
        ("p_val", C.c_void_p),
        ("pt_rowptr", C.c_void_p),
        ("pt_col", C.c_void_p),
        ("pt_val", C.c_void_p),
        ("x", C.c_void_p),
        ("b", C.c_void_p),
        ("t", C.c_void_p),
        ("d", C.c_void_p),
        ("dist", C.c_void_p),
        ("sendbuf", C.c_void_p),
        ("rep_dist", C.c_void_p),
        ("rep_off", C.c_int64),
        ("AP", cb_mat),
        ("coarse_gid", C.c_void_p),
        ("xc_local", C.c_void_p),
        ("rep_gather", C.c_void_p),
        ("rep_tmp", C.c_void_p),
    ]


class cb_amg(C.Structure):
    _fields_ = [
        ("nlevels", C.c_int32),
        ("smooth_degree", C.c_int32),
        ("coarse_degree", C.c_int32),
        ("p_block", C.c_int32),
        ("smooth_ratio", C.c_double),
        ("coarse_ratio", C.c_double),
        ("level_smooth_degree", C.c_int32),
        ("pad3", C.c_int32),
        ("levels", C.c_void_p),
        ("coarse_inv", C.c_void_p),
    ]


class cb_ksp_opts(C.Structure):
    _fields_ = [
        ("ksp_type", C.c_int32),
        ("pc_type", C.c_int32),
        ("rtol", C.c_double),
        ("atol", C.c_double),
        ("dtol", C.c_double),
        ("max_it", C.c_int32),
        ("cheb_degree", C.c_int32),
        ("cheb_ratio", C.c_double),
        ("check_every", C.c_int32),
        ("monitor", C.c_int32),
        ("profile_every", C.c_int32),
        ("use_graph", C.c_int32),
        ("amg", C.c_void_p),
        ("rbm_coords", C.c_void_p),
        ("rbm_einv", C.c_void_p),
        ("rbm_work", C.c_void_p),
        ("flexible", C.c_int32),
        ("pad2", C.c_int32),
    ]


class cb_ksp_result(C.Structure):
    _fields_ = [
        ("reason", C.c_int32),
        ("its", C.c_int32),
        ("rnorm", C.c_double),
        ("rnorm0", C.c_double),
        ("spmv_ms", C.c_double),
        ("spmv_profiled", C.c_int32),
        ("spmv_launches", C.c_int32),
    ]


P = C.c_void_p
I64 = C.c_int64
I32 = C.c_int
DBL = C.c_double

# name -> (restype, argtypes); every symbol declared in include/cashocs_b200.h
SIGNATURES = {
    "cb_last_error": (C.c_char_p, []),
    "cb_version": (I32, []),
    "cb_struct_size": (I64, [I32]),
    "cb_num_sms": (I32, []),
    "cb_launch_count": (C.c_longlong, []),
    "cb_v2c_build": (I32, [P, I64, I64, I32, P, P, P]),
    "cb_pattern_rowptr": (I32, [P, I64, I32, P, P, P, P]),
    "cb_pattern_fill": (I32, [P, I64, I64, I32, P, P, P, P, P, P]),
    "cb_colour_cells_host": (I32, [I64, I64, I32, P, P]),
    "cb_assemble_scalar": (I32, [P, C.POINTER(cb_mesh_view), DBL, DBL, DBL, C.POINTER(cb_poly),
                                 C.POINTER(cb_quad), DBL, P, P, P, P, P]),
    "cb_assemble_elasticity": (I32, [P, C.POINTER(cb_mesh_view), P, DBL, DBL, P, P, P]),
    "cb_assemble_shape_derivative_poisson": (I32, [P, C.POINTER(cb_mesh_view), P, P, C.POINTER(cb_poly),
                                                   C.POINTER(cb_poly), C.POINTER(cb_quad), DBL, P, P]),
    "cb_assemble_shape_derivative": (I32, [P, C.POINTER(cb_mesh_view), P, P, C.POINTER(cb_poly), C.POINTER(cb_poly),
                                           C.POINTER(cb_quad), DBL, DBL, DBL, DBL, P, I32, P, P]),
    "cb_assemble_cubic": (I32, [P, C.POINTER(cb_mesh_view), P, DBL, C.POINTER(cb_quad), P, P, P]),
    "cb_ident_zeros": (I32, [P, I64, I32, P, P, P, P]),
    "cb_p2_assemble_matrix": (I32, [P, C.POINTER(cb_p2_view), DBL, DBL, P, P]),
    "cb_p2_assemble_rhs": (I32, [P, C.POINTER(cb_p2_view), DBL, DBL, DBL, C.POINTER(cb_poly), I32, P, DBL, DBL, P,
                                 P, P, P]),
    "cb_p2_shape_derivative_poisson": (I32, [P, C.POINTER(cb_p2_view), P, P, P, P, C.POINTER(cb_poly),
                                             C.POINTER(cb_poly), I32, P, DBL, P, P]),
    "cb_p2_functional": (I32, [P, C.POINTER(cb_p2_view), DBL, P, DBL, P, P, P, I64]),
    "cb_spmv": (I32, [P, C.POINTER(cb_mat), P, P]),
    "cb_spmv_f32": (I32, [P, C.POINTER(cb_mat), P, P]),
    "cb_scalar_product": (I32, [P, C.POINTER(cb_mat), P, P, P, P, I64]),
    "cb_dot": (I32, [P, I64, P, P, P, P, I64]),
    "cb_reduce_scratch_len": (I64, []),
    "cb_extract_diag_inv": (I32, [P, C.POINTER(cb_mat), P, P, P, I64]),
    "cb_rbm_galerkin": (I32, [P, C.POINTER(cb_mat), P, P, P, P, P, P, P, I64]),
    "cb_rbm_invert": (I32, [P, I32, P, P]),
    "cb_ksp_workspace_len": (I64, [I64, I64, C.POINTER(cb_ksp_opts)]),
    "cb_ksp_solve": (I32, [P, C.POINTER(cb_mat), P, P, C.POINTER(cb_ksp_opts), P, P, I64,
                           C.POINTER(cb_ksp_result), P, I64]),
    "cb_amg_block_trace": (I32, [P, C.POINTER(cb_mat), P]),
    "cb_amg_smooth_prolongator": (I32, [P, C.POINTER(cb_mat), DBL, P, P, P, P, P]),
    "cb_amg_spgemm_terms": (I32, [P, I32, I32, I64, P, P, P, P, P, P]),
    "cb_amg_spgemm_rowwise": (I32, [P, I32, I32, I64, P, P, P, P, P, P, P, P, P]),
    "cb_amg_dense_inverse": (I32, [P, C.POINTER(cb_mat), P, P]),
    "cb_mesh_quality": (I32, [P, I32, I64, P, P, I32, P, P, P, I64]),
    "cb_det_check": (I32, [P, I32, I64, P, P, P, DBL, P, P, P, I64]),
    "cb_gradient_frobenius_max": (I32, [P, I32, I64, P, P, P, P, P, I64]),
    "cb_cell_volumes": (I32, [P, I32, I64, P, P, P, P, P, I64]),
    "cb_move_mesh": (I32, [P, I64, P, P, P, DBL]),
    "cb_functional": (I32, [P, I32, I64, P, P, DBL, P, DBL, P, P, P, I64]),
    "cb_to_f32": (I32, [P, I64, P, P]),
    "cb_power_step": (I32, [P, C.POINTER(cb_mat), P, DBL, P, P, P, P, I64, I32]),
    "cb_amg_rbm_aggregate_geometry": (I32, [P, I64, P, P, P, P, P, P, P]),
    "cb_amg_rbm_tentative": (I32, [P, I64, P, P, P, P, P, P, P, P]),
    "cb_amg_rbm_spgemm_terms": (I32, [P, I32, I64, P, P, P, P, P, P]),
    "cb_amg_rbm_prolongator_finish": (I32, [P, I64, P, I64, P, P, P, DBL, P]),
    "cb_amg_rbm_transpose_blocks": (I32, [P, I64, P, P, P]),
    "cb_amg_rbm_restrict": (I32, [P, I64, P, P, P, P, P, P]),
    "cb_amg_rbm_prolong": (I32, [P, I64, P, P, P, P, P]),
    "cb_boundary_measure": (I32, [P, I32, I64, P, P, P, P, I64]),
    "cb_boundary_measure_gradient": (I32, [P, I32, I64, I64, P, P, P, P, P, P]),
    "cb_plaplace_residual": (I32, [P, I32, I64, P, P, P, P, P, P, I32, DBL, DBL, P, P, P]),
    "cb_plaplace_jacobian": (I32, [P, I32, I64, P, P, P, P, P, P, I32, DBL, DBL, P, P, P, P]),
    "cb_plaplace_product": (I32, [P, I32, I64, P, P, P, P, P, I32, DBL, DBL, P, P, I64]),
    "cb_eikonal_rhs": (I32, [P, I32, I64, P, P, P, P, P, P, P]),
    "cb_eikonal_residual": (I32, [P, I32, I64, P, P, P, P, P, I64]),
    "cb_distance_mu": (I32, [P, I64, P, DBL, DBL, DBL, DBL, I32, P]),
    "cb_boundary_facet_opposite": (I32, [P, I32, I64, P, P, P, P, P]),
    "cb_boundary_normal_rhs": (I32, [P, I32, I32, I64, I64, P, P, P, P, P, P, P, P]),
    "cb_boundary_mass": (I32, [P, I32, I64, P, P, P, P, P, P, P, P]),
    "cb_boundary_curvature_derivative": (I32, [P, I32, I64, I64, P, P, P, P, P, P, P]),
    "cb_intersection_test": (I32, [P, I32, I64, I64, P, P, P, P, P]),
    "cb_dist_unique_id": (I32, [P]),
    "cb_dist_create": (I32, [C.POINTER(P), P, I32, I32, I32, P, P, P, P, I64, I64]),
    "cb_dist_clone_plan": (I32, [P, C.POINTER(P), I32, P, P, P, P, I64, I64]),
    "cb_dist_destroy": (I32, [P]),
    "cb_dist_p2p_export": (I32, [P, P]),
    "cb_dist_p2p_ghost_cap": (I64, [P]),
    "cb_dist_p2p_open": (I32, [P, P, P, P]),
    "cb_dist_p2p_disable": (I32, [P]),
    "cb_dist_p2p_error": (I32, [P]),
    "cb_dist_halo_exchange": (I32, [P, P, P, I32, P]),
    "cb_dist_allreduce": (I32, [P, P, P, I32, I32]),
}

_lib = None


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m cashocs_b200.build` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "cashocs_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    # a stale .so (built from an older header) would silently misread every struct: refuse it
    mirrors = [cb_poly, cb_quad, cb_mat, cb_mesh_view, cb_ksp_opts, cb_amg_level, cb_amg, cb_ksp_result, cb_p2_view]
    for which, mirror in enumerate(mirrors):
        if lib.cb_struct_size(which) != C.sizeof(mirror):
            raise ImportError(f"{LIB_PATH} is stale: sizeof({mirror.__name__}) is {lib.cb_struct_size(which)} in the library, "
                              f"{C.sizeof(mirror)} in the binding -- rebuild with `python -m cashocs_b200.build -f`")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> int:
    """Raise on a negative library return code."""
    if rc < 0:
        msg = load().cb_last_error().decode("utf-8", "replace")
        raise _exceptions.CashocsException(f"cashocs_b200 C-ABI error {rc} in {what}: {msg}")
    return rc


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


def stream_ptr():
    """The current torch CUDA stream as a void*."""
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    """Hot-path entry points accept CUDA fp64/int32/int64 contiguous tensors only."""
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _exceptions.CashocsException(
                "cashocs_b200: this operation runs on the GPU only (got a CPU tensor); "
                "there is no CPU fallback."
            )
        if not t.is_contiguous():
            raise _exceptions.CashocsException("cashocs_b200: tensors must be contiguous")
