"""ctypes binding of ``libreyna_b200.so`` (the C-ABI declared in ``include/reyna_b200.h``).

There is NO CPU fallback: if the CUDA library is missing or a CUDA device is absent, the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("REYNA_B200_LIB") or os.path.join(_HERE, "libreyna_b200.so")   # override: A/B builds
CSRC = os.path.join(_HERE, "csrc")

RB_MAX_QV = 16
RB_MAX_QE = 4

EXPORTS = (
    "reyna_b200_structure_workspace", "reyna_b200_structure", "reyna_b200_structure_known", "reyna_b200_csr_pattern", "reyna_b200_csr_pattern_local",
    "reyna_b200_nccl_unique_id", "reyna_b200_comm_create", "reyna_b200_comm_destroy", "reyna_b200_dist_create", "reyna_b200_dist_destroy", "reyna_b200_dist_counters",
    "reyna_b200_nccl_halo", "reyna_b200_nccl_allreduce",
    "reyna_b200_geometry_tables_bytes", "reyna_b200_geometry_tables", "reyna_b200_assemble_workspace", "reyna_b200_assemble", "reyna_b200_assemble_prepare", "reyna_b200_assemble_range",
    "reyna_b200_assemble_volume_range", "reyna_b200_assemble_faces_range",
    "reyna_b200_boundary", "reyna_b200_compact_workspace", "reyna_b200_compact_count", "reyna_b200_compact_fill",
    "reyna_b200_mg_create", "reyna_b200_mg_destroy", "reyna_b200_mg_info", "reyna_b200_mg_apply",
    "reyna_b200_solve_workspace", "reyna_b200_solve", "reyna_b200_spmv", "reyna_b200_errors_workspace",
    "reyna_b200_errors", "reyna_b200_evaluate", "reyna_b200_fma_probe", "reyna_b200_dmma_probe",
    "reyna_b200_launch_count", "reyna_b200_last_error", "reyna_b200_version",
)


class RbGeom(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_nodes", "n_elem", "n_rows", "n_tri", "n_iface", "n_bface")] + \
               [(n, C.c_void_p) for n in ("nodes", "tri", "tri_elem", "tri_off", "bbox", "area", "poly_off", "poly_vtx",
                                          "iedge", "ielem", "inorm", "bedge", "belem", "bnorm", "elem_gid",
                                          "tri_xy", "vround_tri0", "tri_seg", "elem_seg0")] + \
               [(n, C.c_int32) for n in ("n_vrounds", "n_vsegs")] + [("tables", C.c_void_p)]


class RbQuad(C.Structure):
    _fields_ = [("p", C.c_int32), ("wv", C.c_double * RB_MAX_QV), ("rv", (C.c_double * 2) * RB_MAX_QV),
                ("we", C.c_double * RB_MAX_QE), ("xe", C.c_double * RB_MAX_QE)]


class RbCoefs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("vol_a", "vol_b", "vol_c", "vol_f", "if_a", "if_b", "if_a_mid", "if_upwind",
                                          "bf_a", "bf_b", "bf_g", "bf_a_mid")]


class RbErrSamples(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("vol_u", "vol_gu", "vol_a", "vol_c0", "if_a_mid", "if_b", "bf_u", "bf_a_mid",
                                          "bf_b", "bf_flags")]


class RbStructure(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("adj_off", "adj_item", "adj_nb", "grp_off", "counts")] + \
               [(n, C.c_int32) for n in ("n_items", "n_groups", "n_dup_items", "max_items")]


class RbMgOpts(C.Structure):
    _fields_ = [("nu", C.c_int32), ("omega", C.c_double), ("alpha", C.c_double), ("elem_bbox", C.c_void_p),
                ("fp32_levels", C.c_int32)]


class RbMgDist(C.Structure):
    _fields_ = [("dist", C.c_void_p), ("rank", C.c_int32), ("world", C.c_int32), ("n_ghost_elem", C.c_int32),
                ("n_global_elem", C.c_int32), ("bounds", C.c_void_p), ("p0_indptr", C.c_void_p), ("p0_indices", C.c_void_p),
                ("p0_data", C.c_void_p), ("p0_nnz", C.c_longlong),
                ("l1_indptr", C.c_void_p), ("l1_indices", C.c_void_p), ("l1_data", C.c_void_p), ("l1_nnz", C.c_longlong),
                ("l1_bbox", C.c_void_p), ("l0_tr", C.c_void_p)]


class RbSolveOpts(C.Structure):
    _fields_ = [("method", C.c_int32), ("max_iter", C.c_int32), ("rtol", C.c_double), ("check_every", C.c_int32),
                ("accept_rtol", C.c_double), ("sweep_elems", C.c_void_p), ("sweep_level_off", C.c_void_p),
                ("sweep_levels", C.c_int32)]


class RbSolveInfo(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("rel_residual", C.c_double), ("converged", C.c_int32),
                ("restarts", C.c_int32), ("stagnated", C.c_int32), ("singular_blocks", C.c_int32)]


HALO_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)


class ReynaB200Error(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU) and return the .so path."""
    res = subprocess.run(["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))], capture_output=True, text=True)
    if res.returncode != 0:
        raise ReynaB200Error("building libreyna_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (building it first if it is absent).  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    try:
        L = C.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the machine
        raise ReynaB200Error(f"cannot load {LIB_PATH}: {exc}") from exc
    vp, i32, i64p = C.c_void_p, C.c_int32, C.c_void_p
    L.reyna_b200_structure_workspace.restype = C.c_size_t
    L.reyna_b200_structure_workspace.argtypes = [i32, i32]
    L.reyna_b200_structure.restype = C.c_int
    L.reyna_b200_structure.argtypes = [C.POINTER(RbGeom), C.c_int, vp, C.POINTER(RbStructure), vp, C.c_size_t, vp]
    L.reyna_b200_structure_known.restype = C.c_int
    L.reyna_b200_structure_known.argtypes = [C.POINTER(RbGeom), C.c_int, vp, C.POINTER(RbStructure), vp, C.c_size_t, vp]
    L.reyna_b200_csr_pattern.restype = C.c_int
    L.reyna_b200_csr_pattern.argtypes = [C.POINTER(RbGeom), C.POINTER(RbStructure), C.c_int, vp, vp, vp]
    L.reyna_b200_csr_pattern_local.restype = C.c_int
    L.reyna_b200_csr_pattern_local.argtypes = [C.POINTER(RbGeom), C.POINTER(RbStructure), C.c_int, vp, vp, vp]
    L.reyna_b200_nccl_unique_id.restype = C.c_int
    L.reyna_b200_nccl_unique_id.argtypes = [C.c_char_p, C.c_char_p]
    L.reyna_b200_dist_create.restype = C.c_int
    L.reyna_b200_dist_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), vp, C.POINTER(vp)]
    L.reyna_b200_comm_create.restype = C.c_int
    L.reyna_b200_comm_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
    L.reyna_b200_comm_destroy.restype = None
    L.reyna_b200_comm_destroy.argtypes = [vp]
    L.reyna_b200_dist_destroy.restype = None
    L.reyna_b200_dist_destroy.argtypes = [vp]
    L.reyna_b200_dist_counters.restype = C.c_int
    L.reyna_b200_dist_counters.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    L.reyna_b200_assemble.restype = C.c_int
    L.reyna_b200_assemble.argtypes = [C.POINTER(RbGeom), C.POINTER(RbStructure), C.POINTER(RbCoefs), C.POINTER(RbQuad),
                                      C.c_double, vp, vp, i64p, vp, C.c_size_t, vp]
    L.reyna_b200_geometry_tables_bytes.restype = C.c_size_t
    L.reyna_b200_geometry_tables_bytes.argtypes = [i32, i32]
    L.reyna_b200_geometry_tables.restype = C.c_int
    L.reyna_b200_geometry_tables.argtypes = [C.POINTER(RbGeom), vp, C.c_size_t, vp]
    L.reyna_b200_assemble_prepare.restype = C.c_int
    L.reyna_b200_assemble_prepare.argtypes = [C.POINTER(RbGeom), C.POINTER(RbStructure), C.POINTER(RbCoefs), C.POINTER(RbQuad),
                                              C.c_double, vp, C.c_size_t, vp]
    L.reyna_b200_assemble_range.restype = C.c_int
    L.reyna_b200_assemble_range.argtypes = [C.POINTER(RbGeom), C.POINTER(RbStructure), C.POINTER(RbCoefs), C.POINTER(RbQuad),
                                            C.c_double, vp, vp, i64p, vp, C.c_size_t, i32, i32, i32, i32, vp]
    for fn in (L.reyna_b200_assemble_volume_range, L.reyna_b200_assemble_faces_range):
        fn.restype = C.c_int
        fn.argtypes = L.reyna_b200_assemble_range.argtypes
    L.reyna_b200_assemble_workspace.restype = C.c_size_t
    L.reyna_b200_assemble_workspace.argtypes = [i32, i32, i32, i32]
    L.reyna_b200_boundary.restype = C.c_int
    L.reyna_b200_boundary.argtypes = [C.POINTER(RbGeom), C.POINTER(RbStructure), C.POINTER(RbCoefs), C.POINTER(RbQuad),
                                      C.c_double, i32, vp, vp, vp, vp, vp, vp, vp, i64p, vp]
    L.reyna_b200_compact_workspace.restype = C.c_size_t
    L.reyna_b200_compact_workspace.argtypes = [i32]
    L.reyna_b200_compact_count.restype = C.c_int
    L.reyna_b200_compact_count.argtypes = [i32, vp, vp, vp, vp, C.c_size_t, vp]
    L.reyna_b200_compact_fill.restype = C.c_int
    L.reyna_b200_compact_fill.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp]
    L.reyna_b200_solve_workspace.restype = C.c_size_t
    L.reyna_b200_solve_workspace.argtypes = [i32, i32, i32, C.c_longlong]
    L.reyna_b200_solve.restype = C.c_int
    L.reyna_b200_solve.argtypes = [i32, i32, i32, C.c_int, vp, vp, vp, vp, vp, C.POINTER(RbSolveOpts),
                                   C.POINTER(RbSolveInfo), vp, vp, vp, vp, vp, C.c_size_t, vp]
    L.reyna_b200_mg_create.restype = C.c_int
    L.reyna_b200_mg_create.argtypes = [i32, i32, i32, vp, vp, vp, C.POINTER(RbMgOpts), C.POINTER(RbMgDist), C.POINTER(vp), vp]
    L.reyna_b200_mg_destroy.restype = None
    L.reyna_b200_mg_destroy.argtypes = [vp]
    L.reyna_b200_mg_info.restype = C.c_int
    L.reyna_b200_mg_info.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_longlong)]
    L.reyna_b200_mg_apply.restype = C.c_int
    L.reyna_b200_mg_apply.argtypes = [vp, vp, vp, vp]
    L.reyna_b200_spmv.restype = C.c_int
    L.reyna_b200_spmv.argtypes = [i32, i32, C.c_int, vp, vp, vp, vp, vp, vp]
    L.reyna_b200_errors_workspace.restype = C.c_size_t
    L.reyna_b200_errors_workspace.argtypes = []
    L.reyna_b200_errors.restype = C.c_int
    L.reyna_b200_errors.argtypes = [C.POINTER(RbGeom), C.POINTER(RbQuad), C.POINTER(RbErrSamples), C.c_double, vp, vp,
                                    vp, C.c_size_t, vp]
    L.reyna_b200_evaluate.restype = C.c_int
    L.reyna_b200_evaluate.argtypes = [C.c_int, i32, vp, vp, vp, vp, vp, vp]
    L.reyna_b200_fma_probe.restype = C.c_int
    L.reyna_b200_fma_probe.argtypes = [C.c_int, vp, C.POINTER(C.c_double), vp]
    L.reyna_b200_dmma_probe.restype = C.c_int
    L.reyna_b200_dmma_probe.argtypes = [C.c_int, vp, C.POINTER(C.c_double), vp]
    L.reyna_b200_launch_count.restype = C.c_ulonglong
    L.reyna_b200_launch_count.argtypes = []
    L.reyna_b200_last_error.restype = C.c_char_p
    L.reyna_b200_version.restype = C.c_char_p
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().reyna_b200_last_error().decode("utf-8", "replace")
        raise ReynaB200Error(f"{what} failed (code {rc}): {msg}")


def require_cuda():
    """The product path has no CPU implementation: fail loudly when there is no CUDA device."""
    import torch
    if not torch.cuda.is_available():
        raise ReynaB200Error("reyna_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch
