"""ctypes binding of libpanda_validity.so (include/panda_validity.h) and its build recipe.

The library is built IN-TREE (csrc/libpanda_validity.so) for sm_100a only.  There is no CPU fallback:
`load()` raises if the shared object is missing, and `pv_create` fails on a machine without a B200.
"""
from __future__ import annotations

import ctypes as C
import operator
import os
import shutil
import subprocess
from typing import List

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libpanda_validity.so")
SOURCES = ["pv_kernels.cu", "pv_edge.cu", "pv_rrtc.cu", "pv_ik.cu", "pv_plan.cu", "pv_nn.cu"]
HEADERS = ["pv_device.cuh", "pv_handle.h", "panda_model_gen.h", os.path.join("..", "..", "include", "panda_validity.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--threads", "4",
]

# every symbol include/panda_validity.h declares
EXPORTS = [
    "pv_create", "pv_destroy", "pv_last_error", "pv_version", "pv_model_info", "pv_joint_limits",
    "pv_set_scene", "pv_set_attached", "pv_set_carried", "pv_set_flags", "pv_set_culling", "pv_set_launch_overlap", "pv_set_gather", "pv_fk", "pv_fk_verdict_path", "pv_check_states", "pv_state_margins", "pv_state_contacts",
    "pv_check_edges", "pv_edge_margins", "pv_check_states_host", "pv_check_states_host_arm", "pv_check_edges_host", "pv_sweep",
    "pv_rrtc_batch", "pv_rrtc_batch_packed", "pv_plan_path", "pv_interpolate_path", "pv_obb_from_poses", "pv_simplify_path",
    "pv_simplify_path_cb",
    "pv_nn_candidates", "pv_nn_candidates_gather", "pv_rrtc_steer", "pv_rrtc_samples", "pv_ik_batch", "pv_fp32_peak", "pv_launch_count",
]


PV_ERR_CAPACITY = -6


class PvRrtcParams(C.Structure):
    _fields_ = [
        ("range", C.c_float), ("resolution", C.c_float), ("max_iters", C.c_int), ("max_nodes", C.c_int),
        ("max_path", C.c_int), ("seed", C.c_uint32), ("replicas", C.c_int), ("shortcut_passes", C.c_int),
        ("check_endpoints", C.c_int), ("planner", C.c_int), ("query_offset", C.c_int),
    ]


class PvPlanParams(C.Structure):
    _fields_ = [
        ("range", C.c_float), ("resolution", C.c_float), ("max_iters", C.c_int), ("max_nodes", C.c_int),
        ("seed", C.c_uint32), ("replicas", C.c_int), ("smooth", C.c_int), ("planner", C.c_int), ("validate", C.c_int),
        ("max_attempts", C.c_int), ("timeout_s", C.c_double),
    ]


class PvPlanStats(C.Structure):
    _fields_ = [
        ("solved", C.c_int), ("endpoint_status", C.c_int), ("attempts", C.c_int), ("refinements", C.c_int),
        ("iters", C.c_int), ("checks", C.c_longlong), ("vertices_raw", C.c_int), ("vertices", C.c_int),
        ("partial_rounds", C.c_int), ("bspline_steps", C.c_int), ("reduce_rounds", C.c_int),
        ("simplify_motions", C.c_int), ("fallback_unsimplified", C.c_int), ("validated", C.c_int),
        ("speculative_hit", C.c_int), ("launches", C.c_int), ("ms_solve", C.c_float), ("ms_simplify", C.c_float),
        ("ms_post", C.c_float), ("ms_total", C.c_float),
    ]

    def as_dict(self) -> dict:
        return dict(zip(_PLAN_STAT_NAMES, _PLAN_STAT_GET(self)))


_PLAN_STAT_NAMES = tuple(name for name, _ in PvPlanStats._fields_)
_PLAN_STAT_GET = operator.attrgetter(*_PLAN_STAT_NAMES)  # one call for all fields (as_dict runs inside every plan_path)


EDGE_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int,
                            C.POINTER(C.c_ubyte))


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libpanda_validity.so")
    return exe


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into csrc/libpanda_validity.so (nvcc cross-compiles without a GPU)."""
    from . import panda_model

    panda_model.write_header()
    if not force and not needs_build():
        return LIB_PATH
    nvcc = _nvcc()
    objs: List[str] = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH] + objs + ["-lcudart"], cwd=CSRC, check=True)
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """Load the in-tree library and declare prototypes.  Raises (no fallback) if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "The validity path has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, fp, u32p, i32p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    lib.pv_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    lib.pv_destroy.argtypes = [vp]
    lib.pv_destroy.restype = None
    lib.pv_last_error.argtypes = [vp]
    lib.pv_last_error.restype = C.c_char_p
    lib.pv_version.restype = C.c_char_p
    lib.pv_model_info.argtypes = [C.POINTER(C.c_int)] * 4
    lib.pv_joint_limits.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.pv_set_scene.argtypes = [vp, C.POINTER(C.c_float), C.c_int, C.c_float, C.POINTER(C.c_float)]
    lib.pv_set_attached.argtypes = [vp, C.c_int]
    lib.pv_set_carried.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.c_float]
    lib.pv_set_flags.argtypes = [vp, C.c_uint]
    lib.pv_set_culling.argtypes = [vp, C.c_int]
    lib.pv_set_launch_overlap.argtypes = [vp, C.c_int]
    lib.pv_set_gather.argtypes = [vp, vp, C.c_int, vp, C.c_longlong, C.c_longlong]
    lib.pv_fk.argtypes = [vp, fp, fp, fp, C.c_int64, fp, vp]
    lib.pv_fk_verdict_path.argtypes = [vp, fp, fp, fp, C.c_int64, fp, vp]
    lib.pv_check_states.argtypes = [vp, fp, fp, fp, C.c_int64, u32p, vp]
    lib.pv_state_margins.argtypes = [vp, fp, fp, fp, C.c_int64, fp, i32p, vp]
    lib.pv_state_contacts.argtypes = [vp, fp, fp, fp, C.c_int64, i32p, i32p, vp]
    lib.pv_check_edges.argtypes = [vp, fp, fp, fp, fp, fp, fp, C.c_int64, C.c_int, C.c_float, u32p, vp]
    lib.pv_edge_margins.argtypes = [vp, fp, fp, fp, fp, fp, fp, C.c_int64, C.c_int, C.c_float, fp, vp]
    lib.pv_check_states_host.argtypes = [vp, fp, C.c_int64, u32p]
    lib.pv_check_edges_host.argtypes = [vp, fp, fp, C.c_int64, C.c_int, C.c_float, u32p]
    lib.pv_sweep.argtypes = [vp, C.c_uint64, C.c_int64, C.c_uint32, C.c_int, u32p, vp, fp, vp]
    lib.pv_rrtc_batch.argtypes = [vp, fp, fp, C.c_int, C.POINTER(PvRrtcParams), fp, i32p, i32p, vp]
    lib.pv_rrtc_batch_packed.argtypes = [vp, fp, fp, C.c_int, C.POINTER(PvRrtcParams), fp, C.c_longlong, vp, i32p, i32p, vp, vp]
    lib.pv_plan_path.argtypes = [vp, vp, vp, C.c_int, C.POINTER(PvPlanParams), fp, C.c_int, C.POINTER(C.c_int),
                                 C.POINTER(PvPlanStats)]
    lib.pv_interpolate_path.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int, C.POINTER(C.c_int)]
    lib.pv_check_states_host_arm.argtypes = [vp, vp, C.c_int64, C.c_float, C.c_float, vp]
    lib.pv_obb_from_poses.argtypes = [vp, vp, vp, C.c_int, vp]
    lib.pv_simplify_path.argtypes = [vp, vp, C.c_int, C.c_uint32, C.c_float, vp, C.c_int, C.POINTER(C.c_int), vp]
    lib.pv_simplify_path_cb.argtypes = [vp, C.c_int, C.c_uint32, EDGE_CALLBACK, vp, vp, C.c_int, C.POINTER(C.c_int), vp]
    lib.pv_nn_candidates.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
    lib.pv_nn_candidates_gather.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, vp]
    lib.pv_rrtc_steer.argtypes = [vp, vp, C.c_int, C.c_int, vp, C.c_float, vp, vp, vp, vp, vp]
    lib.pv_rrtc_samples.argtypes = [vp, C.c_uint32, vp, vp, C.c_int, vp, vp]
    lib.pv_ik_batch.argtypes = [vp, fp, fp, C.c_int, fp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_uint32, fp, i32p, fp]
    lib.pv_fp32_peak.argtypes = [vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    lib.pv_launch_count.argtypes = [vp]
    lib.pv_launch_count.restype = C.c_longlong
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int or name in ("pv_create",):
            fn.restype = C.c_int
    _lib = lib
    return lib
