"""ctypes front end of the C oracle (oracle/panda_oracle.c).  Test infrastructure, NOT product code:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libpanda_oracle.so")

FLAG_SELF = 1
FLAG_LIMITS = 2
FLAG_CARRY = 4


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("panda_oracle.c", "panda_oracle_impl.h", "rrtc_oracle_impl.h")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return LIB_PATH


def _model_struct(real):
    class Model(C.Structure):
        _fields_ = [
            ("n_spheres", C.c_int), ("sphere_link", C.POINTER(C.c_int)),
            ("sphere_center", C.POINTER(real)), ("sphere_radius", C.POINTER(real)),
            ("n_boxes", C.c_int), ("box_link", C.POINTER(C.c_int)),
            ("box_center", C.POINTER(real)), ("box_half", C.POINTER(real)),
            ("n_ss", C.c_int), ("ss_pairs", C.POINTER(C.c_int)),
            ("n_sb", C.c_int), ("sb_pairs", C.POINTER(C.c_int)),
            ("q_lower", C.POINTER(real)), ("q_upper", C.POINTER(real)),
        ]
    return Model


class COracle:
    """precision: 'f64' or 'f32'."""

    def __init__(self, model: dict, precision: str = "f64"):
        self.lib = C.CDLL(build())
        self.sfx = "_" + precision
        self.np_t = np.float64 if precision == "f64" else np.float32
        self.c_t = C.c_double if precision == "f64" else C.c_float
        self._keep = {}
        Model = _model_struct(self.c_t)
        m = Model()

        def arr(name, dtype):
            a = np.ascontiguousarray(model[name], dtype=dtype)
            self._keep[name] = a
            return a

        def fp(a):
            return a.ctypes.data_as(C.POINTER(self.c_t))

        def ip(a):
            return a.ctypes.data_as(C.POINTER(C.c_int))

        m.n_spheres = len(model["sphere_link"])
        m.sphere_link = ip(arr("sphere_link", np.int32))
        m.sphere_center = fp(arr("sphere_center", self.np_t))
        m.sphere_radius = fp(arr("sphere_radius", self.np_t))
        m.n_boxes = len(model["box_link"])
        m.box_link = ip(arr("box_link", np.int32))
        m.box_center = fp(arr("box_center", self.np_t))
        m.box_half = fp(arr("box_half", self.np_t))
        m.n_ss = len(model["ss_pairs"])
        m.ss_pairs = ip(arr("ss_pairs", np.int32))
        m.n_sb = len(model["sb_pairs"])
        m.sb_pairs = ip(arr("sb_pairs", np.int32))
        m.q_lower = fp(arr("q_lower", self.np_t))
        m.q_upper = fp(arr("q_upper", self.np_t))
        self.model = m

    def _p(self, a):
        return a.ctypes.data_as(C.POINTER(self.c_t))

    def _scene(self, scene, attached, flags):
        """(obb array, n_obb, attached, flags); a scene["carried"] entry (see panda_oracle.state_margin) becomes the
        extra record obb[n_obb] plus FLAG_CARRY, with `attached` = the carried box."""
        obb = np.ascontiguousarray(scene["obb"], dtype=self.np_t).reshape(-1, 16)
        n = obb.shape[0]
        carried = scene.get("carried")
        if carried is not None:
            k = int(carried["index"])
            rec = np.zeros((1, 16), dtype=self.np_t)
            rec[0, 0:3] = np.asarray(carried["t"], dtype=np.float64)
            rec[0, 3:6] = obb[k, 3:6] - self.np_t(carried.get("shrink", 0.0))
            rec[0, 6:15] = np.asarray(carried["R"], dtype=np.float64).reshape(9)
            obb = np.ascontiguousarray(np.concatenate([obb, rec], axis=0))
            attached, flags = k, flags | FLAG_CARRY
        return obb, n, attached, flags

    def fk(self, q, base=(0.0, 0.0, 0.01)):
        q = np.ascontiguousarray(np.atleast_2d(q), dtype=self.np_t)
        n = q.shape[0]
        R = np.empty((n, 11, 3, 3), dtype=self.np_t)
        p = np.empty((n, 11, 3), dtype=self.np_t)
        b = np.asarray(base, dtype=self.np_t)
        getattr(self.lib, "po_fk" + self.sfx)(self._p(q), C.c_long(n), self._p(b), self._p(R), self._p(p))
        return R, p

    def state_margin(self, q, scene, attached=-1, flags=FLAG_SELF, base=(0.0, 0.0, 0.01), nthreads=0):
        q = np.ascontiguousarray(np.atleast_2d(q), dtype=self.np_t)
        n = q.shape[0]
        obb, n_obb, attached, flags = self._scene(scene, attached, flags)
        out = np.empty(n, dtype=self.np_t)
        b = np.asarray(base, dtype=self.np_t)
        getattr(self.lib, "po_state_margin" + self.sfx)(
            C.byref(self.model), self._p(obb), C.c_int(n_obb), self.c_t(scene["table_z"]), self._p(b),
            C.c_int(attached), C.c_int(flags), self._p(q), C.c_long(n), self._p(out),
            C.c_int(nthreads or os.cpu_count() or 1))
        return out

    def edge_margin(self, qa, qb, scene, n_steps=0, resolution=0.13037159046356686, attached=-1,
                    flags=FLAG_SELF, base=(0.0, 0.0, 0.01), early_exit=False, nthreads=0, return_count=False):
        qa = np.ascontiguousarray(np.atleast_2d(qa), dtype=self.np_t)
        qb = np.ascontiguousarray(np.atleast_2d(qb), dtype=self.np_t)
        n = qa.shape[0]
        obb, n_obb, attached, flags = self._scene(scene, attached, flags)
        out = np.empty(n, dtype=self.np_t)
        b = np.asarray(base, dtype=self.np_t)
        cnt = C.c_long(0)
        getattr(self.lib, "po_edge_margin" + self.sfx)(
            C.byref(self.model), self._p(obb), C.c_int(n_obb), self.c_t(scene["table_z"]), self._p(b),
            C.c_int(attached), C.c_int(flags), self._p(qa), self._p(qb), C.c_long(n), C.c_int(n_steps),
            self.c_t(resolution), C.c_int(1 if early_exit else 0), self._p(out), C.byref(cnt),
            C.c_int(nthreads or os.cpu_count() or 1))
        return (out, cnt.value) if return_count else out

    def rrtc(self, start, goal, scene, seed=1, search=0, max_iters=2000, max_nodes=2048, max_path=128,
             shortcut_passes=2, rrt_range=2.6074318092713376, resolution=0.13037159046356686, attached=-1,
             flags=FLAG_SELF, base=(0.0, 0.0, 0.01), planner="RRTConnect"):
        """CPU restatement of one device RRT-Connect search (fp32 only).  Returns (path (len, 9), iters, checks)."""
        assert self.np_t is np.float32, "the planner restatement mirrors the device arithmetic: use precision 'f32'"
        s = np.ascontiguousarray(start, dtype=np.float32).reshape(9)
        g = np.ascontiguousarray(goal, dtype=np.float32).reshape(9)
        obb, n_obb, attached, flags = self._scene(scene, attached, flags)
        b = np.asarray(base, dtype=np.float32)
        path = np.zeros((max_path, 9), dtype=np.float32)
        iters, checks = C.c_int(0), C.c_longlong(0)
        fn = self.lib.po_rrtc_f32
        fn.restype = C.c_int
        n = fn(C.byref(self.model), self._p(obb), C.c_int(n_obb), C.c_float(scene["table_z"]), self._p(b),
               C.c_int(attached), C.c_int(flags), self._p(s), self._p(g), C.c_float(np.float32(rrt_range)),
               C.c_float(np.float32(resolution)), C.c_int(max_iters), C.c_int(max_nodes), C.c_int(max_path),
               C.c_uint32(seed & 0xFFFFFFFF), C.c_uint32(search), C.c_int(shortcut_passes),
               C.c_int({"RRTConnect": 0, "RRT": 1}[planner]), self._p(path),
               C.byref(iters), C.byref(checks))
        return path[:n].copy(), iters.value, checks.value

    def edge_callback(self, scene, attached=-1, flags=FLAG_SELF, base=(0.0, 0.0, 0.01), resolution=0.13037159046356686):
        """(function pointer, context pointer, keep-alive) for pv_simplify_path_cb: this oracle as the motion validator
        of the product's simplifier, in C on both sides (fp32 only)."""
        assert self.np_t is np.float32

        class Ctx(C.Structure):
            _fields_ = [("m", C.c_void_p), ("obb", C.c_void_p), ("n_obb", C.c_int), ("table_z", C.c_float),
                        ("base", C.c_void_p), ("attached", C.c_int), ("flags", C.c_int), ("resolution", C.c_float),
                        ("motions", C.c_longlong), ("states", C.c_longlong)]

        obb, n_obb, attached, flags = self._scene(scene, attached, flags)
        b = np.ascontiguousarray(base, dtype=np.float32)
        ctx = Ctx(C.addressof(self.model), obb.ctypes.data, n_obb, float(scene["table_z"]), b.ctypes.data, attached, flags,
                  float(np.float32(resolution)), 0, 0)
        fn = C.cast(self.lib.po_edge_callback_f32, C.c_void_p)
        return fn, ctx, (obb, b)
