"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/libep_oracle.so (the CPU restatement).
Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes as C
import os
import subprocess

import numpy as np

from elm_physics_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libep_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.ep_oracle_step.restype = C.c_int
    return _LIB


def step(shapes, bodies, params, warm=None, out=None, n_steps=1):
    """In-place step of `bodies` (abi.Bodies); returns the error code."""
    rc = lib().ep_oracle_step(C.byref(shapes.c), C.byref(bodies.c), C.byref(params.c),
                              C.byref(warm.c) if warm is not None else None,
                              C.byref(out.c) if out is not None else None, C.c_int32(n_steps))
    if rc != 0:
        raise RuntimeError(f"ep_oracle_step failed: {rc}")
    return rc


def _p(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def shape_contacts(shapes, s1, t1, s2, t2, shape_key=0, cap=64):
    """Collision.<A><B>.addContacts for world shapes placeIn t1 s1 / placeIn t2 s2; list order of the reference."""
    fk = np.zeros(cap, np.int64)
    ni, pi, pj = np.zeros((cap, 3)), np.zeros((cap, 3)), np.zeros((cap, 3))
    p1, pp1 = _p(t1[0]); q1, pq1 = _p(t1[1]); p2, pp2 = _p(t2[0]); q2, pq2 = _p(t2[1])
    n = lib().ep_oracle_shape_contacts(C.byref(shapes.c), s1, pp1, pq1, s2, pp2, pq2, shape_key, cap,
                                       fk.ctypes.data_as(C.POINTER(C.c_int64)), ni.ctypes.data_as(C.POINTER(C.c_double)),
                                       pi.ctypes.data_as(C.POINTER(C.c_double)), pj.ctypes.data_as(C.POINTER(C.c_double)))
    if n < 0:
        raise RuntimeError(f"ep_oracle_shape_contacts failed: {n}")
    return [dict(feature_key=int(fk[k]), ni=tuple(ni[k]), pi=tuple(pi[k]), pj=tuple(pj[k])) for k in range(n)]


def test_separating_axis(shapes, s1, t1, s2, t2, axis):
    p1, pp1 = _p(t1[0]); q1, pq1 = _p(t1[1]); p2, pp2 = _p(t2[0]); q2, pq2 = _p(t2[1]); a, pa = _p(axis)
    d = C.c_double(0)
    ok = lib().ep_oracle_test_separating_axis(C.byref(shapes.c), s1, pp1, pq1, s2, pp2, pq2, pa, C.byref(d))
    return d.value if ok else None


def find_separating_axis(shapes, s1, t1, s2, t2):
    p1, pp1 = _p(t1[0]); q1, pq1 = _p(t1[1]); p2, pp2 = _p(t2[0]); q2, pq2 = _p(t2[1])
    out = np.zeros(3)
    ok = lib().ep_oracle_find_separating_axis(C.byref(shapes.c), s1, pp1, pq1, s2, pp2, pq2, out.ctypes.data_as(C.POINTER(C.c_double)))
    return tuple(out) if ok else None


def support_feature(shapes, s, t, axis):
    p, pp = _p(t[0]); q, pq = _p(t[1]); a, pa = _p(axis)
    out = np.zeros(6)
    n = lib().ep_oracle_support_feature(C.byref(shapes.c), s, pp, pq, pa, out.ctypes.data_as(C.POINTER(C.c_double)))
    return [tuple(out[3 * i:3 * i + 3]) for i in range(n)]


def project(shapes, s, t, axis):
    """Collision.ConvexConvex.project over Convex.convexVertices of the placed hull: (min, max)."""
    p, pp = _p(t[0]); q, pq = _p(t[1]); a, pa = _p(axis)
    out = np.zeros(2)
    lib().ep_oracle_project(C.byref(shapes.c), s, pp, pq, pa, out.ctypes.data_as(C.POINTER(C.c_double)))
    return float(out[0]), float(out[1])


def raycast(shapes, bodies, ray_world, origins, directions):
    """ep_oracle_raycast: Physics.raycast (src/Physics.elm:802-841) against the current state of `bodies`."""
    rw = np.ascontiguousarray(ray_world, dtype=np.int32)
    o = np.ascontiguousarray(origins, dtype=np.float64).reshape(-1, 3)
    dr = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
    n = len(rw)
    hit, dist, point, normal = np.zeros(n, np.int32), np.zeros(n), np.zeros((n, 3)), np.zeros((n, 3))
    pf, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    rc = lib().ep_oracle_raycast(C.byref(shapes.c), C.byref(bodies.c), C.c_int32(n), rw.ctypes.data_as(pi), o.ctypes.data_as(pf),
                                 dr.ctypes.data_as(pf), hit.ctypes.data_as(pi), dist.ctypes.data_as(pf),
                                 point.ctypes.data_as(pf), normal.ctypes.data_as(pf))
    if rc != 0:
        raise RuntimeError(f"ep_oracle_raycast failed: {rc}")
    return hit, dist, point, normal


def apply(bodies, kind, body_row, a, b):
    """ep_oracle_apply: body mutators (src/Physics.elm:857-1017) on the host arrays of `bodies`, in list order."""
    k = np.ascontiguousarray(kind, dtype=np.int32); r = np.ascontiguousarray(body_row, dtype=np.int32)
    va = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 3); vb = np.ascontiguousarray(b, dtype=np.float64).reshape(-1, 3)
    pf, pi = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    rc = lib().ep_oracle_apply(C.byref(bodies.c), C.c_int32(len(k)), k.ctypes.data_as(pi), r.ctypes.data_as(pi), va.ctypes.data_as(pf), vb.ctypes.data_as(pf))
    if rc != 0:
        raise RuntimeError(f"ep_oracle_apply failed: {rc}")
