"""ctypes mirror of include/elm_physics_b200.h (struct layouts + numpy-backed holders).

Pure marshalling: every holder owns numpy arrays and exposes a ctypes struct whose pointers
alias them.  Used by the product binding (engine.py) and, with the same structs, by the
oracle binding that lives under oracle/ (tests only).
"""
import ctypes as C

import numpy as np

EP_OK, EP_EINVAL, EP_ECUDA, EP_ECAPACITY, EP_ENODEVICE, EP_ESTATE = 0, -1, -2, -3, -4, -5
EP_CONVEX, EP_PLANE, EP_SPHERE, EP_CAPSULE, EP_PARTICLE = 0, 1, 2, 3, 4
EP_STATIC, EP_DYNAMIC, EP_KINEMATIC = 1, 2, 3
EP_POINT_TO_POINT, EP_HINGE, EP_LOCK, EP_DISTANCE = 0, 1, 2, 3
EP_CONVEX_F64_HEADER, EP_CONVEX_I32_HEADER, EP_GROUP_F64, EP_GROUP_I32, EP_EDGE_I32 = 15, 4, 8, 5, 2
EP_STAT_COUNT = 8
EP_K_COUNT = 18
EP_OPT_PATH, EP_OPT_CONTACT_CAP, EP_OPT_RAW_CAP, EP_OPT_ITEM_CAP = 0, 1, 2, 3
# ep_step_frame_fields masks
EP_F_POS, EP_F_QUAT, EP_F_VEL, EP_F_ANGVEL, EP_F_FORCE, EP_F_TORQUE, EP_F_INV_INERTIA_WORLD, EP_F_CONSTANTS = 1, 2, 4, 8, 16, 32, 64, 128
EP_F_STATE, EP_F_ALL = 15, 255
BODY_FIELD_MASK = {"pos": 1, "quat": 2, "vel": 4, "angvel": 8, "force": 16, "torque": 32, "inv_inertia_world": 64, "inv_mass": 128,
                   "inv_inertia": 128, "lin_damp_pow": 128, "ang_damp_pow": 128, "lin_lock": 128, "ang_lock": 128, "bounding_radius": 128}

_pi32 = C.POINTER(C.c_int32)
_pi64 = C.POINTER(C.c_int64)
_pf64 = C.POINTER(C.c_double)
_pu8 = C.POINTER(C.c_uint8)


class EpShapes(C.Structure):
    _fields_ = [("n_shapes", C.c_int32), ("kind", _pi32), ("f64_off", _pi32), ("i32_off", _pi32),
                ("f64", _pf64), ("i32", _pi32)]


class EpBodies(C.Structure):
    _fields_ = [("n_worlds", C.c_int32), ("n_bodies", C.c_int32), ("world_body_off", _pi32), ("body_id", _pi32),
                ("kind", _pi32), ("pos", _pf64), ("quat", _pf64), ("vel", _pf64), ("angvel", _pf64),
                ("force", _pf64), ("torque", _pf64), ("inv_inertia_world", _pf64), ("inv_mass", _pf64),
                ("inv_inertia", _pf64), ("lin_damp_pow", _pf64), ("ang_damp_pow", _pf64), ("lin_lock", _pf64),
                ("ang_lock", _pf64), ("bounding_radius", _pf64), ("inst_off", _pi32), ("n_insts", C.c_int32),
                ("inst_shape", _pi32), ("inst_friction", _pf64), ("inst_bounciness", _pf64)]


class EpParams(C.Structure):
    _fields_ = [("n_worlds", C.c_int32), ("gravity", _pf64), ("dt", _pf64), ("iterations", _pi32),
                ("collide", _pu8), ("collide_off", _pi64), ("n_constraints", C.c_int32), ("con_body_a", _pi32),
                ("con_body_b", _pi32), ("con_kind", _pi32), ("con_data", _pf64)]


class EpContacts(C.Structure):
    _fields_ = [("n_worlds", C.c_int32), ("group_cap", C.c_int32), ("contact_cap", C.c_int32),
                ("world_group_off", _pi32), ("group_id1", _pi32), ("group_id2", _pi32),
                ("group_n_constraints", _pi32), ("group_contact_off", _pi32), ("group_pre_pos1", _pf64),
                ("group_pre_quat1", _pf64), ("shape_key", _pi32), ("feature_key", _pi64), ("ni", _pf64),
                ("pi", _pf64), ("pj", _pf64), ("friction", _pf64), ("bounciness", _pf64), ("lambda_", _pf64),
                ("t1", _pf64), ("iterations", _pi32), ("world_n_islands", _pi32), ("group_island", _pi32)]


def _ptr(a, typ):
    return None if a is None else a.ctypes.data_as(typ)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a if shape is None else a.reshape(shape)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Shapes:
    """Packed shape table (ep_shapes)."""

    def __init__(self, kind, f64_off, i32_off, f64, i32):
        self.kind, self.f64_off, self.i32_off = _i32(kind), _i32(f64_off), _i32(i32_off)
        self.f64 = _f64(f64) if len(f64) else np.zeros(1)
        self.i32 = _i32(i32) if len(i32) else np.zeros(1, np.int32)
        self.n_shapes = len(self.kind)
        self.c = EpShapes(self.n_shapes, _ptr(self.kind, _pi32), _ptr(self.f64_off, _pi32),
                          _ptr(self.i32_off, _pi32), _ptr(self.f64, _pf64), _ptr(self.i32, _pi32))


BODY_F64_FIELDS = (("pos", 3), ("quat", 4), ("vel", 3), ("angvel", 3), ("force", 3), ("torque", 3),
                   ("inv_inertia_world", 9), ("inv_mass", 1), ("inv_inertia", 3), ("lin_damp_pow", 1),
                   ("ang_damp_pow", 1), ("lin_lock", 3), ("ang_lock", 3), ("bounding_radius", 1))
BODY_INOUT = ("pos", "quat", "vel", "angvel", "force", "torque", "inv_inertia_world")


class Bodies:
    """Packed body state (ep_bodies).  Arrays are attributes: pos (n,3), quat (n,4) ..."""

    def __init__(self, world_body_off, body_id, kind, inst_off, inst_shape, inst_friction, inst_bounciness, alloc=None, **f):
        """`alloc(shape, dtype)` places the per-body f64 arrays (default numpy; engine.PinnedPool.empty for pinned)."""
        self.world_body_off, self.body_id, self.kind = _i32(world_body_off), _i32(body_id), _i32(kind)
        self.n_worlds, self.n_bodies = len(self.world_body_off) - 1, len(self.body_id)
        n = self.n_bodies
        for name, w in BODY_F64_FIELDS:
            src = _f64(f[name], (n, w) if w > 1 else (n,))
            if alloc is None:
                setattr(self, name, src.copy())
            else:
                a = alloc(src.shape, np.float64)
                a[...] = src
                setattr(self, name, a)
        self.inst_off, self.inst_shape = _i32(inst_off), _i32(inst_shape)
        self.inst_friction, self.inst_bounciness = _f64(inst_friction), _f64(inst_bounciness)
        self.n_insts = len(self.inst_shape)
        self._rebuild()

    def _rebuild(self):
        self.c = EpBodies(self.n_worlds, self.n_bodies, _ptr(self.world_body_off, _pi32), _ptr(self.body_id, _pi32),
                          _ptr(self.kind, _pi32),
                          *[_ptr(getattr(self, name), _pf64) for name, _ in BODY_F64_FIELDS],
                          _ptr(self.inst_off, _pi32), self.n_insts, _ptr(self.inst_shape, _pi32),
                          _ptr(self.inst_friction, _pf64), _ptr(self.inst_bounciness, _pf64))

    def copy(self, alloc=None):
        return Bodies(self.world_body_off, self.body_id, self.kind, self.inst_off, self.inst_shape,
                      self.inst_friction, self.inst_bounciness, alloc=alloc,
                      **{name: getattr(self, name) for name, _ in BODY_F64_FIELDS})

    def slice_worlds(self, w0, w1):
        """Worlds [w0, w1) as an independent Bodies (rows, instances and offsets rebased)."""
        r0, r1 = int(self.world_body_off[w0]), int(self.world_body_off[w1])
        i0, i1 = int(self.inst_off[r0]), int(self.inst_off[r1])
        return Bodies(self.world_body_off[w0:w1 + 1] - r0, self.body_id[r0:r1], self.kind[r0:r1],
                      self.inst_off[r0:r1 + 1] - i0, self.inst_shape[i0:i1], self.inst_friction[i0:i1],
                      self.inst_bounciness[i0:i1], **{name: getattr(self, name)[r0:r1] for name, _ in BODY_F64_FIELDS})

    def state_bytes(self):
        """In/out state as one byte string (for bit-exact comparisons)."""
        return b"".join(getattr(self, n).tobytes() for n in BODY_INOUT)

    def nbytes_in(self):
        return sum(getattr(self, n).nbytes for n, _ in BODY_F64_FIELDS) + self.body_id.nbytes + self.kind.nbytes

    def nbytes_out(self):
        return sum(getattr(self, n).nbytes for n in BODY_INOUT)


class Params:
    """Per-world step parameters + callback results (ep_params)."""

    def __init__(self, gravity, dt, iterations, collide=None, collide_off=None, con_body_a=(), con_body_b=(),
                 con_kind=(), con_data=()):
        self.gravity = _f64(gravity).reshape(-1, 3)
        self.n_worlds = len(self.gravity)
        self.dt, self.iterations = _f64(dt), _i32(iterations)
        self.collide = None if collide is None else np.ascontiguousarray(collide, dtype=np.uint8)
        self.collide_off = None if collide is None else np.ascontiguousarray(collide_off, dtype=np.int64)
        self.con_body_a, self.con_body_b, self.con_kind = _i32(con_body_a), _i32(con_body_b), _i32(con_kind)
        self.con_data = _f64(con_data).reshape(-1, 24)
        self.n_constraints = len(self.con_kind)
        nz = self.n_constraints > 0
        self.c = EpParams(self.n_worlds, _ptr(self.gravity, _pf64), _ptr(self.dt, _pf64), _ptr(self.iterations, _pi32),
                          _ptr(self.collide, _pu8), _ptr(self.collide_off, _pi64), self.n_constraints,
                          _ptr(self.con_body_a, _pi32) if nz else None, _ptr(self.con_body_b, _pi32) if nz else None,
                          _ptr(self.con_kind, _pi32) if nz else None, _ptr(self.con_data, _pf64) if nz else None)


CONTACT_FIELDS = (("shape_key", np.int32, 1), ("feature_key", np.int64, 1), ("ni", np.float64, 3),
                  ("pi", np.float64, 3), ("pj", np.float64, 3), ("friction", np.float64, 1),
                  ("bounciness", np.float64, 1), ("lambda_", np.float64, 1), ("t1", np.float64, 3))


# construction order of the arrays of Contacts (see __init__)
ALL_CONTACT_ARRAYS = ("world_group_off", "group_id1", "group_id2", "group_n_constraints", "group_contact_off", "group_pre_pos1",
                      "group_pre_quat1") + tuple(n for n, _, _ in CONTACT_FIELDS) + ("iterations", "world_n_islands", "group_island")
# what Physics.contactPoints consumes (src/Physics.elm:1158-1207)
CONTACT_POINTS_FIELDS = ("group_id1", "group_id2", "group_pre_pos1", "group_pre_quat1", "pi")


def slice_params(p, w0, w1):
    """Worlds [w0, w1) of a Params without callbacks/constraints (batched scenes)."""
    assert p.collide is None and p.n_constraints == 0
    return Params(p.gravity[w0:w1], p.dt[w0:w1], p.iterations[w0:w1])


class Contacts:
    """Pair groups + contacts (ep_contacts) with caller-chosen capacities."""

    def __init__(self, n_worlds, group_cap, contact_cap, alloc=None, fields=None):
        """`fields`: names of the optional arrays to hold (None = all).  The others stay NULL in the C struct and are not
        copied back by the downloads (include/elm_physics_b200.h); CONTACT_POINTS_FIELDS is what contactPoints reads."""
        self.n_worlds, self.group_cap, self.contact_cap = n_worlds, group_cap, contact_cap
        g, c = max(group_cap, 1), max(contact_cap, 1)
        always = ("world_group_off", "group_contact_off")
        made = []

        def zeros(shape, dtype=np.float64):
            made.append(None)
            name = ALL_CONTACT_ARRAYS[len(made) - 1]
            if fields is not None and name not in fields and name not in always:
                return None
            if alloc is None:
                return np.zeros(shape, dtype)
            a = alloc(shape, dtype)
            a[...] = 0
            return a
        self.world_group_off = zeros(n_worlds + 1, np.int32)
        self.group_id1 = zeros(g, np.int32)
        self.group_id2 = zeros(g, np.int32)
        self.group_n_constraints = zeros(g, np.int32)
        self.group_contact_off = zeros(g + 1, np.int32)
        self.group_pre_pos1 = zeros((g, 3))
        self.group_pre_quat1 = zeros((g, 4))
        for name, dt, w in CONTACT_FIELDS:
            setattr(self, name, zeros((c, w) if w > 1 else (c,), dt))
        self.iterations = zeros(max(n_worlds, 1), np.int32)
        self.world_n_islands = zeros(max(n_worlds, 1), np.int32)
        self.group_island = zeros(g, np.int32)
        self.c = EpContacts(
            n_worlds, group_cap, contact_cap, _ptr(self.world_group_off, _pi32), _ptr(self.group_id1, _pi32),
            _ptr(self.group_id2, _pi32), _ptr(self.group_n_constraints, _pi32), _ptr(self.group_contact_off, _pi32),
            _ptr(self.group_pre_pos1, _pf64), _ptr(self.group_pre_quat1, _pf64), _ptr(self.shape_key, _pi32),
            _ptr(self.feature_key, _pi64), _ptr(self.ni, _pf64), _ptr(self.pi, _pf64), _ptr(self.pj, _pf64),
            _ptr(self.friction, _pf64), _ptr(self.bounciness, _pf64), _ptr(self.lambda_, _pf64), _ptr(self.t1, _pf64),
            _ptr(self.iterations, _pi32), _ptr(self.world_n_islands, _pi32), _ptr(self.group_island, _pi32))

    @property
    def n_groups(self):
        return int(self.world_group_off[self.n_worlds])

    @property
    def n_contacts(self):
        return int(self.group_contact_off[self.n_groups])

    def nbytes_out(self):
        """Bytes a download copies into this object (the arrays it holds, trimmed to the used groups / contacts)."""
        ng, nc, nw = self.n_groups, self.n_contacts, self.n_worlds
        per = {"world_group_off": nw + 1, "group_contact_off": ng + 1, "iterations": nw, "world_n_islands": nw}
        total = 0
        for name in ALL_CONTACT_ARRAYS:
            a = getattr(self, name)
            if a is None:
                continue
            rows = per.get(name, ng if name.startswith("group_") else nc)
            total += rows * a.itemsize * (a.shape[1] if a.ndim > 1 else 1)
        return total

    def canonical(self):
        """Everything a frame's contacts are, as comparable numpy arrays trimmed to the used length."""
        ng, nc = self.n_groups, self.n_contacts
        d = dict(world_group_off=self.world_group_off.copy(), group_id1=self.group_id1[:ng].copy(),
                 group_id2=self.group_id2[:ng].copy(), group_n_constraints=self.group_n_constraints[:ng].copy(),
                 group_contact_off=self.group_contact_off[:ng + 1].copy(),
                 group_pre_pos1=self.group_pre_pos1[:ng].copy(), group_pre_quat1=self.group_pre_quat1[:ng].copy(),
                 group_island=self.group_island[:ng].copy(), iterations=self.iterations[:self.n_worlds].copy(),
                 world_n_islands=self.world_n_islands[:self.n_worlds].copy())
        for name, _, _ in CONTACT_FIELDS:
            d[name] = getattr(self, name)[:nc].copy()
        return d
