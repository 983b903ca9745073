"""Packing of host-side `Body` values into the C-ABI SoA buffers (include/elm_physics_b200.h)
and back.  Pure marshalling — no physics (north_star: "the shim must be pure data marshalling").
"""
import numpy as np

from . import abi
from . import math3d as m3
from . import shapes as sh
from .physics import Body, Distance, Hinge, Lock, PointToPoint

M_KEYS = ("m11", "m21", "m31", "m12", "m22", "m32", "m13", "m23", "m33")


class ShapeTable:
    """Deduplicating builder of the immutable shape table (ep_shapes)."""

    def __init__(self):
        self.kind, self.f64_off, self.i32_off, self.f64, self.i32 = [], [0], [0], [], []
        self._index = {}

    def add(self, shape):
        f, i = _serialise(shape)
        key = (shape.kind, tuple(f), tuple(i))
        if key in self._index:
            return self._index[key]
        idx = len(self.kind)
        self._index[key] = idx
        self.kind.append(shape.kind)
        self.f64.extend(f)
        self.i32.extend(i)
        self.f64_off.append(len(self.f64))
        self.i32_off.append(len(self.i32))
        return idx

    def build(self):
        return abi.Shapes(self.kind, self.f64_off, self.i32_off, self.f64, self.i32)


def _serialise(s):
    if s.kind == sh.KIND_SPHERE:
        return [*s.position, s.radius], []
    if s.kind == sh.KIND_CAPSULE:
        return [*s.position, *s.axis, s.radius, s.half_length], []
    if s.kind == sh.KIND_PLANE:
        return [*s.position, *s.normal], []
    if s.kind == sh.KIND_PARTICLE:
        return [*s.position], []
    f = [*s.position]
    if s.box is not None:
        ax, ay, az, he = s.box
        f += [*ax, *ay, *az, *he]
    else:
        f += [0.0] * 12
    for v in s.vertices:
        f += [*v]
    ng, ne = len(s.faces), len(s.unique_edges)
    i = [1 if s.box is not None else 0, len(s.vertices), ng, ne]
    data_off = abi.EP_CONVEX_I32_HEADER + abi.EP_GROUP_I32 * ng + abi.EP_EDGE_I32 * ne
    data = []
    for g in s.faces:
        f += [*g.normal, g.dist]
        if g.two_sided:
            f += [*g.partner_normal, g.partner_dist]
            i += [1, len(g.indices), data_off + len(data)]
            data += list(g.indices)
            i += [len(g.partner_indices), data_off + len(data)]
            data += list(g.partner_indices)
        else:
            f += [0.0, 0.0, 0.0, 0.0]
            i += [0, len(g.indices), data_off + len(data), 0, 0]
            data += list(g.indices)
    for e in s.unique_edges:
        i += [len(e), data_off + len(data)]
        data += list(e)
    return f, i + data


def _constraint_rows(c, com_a, com_b):
    """Constraint.relativeToCenterOfMass (Constraint.elm:16-47): a-side data with body a's CoM frame,
    b-side with body b's.  Returns (kind, 24 doubles)."""
    z = [0.0] * 24
    if isinstance(c, PointToPoint):
        z[0:3] = m3.point_relative_to(com_a, c.pivot1)
        z[12:15] = m3.point_relative_to(com_b, c.pivot2)
        return abi.EP_POINT_TO_POINT, z
    if isinstance(c, Hinge):
        z[0:3] = m3.point_relative_to(com_a, c.pivot1)
        z[3:6] = m3.direction_relative_to(com_a, c.axis1)
        z[12:15] = m3.point_relative_to(com_b, c.pivot2)
        z[15:18] = m3.direction_relative_to(com_b, c.axis2)
        return abi.EP_HINGE, z
    if isinstance(c, Lock):
        z[0:3] = m3.point_relative_to(com_a, c.pivot1)
        z[3:6] = m3.direction_relative_to(com_a, c.x1)
        z[6:9] = m3.direction_relative_to(com_a, c.y1)
        z[9:12] = m3.direction_relative_to(com_a, c.z1)
        z[12:15] = m3.point_relative_to(com_b, c.pivot2)
        z[15:18] = m3.direction_relative_to(com_b, c.x2)
        z[18:21] = m3.direction_relative_to(com_b, c.y2)
        z[21:24] = m3.direction_relative_to(com_b, c.z2)
        return abi.EP_LOCK, z
    if isinstance(c, Distance):
        z[0] = c.length
        return abi.EP_DISTANCE, z
    raise TypeError(c)


def pack_worlds(worlds, table=None):
    """worlds: list of (config, [(ext_id, Body with assigned id)]).  Returns (Shapes, Bodies, Params)."""
    table = table or ShapeTable()
    wbo, bid, kind, inst_off, inst_shape, inst_f, inst_b = [0], [], [], [0], [], [], []
    cols = {name: [] for name, _ in abi.BODY_F64_FIELDS}
    gravity, dts, iters = [], [], []
    collide, collide_off, any_collide = [], [], any(cfg.collide is not None for cfg, _ in worlds)
    con_a, con_b, con_k, con_d = [], [], [], []
    for cfg, bodies in worlds:
        base = len(bid)
        dt = cfg.duration
        gravity.append(tuple(cfg.gravity))
        dts.append(dt)
        iters.append(cfg.solver_iterations)
        for ext, b in bodies:
            bid.append(b.id)
            kind.append(b.kind)
            cols["pos"].append(b.transform3d[0])
            cols["quat"].append(b.transform3d[1])
            cols["vel"].append(b.velocity)
            cols["angvel"].append(b.angular_velocity)
            cols["force"].append(b.force)
            cols["torque"].append(b.torque)
            cols["inv_inertia_world"].append([b.inv_inertia_world[k] for k in M_KEYS])
            cols["inv_mass"].append(b.inv_mass)
            cols["inv_inertia"].append(b.inv_inertia)
            cols["lin_damp_pow"].append((1.0 - b.linear_damping) ** dt)      # SolverBody.elm:149-153
            cols["ang_damp_pow"].append((1.0 - b.angular_damping) ** dt)
            cols["lin_lock"].append(b.linear_lock)
            cols["ang_lock"].append(b.angular_lock)
            cols["bounding_radius"].append(b.bounding_sphere_radius)
            for shape, mat in b.shapes:
                inst_shape.append(table.add(shape))
                inst_f.append(mat.friction)
                inst_b.append(mat.bounciness)
            inst_off.append(len(inst_shape))
        n = len(bodies)
        wbo.append(len(bid))
        if any_collide:
            collide_off.append(len(collide))
            fn = cfg.collide or (lambda a, b: True)
            collide += [1 if fn(ea, eb) else 0 for ea, _ in bodies for eb, _ in bodies]
        for ia, (ea, ba) in enumerate(bodies):      # constrain callbacks, BroadPhase.elm:144-186
            fn = cfg.constrain(ea)
            if fn is None:
                continue
            for ib, (eb, bb) in enumerate(bodies):
                if ib == ia:
                    continue
                for c in fn(eb) or []:
                    k, d = _constraint_rows(c, ba.center_of_mass_transform3d, bb.center_of_mass_transform3d)
                    con_a.append(base + ia)
                    con_b.append(base + ib)
                    con_k.append(k)
                    con_d.append(d)
    nb = len(bid)
    f = {}
    for name, w in abi.BODY_F64_FIELDS:
        f[name] = np.array(cols[name], dtype=np.float64).reshape((nb, w) if w > 1 else (nb,))
    bodies_p = abi.Bodies(wbo, bid, kind, inst_off, inst_shape, inst_f, inst_b, **f)
    params = abi.Params(np.array(gravity, dtype=np.float64).reshape(-1, 3), dts, iters,
                        collide if any_collide else None, collide_off if any_collide else None,
                        con_a, con_b, con_k, np.array(con_d, dtype=np.float64).reshape(-1, 24))
    return table.build(), bodies_p, params


def unpack_bodies(packed, worlds):
    """Write the integrated state back into new Body values; returns list of [(ext_id, Body)] per world."""
    from dataclasses import replace
    out = []
    r = 0
    for cfg, bodies in worlds:
        res = []
        for ext, b in bodies:
            m = packed.inv_inertia_world[r]
            res.append((ext, replace(
                b, transform3d=(tuple(packed.pos[r]), tuple(packed.quat[r])), velocity=tuple(packed.vel[r]),
                angular_velocity=tuple(packed.angvel[r]), force=tuple(packed.force[r]), torque=tuple(packed.torque[r]),
                inv_inertia_world={k: float(m[i]) for i, k in enumerate(M_KEYS)})))
            r += 1
        out.append(res)
    return out


def contacts_world_dict(c, w):
    """Slice one world's groups/contacts out of an abi.Contacts as plain dict-of-arrays (used to feed a
    single world's contacts back as warm start)."""
    g0, g1 = int(c.world_group_off[w]), int(c.world_group_off[w + 1])
    c0, c1 = int(c.group_contact_off[g0]), int(c.group_contact_off[g1])
    d = dict(g0=g0, g1=g1, c0=c0, c1=c1)
    for name in ("group_id1", "group_id2", "group_n_constraints", "group_island", "group_pre_pos1", "group_pre_quat1"):
        d[name] = getattr(c, name)[g0:g1].copy()
    d["group_contact_off"] = c.group_contact_off[g0:g1 + 1].copy() - c0
    for name, _, _ in abi.CONTACT_FIELDS:
        d[name] = getattr(c, name)[c0:c1].copy()
    d["iterations"] = int(c.iterations[w])
    return d


def contacts_from_world_dicts(dicts):
    """Inverse of contacts_world_dict for a list of worlds (None = emptyContacts)."""
    ng = sum(0 if d is None else len(d["group_id1"]) for d in dicts)
    nc = sum(0 if d is None else len(d["shape_key"]) for d in dicts)
    c = abi.Contacts(len(dicts), ng, nc)
    g = k = 0
    for w, d in enumerate(dicts):
        c.world_group_off[w] = g
        if d is not None:
            n = len(d["group_id1"])
            m = len(d["shape_key"])
            for name in ("group_id1", "group_id2", "group_n_constraints", "group_island", "group_pre_pos1", "group_pre_quat1"):
                getattr(c, name)[g:g + n] = d[name]
            c.group_contact_off[g:g + n + 1] = d["group_contact_off"] + k
            for name, _, _ in abi.CONTACT_FIELDS:
                getattr(c, name)[k:k + m] = d[name]
            c.iterations[w] = d["iterations"]
            g += n
            k += m
    c.world_group_off[len(dicts)] = g
    c.group_contact_off[g] = k
    return c


def unpack_pair_groups(d):
    """`pairGroups` view of one world's packed contacts, in the reference's list order
    (reverse enumeration; contacts in pairGroup.contacts order = reverse solver order)."""
    groups = []
    n = len(d["group_id1"])
    for g in range(n - 1, -1, -1):
        c0, c1 = int(d["group_contact_off"][g]), int(d["group_contact_off"][g + 1])
        cs = [dict(shape_key=int(d["shape_key"][k]), feature_key=int(d["feature_key"][k]), ni=tuple(d["ni"][k]),
                   pi=tuple(d["pi"][k]), pj=tuple(d["pj"][k]), friction=float(d["friction"][k]),
                   bounciness=float(d["bounciness"][k]), lambda_=float(d["lambda_"][k]), t1=tuple(d["t1"][k]))
              for k in range(c1 - 1, c0 - 1, -1)]
        groups.append(dict(id1=int(d["group_id1"][g]), id2=int(d["group_id2"][g]), contacts=cs,
                           n_constraints=int(d["group_n_constraints"][g]),
                           pre_transform1=(tuple(d["group_pre_pos1"][g]), tuple(d["group_pre_quat1"][g]))))
    return groups
