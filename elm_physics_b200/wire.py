"""The wire format of the reference-side shim (shim/README.md), a second time in Python.

`encode_world` mirrors shim/elm/Physics/B200/Encode.elm plus the `Math.pow` lines of shim/patch_simulate.js: one world's
`List (id, Body)` -> a dict holding one flat list per field of the C structs, under the C field's name.  `to_abi` mirrors
shim/ep_napi.cc: that dict -> abi.Shapes / abi.Bodies / abi.Params by wrapping the arrays, no arithmetic.  The two files under
shim/ cannot run in the build image (no elm / node); tests/test_wire_format.py checks this mirror, array by array, against
the packer every GPU parity test goes through (pack.pack_worlds), so the documented byte layout is the one the engine reads.
"""
import numpy as np

from . import abi
from . import pack

M_KEYS = ("m11", "m21", "m31", "m12", "m22", "m32", "m13", "m23", "m33")


def _flat(vs):
    return [float(c) for v in vs for c in v]


def encode_shape_table(shapes):
    """Physics.B200.Encode.shapeTable: rows of ep_shapes for a list of distinct centre-of-mass-frame shapes."""
    rows = [pack._serialise(s) for s in shapes]
    f_off, i_off = [0], [0]
    for f, i in rows:
        f_off.append(f_off[-1] + len(f))
        i_off.append(i_off[-1] + len(i))
    return dict(n_shapes=len(shapes), kind=[s.kind for s in shapes], f64_off=f_off, i32_off=i_off,
                f64=[x for f, _ in rows for x in f], i32=[x for _, i in rows for x in i])


def encode_world(config, bodies_with_ids):
    """One `simulate` call: (wire bodies, wire params, wire shape table).  Bodies carry their assigned ids
    (AssignIds.assignIds runs before, on the Elm side)."""
    bs = [b for _, b in bodies_with_ids]
    n = len(bs)
    shapes, index = [], {}
    inst_shape, inst_f, inst_b, inst_off = [], [], [], [0]
    for b in bs:
        for shape, mat in b.shapes:
            key = (shape.kind, tuple(pack._serialise(shape)[0]), tuple(pack._serialise(shape)[1]))
            if key not in index:
                index[key] = len(shapes)
                shapes.append(shape)
            inst_shape.append(index[key])
            inst_f.append(mat.friction)
            inst_b.append(mat.bounciness)
        inst_off.append(len(inst_shape))
    dt = config.duration
    wire_bodies = dict(
        n_worlds=1, n_bodies=n, n_insts=len(inst_shape), world_body_off=[0, n],
        body_id=[b.id for b in bs], kind=[b.kind for b in bs],
        pos=_flat(b.transform3d[0] for b in bs), quat=_flat(b.transform3d[1] for b in bs),
        vel=_flat(b.velocity for b in bs), angvel=_flat(b.angular_velocity for b in bs),
        force=_flat(b.force for b in bs), torque=_flat(b.torque for b in bs),
        inv_inertia_world=[float(b.inv_inertia_world[k]) for b in bs for k in M_KEYS],
        inv_mass=[b.inv_mass for b in bs], inv_inertia=_flat(b.inv_inertia for b in bs),
        # the Elm encoder sends the raw damping; the JS side turns it into the factor with Math.pow (SolverBody.elm:149-153)
        lin_damp_pow=[(1.0 - b.linear_damping) ** dt for b in bs], ang_damp_pow=[(1.0 - b.angular_damping) ** dt for b in bs],
        lin_lock=_flat(b.linear_lock for b in bs), ang_lock=_flat(b.angular_lock for b in bs),
        bounding_radius=[b.bounding_sphere_radius for b in bs],
        inst_off=inst_off, inst_shape=inst_shape, inst_friction=inst_f, inst_bounciness=inst_b)
    rows = []
    for ia, (ea, ba) in enumerate(bodies_with_ids):              # constrain callbacks, every ordered pair (BroadPhase.elm:144-186)
        fn = config.constrain(ea)
        if fn is None:
            continue
        for ib, (eb, bb) in enumerate(bodies_with_ids):
            if ib == ia:
                continue
            for c in fn(eb) or []:
                k, d = pack._constraint_rows(c, ba.center_of_mass_transform3d, bb.center_of_mass_transform3d)
                rows.append((ia, ib, k, d))
    wire_params = dict(gravity=list(config.gravity), dt=[dt], iterations=[config.solver_iterations],
                       n_constraints=len(rows), con_body_a=[r[0] for r in rows], con_body_b=[r[1] for r in rows],
                       con_kind=[r[2] for r in rows], con_data=[x for r in rows for x in r[3]])
    if config.collide is not None:
        wire_params["collide"] = [1 if config.collide(ea, eb) else 0 for ea, _ in bodies_with_ids for eb, _ in bodies_with_ids]
        wire_params["collide_off"] = [0]
    return wire_bodies, wire_params, encode_shape_table(shapes)


def to_abi(wire_bodies, wire_params, wire_shapes):
    """What ep_napi.cc does: wrap the arrays of the wire objects as the C structs (no arithmetic)."""
    s = wire_shapes
    shapes = abi.Shapes(s["kind"], s["f64_off"], s["i32_off"], s["f64"], s["i32"])
    b = wire_bodies
    n = b["n_bodies"]
    cols = {name: np.array(b[name], dtype=np.float64).reshape((n, w) if w > 1 else (n,)) for name, w in abi.BODY_F64_FIELDS}
    bodies = abi.Bodies(b["world_body_off"], b["body_id"], b["kind"], b["inst_off"], b["inst_shape"], b["inst_friction"],
                        b["inst_bounciness"], **cols)
    p = wire_params
    params = abi.Params(np.array(p["gravity"], dtype=np.float64).reshape(-1, 3), p["dt"], p["iterations"],
                        p.get("collide"), p.get("collide_off"), p["con_body_a"], p["con_body_b"], p["con_kind"],
                        np.array(p["con_data"], dtype=np.float64).reshape(-1, 24))
    return shapes, bodies, params
