"""Host-side mirror of the reference's public `Physics` module for the path behind
`Physics.simulate` (reference: src/Physics.elm, src/Internal/Body.elm,
src/Internal/AssignIds.elm, src/Physics/Shape.elm, src/Physics/Material.elm).

Everything here is construction-time / between-step host logic that the reference
also keeps outside the step: building bodies, mass
properties, id assignment, user callbacks.  The step itself is `simulate`, which
packs the bodies (pack.py) and calls the CUDA library through the C ABI
(include/elm_physics_b200.h); there is no CPU stepping code in this package.
"""
import copy
import math
from dataclasses import dataclass, field, replace
from typing import Callable, List, Optional

from . import math3d as m3
from . import shapes as sh


# ---- Material (Internal/Material.elm:36-58) ------------------------------------
@dataclass(frozen=True)
class Material:
    bounciness: float
    friction: float
    density: float


wood = Material(friction=0.4, bounciness=0.3, density=700)
rubber = Material(friction=0.8, bounciness=0.7, density=1100)
steel = Material(friction=0.3, bounciness=0.2, density=7800)
ice = Material(friction=0.03, bounciness=0.1, density=900)
plastic = Material(friction=0.35, bounciness=0.45, density=1050)


# ---- Constraint (Internal/Constraint.elm:9-13; public: Physics/Constraint.elm) ----
@dataclass(frozen=True)
class PointToPoint:
    pivot1: tuple
    pivot2: tuple


@dataclass(frozen=True)
class Hinge:
    pivot1: tuple
    axis1: tuple
    pivot2: tuple
    axis2: tuple


@dataclass(frozen=True)
class Lock:
    pivot1: tuple
    x1: tuple
    y1: tuple
    z1: tuple
    pivot2: tuple
    x2: tuple
    y2: tuple
    z2: tuple


@dataclass(frozen=True)
class Distance:
    length: float


# ---- Body (Internal/Body.elm:36-72) -----------------------------------------------
@dataclass
class Body:
    id: int
    kind: int                        # 1 static, 2 dynamic, 3 kinematic (Body.elm:22-34)
    transform3d: tuple               # CoM frame in world: (origin, quaternion xyzw)
    center_of_mass_transform3d: tuple
    velocity: tuple
    angular_velocity: tuple
    mass: float
    shapes: list                     # geometry.shapesWithMaterials: [(shape in CoM frame, Material)]
    bounding_sphere_radius: float
    volume: float
    force: tuple
    torque: tuple
    linear_damping: float
    angular_damping: float
    inv_mass: float
    inv_inertia: tuple
    inv_inertia_world: dict
    linear_lock: tuple
    angular_lock: tuple


def _place_shapes(inv_com, com, shapes):
    """Body.elm:120-173."""
    inertia = m3.m_zero()
    solid = []
    radius = 0.0
    for shape, mat, sign in shapes:
        moved = sh.place_in(inv_com, shape)
        volume = sh.shape_volume(moved)
        res = m3.inertia_place_in(com, m3.point_place_in(com, sh.shape_center_of_mass(moved)), volume,
                                  sh.shape_inertia(moved))
        inertia = m3.m_add(inertia, m3.m_scale(sign * mat.density, res))
        if sign > 0:
            solid.insert(0, (moved, mat))
            radius = sh.expand_bounding_sphere_radius(moved, radius)
    return inertia, solid, radius


def compound(kind, raw_shapes):
    """Body.elm:176-268.  raw_shapes: [(shape in body coordinates, Material, sign)]."""
    if kind == 2:
        shapes = list(raw_shapes)
    else:
        shapes = [(s, replace(m, density=0), sign) for s, m, sign in raw_shapes]
    total_mass = total_volume = cx = cy = cz = 0.0
    for shape, mat, sign in shapes:
        signed_volume = sign * sh.shape_volume(shape)
        signed_mass = signed_volume * mat.density
        x, y, z = sh.shape_center_of_mass(shape)
        total_mass = total_mass + signed_mass
        total_volume = total_volume + signed_volume
        cx, cy, cz = cx + signed_mass * x, cy + signed_mass * y, cz + signed_mass * z
    com_point = (cx / total_mass, cy / total_mass, cz / total_mass) if total_mass > 0 else m3.ZERO
    init_com = m3.t_at_point(com_point)
    init_inertia, _, _ = _place_shapes(m3.t_inverse(init_com), init_com, shapes)
    eig = m3.eigen_decomposition(init_inertia)
    eigen_rotation = m3.from_origin_and_basis(m3.ZERO, eig["v1"], eig["v2"], eig["v3"])
    com = m3.t_place_in(init_com, eigen_rotation)
    _, solid, radius = _place_shapes(m3.t_inverse(com), com, shapes)
    transform3d = m3.t_place_in(m3.AT_ORIGIN, com)
    ev = eig["eigenvalues"]
    inv_inertia = tuple(0.0 if e == 0 else 1 / e for e in ev)
    return Body(
        id=-1, kind=kind, velocity=m3.ZERO, angular_velocity=m3.ZERO, transform3d=transform3d,
        center_of_mass_transform3d=com, mass=total_mass, shapes=solid, bounding_sphere_radius=radius,
        volume=total_volume, linear_damping=0.01, angular_damping=0.01,
        inv_mass=0.0 if total_mass == 0 else 1 / total_mass, inv_inertia=inv_inertia,
        inv_inertia_world=m3.inverted_inertia_rotate_in(transform3d, inv_inertia),
        force=m3.ZERO, torque=m3.ZERO, linear_lock=(1.0, 1.0, 1.0), angular_lock=(1.0, 1.0, 1.0))


def point_mass(position, mass, material):
    """Body.elm:271-302 / Physics.elm:204-213."""
    mass = max(m3.PRECISION, mass)
    mat = Material(friction=material.friction, bounciness=material.bounciness, density=0)
    return Body(
        id=-1, kind=2, velocity=m3.ZERO, angular_velocity=m3.ZERO, transform3d=m3.t_at_point(tuple(position)),
        center_of_mass_transform3d=m3.AT_ORIGIN, mass=mass, shapes=[(sh.Particle(m3.ZERO), mat)],
        bounding_sphere_radius=0.0, volume=0.0, linear_damping=0.01, angular_damping=0.01,
        inv_mass=1 / mass, inv_inertia=m3.ZERO, inv_inertia_world=m3.m_zero(), force=m3.ZERO, torque=m3.ZERO,
        linear_lock=(1.0, 1.0, 1.0), angular_lock=(1.0, 1.0, 1.0))


# ---- Shapes (Physics/Shape.elm) : a Shape is a list of (internal shape, sign) ------------
def _perpendicular_basis(d):
    """elm-geometry Direction3d.perpendicularBasis (un-vendored dependency)."""
    ax, ay, az = abs(d[0]), abs(d[1]), abs(d[2])
    if ax <= ay:
        if ax <= az:
            s = math.sqrt(d[2] * d[2] + d[1] * d[1])
            a = (0.0, -d[2] / s, d[1] / s)
        else:
            s = math.sqrt(d[1] * d[1] + d[0] * d[0])
            a = (-d[1] / s, d[0] / s, 0.0)
    elif ay <= az:
        s = math.sqrt(d[2] * d[2] + d[0] * d[0])
        a = (d[2] / s, 0.0, -d[0] / s)
    else:
        s = math.sqrt(d[1] * d[1] + d[0] * d[0])
        a = (-d[1] / s, d[0] / s, 0.0)
    return a, m3.cross(d, a)


def _axial_transform(center, axis):
    a, b = _perpendicular_basis(axis)
    return m3.from_origin_and_basis(tuple(center), a, b, tuple(axis))


def shape_block(size, center=m3.ZERO, x=m3.X_AXIS, y=m3.Y_AXIS, z=m3.Z_AXIS):
    """Physics/Shape.elm:52-93."""
    t = m3.from_origin_and_basis(tuple(center), tuple(x), tuple(y), tuple(z))
    return [(sh.convex_place_in(t, sh.from_block(size[0], size[1], size[2])), 1)]


def shape_sphere(radius, center=m3.ZERO):
    """Physics/Shape.elm:96-115."""
    return [(sh.place_in(m3.t_at_point(tuple(center)), sh.sphere_at_origin(radius)), 1)]


def shape_cylinder(subdivisions, radius, length, center=m3.ZERO, axis=m3.Z_AXIS):
    """Physics/Shape.elm:120-142."""
    return [(sh.convex_place_in(_axial_transform(center, axis),
                                sh.from_cylinder(max(3, subdivisions), radius, length)), 1)]


def shape_capsule(radius, length, center=m3.ZERO, axis=m3.Z_AXIS):
    """Physics/Shape.elm:147-170."""
    return [(sh.place_in(_axial_transform(center, axis), sh.capsule_at_origin(radius, length / 2)), 1)]


def shape_cone(subdivisions, radius, length, center=m3.ZERO, axis=m3.Z_AXIS):
    """Physics/Shape.elm:175-197 (center = midpoint of base and tip)."""
    return [(sh.convex_place_in(_axial_transform(center, axis),
                                sh.from_cone(max(3, subdivisions), radius, length)), 1)]


def shape_unsafe_convex(face_indices, vertices):
    """Physics/Shape.elm:207-225."""
    return [(sh.from_triangular_mesh(face_indices, vertices), 1)]


def shape_plus(to_add, base):
    return base + to_add


def shape_minus(to_subtract, base):
    return base + [(s, -sign) for s, sign in to_subtract]


def shape_sum(shapes):
    out = []
    for s in shapes:
        out = shape_plus(s, out)
    return out


# ---- Body constructors (Physics.elm:144-281) --------------------------------------------
def _entries(shapes_with_materials):
    return [(s, mat, sign) for shape, mat in shapes_with_materials for s, sign in shape]


def dynamic(shapes_with_materials):
    return compound(2, _entries(shapes_with_materials))


def static(shapes_with_materials):
    return compound(1, _entries(shapes_with_materials))


def kinematic(shapes_with_materials):
    return compound(3, _entries(shapes_with_materials))


def block(size, material, **kw):
    return dynamic([(shape_block(size, **kw), material)])


def sphere(radius, material, **kw):
    return dynamic([(shape_sphere(radius, **kw), material)])


def cylinder(radius, length, material, **kw):
    return dynamic([(shape_cylinder(12, radius, length, **kw), material)])


def capsule(radius, length, material, **kw):
    return dynamic([(shape_capsule(radius, length, **kw), material)])


def cone(radius, length, material, **kw):
    return dynamic([(shape_cone(12, radius, length, **kw), material)])


def plane(material, origin=m3.ZERO, normal=m3.Z_AXIS):
    """Physics.elm:157-175."""
    return compound(1, [(sh.Plane(tuple(normal), tuple(origin)), material, 1)])


# ---- Mutators (Physics.elm:291-497, 857-1113) -------------------------------------------
def _body_frame(body):
    return m3.t_place_in(body.transform3d, m3.t_inverse(body.center_of_mass_transform3d))


def move_to(point, body):
    bf = _body_frame(body)
    return replace(body, transform3d=m3.t_place_in((tuple(point), bf[1]), body.center_of_mass_transform3d))


def translate_by(vector, body):
    bf = _body_frame(body)
    return replace(body, transform3d=m3.t_place_in((m3.add(tuple(vector), bf[0]), bf[1]),
                                                   body.center_of_mass_transform3d))


def _with_rotated(body, new_body_frame):
    t = m3.t_place_in(new_body_frame, body.center_of_mass_transform3d)
    iiw = m3.inverted_inertia_rotate_in(t, body.inv_inertia) if body.kind == 2 else body.inv_inertia_world
    return replace(body, transform3d=t, inv_inertia_world=iiw)


def rotate_around(axis_origin, axis_direction, angle, body):
    """Physics.elm:380-430 (Point3d.rotateAround from elm-geometry restated with a quaternion)."""
    bf = _body_frame(body)
    q = m3.from_angle_axis(angle, tuple(axis_direction))
    rel = m3.sub(bf[0], tuple(axis_origin))
    rotated_origin = m3.add(tuple(axis_origin), m3.q_rotate(q, rel))
    return _with_rotated(body, m3.rotate_around_own(tuple(axis_direction), angle, (rotated_origin, bf[1])))


def place(origin, x, y, body):
    """Physics.elm:446-497."""
    z = m3.cross(tuple(x), tuple(y))
    return _with_rotated(body, m3.from_origin_and_basis(tuple(origin), tuple(x), tuple(y), z))


def place_quaternion(origin, q, body):
    """Convenience for generators: body frame given directly as origin + unit quaternion."""
    return _with_rotated(body, (tuple(origin), tuple(q)))


def apply_force(force, point, body):
    if body.kind != 2:
        return body
    rel = m3.sub(tuple(point), body.transform3d[0])
    return replace(body, force=m3.add(body.force, tuple(force)),
                   torque=m3.add(body.torque, m3.cross(rel, tuple(force))))


def apply_torque(torque, body):
    if body.kind != 2:
        return body
    return replace(body, torque=m3.add(body.torque, tuple(torque)))


def _iiw_mul(body, v):
    i = body.inv_inertia_world
    x, y, z = v
    w = body.angular_velocity
    return (w[0] + i["m11"] * x + i["m12"] * y + i["m13"] * z,
            w[1] + i["m21"] * x + i["m22"] * y + i["m23"] * z,
            w[2] + i["m31"] * x + i["m32"] * y + i["m33"] * z)


def apply_impulse(impulse, point, body):
    """Body.elm:305-343."""
    if body.kind != 2:
        return body
    impulse = tuple(impulse)
    rel = m3.sub(tuple(point), body.transform3d[0])
    v = body.velocity
    k = body.inv_mass
    return replace(body, velocity=(v[0] + k * impulse[0], v[1] + k * impulse[1], v[2] + k * impulse[2]),
                   angular_velocity=_iiw_mul(body, m3.cross(rel, impulse)))


def apply_angular_impulse(angular_impulse, body):
    if body.kind != 2:
        return body
    return replace(body, angular_velocity=_iiw_mul(body, tuple(angular_impulse)))


def set_velocity_to(v, body):
    return body if body.kind == 1 else replace(body, velocity=tuple(v))


def set_angular_velocity_to(w, body):
    return body if body.kind == 1 else replace(body, angular_velocity=tuple(w))


def scale_mass_to(new_mass, body):
    """Physics.elm:1023-1062."""
    if body.kind == 2 and new_mass > 0:
        s = body.mass / new_mass
        inv_inertia = (body.inv_inertia[0] * s, body.inv_inertia[1] * s, body.inv_inertia[2] * s)
        return replace(body, mass=new_mass, inv_mass=1 / new_mass, inv_inertia=inv_inertia,
                       inv_inertia_world=m3.inverted_inertia_rotate_in(body.transform3d, inv_inertia))
    return body


def damp(linear, angular, body):
    clamp = lambda v: 0.0 if v < 0 else (1.0 if v > 1 else v)
    return replace(body, linear_damping=clamp(linear), angular_damping=clamp(angular))


TRANSLATE_X, TRANSLATE_Y, TRANSLATE_Z, ROTATE_X, ROTATE_Y, ROTATE_Z = range(6)


def lock(locks, body):
    """Internal/Lock.elm:19-46."""
    if body.kind != 2:
        return body
    lin = [1.0, 1.0, 1.0]
    ang = [1.0, 1.0, 1.0]
    for l in locks:
        if l < 3:
            lin[l] = 0.0
        else:
            ang[l - 3] = 0.0
    return replace(body, linear_lock=tuple(lin), angular_lock=tuple(ang))


def origin_point(body):
    """Physics.elm:775-781 : centre of mass in world."""
    return body.transform3d[0]


def frame(body):
    """Body-coordinates frame in world (Physics.elm:671-690)."""
    return _body_frame(body)


# ---- AssignIds (Internal/AssignIds.elm:7-127) --------------------------------------------
def assign_ids(bodies_with_ids):
    """Keeps existing ids; `-1` bodies and the FIRST occurrences of duplicated ids get the
    smallest free ids, in list order.  Returns (list of (ext_id, body), max_id)."""
    existing = [b.id for _, b in bodies_with_ids if b.id != -1]
    new_count = sum(1 for _, b in bodies_with_ids if b.id == -1)
    mx = max(existing) if existing else -1
    srt = sorted(existing)
    dup_ids = []                      # one entry per extra occurrence, DESCENDING (cons of ascending)
    prev = -2
    for x in srt:
        if x == prev:
            dup_ids.insert(0, x)
        prev = x
    needed = new_count + len(dup_ids)
    free = []
    n = 0
    rest = list(srt)
    while needed > 0:
        if not rest:
            free.append(n); n += 1; needed -= 1
        else:
            x = rest[0]
            if x - n > 0:
                free.append(n); n += 1; needed -= 1
            elif x == n:
                n += 1; rest.pop(0)
            else:
                rest.pop(0)
    if not free:
        return list(bodies_with_ids), mx
    out = []
    free_i = 0
    for ext, b in bodies_with_ids:
        if free_i < len(free):
            fresh = free[free_i]
            if b.id == -1:
                out.append((ext, replace(b, id=fresh))); free_i += 1; mx = max(mx, fresh)
                continue
            if b.id in dup_ids:
                dup_ids.remove(b.id)
                out.append((ext, replace(b, id=fresh))); free_i += 1; mx = max(mx, fresh)
                continue
        out.append((ext, b))
    return out, mx


# ---- Config / Contacts / simulate (Physics.elm:522-665) ----------------------------------
@dataclass
class Contacts:
    """Physics/Types.elm:19-25.  `packed` is the C-ABI contact block of the last frame
    (groups in pair-enumeration order, contacts in solver order, solved lambdas + t1): it is
    both the `pairGroups` view and the warm-start cache."""
    packed: Optional[dict] = None
    iterations: int = 0
    bodies: Optional[list] = None     # post-step [(ext_id, Body)] indexed like the input

    def pair_groups(self):
        from .pack import unpack_pair_groups
        return [] if self.packed is None else unpack_pair_groups(self.packed)


def empty_contacts():
    return Contacts()


GEE = 9.80665


@dataclass
class Config:
    gravity: tuple = (0.0, 0.0, -GEE)
    duration: float = 1 / 60
    solver_iterations: int = 20
    contacts: Contacts = field(default_factory=Contacts)
    constrain: Callable = lambda _id: None
    collide: Callable = None          # None == `\_ _ -> True` (lets the packer skip the N^2 mask)


def on_earth():
    """Physics.elm:590-598."""
    return Config()


def contact_points(predicate, contacts):
    """Physics.elm:651-653, 1158-1207 (contactPointsHelp): world contact points on body1 of each pair group that has
    contacts, moved from the pre-step frame of body1 to its post-step frame.  The predicate is tried in both argument
    orders (1177), groups without contacts are skipped (1169-1171), and because the reference conses while walking
    `pairGroups` (itself the reverse of the enumeration) the result comes out in pair-enumeration order."""
    out = []
    if contacts.packed is None or contacts.bodies is None:
        return out
    from .pack import unpack_pair_groups
    by_id = {b.id: (ext, b) for ext, b in contacts.bodies}
    for g in unpack_pair_groups(contacts.packed):
        if not g["contacts"]:
            continue
        if g["id1"] not in by_id or g["id2"] not in by_id:
            continue
        ext1, b1 = by_id[g["id1"]]
        ext2, _ = by_id[g["id2"]]
        if not (predicate(ext1, ext2) or predicate(ext2, ext1)):
            continue
        # Transform3d.atOrigin |> relativeTo pre |> placeIn post, then pointPlaceIn of every contact.pi (1180-1192)
        transform = m3.t_place_in(b1.transform3d, m3.t_relative_to(g["pre_transform1"], (m3.ZERO, m3.IDENTITY_Q)))
        pts = [m3.point_place_in(transform, c["pi"]) for c in g["contacts"]]
        out.insert(0, (ext1, ext2, pts))
    return out


def simulate(config, bodies_with_ids, engine=None):
    """Physics.elm:522-583.  `List (id, Body)` in, `(List (id, Body), Contacts)` out.
    Id assignment, callback evaluation and (un)packing happen here; the step runs on the GPU."""
    from . import engine as eng
    e = engine if engine is not None else eng.default_engine()
    return e.simulate(config, bodies_with_ids)
