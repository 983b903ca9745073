"""Synthetic inputs for the BASELINE.json configs, built with the host-side mirror of the
reference's constructors.  Deterministic: SplitMix64 seeded with 0xE1F0 ^ (config << 32) ^ world index.
Small scenes are lists of (Config, [(ext_id, Body)]); the batched C3 scene is generated directly as
packed arrays (vectorised placement of template bodies) because 270k Python Body values would take minutes.
"""
import math

import numpy as np

from . import abi
from . import math3d as m3
from . import pack
from . import physics as P

MATERIALS = (P.wood, P.rubber, P.steel, P.ice, P.plastic)
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """Vectorised SplitMix64 finaliser over uint64 arrays."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def uniforms(seed, shape):
    """Uniform [0,1) doubles, counter-based: element k of stream `seed`."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        ctr = (np.uint64(seed) + np.arange(n, dtype=np.uint64) * np.uint64(0xD1342543DE82EF95)) & _M64
    bits = splitmix64(splitmix64(ctr))
    return ((bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)).reshape(shape)


def scene_seed(config, world=0):
    return 0xE1F0 ^ (config << 32) ^ world


def random_quaternions(u):
    """Shoemake: u (n,3) uniforms -> unit quaternions (n,4) x,y,z,w."""
    s1, s2 = np.sqrt(1.0 - u[:, 0]), np.sqrt(u[:, 0])
    a, b = 2 * math.pi * u[:, 1], 2 * math.pi * u[:, 2]
    q = np.stack([s1 * np.sin(a), s1 * np.cos(a), s2 * np.sin(b), s2 * np.cos(b)], axis=1)
    return q / np.sqrt((q * q).sum(axis=1, keepdims=True))


# ---- C1: README scene (README.md:9-19) ---------------------------------------------------------------
def readme_scene():
    ball = P.move_to((0, 0, 5), P.sphere(0.5, P.rubber))
    floor = P.plane(P.wood)
    bodies, _ = P.assign_ids([("ball", ball), ("floor", floor)])
    return [(P.on_earth(), bodies)]


# ---- C2: 100 unit wood boxes at rest: 2-D pyramid base 13 (91 boxes) + 9-box stack ---------------------
def boxes_scene(iterations=20):
    bodies = [("floor", P.plane(P.wood))]
    k = 0
    for row in range(13):
        for i in range(13 - row):
            bodies.append((f"p{k}", P.move_to((i + 0.5 * row, 0.0, row + 0.5), P.block((1, 1, 1), P.wood))))
            k += 1
    for j in range(9):
        bodies.append((f"s{j}", P.move_to((20.0, 0.0, j + 0.5), P.block((1, 1, 1), P.wood))))
    cfg = P.on_earth()
    cfg.solver_iterations = iterations
    with_ids, _ = P.assign_ids(bodies)
    return [(cfg, with_ids)]


def stack_scene(n=5, iterations=7):
    """sandbox/src/Stability/Scenarios.elm style stack of n unit boxes on a plane."""
    bodies = [("floor", P.plane(P.wood))]
    for j in range(n):
        bodies.append((f"b{j}", P.move_to((0.0, 0.0, j + 0.5), P.block((1, 1, 1), P.wood))))
    cfg = P.on_earth()
    cfg.solver_iterations = iterations
    with_ids, _ = P.assign_ids(bodies)
    return [(cfg, with_ids)]


# ---- C3: batched mixed worlds ---------------------------------------------------------------------------
MIXED_KINDS = ("block", "sphere", "capsule", "cylinder")


def _mixed_template(kind, mat):
    if kind == "block":
        return P.block((1, 1, 1), mat)
    if kind == "sphere":
        return P.sphere(0.5, mat)
    if kind == "capsule":
        return P.capsule(0.3, 1.0, mat)
    return P.cylinder(0.5, 1.0, mat)


def _qmul(a, b):
    """Transform3d.mul (Transform3d.elm:406-412), vectorised: a (n,4), b (4,)."""
    ax, ay, az, aw = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    bx, by, bz, bw = b
    return np.stack([ax * bw + ay * bz - az * by + aw * bx, -ax * bz + ay * bw + az * bx + aw * by,
                     ax * by - ay * bx + az * bw + aw * bz, -ax * bx - ay * by - az * bz + aw * bw], axis=1)


def _qrot(q, v):
    """Transform3d.rotate (Transform3d.elm:415-432), vectorised: q (n,4), v (3,)."""
    qx, qy, qz, qw = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    x, y, z = v
    ix, iy, iz = 2 * (qy * z - qz * y), 2 * (qz * x - qx * z), 2 * (qx * y - qy * x)
    return np.stack([x + qw * ix + qy * iz - qz * iy, y + qw * iy + qz * ix - qx * iz, z + qw * iz + qx * iy - qy * ix], axis=1)


def _iiw(q, inv):
    """Transform3d.invertedInertiaRotateIn (Transform3d.elm:181-205, 334-345), vectorised -> (n,9) m11,m21,m31,m12.."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    m11, m12, m13 = 1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * x * z + 2 * w * y
    m21, m22, m23 = 2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x
    m31, m32, m33 = 2 * x * z - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y
    a, b, c = inv[:, 0], inv[:, 1], inv[:, 2]
    return np.stack([
        m11 * m11 * a + m12 * m12 * b + m13 * m13 * c, m21 * m11 * a + m22 * m12 * b + m23 * m13 * c,
        m31 * m11 * a + m32 * m12 * b + m33 * m13 * c, m11 * m21 * a + m12 * m22 * b + m13 * m23 * c,
        m21 * m21 * a + m22 * m22 * b + m23 * m23 * c, m31 * m21 * a + m32 * m22 * b + m33 * m23 * c,
        m11 * m31 * a + m12 * m32 * b + m13 * m33 * c, m21 * m31 * a + m22 * m32 * b + m23 * m33 * c,
        m31 * m31 * a + m32 * m32 * b + m33 * m33 * c], axis=1)


def mixed_worlds(n_worlds, bodies_per_world=32, first_world=0, dt=1 / 60, iterations=20, config=3):
    """C3: each world = wood plane + `bodies_per_world` dynamic bodies cycling block / sphere r0.5 / capsule
    r0.3 l1 / 12-gon cylinder r0.5 l1, materials cycling wood/rubber/steel/ice/plastic, on a 4x4xL grid
    (spacing 1.5 m, z0 = 1.0 + 1.5*layer), jitter +-0.1 m, uniformly random orientation.
    Worlds are indexed first_world .. first_world+n_worlds-1 so shards of a batch are reproducible.
    Returns packed (Shapes, Bodies, Params)."""
    nb = bodies_per_world
    table = pack.ShapeTable()
    floor = P.plane(P.wood)
    templates = {}
    for k in MIXED_KINDS:
        for mi, mat in enumerate(MATERIALS):
            templates[(k, mi)] = _mixed_template(k, mat)
    floor_shape = table.add(floor.shapes[0][0])
    shape_of = {k: table.add(templates[(k, 0)].shapes[0][0]) for k in MIXED_KINDS}
    per = nb + 1
    n = n_worlds * per
    kind_idx = np.arange(nb) % 4
    mat_idx = np.arange(nb) % 5
    tb = [templates[(MIXED_KINDS[kind_idx[i]], int(mat_idx[i]))] for i in range(nb)]

    def col(fn, width):
        one = np.array([fn(floor)] + [fn(b) for b in tb], dtype=np.float64).reshape(per, width)
        return np.tile(one, (n_worlds, 1))

    f = {}
    f["inv_mass"] = col(lambda b: [b.inv_mass], 1)[:, 0]
    f["inv_inertia"] = col(lambda b: b.inv_inertia, 3)
    f["lin_damp_pow"] = col(lambda b: [(1.0 - b.linear_damping) ** dt], 1)[:, 0]
    f["ang_damp_pow"] = col(lambda b: [(1.0 - b.angular_damping) ** dt], 1)[:, 0]
    f["lin_lock"] = col(lambda b: b.linear_lock, 3)
    f["ang_lock"] = col(lambda b: b.angular_lock, 3)
    f["bounding_radius"] = col(lambda b: [b.bounding_sphere_radius], 1)[:, 0]
    for name in ("vel", "angvel", "force", "torque"):
        f[name] = np.zeros((n, 3))
    # placement: body frame (origin, q) composed with each template's centre-of-mass frame
    gi = np.arange(nb)
    grid = np.stack([(gi % 4) * 1.5, ((gi // 4) % 4) * 1.5, 1.0 + 1.5 * (gi // 16)], axis=1)
    pos = np.zeros((n_worlds, per, 3))
    quat = np.zeros((n_worlds, per, 4))
    quat[:, :, 3] = 1.0
    iiw = np.zeros((n_worlds, per, 9))
    world_ids = np.arange(first_world, first_world + n_worlds, dtype=np.uint64)
    seeds = np.array([scene_seed(config, int(w)) for w in world_ids], dtype=np.uint64)
    for i in range(nb):
        with np.errstate(over="ignore"):
            base = splitmix64(seeds + np.uint64(i) * np.uint64(0x2545F4914F6CDD1D))
        uu = np.stack([((splitmix64(base + np.uint64(j)) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0) for j in range(6)], axis=1)
        origin = grid[i] + (uu[:, 0:3] * 2.0 - 1.0) * 0.1
        q = random_quaternions(uu[:, 3:6])
        com_p, com_q = tb[i].center_of_mass_transform3d
        pos[:, i + 1, :] = origin + _qrot(q, com_p)          # Transform3d.placeIn (Transform3d.elm:222-230)
        qq = _qmul(q, com_q)
        quat[:, i + 1, :] = qq
        iiw[:, i + 1, :] = _iiw(qq, np.tile(np.array(tb[i].inv_inertia), (n_worlds, 1)))
    fm = floor.inv_inertia_world
    iiw[:, 0, :] = [fm[k] for k in pack.M_KEYS]
    pos[:, 0, :] = floor.transform3d[0]
    quat[:, 0, :] = floor.transform3d[1]
    f["pos"], f["quat"], f["inv_inertia_world"] = pos.reshape(n, 3), quat.reshape(n, 4), iiw.reshape(n, 9)
    body_id = np.tile(np.arange(per, dtype=np.int32), n_worlds)
    kind = np.tile(np.array([1] + [2] * nb, dtype=np.int32), n_worlds)
    inst_shape = np.tile(np.array([floor_shape] + [shape_of[MIXED_KINDS[k]] for k in kind_idx], dtype=np.int32), n_worlds)
    fr = np.tile(np.array([P.wood.friction] + [MATERIALS[m].friction for m in mat_idx]), n_worlds)
    bo = np.tile(np.array([P.wood.bounciness] + [MATERIALS[m].bounciness for m in mat_idx]), n_worlds)
    bodies = abi.Bodies(np.arange(n_worlds + 1, dtype=np.int32) * per, body_id, kind, np.arange(n + 1, dtype=np.int32),
                        inst_shape, fr, bo, **f)
    params = abi.Params(np.tile(np.array([0.0, 0.0, -P.GEE]), (n_worlds, 1)), np.full(n_worlds, dt),
                        np.full(n_worlds, iterations, np.int32))
    return table.build(), bodies, params


# ---- small mixed world built from Body values (tests: every shape kind, compound, kinematic, particle) ---
def small_mixed_world(seed=1, n=12, with_extras=True):
    rng = np.random.RandomState(seed)
    bodies = [("floor", P.plane(P.wood))]
    for i in range(n):
        mat = MATERIALS[i % 5]
        kind = i % 6
        if kind == 0:
            b = P.block((1.0, 0.8, 0.6), mat)
        elif kind == 1:
            b = P.sphere(0.45, mat)
        elif kind == 2:
            b = P.capsule(0.3, 1.0, mat)
        elif kind == 3:
            b = P.cylinder(0.5, 1.0, mat)
        elif kind == 4:
            b = P.cone(0.5, 1.0, mat)
        else:
            b = P.dynamic([(P.shape_block((0.6, 0.6, 0.6), center=(-0.5, 0, 0)), mat),
                           (P.shape_sphere(0.35, center=(0.5, 0, 0)), mat)])
        q = rng.normal(size=4)
        q = q / np.linalg.norm(q)
        origin = ((i % 3) * 1.1 + rng.uniform(-0.1, 0.1), ((i // 3) % 2) * 1.1 + rng.uniform(-0.1, 0.1), 0.8 + 0.9 * (i // 6) + rng.uniform(-0.1, 0.1))
        bodies.append((f"b{i}", P.place_quaternion(origin, tuple(float(c) for c in q), b)))
    if with_extras:
        bodies.append(("pm", P.point_mass((0.5, 0.5, 3.5), 2.0, P.steel)))
        lift = P.kinematic([(P.shape_block((1.5, 1.5, 0.2)), P.steel)])
        lift = P.set_velocity_to((0, 0, 0.3), P.move_to((4.0, 0.0, 0.5), lift))
        bodies.append(("lift", lift))
        bodies.append(("rider", P.move_to((4.0, 0.0, 1.2), P.block((0.5, 0.5, 0.5), P.wood))))
    with_ids, _ = P.assign_ids(bodies)
    return [(P.on_earth(), with_ids)]


# ---- C4: one large scene, convex polyhedra + compound bodies falling into a pit ---------------------------------
def _octahedron(r):
    v = [(r, 0, 0), (-r, 0, 0), (0, r, 0), (0, -r, 0), (0, 0, r), (0, 0, -r)]
    f = [(0, 2, 4), (2, 1, 4), (1, 3, 4), (3, 0, 4), (2, 0, 5), (1, 2, 5), (3, 1, 5), (0, 3, 5)]
    return P.shape_unsafe_convex(f, v)


def _square_pyramid(a, h):
    v = [(-a, -a, 0), (a, -a, 0), (a, a, 0), (-a, a, 0), (0, 0, h)]
    f = [(0, 3, 2), (0, 2, 1), (0, 1, 4), (1, 2, 4), (2, 3, 4), (3, 0, 4)]      # triangular mesh; coplanar faces are merged
    return P.shape_unsafe_convex(f, v)


def _pile_body(kind, mat):
    if kind == 0:
        return P.block((0.9, 0.7, 0.5), mat)
    if kind == 1:
        return P.cylinder(0.4, 0.9, mat)
    if kind == 2:
        return P.cone(0.45, 0.9, mat)
    if kind == 3:
        return P.dynamic([(_octahedron(0.5), mat)])
    if kind == 4:
        return P.dynamic([(_square_pyramid(0.45, 0.8), mat)])
    if kind == 5:      # block + sphere compound
        return P.dynamic([(P.shape_block((0.6, 0.6, 0.6), center=(-0.35, 0, 0)), mat), (P.shape_sphere(0.3, center=(0.35, 0, 0)), mat)])
    if kind == 6:      # dumbbell of three cylinders (Physics/Shape.elm:257-269 style)
        return P.dynamic([(P.shape_cylinder(12, 0.3, 0.25, center=(0, 0, -0.4)), mat), (P.shape_cylinder(12, 0.12, 0.6), mat),
                          (P.shape_cylinder(12, 0.3, 0.25, center=(0, 0, 0.4)), mat)])
    # hollow crate: five thin blocks (examples/src/RaycastCar.elm:150-172 style)
    return P.dynamic([(P.shape_block((0.8, 0.8, 0.1), center=(0, 0, -0.35)), mat), (P.shape_block((0.1, 0.8, 0.6), center=(-0.35, 0, 0)), mat),
                      (P.shape_block((0.1, 0.8, 0.6), center=(0.35, 0, 0)), mat), (P.shape_block((0.6, 0.1, 0.6), center=(0, -0.35, 0)), mat),
                      (P.shape_block((0.6, 0.1, 0.6), center=(0, 0.35, 0)), mat)])


def pile_scene(n_bodies=2000, iterations=20, seed=4, layers=None):
    """C4: plane + 4 static walls forming a square pit + `n_bodies` dynamic bodies (80 % convex hulls: block, 12-gon
    cylinder, 12-gon cone, octahedron, square pyramid; 20 % compounds: block+sphere, dumbbell, hollow crate) dropped from a
    lattice (spacing 2 m, jitter +-0.2 m, random orientation)."""
    side = max(2, int(math.ceil((n_bodies / float(layers or 20)) ** 0.5)))   # side x side x layers lattice, 20 layers at full size
    half = side + 1.0                                             # half width of the pit
    bodies = [("floor", P.plane(P.wood))]
    wall_h = 4.0
    for k, (cx, cy, sx, sy) in enumerate(((half + 0.5, 0, 1.0, 2 * half + 2), (-half - 0.5, 0, 1.0, 2 * half + 2),
                                          (0, half + 0.5, 2 * half + 2, 1.0), (0, -half - 0.5, 2 * half + 2, 1.0))):
        bodies.append((f"wall{k}", P.move_to((cx, cy, wall_h / 2), P.static([(P.shape_block((sx, sy, wall_h)), P.wood)]))))
    u = uniforms(scene_seed(4, seed), (n_bodies, 7))
    q = random_quaternions(u[:, 3:6])
    for i in range(n_bodies):
        r = i % 10
        kind = r if r < 5 else (r - 5 if r < 8 else r - 3)        # 0..4 hulls twice as often as the 3 compounds 5..7
        mat = MATERIALS[i % 5]
        ix, iy, iz = i % side, (i // side) % side, i // (side * side)
        pos = ((ix - (side - 1) / 2) * 2.0 + (u[i, 0] - 0.5) * 0.4, (iy - (side - 1) / 2) * 2.0 + (u[i, 1] - 0.5) * 0.4,
               1.5 + 2.0 * iz + (u[i, 2] - 0.5) * 0.4)
        bodies.append((f"b{i}", P.place_quaternion(pos, tuple(float(c) for c in q[i]), _pile_body(kind, mat))))
    cfg = P.on_earth()
    cfg.solver_iterations = iterations
    with_ids, _ = P.assign_ids(bodies)
    return [(cfg, with_ids)]


# ---- C5: RaycastCar-style worlds with hinge / point-to-point constraints -------------------------------------------
def car_world(seed=0, iterations=20):
    """Plane + static ramp (10 x 16 x 0.5 block tilted pi/16, sandbox/src/Car.elm:264-275) + chassis (compound of two
    blocks) + 4 wheel spheres hinged to it (sandbox/src/Car.elm:177-254) + 6 hollow crates + a point-to-point tether."""
    u = uniforms(scene_seed(5, seed), (16, 3))
    ramp = P.static([(P.shape_block((10.0, 16.0, 0.5)), P.wood)])
    ramp = P.move_to((0.0, 12.0, 0.6), P.rotate_around((0, 0, 0), (1, 0, 0), math.pi / 16, ramp))
    chassis = P.dynamic([(P.shape_block((2.0, 4.0, 0.6)), P.steel), (P.shape_block((1.6, 2.0, 0.5), center=(0, -0.3, 0.55)), P.plastic)])
    chassis = P.scale_mass_to(10.0, chassis)
    cx, cy, cz = (u[0, 0] - 0.5) * 0.4, (u[0, 1] - 0.5) * 0.4, 1.6
    chassis = P.set_velocity_to((0.0, 2.0 + u[0, 2], 0.0), P.move_to((cx, cy, cz), chassis))
    bodies = [("floor", P.plane(P.wood)), ("ramp", ramp), ("chassis", chassis)]
    cons = {}
    for k, (wx, wy) in enumerate(((-1.3, 1.4), (1.3, 1.4), (-1.3, -1.4), (1.3, -1.4))):
        wheel = P.scale_mass_to(2.0, P.sphere(0.6, P.rubber))
        wheel = P.move_to((cx + wx, cy + wy, cz - 0.5), wheel)
        bodies.append((f"wheel{k}", wheel))
        cons[("chassis", f"wheel{k}")] = [P.Hinge((wx, wy, -0.5), (1, 0, 0), (0, 0, 0), (1, 0, 0))]
    for k in range(6):
        crate = _pile_body(7, P.wood)
        crate = P.move_to((-3.0 + 1.2 * k + (u[4 + k, 0] - 0.5) * 0.2, 6.0 + (u[4 + k, 1] - 0.5) * 0.2, 1.0 + 0.9 * (k % 2)), crate)
        bodies.append((f"crate{k}", crate))
    cons[("chassis", "crate0")] = [P.PointToPoint((0.0, 2.0, 0.0), (0.0, -0.4, 0.0))]       # tether
    cfg = P.on_earth()
    cfg.solver_iterations = iterations

    def constrain(a):
        tab = {b: v for (x, b), v in cons.items() if x == a}
        return (lambda b: tab.get(b, [])) if tab else None
    cfg.constrain = constrain
    with_ids, _ = P.assign_ids(bodies)
    return (cfg, with_ids)


def car_worlds(n_worlds, first_world=0, iterations=20):
    return [car_world(first_world + w, iterations) for w in range(n_worlds)]


def replicate_worlds(shapes, bodies, params, times):
    """Tile a packed batch `times` times (same shapes table): the cheap way to a large batch of structurally identical
    worlds (C5 at full size) without building every world through the Python object layer."""
    from . import abi
    assert params.collide is None
    nb, nw, ni = bodies.n_bodies, bodies.n_worlds, bodies.n_insts
    wbo = np.concatenate([bodies.world_body_off[:-1] + k * nb for k in range(times)] + [[times * nb]]).astype(np.int32)
    ioff = np.concatenate([bodies.inst_off[:-1] + k * ni for k in range(times)] + [[times * ni]]).astype(np.int32)
    cols = {name: np.concatenate([getattr(bodies, name)] * times, axis=0) for name, _ in abi.BODY_F64_FIELDS}
    b = abi.Bodies(wbo, np.tile(bodies.body_id, times), np.tile(bodies.kind, times), ioff, np.tile(bodies.inst_shape, times),
                   np.tile(bodies.inst_friction, times), np.tile(bodies.inst_bounciness, times), **cols)
    nc = params.n_constraints
    p = abi.Params(np.tile(params.gravity, (times, 1)), np.tile(params.dt, times), np.tile(params.iterations, times),
                   con_body_a=np.concatenate([params.con_body_a + k * nb for k in range(times)]) if nc else (),
                   con_body_b=np.concatenate([params.con_body_b + k * nb for k in range(times)]) if nc else (),
                   con_kind=np.tile(params.con_kind, times) if nc else (),
                   con_data=np.tile(params.con_data, (times, 1)) if nc else ())
    return shapes, b, p


def slice_batch(bodies, params, worlds):
    """The listed worlds of a packed batch as a batch of their own (rows, shape instances, constraint endpoints rebased;
    same shape table): worlds are independent `simulate` calls (src/Physics.elm:522-525), so a sample of a large batch can
    be stepped on its own -- by the CPU oracle in the full-size parity tests -- and compared with the same worlds of the
    whole batch.  Batches with `collide` masks are not supported."""
    from . import abi
    assert params.collide is None
    worlds = [int(w) for w in worlds]
    rows = np.concatenate([np.arange(bodies.world_body_off[w], bodies.world_body_off[w + 1]) for w in worlds]).astype(np.int64)
    new_row = -np.ones(bodies.n_bodies, np.int64)
    new_row[rows] = np.arange(len(rows))
    counts = [int(bodies.world_body_off[w + 1] - bodies.world_body_off[w]) for w in worlds]
    wbo = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    icount = (bodies.inst_off[rows + 1] - bodies.inst_off[rows]).astype(np.int64)
    ioff = np.concatenate([[0], np.cumsum(icount)]).astype(np.int32)
    inst = np.concatenate([np.arange(bodies.inst_off[r], bodies.inst_off[r + 1]) for r in rows]).astype(np.int64) if len(rows) else np.zeros(0, np.int64)
    b = abi.Bodies(wbo, bodies.body_id[rows], bodies.kind[rows], ioff, bodies.inst_shape[inst], bodies.inst_friction[inst],
                   bodies.inst_bounciness[inst], **{name: getattr(bodies, name)[rows] for name, _ in abi.BODY_F64_FIELDS})
    keep = np.nonzero(new_row[params.con_body_a] >= 0)[0] if params.n_constraints else np.zeros(0, np.int64)
    p = abi.Params(params.gravity[worlds], params.dt[worlds], params.iterations[worlds],
                   con_body_a=new_row[params.con_body_a[keep]] if len(keep) else (), con_body_b=new_row[params.con_body_b[keep]] if len(keep) else (),
                   con_kind=params.con_kind[keep] if len(keep) else (), con_data=params.con_data[keep] if len(keep) else ())
    return b, p, rows


def car_batch(n_worlds, distinct=64, seed=55):
    """C5 at any size without building every world through the Python object layer: `distinct` car worlds (car_world seeds
    0..distinct-1) tiled to n_worlds, then every world made different -- each dynamic body gets its own small seeded
    velocity and spin kick -- so that a sample of worlds is a real test of the whole batch."""
    from . import pack
    distinct = min(distinct, n_worlds)
    assert n_worlds % distinct == 0
    shapes, bodies, params = pack.pack_worlds(car_worlds(distinct))
    shapes, bodies, params = replicate_worlds(shapes, bodies, params, n_worlds // distinct)
    u = uniforms(scene_seed(5, seed), (bodies.n_bodies, 6))
    dyn = (bodies.kind == 2)[:, None]
    bodies.vel[...] = bodies.vel + dyn * (u[:, 0:3] - 0.5) * 0.6
    bodies.angvel[...] = bodies.angvel + dyn * (u[:, 3:6] - 0.5) * 1.0
    return shapes, bodies, params
