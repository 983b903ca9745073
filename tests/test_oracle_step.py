"""Behavioural checks of the CPU oracle's whole-step path (the reference pins this layer only through
end-to-end behaviour: tests/FrictionTest.elm, tests/KinematicTest.elm, sandbox/tests/StabilityTest.elm).
Whole-step fp64 trajectories are PARITY UNPINNED by the reference (no golden trajectories, no elm/node here);
these tests pin the physics the reference's own end-to-end tests assert."""
import math

import numpy as np

from elm_physics_b200 import pack
from elm_physics_b200 import physics as P
from elm_physics_b200 import scenes

import helpers as H


def run(oracle, worlds, steps):
    shapes, bodies, params = pack.pack_worlds(worlds)
    out = H.contacts_for(bodies)
    warm = None
    for _ in range(steps):
        o = H.contacts_for(bodies)
        oracle.step(shapes, bodies, params, warm, o, 1)
        warm = o
    return bodies, warm


def test_readme_ball_comes_to_rest(oracle):
    bodies, contacts = run(oracle, scenes.readme_scene(), 1000)
    assert abs(bodies.pos[0][2] - 0.5) < 5e-3 and abs(bodies.vel[0][2]) < 1e-2
    assert contacts.n_contacts == 1 and contacts.iterations[0] >= 1


def test_free_fall_reports_one_iteration(oracle):
    """Solver.elm:262-276: no pair groups -> iterations = max 1 0 = 1; body integrates under gravity with damping."""
    w = [(P.on_earth(), P.assign_ids([("a", P.move_to((0, 0, 100), P.sphere(0.5, P.rubber)))])[0])]
    bodies, contacts = run(oracle, w, 1)
    assert contacts.iterations[0] == 1 and contacts.n_groups == 0
    dt = 1 / 60
    assert math.isclose(bodies.vel[0][2], -P.GEE * dt, rel_tol=1e-12)
    assert math.isclose(bodies.pos[0][2], 100 - P.GEE * dt * dt, rel_tol=1e-12)


def test_stack_of_five_is_stable_warm_started(oracle):
    """sandbox/tests/StabilityTest.elm:339-360 at reduced length: 7 iterations, 63 warm-up frames with drift < 0.05 m,
    then 2000 frames with max speed < 0.05 m/s."""
    shapes, bodies, params = pack.pack_worlds(scenes.stack_scene(5, 7))
    start = bodies.pos.copy()
    warm = None
    for f in range(2063):
        o = H.contacts_for(bodies)
        oracle.step(shapes, bodies, params, warm, o, 1)
        warm = o
        if f < 63:
            assert float(np.abs(bodies.pos - start).max()) < 0.05
        else:
            assert float(np.sqrt((bodies.vel ** 2).sum(axis=1)).max()) < 0.05, f


def _slope_world(angle_deg, mu_body):
    """tests/FrictionTest.elm:237-300: box on a tilted plane (gravity rotated instead of the plane)."""
    a = math.radians(angle_deg)
    mat = P.Material(friction=mu_body, bounciness=0.0, density=700)
    box = P.move_to((0, 0, 0.5), P.block((1, 1, 1), mat))
    floor = P.plane(P.Material(friction=mu_body, bounciness=0.0, density=700))
    cfg = P.on_earth()
    cfg.gravity = (P.GEE * math.sin(a), 0.0, -P.GEE * math.cos(a))
    return [(cfg, P.assign_ids([("box", box), ("floor", floor)])[0])]


def test_coulomb_friction_holds_below_and_slides_above_the_critical_angle(oracle):
    """Combined friction sqrt(mu*mu) = mu (Material.elm:23-25); critical angle atan(mu)."""
    mu = 0.5
    crit = math.degrees(math.atan(mu))
    held, _ = run(oracle, _slope_world(crit - 2, mu), 120)
    assert abs(held.vel[0][0]) < 0.02
    slid, _ = run(oracle, _slope_world(crit + 6, mu), 120)
    a = math.radians(crit + 6)
    expected = P.GEE * (math.sin(a) - mu * math.cos(a)) * 2.0      # v = a t after 2 s
    assert slid.vel[0][0] > 0.5 * expected and slid.vel[0][0] < 1.15 * expected


def test_kinematic_body_moves_at_its_velocity_and_carries_a_box(oracle):
    """tests/KinematicTest.elm:28-67: kinematic bodies integrate their own velocity and are never pushed."""
    lift = P.kinematic([(P.shape_block((2.0, 2.0, 0.2)), P.steel)])
    lift = P.set_velocity_to((0, 0, 0.5), P.move_to((0, 0, 1.0), lift))
    rider = P.move_to((0, 0, 1.36), P.block((0.5, 0.5, 0.5), P.wood))
    w = [(P.on_earth(), P.assign_ids([("lift", lift), ("rider", rider)])[0])]
    bodies, contacts = run(oracle, w, 120)
    assert math.isclose(bodies.pos[0][2], 1.0 + 0.5 * 2.0, rel_tol=1e-9)
    assert tuple(bodies.vel[0]) == (0.0, 0.0, 0.5)
    assert bodies.pos[1][2] > 2.2      # carried upwards
    assert contacts.world_n_islands[0] == 1


def test_distance_constraint_registered_on_the_later_body_keeps_its_length(oracle):
    """Constraint.relativeToCenterOfMassFlipped maps `Distance length` to itself (Constraint.elm:93-95)."""
    anchor = P.move_to((0, 0, 5), P.static([(P.shape_sphere(0.1), P.steel)]))
    bob = P.move_to((1.0, 0, 5), P.sphere(0.2, P.steel))
    for registrant, partner in (("anchor", "bob"), ("bob", "anchor")):
        cfg = P.on_earth()
        cfg.constrain = (lambda reg, par: (lambda a: ((lambda b: [P.Distance(1.0)] if b == par else []) if a == reg else None)))(registrant, partner)
        bodies, contacts = run(oracle, [(cfg, P.assign_ids([("anchor", anchor), ("bob", bob)])[0])], 240)
        d = float(np.linalg.norm(bodies.pos[1] - bodies.pos[0]))
        assert abs(d - 1.0) < 0.05, (registrant, d)
        assert contacts.group_n_constraints[0] == 1


# ---- tests/ConvexSphereSimTest.elm:32-50 -----------------------------------------------------------------
def test_icosphere_body_bounces_off_the_floor_and_does_not_pass_through_the_cube(oracle):
    """Floor at z = -1, a 2 m cube body (5 kg) dropped from z = 8 onto an icosphere body (5 kg, the 42-vertex mesh of
    tests/ConvexSphereSimTest.elm:142-277) dropped from z = 5 at x = 0.3: over 200 frames the two centres never come closer
    than 0.4 m (the sum of the half extents is ~2 m; deep penetration / tunnelling would be < 0.5)."""
    import json
    import os
    mesh = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "icosphere.json")))
    cube_v = [(1, 1, -1), (1, -1, -1), (1, 1, 1), (1, -1, 1), (-1, 1, -1), (-1, -1, -1), (-1, 1, 1), (-1, -1, 1)]
    cube_f = [(0, 4, 6), (0, 6, 2), (3, 2, 6), (3, 6, 7), (7, 6, 4), (7, 4, 5), (5, 1, 3), (5, 3, 7), (1, 0, 2), (1, 2, 3),
              (5, 4, 0), (5, 0, 1)]
    cube = P.scale_mass_to(5.0, P.dynamic([(P.shape_unsafe_convex(cube_f, cube_v), P.wood)]))
    ico = P.scale_mass_to(5.0, P.dynamic([(P.shape_unsafe_convex([tuple(f) for f in mesh["faces"]],
                                                                 [tuple(v) for v in mesh["vertices"]]), P.wood)]))
    bodies = [(0, P.move_to((0, 0, -1), P.plane(P.wood))), (1, P.move_to((0, 0, 8), cube)), (2, P.move_to((0.3, 0, 5), ico))]
    with_ids, _ = P.assign_ids(bodies)
    shapes, b, params = pack.pack_worlds([(P.on_earth(), with_ids)])
    warm, min_dist = None, 1e9
    for _ in range(200):
        o = H.contacts_for(b)
        oracle.step(shapes, b, params, warm, o, 1)
        warm = o
        min_dist = min(min_dist, float(np.linalg.norm(b.pos[1] - b.pos[2])))
    assert min_dist > 0.4, f"min cube-ico centre distance was {min_dist} m (deep penetration / tunnelling)"
    assert b.pos[2][2] > -1.0 and b.pos[1][2] > b.pos[2][2]        # the icosphere rests on the floor, the cube on top of it or beside


# ---- tests/FrictionTest.elm:184-369 with the reference's own scene, forces, frame counts and tolerances ------------------
_G = 9.80665
_BOX_MASS = 10.0


def _friction_scene(material):
    box = P.move_to((0, 0, 0.5), P.scale_mass_to(_BOX_MASS, P.block((1, 1, 1), material)))
    return P.assign_ids([(0, P.plane(material)), (1, box)])[0]


def _friction_run(oracle, gravity, force, n, state):
    """FrictionTest.run: n frames, force (fx, fz) re-applied at the box's centre of mass before every frame; the contact
    cache is threaded through the frames of ONE run (each run starts from emptyContacts, like the reference's helpers)."""
    shapes, bodies, params = state
    params.gravity[0] = gravity
    warm = None
    for _ in range(n):
        if force != (0.0, 0.0):
            bodies.force[1] += (force[0], 0.0, force[1])
        o = H.contacts_for(bodies)
        oracle.step(shapes, bodies, params, warm, o, 1)
        warm = o
    return state


def _push_vx_fz(oracle, fx, fz):
    state = pack.pack_worlds([(P.on_earth(), _friction_scene(P.wood))])
    _friction_run(oracle, (0.0, 0.0, -_G), (0.0, 0.0), 120, state)
    _friction_run(oracle, (0.0, 0.0, -_G), (float(fx), float(fz)), 30, state)
    return float(state[1].vel[1][0])


def _slide_vx(oracle, material, degs, frames):
    r = math.radians(degs)
    state = pack.pack_worlds([(P.on_earth(), _friction_scene(material))])
    _friction_run(oracle, (_G * math.sin(r), 0.0, -_G * math.cos(r)), (0.0, 0.0), frames, state)
    return float(state[1].vel[1][0])


def test_friction_flat_floor_horizontal_push(oracle):
    f_max = P.wood.friction * _BOX_MASS * _G
    assert abs(_push_vx_fz(oracle, 0, 0)) < 0.005                      # settles to rest under gravity alone
    assert abs(_push_vx_fz(oracle, 0.5 * f_max, 0)) < 0.005            # half Fmax: static friction holds
    force = 2 * f_max
    expected = (force / _BOX_MASS - P.wood.friction * _G) * 0.5         # a = (F - mu m g) / m over 30 frames
    assert abs(_push_vx_fz(oracle, force, 0) - expected) <= 0.05 * expected


def test_friction_clip_follows_the_applied_normal_load(oracle):
    slide_push = 1.5 * P.wood.friction * _BOX_MASS * _G
    alone = _push_vx_fz(oracle, slide_push, 0)
    assert alone > 0.3                                                  # slides under the gravity-only normal
    assert abs(_push_vx_fz(oracle, slide_push, -100)) < 0.01            # pressed down: the clip rises, it holds
    assert _push_vx_fz(oracle, slide_push, 100) > alone                 # lifted: the clip drops, it slides faster


def test_friction_inclined_contact(oracle):
    mu = P.wood.friction
    critical = math.degrees(math.atan(mu))
    assert abs(_slide_vx(oracle, P.wood, 15, 200)) < 0.005
    assert _slide_vx(oracle, P.wood, 24, 200) > 0.5
    breakaway = min(d for d in (18 + 0.5 * i for i in range(13)) if not abs(_slide_vx(oracle, P.wood, d, 200)) < 0.01)
    assert abs(breakaway - critical) <= 1.0
    r = math.radians(30)
    measured = (_slide_vx(oracle, P.wood, 30, 200) - _slide_vx(oracle, P.wood, 30, 100)) / (100 / 60)
    expected = _G * (math.sin(r) - mu * math.cos(r))
    assert abs(measured - expected) <= 0.05 * expected


def test_friction_ice_slides_where_wood_holds(oracle):
    assert abs(_slide_vx(oracle, P.wood, 10, 200)) < 0.005
    assert _slide_vx(oracle, P.ice, 10, 200) > 1


# ---- tests/KinematicTest.elm:27-135, the four cases as written there -------------------------------------------------------
def test_kinematic_cases_of_the_reference(oracle):
    ball = [(P.shape_sphere(0.5), P.wood)]
    platform = P.set_velocity_to((1, 0, 0), P.move_to((0, 0, 5), P.kinematic(ball)))
    bodies, _ = run(oracle, [(P.on_earth(), P.assign_ids([("floor", P.plane(P.wood)), ("platform", platform)])[0])], 1)
    assert abs(bodies.pos[1][0] - 1 / 60) < 1e-4 and abs(bodies.pos[1][2] - 5) < 1e-4          # translates by velocity * dt
    bodies, _ = run(oracle, [(P.on_earth(), P.assign_ids([("platform", P.set_velocity_to((1, 0, 0), P.kinematic(ball)))])[0])], 1)
    assert abs(bodies.vel[0][0] - 1) < 1e-4 and abs(bodies.vel[0][1]) < 1e-4 and abs(bodies.vel[0][2]) < 1e-4   # gravity does not touch it
    moved = P.move_to((100, 0, 5), P.set_velocity_to((2, 0, 0), P.kinematic(ball)))
    assert abs(moved.velocity[0] - 2) < 1e-4                                                   # moveTo keeps the velocity
    wall = P.set_velocity_to((5, 0, 0), P.static(ball))
    assert abs(wall.velocity[0]) < 1e-4                                                        # static bodies ignore setVelocityTo


# ---- sandbox/tests/StabilityTest.elm:383-447 at reduced length (the reference runs 100 000 frames) ------------------------
def test_box_resting_on_a_slope_creeps_but_never_slides(oracle):
    """sandbox/src/Stability/Scenarios.elm:107-133: a wood box flush on a static wood ramp tilted 15 degrees, 10 solver
    iterations.  The reference's limits: max speed < 0.05 m/s, drift < 0.7 m over 100 000 frames, and it documents what it
    observes with friction working: a peak speed of ~4e-4 m/s and a steady creep of ~0.27 mm/s (0.59 m over 100 000
    frames).  6000 frames here: the same speed limit, and the creep RATE has to agree with the documented one."""
    th = math.radians(15)
    box = P.move_to((0.5 * math.sin(th), 0, 2 + 0.5 * math.cos(th)), P.rotate_around((0, 0, 0), (0, 1, 0), th, P.block((1, 1, 1), P.wood)))
    ramp = P.static([(P.shape_block((10.0, 6.0, 0.5)), P.wood)])
    ramp = P.move_to((-0.25 * math.sin(th), 0, 2 - 0.25 * math.cos(th)), P.rotate_around((0, 0, 0), (0, 1, 0), th, ramp))
    cfg = P.on_earth()
    cfg.solver_iterations = 10
    shapes, bodies, params = pack.pack_worlds([(cfg, P.assign_ids([(1, box), (0, ramp)])[0])])
    start = bodies.pos[0].copy()
    warm, max_speed, settled_speed, frames = None, 0.0, 0.0, 6000
    for f in range(frames + 1):
        o = H.contacts_for(bodies)
        oracle.step(shapes, bodies, params, warm, o, 1)
        warm = o
        if f >= 1:              # frame 0 is skipped: initial-contact velocity spike (StabilityTest.elm:391-395)
            max_speed = max(max_speed, float(np.linalg.norm(bodies.vel[0])))
        if f >= 100:
            settled_speed = max(settled_speed, float(np.linalg.norm(bodies.vel[0])))
    drift = float(np.linalg.norm(bodies.pos[0] - start))
    assert max_speed < 0.05, max_speed                      # the reference's limit
    assert 2.0e-4 < settled_speed < 4.5e-4, settled_speed   # the reference's observation: creep ~0.27 mm/s, peak ~4e-4 m/s
    rate_mm_s = 1e3 * drift / (frames / 60.0)
    assert 0.15 < rate_mm_s < 0.45, rate_mm_s


def test_dropped_stack_of_five_lands_without_toppling(oracle):
    """Scenarios.elm:178-200 + StabilityTest.elm:420-447: boxes released with 0.4 m of air between them, 10 iterations: no
    box ever strays more than 0.25 m sideways and the stack is at rest (max speed < 0.05 m/s) at the end; 3000 frames."""
    bodies = [(k, P.move_to((0, 0, k - 0.5 + k * 0.4), P.block((1, 1, 1), P.wood))) for k in range(5, 0, -1)] + [(0, P.plane(P.wood))]
    cfg = P.on_earth()
    cfg.solver_iterations = 10
    shapes, b, params = pack.pack_worlds([(cfg, P.assign_ids(bodies)[0])])
    warm = None
    for f in range(3000):
        o = H.contacts_for(b)
        oracle.step(shapes, b, params, warm, o, 1)
        warm = o
        assert float(np.hypot(b.pos[:5, 0], b.pos[:5, 1]).max()) <= 0.25, f
    assert float(np.sqrt((b.vel[:5] ** 2).sum(axis=1)).max()) < 0.05
    assert np.allclose(np.sort(b.pos[:5, 2]), [0.5, 1.5, 2.5, 3.5, 4.5], atol=0.02)


# ---- SURVEY §9.8 edge cases on the oracle (the GPU suite compares the same kinds of scenes bit for bit) ----------------------
def test_two_dynamic_bodies_on_the_same_static_body_are_separate_islands(oracle):
    """Islands.elm:3-7, Solver.elm:201-206: only dynamic-dynamic pairs are united; the floor does not connect what rests on it."""
    bodies = [("floor", P.plane(P.wood)),
              ("a", P.move_to((-3, 0, 0.5), P.block((1, 1, 1), P.wood))),
              ("b", P.move_to((3, 0, 0.5), P.block((1, 1, 1), P.wood)))]
    _, contacts = run(oracle, [(P.on_earth(), P.assign_ids(bodies)[0])], 3)
    assert contacts.n_groups == 2 and contacts.world_n_islands[0] == 2
    assert contacts.group_island[0] != contacts.group_island[1]


def test_touching_dynamic_bodies_share_an_island_and_a_kinematic_body_does_not_unite(oracle):
    """Islands.elm:23-66: box on box on the floor = one island of two groups; the same two boxes side by side on a kinematic
    platform stay two islands (kinematic-dynamic pairs make rows but never unions, Solver.elm:201-206)."""
    stack = [("floor", P.plane(P.wood)),
             ("lo", P.move_to((0, 0, 0.5), P.block((1, 1, 1), P.wood))),
             ("hi", P.move_to((0, 0, 1.5), P.block((1, 1, 1), P.wood)))]
    _, c1 = run(oracle, [(P.on_earth(), P.assign_ids(stack)[0])], 3)
    assert c1.n_groups == 2 and c1.world_n_islands[0] == 1
    platform = P.move_to((0, 0, 0.25), P.kinematic([(P.shape_block((8, 8, 0.5)), P.wood)]))
    side = [("platform", platform),
            ("a", P.move_to((-2, 0, 1.0), P.block((1, 1, 1), P.wood))),
            ("b", P.move_to((2, 0, 1.0), P.block((1, 1, 1), P.wood)))]
    _, c2 = run(oracle, [(P.on_earth(), P.assign_ids(side)[0])], 3)
    assert c2.n_groups == 2 and c2.world_n_islands[0] == 2


def test_point_masses_collide_with_shapes_but_not_with_each_other(oracle):
    """NarrowPhase.elm:235-237 (particle-particle: nothing), PlaneParticle.elm:10-33 (a particle rests on the plane)."""
    bodies = [("floor", P.plane(P.wood)),
              ("p", P.point_mass((0, 0, 0.0005), 1.0, P.wood)),
              ("q", P.point_mass((0, 0, 0.0005), 1.0, P.wood))]
    _, contacts = run(oracle, [(P.on_earth(), P.assign_ids(bodies)[0])], 2)
    ids = sorted((int(a), int(b)) for a, b in zip(contacts.group_id1[:contacts.n_groups], contacts.group_id2[:contacts.n_groups]))
    assert contacts.n_groups == 2 and contacts.n_contacts == 2          # each particle against the floor, nothing between them
    assert all(0 in pair for pair in ids)


def test_only_the_warm_start_of_the_previous_frame_is_read(oracle):
    """Physics.elm:528-529, 579: `simulate` reads cache.warmStart only -- a step warm-started by the previous frame's contacts
    gives the same bodies whatever that frame reported as iterations / islands."""
    shapes, bodies, params = pack.pack_worlds(scenes.readme_scene())
    warm = None
    for _ in range(40):
        o = H.contacts_for(bodies)
        oracle.step(shapes, bodies, params, warm, o, 1)
        warm = o
    b1, b2 = bodies.copy(), bodies.copy()
    o1, o2 = H.contacts_for(b1), H.contacts_for(b2)
    oracle.step(shapes, b1, params, warm, o1, 1)
    warm.iterations[:] = 17
    warm.world_n_islands[:] = 5
    oracle.step(shapes, b2, params, warm, o2, 1)
    H.assert_bodies_equal(b1, b2, "previous frame's iterations / islands must not matter")
