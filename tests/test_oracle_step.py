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
