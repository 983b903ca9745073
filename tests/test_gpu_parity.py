"""GPU parity tests proper: the CUDA path through the C ABI (include/elm_physics_b200.h) against the
CPU oracle (oracle/, test infrastructure) on the same seeded inputs.  The bar is BIT-EXACT for the whole
state and the whole contact block (pair sets, island roots, keys, ni/pi/pj, lambdas, iteration counts),
which implies the north-star's <= 1e-9 per-step fp64 tolerance; `max_rel_dev` is reported for the record.

Reference behaviour under test: src/Physics.elm:522-583 (simulate), src/Internal/BroadPhase.elm,
src/Internal/NarrowPhase.elm + src/Collision/*.elm, src/Internal/Equation.elm, src/Internal/Islands.elm,
src/Internal/Solver.elm, src/Internal/SolverBody.elm.
"""
import numpy as np
import pytest

from elm_physics_b200 import abi
from elm_physics_b200 import pack
from elm_physics_b200 import physics as P
from elm_physics_b200 import scenes

import helpers as H

pytestmark = pytest.mark.gpu

REL_TOL = 1e-9   # north_star per-step relative tolerance; the tests below assert the stronger bit-exact bar


@pytest.fixture(scope="module")
def engine(request):
    """One context on cuda:0 for the whole module (the library has one device path and no CPU fallback)."""
    from elm_physics_b200 import engine as eng
    e = eng.Engine(0)
    yield e
    e.close()


def _packed(worlds):
    return pack.pack_worlds(worlds)


def test_library_is_the_cuda_build(engine):
    import os
    from elm_physics_b200 import engine as eng
    assert os.path.exists(eng.LIB_PATH)
    assert engine.stream_handle() is not None


def test_readme_scene_lockstep_1000_steps(engine, oracle):
    """C1: README.md:9-19 — rubber sphere on a wood plane, 1000 steps @ 60 Hz, compared every step."""
    shapes, bodies, params = _packed(scenes.readme_scene())
    bo, bg, _, _ = H.lockstep(oracle, engine, shapes, bodies, params, 1000, what="readme")
    assert H.max_rel_dev(bo, bg) <= REL_TOL
    # the ball has come to rest on the plane (z = radius within the contact tolerance)
    assert abs(bg.pos[0][2] - 0.5) < 5e-3


def test_stack_of_five_lockstep(engine, oracle):
    """sandbox/tests/StabilityTest.elm:339-360 scene (7 iterations, warm-started), 240 lockstep frames."""
    shapes, bodies, params = _packed(scenes.stack_scene(5, 7))
    bo, bg, _, wg = H.lockstep(oracle, engine, shapes, bodies, params, 240, what="stack5")
    # stable: max speed < 0.05 m/s as the reference's stability test requires
    assert float(np.abs(bg.vel).max()) < 0.05
    # one island containing the whole stack, rooted at the lowest dynamic id
    assert wg.world_n_islands[0] == 1


def test_boxes_scene_resident(engine, oracle):
    """C2: 100 boxes (pyramid + stack), 40 resident GPU steps (contacts fed back on device) == 40 oracle steps."""
    shapes, bodies, params = _packed(scenes.boxes_scene())
    _, bg, oo, og = H.resident_vs_oracle(oracle, engine, shapes, bodies, params, 40, what="boxes")
    assert og.n_contacts > 500
    assert og.world_n_islands[0] == 2


@pytest.mark.parametrize("seed,gravity", [(11, (0.0, 0.0, -9.80665)), (12, (0.0, 0.0, 0.0)), (13, (3.0, -2.0, -6.0))])
def test_narrowphase_matrix_pair_worlds(engine, oracle, seed, gravity):
    """Every (shape kind x shape kind) cell of NarrowPhase.elm:86-286 incl. compound, static, kinematic, particle:
    726 two-body worlds, 4 lockstep frames each."""
    worlds = H.pair_worlds(seed, per_combo=6, gravity=gravity)
    shapes, bodies, params = _packed(worlds)
    _, _, _, wg = H.lockstep(oracle, engine, shapes, bodies, params, 4, what=f"pairs{seed}")
    assert wg.n_contacts > 200     # the scene generator does make most pairs touch


def test_small_mixed_world_lockstep(engine, oracle):
    """Blocks, spheres, capsules, cylinders, cones, compounds, a point mass and a kinematic lift: 300 frames."""
    shapes, bodies, params = _packed(scenes.small_mixed_world(seed=3))
    H.lockstep(oracle, engine, shapes, bodies, params, 300, check_every=1, what="mixed")


def test_batched_mixed_worlds_resident(engine, oracle):
    """C3 at test size: 48 worlds x 32 bodies, 150 resident steps (drop + first settling)."""
    shapes, bodies, params = scenes.mixed_worlds(48)
    _, _, _, og = H.resident_vs_oracle(oracle, engine, shapes, bodies, params, 150, what="c3x48")
    assert og.n_contacts > 48 * 20


def test_mixed_worlds_far_from_the_origin(engine, oracle):
    """The same piles 20 km away from the origin (translated in x / y; the floor plane is z = 0): the fp32 screening of the
    SAT axes works on vertices relative to the pair, so large coordinates must neither break its bound nor change any
    result -- bit-exact against the oracle, and cylinder-cylinder contacts (the screened path) do occur."""
    shapes, bodies, params = scenes.mixed_worlds(12)
    bodies.pos[:, 0] += 20000.0
    bodies.pos[:, 1] -= 12345.678
    _, _, _, og = H.resident_vs_oracle(oracle, engine, shapes, bodies, params, 200, what="far c3x12")
    assert og.n_contacts > 12 * 20


def test_cylinder_ties_lockstep(engine, oracle):
    """Hull-hull configurations with exactly tied SAT axes (the screened path must keep the reference's first-wins order):
    a coaxial tower of cylinders at exact rest (cap on cap: every side-face axis ties), two parallel cylinders lying side by
    side, a cylinder across another at a right angle, cone on cylinder; 240 frames in lock step."""
    import math
    bodies = [("floor", P.plane(P.wood))]
    for k in range(4):
        bodies.append((f"tower{k}", P.move_to((0.0, 0.0, 0.5 + k * 1.0), P.cylinder(0.5, 1.0, P.wood))))
    lying = P.rotate_around((0, 0, 0), (0, 1, 0), math.pi / 2, P.cylinder(0.5, 1.0, P.steel))
    bodies.append(("lie0", P.move_to((3.0, 0.0, 0.5), lying)))
    bodies.append(("lie1", P.move_to((3.0, 1.0, 0.5), lying)))
    bodies.append(("cross", P.move_to((3.0, 0.5, 1.5), P.rotate_around((0, 0, 0), (1, 0, 0), math.pi / 2, P.cylinder(0.4, 1.6, P.rubber)))))
    bodies.append(("base", P.move_to((-3.0, 0.0, 0.5), P.cylinder(0.6, 1.0, P.wood))))
    bodies.append(("cone", P.move_to((-3.0, 0.0, 1.5), P.cone(0.5, 1.0, P.plastic))))
    with_ids, _ = P.assign_ids(bodies)
    shapes, b, params = _packed([(P.on_earth(), with_ids)])
    H.lockstep(oracle, engine, shapes, b, params, 240, check_every=1, what="cylinder ties")


def test_icosphere_and_cube_meshes_lockstep(engine, oracle):
    """tests/ConvexSphereSimTest.elm's scene (a 42-vertex icosphere body and a cube given as a triangle mesh, both
    `unsafeConvex`, over a floor) 200 frames in lock step: hulls beyond the screening limit take the exact SAT path."""
    import json
    import os
    mesh = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "icosphere.json")))
    cube_v = [(1, 1, -1), (1, -1, -1), (1, 1, 1), (1, -1, 1), (-1, 1, -1), (-1, -1, -1), (-1, 1, 1), (-1, -1, 1)]
    cube_f = [(0, 4, 6), (0, 6, 2), (3, 2, 6), (3, 6, 7), (7, 6, 4), (7, 4, 5), (5, 1, 3), (5, 3, 7), (1, 0, 2), (1, 2, 3),
              (5, 4, 0), (5, 0, 1)]
    cube = P.scale_mass_to(5.0, P.dynamic([(P.shape_unsafe_convex(cube_f, cube_v), P.wood)]))
    ico = P.scale_mass_to(5.0, P.dynamic([(P.shape_unsafe_convex([tuple(f) for f in mesh["faces"]],
                                                                 [tuple(v) for v in mesh["vertices"]]), P.wood)]))
    bodies = [(0, P.move_to((0, 0, -1), P.plane(P.wood))), (1, P.move_to((0, 0, 8), cube)), (2, P.move_to((0.3, 0, 5), ico)),
              (3, P.move_to((0.2, 0.1, 11), cube)), (4, P.move_to((2.5, 0.3, 3), cube))]
    with_ids, _ = P.assign_ids(bodies)
    shapes, b, params = _packed([(P.on_earth(), with_ids)])
    H.lockstep(oracle, engine, shapes, b, params, 200, check_every=1, what="icosphere + cube meshes")


def test_batch_equals_single_worlds(engine):
    """Worlds are independent (src/Physics.elm:522-525): a world stepped inside a batch equals the same world
    stepped alone, bit for bit — the property the multi-GPU sharding relies on."""
    shapes, bodies, params = scenes.mixed_worlds(32)
    engine.upload_shapes(shapes)
    bb = bodies.copy()
    ob = H.contacts_for(bb)
    engine.step(bb, params, None, ob, 60)
    for w in (0, 17, 31):
        s1, b1, p1 = scenes.mixed_worlds(1, first_world=w)
        engine.upload_shapes(s1)
        o1 = H.contacts_for(b1)
        engine.step(b1, p1, None, o1, 60)
        r0, r1 = bodies.world_body_off[w], bodies.world_body_off[w + 1]
        for name in abi.BODY_INOUT:
            assert getattr(bb, name)[r0:r1].tobytes() == getattr(b1, name).tobytes(), (w, name)
        d_b, d_1 = pack.contacts_world_dict(ob, w), pack.contacts_world_dict(o1, 0)
        for k in ("feature_key", "shape_key", "lambda_", "ni", "pi", "pj", "group_id1", "group_id2"):
            assert d_b[k].tobytes() == d_1[k].tobytes(), (w, k)


def test_determinism_run_twice(engine):
    shapes, bodies, params = scenes.mixed_worlds(64)
    engine.upload_shapes(shapes)
    outs = []
    for _ in range(2):
        b = bodies.copy()
        o = H.contacts_for(b)
        engine.step(b, params, None, o, 80)
        outs.append((b.state_bytes(), {k: v.tobytes() for k, v in o.canonical().items()}))
    assert outs[0] == outs[1]


def _joint_world(kind):
    """Two/three-body worlds exercising Equation.elm:157-361 joint rows, incl. a flipped registration."""
    base = P.static([(P.shape_block((1.0, 1.0, 1.0)), P.wood)])
    base = P.move_to((0.0, 0.0, 3.0), base)
    arm = P.move_to((1.2, 0.1, 3.0), P.block((1.0, 0.4, 0.4), P.wood))
    arm = P.set_angular_velocity_to((0.3, 0.5, -0.2), arm)
    bob = P.move_to((2.4, 0.0, 2.5), P.sphere(0.3, P.steel))
    floor = P.plane(P.wood)
    bodies = [("floor", floor), ("base", base), ("arm", arm), ("bob", bob)]
    if kind == "hinge":
        cons = {("base", "arm"): [P.Hinge((0.5, 0, 0), (0, 1, 0), (-0.6, 0, 0), (0, 1, 0))],
                ("bob", "arm"): [P.PointToPoint((-0.5, 0, 0.3), (0.6, 0, 0))]}      # registered on the later body
    elif kind == "lock":
        cons = {("base", "arm"): [P.Lock((0.5, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (-0.6, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1))],
                ("arm", "bob"): [P.Distance(1.3)]}
    else:
        cons = {("arm", "base"): [P.PointToPoint((-0.6, 0, 0), (0.5, 0, 0)), P.Distance(1.25)],
                ("arm", "bob"): [P.Distance(1.4), P.PointToPoint((0.6, 0, 0), (-0.5, 0, 0.3))]}
    cfg = P.on_earth()

    def constrain(a):
        tab = {b: v for (x, b), v in cons.items() if x == a}
        return (lambda b: tab.get(b, [])) if tab else None
    cfg.constrain = constrain
    with_ids, _ = P.assign_ids(bodies)
    return [(cfg, with_ids)]


@pytest.mark.parametrize("kind", ["hinge", "lock", "mixed"])
def test_joint_constraints_lockstep(engine, oracle, kind):
    shapes, bodies, params = _packed(_joint_world(kind))
    assert params.n_constraints >= 2
    _, bg, _, wg = H.lockstep(oracle, engine, shapes, bodies, params, 200, what=f"joints-{kind}")
    assert int(wg.group_n_constraints[:wg.n_groups].sum()) >= 2
    assert np.isfinite(bg.pos).all()


def test_convex_pile_resident(engine, oracle):
    """C4 at test size: 150 convex hulls / compounds dropped into a walled pit (static block walls), 260 resident steps:
    islands of every size class up to ~150 contacts (lane tiles and both team kernels), general convex SAT + clipping,
    compound bodies."""
    shapes, bodies, params = _packed(scenes.pile_scene(150, layers=10))
    _, _, _, og = H.resident_vs_oracle(oracle, engine, shapes, bodies, params, 260, what="pile150")
    assert og.n_contacts > 400
    assert og.world_n_islands[0] < 80


def test_car_worlds_resident(engine, oracle):
    """C5 at test size: 24 car worlds (chassis compound + 4 hinged wheels + ramp + 6 hollow crates + tether),
    120 resident steps: joint rows inside lane-per-island tiles and team islands, batched."""
    shapes, bodies, params = _packed(scenes.car_worlds(24))
    assert params.n_constraints == 24 * 5
    _, bg, _, og = H.resident_vs_oracle(oracle, engine, shapes, bodies, params, 120, what="cars24")
    assert int(og.group_n_constraints[:og.n_groups].sum()) == 24 * 5
    assert np.isfinite(bg.pos).all()


def test_collide_callback_mask(engine, oracle):
    """`collide` evaluated caller-side (BroadPhase.elm:189-215): boxes that ignore each other fall through."""
    worlds = scenes.stack_scene(3, 10)
    cfg, bodies = worlds[0]
    cfg.collide = lambda a, b: not ({a, b} == {"b0", "b1"})
    shapes, pb, params = _packed(worlds)
    assert params.collide is not None
    _, bg, _, wg = H.lockstep(oracle, engine, shapes, pb, params, 90, what="collide")
    ids = {(int(wg.group_id1[g]), int(wg.group_id2[g])) for g in range(wg.n_groups)}
    assert (1, 2) not in ids and (2, 1) not in ids


def test_empty_ragged_and_sparse_ids(engine, oracle):
    """Edge cases of SURVEY §9.8: empty world, single static body, free fall (iterations = 1), sparse ids."""
    from dataclasses import replace
    w_empty = (P.on_earth(), [])
    w_plane = (P.on_earth(), P.assign_ids([("f", P.plane(P.wood))])[0])
    w_fall = (P.on_earth(), P.assign_ids([("a", P.move_to((0, 0, 50), P.sphere(0.5, P.rubber))), ("f", P.plane(P.wood))])[0])
    sparse = [("f", replace(P.plane(P.wood), id=7)), ("a", replace(P.move_to((0, 0, 0.45), P.block((1, 1, 1), P.wood)), id=3)),
              ("b", replace(P.move_to((0, 0, 1.4), P.block((1, 1, 1), P.wood)), id=40))]
    w_sparse = (P.on_earth(), P.assign_ids(sparse)[0])
    shapes, bodies, params = _packed([w_empty, w_plane, w_fall, w_sparse, w_empty])
    _, _, _, wg = H.lockstep(oracle, engine, shapes, bodies, params, 30, what="edge")
    assert list(wg.iterations[:5]) [0] == 0 and wg.iterations[2] == 1
    assert {int(x) for x in wg.group_id1[:wg.n_groups]} | {int(x) for x in wg.group_id2[:wg.n_groups]} == {3, 7, 40}


def test_warm_start_roundtrip_through_host(engine, oracle):
    """Contacts downloaded to the host and uploaded again as warm start == contacts kept resident on the device."""
    shapes, bodies, params = scenes.mixed_worlds(8)
    engine.upload_shapes(shapes)
    a = bodies.copy()
    oa = H.contacts_for(a)
    engine.step(a, params, None, oa, 90)
    b = bodies.copy()
    wb = None
    for _ in range(3):
        ob = H.contacts_for(b)
        engine.step(b, params, wb, ob, 30)
        wb = ob
    H.assert_bodies_equal(a, b, "resident vs host round trip")
    H.assert_contacts_equal(oa, wb, "resident vs host round trip")


def test_step_frame_host_buffers_lockstep(engine, oracle):
    """ep_step_frame (host buffers in / out, resident contacts as the warm start, pinned memory) == the oracle fed its own
    previous contacts, 60 frames, with forces applied on the host between frames (src/Physics.elm:857-1017)."""
    from elm_physics_b200.engine import PinnedPool
    shapes, bodies, params = scenes.mixed_worlds(6)
    pool = PinnedPool()
    try:
        bo, bg = bodies.copy(), bodies.copy(alloc=pool.empty)
        engine.upload_shapes(shapes)
        engine.upload_worlds(bg, params, None)
        og = abi.Contacts(bodies.n_worlds, H.contacts_for(bodies).group_cap, H.contacts_for(bodies).contact_cap, alloc=pool.empty)
        wo = None
        rng = np.random.RandomState(5)
        for s in range(60):
            if s % 7 == 3:
                rows = rng.randint(0, bodies.n_bodies, 9)
                push = rng.uniform(-300, 300, (9, 3))
                for b in (bo, bg):
                    b.force[rows] += push
                    b.torque[rows] += 0.1 * push
            oo = H.contacts_for(bodies)
            oracle.step(shapes, bo, params, wo, oo, 1)
            engine.step_frame(bg, og)
            H.assert_contacts_equal(oo, og, f"step_frame {s}")
            H.assert_bodies_equal(bo, bg, f"step_frame {s}")
            wo = oo
    finally:
        pool.close()


def test_step_frame_fields_masked_lockstep(engine, oracle):
    """ep_step_frame_fields: only the dynamic state goes up (EP_F_STATE) and comes back, forces / torques only on the frames
    that pushed bodies, constants never again after ep_upload_worlds -- 60 frames in lock step with the oracle, which is fed
    every array every frame.  The host copies of the arrays that are not downloaded are reconstructed the way a caller would:
    force / torque are zero after a step (SolverBody.elm:139-141, 221-223)."""
    from elm_physics_b200.engine import PinnedPool
    shapes, bodies, params = scenes.mixed_worlds(6)
    pool = PinnedPool()
    try:
        bo, bg = bodies.copy(), bodies.copy(alloc=pool.empty)
        engine.upload_shapes(shapes)
        engine.upload_worlds(bg, params, None)
        cap = H.contacts_for(bodies)
        og = abi.Contacts(bodies.n_worlds, cap.group_cap, cap.contact_cap, alloc=pool.empty)
        wo = None
        rng = np.random.RandomState(7)
        for s in range(60):
            up = abi.EP_F_STATE
            if s % 7 == 3:
                rows = rng.randint(0, bodies.n_bodies, 9)
                push = rng.uniform(-300, 300, (9, 3))
                for b in (bo, bg):
                    b.force[rows] += push
                    b.torque[rows] += 0.1 * push
                up |= abi.EP_F_FORCE | abi.EP_F_TORQUE
            oo = H.contacts_for(bodies)
            oracle.step(shapes, bo, params, wo, oo, 1)
            engine.step_frame(bg, og, upload=up, download=abi.EP_F_STATE)
            moving = bodies.kind != abi.EP_STATIC
            bg.force[moving] = 0.0; bg.torque[moving] = 0.0          # what the library guarantees for every non-static body; not read back
            H.assert_contacts_equal(oo, og, f"masked step_frame {s}")
            for name in ("pos", "quat", "vel", "angvel"):
                assert getattr(bo, name).tobytes() == getattr(bg, name).tobytes(), (s, name)
            wo = oo
        # invInertiaWorld was never read back on the way; on request it is the oracle's
        engine.step_frame(bg, og, upload=0, download=abi.EP_F_ALL)
        oo = H.contacts_for(bodies)
        oracle.step(shapes, bo, params, wo, oo, 1)
        H.assert_bodies_equal(bo, bg, "masked step_frame, full read-back")
    finally:
        pool.close()


def test_step_frame_contact_points_fields_only(engine):
    """ep_step_frame with only the arrays Physics.contactPoints reads (the other ep_contacts pointers NULL) returns the
    same bodies and, in those arrays, the same bytes as the call that copies every array back; the skipped arrays of the
    frame stay resident (the next frame's warm start is unaffected)."""
    shapes, bodies, params = scenes.mixed_worlds(5)
    cap = H.contacts_for(bodies)
    runs = []
    for fields in (None, abi.CONTACT_POINTS_FIELDS):
        b = bodies.copy()
        engine.upload_shapes(shapes)
        engine.upload_worlds(b, params, None)
        out = abi.Contacts(bodies.n_worlds, cap.group_cap, cap.contact_cap, fields=fields)
        for _ in range(40):
            engine.step_frame(b, out)
        runs.append((b, out))
    (bf, of), (bl, ol) = runs
    H.assert_bodies_equal(bf, bl, "lean vs full read-back")
    assert ol.shape_key is None and ol.lambda_ is None and ol.t1 is None
    ng, nc = of.n_groups, of.n_contacts
    assert (ol.n_groups, ol.n_contacts) == (ng, nc) and nc > 0
    assert ol.nbytes_out() < of.nbytes_out() // 2
    for name, n in (("group_id1", ng), ("group_id2", ng), ("group_pre_pos1", ng), ("group_pre_quat1", ng), ("pi", nc)):
        assert getattr(ol, name)[:n].tobytes() == getattr(of, name)[:n].tobytes(), name


def test_contact_pool_overflow_is_reported_not_fatal():
    """A contact pool that is too small (ep_configure) is reported as EP_ECAPACITY by ep_sync / the downloads -- never a
    crash, never silently -- and the context keeps working after it is given room again."""
    from elm_physics_b200 import engine as eng
    e = eng.Engine(0)
    try:
        shapes, bodies, params = scenes.mixed_worlds(8)
        e.upload_shapes(shapes)
        e.configure(contact_cap=64)
        e.upload_worlds(bodies.copy(), params, None)
        e.step_resident(150)                         # the pile needs ~600 contacts
        with pytest.raises(eng.EngineError, match="overflow"):
            e.sync()
        with pytest.raises(eng.EngineError):
            e.download_contacts(H.contacts_for(bodies))
        e.configure(contact_cap=1 << 16)
        b = bodies.copy()
        e.upload_worlds(b, params, None)
        e.step_resident(150)
        e.sync()
        out = H.contacts_for(bodies)
        e.download_contacts(out)
        assert out.n_contacts > 64
    finally:
        e.close()


def test_python_simulate_mirror(engine):
    """The host-side mirror of Physics.simulate (physics.simulate) steps the README scene like the reference's
    own doc example: contacts fed back, ball ends on the floor; contactPoints reports the touching point."""
    cfg, bodies = scenes.readme_scene()[0]
    contacts = P.empty_contacts()
    for _ in range(320):
        cfg.contacts = contacts
        bodies, contacts = P.simulate(cfg, bodies, engine=engine)
    ball = dict(bodies)["ball"]
    assert abs(P.origin_point(ball)[2] - 0.5) < 2e-2
    pts = P.contact_points(lambda a, b: True, contacts)
    assert len(pts) == 1 and abs(pts[0][2][0][2]) < 2e-2


def test_full_size_c3_properties(engine):
    """BASELINE config 3 at full size (8192 worlds x 32 bodies): size-independent properties —
    (1) worlds [0,16) of the full batch are bit-identical to the same worlds stepped as their own batch,
    (2) no capacity overflow, every world reports 1 <= iterations <= 20, state finite."""
    steps = 30
    shapes, bodies, params = scenes.mixed_worlds(8192)
    engine.upload_shapes(shapes)
    engine.upload_worlds(bodies, params, None)
    engine.step_resident(steps)
    engine.sync()
    engine.download_bodies(bodies)
    assert np.isfinite(bodies.pos).all() and np.isfinite(bodies.vel).all()
    s1, b1, p1 = scenes.mixed_worlds(16)
    engine.upload_shapes(s1)
    o1 = H.contacts_for(b1)
    engine.step(b1, p1, None, o1, steps)
    r1 = b1.n_bodies
    for name in abi.BODY_INOUT:
        assert getattr(bodies, name)[:r1].tobytes() == getattr(b1, name).tobytes(), name
    assert (o1.iterations[:16] >= 1).all() and (o1.iterations[:16] <= 20).all()


def test_upload_rejects_duplicate_ids_and_reports_item_overflow(engine):
    """The dense warm-start table equals bodyKey(id1, id2) (ContactId.elm:62-68) only for ids unique in a world: duplicates
    are refused (the caller runs AssignIds.assignIds first).  A narrowphase work-item pool that is too small is reported
    under its own name."""
    from elm_physics_b200 import engine as eng
    shapes, bodies, params = scenes.mixed_worlds(2)
    engine.upload_shapes(shapes)
    dup = bodies.copy()
    dup.body_id = dup.body_id.copy()          # Bodies.copy shares the topology arrays
    dup.body_id[3] = dup.body_id[2]
    dup._rebuild()
    with pytest.raises(eng.EngineError, match="duplicate body id"):
        engine.upload_worlds(dup, params, None)
    e = eng.Engine(0)
    try:
        e.upload_shapes(shapes)
        e.configure(item_cap=64)
        e.upload_worlds(bodies.copy(), params, None)
        e.step_resident(100)
        with pytest.raises(eng.EngineError, match="work-item pool overflow"):
            e.sync()
    finally:
        e.close()
