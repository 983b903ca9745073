"""Physics.raycast (src/Physics.elm:802-841; next-row component of SURVEY 8f): the oracle against closed-form answers
(the reference ships no raycast test vectors), and the CUDA kernel (ep_raycast) against the oracle, bit for bit."""
import math

import numpy as np
import pytest

import helpers as H
from elm_physics_b200 import pack, scenes
from elm_physics_b200 import physics as P


def _world():
    bodies = [("floor", P.plane(P.wood)),
              ("ball", P.move_to((0.0, 0.0, 2.0), P.sphere(0.5, P.rubber))),
              ("box", P.move_to((3.0, 0.0, 1.0), P.block((1.0, 2.0, 2.0), P.wood))),
              ("cap", P.move_to((6.0, 0.0, 1.0), P.capsule(0.25, 1.0, P.steel))),          # axis along z, half length 0.5
              ("pm", P.point_mass((9.0, 0.0, 1.0), 1.0, P.steel)),
              ("ball2", P.move_to((0.0, 0.0, 4.0), P.sphere(0.5, P.rubber)))]
    with_ids, _ = P.assign_ids(bodies)
    return pack.pack_worlds([(P.on_earth(), with_ids)])


def test_oracle_raycast_closed_form_answers(oracle):
    shapes, bodies, _ = _world()
    down = (0.0, 0.0, -1.0)
    origins = [(0, 0, 10), (3, 0, 10), (6, 0, 10), (9, 0, 10), (20, 20, 10), (0, 0, -1), (3, -10, 1), (6, -5, 1), (6.1, -5, 1.6), (0, 0, 3)]
    dirs = [down, down, down, down, down, (0, 0, 1), (0, 1, 0), (0, 1, 0), (0, 1, 0), (0, 0, 1)]
    hit, dist, point, normal = oracle.raycast(shapes, bodies, [0] * len(origins), origins, dirs)
    # 0: the upper ball first (z = 4.5), 1: box top face z = 2, 2: capsule top cap z = 1.75, 3: point mass ignored -> plane
    assert list(hit[:5]) == [5, 2, 3, 0, 0]
    assert dist[0] == 5.5 and dist[1] == 8.0 and abs(dist[2] - 8.25) < 1e-12 and dist[3] == 10.0 and dist[4] == 10.0
    assert tuple(normal[1]) == (0.0, 0.0, 1.0) and tuple(point[1]) == (3.0, 0.0, 2.0)
    assert np.allclose(normal[0], (0, 0, 1)) and np.allclose(normal[2], (0, 0, 1)) and tuple(normal[3]) == (0.0, 0.0, 1.0)
    # 5: from below the plane looking up: planes are one-sided (Plane.elm:26) but the ball above is hit at its bottom
    assert hit[5] == 1 and abs(dist[5] - 2.5) < 1e-12 and np.allclose(normal[5], (0, 0, -1))
    # 6: box side face y = -1; 7: capsule cylinder wall at y = -0.25; 8: capsule top cap region seen from the side
    assert hit[6] == 2 and dist[6] == 9.0 and tuple(normal[6]) == (0.0, -1.0, 0.0)
    assert hit[7] == 3 and abs(dist[7] - 4.75) < 1e-12 and np.allclose(normal[7], (0, -1, 0))
    assert hit[8] == 3 and abs(np.linalg.norm(normal[8]) - 1) < 1e-12 and normal[8][2] > 0
    # 9: starting inside nothing, between the balls, looking up: the upper ball's bottom at z = 3.5
    assert hit[9] == 5 and abs(dist[9] - 0.5) < 1e-12


def test_oracle_raycast_first_body_wins_exact_ties(oracle):
    a = P.move_to((0.0, 0.0, 1.0), P.block((2, 2, 2), P.wood))
    b = P.move_to((0.5, 0.0, 1.0), P.block((2, 2, 2), P.wood))          # same top face height
    with_ids, _ = P.assign_ids([("a", a), ("b", b)])
    shapes, bodies, _ = pack.pack_worlds([(P.on_earth(), with_ids)])
    hit, dist, _, _ = oracle.raycast(shapes, bodies, [0], [(0.25, 0, 5)], [(0, 0, -1)])
    assert hit[0] == 0 and dist[0] == 3.0
    miss, _, _, _ = oracle.raycast(shapes, bodies, [0], [(10, 10, 5)], [(0, 0, -1)])
    assert miss[0] == -1


def _random_rays(rng, bodies, per_world):
    n = bodies.n_worlds * per_world
    rw = np.repeat(np.arange(bodies.n_worlds, dtype=np.int32), per_world)
    centers = np.array([bodies.pos[bodies.world_body_off[w]:bodies.world_body_off[w + 1]].mean(axis=0) for w in range(bodies.n_worlds)])
    o = centers[rw] + rng.uniform(-6, 6, (n, 3)) + np.array([0, 0, 4.0])
    target = centers[rw] + rng.uniform(-2, 2, (n, 3))
    dr = target - o
    dr /= np.linalg.norm(dr, axis=1, keepdims=True)
    dr[::7] = (0.0, 0.0, -1.0)                                           # RaycastCar-style straight-down rays
    return rw, o, dr


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["mixed", "pile", "cars"])
def test_gpu_raycast_matches_oracle(oracle, scene):
    from elm_physics_b200 import engine as eng
    if scene == "mixed":
        shapes, bodies, params = pack.pack_worlds(scenes.small_mixed_world(seed=3))
        steps = 120
    elif scene == "pile":
        shapes, bodies, params = pack.pack_worlds(scenes.pile_scene(120, layers=8))
        steps = 150
    else:
        shapes, bodies, params = pack.pack_worlds(scenes.car_worlds(12))
        steps = 60
    e = eng.Engine(0)
    try:
        e.upload_shapes(shapes)
        e.upload_worlds(bodies, params, None)
        rng = np.random.RandomState(7)
        for phase in range(3):                       # rays between simulate calls, as examples/src/RaycastCar does
            e.step_resident(steps // 3)
            e.download_bodies(bodies)
            rw, o, dr = _random_rays(rng, bodies, 64)
            got = e.raycast(rw, o, dr)
            want = oracle.raycast(shapes, bodies, rw, o, dr)
            assert (got[0] >= 0).sum() > len(rw) // 4
            for g, w, name in zip(got, want, ("hit_body", "distance", "point", "normal")):
                assert g.tobytes() == w.tobytes(), f"{scene} phase {phase}: {name} differs"
    finally:
        e.close()


@pytest.mark.gpu
def test_gpu_body_mutators_match_oracle(oracle):
    """ep_apply (applyForce / applyImpulse / applyTorque / applyAngularImpulse / setVelocityTo / setAngularVelocityTo on the
    resident state, several operations per body in list order) == the oracle's mutators on host buffers, then 40 more
    steps on both sides stay bit-identical."""
    from elm_physics_b200 import engine as eng
    shapes, bodies, params = pack.pack_worlds(scenes.small_mixed_world(seed=5))
    e = eng.Engine(0)
    try:
        e.upload_shapes(shapes)
        e.upload_worlds(bodies, params, None)
        bo = bodies.copy()
        oo, og = H.contacts_for(bodies), H.contacts_for(bodies)
        warm = None
        rng = np.random.RandomState(11)
        for rnd in range(4):
            n = 40
            kind = rng.randint(0, 6, n)
            rows = rng.randint(0, bodies.n_bodies, n)          # includes the static plane and the kinematic lift
            a = rng.uniform(-50, 50, (n, 3)); b = rng.uniform(-2, 2, (n, 3))
            e.apply(kind, rows, a, b)
            oracle.apply(bo, kind, rows, a, b)
            e.step_resident(10)
            nxt = H.contacts_for(bodies)
            oracle.step(shapes, bo, params, warm, nxt, 10)
            warm = nxt
            bg = bodies.copy()
            e.download_bodies(bg)
            H.assert_bodies_equal(bo, bg, f"mutators round {rnd}")
    finally:
        e.close()


def _place_rays(pos, quat, local_from, local_dir):
    """Transform3d.pointPlaceIn (Transform3d.elm:134-153) / directionPlaceIn -> rotate (218-220, 415-432) with numpy, one IEEE
    operation per reference operation in the reference's order (numpy never fuses), so the result equals the device's."""
    qx, qy, qz, qw = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]

    def place(v, origin):                     # Transform3d.pointPlaceIn
        x, y, z = v[:, 0], v[:, 1], v[:, 2]
        ix = qw * x + qy * z - qz * y
        iy = qw * y + qz * x - qx * z
        iz = qw * z + qx * y - qy * x
        iw = -qx * x - qy * y - qz * z
        return np.stack([ix * qw + iw * -qx + iy * -qz - iz * -qy + origin[:, 0],
                         iy * qw + iw * -qy + iz * -qx - ix * -qz + origin[:, 1],
                         iz * qw + iw * -qz + ix * -qy - iy * -qx + origin[:, 2]], axis=1)

    def rotate(v):                            # Transform3d.rotate: a DIFFERENT formula from pointPlaceIn
        x, y, z = v[:, 0], v[:, 1], v[:, 2]
        ix = 2 * (qy * z - qz * y)
        iy = 2 * (qz * x - qx * z)
        iz = 2 * (qx * y - qy * x)
        return np.stack([x + qw * ix + qy * iz - qz * iy, y + qw * iy + qz * ix - qx * iz, z + qw * iz + qx * iy - qy * ix], axis=1)
    return place(local_from, pos), rotate(local_dir)


@pytest.mark.gpu
def test_gpu_attached_rays_match_oracle(oracle):
    """ep_raycast_attached (the RaycastCar pattern, examples/src/RaycastCar/Car.elm:116-128): four rays per car given in the
    chassis frame, placed on the device with the chassis' current transform, cast against the bodies WITHOUT the car.
    Oracle side: the same rays placed on the host and cast against a copy of the state in which every chassis is moved out of
    reach (that is what `bodiesWithoutCar` amounts to).  Also: a second cast against the same state reuses the placed shapes."""
    from elm_physics_b200 import engine as eng
    shapes, bodies, params = pack.pack_worlds(scenes.car_worlds(12))
    e = eng.Engine(0)
    try:
        e.upload_shapes(shapes)
        b = bodies.copy()
        e.upload_worlds(b, params, None)
        e.step_resident(45)
        e.sync()
        e.download_bodies(b)
        nw = bodies.n_worlds
        chassis = np.array([int(bodies.world_body_off[w]) + 2 for w in range(nw)], dtype=np.int32)
        corners = np.array([[-0.9, 1.4, 0.3], [0.9, 1.4, 0.3], [-0.9, -1.4, 0.3], [0.9, -1.4, 0.3]])       # inside the chassis box: the own body must be skipped
        rb = np.repeat(chassis, 4)
        lf = np.tile(corners, (nw, 1))
        ld = np.tile(np.array([[0.0, 0.0, -1.0]]), (4 * nw, 1))
        launches0 = e.stats()["kernel_launches"]
        hg = [a.copy() for a in e.raycast_attached(rb, lf, ld)]
        launches1 = e.stats()["kernel_launches"]
        hg2 = [a.copy() for a in e.raycast_attached(rb, lf, ld)]
        assert e.stats()["kernel_launches"] - launches1 == 1 and launches1 - launches0 == 2       # k_place only before the first cast
        wf, wd = _place_rays(b.pos[rb], b.quat[rb], lf, ld)
        away = b.copy()
        away.pos[chassis] += 1.0e6
        ho = oracle.raycast(shapes, away, np.repeat(np.arange(nw, dtype=np.int32), 4), wf, wd)
        for k in range(4):
            assert hg[k].tobytes() == ho[k].tobytes(), k
            assert hg2[k].tobytes() == hg[k].tobytes(), k
        assert (hg[0] >= 0).mean() > 0.8 and not np.isin(hg[0], chassis).any()
    finally:
        e.close()
