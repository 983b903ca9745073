"""Host-side logic that stays on the caller's side of the C ABI (id assignment, shape construction).
Ports tests/SimulateTest.elm:47-176 (exact expected ids) and spot checks of tests/BodyTest.elm /
tests/Shapes/ConvexTest.elm for the constructors the scene generators rely on."""
from dataclasses import replace

from elm_physics_b200 import math3d as m3
from elm_physics_b200 import physics as P
from elm_physics_b200 import shapes as sh


def ids(pairs):
    return [(e, b.id) for e, b in pairs]


def plane():
    return P.plane(P.wood)


def with_id(i, b):
    return replace(b, id=i)


def step(bodies):
    return P.assign_ids(bodies)[0]


def test_assign_ids_cases():
    assert ids(step([])) == []
    assert ids(step([("a", plane())])) == [("a", 0)]
    assert ids(step([("a", plane()), ("b", plane()), ("c", plane())])) == [("a", 0), ("b", 1), ("c", 2)]
    s1 = step([("a", plane()), ("b", plane())])
    assert ids(step(s1)) == ids(s1)
    assert ids(step(s1 + [("c", plane())])) == [("a", 0), ("b", 1), ("c", 2)]
    existing = [("a", with_id(0, plane())), ("b", with_id(2, plane()))]
    assert ids(step(existing + [("c", plane())])) == [("a", 0), ("b", 2), ("c", 1)]
    s1 = step([("a", plane())])
    assert ids(step([("b", s1[0][1])] + s1)) == [("b", 1), ("a", 0)]
    bodies = [("three", with_id(0, plane())), ("two", with_id(1, plane())), ("one", with_id(0, plane()))]
    assert ids(step(bodies)) == [("three", 2), ("two", 1), ("one", 0)]
    assert ids(step([("a", with_id(0, plane())), ("b", with_id(5, plane()))])) == [("a", 0), ("b", 5)]
    assert P.assign_ids([("a", with_id(0, plane())), ("b", with_id(5, plane()))])[1] == 5


def test_block_is_box_with_six_faces_three_edge_groups():
    c = sh.from_block(1, 2, 3)
    assert c.box is not None and len(c.faces) == 3 and all(g.two_sided for g in c.faces)
    assert len(c.unique_edges) == 3 and all(len(e) == 8 for e in c.unique_edges)
    assert c.box[3] == (1.5, 1.0, 0.5)      # he pairs with (z, y, x) group normals: Convex.elm:241-252


def test_cylinder_12_has_seven_face_and_edge_groups():
    c = sh.from_cylinder(12, 0.5, 1)
    # side faces recompute their ring points; the wrap-around can differ from the caps in the last bits,
    # so structural dedup may keep 1-2 near-duplicate vertices (Convex.elm:1013-1017) — kept as is.
    assert c.box is None and 24 <= len(c.vertices) <= 26
    assert len(c.faces) == 7 and all(g.two_sided for g in c.faces)
    assert len(c.unique_edges) == 7
    assert 36 <= sum(len(e) // 2 for e in c.unique_edges) <= 39


def test_block_body_mass_properties():
    b = P.block((1, 1, 1), P.wood)
    assert abs(b.mass - 700) < 1e-9
    assert abs(b.bounding_sphere_radius - (0.75 ** 0.5)) < 1e-12
    i = 700 / 12 * 2
    assert all(abs(1 / x - i) < 1e-9 for x in b.inv_inertia)


def test_sphere_and_plane_bodies():
    s = P.move_to((0, 0, 5), P.sphere(0.5, P.rubber))
    assert s.transform3d[0] == (0.0, 0.0, 5.0) and s.kind == 2
    f = P.plane(P.wood)
    assert f.kind == 1 and f.inv_mass == 0 and f.bounding_sphere_radius == m3.MAX_NUMBER


def test_contacts_optional_fields():
    """abi.Contacts(fields=...) leaves the unlisted arrays NULL in ep_contacts (include/elm_physics_b200.h: a NULL output
    array is not copied back); the two offset arrays are always held."""
    from elm_physics_b200 import abi
    full = abi.Contacts(2, 4, 8)
    lean = abi.Contacts(2, 4, 8, fields=abi.CONTACT_POINTS_FIELDS)
    for name in abi.ALL_CONTACT_ARRAYS:
        assert getattr(full, name) is not None
        held = name in abi.CONTACT_POINTS_FIELDS or name in ("world_group_off", "group_contact_off")
        assert (getattr(lean, name) is not None) == held, name
    assert not lean.c.lambda_ if hasattr(lean.c, "lambda_") else True
    assert lean.nbytes_out() <= full.nbytes_out()


# ---- tests/PlaceInTest.elm:20-119 (Physics.place / Physics.frame of the host-side mirror) -------------------
def _frame_close(f, origin, x, y, z, tol=1e-12):
    from elm_physics_b200 import math3d as m3
    o, q = f
    got = (o, m3.q_rotate(q, (1.0, 0.0, 0.0)), m3.q_rotate(q, (0.0, 1.0, 0.0)), m3.q_rotate(q, (0.0, 0.0, 1.0)))
    for a, b in zip(got, (origin, x, y, z)):
        assert all(abs(p - r) < tol for p, r in zip(a, b)), (got, (origin, x, y, z))


def test_place_sets_position_and_orientation_and_overwrites_previous_moves():
    import math
    from elm_physics_b200 import physics as P
    block = P.block((1.0, 1.0, 1.0), P.wood)
    X, Y, Z = (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)
    _frame_close(P.frame(P.place((0, 0, 0), X, Y, block)), (0, 0, 0), X, Y, Z)                 # place atOrigin is identity
    _frame_close(P.frame(P.place((1, 2, 3), X, Y, block)), (1, 2, 3), X, Y, Z)                 # place sets position
    rz = ((0.0, 1.0, 0.0), (-1.0, 0.0, 0.0))                                                    # rotateAround Axis3d.z 90 deg
    _frame_close(P.frame(P.place((0, 0, 0), rz[0], rz[1], block)), (0, 0, 0), rz[0], rz[1], Z)
    c, s = math.cos(math.pi / 4), math.sin(math.pi / 4)                                         # x axis, 45 deg, at (3, 4, 5)
    _frame_close(P.frame(P.place((3, 4, 5), X, (0.0, c, s), block)), (3, 4, 5), X, (0.0, c, s), (0.0, -s, c))
    moved = P.move_to((1, 2, 3), block)                                                         # place overwrites moveTo
    _frame_close(P.frame(P.place((10, 0, 0), X, Y, moved)), (10, 0, 0), X, Y, Z)
    turned = P.rotate_around((0, 0, 0), Z, math.pi / 2, block)                                  # ... and rotateAround
    c30, s30 = math.cos(math.pi / 6), math.sin(math.pi / 6)
    _frame_close(P.frame(P.place((0, 0, 0), (c30, 0.0, -s30), Y, turned)), (0, 0, 0), (c30, 0.0, -s30), Y, (s30, 0.0, c30))


def test_place_roundtrips_with_frame():
    import math
    from elm_physics_b200 import math3d as m3, physics as P
    block = P.block((1.0, 1.0, 1.0), P.wood)
    moved = P.rotate_around((0, 0, 0), (0.0, 0.0, 1.0), math.pi / 4, P.move_to((1, 2, 3), block))
    o, q = P.frame(moved)
    x, y = m3.q_rotate(q, (1.0, 0.0, 0.0)), m3.q_rotate(q, (0.0, 1.0, 0.0))
    o2, q2 = P.frame(P.place(o, x, y, block))
    assert all(abs(a - b) < 1e-12 for a, b in zip(o, o2))
    for axis in ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)):
        assert all(abs(a - b) < 1e-12 for a, b in zip(m3.q_rotate(q, axis), m3.q_rotate(q2, axis)))


# ---- tests/BodyTest.elm:16-152 (Body.compound of the host-side mirror: the bodies the C4 / C5 scenes are made of) --------
def _mat(density=700.0, friction=0.0, bounciness=0.0):
    from elm_physics_b200 import physics as P
    return P.Material(friction=friction, bounciness=bounciness, density=density)


def test_compound_bounding_sphere_radius():
    from elm_physics_b200 import math3d as m3, physics as P, shapes as sh
    box = lambda x, y, z: sh.from_block(x, y, z)
    assert P.compound(2, []).bounding_sphere_radius == 0
    r = P.compound(2, [(box(2, 2, 2), P.wood, 1)]).bounding_sphere_radius
    assert abs(r - m3.length((1, 1, 1))) < 1e-5
    r = P.compound(2, [(box(2, 2, 2), P.wood, 1), (box(4, 4, 4), P.wood, 1)]).bounding_sphere_radius
    assert abs(r - m3.length((2, 2, 2))) < 1e-5
    r = P.compound(2, [(sh.Plane(m3.ZERO, m3.Z_AXIS), P.wood, 1)]).bounding_sphere_radius
    assert r >= 3.40282347e38


def test_compound_inertia():
    from elm_physics_b200 import math3d as m3, physics as P, shapes as sh
    mat = _mat()
    two = P.compound(2, [(sh.convex_place_in(m3.t_at_point((-1.0, 0.0, 0.0)), sh.from_block(2, 2, 2)), mat, 1),
                         (sh.convex_place_in(m3.t_at_point((1.0, 0.0, 0.0)), sh.from_block(2, 2, 2)), mat, 1)])
    one = P.compound(2, [(sh.from_block(4, 2, 2), mat, 1)])
    assert sorted(round(v, 12) for v in two.inv_inertia) == sorted(round(v, 12) for v in one.inv_inertia)
    lx, ly, lz = 1.0, 1.5, 0.5          # blockOfTetrahedrons 2 3 1, tests/Fixtures/Convex.elm:62-96
    corners = [((lx, ly, -lz), (-lx, ly, lz), (lx, ly, lz)), ((lx, ly, -lz), (-lx, ly, -lz), (-lx, ly, lz)),
               ((lx, -ly, lz), (-lx, ly, lz), (-lx, -ly, lz)), ((lx, -ly, lz), (lx, ly, lz), (-lx, ly, lz)),
               ((-lx, -ly, lz), (-lx, ly, -lz), (-lx, -ly, -lz)), ((-lx, -ly, lz), (-lx, ly, lz), (-lx, ly, -lz)),
               ((-lx, -ly, -lz), (lx, -ly, lz), (-lx, -ly, lz)), ((-lx, -ly, -lz), (lx, -ly, -lz), (lx, -ly, lz)),
               ((lx, -ly, -lz), (lx, ly, lz), (lx, -ly, lz)), ((lx, -ly, -lz), (lx, ly, -lz), (lx, ly, lz)),
               ((-lx, -ly, -lz), (lx, ly, -lz), (lx, -ly, -lz)), ((-lx, -ly, -lz), (-lx, ly, -lz), (lx, ly, -lz))]
    tets = [sh.from_triangular_mesh([(1, 0, 2), (2, 0, 3), (3, 0, 1), (3, 1, 2)], [(0.0, 0.0, 0.0), p1, p2, p3])
            for p1, p2, p3 in corners]
    from_tets = P.compound(2, [(t, mat, 1) for t in tets])
    block = P.compound(2, [(sh.from_block(2, 3, 1), mat, 1)])
    for a, b in zip(sorted(from_tets.inv_inertia), sorted(block.inv_inertia)):
        assert abs(a - b) <= 1e-9 * abs(b)


def test_compound_volume_and_hollow_crate():
    from elm_physics_b200 import physics as P, shapes as sh
    assert abs(P.compound(2, [(sh.from_block(2, 3, 4), P.wood, 1)]).volume - 24) < 1e-5
    hollow = P.compound(2, [(sh.from_block(1, 1, 1), P.wood, 1), (sh.from_block(0.8, 0.8, 0.8), P.wood, -1)])
    assert abs(hollow.mass - 700 * (1 - 0.8 ** 3)) < 1e-3
    assert abs(hollow.volume - (1 - 0.8 ** 3)) < 1e-5
    expected = 6 / (700 * (1 - 0.8 ** 5))
    assert abs(hollow.inv_inertia[0] - expected) <= 1e-4 * expected


# ---- tests/Shapes/ConvexTest.elm (the hull data the hot path consumes: group / edge order decides the feature keys) --------
def _mesh_block(t, sx, sy, sz):
    """tests/Fixtures/Convex.elm:20-59."""
    from elm_physics_b200 import math3d as m3
    hx, hy, hz = sx / 2, sy / 2, sz / 2
    v = [(hx, -hy, -hz), (hx, hy, -hz), (-hx, hy, -hz), (-hx, -hy, -hz), (hx, -hy, hz), (hx, hy, hz), (-hx, -hy, hz), (-hx, hy, hz)]
    f = [(1, 7, 5), (1, 2, 7), (4, 7, 6), (4, 5, 7), (6, 2, 3), (6, 7, 2), (3, 4, 6), (3, 0, 4), (0, 5, 4), (0, 1, 5), (3, 1, 0), (3, 2, 1)]
    return sh.from_triangular_mesh(f, [m3.point_place_in(t, p) for p in v])


def _square_like_pyramid(eps):
    """tests/Fixtures/Convex.elm:138-175."""
    zo = 0.5 ** (1.0 / 3.0)
    v = [(-1.0, -1.0, -zo), (1.0, -1.0, -zo), (1.0 + eps, 1.0 + eps, -zo), (-1.0, 1.0, -zo), (0.0, 0.0, 1.0 - zo)]
    return sh.from_triangular_mesh([(3, 2, 1), (3, 1, 0), (0, 1, 4), (1, 2, 4), (2, 3, 4), (3, 0, 4)], v)


def _flat_faces(c):
    out = []
    for g in c.faces:
        out.append((g.normal, g.indices))
        if g.two_sided:
            out.append((g.partner_normal, g.partner_indices))
    return out


def _mat_close(a, b, tol=1e-9):
    for k in a:
        assert abs(a[k] - b[k]) <= tol * max(1.0, abs(b[k])), (k, a[k], b[k])


def test_convex_inertia_center_of_mass_volume():
    import math
    from elm_physics_b200 import math3d as m3
    _mat_close(_mesh_block(m3.AT_ORIGIN, 2, 3, 5).inertia, sh.from_block(2, 3, 5).inertia)
    t = m3.rotate_around_own(m3.X_AXIS, math.pi / 5, m3.rotate_around_own(m3.Z_AXIS, math.pi / 5, m3.t_at_point((12.0, 2.0, -3.0))))
    _mat_close(_mesh_block(t, 2, 3, 5).inertia, m3.inertia_rotate_in(t, sh.from_block(2, 3, 5).inertia))
    t = m3.rotate_around_own(m3.X_AXIS, math.pi / 5, m3.rotate_around_own(m3.Z_AXIS, math.pi / 5, m3.t_at_point((2.0, 1.0, -2.0))))
    a, b = _mesh_block(t, 2, 3, 5).position, sh.convex_place_in(t, sh.from_block(2, 3, 5)).position
    assert all(abs(p - q) < 1e-9 for p, q in zip(a, b))
    assert abs(_mesh_block(m3.AT_ORIGIN, 2, 3, 1).volume - sh.from_block(2, 3, 1).volume) < 1e-12


def test_convex_faces_normals_and_winding():
    from elm_physics_b200 import math3d as m3
    import json
    import os
    block = sh.from_block(2, 2, 2)
    assert [n for n, _ in _flat_faces(block)] == [(0.0, 0.0, 1.0), (0.0, 0.0, -1.0), (0.0, 1.0, 0.0), (0.0, -1.0, 0.0),
                                                  (1.0, 0.0, 0.0), (-1.0, 0.0, 0.0)]
    assert [g.normal for g in block.faces] == [(0.0, 0.0, 1.0), (0.0, 1.0, 0.0), (1.0, 0.0, 0.0)]      # .uniqeNormals
    huge = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "huge_convex.json")))
    hulls = [block, _square_like_pyramid(0.0), sh.from_cylinder(6, 4, 5), _mesh_block(m3.AT_ORIGIN, 4, 4, 4),
             sh.from_triangular_mesh([tuple(f) for f in huge["faces"]], [tuple(v) for v in huge["vertices"]])]
    for c in hulls:
        for normal, idx in _flat_faces(c):
            v = [c.vertices[i] for i in idx]
            # the normal points away from the hull's position ...
            assert m3.dot(normal, m3.sub(v[0], c.position)) > 0
            # ... and every fan triangle of the face winds counter-clockwise around it
            for k in range(1, len(v) - 1):
                n = sh.compute_normal(v[0], v[k], v[k + 1])
                assert all(abs(p - q) < 1e-6 for p, q in zip(n, normal)), (normal, n)
    assert len(_square_like_pyramid(0.0).faces) == 5


def test_convex_unique_edges():
    from elm_physics_b200 import math3d as m3
    assert sh.from_block(2, 2, 2).unique_edges == [[0, 4, 3, 7, 2, 6, 1, 5], [0, 3, 5, 6, 4, 7, 1, 2], [0, 1, 7, 6, 4, 5, 3, 2]]
    assert [len(g) // 2 for g in sh.from_block(2, 3, 5).unique_edges] == [4, 4, 4]
    assert sorted(len(g) // 2 for g in _mesh_block(m3.AT_ORIGIN, 2, 3, 5).unique_edges) == [4, 4, 4]
    assert len(_square_like_pyramid(0.0).unique_edges) == 6
    assert len(_square_like_pyramid(1.0e-5).unique_edges) == 6        # askew: below the parallel tolerance
    assert len(_square_like_pyramid(1.0e-2).unique_edges) == 8        # above it: all edges unique, none parallel


# ---- tests/Transform3dTest.elm:9-92 (the placement math of the path, host-side mirror) -----------------------------------
def test_transform3d_round_trips():
    import math
    from elm_physics_b200 import math3d as m3
    t = m3.rotate_around_own(m3.X_AXIS, math.pi / 5, m3.rotate_around_own(m3.Z_AXIS, math.pi / 5, m3.t_at_point((0.5, 0.6, 0.7))))
    close = lambda a, b: all(abs(p - q) < 1e-12 for p, q in zip(a, b))
    point, direction = (0.4, 0.6, 0.8), (0.4, 0.6, 0.8)
    assert close(m3.point_relative_to(t, m3.point_place_in(t, point)), point)
    assert close(m3.direction_relative_to(t, m3.direction_place_in(t, direction)), direction)
    inv = m3.t_inverse(t)          # Transform3d.relativeTo t atOrigin
    assert close(m3.direction_place_in(inv, m3.direction_place_in(t, direction)), direction)
    assert close(m3.point_place_in(inv, m3.point_place_in(t, (0.3, 0.5, 0.7))), (0.3, 0.5, 0.7))


# ---- tests/EigenDecompositionTest.elm (properties): Body.compound's principal frame comes from this ----------------------
def test_eigen_decomposition_properties():
    import numpy as np
    from elm_physics_b200 import math3d as m3
    rng = np.random.RandomState(7)
    for trial in range(40):
        a = rng.uniform(-1, 1, (3, 3))
        sym = a @ a.T + (np.eye(3) * rng.uniform(0, 2) if trial % 3 else 0)
        if trial % 5 == 0:
            sym = np.diag(rng.uniform(0.1, 3, 3))                     # already diagonal
        m = dict(m11=sym[0, 0], m21=sym[1, 0], m31=sym[2, 0], m12=sym[0, 1], m22=sym[1, 1], m32=sym[2, 1],
                 m13=sym[0, 2], m23=sym[1, 2], m33=sym[2, 2])
        e = m3.eigen_decomposition(m)
        vs = [np.array(e[k]) for k in ("v1", "v2", "v3")]
        ev = e["eigenvalues"]
        for lam, v in zip(ev, vs):
            assert abs(np.linalg.norm(v) - 1) < 1e-9
            assert np.linalg.norm(sym @ v - lam * v) < 1e-8 * max(1.0, abs(lam))
        assert abs(np.dot(vs[0], vs[1])) < 1e-9 and abs(np.dot(vs[0], vs[2])) < 1e-9 and abs(np.dot(vs[1], vs[2])) < 1e-9
        assert np.dot(np.cross(vs[0], vs[1]), vs[2]) > 0.999          # right-handed: usable as a body frame
        assert np.allclose(sorted(ev), np.linalg.eigvalsh(sym), atol=1e-9)


def test_contact_points_follows_contactPointsHelp():
    """Physics.elm:1158-1207: the predicate is tried in both argument orders, pair groups without contacts (constraint-only
    pairs) are skipped, the result is in pair-enumeration order (the reference conses while walking the reversed list),
    and the points go through the COMPOSED transform `atOrigin |> relativeTo pre |> placeIn post`."""
    import numpy as np
    a = with_id(0, plane()); b = with_id(1, plane()); c = with_id(2, plane())
    qz = (0.0, 0.0, 0.14943813247359922, 0.9887710779360422)         # 0.3 rad about z
    pre_a = ((0.0, 0.0, 1.0), qz)
    moved_a = replace(a, transform3d=((0.5, -1.0, 2.0), (0.0, 0.0, 0.0, 1.0)))
    packed = dict(
        group_id1=np.array([0, 0, 1]), group_id2=np.array([1, 2, 2]), group_n_constraints=np.array([0, 1, 0]),
        group_contact_off=np.array([0, 1, 1, 3]),
        group_pre_pos1=np.array([pre_a[0], pre_a[0], (0.0, 0.0, 0.0)]), group_pre_quat1=np.array([qz, qz, (0.0, 0.0, 0.0, 1.0)]),
        shape_key=np.array([257, 257, 257]), feature_key=np.array([1, 2, 3]),
        ni=np.zeros((3, 3)), pi=np.array([(1.0, 0.0, 1.0), (0.0, 2.0, 0.0), (0.0, 3.0, 0.0)]), pj=np.zeros((3, 3)),
        friction=np.zeros(3), bounciness=np.zeros(3), lambda_=np.zeros(3), t1=np.zeros((3, 3)), iterations=1)
    contacts = P.Contacts(packed=packed, iterations=1, bodies=[("jeep", moved_a), ("wall", b), ("crate", c)])
    everything = P.contact_points(lambda x, y: True, contacts)
    assert [(e1, e2, len(p)) for e1, e2, p in everything] == [("jeep", "wall", 1), ("wall", "crate", 2)]     # (0,2) has no contacts
    # pairGroup.contacts is the reverse of the solver order the arrays hold
    assert everything[1][2] == [(0.0, 3.0, 0.0), (0.0, 2.0, 0.0)]
    # asymmetric predicates match whichever body is body1 (Physics.elm:1177)
    assert [(e1, e2) for e1, e2, _ in P.contact_points(lambda x, y: x == "wall" and y == "jeep", contacts)] == [("jeep", "wall")]
    assert [(e1, e2) for e1, e2, _ in P.contact_points(lambda x, y: x == "crate" and y == "wall", contacts)] == [("wall", "crate")]
    assert P.contact_points(lambda x, y: x == "jeep" and y == "crate", contacts) == []
    # the point: relative to the pre-step frame of body1, placed in its post-step frame, as one composed transform
    t = m3.t_place_in(moved_a.transform3d, m3.t_relative_to(pre_a, m3.AT_ORIGIN))
    assert everything[0][2][0] == m3.point_place_in(t, (1.0, 0.0, 1.0))
    x, y, z = everything[0][2][0]
    # (1,0,1) is (1,0,0) from the old origin; derotated by 0.3 rad about z, then put at the new origin
    import math
    assert abs(x - (0.5 + math.cos(0.3))) < 1e-12 and abs(y - (-1.0 - math.sin(0.3))) < 1e-12 and abs(z - 2.0) < 1e-12


def test_level_schedule_parallel_formulation_equals_the_walk():
    """The level schedule of a giant island (csrc/ep_solve_cluster.cuh): level(g) = max over the two dynamic bodies of g of
    (1 + level of the previous group touching that body).  The kernel builds it in parallel -- unordered per-body group lists,
    predecessor = the largest earlier group of the list, relaxation rounds in any order until nothing changes -- and that must be
    the one-thread walk's result, in at most (levels + 1) rounds.  A model of both, on random islands."""
    import random

    def walk(groups, nb):
        last, lv = [0] * (nb + 1), []
        for a, b in groups:
            l = max(last[a] if a else 0, last[b] if b else 0)
            lv.append(l)
            if a:
                last[a] = l + 1
            if b:
                last[b] = l + 1
        return lv

    def parallel(groups, nb, rng):
        gn = len(groups)
        lists = [[] for _ in range(nb + 1)]
        for g in rng.sample(range(gn), gn):                 # the fill order of the atomics is arbitrary
            for x in groups[g]:
                if x:
                    lists[x].append(g)
        pred = [[max([o for o in lists[x] if o < g], default=-1) if x else -1 for x in groups[g]] for g in range(gn)]
        lv, rounds = [0] * gn, 0
        while True:
            rounds += 1
            changed = False
            for g in rng.sample(range(gn), gn):             # threads race: any order, stale reads allowed
                l = max((lv[p] + 1 if p >= 0 else 0) for p in pred[g])
                if l != lv[g]:
                    lv[g], changed = l, True
            if not changed:
                return lv, rounds

    rng = random.Random(7)
    for _ in range(200):
        nb, gn = rng.randint(1, 30), rng.randint(1, 150)
        groups = []
        while len(groups) < gn:
            a, b = rng.randint(0, nb), rng.randint(0, nb)       # 0 = a non-dynamic body (the stand-in)
            if a != b:
                groups.append((a, b))
        expect = walk(groups, nb)
        got, rounds = parallel(groups, nb, rng)
        assert got == expect
        assert rounds <= max(expect) + 2


def test_from_origin_and_basis_round_trips_a_rotated_frame():
    """tests/Transform3dFromFrame3dTest.elm:43-71: a right-handed frame at (0.5, 0.6, 0.7) rotated by pi/5 around the world x axis and
    then the world y axis -> Transform3d.fromOriginAndBasis -> the same basis back out of the quaternion, and the same placed point
    as origin + p.x X + p.y Y + p.z Z.  All four branches of the matrix-to-quaternion conversion (Transform3d.elm:34-121) as well."""
    import math

    def rot(axis, ang, v):
        c, s = math.cos(ang), math.sin(ang)
        x, y, z = v
        return (x, c * y - s * z, s * y + c * z) if axis == "x" else (c * x + s * z, y, -s * x + c * z)

    def close(a, b, tol=1e-12):
        return all(abs(p - q) < tol for p, q in zip(a, b))

    def check(origin, X, Y, Z):
        t = m3.from_origin_and_basis(origin, X, Y, Z)
        o = m3.orientation(t)
        assert close((o["m11"], o["m21"], o["m31"]), X) and close((o["m12"], o["m22"], o["m32"]), Y) and close((o["m13"], o["m23"], o["m33"]), Z)
        p = (0.234, 0.234, 0.234)
        expect = tuple(origin[k] + p[0] * X[k] + p[1] * Y[k] + p[2] * Z[k] for k in range(3))
        assert close(m3.point_place_in(t, p), expect)

    frame = [(0.5, 0.6, 0.7), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0)]
    frame = [rot("y", math.pi / 5, rot("x", math.pi / 5, v)) for v in frame]
    check(*frame)
    # the other three branches: half turns about x, y, z (trace = -1, largest diagonal element m00 / m11 / m22) slightly perturbed
    for axis, ang in (("x", math.pi - 0.01), ("y", math.pi - 0.01)):
        basis = [rot(axis, ang, v) for v in ((1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0))]
        check((0.1, -0.2, 0.3), *basis)
    c, s = math.cos(math.pi - 0.01), math.sin(math.pi - 0.01)
    check((0.0, 0.0, 0.0), (c, s, 0.0), (-s, c, 0.0), (0.0, 0.0, 1.0))
