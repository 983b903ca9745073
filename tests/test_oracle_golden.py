"""Pins the CPU oracle (oracle/) against the reference's own known-answer tests.
Each test names the reference test it ports.  Tolerance is the reference's (Extra/Expect.elm:95-97,
absolute 1e-4) unless the reference gives full-precision digits, in which case it is 1e-12.
"""
import json
import math
import os

import pytest

from elm_physics_b200 import contact_id as cid
from elm_physics_b200 import math3d as m3
from elm_physics_b200 import pack
from elm_physics_b200 import shapes as sh

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-4


def at_point(p):
    return m3.t_at_point(tuple(float(c) for c in p))


def rot(t, axis, angle):
    return m3.rotate_around_own(axis, angle, t)


def translate(t, v):
    return (m3.add(v, t[0]), t[1])


class Scene:
    """Shape table + oracle calls for world shapes `placeIn transform shape`."""

    def __init__(self, oracle):
        self.o = oracle
        self.table = pack.ShapeTable()
        self.built = None

    def add(self, shape):
        self.built = None
        return self.table.add(shape)

    def shapes(self):
        if self.built is None:
            self.built = self.table.build()
        return self.built

    def contacts(self, s1, t1, s2, t2):
        return self.o.shape_contacts(self.shapes(), s1, t1, s2, t2)


def assert_vec(a, b, tol=TOL):
    assert all(abs(x - y) <= tol for x, y in zip(a, b)), (a, b)


def assert_contacts(actual, expected, tol=TOL, ids=True):
    assert len(actual) == len(expected), (actual, expected)
    for a, e in zip(actual, expected):
        if ids:
            assert cid.feature_string(a["feature_key"]) == cid.feature_string(e["feature_key"]), (a, e)
        for k in ("pi", "pj", "ni"):
            assert_vec(a[k], e[k], tol)


# ---- tests/Collision/PlaneSphereTest.elm:36-61 ----------------------------------------------------
def test_plane_sphere(oracle):
    sc = Scene(oracle)
    plane = sc.add(sh.Plane((0.0, 0.0, 1.0), (0.0, 0.0, 0.0)))
    sphere = sc.add(sh.sphere_at_origin(1))
    I = m3.AT_ORIGIN
    assert_contacts(sc.contacts(plane, I, sphere, at_point((0, 0, 1))),
                    [dict(feature_key=0, ni=(0, 0, 1), pi=(0, 0, 0), pj=(0, 0, 0))])
    assert_contacts(sc.contacts(plane, I, sphere, at_point((0, 0, 0.7))),
                    [dict(feature_key=0, ni=(0, 0, 1), pi=(0, 0, 0), pj=(0, 0, -0.3))])
    assert sc.contacts(plane, I, sphere, at_point((0, 0, 1.3))) == []


# ---- tests/Collision/PlaneConvexTest.elm:13-29 ----------------------------------------------------
def test_plane_convex_cylinder_cap_culled_to_4(oracle):
    sc = Scene(oracle)
    plane = sc.add(sh.Plane((0.0, 0.0, 1.0), (0.0, 0.0, 0.0)))
    cyl = sc.add(sh.from_cylinder(12, 0.5, 1))
    assert len(sc.contacts(plane, m3.AT_ORIGIN, cyl, at_point((0, 0, 0.4995)))) == 4


# ---- tests/Collision/ConvexConvexTest.elm:18-160 --------------------------------------------------
def tilted_box_transform(origin):
    return m3.from_origin_and_basis(
        origin,
        (-0.001580161067349219, 0.9999882020975505, -0.004593338296621986),
        (-0.6852663896092609, -0.004428115539111141, -0.7282790447793068),
        (-0.7282907924468673, 0.0019968621580540736, 0.6852653020390244))


def test_convex_convex_counts(oracle):
    sc = Scene(oracle)
    b2 = sc.add(sh.from_block(2, 2, 2))
    b12 = sc.add(sh.from_block(1.2, 1.2, 1.2))
    b1 = sc.add(sh.from_block(1, 1, 1))
    # "should return 4 results" (39-55)
    t1 = rot(at_point((0, 0, 2.1)), m3.Y_AXIS, math.pi / 2)
    t2 = rot(at_point((0, 0, 4)), m3.Y_AXIS, math.pi / 2)
    assert len(sc.contacts(b2, t1, b2, t2)) == 4
    # "should return 2 results" (56-79)
    t1 = rot(at_point((-0.5, 0, 0)), m3.Z_AXIS, math.pi / 2)
    t2 = rot(at_point((0.5, 0, 0)), m3.Z_AXIS, math.pi / 4)
    assert len(sc.contacts(b12, t1, b1, t2)) == 2
    # single edge-edge contact (80-100)
    t1 = rot(m3.AT_ORIGIN, m3.Y_AXIS, math.pi / 6)
    t2 = rot(at_point((0.6, 0.6, 0.6)), m3.X_AXIS, math.pi / 6)
    assert len(sc.contacts(b1, t1, b1, t2)) == 1
    # corner-on-corner stays on the face-clip path (101-115)
    cs = sc.contacts(b1, m3.AT_ORIGIN, b1, at_point((0.7, 0.7, 0.7)))
    assert all(not cid.feature_string(c["feature_key"]).startswith("-e") for c in cs)


def test_convex_convex_near_coplanar_stack_interface(oracle):
    """ConvexConvexTest.elm:116-141 — 16-digit settling poses must give exactly 4 contacts."""
    sc = Scene(oracle)
    b1 = sc.add(sh.from_block(1, 1, 1))
    t2 = m3.from_origin_and_basis(
        (0.0002598859790008731, 0.00013646414197537694, 1.4996999282620576),
        (0.999999999331571, 0.0000032595970446131582, -0.00003641748174313937),
        (-0.0000032582908802255252, 0.999999999351492, 0.00003586641330410744),
        (0.00003641759862957714, -0.000035866294621384654, 0.9999999986936836))
    t3 = m3.from_origin_and_basis(
        (0.0001967863351834402, 0.000008252933652183521, 2.499616662498541),
        (0.9999999861179005, 0.000002032073717667724, 0.00016661353315350634),
        (-0.0000020688652840958605, 0.9999999756171408, 0.00022081992116338665),
        (-0.00016661308036863387, -0.00022082026279889716, 0.9999999617392458))
    assert len(sc.contacts(b1, t2, b1, t3)) == 4


@pytest.mark.parametrize("origin", [(-0.11608751888227789, 0.5683352706207844, 1.9053833288923054),
                                    (-0.06790182997043828, -0.6276944695546394, 1.9053833288923054)])
def test_convex_convex_tilted_box_ids(oracle, origin):
    """ConvexConvexTest.elm:142-159."""
    sc = Scene(oracle)
    tilted = sc.add(sh.from_block(1.5, 1.5, 1.5))
    target = sc.add(sh.from_block(2, 2, 2))
    cs = sc.contacts(tilted, tilted_box_transform(origin), target, m3.AT_ORIGIN)
    assert sorted(cid.feature_string(c["feature_key"]) for c in cs) == ["-f3-f5-v5", "-f3-f5-v6"]


def test_test_separating_axis(oracle):
    """ConvexConvexTest.elm:163-215."""
    sc = Scene(oracle)
    b = sc.add(sh.from_block(1, 1, 1))
    S = sc.shapes()
    assert oracle.test_separating_axis(S, b, at_point((-0.2, 0, 0)), b, at_point((0.2, 0, 0)), m3.X_AXIS) == 0.6
    assert oracle.test_separating_axis(S, b, at_point((-5.2, 0, 0)), b, at_point((0.2, 0, 0)), m3.X_AXIS) is None
    v = oracle.test_separating_axis(S, b, at_point((1, 0, 0)), b, rot(at_point((0.2, 0, 0)), m3.Z_AXIS, math.pi / 4), m3.X_AXIS)
    assert abs(v - 0.4071067) < 1e-5


def test_find_separating_axis(oracle):
    """ConvexConvexTest.elm:219-253."""
    sc = Scene(oracle)
    b = sc.add(sh.from_block(1, 1, 1))
    S = sc.shapes()
    assert oracle.find_separating_axis(S, b, at_point((-0.2, 0, 0)), b, at_point((0.2, 0, 0))) == (-1.0, 0.0, 0.0)
    assert oracle.find_separating_axis(S, b, at_point((-0.2, 0, 0)), b,
                                       rot(at_point((0.2, 0, 0)), m3.Z_AXIS, math.pi / 4)) == (-1.0, 0.0, 0.0)


# ---- tests/Collision/ConvexConvexTest.elm:256-319 ------------------------------------------------------
def test_project(oracle):
    sc = Scene(oracle)
    b = sc.add(sh.from_block(1, 1, 1))
    S = sc.shapes()
    assert oracle.project(S, b, m3.AT_ORIGIN, (1.0, 0.0, 0.0)) == (-0.5, 0.5)
    assert oracle.project(S, b, m3.AT_ORIGIN, (-1.0, 0.0, 0.0)) == (-0.5, 0.5)
    assert oracle.project(S, b, m3.AT_ORIGIN, (0.0, 1.0, 0.0)) == (-0.5, 0.5)
    assert oracle.project(S, b, at_point((0, 1, 0)), (0.0, 1.0, 0.0)) == (0.5, 1.5)
    mn, mx = oracle.project(S, b, rot(at_point((0, 1, 0)), m3.X_AXIS, math.pi / 2), (0.0, 1.0, 0.0))
    assert abs(mn - 0.5) < 1e-5 and abs(mx - 1.5) < 1e-5


# ---- tests/ConvexSphereBoxTest.elm:21-182 (full-precision expectation) -------------------------------
def test_icosphere_box_edge_contact_full_precision(oracle):
    mesh = json.load(open(os.path.join(GOLDEN, "icosphere.json")))
    sc = Scene(oracle)
    ico = sc.add(sh.from_triangular_mesh(mesh["faces"], mesh["vertices"]))
    box = sc.add(sh.from_block(2, 2, 2))
    t = m3.from_origin_and_basis((0.08070649126382506, -1.5245882476123196, 1.7289522010617453),
                                 (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0))
    cs = sc.contacts(ico, t, box, m3.AT_ORIGIN)
    expected = [dict(feature_key=cid.convex_convex_edge(20, 3, 3, 3),
                     ni=(0, 0.7246131188598045, -0.6891558807528719),
                     pi=(0.01797561156268518, -0.9521386364340609, 0.9544806196522924),
                     pj=(0.017975611562685012, -1, 1))]
    assert len(cs) == 1 and cs[0]["feature_key"] == expected[0]["feature_key"]
    assert_contacts(cs, expected, tol=1e-12)


# ---- tests/Collision/SphereConvexTest.elm:34-111 + tests/Fixtures/NarrowPhase.elm -------------------
def octo_hull(he):
    """tests/Fixtures/Convex.elm:98-118."""
    return sh.from_triangular_mesh(
        [(2, 1, 0), (0, 5, 2), (1, 2, 4), (3, 0, 1), (2, 5, 4), (4, 3, 1), (5, 0, 3), (3, 4, 5)],
        [(0, 0, he), (0, he, 0), (he, 0, 0), (-he, 0, 0), (0, 0, -he), (0, -he, 0)])


def test_sphere_convex_fixtures(oracle):
    fx = json.load(open(os.path.join(GOLDEN, "sphere_convex.json")))
    sc = Scene(oracle)
    sphere = sc.add(sh.sphere_at_origin(fx["radius"]))
    box = sc.add(sh.from_block(fx["box_size"], fx["box_size"], fx["box_size"]))
    # `placeInWithCorrectWinding` places twice (atPoint, then atOrigin): face order back to source order.
    octo = sc.add(sh.convex_place_in(m3.AT_ORIGIN, octo_hull(fx["octo_half_extent"])))
    center = at_point(fx["center"])
    for e in fx["box"]:
        assert_contacts(sc.contacts(sphere, center, box, at_point(e["position"])),
                        [dict(ni=e["ni"], pi=e["pi"], pj=e["pj"])], ids=False)
    for e in fx["octo"]:
        assert_contacts(sc.contacts(sphere, center, octo, at_point(e["position"])),
                        [dict(ni=e["ni"], pi=e["pi"], pj=e["pj"])], ids=False)
    for e in fx["box_far"]:
        assert sc.contacts(sphere, center, box, at_point(e["position"])) == []
    for e in fx["octo_far"]:
        assert sc.contacts(sphere, center, octo, at_point(e["position"])) == []


# ---- tests/Collision/CapsuleConvexTest.elm:135-331 ---------------------------------------------------
def test_capsule_convex_table(oracle):
    sc = Scene(oracle)
    box = sc.add(sh.from_block(2, 2, 2))
    c05 = sc.add(sh.capsule_at_origin(0.5, 0.5))
    c10 = sc.add(sh.capsule_at_origin(0.5, 1.0))
    I = m3.AT_ORIGIN
    pi_ = math.pi
    inv2, inv3 = 1 / math.sqrt(2), 1 / math.sqrt(3)
    F = cid.on_convex_face

    def run(cap, t):
        return sc.contacts(cap, t, box, at_point((0, 0, 0)))

    # Cap on face F
    assert_contacts(run(c05, at_point((0, 0, 1.9))),
                    [dict(feature_key=cid.capsule_cap_end1(F(5)), ni=(0, 0, -1), pi=(0, 0, 0.9), pj=(0, 0, 1.0))])
    # Cylinder on face F (segment clip)
    assert_contacts(run(c10, translate(rot(I, m3.X_AXIS, pi_ / 2), (0, 0, 1.4))),
                    [dict(feature_key=cid.capsule_cylinder2(F(5)), ni=(0, 0, -1), pi=(0, -1, 0.9), pj=(0, -1, 1.0)),
                     dict(feature_key=cid.capsule_cylinder1(F(5)), ni=(0, 0, -1), pi=(0, 1, 0.9), pj=(0, 1, 1.0))])
    # too far
    assert run(c05, at_point((0, 0, 2.6))) == []
    # tilted capsule's lower cap
    half = math.sqrt(2) / 2
    assert_contacts(run(c10, translate(rot(I, m3.Y_AXIS, pi_ / 4), (0, 0, 2))),
                    [dict(feature_key=cid.capsule_cap_end1(F(5)), ni=(0, 0, -1), pi=(-half, 0, 2 - half - 0.5), pj=(-half, 0, 1))])
    # vertical capsule overhanging +x edge
    assert_contacts(run(c05, at_point((1.4, 0, 1.4))),
                    [dict(feature_key=cid.capsule_cylinder2(F(1)), ni=(-1, 0, 0), pi=(0.9, 0, 1), pj=(1, 0, 1)),
                     dict(feature_key=cid.capsule_cylinder1(F(1)), ni=(-1, 0, 0), pi=(0.9, 0, 0.9), pj=(1, 0, 0.9))])
    # just outside a corner
    assert run(c05, translate(rot(I, m3.Y_AXIS, pi_ / 4), (1.8, 0, 1.8))) == []
    # parallel to +y +z edge
    inv_len = 1 / math.sqrt(0.18)
    ni = (0, -0.3 * inv_len, -0.3 * inv_len)

    def at(id_, x):
        return dict(feature_key=id_, ni=ni, pi=(x, 1.3 + 0.5 * ni[1], 1.3 + 0.5 * ni[2]), pj=(x, 1, 1))

    assert_contacts(run(c10, translate(rot(I, m3.Y_AXIS, pi_ / 2), (0, 1.3, 1.3))),
                    [at(cid.capsule_cylinder2(cid.on_convex_edge), 1), at(cid.capsule_cylinder1(cid.on_convex_edge), -1)])
    # cap on convex edge
    pic = 1.8 - 1.5 * inv2
    assert_contacts(run(c10, translate(rot(I, m3.Y_AXIS, pi_ / 4), (1.8, 0, 1.8))),
                    [dict(feature_key=cid.capsule_cap_end1(cid.on_convex_edge), ni=(-inv2, 0, -inv2), pi=(pic, 0, pic), pj=(1, 0, 1))])
    # cap on convex vertex
    pic = 1.4 - inv3
    t = translate(rot(rot(I, m3.Y_AXIS, math.acos(1 / math.sqrt(3))), m3.Z_AXIS, pi_ / 4), (1.4, 1.4, 1.4))
    assert_contacts(run(c05, t),
                    [dict(feature_key=cid.capsule_cap_end1(cid.on_convex_vertex), ni=(-inv3, -inv3, -inv3), pi=(pic, pic, pic), pj=(1, 1, 1))])
    # closest-edge fallback
    assert_contacts(run(c05, translate(rot(I, m3.Y_AXIS, pi_ / 2), (1.5, 0.95, 1.4))),
                    [dict(feature_key=cid.capsule_cylinder(F(5)), ni=(0, 0, -1), pi=(1, 0.95, 0.9), pj=(1, 0.95, 1.0))])
    # skew edge
    pic = 1.3 - 0.5 * inv2
    assert_contacts(run(c10, translate(rot(I, m3.Y_AXIS, 3 * pi_ / 4), (1.3, 0, 1.3))),
                    [dict(feature_key=cid.capsule_cylinder(cid.on_convex_edge), ni=(-inv2, 0, -inv2), pi=(pic, 0, pic), pj=(1, 0, 1))])
    # cylinder on vertex
    pic = 1.2 - 0.5 * inv3
    t = translate(rot(rot(I, m3.Y_AXIS, pi_ / 2), m3.Z_AXIS, -pi_ / 4), (1.2, 1.2, 1.2))
    assert_contacts(run(c10, t),
                    [dict(feature_key=cid.capsule_cylinder(cid.on_convex_vertex), ni=(-inv3, -inv3, -inv3), pi=(pic, pic, pic), pj=(1, 1, 1))])


def test_capsule_convex_flip(oracle):
    """CapsuleConvexTest.elm:165-175 — convex is body1: NarrowPhase passes Contact.flip."""
    sc = Scene(oracle)
    box = sc.add(sh.from_block(2, 2, 2))
    c05 = sc.add(sh.capsule_at_origin(0.5, 0.5))
    cs = sc.contacts(box, m3.AT_ORIGIN, c05, at_point((0, 0, 1.9)))
    assert_contacts(cs, [dict(feature_key=cid.capsule_cap_end1(cid.on_convex_face(5)), ni=(0, 0, 1), pi=(0, 0, 1.0), pj=(0, 0, 0.9))])


def test_support_feature_returns_edge_adjacent_pair(oracle):
    """CapsuleConvexTest.elm:340-401."""
    d1, d2, a1, a2, apex = (1.0, 1.0, 0.0), (-1.0, -1.0, 0.0), (-1.0, 1.0, 0.0), (1.0, -1.0, 0.0), (0.0, 0.0, 1.0)
    n = m3.normalize
    faces = [(([d1, a2, d2, a1], (0.0, 0.0, -1.0)), None), (([d1, apex, a2], n((1, 0, 1))), None),
             (([a2, apex, d2], n((0, -1, 1))), None), (([d2, apex, a1], n((-1, 0, 1))), None),
             (([a1, apex, d1], n((0, 1, 1))), None)]
    pyramid = sh.convex_init(faces, [d1, d2, a1, a2, apex], [[d1, a2, d2, a1], [a2, d2, a1, d1]], m3.ZERO, m3.m_zero(), 0)
    sc = Scene(oracle)
    p = sc.add(pyramid)
    assert oracle.support_feature(sc.shapes(), p, m3.AT_ORIGIN, (0, 0, -1)) == [d1, a1]


# ---- tests/SolverIslandsTest.elm:136-163 -------------------------------------------------------------
def test_islands_visit_order(oracle):
    import ctypes as C
    import numpy as np
    groups = [((0, 1), (1, 2)), ((1, 2), (2, 2)), ((1, 2), (3, 2)), ((2, 2), (3, 2)), ((5, 2), (6, 2))]
    i1 = np.array([g[0][0] for g in groups], np.int32); k1 = np.array([g[0][1] for g in groups], np.int32)
    i2 = np.array([g[1][0] for g in groups], np.int32); k2 = np.array([g[1][1] for g in groups], np.int32)
    root = np.zeros(5, np.int32); order = np.zeros(5, np.int32)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    oracle.lib().ep_oracle_islands(5, p(i1), p(k1), p(i2), p(k2), 6, p(root), p(order))
    visits, cur = [], None
    for g in order:
        if root[g] != cur:
            visits.append([])
            cur = root[g]
        visits[-1].append((int(i1[g]), int(i2[g])))
    assert visits == [[(0, 1), (1, 2), (1, 3), (2, 3)], [(5, 6)]]


def test_islands_fuzz_properties(oracle):
    """SolverIslandsTest.elm:164-236 — bottom-first inside an island, disjoint islands, no lost groups."""
    import ctypes as C
    import random
    import numpy as np
    rng = random.Random(7)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    for _ in range(200):
        n = rng.randint(4, 8)
        split = n // 2
        pairs = [(i, j) for i in range(0, split - 1) for j in range(i + 1, split)] + \
                [(i, j) for i in range(split, n - 1) for j in range(i + 1, n)]
        i1 = np.array([a for a, _ in pairs], np.int32); i2 = np.array([b for _, b in pairs], np.int32)
        k = np.full(len(pairs), 2, np.int32)
        root = np.zeros(len(pairs), np.int32); order = np.zeros(len(pairs), np.int32)
        oracle.lib().ep_oracle_islands(len(pairs), p(i1), p(k), p(i2), p(k), n - 1, p(root), p(order))
        islands = {}
        for g in order:
            islands.setdefault(int(root[g]), []).append((int(i1[g]), int(i2[g])))
        assert 1 < len(islands) < n
        assert sum(len(v) for v in islands.values()) == len(pairs)
        sets = [set(x for pr in v for x in pr) for v in islands.values()]
        assert sum(len(s) for s in sets) == len(set().union(*sets))
        for v in islands.values():
            assert [a for a, _ in v] == sorted(a for a, _ in v)     # body1 id == gravity rank: non-decreasing
