"""Packed integer contact ids (reference: src/Internal/ContactId.elm).  Host-side helpers only:
the device builds the same integers itself (csrc/), these encode/decode them for tests and debug
overlays (`featureString`, ContactId.elm:316-389)."""

BODY_RANGE, SHAPE_RANGE, FEAT_RANGE = 262144, 256, 2048


def body_key(a, b):
    return a * BODY_RANGE + b if a - b <= 0 else b * BODY_RANGE + a


def shape_key(s1, s2):
    return s1 * SHAPE_RANGE + s2


def feature(tag, a, b, c, d):
    return (((tag * FEAT_RANGE + a) * FEAT_RANGE + b) * FEAT_RANGE + c) * FEAT_RANGE + d


simple = 0
on_convex_none, on_convex_edge, on_convex_vertex = 0, 1, 2


def on_convex_face(face_id):
    return face_id + 2


def convex_convex_face(f1, f2, v):
    return feature(1, f1, f2, v, 0)


def convex_convex_edge(d1, e1, d2, e2):
    return feature(2, d1, e1, d2, e2)


def plane_vertex(v):
    return feature(3, v, 0, 0, 0)


def plane_cap_end(e):
    return feature(4, e, 0, 0, 0)


def sphere_on_convex(cf):
    return feature(5, cf, 0, 0, 0)


def capsule_cap_end1(cf):
    return feature(6, 1, cf, 0, 0)


def capsule_cap_end2(cf):
    return feature(6, 2, cf, 0, 0)


def capsule_cylinder1(cf):
    return feature(6, 3, cf, 0, 0)


def capsule_cylinder2(cf):
    return feature(6, 4, cf, 0, 0)


def capsule_cylinder(cf):
    return feature(6, 5, cf, 0, 0)


def _convex_feature_string(cf):
    return {0: "", 1: "-e", 2: "-v"}.get(cf, "-f%d" % (cf - 2))


def feature_string(fk):
    """ContactId.elm:316-356."""
    d = fk % FEAT_RANGE
    c = (fk // FEAT_RANGE) % FEAT_RANGE
    b = (fk // FEAT_RANGE ** 2) % FEAT_RANGE
    a = (fk // FEAT_RANGE ** 3) % FEAT_RANGE
    tag = fk // FEAT_RANGE ** 4
    if tag == 1:
        return "-f%d-f%d-v%d" % (a, b, c)
    if tag == 2:
        return "-e%d.%d-e%d.%d" % (a, b, c, d)
    if tag == 3:
        return "-v%d" % a
    if tag == 4:
        return "-e%d" % a
    if tag == 5:
        return _convex_feature_string(a)
    if tag == 6:
        site = {1: "-e1", 2: "-e2", 3: "-c1", 4: "-c2"}.get(a, "-c")
        return site + _convex_feature_string(b)
    return ""


def to_string(shapes, fk):
    return "%d-%d%s" % (shapes // SHAPE_RANGE, shapes % SHAPE_RANGE, feature_string(fk))
