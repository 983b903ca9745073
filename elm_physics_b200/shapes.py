"""Construction-time shape builders for the HOST side (reference: src/Shapes/*.elm,
src/Internal/Shape.elm).  These produce the centre-of-mass-frame shape templates that
are serialised into the packed shape table (`pack.py`) — the per-step placement of the
templates is device work (csrc/), not done here.

A Convex keeps the reference's orders exactly, because contact ids encode them
(Convex.elm:76-99): `faces` is the CoM-frame face-group list (the step walks it
REVERSED, Convex.elm:265-287), `unique_edges` is never reversed (Convex.elm:173),
`vertices` is the vertex buffer in key order (VertexBuffer.elm:39-46).
"""
import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

from . import math3d as m3

ORTHOGONAL_TOLERANCE = 1.0e-6       # Convex.elm:36-38

KIND_CONVEX, KIND_PLANE, KIND_SPHERE, KIND_CAPSULE, KIND_PARTICLE = 0, 1, 2, 3, 4


@dataclass
class FaceGroup:
    """Convex.elm:62-64 : OneSidedFace (partner_* None) | TwoSidedFace."""
    normal: tuple
    indices: List[int]
    dist: float
    partner_normal: Optional[tuple] = None
    partner_indices: Optional[List[int]] = None
    partner_dist: float = 0.0

    @property
    def two_sided(self):
        return self.partner_normal is not None


@dataclass
class Convex:
    faces: List[FaceGroup]
    unique_edges: List[List[int]]
    vertices: List[tuple]
    box: Optional[tuple]             # (ax, ay, az, he) or None for NotBox (Convex.elm:110-112)
    position: tuple
    inertia: dict
    volume: float
    kind: int = KIND_CONVEX


@dataclass
class Sphere:
    radius: float
    position: tuple
    volume: float
    inertia: dict
    kind: int = KIND_SPHERE


@dataclass
class Capsule:
    radius: float
    half_length: float
    axis: tuple
    position: tuple
    volume: float
    inertia: dict
    kind: int = KIND_CAPSULE


@dataclass
class Plane:
    normal: tuple
    position: tuple
    kind: int = KIND_PLANE


@dataclass
class Particle:
    position: tuple
    kind: int = KIND_PARTICLE


# ---- per-kind helpers (Internal/Shape.elm:35-129) -------------------------------
def shape_volume(s):
    return s.volume if s.kind in (KIND_CONVEX, KIND_SPHERE, KIND_CAPSULE) else 0.0


def shape_inertia(s):
    return s.inertia if s.kind in (KIND_CONVEX, KIND_SPHERE, KIND_CAPSULE) else m3.m_zero()


def shape_center_of_mass(s):
    return s.position if s.kind in (KIND_CONVEX, KIND_SPHERE, KIND_CAPSULE) else m3.ZERO


def _max(x, y):          # elm/core Basics.max
    return x if x > y else y


def expand_bounding_sphere_radius(s, r):
    """Internal/Shape.elm:113-129."""
    if s.kind == KIND_CONVEX:
        acc = r * r
        for v in s.vertices:     # ascending key order (VertexBuffer.foldl), Convex.elm:1107-1114
            acc = _max(m3.length_squared(v), acc)
        return math.sqrt(acc)
    if s.kind == KIND_SPHERE:
        return _max(m3.length(s.position) + s.radius, r)
    if s.kind == KIND_CAPSULE:
        return _max(m3.length(s.position) + s.half_length + s.radius, r)
    if s.kind == KIND_PLANE:
        return m3.MAX_NUMBER
    return _max(r, m3.length(s.position))


def place_in(t, s):
    """Internal/Shape.elm:94-110 (construction-time use: body -> CoM frame)."""
    if s.kind == KIND_CONVEX:
        return convex_place_in(t, s)
    if s.kind == KIND_PLANE:
        return Plane(m3.direction_place_in(t, s.normal), m3.point_place_in(t, s.position))
    if s.kind == KIND_SPHERE:
        return Sphere(s.radius, m3.point_place_in(t, s.position), s.volume, s.inertia)
    if s.kind == KIND_CAPSULE:
        return Capsule(s.radius, s.half_length, m3.direction_place_in(t, s.axis),
                       m3.point_place_in(t, s.position), s.volume, m3.inertia_rotate_in(t, s.inertia))
    return Particle(m3.point_place_in(t, s.position))


def convex_place_in(t, c):
    """Convex.elm:166-179; face groups come out REVERSED (placeFaces 265-287)."""
    faces = []
    for g in c.faces:
        faces.insert(0, FaceGroup(
            m3.direction_place_in(t, g.normal), g.indices, g.dist,
            m3.direction_place_in(t, g.partner_normal) if g.two_sided else None,
            g.partner_indices, g.partner_dist))
    box = None
    if c.box is not None:
        ax, ay, az, he = c.box
        box = (m3.direction_place_in(t, ax), m3.direction_place_in(t, ay), m3.direction_place_in(t, az), he)
    return Convex(faces, c.unique_edges, [m3.point_place_in(t, v) for v in c.vertices], box,
                  m3.point_place_in(t, c.position), m3.inertia_rotate_in(t, c.inertia), c.volume)


# ---- Sphere / Capsule constructors -------------------------------------------
def sphere_at_origin(radius):
    """Sphere.elm:23-32."""
    volume = 4 / 3 * math.pi * (radius ** 3)
    return Sphere(radius, m3.ZERO, volume, m3.sphere_inertia(volume, radius))


def capsule_at_origin(radius, half_length):
    """Capsule.elm:33-46."""
    volume = math.pi * radius * radius * (2 * half_length + 4 / 3 * radius)
    return Capsule(radius, half_length, m3.Z_AXIS, m3.ZERO, volume,
                   m3.capsule_inertia(volume, radius, half_length))


# ---- Convex.init and friends (Convex.elm:297-409) --------------------------------
def _index_of(target, vertices):
    for i, v in enumerate(vertices):
        if v == target:
            return i
    return -1


def _orthogonal(a, b):
    return abs(m3.dot(a, b)) - ORTHOGONAL_TOLERANCE < 0


def _face_distance(normal, centroid, vertices):
    if vertices:
        return m3.dot(normal, m3.sub(vertices[0], centroid))
    return 0.0


def convex_init(faces, vertices, unique_edges, position, inertia, volume):
    """Convex.elm:297-323.  faces : list of (primary, partner|None), each face = (vertices, normal)."""
    groups = []
    for primary, partner in faces:
        pv, pn = primary
        if partner is not None:
            qv, qn = partner
            groups.append(FaceGroup(pn, [_index_of(v, vertices) for v in pv], _face_distance(pn, position, pv),
                                    qn, [_index_of(v, vertices) for v in qv], _face_distance(pn, position, qv)))
        else:
            groups.append(FaceGroup(pn, [_index_of(v, vertices) for v in pv], _face_distance(pn, position, pv)))
    # indexEdges: outer list reversed, each group inner-reversed (Convex.elm:318, 372-374)
    edges = [list(reversed([_index_of(v, vertices) for v in grp])) for grp in reversed(unique_edges)]
    box = None
    if len(groups) == 3 and all(g.two_sided for g in groups):      # detectObb, Convex.elm:241-252
        n0, n1, n2 = groups[0].normal, groups[1].normal, groups[2].normal
        if _orthogonal(n0, n1) and _orthogonal(n1, n2) and _orthogonal(n0, n2):
            box = (n0, n1, n2, (groups[0].dist, groups[1].dist, groups[2].dist))
    return Convex(groups, edges, list(vertices), box, position, inertia, volume)


def _fold_face_edges(fn, seed, vertices):
    """Convex.elm:1217-1245 : (v0,v1), (v1,v2) ... (last, v0)."""
    if len(vertices) < 2:
        return seed
    acc = seed
    n = len(vertices)
    for i in range(n):
        acc = fn(vertices[i], vertices[(i + 1) % n], acc)
    return acc


def _has_edge(v1, v2, points):
    for i in range(0, len(points) - 1, 2):
        e1, e2 = points[i], points[i + 1]
        if (v1 == e1 and v2 == e2) or (v1 == e2 and v2 == e1):
            return True
    return False


def _add_edge_to_groups(v1, v2, groups):
    """Convex.elm:453-504 : returns the new group list (insertion order preserved)."""
    d = m3.direction(v1, v2)
    for gi, group in enumerate(groups):
        if len(group) < 2:
            continue
        gd = m3.direction(group[0], group[1])
        if m3.length_squared(m3.cross(d, gd)) - m3.PARALLEL_TOLERANCE < 0:
            if _has_edge(v1, v2, group):
                return groups
            return groups[:gi] + [[v2, v1] + group] + groups[gi + 1:]
    return groups + [[v2, v1]]


def group_edges_by_direction(faces):
    """Convex.elm:443-451."""
    groups = []
    for fv, _ in faces:
        groups = _fold_face_edges(_add_edge_to_groups, groups, fv)
    return groups


def group_faces_by_normal(all_faces):
    """Convex.elm:536-575 : first face of a direction is primary, a later (anti)parallel one its partner."""
    groups = []
    for face in all_faces:
        for gi, (primary, _) in enumerate(groups):
            if m3.almost_zero(m3.cross(primary[1], face[1])):
                groups[gi] = (primary, face)
                break
        else:
            groups.append((face, None))
    return groups


def _dedup_vertices(remaining):
    """Convex.elm:380-391 : result is in reverse first-seen order."""
    acc = []
    for v in remaining:
        if _index_of(v, acc) < 0:
            acc.insert(0, v)
    return acc


def from_block(size_x, size_y, size_z):
    """Convex.elm:857-952."""
    x, y, z = size_x / 2, size_y / 2, size_z / 2
    v0, v1, v2, v3 = (-x, -y, -z), (x, -y, -z), (x, y, -z), (-x, y, -z)
    v4, v5, v6, v7 = (-x, -y, z), (x, -y, z), (x, y, z), (-x, y, z)
    volume = size_x * size_y * size_z
    inertia = m3.m_zero()
    inertia["m11"] = volume / 12 * (size_y * size_y + size_z * size_z)
    inertia["m22"] = volume / 12 * (size_x * size_x + size_z * size_z)
    inertia["m33"] = volume / 12 * (size_y * size_y + size_x * size_x)
    xn, yn, zn = (-1.0, 0.0, 0.0), (0.0, -1.0, 0.0), (0.0, 0.0, -1.0)
    faces = [
        (([v4, v5, v6, v7], m3.Z_AXIS), ([v3, v2, v1, v0], zn)),
        (([v2, v3, v7, v6], m3.Y_AXIS), ([v5, v4, v0, v1], yn)),
        (([v1, v2, v6, v5], m3.X_AXIS), ([v0, v4, v7, v3], xn)),
    ]
    unique_edges = [
        [v2, v3, v5, v4, v6, v7, v1, v0],
        [v2, v1, v7, v4, v6, v5, v3, v0],
        [v5, v1, v6, v2, v7, v3, v4, v0],
    ]
    return convex_init(faces, [v0, v1, v2, v3, v4, v5, v6, v7], unique_edges, m3.ZERO, inertia, volume)


def from_cylinder(subdivisions, radius, length):
    """Convex.elm:955-1030 (volume is 2*pi*l*r^2 in the reference — kept)."""
    top, bottom = length * 0.5, length * -0.5
    n = subdivisions
    sides = []
    for value in range(n):
        r0 = 2 * math.pi * (float(value) - 0.5) / float(n)
        r1 = 2 * math.pi * float(value) / float(n)
        r2 = 2 * math.pi * (float(value) + 0.5) / float(n)
        normal = (math.sin(r1), math.cos(r1), 0.0)
        a = (math.sin(r0) * radius, math.cos(r0) * radius, top)
        b = (math.sin(r2) * radius, math.cos(r2) * radius, top)
        c = (math.sin(r2) * radius, math.cos(r2) * radius, bottom)
        d = (math.sin(r0) * radius, math.cos(r0) * radius, bottom)
        sides.append(([a, b, c, d], normal))
    volume = 2 * math.pi * length * radius ** 2

    def cap(z):
        out = []
        for value in range(n):
            r0 = 2 * math.pi * (float(value) - 0.5) / float(n)
            out.append((math.sin(r0) * radius, math.cos(r0) * radius, z))
        return out

    bottom_cap = (cap(bottom), (0.0, 0.0, -1.0))
    top_cap = (list(reversed(cap(top))), m3.Z_AXIS)
    all_faces = [top_cap, bottom_cap] + sides
    vertices = _dedup_vertices([v for f in all_faces for v in f[0]])
    return convex_init(group_faces_by_normal(all_faces), vertices, group_edges_by_direction(all_faces),
                       m3.ZERO, m3.cylinder_inertia(volume, radius, length), volume)


def from_cone(subdivisions, radius, length):
    """Convex.elm:1033-1104."""
    top, bottom = length * 0.5, length * -0.5
    n = subdivisions
    sides = []
    for value in range(n):
        r0 = 2 * math.pi * (float(value) - 0.5) / float(n)
        r1 = 2 * math.pi * float(value) / float(n)
        r2 = 2 * math.pi * (float(value) + 0.5) / float(n)
        normal = (math.sin(r1), math.cos(r1), 0.0)
        a = (0.0, 0.0, top)
        b = (math.sin(r2) * radius, math.cos(r2) * radius, bottom)
        c = (math.sin(r0) * radius, math.cos(r0) * radius, bottom)
        sides.append(([a, b, c], normal))
    volume = (1 / 3) * math.pi * length * radius ** 2
    cap = []
    for value in range(n):
        r0 = 2 * math.pi * (float(value) - 0.5) / float(n)
        cap.append((math.sin(r0) * radius, math.cos(r0) * radius, bottom))
    all_faces = [(cap, (0.0, 0.0, -1.0))] + sides
    vertices = _dedup_vertices([v for f in all_faces for v in f[0]])
    return convex_init(group_faces_by_normal(all_faces), vertices, group_edges_by_direction(all_faces),
                       m3.ZERO, m3.cone_inertia(volume, radius, length), volume)


# ---- fromTriangularMesh (Convex.elm:412-849) -------------------------------------
def compute_normal(v1, v2, v3):
    return m3.normalize(m3.cross(m3.sub(v3, v2), m3.sub(v1, v2)))


def _extend_contour(triangle, contour):
    """Convex.elm:791-849."""
    if len(contour) < 2:
        return contour
    ti1, ti2, ti3 = triangle
    i1 = contour[0]
    result = []
    rest = contour
    while rest:
        ci1 = rest[0]
        rest1 = rest[1:]
        if rest1:
            ci2 = rest1[0]
            if ci1 == ti2 and ci2 == ti1:
                return result + [ci1, ti3] + rest1
            elif ci1 == ti3 and ci2 == ti2:
                return result + [ci1, ti1] + rest1
            elif ci1 == ti1 and ci2 == ti3:
                return result + [ci1, ti2] + rest1
            result.append(ci1)
            rest = rest1
        else:
            if ci1 == ti2 and i1 == ti1:
                return result + [ci1, ti3]
            elif ci1 == ti3 and i1 == ti2:
                return result + [ci1, ti1]
            elif ci1 == ti1 and i1 == ti3:
                return result + [ci1, ti2]
            return result            # reference drops ci1 here (Convex.elm:843-844) — kept
    return result


def init_faces(face_indices, mesh_vertices):
    """Convex.elm:644-788 : merge coplanar adjacent triangles into polygon faces.
    Returns list of (vertices, normal) in the reference's (cons) order."""
    nv = len(mesh_vertices)

    def ok(i):
        return 0 <= i < nv

    face_by_edge = {}
    for tri in face_indices:
        i1, i2, i3 = tri
        if ok(i1) and ok(i2) and ok(i3):
            face = (tri, compute_normal(mesh_vertices[i1], mesh_vertices[i2], mesh_vertices[i3]))
            face_by_edge[(i1, i2)] = face
            face_by_edge[(i2, i3)] = face
            face_by_edge[(i3, i1)] = face
    if not face_indices:
        return []
    i1, i2, i3 = face_indices[0]
    if not (ok(i1) and ok(i2) and ok(i3)):
        return []
    visited = {tuple(face_indices[0])}
    faces_to_check = []
    edges_to_check = [(i2, i1), (i3, i2), (i1, i3)]
    current_normal = compute_normal(mesh_vertices[i1], mesh_vertices[i2], mesh_vertices[i3])
    contour = [i1, i2, i3]
    result = []
    while True:
        adjacent = [face_by_edge[e] for e in edges_to_check if e in face_by_edge]
        adjacent = [f for f in adjacent if tuple(f[0]) not in visited]
        coplanar = [f for f in adjacent if m3.almost_zero(m3.sub(f[1], current_normal))]
        non_coplanar = [f for f in adjacent if not m3.almost_zero(m3.sub(f[1], current_normal))]
        new_visited = set(visited)
        for f in coplanar:
            new_visited.add(tuple(f[0]))
        new_edges = []
        for f in coplanar:
            a, b, c = f[0]
            new_edges = [(b, a), (c, b), (a, c)] + new_edges
        new_faces_to_check = non_coplanar + faces_to_check
        new_contour = contour
        for f in coplanar:
            new_contour = _extend_contour(f[0], new_contour)
        if coplanar:
            visited, faces_to_check, edges_to_check, contour = new_visited, new_faces_to_check, new_edges, new_contour
            continue
        face_to_add = ([mesh_vertices[i] for i in new_contour if ok(i)], current_normal)
        updated = [f for f in new_faces_to_check if tuple(f[0]) not in new_visited]
        if updated:
            tri, normal = updated[0]
            a, b, c = tri
            visited = set(new_visited)
            visited.add(tuple(tri))
            faces_to_check = updated[1:]
            edges_to_check = [(b, a), (c, b), (a, c)]
            current_normal = normal
            contour = [a, b, c]
            result.insert(0, face_to_add)
        else:
            result.insert(0, face_to_add)
            return result


def _convex_mass_properties(center, face_indices, vertices):
    """Convex.elm:578-641."""
    cx = cy = cz = 0.0
    total_volume = 0.0
    total_inertia = m3.m_zero()
    nv = len(vertices)
    for i1, i2, i3 in face_indices:
        if not (0 <= i1 < nv and 0 <= i2 < nv and 0 <= i3 < nv):
            continue
        p1, p2, p3 = vertices[i1], vertices[i2], vertices[i3]
        new_x = (center[0] + p1[0] + p2[0] + p3[0]) / 4
        new_y = (center[1] + p1[1] + p2[1] + p3[1]) / 4
        new_z = (center[2] + p1[2] + p2[2] + p3[2]) / 4
        volume = m3.dot(m3.sub(p1, center), m3.cross(m3.sub(p2, center), m3.sub(p3, center))) / 6
        new_inertia = m3.tetrahedron_inertia(volume, center, p1, p2, p3)
        cx, cy, cz = cx + new_x * volume, cy + new_y * volume, cz + new_z * volume
        total_volume = total_volume + volume
        total_inertia = m3.m_add(total_inertia, new_inertia)
    com = (cx / total_volume, cy / total_volume, cz / total_volume)
    pin = m3.point_inertia(total_volume, com[0] - center[0], com[1] - center[1], com[2] - center[2])
    return total_volume, com, m3.m_sub(total_inertia, pin)


def from_triangular_mesh(face_indices, vertices):
    """Convex.elm:412-440.  vertices: list of (x,y,z); face_indices: list of (i1,i2,i3) CCW."""
    vertices = [tuple(float(c) for c in v) for v in vertices]
    face_indices = [tuple(f) for f in face_indices]
    faces = init_faces(face_indices, vertices)
    if vertices:
        cx, cy, cz = vertices[0]
        n = 1.0
        for v in vertices[1:]:
            n, cx, cy, cz = n + 1, cx + v[0], cy + v[1], cz + v[2]
        avg = (cx / n, cy / n, cz / n)
    else:
        avg = m3.ZERO
    volume, position, inertia = _convex_mass_properties(avg, face_indices, vertices)
    return convex_init(group_faces_by_normal(faces), vertices, group_edges_by_direction(faces),
                       position, inertia, volume)
