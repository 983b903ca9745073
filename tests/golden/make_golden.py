"""Generates tests/golden/*.json from the reference's own test fixtures.

Runs ONLY in the build container (reads /root/reference, which does not exist on the GPU box);
the JSON it writes is committed.  Nothing is computed by the reference here (it cannot run: no
elm/node) — the script transcribes the reference's hand-written expectations:
  * tests/Fixtures/NarrowPhase.elm:11-177  -> sphere_convex.json (38 box + 32 octahedron placements)
  * tests/ConvexSphereBoxTest.elm:31-147   -> icosphere.json (mesh used by the full-precision edge test)
  * tests/Fixtures/Convex.elm:178-369      -> huge_convex.json
"""
import json
import math
import os
import re

REF = "/root/reference/tests"
HERE = os.path.dirname(os.path.abspath(__file__))

AXES = {"Vec3.xAxis": (1, 0, 0), "Vec3.yAxis": (0, 1, 0), "Vec3.zAxis": (0, 0, 1),
        "Vec3.xNegative": (-1, 0, 0), "Vec3.yNegative": (0, -1, 0), "Vec3.zNegative": (0, 0, -1)}
REC = r"\{ x = (?P<x>[^,]+), y = (?P<y>[^,]+), z = (?P<z>[^}]+) \}"


def _vec(txt, env):
    txt = txt.strip()
    if txt in AXES:
        return [float(c) for c in AXES[txt]]
    m = re.fullmatch(REC, txt)
    return [float(eval(m.group(k), {}, env)) for k in "xyz"]


def _ceq_lines(src, fn_name):
    start = src.index(fn_name + " center radius")
    end = src.index("\n\n\n", start)
    return [l for l in src[start:end].splitlines() if "ceq {" in l]


def _field(line, name):
    m = re.search(name + r" = (" + REC + r"|Vec3\.\w+|center)", line)
    return m.group(1)


def sphere_fixture(fn_name, env, center, radius):
    src = open(os.path.join(REF, "Fixtures/NarrowPhase.elm")).read()
    out = []
    for line in _ceq_lines(src, fn_name):
        cj = _vec(_field(line, "cj"), env)
        cj = [cj[i] + center[i] for i in range(3)]        # `|> Vec3.add center`  (cj + center)
        ni = _vec(_field(line, "ni"), env)
        rj = _vec(_field(line, "rj"), env)
        pi = [center[i] + radius * ni[i] for i in range(3)]
        pj = [cj[i] + rj[i] for i in range(3)]
        out.append(dict(position=cj, ni=ni, pi=pi, pj=pj))
    return out


def box_env(radius, box_size):
    precision = 1.0e-6
    h = box_size / 2
    delta = 3 * precision
    vd = h + radius / math.sqrt(3) - precision
    return dict(boxHalfExtent=h, delta=delta, nearEdgeOffset=h - delta, vertexDimension=vd,
                offDiagonalFactor=(vd + delta) / vd, edgeDimension=h + radius / math.sqrt(2) - precision,
                faceDimension=h + radius - precision, invSqrt3=1 / math.sqrt(3), invSqrt2=1 / math.sqrt(2))


def octo_env(radius, he):
    precision = 1.0e-6
    return dict(octoHalfExtent=he, delta=3 * precision, vertexDimension=he + radius - precision,
                edgeDimension=he / 2 + radius / math.sqrt(2) - precision,
                faceDimension=he / 3 + radius / math.sqrt(3) - precision, invSqrt3=1 / math.sqrt(3),
                invSqrt2=1 / math.sqrt(2))


def mesh(path, start_marker):
    src = open(os.path.join(REF, path)).read()
    src = src[src.index(start_marker):]
    verts = [[float(m.group(k)) for k in "xyz"] for m in re.finditer(REC, src[:src.index("\n\n\n")] if False else src)]
    faces = [[int(a), int(b), int(c)] for a, b, c in re.findall(r"\( (\d+), (\d+), (\d+) \)", src)]
    return verts, faces


def main():
    center, radius, box_size, octo_he = [0.0, 0.0, 7.0], 5.0, 6.0, 1.0
    data = dict(center=center, radius=radius, box_size=box_size, octo_half_extent=octo_he,
                box=sphere_fixture("sphereContactBoxPositions", box_env(radius, box_size), center, radius),
                box_far=sphere_fixture("sphereContactBoxPositions", box_env(radius * 2, box_size), center, radius * 2),
                octo=sphere_fixture("sphereContactOctohedronPositions", octo_env(radius, octo_he), center, radius),
                octo_far=sphere_fixture("sphereContactOctohedronPositions", octo_env(radius * 2, octo_he), center, radius * 2))
    assert len(data["box"]) == 38 and len(data["octo"]) == 32, (len(data["box"]), len(data["octo"]))
    json.dump(data, open(os.path.join(HERE, "sphere_convex.json"), "w"), indent=0)

    src = open(os.path.join(REF, "ConvexSphereBoxTest.elm")).read()
    a, b = src.index("sphereVertices ="), src.index("controlledConvex =")
    part = src[a:b]
    verts = [[float(m.group(k)) for k in "xyz"] for m in re.finditer(REC, part)]
    faces = [[int(x), int(y), int(z)] for x, y, z in re.findall(r"\( (\d+), (\d+), (\d+) \)", part)]
    assert len(verts) == 42 and len(faces) == 80, (len(verts), len(faces))
    json.dump(dict(vertices=verts, faces=faces), open(os.path.join(HERE, "icosphere.json"), "w"))

    src = open(os.path.join(REF, "Fixtures/Convex.elm")).read()
    part = src[src.index("hugeConvex =\n"):]
    verts = [[float(m.group(k)) for k in "xyz"] for m in re.finditer(REC, part)]
    faces = [[int(x), int(y), int(z)] for x, y, z in re.findall(r"\( (\d+), (\d+), (\d+) \)", part)]
    json.dump(dict(vertices=verts, faces=faces), open(os.path.join(HERE, "huge_convex.json"), "w"))
    print("icosphere", 42, 80, "huge", len(verts), len(faces))


if __name__ == "__main__":
    main()
