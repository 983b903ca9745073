"""fp64 vector / quaternion / 3x3 helpers for the HOST side (scene construction, packing).

Construction-time only: nothing here runs per step.  Every expression keeps the
operand order of the reference so that shapes built here equal the ones the Elm
constructors build (reference: src/Internal/Vector3.elm, src/Internal/Transform3d.elm,
src/Internal/Matrix3.elm).  Vec3 = (x, y, z) tuple, quaternion = (x, y, z, w) tuple,
Transform3d = (origin, quaternion), Mat3 = dict m11..m33.
"""
import math

MAX_NUMBER = 3.40282347e38          # Const.elm:4-6
PRECISION = 1.0e-6                  # Const.elm:12-14
PARALLEL_TOLERANCE = 1.0e-6         # Const.elm:36-38

ZERO = (0.0, 0.0, 0.0)
X_AXIS = (1.0, 0.0, 0.0)
Y_AXIS = (0.0, 1.0, 0.0)
Z_AXIS = (0.0, 0.0, 1.0)
IDENTITY_Q = (0.0, 0.0, 0.0, 1.0)
AT_ORIGIN = (ZERO, IDENTITY_Q)


# ---- Vector3.elm -----------------------------------------------------------
def add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def neg(a):
    return (-a[0], -a[1], -a[2])


def scale(s, v):
    return (s * v[0], s * v[1], s * v[2])


def dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def length_squared(v):
    return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]


def length(v):
    return math.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


def normalize(v):
    ln = length(v)
    return (v[0] / ln, v[1] / ln, v[2] / ln)


def direction(a, b):
    """Vector3.elm:146-156 : (a - b) / |a - b|."""
    return normalize(sub(a, b))


def almost_zero(v):
    """Vector3.elm:42-46."""
    return (abs(v[0]) - PRECISION <= 0) and (abs(v[1]) - PRECISION <= 0) and (abs(v[2]) - PRECISION <= 0)


# ---- Transform3d.elm -------------------------------------------------------
def q_mul(q1, q2):
    """Transform3d.elm:406-412."""
    q1x, q1y, q1z, q1w = q1
    q2x, q2y, q2z, q2w = q2
    return (
        q1x * q2w + q1y * q2z - q1z * q2y + q1w * q2x,
        -q1x * q2z + q1y * q2w + q1z * q2x + q1w * q2y,
        q1x * q2y - q1y * q2x + q1z * q2w + q1w * q2z,
        -q1x * q2x - q1y * q2y - q1z * q2z + q1w * q2w,
    )


def q_rotate(q, v):
    """Transform3d.elm:415-432 (directions, normals, axes)."""
    qx, qy, qz, qw = q
    x, y, z = v
    ix = 2 * (qy * z - qz * y)
    iy = 2 * (qz * x - qx * z)
    iz = 2 * (qx * y - qy * x)
    return (
        x + qw * ix + qy * iz - qz * iy,
        y + qw * iy + qz * ix - qx * iz,
        z + qw * iz + qx * iy - qy * ix,
    )


def q_derotate(q, v):
    """Transform3d.elm:435-453."""
    qx, qy, qz, qw = q
    x, y, z = v
    ix = -qw * x + qy * z - qz * y
    iy = -qw * y + qz * x - qx * z
    iz = -qw * z + qx * y - qy * x
    iw = -qx * x - qy * y - qz * z
    return (
        ix * -qw + iw * -qx + iy * -qz - iz * -qy,
        iy * -qw + iw * -qy + iz * -qx - ix * -qz,
        iz * -qw + iw * -qz + ix * -qy - iy * -qx,
    )


def point_place_in(t, p):
    """Transform3d.elm:134-153 (points, vertices)."""
    (ox, oy, oz), (qx, qy, qz, qw) = t
    x, y, z = p
    ix = qw * x + qy * z - qz * y
    iy = qw * y + qz * x - qx * z
    iz = qw * z + qx * y - qy * x
    iw = -qx * x - qy * y - qz * z
    return (
        ix * qw + iw * -qx + iy * -qz - iz * -qy + ox,
        iy * qw + iw * -qy + iz * -qx - ix * -qz + oy,
        iz * qw + iw * -qz + ix * -qy - iy * -qx + oz,
    )


def direction_place_in(t, v):
    return q_rotate(t[1], v)


def point_relative_to(t, p):
    return q_derotate(t[1], sub(p, t[0]))


def direction_relative_to(t, v):
    return q_derotate(t[1], v)


def t_place_in(g, l):
    """Transform3d.elm:222-230."""
    return (add(g[0], q_rotate(g[1], l[0])), q_mul(g[1], l[1]))


def t_relative_to(t1, t2):
    """Transform3d.elm:233-241 : t2 expressed in the frame t1 (conjugate of a normalised quaternion as its inverse)."""
    x, y, z, w = t1[1]
    return (point_relative_to(t1, t2[0]), q_mul((-x, -y, -z, w), t2[1]))


def t_inverse(t):
    """Transform3d.elm:245-253."""
    x, y, z, w = t[1]
    return (point_relative_to(t, ZERO), (-x, -y, -z, w))


def t_at_point(p):
    return (p, IDENTITY_Q)


def t_normalize(t):
    x, y, z, w = t[1]
    ln = math.sqrt(x * x + y * y + z * z + w * w)
    return (t[0], (x / ln, y / ln, z / ln, w / ln))


def from_angle_axis(angle, axis):
    """Transform3d.elm:388-403."""
    x, y, z = normalize(axis)
    theta = angle * 0.5
    c = math.cos(theta)
    s = math.sin(theta)
    return (x * s, y * s, z * s, c)


def rotate_around_own(axis, angle, t):
    return (t[0], q_mul(from_angle_axis(angle, axis), t[1]))


def from_origin_and_basis(origin, x, y, z):
    """Transform3d.elm:34-121."""
    m00, m10, m20 = x
    m01, m11, m21 = y
    m02, m12, m22 = z
    tr = m00 + m11 + m22
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        return (origin, ((m21 - m12) / s, (m02 - m20) / s, (m10 - m01) / s, 0.25 * s))
    elif (m00 - m11 > 0) and (m00 - m22 > 0):
        s = math.sqrt(1.0 + m00 - m11 - m22) * 2
        return (origin, (0.25 * s, (m01 + m10) / s, (m02 + m20) / s, (m21 - m12) / s))
    elif m11 - m22 > 0:
        s = math.sqrt(1.0 + m11 - m00 - m22) * 2
        return (origin, ((m01 + m10) / s, 0.25 * s, (m12 + m21) / s, (m02 - m20) / s))
    else:
        s = math.sqrt(1.0 + m22 - m00 - m11) * 2
        return (origin, ((m02 + m20) / s, (m12 + m21) / s, 0.25 * s, (m10 - m01) / s))


def orientation(t):
    """Transform3d.elm:334-345 : rotation matrix of the quaternion."""
    x, y, z, w = t[1]
    return dict(
        m11=1 - 2 * y * y - 2 * z * z, m12=2 * x * y - 2 * w * z, m13=2 * x * z + 2 * w * y,
        m21=2 * x * y + 2 * w * z, m22=1 - 2 * x * x - 2 * z * z, m23=2 * y * z - 2 * w * x,
        m31=2 * x * z - 2 * w * y, m32=2 * y * z + 2 * w * x, m33=1 - 2 * x * x - 2 * y * y,
    )


def inverted_inertia_rotate_in(t, inv_inertia):
    """Transform3d.elm:181-205."""
    o = orientation(t)
    m11, m21, m31 = o["m11"], o["m21"], o["m31"]
    m12, m22, m32 = o["m12"], o["m22"], o["m32"]
    m13, m23, m33 = o["m13"], o["m23"], o["m33"]
    a, b, c = inv_inertia
    return dict(
        m11=m11 * m11 * a + m12 * m12 * b + m13 * m13 * c,
        m21=m21 * m11 * a + m22 * m12 * b + m23 * m13 * c,
        m31=m31 * m11 * a + m32 * m12 * b + m33 * m13 * c,
        m12=m11 * m21 * a + m12 * m22 * b + m13 * m23 * c,
        m22=m21 * m21 * a + m22 * m22 * b + m23 * m23 * c,
        m32=m31 * m21 * a + m32 * m22 * b + m33 * m23 * c,
        m13=m11 * m31 * a + m12 * m32 * b + m13 * m33 * c,
        m23=m21 * m31 * a + m22 * m32 * b + m23 * m33 * c,
        m33=m31 * m31 * a + m32 * m32 * b + m33 * m33 * c,
    )


# ---- Matrix3.elm -----------------------------------------------------------
_KEYS = ("m11", "m21", "m31", "m12", "m22", "m32", "m13", "m23", "m33")


def m_zero():
    return {k: 0.0 for k in _KEYS}


def m_add(a, b):
    return {k: a[k] + b[k] for k in _KEYS}


def m_sub(a, b):
    return {k: a[k] - b[k] for k in _KEYS}


def m_scale(s, m):
    return {k: s * m[k] for k in _KEYS}


def m_transpose(m):
    return dict(m11=m["m11"], m21=m["m12"], m31=m["m13"], m12=m["m21"], m22=m["m22"], m32=m["m23"],
                m13=m["m31"], m23=m["m32"], m33=m["m33"])


def m_mul(a, b):
    """Matrix3.elm:88-99."""
    return dict(
        m11=a["m11"] * b["m11"] + a["m12"] * b["m21"] + a["m13"] * b["m31"],
        m21=a["m21"] * b["m11"] + a["m22"] * b["m21"] + a["m23"] * b["m31"],
        m31=a["m31"] * b["m11"] + a["m32"] * b["m21"] + a["m33"] * b["m31"],
        m12=a["m11"] * b["m12"] + a["m12"] * b["m22"] + a["m13"] * b["m32"],
        m22=a["m21"] * b["m12"] + a["m22"] * b["m22"] + a["m23"] * b["m32"],
        m32=a["m31"] * b["m12"] + a["m32"] * b["m22"] + a["m33"] * b["m32"],
        m13=a["m11"] * b["m13"] + a["m12"] * b["m23"] + a["m13"] * b["m33"],
        m23=a["m21"] * b["m13"] + a["m22"] * b["m23"] + a["m23"] * b["m33"],
        m33=a["m31"] * b["m13"] + a["m32"] * b["m23"] + a["m33"] * b["m33"],
    )


def point_inertia(m, x, y, z):
    """Matrix3.elm:160-181."""
    m21 = -m * x * y
    m31 = -m * x * z
    m32 = -m * y * z
    return dict(m11=m * (y * y + z * z), m21=m21, m31=m31, m12=m21, m22=m * (z * z + x * x), m32=m32,
                m13=m31, m23=m32, m33=m * (x * x + y * y))


def _diag(a, b, c):
    m = m_zero()
    m["m11"], m["m22"], m["m33"] = a, b, c
    return m


def sphere_inertia(m, radius):
    i = m * 2 / 5 * radius * radius
    return _diag(i, i, i)


def cylinder_inertia(m, radius, height):
    a = (m * (3 * radius ** 2 + height ** 2)) / 12
    return _diag(a, a, (m * radius ** 2) / 2)


def cone_inertia(m, radius, height):
    a = (3 * m * (4 * radius ** 2 + height ** 2)) / 80
    return _diag(a, a, (3 * m * radius ** 2) / 10)


def capsule_inertia(m, radius, half_length):
    """Matrix3.elm:470-521."""
    r, h = radius, half_length
    v_cyl = 2 * math.pi * r * r * h
    v_sph = (4 / 3) * math.pi * r * r * r
    v_total = v_cyl + v_sph
    m_cyl = m * v_cyl / v_total
    m_sph = m * v_sph / v_total
    iz = m_cyl * r * r / 2 + m_sph * r * r * 2 / 5
    ixy = m_cyl * (3 * r * r + 4 * h * h) / 12 + m_sph * (2 * r * r / 5 + h * h + 3 * r * h / 4)
    return _diag(ixy, ixy, iz)


def tetrahedron_inertia(m, p0, p1, p2, p3):
    """Matrix3.elm:239-306."""
    x1, x2, x3 = p1[0] - p0[0], p2[0] - p0[0], p3[0] - p0[0]
    y1, y2, y3 = p1[1] - p0[1], p2[1] - p0[1], p3[1] - p0[1]
    z1, z2, z3 = p1[2] - p0[2], p2[2] - p0[2], p3[2] - p0[2]
    ix = m / 10 * (x1 * x1 + x2 * x2 + x3 * x3 + x1 * x2 + x1 * x3 + x2 * x3)
    iy = m / 10 * (y1 * y1 + y2 * y2 + y3 * y3 + y1 * y2 + y1 * y3 + y2 * y3)
    iz = m / 10 * (z1 * z1 + z2 * z2 + z3 * z3 + z1 * z2 + z1 * z3 + z2 * z3)
    ixx, iyy, izz = iy + iz, ix + iz, ix + iy
    ixy = m / 20 * (2 * (x1 * y1 + x2 * y2 + x3 * y3) + x1 * y2 + x2 * y1 + x1 * y3 + x3 * y1 + x2 * y3 + x3 * y2)
    iyz = m / 20 * (2 * (z1 * y1 + z2 * y2 + z3 * y3) + z1 * y2 + z2 * y1 + z1 * y3 + z3 * y1 + z2 * y3 + z3 * y2)
    izx = m / 20 * (2 * (x1 * z1 + x2 * z2 + x3 * z3) + x1 * z2 + x2 * z1 + x1 * z3 + x3 * z1 + x2 * z3 + x3 * z2)
    return dict(m11=ixx, m12=-ixy, m13=-izx, m21=-ixy, m22=iyy, m23=-iyz, m31=-izx, m32=-iyz, m33=izz)


def inertia_rotate_in(t, inertia):
    """Transform3d.elm:172-178."""
    rot = orientation(t)
    return m_mul(rot, m_mul(inertia, m_transpose(rot)))


def inertia_place_in(t, center_of_mass, mass, inertia):
    """Transform3d.elm:156-169."""
    x, y, z = t[0]
    rotated = inertia_rotate_in(t, inertia)
    offset = point_inertia(mass, x - center_of_mass[0], y - center_of_mass[1], z - center_of_mass[2])
    return m_add(rotated, offset)


_JACOBI_TOL = 1.0e-12


def _jacobi_t(theta):
    if theta >= 0:
        return 1 / (theta + math.sqrt(1 + theta * theta))
    return 1 / (theta - math.sqrt(1 + theta * theta))


def eigen_decomposition(m):
    """Matrix3.elm:311-464 : Jacobi sweeps on a symmetric 3x3 (body-construction only)."""
    d11, d22, d33 = m["m11"], m["m22"], m["m33"]
    a12, a13, a23 = m["m12"], m["m13"], m["m23"]
    r11, r21, r31, r12, r22, r32, r13, r23, r33 = 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0
    steps = 30
    while True:
        abs12, abs13, abs23 = abs(a12), abs(a13), abs(a23)
        if steps <= 0 or (abs12 - _JACOBI_TOL < 0 and abs13 - _JACOBI_TOL < 0 and abs23 - _JACOBI_TOL < 0):
            break
        steps -= 1
        if abs12 - abs13 >= 0 and abs12 - abs23 >= 0:
            theta = (d22 - d11) / (2 * a12)
            t = _jacobi_t(theta)
            c = 1 / math.sqrt(1 + t * t)
            k = t * c
            (d11, d22, a12, a13, a23,
             r11, r21, r31, r12, r22, r32) = (
                d11 - t * a12, d22 + t * a12, 0.0, c * a13 - k * a23, k * a13 + c * a23,
                c * r11 - k * r12, c * r21 - k * r22, c * r31 - k * r32,
                k * r11 + c * r12, k * r21 + c * r22, k * r31 + c * r32)
        elif abs13 - abs23 >= 0:
            theta = (d33 - d11) / (2 * a13)
            t = _jacobi_t(theta)
            c = 1 / math.sqrt(1 + t * t)
            k = t * c
            (d11, d33, a12, a13, a23,
             r11, r21, r31, r13, r23, r33) = (
                d11 - t * a13, d33 + t * a13, c * a12 - k * a23, 0.0, k * a12 + c * a23,
                c * r11 - k * r13, c * r21 - k * r23, c * r31 - k * r33,
                k * r11 + c * r13, k * r21 + c * r23, k * r31 + c * r33)
        else:
            theta = (d33 - d22) / (2 * a23)
            t = _jacobi_t(theta)
            c = 1 / math.sqrt(1 + t * t)
            k = t * c
            (d22, d33, a12, a13, a23,
             r12, r22, r32, r13, r23, r33) = (
                d22 - t * a23, d33 + t * a23, c * a12 - k * a13, k * a12 + c * a13, 0.0,
                c * r12 - k * r13, c * r22 - k * r23, c * r32 - k * r33,
                k * r12 + c * r13, k * r22 + c * r23, k * r32 + c * r33)
    return dict(eigenvalues=(d11, d22, d33), v1=(r11, r21, r31), v2=(r12, r22, r32), v3=(r13, r23, r33))
