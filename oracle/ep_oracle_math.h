// ORACLE — TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).
// Never linked into or called from the product library (elm_physics_b200/csrc).
//
// CPU restatement of the reference's fp64 math layer.  Every function cites the reference
// file:line it follows; expressions keep the reference's operand order and association
// (JS doubles, no FMA: compile with -ffp-contract=off).  Parity status: pinned against the
// reference's own known-answer tests (tests/test_oracle_golden.py); whole-step trajectories are
// "parity unpinned" (the reference ships none and cannot run here — DESIGN.md).
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace epo {

// src/Internal/Const.elm:4-38
constexpr double kMaxNumber = 3.40282347e38;
constexpr double kPrecision = 1.0e-6;
constexpr double kSolverTolerance = 1.0e-6;
constexpr double kContactBreakingThreshold = 1.0e-3;
constexpr double kParallelTolerance = 1.0e-6;

// elm/core Basics (not vendored): abs/min/max defined through < and >.
inline double eabs(double n) { return n < 0 ? -n : n; }
inline double emin(double x, double y) { return x < y ? x : y; }
inline double emax(double x, double y) { return x > y ? x : y; }
inline double eclamp(double lo, double hi, double n) { return n < lo ? lo : (n > hi ? hi : n); }

struct V3 { double x, y, z; };
struct Quat { double x, y, z, w; };
struct Tf { V3 o; Quat q; };
struct M3 { double m11, m21, m31, m12, m22, m32, m13, m23, m33; };

// src/Internal/Vector3.elm:119-236
inline V3 add(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 neg(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 scale(double s, V3 v) { return {s * v.x, s * v.y, s * v.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double lengthSquared(V3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
inline double length(V3 v) { return std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }
inline V3 normalize(V3 v) { double len = length(v); return {v.x / len, v.y / len, v.z / len}; }
// Vector3.elm:146-156 : (a - b) / |a - b|
inline V3 direction(V3 a, V3 b) { V3 c = sub(a, b); double len = length(c); return {c.x / len, c.y / len, c.z / len}; }
// Vector3.elm:178-191 : note b - a
inline double distance(V3 a, V3 b) { double x = b.x - a.x, y = b.y - a.y, z = b.z - a.z; return std::sqrt(x * x + y * y + z * z); }
inline double distanceSquared(V3 a, V3 b) { double x = b.x - a.x, y = b.y - a.y, z = b.z - a.z; return x * x + y * y + z * z; }
// Vector3.elm:42-46
inline bool almostZero(V3 v) { return (eabs(v.x) - kPrecision <= 0) && (eabs(v.y) - kPrecision <= 0) && (eabs(v.z) - kPrecision <= 0); }
// Vector3.elm:293-298
inline V3 lerp(double t, V3 a, V3 b) { return {a.x + t * (b.x - a.x), a.y + t * (b.y - a.y), a.z + t * (b.z - a.z)}; }

// Vector3.elm:239-258
inline void tangents(V3 vec, V3& t1, V3& t2) {
  if (lengthSquared(vec) > 0) {
    V3 n = normalize(vec);
    V3 v = normalize((eabs(n.x) - 0.9 < 0) ? cross(n, V3{1, 0, 0}) : cross(n, V3{0, 1, 0}));
    t1 = v;
    t2 = cross(n, v);
  } else {
    t1 = V3{1, 0, 0};
    t2 = V3{0, 1, 0};
  }
}

// Vector3.elm:270-290
inline void stableTangents(V3 cachedT1, V3 ni, V3& t1, V3& t2) {
  double d = dot(cachedT1, ni);
  V3 projected = sub(cachedT1, scale(d, ni));
  double lenSq = lengthSquared(projected);
  if (lenSq - kPrecision < 0) {
    tangents(ni, t1, t2);
  } else {
    t1 = scale(1 / std::sqrt(lenSq), projected);
    t2 = cross(ni, t1);
  }
}

// Vector3.elm:305-370
inline void closestPointsBetweenSegments(V3 p1, V3 q1, V3 p2, V3 q2, V3& c1, V3& c2) {
  V3 d1 = sub(q1, p1), d2 = sub(q2, p2), r = sub(p1, p2);
  double a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
  double s, t;
  if (a - kPrecision <= 0 && e - kPrecision <= 0) {
    s = 0; t = 0;
  } else if (a - kPrecision <= 0) {
    s = 0; t = eclamp(0, 1, f / e);
  } else {
    double c = dot(d1, r);
    if (e - kPrecision <= 0) {
      s = eclamp(0, 1, -c / a); t = 0;
    } else {
      double b = dot(d1, d2);
      double denom = a * e - b * b;
      double sInit = (denom != 0.0) ? eclamp(0, 1, (b * f - c * e) / denom) : 0.0;
      double tNom = b * sInit + f;
      if (tNom < 0.0) { s = eclamp(0, 1, -c / a); t = 0; }
      else if (tNom - e > 0) { s = eclamp(0, 1, (b - c) / a); t = 1; }
      else { s = sInit; t = tNom / e; }
    }
  }
  c1 = add(p1, scale(s, d1));
  c2 = add(p2, scale(t, d2));
}

// src/Internal/Transform3d.elm:134-153
inline V3 pointPlaceIn(const Tf& t, V3 p) {
  double qx = t.q.x, qy = t.q.y, qz = t.q.z, qw = t.q.w, x = p.x, y = p.y, z = p.z;
  double ix = qw * x + qy * z - qz * y;
  double iy = qw * y + qz * x - qx * z;
  double iz = qw * z + qx * y - qy * x;
  double iw = -qx * x - qy * y - qz * z;
  return {ix * qw + iw * -qx + iy * -qz - iz * -qy + t.o.x,
          iy * qw + iw * -qy + iz * -qx - ix * -qz + t.o.y,
          iz * qw + iw * -qz + ix * -qy - iy * -qx + t.o.z};
}
// Transform3d.elm:415-432
inline V3 rotate(Quat q, V3 v) {
  double ix = 2 * (q.y * v.z - q.z * v.y);
  double iy = 2 * (q.z * v.x - q.x * v.z);
  double iz = 2 * (q.x * v.y - q.y * v.x);
  return {v.x + q.w * ix + q.y * iz - q.z * iy, v.y + q.w * iy + q.z * ix - q.x * iz, v.z + q.w * iz + q.x * iy - q.y * ix};
}
inline V3 directionPlaceIn(const Tf& t, V3 v) { return rotate(t.q, v); }
// Transform3d.elm:368-376
inline Tf rotateBy(V3 w, const Tf& t) {
  double qx = t.q.x, qy = t.q.y, qz = t.q.z, qw = t.q.w;
  return {t.o, {qx + (w.x * qw + w.y * qz - w.z * qy) * 0.5, qy + (w.y * qw + w.z * qx - w.x * qz) * 0.5,
                qz + (w.z * qw + w.x * qy - w.y * qx) * 0.5, qw + (-w.x * qx - w.y * qy - w.z * qz) * 0.5}};
}
inline Tf translateBy(V3 v, const Tf& t) { return {add(v, t.o), t.q}; }
// Transform3d.elm:256-262
inline Tf normalizeTf(const Tf& t) {
  double x = t.q.x, y = t.q.y, z = t.q.z, w = t.q.w;
  double len = std::sqrt(x * x + y * y + z * z + w * w);
  return {t.o, {x / len, y / len, z / len, w / len}};
}
// Transform3d.elm:334-345 + 181-205
inline M3 invertedInertiaRotateIn(const Tf& t, V3 inv) {
  double x = t.q.x, y = t.q.y, z = t.q.z, w = t.q.w;
  double m11 = 1 - 2 * y * y - 2 * z * z, m12 = 2 * x * y - 2 * w * z, m13 = 2 * x * z + 2 * w * y;
  double m21 = 2 * x * y + 2 * w * z, m22 = 1 - 2 * x * x - 2 * z * z, m23 = 2 * y * z - 2 * w * x;
  double m31 = 2 * x * z - 2 * w * y, m32 = 2 * y * z + 2 * w * x, m33 = 1 - 2 * x * x - 2 * y * y;
  double a = inv.x, b = inv.y, c = inv.z;
  M3 r;
  r.m11 = m11 * m11 * a + m12 * m12 * b + m13 * m13 * c;
  r.m21 = m21 * m11 * a + m22 * m12 * b + m23 * m13 * c;
  r.m31 = m31 * m11 * a + m32 * m12 * b + m33 * m13 * c;
  r.m12 = m11 * m21 * a + m12 * m22 * b + m13 * m23 * c;
  r.m22 = m21 * m21 * a + m22 * m22 * b + m23 * m23 * c;
  r.m32 = m31 * m21 * a + m32 * m22 * b + m33 * m23 * c;
  r.m13 = m11 * m31 * a + m12 * m32 * b + m13 * m33 * c;
  r.m23 = m21 * m31 * a + m22 * m32 * b + m23 * m33 * c;
  r.m33 = m31 * m31 * a + m32 * m32 * b + m33 * m33 * c;
  return r;
}

}  // namespace epo
