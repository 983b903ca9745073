// fp64 device math for the stepping kernels.  Compiled with -fmad=false: the reference runs on JS
// doubles (no FMA, round-to-nearest), so every expression keeps the reference's operand order and
// association; DDIV / DSQRT are IEEE-rounded on sm_100a.  Reference: src/Internal/Vector3.elm,
// src/Internal/Transform3d.elm, src/Internal/Const.elm (file:line beside each function).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define EPD __device__ __forceinline__

namespace ep {

// Const.elm:4-38
constexpr double kMaxNumber = 3.40282347e38;
constexpr double kPrecision = 1.0e-6;
constexpr double kSolverTolerance = 1.0e-6;
constexpr double kBreaking = 1.0e-3;          // contactBreakingThreshold
constexpr double kParallelTol = 1.0e-6;

struct V3 { double x, y, z; };
struct Q4 { double x, y, z, w; };
struct M33 { double m11, m21, m31, m12, m22, m32, m13, m23, m33; };

// elm/core Basics: abs/min/max/clamp via < and > (ties and NaN return the second argument)
EPD double eabs(double n) { return n < 0 ? -n : n; }
EPD double emin(double x, double y) { return x < y ? x : y; }
EPD double emax(double x, double y) { return x > y ? x : y; }
EPD double eclamp01(double n) { return n < 0 ? 0.0 : (n > 1 ? 1.0 : n); }

EPD V3 mk(double x, double y, double z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
EPD V3 ld3(const double* p) { return mk(p[0], p[1], p[2]); }
EPD void st3(double* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
// Vector3.elm:119-236
EPD V3 vadd(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
EPD V3 vsub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
EPD V3 vneg(V3 a) { return mk(-a.x, -a.y, -a.z); }
EPD V3 vscale(double s, V3 v) { return mk(s * v.x, s * v.y, s * v.z); }
EPD double vdot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
EPD V3 vcross(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
EPD double vlen2(V3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
EPD double vlen(V3 v) { return sqrt(v.x * v.x + v.y * v.y + v.z * v.z); }
EPD V3 vnormalize(V3 v) { double l = vlen(v); return mk(v.x / l, v.y / l, v.z / l); }
EPD V3 vdirection(V3 a, V3 b) { return vnormalize(vsub(a, b)); }                      // (a-b)/|a-b|
EPD double vdist2(V3 a, V3 b) { double x = b.x - a.x, y = b.y - a.y, z = b.z - a.z; return x * x + y * y + z * z; }
EPD double vdist(V3 a, V3 b) { return sqrt(vdist2(a, b)); }
EPD bool valmost0(V3 v) { return (eabs(v.x) - kPrecision <= 0) && (eabs(v.y) - kPrecision <= 0) && (eabs(v.z) - kPrecision <= 0); }
EPD V3 vlerp(double t, V3 a, V3 b) { return mk(a.x + t * (b.x - a.x), a.y + t * (b.y - a.y), a.z + t * (b.z - a.z)); }

// Vector3.elm:239-258
EPD void tangents(V3 vec, V3& t1, V3& t2) {
  if (vlen2(vec) > 0) {
    V3 n = vnormalize(vec);
    V3 c = (eabs(n.x) - 0.9 < 0) ? vcross(n, mk(1, 0, 0)) : vcross(n, mk(0, 1, 0));
    t1 = vnormalize(c);
    t2 = vcross(n, t1);
  } else { t1 = mk(1, 0, 0); t2 = mk(0, 1, 0); }
}
// Vector3.elm:270-290
EPD void stable_tangents(V3 cached, V3 ni, V3& t1, V3& t2) {
  double d = vdot(cached, ni);
  V3 proj = vsub(cached, vscale(d, ni));
  double l2 = vlen2(proj);
  if (l2 - kPrecision < 0) tangents(ni, t1, t2);
  else { t1 = vscale(1 / sqrt(l2), proj); t2 = vcross(ni, t1); }
}
// Vector3.elm:305-370 (Ericson RTCD 5.1.9)
EPD void closest_segments(V3 p1, V3 q1, V3 p2, V3 q2, V3& c1, V3& c2) {
  V3 d1 = vsub(q1, p1), d2 = vsub(q2, p2), r = vsub(p1, p2);
  double a = vdot(d1, d1), e = vdot(d2, d2), f = vdot(d2, r), s, t;
  bool a0 = a - kPrecision <= 0, e0 = e - kPrecision <= 0;
  if (a0 && e0) { s = 0; t = 0; }
  else if (a0) { s = 0; t = eclamp01(f / e); }
  else {
    double c = vdot(d1, r);
    if (e0) { s = eclamp01(-c / a); t = 0; }
    else {
      double b = vdot(d1, d2), denom = a * e - b * b;
      double s0 = (denom != 0.0) ? eclamp01((b * f - c * e) / denom) : 0.0;
      double tn = b * s0 + f;
      if (tn < 0.0) { s = eclamp01(-c / a); t = 0; }
      else if (tn - e > 0) { s = eclamp01((b - c) / a); t = 1; }
      else { s = s0; t = tn / e; }
    }
  }
  c1 = vadd(p1, vscale(s, d1));
  c2 = vadd(p2, vscale(t, d2));
}

// Transform3d.elm:134-153 (points) and 415-432 (directions) — two different formulas on purpose
EPD V3 point_place(V3 o, Q4 q, V3 p) {
  double ix = q.w * p.x + q.y * p.z - q.z * p.y;
  double iy = q.w * p.y + q.z * p.x - q.x * p.z;
  double iz = q.w * p.z + q.x * p.y - q.y * p.x;
  double iw = -q.x * p.x - q.y * p.y - q.z * p.z;
  return mk(ix * q.w + iw * -q.x + iy * -q.z - iz * -q.y + o.x,
            iy * q.w + iw * -q.y + iz * -q.x - ix * -q.z + o.y,
            iz * q.w + iw * -q.z + ix * -q.y - iy * -q.x + o.z);
}
EPD V3 rotate(Q4 q, V3 v) {
  double ix = 2 * (q.y * v.z - q.z * v.y);
  double iy = 2 * (q.z * v.x - q.x * v.z);
  double iz = 2 * (q.x * v.y - q.y * v.x);
  return mk(v.x + q.w * ix + q.y * iz - q.z * iy, v.y + q.w * iy + q.z * ix - q.x * iz, v.z + q.w * iz + q.x * iy - q.y * ix);
}
// Transform3d.elm:334-345 then 181-205
EPD M33 inv_inertia_world(Q4 q, V3 inv) {
  double x = q.x, y = q.y, z = q.z, w = q.w;
  double m11 = 1 - 2 * y * y - 2 * z * z, m12 = 2 * x * y - 2 * w * z, m13 = 2 * x * z + 2 * w * y;
  double m21 = 2 * x * y + 2 * w * z, m22 = 1 - 2 * x * x - 2 * z * z, m23 = 2 * y * z - 2 * w * x;
  double m31 = 2 * x * z - 2 * w * y, m32 = 2 * y * z + 2 * w * x, m33 = 1 - 2 * x * x - 2 * y * y;
  double a = inv.x, b = inv.y, c = inv.z;
  M33 r;
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

}  // namespace ep
