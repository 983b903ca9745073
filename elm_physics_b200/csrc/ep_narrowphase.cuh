// Narrowphase shape matrix, one thread per candidate body pair (device functions).
// Reference: src/Internal/NarrowPhase.elm:86-286 (dispatch), src/Collision/*.elm (13 modules),
// src/Internal/Manifold.elm, src/Internal/ContactId.elm.  Contacts are appended in EMISSION order;
// the caller reverses per shape pair to get the solver order (src/Internal/Equation.elm:135-154).
#pragma once
#include "ep_math.cuh"

namespace ep {

enum { SH_CONVEX = 0, SH_PLANE = 1, SH_SPHERE = 2, SH_CAPSULE = 3, SH_PARTICLE = 4 };
constexpr int kCvxF64Header = 15, kCvxI32Header = 4, kGroupF64 = 8, kGroupI32 = 5, kEdgeI32 = 2;
constexpr int kMaxPoly = 40;      // clip polygon / manifold candidate capacity per thread
constexpr int kMaxPairContacts = 24;

struct RawContact { int64_t fk; V3 ni, pi, pj; };

// hint / sep: temporal coherence of the SAT modules (ConvexConvex, CapsuleConvex).  Every axis of the reference's axis set that
// separates yields the same result -- no contacts -- whichever comes first, so the axis that separated this shape pair LAST frame
// (hint, -1 = none) is tested first, with exactly the arithmetic the full walk applies to it; a pair that is still apart leaves
// after one axis instead of walking the axis list up to its first separating one.  sep returns the axis that separated (-1: none).
// The hint only skips work: a stale or foreign hint names some axis of the set (or none) and the verdict is the same.
struct ContactSink {
  RawContact* buf; int n; int cap; bool overflow; int hint = -1; int sep = -1;
  EPD void emit(bool flip, int64_t fk, V3 ni, V3 pi, V3 pj) {      // Contact.flip, Contact.elm:36-43
    if (n >= cap) { overflow = true; return; }
    RawContact c;
    c.fk = fk;
    if (flip) { c.ni = vneg(ni); c.pi = pj; c.pj = pi; } else { c.ni = ni; c.pi = pi; c.pj = pj; }
    buf[n++] = c;
  }
};

// ContactId.elm:94-287
EPD int64_t feat(int64_t tag, int64_t a, int64_t b, int64_t c, int64_t d) { return (((tag * 2048 + a) * 2048 + b) * 2048 + c) * 2048 + d; }
EPD int on_face(int faceId) { return faceId + 2; }
constexpr int kOnEdge = 1, kOnVertex = 2, kOnNone = 0;

// World-placed convex = placed f64 block (same layout as the template) + the template's index block.
struct Cvx {
  const double* f; const int* ib; int nv, ng, ne; bool box;
  EPD V3 pos() const { return ld3(f); }
  EPD V3 ax() const { return ld3(f + 3); }
  EPD V3 ay() const { return ld3(f + 6); }
  EPD V3 az() const { return ld3(f + 9); }
  EPD V3 he() const { return ld3(f + 12); }
  EPD V3 vert(int i) const { return ld3(f + kCvxF64Header + 3 * i); }
  // world group k (0-based) = template group ng-1-k: placeFaces reverses (Convex.elm:265-287)
  EPD const double* gf(int k) const { return f + kCvxF64Header + 3 * nv + kGroupF64 * (ng - 1 - k); }
  EPD const int* gi(int k) const { return ib + kCvxI32Header + kGroupI32 * (ng - 1 - k); }
  EPD const int* edges(int e, int& n) const { const int* ie = ib + kCvxI32Header + kGroupI32 * ng + kEdgeI32 * e; n = ie[0]; return ib + ie[1]; }
};
EPD Cvx make_cvx(const double* f, const int* ib) { Cvx c; c.f = f; c.ib = ib; c.box = ib[0] != 0; c.nv = ib[1]; c.ng = ib[2]; c.ne = ib[3]; return c; }

struct Face { V3 n; const int* idx; int cnt; };
EPD Face group_face(const Cvx& c, int k, bool partner) {
  const double* g = c.gf(k); const int* gi = c.gi(k);
  Face f;
  if (!partner) { f.n = ld3(g); f.cnt = gi[1]; f.idx = c.ib + gi[2]; }
  else { f.n = ld3(g + 4); f.cnt = gi[3]; f.idx = c.ib + gi[4]; }
  return f;
}
// flat world-order face id (1-based) -> face; returns false past the end
EPD bool flat_face(const Cvx& c, int flat0, Face& f) {
  int id = 0;
  for (int k = 0; k < c.ng; k++) {
    bool two = c.gi(k)[0] != 0;
    if (flat0 == id) { f = group_face(c, k, false); return true; }
    if (two && flat0 == id + 1) { f = group_face(c, k, true); return true; }
    id += two ? 2 : 1;
  }
  return false;
}

struct TagPt { int tag; V3 p; };

// ---- Manifold.elm:25-169 : in-place on a forward list `pts[0..n)`; returns the new count ---------
EPD int farthest_from(V3 a, const TagPt* p, int n) {
  int best = 0; double bs = vdist2(a, p[0].p);
  for (int i = 1; i < n; i++) { double s = vdist2(a, p[i].p); if (s - bs > 0) { best = i; bs = s; } }
  return best;
}
EPD int manifold_reduce(V3 normal, double planeConstant, TagPt* pts, int n, TagPt* tmp) {
  double margin = planeConstant - kBreaking;
  int m = 0;
  for (int i = 0; i < n; i++) if (vdot(normal, pts[i].p) + margin < 0) tmp[m++] = pts[i];
  // withinMargin conses => candidates are in REVERSE scan order
  for (int i = 0; i < m; i++) pts[i] = tmp[m - 1 - i];
  if (m < 5) return m;
  TagPt bestA = pts[0], bestB = pts[0]; double bestD = -1;
  for (int i = 0; i < m; i++) {
    int q = farthest_from(pts[i].p, pts, m);
    double d = vdist2(pts[i].p, pts[q].p);
    if (d - bestD - 1.0e-8 > 0) { bestA = pts[i]; bestB = pts[q]; bestD = d; }
  }
  int ci = 0; double cs = emin(vdist2(bestA.p, pts[0].p), vdist2(bestB.p, pts[0].p));
  for (int i = 1; i < m; i++) { double s = emin(vdist2(bestA.p, pts[i].p), vdist2(bestB.p, pts[i].p)); if (s - cs > 0) { ci = i; cs = s; } }
  TagPt c = pts[ci];
  int ei = 0; double es = emin(vdist2(bestA.p, pts[0].p), emin(vdist2(bestB.p, pts[0].p), vdist2(c.p, pts[0].p)));
  for (int i = 1; i < m; i++) {
    double s = emin(vdist2(bestA.p, pts[i].p), emin(vdist2(bestB.p, pts[i].p), vdist2(c.p, pts[i].p)));
    if (s - es > 0) { ei = i; es = s; }
  }
  TagPt e = pts[ei];
  pts[0] = bestA; pts[1] = bestB; pts[2] = c; pts[3] = e;
  return 4;
}

// ---- primitives -----------------------------------------------------------------------------------
struct Sph { V3 pos; double r; };
struct Cap { V3 pos, axis; double r, h; };
struct Pln { V3 pos, n; };

// shared by PlaneSphere.elm:10-42, PlaneParticle.elm:10-33, PlaneCapsule.elm:39-61
EPD void plane_point(bool flip, int64_t fk, const Pln& pl, V3 vertex, ContactSink& out) {
  double d = ((vertex.x - pl.pos.x) * pl.n.x) + ((vertex.y - pl.pos.y) * pl.n.y) + ((vertex.z - pl.pos.z) * pl.n.z);
  if (d - kBreaking < 0) out.emit(flip, fk, pl.n, mk(vertex.x - d * pl.n.x, vertex.y - d * pl.n.y, vertex.z - d * pl.n.z), vertex);
}
EPD void plane_sphere(bool flip, const Pln& pl, const Sph& s, ContactSink& out) {
  plane_point(flip, 0, pl, mk(s.pos.x - s.r * pl.n.x, s.pos.y - s.r * pl.n.y, s.pos.z - s.r * pl.n.z), out);
}
EPD void plane_capsule(bool flip, const Pln& pl, const Cap& c, ContactSink& out) {
  V3 ep1 = mk(c.pos.x - c.h * c.axis.x, c.pos.y - c.h * c.axis.y, c.pos.z - c.h * c.axis.z);
  V3 ep2 = mk(c.pos.x + c.h * c.axis.x, c.pos.y + c.h * c.axis.y, c.pos.z + c.h * c.axis.z);
  plane_point(flip, feat(4, 1, 0, 0, 0), pl, mk(ep1.x - c.r * pl.n.x, ep1.y - c.r * pl.n.y, ep1.z - c.r * pl.n.z), out);
  plane_point(flip, feat(4, 2, 0, 0, 0), pl, mk(ep2.x - c.r * pl.n.x, ep2.y - c.r * pl.n.y, ep2.z - c.r * pl.n.z), out);
}
// SphereSphere.elm:9-42
EPD void sphere_sphere(const Sph& a, const Sph& b, ContactSink& out) {
  double dist = vdist(b.pos, a.pos) - a.r - b.r;
  V3 n = vdirection(b.pos, a.pos);
  if (dist > 0) return;
  out.emit(false, 0, n, vadd(a.pos, vscale(a.r - dist, n)), vadd(b.pos, vscale(-b.r, n)));
}
// SphereParticle.elm:9-29
EPD void sphere_particle(bool flip, const Sph& s, V3 pp, ContactSink& out) {
  double dist = vdist(pp, s.pos) - s.r;
  V3 n = vdirection(pp, s.pos);
  if (dist > 0) return;
  out.emit(flip, 0, n, vadd(s.pos, vscale(s.r - dist, n)), pp);
}
// CapsuleSphere.elm:10-39 / CapsuleParticle.elm:9-38
EPD void capsule_sphere(bool flip, const Cap& c, const Sph& s, ContactSink& out) {
  double t = emax(-c.h, emin(c.h, vdot(vsub(s.pos, c.pos), c.axis)));
  V3 cp = vadd(c.pos, vscale(t, c.axis));
  double dist = vdist(s.pos, cp) - c.r - s.r;
  V3 n = vdirection(s.pos, cp);
  if (dist > 0) return;
  out.emit(flip, 0, n, vadd(cp, vscale(c.r, n)), vsub(s.pos, vscale(s.r, n)));
}
EPD void capsule_particle(bool flip, const Cap& c, V3 pp, ContactSink& out) {
  double t = emax(-c.h, emin(c.h, vdot(vsub(pp, c.pos), c.axis)));
  V3 cp = vadd(c.pos, vscale(t, c.axis));
  double dist = vdist(pp, cp) - c.r;
  V3 n = vdirection(pp, cp);
  if (dist > 0) return;
  out.emit(flip, 0, n, vadd(cp, vscale(c.r, n)), pp);
}
// CapsuleCapsule.elm:10-97
EPD void capsule_capsule(const Cap& a, const Cap& b, ContactSink& out) {
  V3 r = vsub(a.pos, b.pos);
  double bb = vdot(a.axis, b.axis), c = vdot(a.axis, r), f = vdot(b.axis, r);
  double denom = 1 - bb * bb, s, t;
  if (denom - kParallelTol < 0) {
    t = emax(-b.h, emin(b.h, f));
    s = emax(-a.h, emin(a.h, bb * t - c));
  } else {
    double s0 = emax(-a.h, emin(a.h, (bb * f - c) / denom));
    t = emax(-b.h, emin(b.h, bb * s0 + f));
    s = emax(-a.h, emin(a.h, bb * t - c));
  }
  V3 p1 = vadd(a.pos, vscale(s, a.axis)), p2 = vadd(b.pos, vscale(t, b.axis));
  double dist = vdist(p2, p1) - a.r - b.r;
  V3 n = vdirection(p2, p1);
  if (dist > 0) return;
  out.emit(false, 0, n, vadd(p1, vscale(a.r, n)), vsub(p2, vscale(b.r, n)));
}

// Convex.elm:217-234 box corner k (list order +++,++-,+-+,+--,-++,-+-,--+,---)
EPD V3 box_corner(V3 c, V3 ax, V3 ay, V3 az, V3 he, int k) {
  double sx = (k & 4) ? -1.0 : 1.0, sy = (k & 2) ? -1.0 : 1.0, sz = (k & 1) ? -1.0 : 1.0;
  return mk(c.x + sx * he.x * ax.x + sy * he.y * ay.x + sz * he.z * az.x,
            c.y + sx * he.x * ax.y + sy * he.y * ay.y + sz * he.z * az.y,
            c.z + sx * he.x * ax.z + sy * he.y * ay.z + sz * he.z * az.z);
}
// i-th vertex of Convex.convexVertices (Convex.elm:205-214): box corners, or buffer keys DESCENDING
EPD int cvx_nverts(const Cvx& c) { return c.box ? 8 : c.nv; }
EPD V3 cvx_vertex(const Cvx& c, int i) { return c.box ? box_corner(c.pos(), c.ax(), c.ay(), c.az(), c.he(), i) : c.vert(c.nv - 1 - i); }

// ---- PlaneConvex.elm:12-117 -----------------------------------------------------------------------
EPD void plane_convex(bool flip, const Pln& pl, const Cvx& c, ContactSink& out, TagPt* pa, TagPt* pb, bool& polyOverflow) {
  if (c.box) {
    V3 ctr = c.pos(), ax = c.ax(), ay = c.ay(), az = c.az(), he = c.he();
    for (int k = 0; k < 8; k++) plane_point(flip, feat(3, k + 1, 0, 0, 0), pl, box_corner(ctr, ax, ay, az, he, k), out);
  } else {
    // Manifold.reduce over (i+1, v_i), v in descending-key order; only in-margin points are kept
    double pc = -(vdot(pl.n, pl.pos));
    double margin = pc - kBreaking;
    int m = 0;
    for (int i = 0; i < c.nv; i++) {
      V3 v = c.vert(c.nv - 1 - i);
      if (vdot(pl.n, v) + margin < 0) { if (m >= kMaxPoly) { polyOverflow = true; break; } pa[m].tag = i + 1; pa[m].p = v; m++; }
    }
    m = manifold_reduce(pl.n, pc, pa, m, pb);
    for (int i = 0; i < m; i++) {           // emitPlaneContacts, PlaneConvex.elm:84-117 (no threshold re-test)
      V3 v = pa[i].p;
      double d = ((v.x - pl.pos.x) * pl.n.x) + ((v.y - pl.pos.y) * pl.n.y) + ((v.z - pl.pos.z) * pl.n.z);
      out.emit(flip, feat(3, pa[i].tag, 0, 0, 0), pl.n, mk(v.x - d * pl.n.x, v.y - d * pl.n.y, v.z - d * pl.n.z), v);
    }
  }
}

// ---- ParticleConvex.elm:11-80 ---------------------------------------------------------------------
EPD void particle_convex(bool flip, V3 pp, const Cvx& c, ContactSink& out) {
  if (c.ng == 0) return;
  double bestDepth = kMaxNumber; V3 bestN = mk(0, 0, 0);
  for (int k = 0; k < c.ng; k++) {
    bool two = c.gi(k)[0] != 0;
    for (int s = 0; s < (two ? 2 : 1); s++) {
      Face f = group_face(c, k, s == 1);
      V3 point = f.cnt > 0 ? c.vert(f.idx[0]) : mk(0, 0, 0);
      double d = vdot(f.n, vsub(point, pp));
      if (d >= 0) { if (d - bestDepth < 0) { bestDepth = d; bestN = f.n; } }
      else return;
    }
  }
  if (bestDepth - kMaxNumber < 0) out.emit(flip, 0, vneg(bestN), pp, vadd(pp, vscale(bestDepth, bestN)));
}

// ---- SphereConvex.elm:15-250 ----------------------------------------------------------------------
EPD void sphere_emit(bool flip, int64_t fk, V3 center, V3 cp, double pen, ContactSink& out) {
  V3 n = vdirection(cp, center);
  out.emit(flip, fk, n, mk(cp.x + pen * n.x, cp.y + pen * n.y, cp.z + pen * n.z), cp);
}
EPD double sphere_face_distance(const Cvx& c, const Face& f, V3 center) {
  if (f.cnt == 0) return -1;
  V3 first = c.vert(f.idx[0]);
  return f.n.x * (center.x - first.x) + f.n.y * (center.y - first.y) + f.n.z * (center.z - first.z);
}
EPD double sphere_edge_side(const Face& f, V3 v1, V3 v2, V3 center) {
  double ex = v2.x - v1.x, ey = v2.y - v1.y, ez = v2.z - v1.z;
  double cnX = f.n.y * ez - f.n.z * ey, cnY = f.n.z * ex - f.n.x * ez, cnZ = f.n.x * ey - f.n.y * ex;
  return cnX * (v1.x - center.x) + cnY * (v1.y - center.y) + cnZ * (v1.z - center.z);
}
EPD void sphere_convex(bool flip, const Sph& s, const Cvx& c, ContactSink& out) {
  if (c.ng == 0) return;
  V3 center = s.pos; double radius = s.r;
  // pass 1 (walkFaces): first face whose polygon contains the centre's projection wins
  int faceId = 0;
  for (int k = 0; k < c.ng; k++) {
    bool two = c.gi(k)[0] != 0;
    for (int sd = 0; sd < (two ? 2 : 1); sd++) {
      faceId++;
      Face f = group_face(c, k, sd == 1);
      double fd = sphere_face_distance(c, f, center);
      if (fd > 0 && fd - radius < 0) {
        bool anyOutside = false;
        if (f.cnt >= 2)
          for (int e = 0; e < f.cnt; e++) {
            V3 v1 = c.vert(f.idx[e]), v2 = c.vert(f.idx[e + 1 < f.cnt ? e + 1 : 0]);
            if (sphere_edge_side(f, v1, v2, center) > 0) { anyOutside = true; break; }
          }
        if (!anyOutside) {
          sphere_emit(flip, feat(5, on_face(faceId), 0, 0, 0), center, mk(center.x - fd * f.n.x, center.y - fd * f.n.y, center.z - fd * f.n.z), radius - fd, out);
          return;
        }
      }
    }
  }
  // pass 2 (walkBoundaries): the candidate-edge stack is popped newest first = faces and edges in
  // reverse order, recomputing the same classification instead of storing the list
  V3 bestP = mk(0, 0, 0); double bestD2 = radius * radius;
  for (int k = c.ng - 1; k >= 0; k--) {
    bool two = c.gi(k)[0] != 0;
    for (int sd = (two ? 1 : 0); sd >= 0; sd--) {
      Face f = group_face(c, k, sd == 1);
      double fd = sphere_face_distance(c, f, center);
      if (!(fd > 0 && fd - radius < 0) || f.cnt < 2) continue;
      for (int e = f.cnt - 1; e >= 0; e--) {
        V3 pv = c.vert(f.idx[e]), v = c.vert(f.idx[e + 1 < f.cnt ? e + 1 : 0]);
        if (!(sphere_edge_side(f, pv, v, center) > 0)) continue;
        double ex = v.x - pv.x, ey = v.y - pv.y, ez = v.z - pv.z;
        double len2 = ex * ex + ey * ey + ez * ez;
        double off = (center.x - pv.x) * ex + (center.y - pv.y) * ey + (center.z - pv.z) * ez;
        if (off < 0) {
          double d2 = vdist2(pv, center);
          if (d2 - bestD2 < 0) { bestP = pv; bestD2 = d2; }
        } else if (off - len2 > 0) {
          double d2 = vdist2(v, center);
          if (d2 - bestD2 < 0) { bestP = v; bestD2 = d2; }
        } else {
          double fr = off / len2;
          V3 cp = mk(pv.x + fr * ex, pv.y + fr * ey, pv.z + fr * ez);
          double d2 = vdist2(cp, center);
          if (d2 - bestD2 < 0) { sphere_emit(flip, feat(5, kOnEdge, 0, 0, 0), center, cp, radius - sqrt(d2), out); return; }
        }
      }
    }
  }
  if (bestD2 - radius * radius < 0) sphere_emit(flip, feat(5, kOnVertex, 0, 0, 0), center, bestP, radius - sqrt(bestD2), out);
}

// ---- fp32 screening of SAT axes ------------------------------------------------------------------------------------
// The SAT needs the exact overlap only of the axes that can still matter.  If T is any upper bound of the smallest TRUE
// overlap of a set of axes, an axis whose true overlap is > T neither wins (`dist - best < 0` is a strict improvement,
// first wins ties: ConvexConvex.elm:447-470, 540-573) nor changes the verdict "separated" (the smallest one decides it as
// well: every separating axis yields the same empty result).  Bounds come from fp32 projections of an fp32 copy of the
// hull's vertices RELATIVE to an origin o near the pair (v.n = (v - o).n + o.n; differences of two hulls' projections do
// not depend on o): pass 1 walks ALL axes in fp32 and keeps a bit mask of the survivors, pass 2 evaluates the survivors in
// fp64 exactly as the reference does, in the reference's order.  The lanes of a warp walk pass 1 in lock step and pass 2
// over their OWN survivor lists, so a warp runs max-over-lanes(survivors) exact projections instead of all of them.
// The copy lives in local memory (interleaved per thread: the lanes read it coalesced), 16 B per vertex so that a vertex is ONE
// load, and pass 1 projects two axes per walk over the vertices (k_narrowphase 0.82 -> 0.78 ms at 8192 worlds, 0.285 -> 0.26 at
// 1024).  Normalising the edge axes in fp32 for the screen (rsqrtf, bound widened accordingly) was measured SLOWER (0.77 -> 0.82).
// Error of one projected value against the reference's fp64 value: |fl32(v - o) - (v - o)| <= 2^-24 D per component
// (D = largest |component| of any v - o), |fl32(n) - n| <= 2^-24, three products and two sums in fp32 on terms <= D
// =>  <= 15 * 2^-24 * D < 9e-7 D per value, < 2e-6 D per overlap (a difference of two values).  kScrE doubles that.
// Screening only skips work: every value that is USED is the exact fp64 one.
constexpr int kScrVerts = 32;
constexpr int kScrEdges = 24;     // edge directions per hull the screened SAT takes (a 12-gon cone has 18); axes go through it 64 at a time
constexpr double kScrE = 4.0e-6;
struct Scr { float4 r[kScrVerts]; int nv; };      // (x, y, z, -) per vertex: one 16 B local load each
EPD void scr_build(Scr& s, const Cvx& c, V3 o, double& D) {
  for (int i = 0; i < c.nv; i++) {
    V3 v = c.vert(i);
    double dx = v.x - o.x, dy = v.y - o.y, dz = v.z - o.z;
    s.r[i] = make_float4((float)dx, (float)dy, (float)dz, 0.0f);
    D = emax(D, emax(eabs(dx), emax(eabs(dy), eabs(dz))));
  }
  s.nv = c.nv;
}
EPD float scr_dot(float4 v, float nx, float ny, float nz) { return fmaf(v.z, nz, fmaf(v.y, ny, v.x * nx)); }
EPD void scr_project(const Scr& s, V3 n, float& lo, float& hi) {
  const float nx = (float)n.x, ny = (float)n.y, nz = (float)n.z;
  float lo0 = 3.0e38f, lo1 = 3.0e38f, hi0 = -3.0e38f, hi1 = -3.0e38f;
  int i = 0;
  for (; i + 1 < s.nv; i += 2) {
    const float a = scr_dot(s.r[i], nx, ny, nz), b = scr_dot(s.r[i + 1], nx, ny, nz);
    lo0 = fminf(lo0, a); hi0 = fmaxf(hi0, a); lo1 = fminf(lo1, b); hi1 = fmaxf(hi1, b);
  }
  if (i < s.nv) { const float a = scr_dot(s.r[i], nx, ny, nz); lo0 = fminf(lo0, a); hi0 = fmaxf(hi0, a); }
  lo = fminf(lo0, lo1); hi = fmaxf(hi0, hi1);
}
// two axes in one walk over the vertices (each vertex is loaded once): the same values as two scr_project calls
EPD void scr_project2(const Scr& s, V3 na, V3 nb, float& loa, float& hia, float& lob, float& hib) {
  const float ax = (float)na.x, ay = (float)na.y, az = (float)na.z, bx = (float)nb.x, by = (float)nb.y, bz = (float)nb.z;
  float la0 = 3.0e38f, la1 = 3.0e38f, ha0 = -3.0e38f, ha1 = -3.0e38f, lb0 = 3.0e38f, lb1 = 3.0e38f, hb0 = -3.0e38f, hb1 = -3.0e38f;
  int i = 0;
  for (; i + 1 < s.nv; i += 2) {
    const float4 v = s.r[i], w = s.r[i + 1];
    const float a = scr_dot(v, ax, ay, az), b = scr_dot(w, ax, ay, az), c = scr_dot(v, bx, by, bz), e = scr_dot(w, bx, by, bz);
    la0 = fminf(la0, a); ha0 = fmaxf(ha0, a); la1 = fminf(la1, b); ha1 = fmaxf(ha1, b);
    lb0 = fminf(lb0, c); hb0 = fmaxf(hb0, c); lb1 = fminf(lb1, e); hb1 = fmaxf(hb1, e);
  }
  if (i < s.nv) {
    const float4 v = s.r[i];
    const float a = scr_dot(v, ax, ay, az), c = scr_dot(v, bx, by, bz);
    la0 = fminf(la0, a); ha0 = fmaxf(ha0, a); lb0 = fminf(lb0, c); hb0 = fmaxf(hb0, c);
  }
  loa = fminf(la0, la1); hia = fmaxf(ha0, ha1); lob = fminf(lb0, lb1); hib = fmaxf(hb0, hb1);
}

// ---- ConvexConvex.elm -----------------------------------------------------------------------------
// projectConvex 596-610 / project 689-710
EPD void project_convex(V3 axis, const Cvx& c, double& mn, double& mx) {
  if (c.box) {
    double cc = vdot(axis, c.pos());
    double e = eabs(vdot(axis, c.ax())) * c.f[12] + eabs(vdot(axis, c.ay())) * c.f[13] + eabs(vdot(axis, c.az())) * c.f[14];
    mn = cc - e; mx = cc + e;
  } else {
    mn = kMaxNumber; mx = -kMaxNumber;
#pragma unroll 4
    for (int i = c.nv - 1; i >= 0; i--) {
      V3 v = c.vert(i);
      double val = v.x * axis.x + v.y * axis.y + v.z * axis.z;
      mn = (mn - val > 0) ? val : mn;
      mx = (mx - val > 0) ? mx : val;
    }
  }
}
// overlap 646-662
EPD bool overlap(double min1, double max1, double min2, double max2, double& depth) {
  double d1 = max1 - min2, d2 = max2 - min1;
  if (d1 + kBreaking < 0 || d2 + kBreaking < 0) return false;
  depth = (d1 - d2 > 0) ? d2 : d1;
  return true;
}
// faceExtent 669-680 : owning convex along its own group normal
EPD void face_extent(V3 axis, const Cvx& c, int k, double& mn, double& mx) {
  if (c.gi(k)[0] != 0) { const double* g = c.gf(k); double pd = vdot(axis, c.pos()); mx = g[3] + pd; mn = g[7] + pd; }
  else {
    mn = kMaxNumber; mx = -kMaxNumber;
    int n = cvx_nverts(c);
    for (int i = 0; i < n; i++) {
      V3 v = cvx_vertex(c, i);
      double val = v.x * axis.x + v.y * axis.y + v.z * axis.z;
      mn = (mn - val > 0) ? val : mn;
      mx = (mx - val > 0) ? mx : val;
    }
  }
}
EPD V3 orient_axis(const Cvx& c1, const Cvx& c2, V3 axis) { return vdot(vsub(c2.pos(), c1.pos()), axis) > 0 ? vneg(axis) : axis; }

// bestFace 276-324 : returns the 1-based flat face id (or -1)
EPD int best_face(const Cvx& c, V3 sep, Face& out) {
  int faceId = 1, bestId = -1; double bestDist = kMaxNumber;
  out.n = mk(0, 0, 0); out.idx = nullptr; out.cnt = 0;
  for (int k = 0; k < c.ng; k++) {
    const double* g = c.gf(k);
    if (c.gi(k)[0] != 0) {
      double pd = vdot(ld3(g), sep), qd = -pd;
      int id1 = bestId; double d1 = bestDist; bool took1 = false;
      if (bestDist - pd > 0) { id1 = faceId; d1 = pd; took1 = true; }
      if (d1 - qd > 0) { bestId = faceId + 1; out = group_face(c, k, true); bestDist = qd; }
      else if (took1) { bestId = id1; out = group_face(c, k, false); bestDist = d1; }
      faceId += 2;
    } else {
      double d = vdot(ld3(g), sep);
      if (bestDist - d > 0) { bestId = faceId; out = group_face(c, k, false); bestDist = d; }
      faceId += 1;
    }
  }
  return bestId;
}
// pickSupportEdge 175-201
EPD int pick_support_edge(V3 dir, const int* e, int n, const Cvx& c, V3& p, V3& q) {
  int bestIdx = 0, idx = 1; double bestDot = -kMaxNumber;
  p = mk(0, 0, 0); q = mk(0, 0, 0);
  for (int k = 0; k + 1 < n; k += 2, idx++) {
    V3 v1 = c.vert(e[k]), v2 = c.vert(e[k + 1]);
    double md = dir.x * (v1.x + v2.x) + dir.y * (v1.y + v2.y) + dir.z * (v1.z + v2.z);
    if (md - bestDot > 0) { bestIdx = idx; p = v1; q = v2; bestDot = md; }
  }
  return bestIdx;
}
// clipFaceAgainstPlaneAdd 349-395 over a forward polygon src[0..n) -> dst (forward list), returns count
EPD int clip_polygon(V3 pn, double pc, const TagPt* src, int n, TagPt* dst, bool& ovf) {
  if (n < 2) return 0;
  // result is built by cons while walking edges (p_i, p_i+1): final forward list = reverse of pushes
  int m = 0;
  for (int i = 0; i < n; i++) {
    const TagPt& prev = src[i];
    const TagPt& next = src[i + 1 < n ? i + 1 : 0];
    double dp = vdot(pn, prev.p) + pc, dn = vdot(pn, next.p) + pc;
    bool pin = dp < 0, nin = dn < 0;
    if (!pin && !nin) continue;
    TagPt cr;
    if (pin != nin) { double t = dp / (dp - dn); cr.tag = t < 0.5 ? prev.tag : next.tag; cr.p = vlerp(t, prev.p, next.p); }
    if (m + 2 > kMaxPoly) { ovf = true; break; }
    if (pin) dst[m++] = nin ? next : cr;
    else { dst[m++] = cr; dst[m++] = next; }
  }
  for (int i = 0; i < m / 2; i++) { TagPt t = dst[i]; dst[i] = dst[m - 1 - i]; dst[m - 1 - i] = t; }
  return m;
}
struct SatResult { int winIdx, winSide, winK; double dmin; bool edgeBeats; V3 eAxis; int dir1, dir2; };
// findFaceSAT 420-479 + findEdgeSAT 517-580 (threshold pre-divided by edgeBiasFactor 1.05), every axis in fp64; false = separated
// Axis ids of a hull pair (ContactSink::hint / sep): face axes 0 .. ng1 + ng2 - 1 (convex1's groups, then convex2's), edge axis
// (a, b) = ng1 + ng2 + a * ne2 + b.  sat_axis_separates evaluates ONE axis as the walks below do: true = it separates.
EPD bool sat_axis_separates(const Cvx& c1, const Cvx& c2, int id) {
  double dist;
  if (id < 0) return false;
  if (id < c1.ng + c2.ng) {
    const bool first = id < c1.ng;
    const Cvx& own = first ? c1 : c2;
    const Cvx& other = first ? c2 : c1;
    const int k = first ? id : id - c1.ng;
    const V3 axis = ld3(own.gf(k));
    double omn, omx, pmn, pmx;
    face_extent(axis, own, k, omn, omx);
    project_convex(axis, other, pmn, pmx);
    return !(first ? overlap(omn, omx, pmn, pmx, dist) : overlap(pmn, pmx, omn, omx, dist));
  }
  const int e = id - c1.ng - c2.ng;
  if (c2.ne <= 0 || e >= c1.ne * c2.ne) return false;
  const int a = e / c2.ne, b = e - a * c2.ne;
  int n1, n2; const int* e1 = c1.edges(a, n1); const int* e2 = c2.edges(b, n2);
  if (n1 < 2 || n2 < 2) return false;
  const V3 cr = vcross(vdirection(c1.vert(e1[0]), c1.vert(e1[1])), vdirection(c2.vert(e2[0]), c2.vert(e2[1])));
  if (valmost0(cr)) return false;
  const V3 n = vnormalize(cr);
  double mn1, mx1, mn2, mx2;
  project_convex(n, c1, mn1, mx1);
  project_convex(n, c2, mn2, mx2);
  return !overlap(mn1, mx1, mn2, mx2, dist);
}
EPD bool convex_sat(const Cvx& c1, const Cvx& c2, SatResult& R, int& sep) {
  // face SAT: convex1's groups then convex2's, strict improvement => first wins ties
  R.winIdx = -1; R.winSide = 1; R.winK = 0; R.dmin = kMaxNumber;
  for (int side = 1; side <= 2; side++) {
    const Cvx& own = side == 1 ? c1 : c2;
    const Cvx& other = side == 1 ? c2 : c1;
    int nextIdx = 1;
    for (int k = 0; k < own.ng; k++) {
      V3 axis = ld3(own.gf(k));
      double omn, omx, pmn, pmx, dist;
      face_extent(axis, own, k, omn, omx);
      project_convex(axis, other, pmn, pmx);
      bool ok = side == 1 ? overlap(omn, omx, pmn, pmx, dist) : overlap(pmn, pmx, omn, omx, dist);
      if (!ok) { sep = (side == 1 ? 0 : c1.ng) + k; return false; }
      if (dist - R.dmin < 0) { R.winIdx = nextIdx; R.winSide = side; R.winK = k; R.dmin = dist; }
      nextIdx += own.gi(k)[0] != 0 ? 2 : 1;
    }
  }
  if (R.winIdx == -1) return false;
  R.edgeBeats = false; R.eAxis = mk(0, 0, 0); R.dir1 = 0; R.dir2 = 0;
  double emin_ = R.dmin / 1.05;
  for (int a = 0; a < c1.ne; a++) {
    int n1; const int* e1 = c1.edges(a, n1);
    if (n1 < 2) continue;
    V3 d1 = vdirection(c1.vert(e1[0]), c1.vert(e1[1]));
    for (int b = 0; b < c2.ne; b++) {
      int n2; const int* e2 = c2.edges(b, n2);
      if (n2 < 2) continue;
      V3 d2 = vdirection(c2.vert(e2[0]), c2.vert(e2[1]));
      V3 cr = vcross(d1, d2);
      if (valmost0(cr)) continue;
      V3 n = vnormalize(cr);
      double mn1, mx1, mn2, mx2, dist;
      project_convex(n, c1, mn1, mx1);
      project_convex(n, c2, mn2, mx2);
      if (!overlap(mn1, mx1, mn2, mx2, dist)) { sep = c1.ng + c2.ng + a * c2.ne + b; return false; }       // EdgeSeparates
      if (dist - emin_ < 0) { R.edgeBeats = true; R.eAxis = n; R.dir1 = a + 1; R.dir2 = b + 1; emin_ = dist; }
    }
  }
  return true;
}
// The same verdict and the same winners with the axes screened in fp32 first (see above).  For two non-box hulls of at most
// kScrVerts vertices, 32 face groups and kScrEdges edge directions each.
EPD bool convex_sat_screened(const Cvx& c1, const Cvx& c2, SatResult& R, int& sep) {
  const V3 so = c1.pos();
  Scr s1, s2; double sD = 0;
  scr_build(s1, c1, so, sD); scr_build(s2, c2, so, sD);
  const double sE = kScrE * sD + 1.0e-15 * (eabs(so.x) + eabs(so.y) + eabs(so.z)) + 1.0e-12;
  // ---- face axes, pass 1: axis j = group k of convex1 (j = k) or of convex2 (j = ng1 + k)
  unsigned fmask = 0; double thr = kMaxNumber;
  {
    // the axes are walked two at a time (both of the same hull: they project the same fp32 copy); the running threshold is
    // applied in ascending j afterwards, so the mask is the one the one-by-one walk produces
    auto face_est = [&](const Cvx& own, int k, V3 axis, float lo, float hi) -> double {
      double omn, omx;
      face_extent(axis, own, k, omn, omx);
      const double on = vdot(axis, so);
      const double a = omx - ((double)lo + on), b = ((double)hi + on) - omn;
      return (a - b > 0) ? b : a;
    };
    auto apply = [&](int j, double est) {
      if (!(est - sE > thr)) fmask |= 1u << j;
      if (est + sE < thr) thr = est + sE;
    };
    for (int side = 0; side < 2; side++) {
      const Cvx& own = side == 0 ? c1 : c2;
      const Scr& other = side == 0 ? s2 : s1;
      const int j0 = side == 0 ? 0 : c1.ng;
      int k = 0;
      for (; k + 1 < own.ng; k += 2) {
        const V3 xa = ld3(own.gf(k)), xb = ld3(own.gf(k + 1));
        float loa, hia, lob, hib; scr_project2(other, xa, xb, loa, hia, lob, hib);
        apply(j0 + k, face_est(own, k, xa, loa, hia));
        apply(j0 + k + 1, face_est(own, k + 1, xb, lob, hib));
      }
      if (k < own.ng) {
        const V3 xa = ld3(own.gf(k));
        float lo, hi; scr_project(other, xa, lo, hi);
        apply(j0 + k, face_est(own, k, xa, lo, hi));
      }
    }
  }
  // ---- face axes, pass 2: the survivors in fp64, ascending j = the reference's order
  R.winIdx = -1; R.winSide = 1; R.winK = 0; R.dmin = kMaxNumber;
  while (fmask) {
    const int j = __ffs(fmask) - 1; fmask &= fmask - 1;
    const bool first = j < c1.ng;
    const Cvx& own = first ? c1 : c2;
    const Cvx& other = first ? c2 : c1;
    const int k = first ? j : j - c1.ng;
    int idx = 1;
    for (int q = 0; q < k; q++) idx += own.gi(q)[0] != 0 ? 2 : 1;
    const V3 axis = ld3(own.gf(k));
    double omn, omx, pmn, pmx, dist;
    face_extent(axis, own, k, omn, omx);
    project_convex(axis, other, pmn, pmx);
    const bool ok = first ? overlap(omn, omx, pmn, pmx, dist) : overlap(pmn, pmx, omn, omx, dist);
    if (!ok) { sep = j; return false; }
    if (dist - R.dmin < 0) { R.winIdx = idx; R.winSide = first ? 1 : 2; R.winK = k; R.dmin = dist; }
  }
  if (R.winIdx == -1) return false;
  // ---- edge axes, pass 1: axis e = a * ne2 + b
  R.edgeBeats = false; R.eAxis = mk(0, 0, 0); R.dir1 = 0; R.dir2 = 0;
  double emin_ = R.dmin / 1.05;
  double d2s[3 * kScrEdges]; unsigned ok2 = 0;
  for (int b = 0; b < c2.ne; b++) {
    int n2; const int* e2 = c2.edges(b, n2);
    if (n2 < 2) continue;
    const V3 d2 = vdirection(c2.vert(e2[0]), c2.vert(e2[1]));
    d2s[3 * b] = d2.x; d2s[3 * b + 1] = d2.y; d2s[3 * b + 2] = d2.z; ok2 |= 1u << b;
  }
  // The rows of convex1's directions go through the two passes in blocks of at most 64 axes (the survivor mask): pass 2 of a
  // block runs before pass 1 of the next, so the survivors still come in the reference's order, and every threshold in use (the
  // exact running minimum, or a screened estimate + its error bound) is an upper bound of the final minimum as before.
  thr = emin_;
  const int rows = 64 / c2.ne;                           // >= 2 (ne <= kScrEdges)
  for (int a0 = 0; a0 < c1.ne; a0 += rows) {
    const int a1 = min(c1.ne, a0 + rows);
    unsigned long long emask = 0;
    {
      auto edge_est = [&](float lo1, float hi1, float lo2, float hi2) -> double {
        const double x = (double)hi1 - (double)lo2, y = (double)hi2 - (double)lo1;
        return (x - y > 0) ? y : x;
      };
      auto apply = [&](int e, double est) {
        if (!(est - 2 * sE > thr)) emask |= 1ull << e;
        if (est + 2 * sE < thr) thr = est + 2 * sE;
      };
      for (int a = a0; a < a1; a++) {
        int n1; const int* e1 = c1.edges(a, n1);
        if (n1 < 2) continue;
        const V3 d1 = vdirection(c1.vert(e1[0]), c1.vert(e1[1]));
        const int bit0 = (a - a0) * c2.ne;
        // the valid axes of this row two at a time (each vertex of a copy is loaded once per pair); estimates applied in ascending b
        int pb = -1; V3 pn = mk(0, 0, 0);                 // a valid axis waiting for a partner
        for (int b = 0; b < c2.ne; b++) {
          if (!((ok2 >> b) & 1u)) continue;
          const V3 cr = vcross(d1, mk(d2s[3 * b], d2s[3 * b + 1], d2s[3 * b + 2]));
          if (valmost0(cr)) continue;
          const V3 n = vnormalize(cr);
          if (pb < 0) { pb = b; pn = n; continue; }
          float l1a, h1a, l1b, h1b, l2a, h2a, l2b, h2b;
          scr_project2(s1, pn, n, l1a, h1a, l1b, h1b); scr_project2(s2, pn, n, l2a, h2a, l2b, h2b);
          apply(bit0 + pb, edge_est(l1a, h1a, l2a, h2a));
          apply(bit0 + b, edge_est(l1b, h1b, l2b, h2b));
          pb = -1;
        }
        if (pb >= 0) {
          float lo1, hi1, lo2, hi2; scr_project(s1, pn, lo1, hi1); scr_project(s2, pn, lo2, hi2);
          apply(bit0 + pb, edge_est(lo1, hi1, lo2, hi2));
        }
      }
    }
    // ---- edge axes, pass 2
    while (emask) {
      const int bit = __ffsll((long long)emask) - 1; emask &= emask - 1;
      const int a = a0 + bit / c2.ne, b = bit % c2.ne;
      int n1; const int* e1 = c1.edges(a, n1);
      const V3 d1 = vdirection(c1.vert(e1[0]), c1.vert(e1[1]));
      const V3 n = vnormalize(vcross(d1, mk(d2s[3 * b], d2s[3 * b + 1], d2s[3 * b + 2])));
      double mn1, mx1, mn2, mx2, dist;
      project_convex(n, c1, mn1, mx1);
      project_convex(n, c2, mn2, mx2);
      if (!overlap(mn1, mx1, mn2, mx2, dist)) { sep = c1.ng + c2.ng + a * c2.ne + b; return false; }       // EdgeSeparates
      if (dist - emin_ < 0) { R.edgeBeats = true; R.eAxis = n; R.dir1 = a + 1; R.dir2 = b + 1; emin_ = dist; }
    }
    if (emin_ < thr) thr = emin_;
  }
  return true;
}
// the contacts of a hull pair from the SAT's winners: addEdgeContact 148-169, dispatchBestFaces 69-114, clipTwoFaces 204-269
EPD void convex_contacts_from_sat(const Cvx& c1, const Cvx& c2, const SatResult& R, ContactSink& out, TagPt* pa, TagPt* pb, bool& polyOverflow) {
  const int winIdx = R.winIdx, winSide = R.winSide, winK = R.winK, dir1 = R.dir1, dir2 = R.dir2;
  const bool edgeBeats = R.edgeBeats; const V3 eAxis = R.eAxis;
  if (edgeBeats) {   // addEdgeContact 148-169
    V3 sep = orient_axis(c1, c2, eAxis), rev = vneg(sep);
    int n1, n2; const int* e1 = c1.edges(dir1 - 1, n1); const int* e2 = c2.edges(dir2 - 1, n2);
    V3 p1, q1, p2, q2;
    int i1 = pick_support_edge(rev, e1, n1, c1, p1, q1);
    int i2 = pick_support_edge(sep, e2, n2, c2, p2, q2);
    V3 pi, pj; closest_segments(p1, q1, p2, q2, pi, pj);
    out.emit(false, feat(2, dir1, i1, dir2, i2), rev, pi, pj);
    return;
  }
  const Cvx& wc = winSide == 1 ? c1 : c2;
  V3 sep = orient_axis(c1, c2, ld3(wc.gf(winK))), rev = vneg(sep);
  Face f1, f2; int id1, id2;
  bool two = wc.gi(winK)[0] != 0;
  if (winSide == 1) {      // pickWinningFace 117-132
    bool partner = two && !(vdot(ld3(wc.gf(winK)), sep) <= 0);
    f1 = group_face(wc, winK, partner); id1 = winIdx + (partner ? 1 : 0);
    id2 = best_face(c2, rev, f2);
  } else {
    id1 = best_face(c1, sep, f1);
    bool partner = two && !(vdot(ld3(wc.gf(winK)), rev) <= 0);
    f2 = group_face(wc, winK, partner); id2 = winIdx + (partner ? 1 : 0);
  }
  if (id1 == -1 || id2 == -1) return;
  // clipTwoFaces: reference face f1 (convex1), incident polygon f2 (convex2), ni = rev
  if (f2.cnt > kMaxPoly) { polyOverflow = true; return; }
  int m = f2.cnt;
  for (int i = 0; i < m; i++) { pa[i].tag = f2.idx[i]; pa[i].p = c2.vert(f2.idx[i]); }
  V3 point = f1.cnt > 0 ? c1.vert(f1.idx[0]) : mk(0, 0, 0);
  double fpc = -(vdot(f1.n, point));
  TagPt* src = pa; TagPt* dst = pb;
  if (f1.cnt >= 2)
    for (int e = 0; e < f1.cnt; e++) {    // clipAgainstAdjacentFaces 327-346
      V3 v1 = c1.vert(f1.idx[e]), v2 = c1.vert(f1.idx[e + 1 < f1.cnt ? e + 1 : 0]);
      V3 edge = vnormalize(vsub(v1, v2));
      V3 pn = vcross(f1.n, edge);
      double pc = -(vdot(v1, pn));
      m = clip_polygon(pn, pc, src, m, dst, polyOverflow);
      TagPt* t = src; src = dst; dst = t;
    }
  m = manifold_reduce(f1.n, fpc, src, m, dst);
  for (int i = 0; i < m; i++) {           // emitManifold 240-269
    V3 v = src[i].p;
    double depth = vdot(f1.n, v) + fpc;
    out.emit(false, feat(1, id1, id2, src[i].tag, 0), rev, mk(v.x - depth * f1.n.x, v.y - depth * f1.n.y, v.z - depth * f1.n.z), v);
  }
}
// addContacts 40-63 with dispatchBestFaces 69-114, clipTwoFaces 204-269
// MAY_SCREEN false: an instance for pairs with a box on either side (never screened): it carries neither the fp32 copies nor
// the screening code
template <bool MAY_SCREEN>
EPD void convex_convex_t(const Cvx& c1, const Cvx& c2, ContactSink& out, TagPt* pa, TagPt* pb, bool& polyOverflow) {
  SatResult R;
  // (a box against another hull is screened too: the fp32 copy of its eight placed corners bounds the exact centre +- extent
  // projection of pass 2 like any other hull's -- the corners and the axes come from the same transform, rounding apart; two boxes
  // have 15 axes and are not worth the copies)
  const bool screened = MAY_SCREEN && !(c1.box && c2.box) && c1.nv <= kScrVerts && c2.nv <= kScrVerts && c1.nv > 0 && c2.nv > 0 &&
                        c1.ng + c2.ng <= 32 && c1.ne <= kScrEdges && c2.ne <= kScrEdges && c1.ne > 0 && c2.ne > 0;
  if (sat_axis_separates(c1, c2, out.hint)) { out.sep = out.hint; return; }      // still apart along last frame's axis
  bool hit;
  if (MAY_SCREEN) hit = screened ? convex_sat_screened(c1, c2, R, out.sep) : convex_sat(c1, c2, R, out.sep);
  else hit = convex_sat(c1, c2, R, out.sep);
  if (!hit) return;
  convex_contacts_from_sat(c1, c2, R, out, pa, pb, polyOverflow);
}
EPD void convex_convex(const Cvx& c1, const Cvx& c2, ContactSink& out, TagPt* pa, TagPt* pb, bool& polyOverflow) { convex_convex_t<true>(c1, c2, out, pa, pb, polyOverflow); }

// ---- the screened hull-hull SAT by a QUAD of lanes ---------------------------------------------------------------------------
// Small batches: the narrowphase is the latency of its heaviest warp -- 16 touching 12-gon cylinder pairs, one per lane, ~0.25 ms.
// Here four lanes (q = lane & 3, all holding the same item) share one pair:
//   pass 1 (fp32 screen): the axis pairs are dealt round-robin over the four lanes; every lane keeps its estimates, the threshold
//          is the minimum over the quad (any upper bound of the smallest true overlap is a valid threshold, see above) and the
//          survivor masks are OR-ed -- the survivors are a superset of the axes that can matter, as in the one-lane walk;
//   pass 2 (exact): the survivors in the reference's order, one at a time, each projection split over the four lanes by vertex
//          (min / max are exact and order-independent, so every lane ends with the very value the one-lane loop computes) and
//          the overlap / winner logic run identically by all four;
//   the winners go to lane 0, which forms the contacts (clipping stays serial).
EPD double quad_min(double v, unsigned qm) {
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) { const double a = __shfl_xor_sync(qm, v, o); v = (v - a > 0) ? a : v; }
  return v;
}
EPD double quad_max(double v, unsigned qm) {
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) { const double a = __shfl_xor_sync(qm, v, o); v = (v - a > 0) ? v : a; }
  return v;
}
// projectConvex 596-610 of a non-box hull, vertices dealt over the quad
EPD void project_convex_quad(V3 axis, const Cvx& c, int q, unsigned qm, double& mn, double& mx) {
  mn = kMaxNumber; mx = -kMaxNumber;
  for (int i = c.nv - 1 - q; i >= 0; i -= 4) {
    V3 v = c.vert(i);
    double val = v.x * axis.x + v.y * axis.y + v.z * axis.z;
    mn = (mn - val > 0) ? val : mn;
    mx = (mx - val > 0) ? mx : val;
  }
  mn = quad_min(mn, qm); mx = quad_max(mx, qm);
}
// sat_axis_separates for two non-box hulls, by the quad (the same verdict in all four lanes)
EPD bool sat_axis_separates_quad(const Cvx& c1, const Cvx& c2, int id, int q, unsigned qm) {
  double dist;
  if (id < 0) return false;
  if (id < c1.ng + c2.ng) {
    const bool first = id < c1.ng;
    const Cvx& own = first ? c1 : c2;
    const Cvx& other = first ? c2 : c1;
    const int k = first ? id : id - c1.ng;
    const V3 axis = ld3(own.gf(k));
    double omn, omx, pmn, pmx;
    face_extent(axis, own, k, omn, omx);
    project_convex_quad(axis, other, q, qm, pmn, pmx);
    return !(first ? overlap(omn, omx, pmn, pmx, dist) : overlap(pmn, pmx, omn, omx, dist));
  }
  const int e = id - c1.ng - c2.ng;
  if (c2.ne <= 0 || e >= c1.ne * c2.ne) return false;
  const int a = e / c2.ne, b = e - a * c2.ne;
  int n1, n2; const int* e1 = c1.edges(a, n1); const int* e2 = c2.edges(b, n2);
  if (n1 < 2 || n2 < 2) return false;
  const V3 cr = vcross(vdirection(c1.vert(e1[0]), c1.vert(e1[1])), vdirection(c2.vert(e2[0]), c2.vert(e2[1])));
  if (valmost0(cr)) return false;
  const V3 n = vnormalize(cr);
  double mn1, mx1, mn2, mx2;
  project_convex_quad(n, c1, q, qm, mn1, mx1);
  project_convex_quad(n, c2, q, qm, mn2, mx2);
  return !overlap(mn1, mx1, mn2, mx2, dist);
}
EPD bool convex_sat_screened_quad(const Cvx& c1, const Cvx& c2, SatResult& R, int& sep, int q, unsigned qm) {
  const V3 so = c1.pos();
  Scr s1, s2; double sD = 0;
  scr_build(s1, c1, so, sD); scr_build(s2, c2, so, sD);      // every lane its own copy (local memory)
  const double sE = kScrE * sD + 1.0e-15 * (eabs(so.x) + eabs(so.y) + eabs(so.z)) + 1.0e-12;
  // ---- face axes, pass 1: pairs of axes of one hull, pair p to lane p & 3
  unsigned fmask = 0;
  {
    double estv[10]; int jv[10]; int cnt = 0; double thr = kMaxNumber;
    auto face_est = [&](const Cvx& own, int k, V3 axis, float lo, float hi) -> double {
      double omn, omx;
      face_extent(axis, own, k, omn, omx);
      const double on = vdot(axis, so);
      const double a = omx - ((double)lo + on), b = ((double)hi + on) - omn;
      return (a - b > 0) ? b : a;
    };
    auto keep = [&](int j, double est) {
      if (cnt < 10) { estv[cnt] = est; jv[cnt] = j; cnt++; } else fmask |= 1u << j;      // cannot overflow (<= 17 pairs); survive rather than drop
      if (est + sE < thr) thr = est + sE;
    };
    int pairNo = 0;
    for (int side = 0; side < 2; side++) {
      const Cvx& own = side == 0 ? c1 : c2;
      const Scr& other = side == 0 ? s2 : s1;
      const int j0 = side == 0 ? 0 : c1.ng;
      for (int k = 0; k < own.ng; k += 2, pairNo++) {
        if ((pairNo & 3) != q) continue;
        const V3 xa = ld3(own.gf(k));
        if (k + 1 < own.ng) {
          const V3 xb = ld3(own.gf(k + 1));
          float loa, hia, lob, hib; scr_project2(other, xa, xb, loa, hia, lob, hib);
          keep(j0 + k, face_est(own, k, xa, loa, hia));
          keep(j0 + k + 1, face_est(own, k + 1, xb, lob, hib));
        } else {
          float lo, hi; scr_project(other, xa, lo, hi);
          keep(j0 + k, face_est(own, k, xa, lo, hi));
        }
      }
    }
    thr = quad_min(thr, qm);
    for (int i = 0; i < cnt; i++) if (!(estv[i] - sE > thr)) fmask |= 1u << jv[i];
    fmask |= __shfl_xor_sync(qm, fmask, 1); fmask |= __shfl_xor_sync(qm, fmask, 2);
  }
  // ---- face axes, pass 2: the survivors in fp64, ascending j = the reference's order
  R.winIdx = -1; R.winSide = 1; R.winK = 0; R.dmin = kMaxNumber;
  while (fmask) {
    const int j = __ffs(fmask) - 1; fmask &= fmask - 1;
    const bool first = j < c1.ng;
    const Cvx& own = first ? c1 : c2;
    const Cvx& other = first ? c2 : c1;
    const int k = first ? j : j - c1.ng;
    int idx = 1;
    for (int x = 0; x < k; x++) idx += own.gi(x)[0] != 0 ? 2 : 1;
    const V3 axis = ld3(own.gf(k));
    double omn, omx, pmn, pmx, dist;
    face_extent(axis, own, k, omn, omx);
    project_convex_quad(axis, other, q, qm, pmn, pmx);
    const bool ok = first ? overlap(omn, omx, pmn, pmx, dist) : overlap(pmn, pmx, omn, omx, dist);
    if (!ok) { sep = j; return false; }
    if (dist - R.dmin < 0) { R.winIdx = idx; R.winSide = first ? 1 : 2; R.winK = k; R.dmin = dist; }
  }
  if (R.winIdx == -1) return false;
  // ---- edge axes, pass 1: axis e = a * ne2 + b, pairs (2p, 2p + 1) to lane p & 3
  R.edgeBeats = false; R.eAxis = mk(0, 0, 0); R.dir1 = 0; R.dir2 = 0;
  double emin_ = R.dmin / 1.05;
  double d1s[3 * kScrEdges], d2s[3 * kScrEdges]; unsigned ok1 = 0, ok2 = 0;
  for (int a = 0; a < c1.ne; a++) {
    int n1; const int* e1 = c1.edges(a, n1);
    if (n1 < 2) continue;
    const V3 d1 = vdirection(c1.vert(e1[0]), c1.vert(e1[1]));
    d1s[3 * a] = d1.x; d1s[3 * a + 1] = d1.y; d1s[3 * a + 2] = d1.z; ok1 |= 1u << a;
  }
  for (int b = 0; b < c2.ne; b++) {
    int n2; const int* e2 = c2.edges(b, n2);
    if (n2 < 2) continue;
    const V3 d2 = vdirection(c2.vert(e2[0]), c2.vert(e2[1]));
    d2s[3 * b] = d2.x; d2s[3 * b + 1] = d2.y; d2s[3 * b + 2] = d2.z; ok2 |= 1u << b;
  }
  // 64 axes (one survivor mask) at a time, pass 2 of a block before pass 1 of the next: see convex_sat_screened
  const int ne = c1.ne * c2.ne;
  for (int base = 0; base < ne; base += 64) {
  const int lim = min(ne, base + 64);
  unsigned long long emask = 0;
  {
    double estv[16]; int ev[16]; int cnt = 0; double thr = emin_;
    auto edge_est = [&](float lo1, float hi1, float lo2, float hi2) -> double {
      const double x = (double)hi1 - (double)lo2, y = (double)hi2 - (double)lo1;
      return (x - y > 0) ? y : x;
    };
    auto keep = [&](int e, double est) {
      if (cnt < 16) { estv[cnt] = est; ev[cnt] = e; cnt++; } else emask |= 1ull << e;     // cannot overflow (<= 32 pairs)
      if (est + 2 * sE < thr) thr = est + 2 * sE;
    };
    auto axis_of = [&](int e, V3& n) -> bool {
      const int a = e / c2.ne, b = e - a * c2.ne;
      if (!((ok1 >> a) & 1u) || !((ok2 >> b) & 1u)) return false;
      const V3 cr = vcross(mk(d1s[3 * a], d1s[3 * a + 1], d1s[3 * a + 2]), mk(d2s[3 * b], d2s[3 * b + 1], d2s[3 * b + 2]));
      if (valmost0(cr)) return false;
      n = vnormalize(cr);
      return true;
    };
    for (int e0 = base + 2 * q; e0 < lim; e0 += 8) {
      V3 na = mk(0, 0, 0), nb = mk(0, 0, 0);
      const bool va = axis_of(e0, na), vb = e0 + 1 < lim && axis_of(e0 + 1, nb);
      if (va && vb) {
        float l1a, h1a, l1b, h1b, l2a, h2a, l2b, h2b;
        scr_project2(s1, na, nb, l1a, h1a, l1b, h1b); scr_project2(s2, na, nb, l2a, h2a, l2b, h2b);
        keep(e0 - base, edge_est(l1a, h1a, l2a, h2a));
        keep(e0 + 1 - base, edge_est(l1b, h1b, l2b, h2b));
      } else if (va || vb) {
        const V3 n = va ? na : nb;
        float lo1, hi1, lo2, hi2; scr_project(s1, n, lo1, hi1); scr_project(s2, n, lo2, hi2);
        keep((va ? e0 : e0 + 1) - base, edge_est(lo1, hi1, lo2, hi2));
      }
    }
    thr = quad_min(thr, qm);
    for (int i = 0; i < cnt; i++) if (!(estv[i] - 2 * sE > thr)) emask |= 1ull << ev[i];
    unsigned lo = (unsigned)emask, hi = (unsigned)(emask >> 32);
    lo |= __shfl_xor_sync(qm, lo, 1); lo |= __shfl_xor_sync(qm, lo, 2);
    hi |= __shfl_xor_sync(qm, hi, 1); hi |= __shfl_xor_sync(qm, hi, 2);
    emask = ((unsigned long long)hi << 32) | lo;
  }
  // ---- edge axes, pass 2
  while (emask) {
    const int e = base + __ffsll((long long)emask) - 1; emask &= emask - 1;
    const int a = e / c2.ne, b = e - a * c2.ne;
    const V3 n = vnormalize(vcross(mk(d1s[3 * a], d1s[3 * a + 1], d1s[3 * a + 2]), mk(d2s[3 * b], d2s[3 * b + 1], d2s[3 * b + 2])));
    double mn1, mx1, mn2, mx2, dist;
    project_convex_quad(n, c1, q, qm, mn1, mx1);
    project_convex_quad(n, c2, q, qm, mn2, mx2);
    if (!overlap(mn1, mx1, mn2, mx2, dist)) { sep = c1.ng + c2.ng + e; return false; }       // EdgeSeparates
    if (dist - emin_ < 0) { R.edgeBeats = true; R.eAxis = n; R.dir1 = a + 1; R.dir2 = b + 1; emin_ = dist; }
  }
  }
  return true;
}
// one hull pair by the four lanes of a quad; `out` of lane 0 receives the contacts, hint must be the same in all four lanes
EPD void convex_convex_quad(const Cvx& c1, const Cvx& c2, ContactSink& out, TagPt* pa, TagPt* pb, bool& polyOverflow, int q, unsigned qm) {
  const bool screened = !c1.box && !c2.box && c1.nv <= kScrVerts && c2.nv <= kScrVerts && c1.nv > 0 && c2.nv > 0 &&
                        c1.ng + c2.ng <= 32 && c1.ne <= kScrEdges && c2.ne <= kScrEdges && c1.ne > 0 && c2.ne > 0;
  if (!screened) { if (q == 0) convex_convex_t<true>(c1, c2, out, pa, pb, polyOverflow); return; }      // the one-lane path
  if (sat_axis_separates_quad(c1, c2, out.hint, q, qm)) { out.sep = out.hint; return; }
  SatResult R;
  if (!convex_sat_screened_quad(c1, c2, R, out.sep, q, qm)) return;
  if (q == 0) convex_contacts_from_sat(c1, c2, R, out, pa, pb, polyOverflow);
}

// ---- CapsuleConvex.elm ----------------------------------------------------------------------------
// testCapsuleConvexAxis 718-749
EPD bool capsule_axis(const Cap& cap, const Cvx& c, V3 n, double& depth) {
  double cp = vdot(cap.pos, n);
  double ac = eabs(vdot(cap.axis, n)) * cap.h;
  double cmin = cp - ac - cap.r, cmax = cp + ac + cap.r;
  double mn, mx; project_convex(n, c, mn, mx);
  double d1 = cmax - mn, d2 = mx - cmin;
  if (d1 + kBreaking < 0 || d2 + kBreaking < 0) return false;
  depth = (d1 - d2 > 0) ? d2 : d1;
  return true;
}
// addDirectContact 237-259
EPD void cap_direct(bool flip, int64_t fk, V3 sp, V3 sep, double pen, const Cap& cap, ContactSink& out) {
  V3 n = vneg(sep);
  V3 pi = mk(sp.x + cap.r * n.x, sp.y + cap.r * n.y, sp.z + cap.r * n.z);
  V3 pj = mk(pi.x - pen * n.x, pi.y - pen * n.y, pi.z - pen * n.z);
  out.emit(flip, fk, n, pi, pj);
}
// clipSegmentAgainstFace 333-388 ; false == ClipDead
EPD bool clip_segment(const Cvx& c, const Face& f, V3& p1, V3& p2) {
  if (f.cnt < 2) return true;
  for (int e = 0; e < f.cnt; e++) {
    V3 v1 = c.vert(f.idx[e]), v2 = c.vert(f.idx[e + 1 < f.cnt ? e + 1 : 0]);
    V3 edge = vnormalize(vsub(v1, v2));
    V3 pn = vcross(f.n, edge);
    double pc = -(vdot(v1, pn));
    double d1 = vdot(pn, p1) + pc, d2 = vdot(pn, p2) + pc;
    if (d1 < 0 && d2 < 0) continue;
    else if (d1 >= 0 && d2 >= 0) return false;
    else if (d1 < 0) p2 = vlerp(d1 / (d1 - d2), p1, p2);
    else p1 = vlerp(d1 / (d1 - d2), p1, p2);
  }
  return true;
}
// supportFeature 607-716 ; returns 0, 1 or 2
EPD int support_feature(V3 axis, const Cvx& c, V3& s1, V3& s2) {
  int n = cvx_nverts(c);
  double maxProj = -kMaxNumber;
  for (int i = 0; i < n; i++) { double p = vdot(cvx_vertex(c, i), axis); maxProj = (p - maxProj > 0) ? p : maxProj; }
  int count = 0; bool two = false;
  for (int i = 0; i < n; i++) {
    V3 v = cvx_vertex(c, i);
    if (maxProj - vdot(v, axis) - 1.0e-4 < 0) { if (count == 0) { count = 1; s1 = v; } else { two = true; break; } }
  }
  if (count == 0) return 0;
  if (!two) return 1;
  for (int g = 0; g < c.ne; g++) {
    int ne; const int* e = c.edges(g, ne);
    for (int k = 0; k + 1 < ne; k += 2) {
      V3 a = c.vert(e[k]), b = c.vert(e[k + 1]);
      if ((maxProj - vdot(a, axis) - 1.0e-4 < 0) && (maxProj - vdot(b, axis) - 1.0e-4 < 0)) { s1 = a; s2 = b; return 2; }
    }
  }
  return 1;
}
// addBodyContacts 262-312
EPD void cap_body_contacts(bool flip, int cf, const Cvx& c, bool haveFace, const Face& f, const Cap& cap, V3 ep1, V3 ep2, V3 sep, V3 capPoint, ContactSink& out) {
  if (!haveFace) return;
  V3 q1 = ep1, q2 = ep2;
  if (!clip_segment(c, f, q1, q2)) return;
  double fpc = f.cnt > 0 ? -(vdot(f.n, c.vert(f.idx[0]))) : 0;
  for (int k = 0; k < 2; k++) {
    V3 pt = k == 0 ? q1 : q2;
    if (vdist(pt, capPoint) < kPrecision) continue;
    double sd = vdot(f.n, pt) + fpc;
    double pen = cap.r - sd;
    if (pen <= 0) continue;
    cap_direct(flip, feat(6, k == 0 ? 3 : 4, cf, 0, 0), pt, sep, pen, cap, out);
  }
}
EPD void capsule_convex_tail(bool flip, const Cap& cap, const Cvx& c, V3 ep1, V3 ep2, V3 target, ContactSink& out);
// addContacts 58-190
EPD void capsule_convex(bool flip, const Cap& cap, const Cvx& c, ContactSink& out) {
  V3 ep1 = mk(cap.pos.x - cap.h * cap.axis.x, cap.pos.y - cap.h * cap.axis.y, cap.pos.z - cap.h * cap.axis.z);
  V3 ep2 = mk(cap.pos.x + cap.h * cap.axis.x, cap.pos.y + cap.h * cap.axis.y, cap.pos.z + cap.h * cap.axis.z);
  // findSeparatingAxis 494-572: group normals, then closest-point axes of every unique edge
  // (screening these axes in fp32 like convex_sat_screened was measured SLOWER: forming an edge's axis -- closest points
  // of two segments -- costs as much as projecting the hull on it, and the two passes form every axis twice)
  // axis ids (ContactSink::hint / sep): group normal k = k, edge (g, k) = ng + (g << 8 | k >> 1)
  auto edge_axis = [&](const int* e, int k, V3& ax) -> bool {      // edgeAxis 575-599; false = the edge yields no axis
    V3 v1 = c.vert(e[k]), v2 = c.vert(e[k + 1]);
    V3 pc, pe; closest_segments(ep1, ep2, v1, v2, pc, pe);
    V3 diff = vsub(pc, pe); double d2 = vlen2(diff);
    if (d2 - kPrecision > 0) ax = vscale(1 / sqrt(d2), diff);
    else { V3 cr = vcross(cap.axis, vsub(v2, v1)); if (valmost0(cr)) return false; ax = vnormalize(cr); }
    return true;
  };
  if (out.hint >= 0) {       // still apart along the axis that separated last frame?
    V3 ax; bool have = false; double dist;
    if (out.hint < c.ng) { ax = ld3(c.gf(out.hint)); have = true; }
    else {
      const int h = out.hint - c.ng, g = h >> 8, k = (h & 0xff) * 2;
      if (g < c.ne) { int ne; const int* e = c.edges(g, ne); if (k + 1 < ne) have = edge_axis(e, k, ax); }
    }
    if (have && !capsule_axis(cap, c, ax, dist)) { out.sep = out.hint; return; }
  }
  V3 target = mk(0, 0, 0); double dmin = kMaxNumber;
  for (int k = 0; k < c.ng; k++) {
    V3 n = ld3(c.gf(k)); double dist;
    if (!capsule_axis(cap, c, n, dist)) { out.sep = k; return; }
    if (dist - dmin < 0) { target = n; dmin = dist; }
  }
  for (int g = 0; g < c.ne; g++) {
    int ne; const int* e = c.edges(g, ne);
    for (int k = 0; k + 1 < ne; k += 2) {
      V3 ax;
      if (!edge_axis(e, k, ax)) continue;
      double dist;
      if (!capsule_axis(cap, c, ax, dist)) { if (g < (1 << 20) && k < 512) out.sep = c.ng + (g << 8 | k >> 1); return; }
      if (dist - dmin < 0) { target = ax; dmin = dist; }
    }
  }
  capsule_convex_tail(flip, cap, c, ep1, ep2, target, out);
}
// addContacts 58-190 behind findSeparatingAxis: the contacts along `target`
EPD void capsule_convex_tail(bool flip, const Cap& cap, const Cvx& c, V3 ep1, V3 ep2, V3 target, ContactSink& out) {
  V3 sep = vdot(vsub(c.pos(), cap.pos), target) > 0 ? vneg(target) : target;
  double pen;
  if (!capsule_axis(cap, c, sep, pen)) return;
  double t = vdot(cap.axis, sep);
  Face f; int fid = best_face(c, vneg(sep), f);
  bool haveFace = (fid != -1) && (vdot(f.n, sep) - 0.999 > 0);
  V3 s1 = mk(0, 0, 0), s2 = mk(0, 0, 0); int ns = 0;
  if (!haveFace) ns = support_feature(sep, c, s1, s2);
  int cf = haveFace ? on_face(fid) : (ns >= 2 ? kOnEdge : (ns == 1 ? kOnVertex : kOnNone));
  if (t - 1.0e-3 > 0) {
    cap_direct(flip, feat(6, 1, cf, 0, 0), ep1, sep, pen, cap, out);
    cap_body_contacts(flip, cf, c, haveFace, f, cap, ep1, ep2, sep, ep1, out);
  } else if (t + 1.0e-3 < 0) {
    cap_direct(flip, feat(6, 2, cf, 0, 0), ep2, sep, pen, cap, out);
    cap_body_contacts(flip, cf, c, haveFace, f, cap, ep1, ep2, sep, ep2, out);
  } else if (haveFace) {
    V3 p1 = ep1, p2 = ep2;
    if (clip_segment(c, f, p1, p2)) {
      cap_direct(flip, feat(6, 3, cf, 0, 0), p1, sep, pen, cap, out);
      cap_direct(flip, feat(6, 4, cf, 0, 0), p2, sep, pen, cap, out);
    } else {   // addClosestEdgeContact 395-491
      double r2 = cap.r * cap.r, bestD2 = r2; V3 bc = mk(0, 0, 0), be = mk(0, 0, 0);
      if (f.cnt >= 2)
        for (int e = 0; e < f.cnt; e++) {
          V3 v1 = c.vert(f.idx[e]), v2 = c.vert(f.idx[e + 1 < f.cnt ? e + 1 : 0]);
          V3 pc, pe; closest_segments(ep1, ep2, v1, v2, pc, pe);
          double dx = pc.x - pe.x, dy = pc.y - pe.y, dz = pc.z - pe.z, d2 = dx * dx + dy * dy + dz * dz;
          if (d2 - bestD2 < 0) { bc = pc; be = pe; bestD2 = d2; }
        }
      if (bestD2 - r2 >= 0) return;
      double dist = sqrt(bestD2);
      if (dist - kPrecision <= 0) return;
      double inv = 1 / dist;
      cap_direct(flip, feat(6, 5, cf, 0, 0), bc, mk((bc.x - be.x) * inv, (bc.y - be.y) * inv, (bc.z - be.z) * inv), cap.r - dist, cap, out);
    }
  } else if (ns >= 2) {
    if (valmost0(vcross(cap.axis, vsub(s2, s1)))) {   // addParallelEdgeContacts 197-234
      double a1 = vdot(ep1, cap.axis), a2 = vdot(ep2, cap.axis), b1 = vdot(s1, cap.axis), b2 = vdot(s2, cap.axis);
      double lo = emax(emin(a1, a2), emin(b1, b2)), hi = emin(emax(a1, a2), emax(b1, b2));
      if (hi - lo - kPrecision > 0) {
        cap_direct(flip, feat(6, 3, cf, 0, 0), mk(ep1.x + (lo - a1) * cap.axis.x, ep1.y + (lo - a1) * cap.axis.y, ep1.z + (lo - a1) * cap.axis.z), sep, pen, cap, out);
        cap_direct(flip, feat(6, 4, cf, 0, 0), mk(ep1.x + (hi - a1) * cap.axis.x, ep1.y + (hi - a1) * cap.axis.y, ep1.z + (hi - a1) * cap.axis.z), sep, pen, cap, out);
      } else {
        V3 pc, pe; closest_segments(ep1, ep2, s1, s2, pc, pe);
        cap_direct(flip, feat(6, 5, cf, 0, 0), pc, sep, pen, cap, out);
      }
    } else {
      V3 pc, pe; closest_segments(ep1, ep2, s1, s2, pc, pe);
      cap_direct(flip, feat(6, 5, cf, 0, 0), pc, sep, pen, cap, out);
    }
  } else if (ns == 1) {
    V3 pc, pe; closest_segments(ep1, ep2, s1, s1, pc, pe);
    cap_direct(flip, feat(6, 5, cf, 0, 0), pc, sep, pen, cap, out);
  }
}

// ---- CapsuleConvex by a QUAD of lanes: the axes of findSeparatingAxis dealt round-robin over the four lanes -------------------
// Axis t of the reference's walk (group normals, then every (edge group, edge) in order, t counting the edges that yield no axis
// as well) goes to lane t & 3, which evaluates it exactly as the one-lane walk does.  A separating axis found by any lane ends the
// pair (every separating axis means "no contacts"); otherwise the target is the axis of smallest depth, the FIRST one among equals
// (the walk replaces on strict improvement only) = the lexicographic minimum of (depth, t) over the quad.  Lane 0 then forms the
// contacts.  hint must be the same in all four lanes.
EPD void capsule_convex_quad(bool flip, const Cap& cap, const Cvx& c, ContactSink& out, int q, unsigned qm) {
  V3 ep1 = mk(cap.pos.x - cap.h * cap.axis.x, cap.pos.y - cap.h * cap.axis.y, cap.pos.z - cap.h * cap.axis.z);
  V3 ep2 = mk(cap.pos.x + cap.h * cap.axis.x, cap.pos.y + cap.h * cap.axis.y, cap.pos.z + cap.h * cap.axis.z);
  auto edge_axis = [&](const int* e, int k, V3& ax) -> bool {      // edgeAxis 575-599; false = the edge yields no axis
    V3 v1 = c.vert(e[k]), v2 = c.vert(e[k + 1]);
    V3 pc, pe; closest_segments(ep1, ep2, v1, v2, pc, pe);
    V3 diff = vsub(pc, pe); double d2 = vlen2(diff);
    if (d2 - kPrecision > 0) ax = vscale(1 / sqrt(d2), diff);
    else { V3 cr = vcross(cap.axis, vsub(v2, v1)); if (valmost0(cr)) return false; ax = vnormalize(cr); }
    return true;
  };
  if (out.hint >= 0) {       // still apart along the axis that separated last frame?  (one lane's work, done by all four alike)
    V3 ax; bool have = false; double dist;
    if (out.hint < c.ng) { ax = ld3(c.gf(out.hint)); have = true; }
    else {
      const int h = out.hint - c.ng, g = h >> 8, k = (h & 0xff) * 2;
      if (g < c.ne) { int ne; const int* e = c.edges(g, ne); if (k + 1 < ne) have = edge_axis(e, k, ax); }
    }
    if (have && !capsule_axis(cap, c, ax, dist)) { out.sep = out.hint; return; }
  }
  V3 target = mk(0, 0, 0); double dmin = kMaxNumber; int tmin = 0x7fffffff;
  int nedges = 0;
  for (int g = 0; g < c.ne; g++) { int ne; c.edges(g, ne); nedges += ne >> 1; }
  const int total = c.ng + nedges;
  // rounds of four axes, one per lane, and a vote after each: the pair ends at the first round that holds a separating axis
  // (the one-lane walk stops at its first separating axis; without the vote the other three lanes would walk on to the end)
  for (int t0 = 0; t0 < total; t0 += 4) {
    const int t = t0 + q;
    int sepId = -1;
    if (t < total) {
      V3 ax; bool have = true; int id = t;
      if (t < c.ng) ax = ld3(c.gf(t));
      else {
        int te = t - c.ng, g = 0, ne = 0; const int* e = c.edges(0, ne);
        while (te >= (ne >> 1)) { te -= ne >> 1; g++; e = c.edges(g, ne); }
        have = edge_axis(e, 2 * te, ax);
        id = (g < (1 << 20) && te < 256) ? c.ng + (g << 8 | te) : 0x7ffffffe;
      }
      if (have) {
        double dist;
        if (!capsule_axis(cap, c, ax, dist)) sepId = id;
        else if (dist - dmin < 0) { target = ax; dmin = dist; tmin = t; }
      }
    }
    unsigned sid = sepId < 0 ? 0xffffffffu : (unsigned)sepId;
    sid = min(sid, __shfl_xor_sync(qm, sid, 1)); sid = min(sid, __shfl_xor_sync(qm, sid, 2));
    if (sid != 0xffffffffu) { if (sid < 0x7ffffffeu) out.sep = (int)sid; return; }
  }
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1) {      // lexicographic minimum of (depth, t)
    const double od = __shfl_xor_sync(qm, dmin, o); const int ot = __shfl_xor_sync(qm, tmin, o);
    const double ox = __shfl_xor_sync(qm, target.x, o), oy = __shfl_xor_sync(qm, target.y, o), oz = __shfl_xor_sync(qm, target.z, o);
    const bool take = (od - dmin < 0) || (!(dmin - od < 0) && ot < tmin);
    if (take) { dmin = od; tmin = ot; target = mk(ox, oy, oz); }
  }
  if (q == 0) capsule_convex_tail(flip, cap, c, ep1, ep2, target, out);
}

// ---- NarrowPhase.elm:86-286 : one (shape1, shape2) cell of the 5x5 matrix ---------------------------
struct ShapeRef { int kind; const double* f; const int* ib; };
EPD Sph as_sphere(const ShapeRef& s) { Sph r; r.pos = ld3(s.f); r.r = s.f[3]; return r; }
EPD Cap as_capsule(const ShapeRef& s) { Cap c; c.pos = ld3(s.f); c.axis = ld3(s.f + 3); c.r = s.f[6]; c.h = s.f[7]; return c; }
EPD Pln as_plane(const ShapeRef& s) { Pln p; p.pos = ld3(s.f); p.n = ld3(s.f + 3); return p; }

// The same dispatch restricted to one GROUP of modules, so that a kernel instantiated for the group carries only that code (and
// only the hull-hull group carries the clip-polygon buffers): 0 = convex-convex (ConvexConvex.elm), 1 = capsule-convex
// (CapsuleConvex.elm), 2 = everything else (plane / sphere / particle against anything, capsule-capsule), 3 = convex-convex
// with a box on either side (group 0 then holds the general hull pairs only).
template <int G>
EPD void shape_pair_contacts_group(const ShapeRef& a, const ShapeRef& b, ContactSink& out, TagPt* pa, TagPt* pb, bool& polyOverflow) {
  const int key = a.kind * 5 + b.kind;
  if (G == 0) { if (key == SH_CONVEX * 5 + SH_CONVEX) convex_convex_t<true>(make_cvx(a.f, a.ib), make_cvx(b.f, b.ib), out, pa, pb, polyOverflow); return; }
  if (G == 3) { if (key == SH_CONVEX * 5 + SH_CONVEX) convex_convex_t<false>(make_cvx(a.f, a.ib), make_cvx(b.f, b.ib), out, pa, pb, polyOverflow); return; }
  if (G == 1) {
    if (key == SH_CONVEX * 5 + SH_CAPSULE) capsule_convex(true, as_capsule(b), make_cvx(a.f, a.ib), out);
    else if (key == SH_CAPSULE * 5 + SH_CONVEX) capsule_convex(false, as_capsule(a), make_cvx(b.f, b.ib), out);
    return;
  }
  switch (key) {
    case SH_CONVEX * 5 + SH_PLANE: plane_convex(true, as_plane(b), make_cvx(a.f, a.ib), out, pa, pb, polyOverflow); break;
    case SH_CONVEX * 5 + SH_SPHERE: sphere_convex(true, as_sphere(b), make_cvx(a.f, a.ib), out); break;
    case SH_CONVEX * 5 + SH_PARTICLE: particle_convex(true, ld3(b.f), make_cvx(a.f, a.ib), out); break;
    case SH_PLANE * 5 + SH_CONVEX: plane_convex(false, as_plane(a), make_cvx(b.f, b.ib), out, pa, pb, polyOverflow); break;
    case SH_PLANE * 5 + SH_SPHERE: plane_sphere(false, as_plane(a), as_sphere(b), out); break;
    case SH_PLANE * 5 + SH_PARTICLE: plane_point(false, 0, as_plane(a), ld3(b.f), out); break;
    case SH_PLANE * 5 + SH_CAPSULE: plane_capsule(false, as_plane(a), as_capsule(b), out); break;
    case SH_SPHERE * 5 + SH_PLANE: plane_sphere(true, as_plane(b), as_sphere(a), out); break;
    case SH_SPHERE * 5 + SH_CONVEX: sphere_convex(false, as_sphere(a), make_cvx(b.f, b.ib), out); break;
    case SH_SPHERE * 5 + SH_SPHERE: sphere_sphere(as_sphere(a), as_sphere(b), out); break;
    case SH_SPHERE * 5 + SH_PARTICLE: sphere_particle(false, as_sphere(a), ld3(b.f), out); break;
    case SH_SPHERE * 5 + SH_CAPSULE: capsule_sphere(true, as_capsule(b), as_sphere(a), out); break;
    case SH_PARTICLE * 5 + SH_PLANE: plane_point(true, 0, as_plane(b), ld3(a.f), out); break;
    case SH_PARTICLE * 5 + SH_CONVEX: particle_convex(false, ld3(a.f), make_cvx(b.f, b.ib), out); break;
    case SH_PARTICLE * 5 + SH_SPHERE: sphere_particle(true, as_sphere(b), ld3(a.f), out); break;
    case SH_PARTICLE * 5 + SH_CAPSULE: capsule_particle(true, as_capsule(b), ld3(a.f), out); break;
    case SH_CAPSULE * 5 + SH_PLANE: plane_capsule(true, as_plane(b), as_capsule(a), out); break;
    case SH_CAPSULE * 5 + SH_SPHERE: capsule_sphere(false, as_capsule(a), as_sphere(b), out); break;
    case SH_CAPSULE * 5 + SH_CAPSULE: capsule_capsule(as_capsule(a), as_capsule(b), out); break;
    case SH_CAPSULE * 5 + SH_PARTICLE: capsule_particle(false, as_capsule(a), ld3(b.f), out); break;
    default: break;   // plane-plane, particle-particle: NarrowPhase.elm:132-134, 235-237
  }
}

EPD void shape_pair_contacts(const ShapeRef& a, const ShapeRef& b, ContactSink& out, TagPt* pa, TagPt* pb, bool& polyOverflow) {
  switch (a.kind * 5 + b.kind) {
    case SH_CONVEX * 5 + SH_CONVEX: convex_convex(make_cvx(a.f, a.ib), make_cvx(b.f, b.ib), out, pa, pb, polyOverflow); break;
    case SH_CONVEX * 5 + SH_PLANE: plane_convex(true, as_plane(b), make_cvx(a.f, a.ib), out, pa, pb, polyOverflow); break;
    case SH_CONVEX * 5 + SH_SPHERE: sphere_convex(true, as_sphere(b), make_cvx(a.f, a.ib), out); break;
    case SH_CONVEX * 5 + SH_PARTICLE: particle_convex(true, ld3(b.f), make_cvx(a.f, a.ib), out); break;
    case SH_CONVEX * 5 + SH_CAPSULE: capsule_convex(true, as_capsule(b), make_cvx(a.f, a.ib), out); break;
    case SH_PLANE * 5 + SH_CONVEX: plane_convex(false, as_plane(a), make_cvx(b.f, b.ib), out, pa, pb, polyOverflow); break;
    case SH_PLANE * 5 + SH_SPHERE: plane_sphere(false, as_plane(a), as_sphere(b), out); break;
    case SH_PLANE * 5 + SH_PARTICLE: plane_point(false, 0, as_plane(a), ld3(b.f), out); break;
    case SH_PLANE * 5 + SH_CAPSULE: plane_capsule(false, as_plane(a), as_capsule(b), out); break;
    case SH_SPHERE * 5 + SH_PLANE: plane_sphere(true, as_plane(b), as_sphere(a), out); break;
    case SH_SPHERE * 5 + SH_CONVEX: sphere_convex(false, as_sphere(a), make_cvx(b.f, b.ib), out); break;
    case SH_SPHERE * 5 + SH_SPHERE: sphere_sphere(as_sphere(a), as_sphere(b), out); break;
    case SH_SPHERE * 5 + SH_PARTICLE: sphere_particle(false, as_sphere(a), ld3(b.f), out); break;
    case SH_SPHERE * 5 + SH_CAPSULE: capsule_sphere(true, as_capsule(b), as_sphere(a), out); break;
    case SH_PARTICLE * 5 + SH_PLANE: plane_point(true, 0, as_plane(b), ld3(a.f), out); break;
    case SH_PARTICLE * 5 + SH_CONVEX: particle_convex(false, ld3(a.f), make_cvx(b.f, b.ib), out); break;
    case SH_PARTICLE * 5 + SH_SPHERE: sphere_particle(true, as_sphere(b), ld3(a.f), out); break;
    case SH_PARTICLE * 5 + SH_CAPSULE: capsule_particle(true, as_capsule(b), ld3(a.f), out); break;
    case SH_CAPSULE * 5 + SH_PLANE: plane_capsule(true, as_plane(b), as_capsule(a), out); break;
    case SH_CAPSULE * 5 + SH_CONVEX: capsule_convex(false, as_capsule(a), make_cvx(b.f, b.ib), out); break;
    case SH_CAPSULE * 5 + SH_SPHERE: capsule_sphere(false, as_capsule(a), as_sphere(b), out); break;
    case SH_CAPSULE * 5 + SH_CAPSULE: capsule_capsule(as_capsule(a), as_capsule(b), out); break;
    case SH_CAPSULE * 5 + SH_PARTICLE: capsule_particle(false, as_capsule(a), ld3(b.f), out); break;
    default: break;   // plane-plane, particle-particle: NarrowPhase.elm:132-134, 235-237
  }
}

}  // namespace ep
