// ORACLE — TEST INFRASTRUCTURE ONLY (see ep_oracle_math.h).
// CPU restatement of the narrowphase: src/Internal/NarrowPhase.elm, src/Internal/Manifold.elm,
// src/Internal/ContactId.elm and the 13 modules of src/Collision/.  Elm cons-lists are std::vectors
// that grow with push_back ("cons"), i.e. a vector holds a list in REVERSE (emission) order.
#include "ep_oracle_internal.h"

namespace epo {

// ---- ContactId.elm:94-287 -----------------------------------------------------------------------
static inline int64_t feature(int64_t tag, int64_t a, int64_t b, int64_t c, int64_t d) {
  return (((tag * 2048 + a) * 2048 + b) * 2048 + c) * 2048 + d;
}
static const int64_t kSimple = 0;
static inline int64_t convexConvexFace(int f1, int f2, int v) { return feature(1, f1, f2, v, 0); }
static inline int64_t convexConvexEdge(int d1, int e1, int d2, int e2) { return feature(2, d1, e1, d2, e2); }
static inline int64_t planeVertex(int v) { return feature(3, v, 0, 0, 0); }
static inline int64_t planeCapEnd(int e) { return feature(4, e, 0, 0, 0); }
static inline int onConvexFace(int faceId) { return faceId + 2; }
static const int onConvexEdge = 1, onConvexVertex = 2, onConvexNone = 0;
static inline int64_t sphereOnConvex(int cf) { return feature(5, cf, 0, 0, 0); }
static inline int64_t capsuleCapEnd1(int cf) { return feature(6, 1, cf, 0, 0); }
static inline int64_t capsuleCapEnd2(int cf) { return feature(6, 2, cf, 0, 0); }
static inline int64_t capsuleCylinder1(int cf) { return feature(6, 3, cf, 0, 0); }
static inline int64_t capsuleCylinder2(int cf) { return feature(6, 4, cf, 0, 0); }
static inline int64_t capsuleCylinder(int cf) { return feature(6, 5, cf, 0, 0); }

// Contact.elm:36-43
static inline Contact orderContact(bool flip, Contact c) {
  if (!flip) return c;
  return Contact{c.shapeKey, c.featureKey, neg(c.ni), c.pj, c.pi};
}

typedef std::vector<Contact> CL;   // emission order (push_back == cons)

// ---- Manifold.elm:25-169 ------------------------------------------------------------------------
struct TagPoint { int tag; V3 p; };
typedef std::vector<TagPoint> PL;  // FORWARD list order here (index 0 == head)

static TagPoint farthestFrom(const TagPoint& a, const PL& all) {
  if (all.empty()) return a;
  TagPoint best = all[0]; double bestScore = distanceSquared(a.p, all[0].p);
  for (size_t i = 1; i < all.size(); i++) { double s = distanceSquared(a.p, all[i].p); if (s - bestScore > 0) { best = all[i]; bestScore = s; } }
  return best;
}
static TagPoint farthestFrom2(const TagPoint& a, const TagPoint& b, const PL& all) {
  if (all.empty()) return a;
  TagPoint best = all[0]; double bestScore = emin(distanceSquared(a.p, all[0].p), distanceSquared(b.p, all[0].p));
  for (size_t i = 1; i < all.size(); i++) {
    double s = emin(distanceSquared(a.p, all[i].p), distanceSquared(b.p, all[i].p));
    if (s - bestScore > 0) { best = all[i]; bestScore = s; }
  }
  return best;
}
static TagPoint farthestFrom3(const TagPoint& a, const TagPoint& b, const TagPoint& c, const PL& all) {
  if (all.empty()) return a;
  auto sc = [&](const TagPoint& q) { return emin(distanceSquared(a.p, q.p), emin(distanceSquared(b.p, q.p), distanceSquared(c.p, q.p))); };
  TagPoint best = all[0]; double bestScore = sc(all[0]);
  for (size_t i = 1; i < all.size(); i++) { double s = sc(all[i]); if (s - bestScore > 0) { best = all[i]; bestScore = s; } }
  return best;
}

// reduce: input `points` in forward list order; output forward list order.
static PL manifoldReduce(V3 normal, double planeConstant, const PL& points) {
  double marginConstant = planeConstant - kContactBreakingThreshold;
  PL cand;   // withinMargin reverses (Manifold.elm:37-49)
  for (const TagPoint& v : points) if (dot(normal, v.p) + marginConstant < 0) cand.insert(cand.begin(), v);
  if (cand.size() < 5) return cand;
  TagPoint bestA = cand[0], bestB = cand[0]; double bestD = -1;
  for (const TagPoint& p : cand) {          // farthestPair, Manifold.elm:54-79
    TagPoint q = farthestFrom(p, cand);
    double d = distanceSquared(p.p, q.p);
    if (d - bestD - 1.0e-8 > 0) { bestA = p; bestB = q; bestD = d; }
  }
  TagPoint c = farthestFrom2(bestA, bestB, cand);
  TagPoint e = farthestFrom3(bestA, bestB, c, cand);
  return PL{bestA, bestB, c, e};
}

// ---- primitive pairs ----------------------------------------------------------------------------
// PlaneSphere.elm:10-42
static void planeSphere(int shapeKey, bool flip, const WShape& plane, const WShape& sphere, CL& out) {
  V3 normal = plane.normal, position = plane.position;
  V3 vertex{sphere.position.x - sphere.radius * normal.x, sphere.position.y - sphere.radius * normal.y, sphere.position.z - sphere.radius * normal.z};
  double d = ((vertex.x - position.x) * normal.x) + ((vertex.y - position.y) * normal.y) + ((vertex.z - position.z) * normal.z);
  if (d - kContactBreakingThreshold < 0)
    out.push_back(orderContact(flip, Contact{shapeKey, kSimple, normal, V3{vertex.x - d * normal.x, vertex.y - d * normal.y, vertex.z - d * normal.z}, vertex}));
}
// PlaneParticle.elm:10-33
static void planeParticle(int shapeKey, bool flip, const WShape& plane, V3 pp, CL& out) {
  V3 normal = plane.normal, position = plane.position;
  double d = ((pp.x - position.x) * normal.x) + ((pp.y - position.y) * normal.y) + ((pp.z - position.z) * normal.z);
  if (d - kContactBreakingThreshold < 0)
    out.push_back(orderContact(flip, Contact{shapeKey, kSimple, normal, V3{pp.x - d * normal.x, pp.y - d * normal.y, pp.z - d * normal.z}, pp}));
}
// PlaneCapsule.elm:11-61
static void planeCapsule(int shapeKey, bool flip, const WShape& plane, const WShape& cap, CL& out) {
  V3 normal = plane.normal, pos = plane.position;
  V3 ep1{cap.position.x - cap.halfLength * cap.axis.x, cap.position.y - cap.halfLength * cap.axis.y, cap.position.z - cap.halfLength * cap.axis.z};
  V3 ep2{cap.position.x + cap.halfLength * cap.axis.x, cap.position.y + cap.halfLength * cap.axis.y, cap.position.z + cap.halfLength * cap.axis.z};
  V3 eps[2] = {ep1, ep2};
  for (int k = 0; k < 2; k++) {
    V3 ep = eps[k];
    V3 vertex{ep.x - cap.radius * normal.x, ep.y - cap.radius * normal.y, ep.z - cap.radius * normal.z};
    double d = ((vertex.x - pos.x) * normal.x) + ((vertex.y - pos.y) * normal.y) + ((vertex.z - pos.z) * normal.z);
    if (d - kContactBreakingThreshold < 0)
      out.push_back(orderContact(flip, Contact{shapeKey, planeCapEnd(k + 1), normal, V3{vertex.x - d * normal.x, vertex.y - d * normal.y, vertex.z - d * normal.z}, vertex}));
  }
}
// SphereSphere.elm:9-42
static void sphereSphere(int shapeKey, const WShape& s1, const WShape& s2, CL& out) {
  V3 c1 = s1.position, c2 = s2.position;
  double dist = distance(c2, c1) - s1.radius - s2.radius;
  V3 normal = direction(c2, c1);
  if (dist > 0) return;
  out.push_back(Contact{shapeKey, kSimple, normal, add(c1, scale(s1.radius - dist, normal)), add(c2, scale(-s2.radius, normal))});
}
// SphereParticle.elm:9-29
static void sphereParticle(int shapeKey, bool flip, const WShape& s, V3 pp, CL& out) {
  double dist = distance(pp, s.position) - s.radius;
  V3 normal = direction(pp, s.position);
  if (dist > 0) return;
  out.push_back(orderContact(flip, Contact{shapeKey, kSimple, normal, add(s.position, scale(s.radius - dist, normal)), pp}));
}
// CapsuleSphere.elm:10-39
static void capsuleSphere(int shapeKey, bool flip, const WShape& cap, const WShape& sph, CL& out) {
  double t = emax(-cap.halfLength, emin(cap.halfLength, dot(sub(sph.position, cap.position), cap.axis)));
  V3 closest = add(cap.position, scale(t, cap.axis));
  double dist = distance(sph.position, closest) - cap.radius - sph.radius;
  V3 normal = direction(sph.position, closest);
  if (dist > 0) return;
  out.push_back(orderContact(flip, Contact{shapeKey, kSimple, normal, add(closest, scale(cap.radius, normal)), sub(sph.position, scale(sph.radius, normal))}));
}
// CapsuleParticle.elm:9-38
static void capsuleParticle(int shapeKey, bool flip, const WShape& cap, V3 pp, CL& out) {
  double t = emax(-cap.halfLength, emin(cap.halfLength, dot(sub(pp, cap.position), cap.axis)));
  V3 closest = add(cap.position, scale(t, cap.axis));
  double dist = distance(pp, closest) - cap.radius;
  V3 normal = direction(pp, closest);
  if (dist > 0) return;
  out.push_back(orderContact(flip, Contact{shapeKey, kSimple, normal, add(closest, scale(cap.radius, normal)), pp}));
}
// CapsuleCapsule.elm:10-97
static void capsuleCapsule(int shapeKey, const WShape& c1, const WShape& c2, CL& out) {
  V3 p1 = c1.position, p2 = c2.position, d1 = c1.axis, d2 = c2.axis;
  double h1 = c1.halfLength, h2 = c2.halfLength;
  V3 r = sub(p1, p2);
  double b = dot(d1, d2), c = dot(d1, r), f = dot(d2, r);
  double denom = 1 - b * b;
  V3 pt1, pt2;
  if (denom - kParallelTolerance < 0) {
    double t = emax(-h2, emin(h2, f));
    double s = emax(-h1, emin(h1, b * t - c));
    pt1 = add(p1, scale(s, d1)); pt2 = add(p2, scale(t, d2));
  } else {
    double s0 = emax(-h1, emin(h1, (b * f - c) / denom));
    double t0 = emax(-h2, emin(h2, b * s0 + f));
    double s = emax(-h1, emin(h1, b * t0 - c));
    pt1 = add(p1, scale(s, d1)); pt2 = add(p2, scale(t0, d2));
  }
  double dist = distance(pt2, pt1) - c1.radius - c2.radius;
  V3 normal = direction(pt2, pt1);
  if (dist > 0) return;
  out.push_back(Contact{shapeKey, kSimple, normal, add(pt1, scale(c1.radius, normal)), sub(pt2, scale(c2.radius, normal))});
}

// ---- PlaneConvex.elm:12-117 ---------------------------------------------------------------------
static void planeConvex(int shapeKey, bool flip, const WShape& plane, const WConvex& cv, CL& out) {
  V3 pp = plane.position, pn = plane.normal;
  if (cv.isBox) {
    static const double S[8][3] = {{1, 1, 1}, {1, 1, -1}, {1, -1, 1}, {1, -1, -1}, {-1, 1, 1}, {-1, 1, -1}, {-1, -1, 1}, {-1, -1, -1}};
    V3 c = cv.position, ax = cv.ax, ay = cv.ay, az = cv.az, he = cv.he;
    for (int k = 0; k < 8; k++) {
      double sx = S[k][0], sy = S[k][1], sz = S[k][2];
      double vx = c.x + sx * he.x * ax.x + sy * he.y * ay.x + sz * he.z * az.x;
      double vy = c.y + sx * he.x * ax.y + sy * he.y * ay.y + sz * he.z * az.y;
      double vz = c.z + sx * he.x * ax.z + sy * he.y * ay.z + sz * he.z * az.z;
      double d = ((vx - pp.x) * pn.x) + ((vy - pp.y) * pn.y) + ((vz - pp.z) * pn.z);
      if (d - kContactBreakingThreshold < 0)
        out.push_back(orderContact(flip, Contact{shapeKey, planeVertex(k + 1), pn, V3{vx - d * pn.x, vy - d * pn.y, vz - d * pn.z}, V3{vx, vy, vz}}));
    }
  } else {
    PL pts;
    for (size_t i = 0; i < cv.vs.size(); i++) pts.push_back(TagPoint{(int)i + 1, cv.vs[i]});
    PL red = manifoldReduce(pn, -(dot(pn, pp)), pts);
    for (const TagPoint& tp : red) {   // emitPlaneContacts conses in list order
      V3 v = tp.p;
      double d = ((v.x - pp.x) * pn.x) + ((v.y - pp.y) * pn.y) + ((v.z - pp.z) * pn.z);
      out.push_back(orderContact(flip, Contact{shapeKey, planeVertex(tp.tag), pn, V3{v.x - d * pn.x, v.y - d * pn.y, v.z - d * pn.z}, v}));
    }
  }
}

// flat world-order face walk helpers: face f (0-based flat) of a convex
struct FlatFace { V3 normal; const int32_t* idx; int n; };
static std::vector<FlatFace> flatFaces(const WConvex& c) {
  std::vector<FlatFace> out;
  for (const WGroup& g : c.faces) {
    out.push_back({g.n1, g.i1, g.n_i1});
    if (g.twoSided) out.push_back({g.n2, g.i2, g.n_i2});
  }
  return out;
}

// ---- ParticleConvex.elm:11-80 -------------------------------------------------------------------
static void particleConvex(int shapeKey, bool flip, V3 pp, const WConvex& cv, CL& out) {
  std::vector<FlatFace> faces = flatFaces(cv);
  if (faces.empty()) return;
  double bestDepth = kMaxNumber; V3 bestNormal{0, 0, 0};
  for (const FlatFace& f : faces) {
    V3 point = f.n > 0 ? cv.buffer[f.idx[0]] : V3{0, 0, 0};
    double d = dot(f.normal, sub(point, pp));
    if (d >= 0) { if (d - bestDepth < 0) { bestDepth = d; bestNormal = f.normal; } }
    else return;
  }
  if (bestDepth - kMaxNumber < 0)
    out.push_back(orderContact(flip, Contact{shapeKey, kSimple, neg(bestNormal), pp, add(pp, scale(bestDepth, bestNormal))}));
}

// ---- SphereConvex.elm:15-250 --------------------------------------------------------------------
static void sphereEmit(int shapeKey, bool flip, int64_t fk, V3 center, V3 cp, double pen, CL& out) {
  V3 normal = direction(cp, center);
  out.push_back(orderContact(flip, Contact{shapeKey, fk, normal, V3{cp.x + pen * normal.x, cp.y + pen * normal.y, cp.z + pen * normal.z}, cp}));
}
static void sphereConvex(int shapeKey, bool flip, const WShape& sph, const WConvex& cv, CL& out) {
  V3 center = sph.position; double radius = sph.radius;
  std::vector<FlatFace> faces = flatFaces(cv);
  if (faces.empty()) return;
  std::vector<std::pair<V3, V3>> cand;   // candidateEdges as a stack: back == head
  for (size_t fi = 0; fi < faces.size(); fi++) {
    const FlatFace& f = faces[fi];
    double faceDistance = -1;
    if (f.n > 0) {
      V3 first = cv.buffer[f.idx[0]];
      faceDistance = f.normal.x * (center.x - first.x) + f.normal.y * (center.y - first.y) + f.normal.z * (center.z - first.z);
    }
    if (faceDistance > 0 && faceDistance - radius < 0) {
      bool anyOutside = false;
      if (f.n >= 2) {
        for (int k = 0; k < f.n; k++) {
          V3 v1 = cv.buffer[f.idx[k]];
          V3 v2 = (k + 1 < f.n) ? cv.buffer[f.idx[k + 1]] : cv.buffer[f.idx[0]];
          double ex = v2.x - v1.x, ey = v2.y - v1.y, ez = v2.z - v1.z;
          double cnX = f.normal.y * ez - f.normal.z * ey;
          double cnY = f.normal.z * ex - f.normal.x * ez;
          double cnZ = f.normal.x * ey - f.normal.y * ex;
          double d = cnX * (v1.x - center.x) + cnY * (v1.y - center.y) + cnZ * (v1.z - center.z);
          if (d > 0) { cand.push_back({v1, v2}); anyOutside = true; }
        }
      }
      if (!anyOutside) {
        V3 cp{center.x - faceDistance * f.normal.x, center.y - faceDistance * f.normal.y, center.z - faceDistance * f.normal.z};
        sphereEmit(shapeKey, flip, sphereOnConvex(onConvexFace((int)fi + 1)), center, cp, radius - faceDistance, out);
        return;
      }
    }
  }
  // walkBoundaries, SphereConvex.elm:158-229 — list head == most recently consed
  V3 bestPoint{0, 0, 0}; double bestDistSq = radius * radius;
  for (int k = (int)cand.size() - 1; k >= 0; k--) {
    V3 pv = cand[k].first, v = cand[k].second;
    double ex = v.x - pv.x, ey = v.y - pv.y, ez = v.z - pv.z;
    double edgeLenSq = ex * ex + ey * ey + ez * ez;
    double off = (center.x - pv.x) * ex + (center.y - pv.y) * ey + (center.z - pv.z) * ez;
    if (off < 0) {
      double dsq = distanceSquared(pv, center);
      if (dsq - bestDistSq < 0) { bestPoint = pv; bestDistSq = dsq; }
    } else if (off - edgeLenSq > 0) {
      double dsq = distanceSquared(v, center);
      if (dsq - bestDistSq < 0) { bestPoint = v; bestDistSq = dsq; }
    } else {
      double fr = off / edgeLenSq;
      V3 cp{pv.x + fr * ex, pv.y + fr * ey, pv.z + fr * ez};
      double dsq = distanceSquared(cp, center);
      if (dsq - bestDistSq < 0) { sphereEmit(shapeKey, flip, sphereOnConvex(onConvexEdge), center, cp, radius - std::sqrt(dsq), out); return; }
    }
  }
  if (bestDistSq - radius * radius < 0)
    sphereEmit(shapeKey, flip, sphereOnConvex(onConvexVertex), center, bestPoint, radius - std::sqrt(bestDistSq), out);
}

// ---- ConvexConvex.elm ---------------------------------------------------------------------------
struct MinMax { double min, max; };
// ConvexConvex.elm:689-710
static MinMax project(V3 axis, const std::vector<V3>& vs) {
  double minVal = kMaxNumber, maxVal = -kMaxNumber;
  for (V3 v : vs) {
    double val = v.x * axis.x + v.y * axis.y + v.z * axis.z;
    minVal = (minVal - val > 0) ? val : minVal;
    maxVal = (maxVal - val > 0) ? maxVal : val;
  }
  return {minVal, maxVal};
}
// Collision.ConvexConvex.project over Convex.convexVertices (tests/Collision/ConvexConvexTest.elm:256-319)
void projectVertices(V3 axis, const WConvex& c, double* out2) {
  MinMax m = project(axis, convexVertices(c));
  out2[0] = m.min; out2[1] = m.max;
}
// ConvexConvex.elm:596-610
MinMax projectConvex(V3 axis, const WConvex& c) {
  if (c.isBox) {
    double cc = dot(axis, c.position);
    double e = eabs(dot(axis, c.ax)) * c.he.x + eabs(dot(axis, c.ay)) * c.he.y + eabs(dot(axis, c.az)) * c.he.z;
    return {cc - e, cc + e};
  }
  return project(axis, c.vs);
}
// ConvexConvex.elm:646-662 — returns false for Nothing
static bool overlap(MinMax p1, MinMax p2, double& depth) {
  double d1 = p1.max - p2.min, d2 = p2.max - p1.min;
  if (d1 + kContactBreakingThreshold < 0 || d2 + kContactBreakingThreshold < 0) return false;
  depth = (d1 - d2 > 0) ? d2 : d1;
  return true;
}
// ConvexConvex.elm:669-680
static MinMax faceExtent(V3 axis, const WGroup& g, const WConvex& c) {
  if (g.twoSided) { double posDot = dot(axis, c.position); return {g.d2 + posDot, g.d1 + posDot}; }
  return project(axis, convexVertices(c));
}
bool testSeparatingAxis(const WConvex& c1, const WConvex& c2, V3 axis, double& depth) {
  return overlap(projectConvex(axis, c1), projectConvex(axis, c2), depth);
}
// ConvexConvex.elm:135-141
static V3 orientAxis(const WConvex& c1, const WConvex& c2, V3 axis) {
  return dot(sub(c2.position, c1.position), axis) > 0 ? neg(axis) : axis;
}

struct FaceWinner { V3 axis; double dmin; int fromSide; int groupIdx; const WGroup* group; };
// ConvexConvex.elm:420-479
static bool findFaceSAT(const WConvex& c1, const WConvex& c2, FaceWinner& w) {
  int winnerIdx = -1, winnerSide = 1; const WGroup* winnerGroup = nullptr; double dmin = kMaxNumber;
  for (int side = 1; side <= 2; side++) {
    const WConvex& own = side == 1 ? c1 : c2;
    int nextIdx = 1;
    for (const WGroup& g : own.faces) {
      V3 axis = g.n1; double dist;
      bool ok = side == 1 ? overlap(faceExtent(axis, g, c1), projectConvex(axis, c2), dist)
                          : overlap(projectConvex(axis, c1), faceExtent(axis, g, c2), dist);
      if (!ok) return false;
      if (dist - dmin < 0) { winnerIdx = nextIdx; winnerSide = side; winnerGroup = &g; dmin = dist; }
      nextIdx += g.twoSided ? 2 : 1;
    }
  }
  if (winnerIdx == -1) return false;
  w = FaceWinner{winnerGroup->n1, dmin, winnerSide, winnerIdx, winnerGroup};
  return true;
}

enum EdgeKind { kEdgeSeparates, kEdgeBeats, kNoEdgeBeats };
struct EdgeResult { EdgeKind kind; V3 axis; int dir1, dir2; const int32_t* e1; int n1; const int32_t* e2; int n2; };
// ConvexConvex.elm:517-580
static EdgeResult findEdgeSAT(const WConvex& c1, const WConvex& c2, double faceDmin) {
  EdgeResult best{kNoEdgeBeats, {0, 0, 0}, 0, 0, nullptr, 0, nullptr, 0};
  double dmin = faceDmin / 1.05;
  for (size_t a = 0; a < c1.edges.size(); a++) {
    if (c1.edges[a].second < 2) continue;
    for (size_t b = 0; b < c2.edges.size(); b++) {
      if (c2.edges[b].second < 2) continue;
      V3 dir1 = direction(c1.buffer[c1.edges[a].first[0]], c1.buffer[c1.edges[a].first[1]]);
      V3 dir2 = direction(c2.buffer[c2.edges[b].first[0]], c2.buffer[c2.edges[b].first[1]]);
      V3 cr = cross(dir1, dir2);
      if (almostZero(cr)) continue;
      V3 n = normalize(cr);
      double dist;
      if (!testSeparatingAxis(c1, c2, n, dist)) { best.kind = kEdgeSeparates; return best; }
      if (dist - dmin < 0) {
        best = EdgeResult{kEdgeBeats, n, (int)a + 1, (int)b + 1, c1.edges[a].first, c1.edges[a].second, c2.edges[b].first, c2.edges[b].second};
        dmin = dist;
      }
    }
  }
  return best;
}

// ConvexConvex.elm:175-201
static int pickSupportEdge(V3 dir, const int32_t* edges, int n, const std::vector<V3>& buf, V3& p, V3& q) {
  int bestIdx = 0; p = V3{0, 0, 0}; q = V3{0, 0, 0}; double bestDot = -kMaxNumber;
  int idx = 1;
  for (int k = 0; k + 1 < n; k += 2, idx++) {
    V3 v1 = buf[edges[k]], v2 = buf[edges[k + 1]];
    double midDot = dir.x * (v1.x + v2.x) + dir.y * (v1.y + v2.y) + dir.z * (v1.z + v2.z);
    if (midDot - bestDot > 0) { bestIdx = idx; p = v1; q = v2; bestDot = midDot; }
  }
  return bestIdx;
}

struct Face { V3 normal; const int32_t* idx; int n; };
// ConvexConvex.elm:276-324 — separatingAxis argument as given by the caller
int bestFace(const std::vector<WGroup>& groups, V3 sepAxis, Face& out) {
  int faceId = 1, bestId = -1; double bestDist = kMaxNumber; out = Face{{0, 0, 0}, nullptr, 0};
  for (const WGroup& g : groups) {
    if (g.twoSided) {
      double primaryDot = dot(g.n1, sepAxis), partnerDot = -primaryDot;
      int id1 = bestId; Face f1 = out; double d1 = bestDist;
      if (bestDist - primaryDot > 0) { id1 = faceId; f1 = Face{g.n1, g.i1, g.n_i1}; d1 = primaryDot; }
      if (d1 - partnerDot > 0) { bestId = faceId + 1; out = Face{g.n2, g.i2, g.n_i2}; bestDist = partnerDot; }
      else { bestId = id1; out = f1; bestDist = d1; }
      faceId += 2;
    } else {
      double d = dot(g.n1, sepAxis);
      if (bestDist - d > 0) { bestId = faceId; out = Face{g.n1, g.i1, g.n_i1}; bestDist = d; }
      faceId += 1;
    }
  }
  return bestId;
}

// ConvexConvex.elm:349-395 ; polygons in FORWARD list order
static TagPoint crossing(double nDotPrev, double nDotNext, const TagPoint& prev, const TagPoint& next) {
  double t = nDotPrev / (nDotPrev - nDotNext);
  return TagPoint{t < 0.5 ? prev.tag : next.tag, lerp(t, prev.p, next.p)};
}
static PL clipPolygon(V3 planeNormal, double planeConstant, const PL& poly) {
  PL result;   // built by cons: keep as forward list => insert at front
  if (poly.size() < 2) return result;
  size_t n = poly.size();
  for (size_t i = 0; i < n; i++) {
    const TagPoint& prev = poly[i];
    const TagPoint& next = poly[(i + 1) % n];
    double nDotPrev = dot(planeNormal, prev.p) + planeConstant;
    double nDotNext = dot(planeNormal, next.p) + planeConstant;
    if (nDotPrev < 0) {
      if (nDotNext < 0) result.insert(result.begin(), next);
      else result.insert(result.begin(), crossing(nDotPrev, nDotNext, prev, next));
    } else if (nDotNext < 0) {
      result.insert(result.begin(), crossing(nDotPrev, nDotNext, prev, next));
      result.insert(result.begin(), next);
    }
  }
  return result;
}
// ConvexConvex.elm:327-346
static PL clipAgainstAdjacentFaces(V3 normal, const std::vector<V3>& ref, PL poly) {
  if (ref.size() < 2) return poly;
  size_t n = ref.size();
  for (size_t i = 0; i < n; i++) {
    V3 v1 = ref[i], v2 = ref[(i + 1) % n];
    V3 edge = normalize(sub(v1, v2));
    V3 planeNormal = cross(normal, edge);
    double planeConstant = -(dot(v1, planeNormal));
    poly = clipPolygon(planeNormal, planeConstant, poly);
  }
  return poly;
}
// ConvexConvex.elm:204-269
static void clipTwoFaces(int shapeKey, int faceId1, int faceId2, const Face& face, const std::vector<V3>& buf1,
                         const Face& incident, const std::vector<V3>& buf2, V3 sepAxis, CL& out) {
  std::vector<V3> ref; for (int i = 0; i < face.n; i++) ref.push_back(buf1[face.idx[i]]);
  PL poly; for (int i = 0; i < incident.n; i++) poly.push_back(TagPoint{incident.idx[i], buf2[incident.idx[i]]});
  V3 point = ref.empty() ? V3{0, 0, 0} : ref[0];
  double fpc = -(dot(face.normal, point));
  PL red = manifoldReduce(face.normal, fpc, clipAgainstAdjacentFaces(face.normal, ref, poly));
  for (const TagPoint& tp : red) {
    double depth = dot(face.normal, tp.p) + fpc;
    out.push_back(Contact{shapeKey, convexConvexFace(faceId1, faceId2, tp.tag), sepAxis,
                          V3{tp.p.x - depth * face.normal.x, tp.p.y - depth * face.normal.y, tp.p.z - depth * face.normal.z}, tp.p});
  }
}
// ConvexConvex.elm:117-132
static int pickWinningFace(int groupIdx, const WGroup& g, V3 axisToward, Face& out) {
  if (g.twoSided) {
    if (dot(g.n1, axisToward) <= 0) { out = Face{g.n1, g.i1, g.n_i1}; return groupIdx; }
    out = Face{g.n2, g.i2, g.n_i2}; return groupIdx + 1;
  }
  out = Face{g.n1, g.i1, g.n_i1}; return groupIdx;
}
// ConvexConvex.elm:40-114, 148-169
static void convexConvex(int shapeKey, const WConvex& c1, const WConvex& c2, CL& out) {
  FaceWinner w;
  if (!findFaceSAT(c1, c2, w)) return;
  EdgeResult er = findEdgeSAT(c1, c2, w.dmin);
  if (er.kind == kEdgeSeparates) return;
  if (er.kind == kEdgeBeats) {
    V3 sepAxis = orientAxis(c1, c2, er.axis);
    V3 rev = neg(sepAxis);
    V3 e1p, e1q, e2p, e2q;
    int edge1 = pickSupportEdge(rev, er.e1, er.n1, c1.buffer, e1p, e1q);
    int edge2 = pickSupportEdge(sepAxis, er.e2, er.n2, c2.buffer, e2p, e2q);
    V3 pi, pj; closestPointsBetweenSegments(e1p, e1q, e2p, e2q, pi, pj);
    out.push_back(Contact{shapeKey, convexConvexEdge(er.dir1, edge1, er.dir2, edge2), rev, pi, pj});
    return;
  }
  V3 sepAxis = orientAxis(c1, c2, w.axis);
  V3 rev = neg(sepAxis);
  int id1, id2; Face f1, f2;
  if (w.fromSide == 1) { id1 = pickWinningFace(w.groupIdx, *w.group, sepAxis, f1); id2 = bestFace(c2.faces, rev, f2); }
  else { id1 = bestFace(c1.faces, sepAxis, f1); id2 = pickWinningFace(w.groupIdx, *w.group, rev, f2); }
  if (id1 == -1 || id2 == -1) return;
  clipTwoFaces(shapeKey, id1, id2, f1, c1.buffer, f2, c2.buffer, rev, out);
}

// exported test hooks (tests/Collision/ConvexConvexTest.elm:163-253)
bool findSeparatingAxis(const WConvex& c1, const WConvex& c2, V3& axis) {
  FaceWinner w;
  if (!findFaceSAT(c1, c2, w)) return false;
  EdgeResult er = findEdgeSAT(c1, c2, w.dmin);
  if (er.kind == kEdgeSeparates) return false;
  axis = orientAxis(c1, c2, er.kind == kEdgeBeats ? er.axis : w.axis);
  return true;
}

// ---- CapsuleConvex.elm --------------------------------------------------------------------------
// CapsuleConvex.elm:718-749
static bool testCapsuleConvexAxis(const WShape& cap, const WConvex& cv, V3 n, double& depth) {
  double centerProj = dot(cap.position, n);
  double axisContrib = eabs(dot(cap.axis, n)) * cap.halfLength;
  double capsuleMin = centerProj - axisContrib - cap.radius;
  double capsuleMax = centerProj + axisContrib + cap.radius;
  MinMax p2 = projectConvex(n, cv);
  double d1 = capsuleMax - p2.min, d2 = p2.max - capsuleMin;
  if (d1 + kContactBreakingThreshold < 0 || d2 + kContactBreakingThreshold < 0) return false;
  depth = (d1 - d2 > 0) ? d2 : d1;
  return true;
}
// CapsuleConvex.elm:575-599
static bool edgeAxis(const WShape& cap, V3 ep1, V3 ep2, V3 v1, V3 v2, V3& axis) {
  V3 pCap, pEdge; closestPointsBetweenSegments(ep1, ep2, v1, v2, pCap, pEdge);
  V3 diff = sub(pCap, pEdge);
  double distSq = lengthSquared(diff);
  if (distSq - kPrecision > 0) { axis = scale(1 / std::sqrt(distSq), diff); return true; }
  V3 cr = cross(cap.axis, sub(v2, v1));
  if (almostZero(cr)) return false;
  axis = normalize(cr);
  return true;
}
// CapsuleConvex.elm:494-572
static bool capsuleFindSeparatingAxis(const WShape& cap, const WConvex& cv, V3 ep1, V3 ep2, V3& axis) {
  V3 target{0, 0, 0}; double dmin = kMaxNumber;
  for (const WGroup& g : cv.faces) {
    double dist;
    if (!testCapsuleConvexAxis(cap, cv, g.n1, dist)) return false;
    if (dist - dmin < 0) { target = g.n1; dmin = dist; }
  }
  for (auto& grp : cv.edges) {
    for (int k = 0; k + 1 < grp.second; k += 2) {
      V3 v1 = cv.buffer[grp.first[k]], v2 = cv.buffer[grp.first[k + 1]];
      V3 ax;
      if (!edgeAxis(cap, ep1, ep2, v1, v2, ax)) continue;
      double dist;
      if (!testCapsuleConvexAxis(cap, cv, ax, dist)) return false;
      if (dist - dmin < 0) { target = ax; dmin = dist; }
    }
  }
  axis = dot(sub(cv.position, cap.position), target) > 0 ? neg(target) : target;
  return true;
}
// CapsuleConvex.elm:607-716 ; returns 0, 1 or 2 vertices
std::vector<V3> supportFeature(V3 axis, const WConvex& cv) {
  std::vector<V3> vertices = convexVertices(cv);
  double maxProj = -kMaxNumber;
  for (V3 v : vertices) { double p = dot(v, axis); maxProj = (p - maxProj > 0) ? p : maxProj; }
  int count = 0; V3 v1{0, 0, 0}; bool two = false;
  for (V3 v : vertices) {
    if (maxProj - dot(v, axis) - 1.0e-4 < 0) { if (count == 0) { count = 1; v1 = v; } else { two = true; break; } }
  }
  if (count == 0) return {};
  if (!two) return {v1};
  for (auto& grp : cv.edges)
    for (int k = 0; k + 1 < grp.second; k += 2) {
      V3 a = cv.buffer[grp.first[k]], b = cv.buffer[grp.first[k + 1]];
      if ((maxProj - dot(a, axis) - 1.0e-4 < 0) && (maxProj - dot(b, axis) - 1.0e-4 < 0)) return {a, b};
    }
  return {v1};
}
// CapsuleConvex.elm:237-259
static void addDirectContact(int shapeKey, int64_t fk, bool flip, V3 sp, V3 sepAxis, double pen, const WShape& cap, CL& out) {
  V3 normal = neg(sepAxis);
  V3 pi{sp.x + cap.radius * normal.x, sp.y + cap.radius * normal.y, sp.z + cap.radius * normal.z};
  V3 pj{pi.x - pen * normal.x, pi.y - pen * normal.y, pi.z - pen * normal.z};
  out.push_back(orderContact(flip, Contact{shapeKey, fk, normal, pi, pj}));
}
// CapsuleConvex.elm:333-388 ; returns false for ClipDead
static bool clipSegmentAgainstFace(const std::vector<V3>& buf, const Face& face, V3 ep1, V3 ep2, V3& o1, V3& o2) {
  o1 = ep1; o2 = ep2;
  if (face.n < 2) return true;
  for (int k = 0; k < face.n; k++) {
    V3 v1 = buf[face.idx[k]];
    V3 v2 = (k + 1 < face.n) ? buf[face.idx[k + 1]] : buf[face.idx[0]];
    V3 edge = normalize(sub(v1, v2));
    V3 planeNormal = cross(face.normal, edge);
    double planeConstant = -(dot(v1, planeNormal));
    double d1 = dot(planeNormal, o1) + planeConstant, d2 = dot(planeNormal, o2) + planeConstant;
    if (d1 < 0 && d2 < 0) continue;
    else if (d1 >= 0 && d2 >= 0) return false;
    else if (d1 < 0) o2 = lerp(d1 / (d1 - d2), o1, o2);
    else o1 = lerp(d1 / (d1 - d2), o1, o2);
  }
  return true;
}
// CapsuleConvex.elm:262-312
static void addBodyContacts(int shapeKey, int cf, bool flip, const std::vector<V3>& buf, bool haveFace, const Face& face,
                            const WShape& cap, V3 ep1, V3 ep2, V3 sepAxis, V3 capPoint, CL& out) {
  if (!haveFace) return;
  V3 q1, q2;
  if (!clipSegmentAgainstFace(buf, face, ep1, ep2, q1, q2)) return;
  double fpc = face.n > 0 ? -(dot(face.normal, buf[face.idx[0]])) : 0;
  V3 qs[2] = {q1, q2};
  for (int k = 0; k < 2; k++) {
    V3 point = qs[k];
    if (distance(point, capPoint) < kPrecision) continue;
    double signedDist = dot(face.normal, point) + fpc;
    double bodyPen = cap.radius - signedDist;
    if (bodyPen <= 0) continue;
    addDirectContact(shapeKey, k == 0 ? capsuleCylinder1(cf) : capsuleCylinder2(cf), flip, point, sepAxis, bodyPen, cap, out);
  }
}
// CapsuleConvex.elm:58-234, 395-491
static void capsuleConvex(int shapeKey, bool flip, const WShape& cap, const WConvex& cv, CL& out) {
  V3 ep1{cap.position.x - cap.halfLength * cap.axis.x, cap.position.y - cap.halfLength * cap.axis.y, cap.position.z - cap.halfLength * cap.axis.z};
  V3 ep2{cap.position.x + cap.halfLength * cap.axis.x, cap.position.y + cap.halfLength * cap.axis.y, cap.position.z + cap.halfLength * cap.axis.z};
  V3 sepAxis;
  if (!capsuleFindSeparatingAxis(cap, cv, ep1, ep2, sepAxis)) return;
  double penetration;
  if (!testCapsuleConvexAxis(cap, cv, sepAxis, penetration)) return;
  double t = dot(cap.axis, sepAxis);
  Face f; int fid = bestFace(cv.faces, neg(sepAxis), f);
  bool haveFace = (fid != -1) && (dot(f.normal, sepAxis) - 0.999 > 0);
  std::vector<V3> supportVerts;
  if (!haveFace) supportVerts = supportFeature(sepAxis, cv);
  int cf = haveFace ? onConvexFace(fid) : (supportVerts.size() >= 2 ? onConvexEdge : (supportVerts.size() == 1 ? onConvexVertex : onConvexNone));
  if (t - 1.0e-3 > 0) {
    addDirectContact(shapeKey, capsuleCapEnd1(cf), flip, ep1, sepAxis, penetration, cap, out);
    addBodyContacts(shapeKey, cf, flip, cv.buffer, haveFace, f, cap, ep1, ep2, sepAxis, ep1, out);
  } else if (t + 1.0e-3 < 0) {
    addDirectContact(shapeKey, capsuleCapEnd2(cf), flip, ep2, sepAxis, penetration, cap, out);
    addBodyContacts(shapeKey, cf, flip, cv.buffer, haveFace, f, cap, ep1, ep2, sepAxis, ep2, out);
  } else if (haveFace) {
    V3 p1, p2;
    if (clipSegmentAgainstFace(cv.buffer, f, ep1, ep2, p1, p2)) {
      addDirectContact(shapeKey, capsuleCylinder1(cf), flip, p1, sepAxis, penetration, cap, out);
      addDirectContact(shapeKey, capsuleCylinder2(cf), flip, p2, sepAxis, penetration, cap, out);
    } else {
      // addClosestEdgeContact, CapsuleConvex.elm:395-491
      double radiusSq = cap.radius * cap.radius;
      V3 bestPCap{0, 0, 0}, bestPEdge{0, 0, 0}; double bestDistSq = radiusSq;
      if (f.n >= 2) {
        for (int k = 0; k < f.n; k++) {
          V3 v1 = cv.buffer[f.idx[k]];
          V3 v2 = (k + 1 < f.n) ? cv.buffer[f.idx[k + 1]] : cv.buffer[f.idx[0]];
          V3 pCap, pEdge; closestPointsBetweenSegments(ep1, ep2, v1, v2, pCap, pEdge);
          double dx = pCap.x - pEdge.x, dy = pCap.y - pEdge.y, dz = pCap.z - pEdge.z;
          double distSq = dx * dx + dy * dy + dz * dz;
          if (distSq - bestDistSq < 0) { bestPCap = pCap; bestPEdge = pEdge; bestDistSq = distSq; }
        }
      }
      if (bestDistSq - radiusSq >= 0) return;
      double dist = std::sqrt(bestDistSq);
      if (dist - kPrecision <= 0) return;
      double inv = 1 / dist;
      V3 esa{(bestPCap.x - bestPEdge.x) * inv, (bestPCap.y - bestPEdge.y) * inv, (bestPCap.z - bestPEdge.z) * inv};
      addDirectContact(shapeKey, capsuleCylinder(cf), flip, bestPCap, esa, cap.radius - dist, cap, out);
    }
  } else {
    if (supportVerts.size() >= 2) {
      V3 v1 = supportVerts[0], v2 = supportVerts[1];
      if (almostZero(cross(cap.axis, sub(v2, v1)))) {
        // addParallelEdgeContacts, CapsuleConvex.elm:197-234
        double sEp1 = dot(ep1, cap.axis), sEp2 = dot(ep2, cap.axis), sV1 = dot(v1, cap.axis), sV2 = dot(v2, cap.axis);
        double sLo = emax(emin(sEp1, sEp2), emin(sV1, sV2));
        double sHi = emin(emax(sEp1, sEp2), emax(sV1, sV2));
        auto pointAt = [&](double s) { return V3{ep1.x + (s - sEp1) * cap.axis.x, ep1.y + (s - sEp1) * cap.axis.y, ep1.z + (s - sEp1) * cap.axis.z}; };
        if (sHi - sLo - kPrecision > 0) {
          addDirectContact(shapeKey, capsuleCylinder1(cf), flip, pointAt(sLo), sepAxis, penetration, cap, out);
          addDirectContact(shapeKey, capsuleCylinder2(cf), flip, pointAt(sHi), sepAxis, penetration, cap, out);
        } else {
          V3 pc, pe; closestPointsBetweenSegments(ep1, ep2, v1, v2, pc, pe);
          addDirectContact(shapeKey, capsuleCylinder(cf), flip, pc, sepAxis, penetration, cap, out);
        }
      } else {
        V3 pc, pe; closestPointsBetweenSegments(ep1, ep2, v1, v2, pc, pe);
        addDirectContact(shapeKey, capsuleCylinder(cf), flip, pc, sepAxis, penetration, cap, out);
      }
    } else if (supportVerts.size() == 1) {
      V3 v = supportVerts[0];
      V3 pc, pe; closestPointsBetweenSegments(ep1, ep2, v, v, pc, pe);
      addDirectContact(shapeKey, capsuleCylinder(cf), flip, pc, sepAxis, penetration, cap, out);
    }
  }
}

// ---- NarrowPhase.elm:86-286 : emission-order contacts of one shape pair -----------------------------
void addRawShapeContacts(int shapeKey, const WShape& s1, const WShape& s2, std::vector<Contact>& out) {
  switch (s1.kind) {
    case EP_CONVEX:
      switch (s2.kind) {
        case EP_CONVEX: convexConvex(shapeKey, s1.convex, s2.convex, out); break;
        case EP_PLANE: planeConvex(shapeKey, true, s2, s1.convex, out); break;
        case EP_SPHERE: sphereConvex(shapeKey, true, s2, s1.convex, out); break;
        case EP_PARTICLE: particleConvex(shapeKey, true, s2.position, s1.convex, out); break;
        case EP_CAPSULE: capsuleConvex(shapeKey, true, s2, s1.convex, out); break;
      }
      break;
    case EP_PLANE:
      switch (s2.kind) {
        case EP_PLANE: break;
        case EP_CONVEX: planeConvex(shapeKey, false, s1, s2.convex, out); break;
        case EP_SPHERE: planeSphere(shapeKey, false, s1, s2, out); break;
        case EP_PARTICLE: planeParticle(shapeKey, false, s1, s2.position, out); break;
        case EP_CAPSULE: planeCapsule(shapeKey, false, s1, s2, out); break;
      }
      break;
    case EP_SPHERE:
      switch (s2.kind) {
        case EP_PLANE: planeSphere(shapeKey, true, s2, s1, out); break;
        case EP_CONVEX: sphereConvex(shapeKey, false, s1, s2.convex, out); break;
        case EP_SPHERE: sphereSphere(shapeKey, s1, s2, out); break;
        case EP_PARTICLE: sphereParticle(shapeKey, false, s1, s2.position, out); break;
        case EP_CAPSULE: capsuleSphere(shapeKey, true, s2, s1, out); break;
      }
      break;
    case EP_PARTICLE:
      switch (s2.kind) {
        case EP_PLANE: planeParticle(shapeKey, true, s2, s1.position, out); break;
        case EP_CONVEX: particleConvex(shapeKey, false, s1.position, s2.convex, out); break;
        case EP_SPHERE: sphereParticle(shapeKey, true, s2, s1.position, out); break;
        case EP_PARTICLE: break;
        case EP_CAPSULE: capsuleParticle(shapeKey, true, s2, s1.position, out); break;
      }
      break;
    case EP_CAPSULE:
      switch (s2.kind) {
        case EP_PLANE: planeCapsule(shapeKey, true, s2, s1, out); break;
        case EP_CONVEX: capsuleConvex(shapeKey, false, s1, s2.convex, out); break;
        case EP_SPHERE: capsuleSphere(shapeKey, false, s1, s2, out); break;
        case EP_CAPSULE: capsuleCapsule(shapeKey, s1, s2, out); break;
        case EP_PARTICLE: capsuleParticle(shapeKey, false, s1, s2.position, out); break;
      }
      break;
  }
}

// NarrowPhase.elm:23-83 — contacts of a body pair in SOLVER order (reverse of pairGroup.contacts):
// shape pairs in (s1 outer, s2 inner) order, each shape pair's contacts in reverse emission order.
void getContactsSolverOrder(const std::vector<WShape>& shapes1, const std::vector<WShape>& shapes2, std::vector<SolverContact>& out) {
  for (size_t a = 0; a < shapes1.size(); a++)
    for (size_t b = 0; b < shapes2.size(); b++) {
      std::vector<Contact> raw;
      addRawShapeContacts((int)(a + 1) * 256 + (int)(b + 1), shapes1[a], shapes2[b], raw);
      if (raw.empty()) continue;
      // Material.elm:23-33
      double bounciness = (shapes1[a].bounciness + shapes2[b].bounciness + eabs(shapes1[a].bounciness - shapes2[b].bounciness)) * 0.5;
      double friction = std::sqrt(shapes1[a].friction * shapes2[b].friction);
      for (int k = (int)raw.size() - 1; k >= 0; k--) out.push_back(SolverContact{friction, bounciness, raw[k]});
    }
}

}  // namespace epo
