// ORACLE — TEST INFRASTRUCTURE ONLY (see ep_oracle_math.h).
// CPU restatement of Physics.raycast (src/Physics.elm:802-841): Body.raycast (src/Internal/Body.elm:465-489) over
// Shape.raycast (src/Internal/Shape.elm:132-148) = Plane.raycast (src/Shapes/Plane.elm:20-49), Sphere.raycast
// (src/Shapes/Sphere.elm:49-86), Convex.raycast (src/Shapes/Convex.elm:1117-1212, foldFaceEdges 1215-1245),
// Capsule.raycast (src/Shapes/Capsule.elm:65-204).  Point masses are never hit.
#include <vector>

#include "ep_oracle.h"
#include "ep_oracle_internal.h"

namespace epo {

struct Hit { bool some; double distance; V3 point, normal; };

static Hit planeRaycast(V3 from, V3 direction, const WShape& s) {           // Plane.elm:20-49
  double d = dot(direction, s.normal);
  if (d < 0) {
    V3 pointToFrom = sub(s.position, from);
    double scalar = dot(s.normal, pointToFrom) / d;
    if (scalar >= 0)
      return {true, scalar, {direction.x * scalar + from.x, direction.y * scalar + from.y, direction.z * scalar + from.z}, s.normal};
  }
  return {false, 0, {0, 0, 0}, {0, 0, 0}};
}

static Hit sphereRaycast(V3 from, V3 direction, const WShape& s) {          // Sphere.elm:49-86
  V3 p = s.position; double radius = s.radius;
  double b = 2 * (direction.x * (from.x - p.x) + direction.y * (from.y - p.y) + direction.z * (from.z - p.z));
  double c = (from.x - p.x) * (from.x - p.x) + (from.y - p.y) * (from.y - p.y) + (from.z - p.z) * (from.z - p.z) - radius * radius;
  double delta = b * b - 4 * c;
  if (delta < 0) return {false, 0, {0, 0, 0}, {0, 0, 0}};
  double distance = (-b - sqrt(delta)) / 2;
  if (distance >= 0) {
    V3 point = {from.x + direction.x * distance, from.y + direction.y * distance, from.z + direction.z * distance};
    return {true, distance, point, sub(point, p)};      // normalized at the top level
  }
  return {false, 0, {0, 0, 0}, {0, 0, 0}};
}

static Hit raycastFace(const WConvex& c, V3 direction, V3 from, V3 normal, const int32_t* idx, int n, Hit maybeHit) {   // Convex.elm:1138-1212
  double d = dot(direction, normal);
  V3 point = n > 0 ? c.buffer[idx[0]] : V3{0, 0, 0};
  if (d < 0) {
    V3 pointToFrom = sub(point, from);
    double scalar = dot(normal, pointToFrom) / d;
    if (scalar >= 0) {
      V3 ip = {direction.x * scalar + from.x, direction.y * scalar + from.y, direction.z * scalar + from.z};
      bool inside = true;
      if (n >= 2)                                            // foldFaceEdges: (v0,v1) ... (v_last, v0)
        for (int i = 0; i < n; i++) {
          V3 p1 = c.buffer[idx[i]], p2 = c.buffer[idx[i + 1 < n ? i + 1 : 0]];
          inside = inside && (dot(sub(ip, p1), cross(normal, sub(p2, p1))) > 0);
        }
      if (inside) {
        if (maybeHit.some) { if (scalar - maybeHit.distance < 0) return {true, scalar, ip, normal}; return maybeHit; }
        return {true, scalar, ip, normal};
      }
    }
  }
  return maybeHit;
}

static Hit convexRaycast(V3 from, V3 direction, const WConvex& c) {         // Convex.elm:1117-1135
  Hit h = {false, 0, {0, 0, 0}, {0, 0, 0}};
  for (const WGroup& g : c.faces) {
    h = raycastFace(c, direction, from, g.n1, g.i1, g.n_i1, h);
    if (g.twoSided) h = raycastFace(c, direction, from, g.n2, g.i2, g.n_i2, h);
  }
  return h;
}

static bool capEntry(V3 p, V3 direction, V3 axis, double radius, double capOffset, double& t) {   // Capsule.elm:108-137
  double qx = p.x - capOffset * axis.x, qy = p.y - capOffset * axis.y, qz = p.z - capOffset * axis.z;
  double b = 2 * (direction.x * qx + direction.y * qy + direction.z * qz);
  double c = qx * qx + qy * qy + qz * qz - radius * radius;
  double delta = b * b - 4 * c;
  if (delta < 0) return false;
  t = (-b - sqrt(delta)) / 2;
  return !(t < 0);
}

static Hit capsuleRaycast(V3 from, V3 direction, const WShape& s) {         // Capsule.elm:65-204
  const double radius = s.radius, halfLength = s.halfLength; const V3 axis = s.axis, position = s.position;
  V3 p = sub(from, position);
  double pDotAxis = dot(p, axis), dDotAxis = dot(direction, axis);
  double pPerpX = p.x - pDotAxis * axis.x, pPerpY = p.y - pDotAxis * axis.y, pPerpZ = p.z - pDotAxis * axis.z;
  double dPerpX = direction.x - dDotAxis * axis.x, dPerpY = direction.y - dDotAxis * axis.y, dPerpZ = direction.z - dDotAxis * axis.z;
  double cylA = dPerpX * dPerpX + dPerpY * dPerpY + dPerpZ * dPerpZ;
  double cylB = 2 * (pPerpX * dPerpX + pPerpY * dPerpY + pPerpZ * dPerpZ);
  double cylC = pPerpX * pPerpX + pPerpY * pPerpY + pPerpZ * pPerpZ - radius * radius;
  bool some = false; double t = 0;
  if (cylA - 1.0e-10 < 0) {                                 // rayParallelTolerance, Capsule.elm:17-19
    if (cylC > 0) some = false;
    else if (dDotAxis < 0) some = capEntry(p, direction, axis, radius, halfLength, t);
    else some = capEntry(p, direction, axis, radius, -halfLength, t);
  } else {
    double delta = cylB * cylB - 4 * cylA * cylC;
    if (delta < 0) some = false;
    else {
      double tt = (-cylB - sqrt(delta)) / (2 * cylA);
      if (tt < 0) some = false;
      else {
        double axial = pDotAxis + tt * dDotAxis;
        if (axial - halfLength > 0) some = capEntry(p, direction, axis, radius, halfLength, t);
        else if (axial + halfLength < 0) some = capEntry(p, direction, axis, radius, -halfLength, t);
        else { some = true; t = tt; }
      }
    }
  }
  if (!some) return {false, 0, {0, 0, 0}, {0, 0, 0}};
  V3 hitPoint = {from.x + direction.x * t, from.y + direction.y * t, from.z + direction.z * t};
  double x = pDotAxis + t * dDotAxis;
  double clampedAxial = (x < -halfLength) ? -halfLength : ((x > halfLength) ? halfLength : x);   // Basics.clamp
  return {true, t, hitPoint, {hitPoint.x - position.x - clampedAxial * axis.x, hitPoint.y - position.y - clampedAxial * axis.y,
                              hitPoint.z - position.z - clampedAxial * axis.z}};
}

static Hit shapeRaycast(V3 from, V3 direction, const WShape& s) {           // Internal/Shape.elm:132-148
  switch (s.kind) {
    case EP_PLANE: return planeRaycast(from, direction, s);
    case EP_SPHERE: return sphereRaycast(from, direction, s);
    case EP_CONVEX: return convexRaycast(from, direction, s.convex);
    case EP_CAPSULE: return capsuleRaycast(from, direction, s);
    default: return {false, 0, {0, 0, 0}, {0, 0, 0}};
  }
}

}  // namespace epo

using namespace epo;

// Physics.raycast for every ray: ray r runs against the bodies of world ray_world[r] in list order.
extern "C" int ep_oracle_raycast(const ep_shapes* S, const ep_bodies* B, int32_t n_rays, const int32_t* ray_world, const double* from,
                                 const double* direction, int32_t* hit_body, double* hit_distance, double* hit_point, double* hit_normal) {
  // worldShapesWithMaterials of every body, placed from its current transform (Body.elm:253, SolverBody.elm:212)
  std::vector<std::vector<WShape>> world(B->n_bodies);
  for (int r = 0; r < B->n_bodies; r++) {
    Tf tf{{B->pos[3 * r], B->pos[3 * r + 1], B->pos[3 * r + 2]}, {B->quat[4 * r], B->quat[4 * r + 1], B->quat[4 * r + 2], B->quat[4 * r + 3]}};
    for (int k = B->inst_off[r]; k < B->inst_off[r + 1]; k++) world[r].push_back(placeShape(S, B->inst_shape[k], tf, 0, 0));
  }
  for (int i = 0; i < n_rays; i++) {
    int w = ray_world[i];
    if (w < 0 || w >= B->n_worlds) return EP_EINVAL;
    V3 f = rd3(from + 3 * i), dir = rd3(direction + 3 * i);
    bool any = false; int bestBody = -1; Hit best{false, 0, {0, 0, 0}, {0, 0, 0}};
    for (int r = B->world_body_off[w]; r < B->world_body_off[w + 1]; r++) {
      Hit bodyHit{false, 0, {0, 0, 0}, {0, 0, 0}};                   // Body.raycast: closest shape, first wins ties
      for (const WShape& s : world[r]) {
        Hit h = shapeRaycast(f, dir, s);
        if (h.some && (!bodyHit.some || h.distance - bodyHit.distance < 0)) bodyHit = h;
      }
      if (bodyHit.some && (!any || bodyHit.distance - best.distance < 0)) { any = true; best = bodyHit; bestBody = r; }
    }
    hit_body[i] = any ? bestBody : -1;
    V3 n = any ? normalize(best.normal) : V3{0, 0, 0};               // Physics.elm:838
    hit_distance[i] = any ? best.distance : 0;
    hit_point[3 * i] = any ? best.point.x : 0; hit_point[3 * i + 1] = any ? best.point.y : 0; hit_point[3 * i + 2] = any ? best.point.z : 0;
    hit_normal[3 * i] = n.x; hit_normal[3 * i + 1] = n.y; hit_normal[3 * i + 2] = n.z;
  }
  return EP_OK;
}

// Between-step body mutators (src/Physics.elm:857-1017 over src/Internal/Body.elm:305-430), operations applied in list order.
extern "C" int ep_oracle_apply(ep_bodies* B, int32_t n_ops, const int32_t* kind, const int32_t* body_row, const double* a, const double* b) {
  for (int k = 0; k < n_ops; k++) {
    int r = body_row[k];
    if (r < 0 || r >= B->n_bodies) return EP_EINVAL;
    int bk = B->kind[r];
    V3 va = rd3(a + 3 * k), vb = rd3(b + 3 * k);
    double* v = B->vel + 3 * r; double* w = B->angvel + 3 * r; double* f = B->force + 3 * r; double* tq = B->torque + 3 * r;
    const double* m = B->inv_inertia_world + 9 * r;      // m11,m21,m31,m12,m22,m32,m13,m23,m33
    V3 x = rd3(B->pos + 3 * r);
    double im = B->inv_mass[r];
    switch (kind[k]) {
      case EP_APPLY_FORCE:            // Body.elm:341-368, dynamic only (Physics.elm:866-878)
        if (bk == 2) { V3 t = cross(sub(vb, x), va); f[0] = f[0] + va.x; f[1] = f[1] + va.y; f[2] = f[2] + va.z; tq[0] = tq[0] + t.x; tq[1] = tq[1] + t.y; tq[2] = tq[2] + t.z; }
        break;
      case EP_APPLY_IMPULSE:          // Body.elm:305-338
        if (bk == 2) {
          V3 c = cross(sub(vb, x), va);
          v[0] = v[0] + im * va.x; v[1] = v[1] + im * va.y; v[2] = v[2] + im * va.z;
          double wx = w[0] + m[0] * c.x + m[3] * c.y + m[6] * c.z, wy = w[1] + m[1] * c.x + m[4] * c.y + m[7] * c.z, wz = w[2] + m[2] * c.x + m[5] * c.y + m[8] * c.z;
          w[0] = wx; w[1] = wy; w[2] = wz;
        }
        break;
      case EP_APPLY_TORQUE:           // Body.elm:371-391
        if (bk == 2) { tq[0] = tq[0] + va.x; tq[1] = tq[1] + va.y; tq[2] = tq[2] + va.z; }
        break;
      case EP_APPLY_ANGULAR_IMPULSE:  // Body.elm:394-430
        if (bk == 2) {
          double wx = w[0] + m[0] * va.x + m[3] * va.y + m[6] * va.z, wy = w[1] + m[1] * va.x + m[4] * va.y + m[7] * va.z, wz = w[2] + m[2] * va.x + m[5] * va.y + m[8] * va.z;
          w[0] = wx; w[1] = wy; w[2] = wz;
        }
        break;
      case EP_SET_VELOCITY:           // Physics.elm:958-983: no effect on static bodies
        if (bk != 1) { v[0] = va.x; v[1] = va.y; v[2] = va.z; }
        break;
      case EP_SET_ANGULAR_VELOCITY:   // Physics.elm:989-1014
        if (bk != 1) { w[0] = va.x; w[1] = va.y; w[2] = va.z; }
        break;
      default: return EP_EINVAL;
    }
  }
  return EP_OK;
}
