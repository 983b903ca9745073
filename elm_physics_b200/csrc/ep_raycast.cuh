// Physics.raycast (src/Physics.elm:802-841), batched: one thread per ray, closest hit over every shape of every body of
// the ray's world (Body.raycast, src/Internal/Body.elm:465-489; Shape.raycast, src/Internal/Shape.elm:132-148).  Bodies in
// list order, shapes in worldShapesWithMaterials order, faces in world order; every comparison is the reference's strict
// "new - best < 0", so the first hit wins ties.  Point masses are never hit; planes only from their front side.
#pragma once
#include "ep_kernels.cuh"

namespace ep {

struct RayHit { bool some; double distance; V3 point, normal; };

// Plane.raycast (src/Shapes/Plane.elm:20-49)
EPD RayHit ray_plane(V3 from, V3 dir, V3 position, V3 normal) {
  RayHit h; h.some = false; h.distance = 0; h.point = mk(0, 0, 0); h.normal = mk(0, 0, 0);
  double d = vdot(dir, normal);
  if (d < 0) {
    V3 pointToFrom = vsub(position, from);
    double scalar = vdot(normal, pointToFrom) / d;
    if (scalar >= 0) { h.some = true; h.distance = scalar; h.point = mk(dir.x * scalar + from.x, dir.y * scalar + from.y, dir.z * scalar + from.z); h.normal = normal; }
  }
  return h;
}
// Sphere.raycast (src/Shapes/Sphere.elm:49-86); the normal is normalised at the top level
EPD RayHit ray_sphere(V3 from, V3 dir, V3 p, double radius) {
  RayHit h; h.some = false; h.distance = 0; h.point = mk(0, 0, 0); h.normal = mk(0, 0, 0);
  double b = 2 * (dir.x * (from.x - p.x) + dir.y * (from.y - p.y) + dir.z * (from.z - p.z));
  double c = (from.x - p.x) * (from.x - p.x) + (from.y - p.y) * (from.y - p.y) + (from.z - p.z) * (from.z - p.z) - radius * radius;
  double delta = b * b - 4 * c;
  if (delta < 0) return h;
  double distance = (-b - sqrt(delta)) / 2;
  if (distance >= 0) {
    h.some = true; h.distance = distance;
    h.point = mk(from.x + dir.x * distance, from.y + dir.y * distance, from.z + dir.z * distance);
    h.normal = vsub(h.point, p);
  }
  return h;
}
// raycastFace (src/Shapes/Convex.elm:1138-1212) with foldFaceEdges (1215-1245): (v0,v1) ... (v_last,v0)
EPD void ray_face(const Cvx& c, V3 dir, V3 from, V3 normal, const int* idx, int n, RayHit& hit) {
  double d = vdot(dir, normal);
  if (!(d < 0)) return;
  V3 point = n > 0 ? c.vert(idx[0]) : mk(0, 0, 0);
  double scalar = vdot(normal, vsub(point, from)) / d;
  if (!(scalar >= 0)) return;
  V3 ip = mk(dir.x * scalar + from.x, dir.y * scalar + from.y, dir.z * scalar + from.z);
  bool inside = true;
  if (n >= 2)
    for (int i = 0; i < n; i++) {
      V3 p1 = c.vert(idx[i]), p2 = c.vert(idx[i + 1 < n ? i + 1 : 0]);
      inside = inside && (vdot(vsub(ip, p1), vcross(normal, vsub(p2, p1))) > 0);
    }
  if (!inside) return;
  if (hit.some && !(scalar - hit.distance < 0)) return;
  hit.some = true; hit.distance = scalar; hit.point = ip; hit.normal = normal;
}
// Convex.raycast (src/Shapes/Convex.elm:1117-1135): world group order, primary then partner
EPD RayHit ray_convex(V3 from, V3 dir, const Cvx& c) {
  RayHit h; h.some = false; h.distance = 0; h.point = mk(0, 0, 0); h.normal = mk(0, 0, 0);
  for (int k = 0; k < c.ng; k++) {
    Face f = group_face(c, k, false);
    ray_face(c, dir, from, f.n, f.idx, f.cnt, h);
    if (c.gi(k)[0] != 0) { Face g = group_face(c, k, true); ray_face(c, dir, from, g.n, g.idx, g.cnt, h); }
  }
  return h;
}
// capEntry (src/Shapes/Capsule.elm:108-137)
EPD bool ray_cap_entry(V3 p, V3 dir, V3 axis, double radius, double capOffset, double& t) {
  double qx = p.x - capOffset * axis.x, qy = p.y - capOffset * axis.y, qz = p.z - capOffset * axis.z;
  double b = 2 * (dir.x * qx + dir.y * qy + dir.z * qz);
  double c = qx * qx + qy * qy + qz * qz - radius * radius;
  double delta = b * b - 4 * c;
  if (delta < 0) return false;
  t = (-b - sqrt(delta)) / 2;
  return !(t < 0);
}
// Capsule.raycast (src/Shapes/Capsule.elm:65-204), rayParallelTolerance 1e-10 (17-19)
EPD RayHit ray_capsule(V3 from, V3 dir, const Cap& s) {
  RayHit h; h.some = false; h.distance = 0; h.point = mk(0, 0, 0); h.normal = mk(0, 0, 0);
  const double radius = s.r, halfLength = s.h; const V3 axis = s.axis, position = s.pos;
  V3 p = vsub(from, position);
  double pDotAxis = vdot(p, axis), dDotAxis = vdot(dir, axis);
  double pPerpX = p.x - pDotAxis * axis.x, pPerpY = p.y - pDotAxis * axis.y, pPerpZ = p.z - pDotAxis * axis.z;
  double dPerpX = dir.x - dDotAxis * axis.x, dPerpY = dir.y - dDotAxis * axis.y, dPerpZ = dir.z - dDotAxis * axis.z;
  double cylA = dPerpX * dPerpX + dPerpY * dPerpY + dPerpZ * dPerpZ;
  double cylB = 2 * (pPerpX * dPerpX + pPerpY * dPerpY + pPerpZ * dPerpZ);
  double cylC = pPerpX * pPerpX + pPerpY * pPerpY + pPerpZ * pPerpZ - radius * radius;
  bool some = false; double t = 0;
  if (cylA - 1.0e-10 < 0) {
    if (cylC > 0) some = false;
    else if (dDotAxis < 0) some = ray_cap_entry(p, dir, axis, radius, halfLength, t);
    else some = ray_cap_entry(p, dir, axis, radius, -halfLength, t);
  } else {
    double delta = cylB * cylB - 4 * cylA * cylC;
    if (!(delta < 0)) {
      double tt = (-cylB - sqrt(delta)) / (2 * cylA);
      if (!(tt < 0)) {
        double axial = pDotAxis + tt * dDotAxis;
        if (axial - halfLength > 0) some = ray_cap_entry(p, dir, axis, radius, halfLength, t);
        else if (axial + halfLength < 0) some = ray_cap_entry(p, dir, axis, radius, -halfLength, t);
        else { some = true; t = tt; }
      }
    }
  }
  if (!some) return h;
  h.some = true; h.distance = t;
  h.point = mk(from.x + dir.x * t, from.y + dir.y * t, from.z + dir.z * t);
  double x = pDotAxis + t * dDotAxis;
  double clampedAxial = (x < -halfLength) ? -halfLength : ((x > halfLength) ? halfLength : x);      // Basics.clamp
  h.normal = mk(h.point.x - position.x - clampedAxial * axis.x, h.point.y - position.y - clampedAxial * axis.y, h.point.z - position.z - clampedAxial * axis.z);
  return h;
}

// attached != 0: ray_world holds BODY rows; the ray is given in the centre-of-mass frame of that body, placed in the world with the
// body's current transform3d (Transform3d.pointPlaceIn / directionPlaceIn) and never hits its own body -- the RaycastCar pattern
// (examples/src/RaycastCar/Car.elm:116-128: `Axis3d.placeIn frame`, cast against the bodies without the car)
__global__ void __launch_bounds__(128) k_raycast(Dev d, int n_rays, const int* ray_world, const double* from, const double* direction,
                                                  int* hit_body, double* hit_distance, double* hit_point, double* hit_normal, int attached) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rays; i += gridDim.x * blockDim.x) {
    const int own = attached ? ray_world[i] : -1;
    const int w = attached ? d.body_world[own] : ray_world[i];
    V3 f = ld3(from + 3 * (size_t)i), dir = ld3(direction + 3 * (size_t)i);
    if (attached) {
      Q4 q; q.x = d.quat[4 * (size_t)own]; q.y = d.quat[4 * (size_t)own + 1]; q.z = d.quat[4 * (size_t)own + 2]; q.w = d.quat[4 * (size_t)own + 3];
      f = point_place(ld3(d.pos + 3 * (size_t)own), q, f);
      dir = rotate(q, dir);
    }
    bool any = false; int bestBody = -1; RayHit best; best.some = false; best.distance = 0; best.point = mk(0, 0, 0); best.normal = mk(0, 0, 0);
    for (int r = d.world_body_off[w]; r < d.world_body_off[w + 1]; r++) {
      if (r == own) continue;
      RayHit bodyHit; bodyHit.some = false; bodyHit.distance = 0; bodyHit.point = mk(0, 0, 0); bodyHit.normal = mk(0, 0, 0);
      for (int inst = d.inst_off[r]; inst < d.inst_off[r + 1]; inst++) {
        const ShapeRef s = shape_ref(d, inst);
        RayHit h; h.some = false;
        switch (s.kind) {
          case SH_PLANE: { Pln p = as_plane(s); h = ray_plane(f, dir, p.pos, p.n); break; }
          case SH_SPHERE: { Sph sp = as_sphere(s); h = ray_sphere(f, dir, sp.pos, sp.r); break; }
          case SH_CONVEX: h = ray_convex(f, dir, make_cvx(s.f, s.ib)); break;
          case SH_CAPSULE: h = ray_capsule(f, dir, as_capsule(s)); break;
          default: break;                                   // particles: Nothing
        }
        if (h.some && (!bodyHit.some || h.distance - bodyHit.distance < 0)) bodyHit = h;
      }
      if (bodyHit.some && (!any || bodyHit.distance - best.distance < 0)) { any = true; best = bodyHit; bestBody = r; }
    }
    hit_body[i] = any ? bestBody : -1;
    const V3 n = any ? vnormalize(best.normal) : mk(0, 0, 0);         // Physics.elm:838
    hit_distance[i] = any ? best.distance : 0;
    hit_point[3 * (size_t)i] = any ? best.point.x : 0; hit_point[3 * (size_t)i + 1] = any ? best.point.y : 0; hit_point[3 * (size_t)i + 2] = any ? best.point.z : 0;
    hit_normal[3 * (size_t)i] = n.x; hit_normal[3 * (size_t)i + 1] = n.y; hit_normal[3 * (size_t)i + 2] = n.z;
  }
}

// ---- between-step body mutators (src/Internal/Body.elm:305-430 behind src/Physics.elm:857-1017) ---------------------
// ops are sorted by body (stable) on the host; one thread per body runs its operations in list order
__global__ void __launch_bounds__(128) k_apply(Dev d, int n_runs, const int* run_off, const int* kind, const int* body_row, const double* a, const double* b) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_runs; t += gridDim.x * blockDim.x) {
    const int r = body_row[run_off[t]];
    const int bk = d.kind[r];
    V3 v = ld3(d.vel + 3 * (size_t)r), w = ld3(d.angvel + 3 * (size_t)r), f = ld3(d.force + 3 * (size_t)r), tq = ld3(d.torque + 3 * (size_t)r);
    const V3 x = ld3(d.pos + 3 * (size_t)r);
    const double im = d.inv_mass[r];
    const double* m = d.iiw + 9 * (size_t)r;      // m11,m21,m31,m12,m22,m32,m13,m23,m33
    for (int k = run_off[t]; k < run_off[t + 1]; k++) {
      const V3 va = ld3(a + 3 * (size_t)k), vb = ld3(b + 3 * (size_t)k);
      switch (kind[k]) {
        case 0: if (bk == 2) { V3 rp = vsub(vb, x); V3 t2 = vcross(rp, va); f = vadd(f, va); tq = vadd(tq, t2); } break;
        case 1: if (bk == 2) {
          V3 rp = vsub(vb, x); V3 c = vcross(rp, va);
          v = mk(v.x + im * va.x, v.y + im * va.y, v.z + im * va.z);
          w = mk(w.x + m[0] * c.x + m[3] * c.y + m[6] * c.z, w.y + m[1] * c.x + m[4] * c.y + m[7] * c.z, w.z + m[2] * c.x + m[5] * c.y + m[8] * c.z);
        } break;
        case 2: if (bk == 2) tq = vadd(tq, va); break;
        case 3: if (bk == 2) w = mk(w.x + m[0] * va.x + m[3] * va.y + m[6] * va.z, w.y + m[1] * va.x + m[4] * va.y + m[7] * va.z, w.z + m[2] * va.x + m[5] * va.y + m[8] * va.z); break;
        case 4: if (bk != 1) v = va; break;
        case 5: if (bk != 1) w = va; break;
        default: break;
      }
    }
    st3(d.vel + 3 * (size_t)r, v); st3(d.angvel + 3 * (size_t)r, w); st3(d.force + 3 * (size_t)r, f); st3(d.torque + 3 * (size_t)r, tq);
  }
}

}  // namespace ep
