// ORACLE — TEST INFRASTRUCTURE ONLY (see ep_oracle_math.h).
// World-placed shapes: restates Shape.placeIn (src/Internal/Shape.elm:94-110) over the packed
// shape table of include/elm_physics_b200.h.
#pragma once
#include "../include/elm_physics_b200.h"
#include "ep_oracle_math.h"

namespace epo {

// src/Shapes/Convex.elm:62-64 — one direction group, already in WORLD order.
struct WGroup {
  bool twoSided;
  V3 n1; const int32_t* i1; int n_i1; double d1;
  V3 n2; const int32_t* i2; int n_i2; double d2;
};

// src/Shapes/Convex.elm:76-99 after placeIn (166-179)
struct WConvex {
  V3 position;
  bool isBox;
  V3 ax, ay, az, he;                 // Obb.Box payload (Convex.elm:110-112)
  std::vector<V3> buffer;            // placed vertex buffer, key order (VertexBuffer.map)
  std::vector<V3> vs;                // NotBox list = buffer in DESCENDING key order (Convex.elm:188-199)
  std::vector<WGroup> faces;         // world order = reverse of the CoM-frame list (Convex.elm:265-287)
  std::vector<std::pair<const int32_t*, int>> edges;   // uniqueEdges, stored order, flat endpoint pairs
};

struct WShape {
  int kind;
  // plane: normal, position; sphere: radius, position; capsule: radius, halfLength, axis, position; particle: position
  V3 position, normal, axis;
  double radius, halfLength;
  WConvex convex;
  double friction, bounciness;
};

// Convex.elm:217-234 boxCorners; list order (+++,++-,+-+,+--,-++,-+-,--+,---)
inline std::vector<V3> boxCorners(V3 c, V3 ax, V3 ay, V3 az, V3 he) {
  std::vector<V3> out;
  static const double S[8][3] = {{1, 1, 1}, {1, 1, -1}, {1, -1, 1}, {1, -1, -1}, {-1, 1, 1}, {-1, 1, -1}, {-1, -1, 1}, {-1, -1, -1}};
  for (auto& s : S) {
    double sx = s[0], sy = s[1], sz = s[2];
    out.push_back({c.x + sx * he.x * ax.x + sy * he.y * ay.x + sz * he.z * az.x,
                   c.y + sx * he.x * ax.y + sy * he.y * ay.y + sz * he.z * az.y,
                   c.z + sx * he.x * ax.z + sy * he.y * ay.z + sz * he.z * az.z});
  }
  return out;
}

// Convex.elm:205-214
inline std::vector<V3> convexVertices(const WConvex& c) {
  return c.isBox ? boxCorners(c.position, c.ax, c.ay, c.az, c.he) : c.vs;
}

inline V3 rd3(const double* p) { return {p[0], p[1], p[2]}; }

// Shape.placeIn for one (template, material) instance.
inline WShape placeShape(const ep_shapes* S, int shape, const Tf& t, double friction, double bounciness) {
  WShape w{};
  w.kind = S->kind[shape];
  w.friction = friction;
  w.bounciness = bounciness;
  const double* f = S->f64 + S->f64_off[shape];
  const int32_t* ib = S->i32 + S->i32_off[shape];
  switch (w.kind) {
    case EP_SPHERE:   // Sphere.elm:35-41
      w.position = pointPlaceIn(t, rd3(f)); w.radius = f[3]; break;
    case EP_CAPSULE:  // Capsule.elm:49-57
      w.position = pointPlaceIn(t, rd3(f)); w.axis = directionPlaceIn(t, rd3(f + 3)); w.radius = f[6]; w.halfLength = f[7]; break;
    case EP_PLANE:    // Plane.elm:13-17
      w.position = pointPlaceIn(t, rd3(f)); w.normal = directionPlaceIn(t, rd3(f + 3)); break;
    case EP_PARTICLE:
      w.position = pointPlaceIn(t, rd3(f)); break;
    case EP_CONVEX: {  // Convex.elm:166-199, 265-287
      WConvex& c = w.convex;
      int isBox = ib[0], nv = ib[1], ng = ib[2], ne = ib[3];
      c.position = pointPlaceIn(t, rd3(f));
      c.isBox = isBox != 0;
      if (c.isBox) {
        c.ax = directionPlaceIn(t, rd3(f + 3)); c.ay = directionPlaceIn(t, rd3(f + 6)); c.az = directionPlaceIn(t, rd3(f + 9));
        c.he = rd3(f + 12);
      }
      const double* fv = f + EP_CONVEX_F64_HEADER;
      c.buffer.resize(nv);
      for (int i = 0; i < nv; i++) c.buffer[i] = pointPlaceIn(t, rd3(fv + 3 * i));
      if (!c.isBox) { c.vs.resize(nv); for (int i = 0; i < nv; i++) c.vs[i] = c.buffer[nv - 1 - i]; }
      const double* fg = fv + 3 * nv;
      const int32_t* ig = ib + EP_CONVEX_I32_HEADER;
      c.faces.resize(ng);
      for (int g = 0; g < ng; g++) {
        WGroup wg{};
        const double* gf = fg + EP_GROUP_F64 * g;
        const int32_t* gi = ig + EP_GROUP_I32 * g;
        wg.twoSided = gi[0] != 0;
        wg.n1 = directionPlaceIn(t, rd3(gf)); wg.d1 = gf[3]; wg.n_i1 = gi[1]; wg.i1 = ib + gi[2];
        if (wg.twoSided) { wg.n2 = directionPlaceIn(t, rd3(gf + 4)); wg.d2 = gf[7]; wg.n_i2 = gi[3]; wg.i2 = ib + gi[4]; }
        c.faces[ng - 1 - g] = wg;   // placeFaces reverses the group order
      }
      const int32_t* ie = ig + EP_GROUP_I32 * ng;
      c.edges.resize(ne);
      for (int e = 0; e < ne; e++) c.edges[e] = {ib + ie[EP_EDGE_I32 * e + 1], ie[EP_EDGE_I32 * e]};
      break;
    }
  }
  return w;
}

}  // namespace epo
