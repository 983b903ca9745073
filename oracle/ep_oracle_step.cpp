// ORACLE — TEST INFRASTRUCTURE ONLY (see ep_oracle_math.h).
// CPU restatement of the step pipeline behind Physics.simulate: gravity sort (src/Physics.elm:548-573),
// BroadPhase.getPairs (src/Internal/BroadPhase.elm), Equation.equationsForPair (src/Internal/Equation.elm),
// Islands (src/Internal/Islands.elm), Solver.solve (src/Internal/Solver.elm) and SolverBody.solved
// (src/Internal/SolverBody.elm).  Single-threaded, one world at a time, as the reference is.
#include <algorithm>
#include <cstring>
#include <map>
#include <unordered_map>

#include "ep_oracle.h"
#include "ep_oracle_internal.h"

namespace epo {

struct OBody {
  int row, id, kind;
  Tf tf; V3 vel, angvel, force, torque;
  double invMass; V3 invInertia; M3 iiw;
  double ldp, adp; V3 linLock, angLock; double radius;
  std::vector<WShape> worldShapes;
};

// Equation.elm:26-36
struct Jac { double wAx, wAy, wAz, vBx, vBy, vBz, wBx, wBy, wBz; };
// Equation.elm:63-93
struct ContactEq {
  double normalLambda, friction1Lambda, friction2Lambda;
  Jac normal, friction1, friction2;
  double normalSolverB, normalSolverInvC, normalMinImpulse, normalMaxImpulse;
  double friction1SolverB, friction1SolverInvC, friction2SolverB, friction2SolverInvC, spookEps, frictionCoefficient;
  int shapeKey; int64_t featureKey;
};
// Equation.elm:503-511
struct ConEq { Jac jacobian; double solverB, solverInvC, spookEps, minImpulse, maxImpulse, solverLambda; };
struct ConstraintIn { int kind; double side1[12], side2[12]; };   // Constraint.elm:9-13, role-ordered (body1, body2)

struct Group {
  OBody* b1; OBody* b2;
  std::vector<SolverContact> contacts;     // solver order
  std::vector<ConstraintIn> constraints;   // list order of constraintsBetween
  std::vector<ContactEq> ceq;              // solver order
  std::vector<ConEq> jeq;                  // solver order
  int root;
};
struct SB { double vX, vY, vZ, wX, wY, wZ; };
struct WarmEntry { int shapeKey; int64_t featureKey; double lambda; V3 t1; };
typedef std::unordered_map<int64_t, std::vector<WarmEntry>> WarmCache;   // value: pairGroup.contacts (L) order

static const double kDefaultMaxImpulse = 1000000, kWarmStartFactor = 0.85, kDefaultRelaxation = 3, kDefaultStiffness = 10000000;

// ContactId.elm:62-68
static inline int64_t bodyKey(int64_t a, int64_t b) { return (a - b <= 0) ? a * 262144 + b : b * 262144 + a; }

// ---- Equation.elm:520-631 -------------------------------------------------------------------------
static double computeGW(const OBody& bi, const OBody& bj, const Jac& j) {
  return -(j.vBx * bi.vel.x + j.vBy * bi.vel.y + j.vBz * bi.vel.z)
         + (j.wAx * bi.angvel.x + j.wAy * bi.angvel.y + j.wAz * bi.angvel.z)
         + (j.vBx * bj.vel.x + j.vBy * bj.vel.y + j.vBz * bj.vel.z)
         + (j.wBx * bj.angvel.x + j.wBy * bj.angvel.y + j.wBz * bj.angvel.z);
}
static double computeGimgt(const OBody& bi, const OBody& bj, const Jac& j) {
  const M3 &I = bi.iiw, &J = bj.iiw;
  return bi.invMass + bj.invMass
         + (j.wAx * (I.m11 * j.wAx + I.m12 * j.wAy + I.m13 * j.wAz))
         + (j.wAy * (I.m21 * j.wAx + I.m22 * j.wAy + I.m23 * j.wAz))
         + (j.wAz * (I.m31 * j.wAx + I.m32 * j.wAy + I.m33 * j.wAz))
         + (j.wBx * (J.m11 * j.wBx + J.m12 * j.wBy + J.m13 * j.wBz))
         + (j.wBy * (J.m21 * j.wBx + J.m22 * j.wBy + J.m23 * j.wBz))
         + (j.wBz * (J.m31 * j.wBx + J.m32 * j.wBy + J.m33 * j.wBz));
}
static double computeGiMf(V3 gravity, const OBody& bi, const OBody& bj, const Jac& j) {
  V3 gi = bi.kind == 2 ? gravity : V3{0, 0, 0};
  V3 gj = bj.kind == 2 ? gravity : V3{0, 0, 0};
  const M3 &I = bi.iiw, &J = bj.iiw;
  return -(j.vBx * (bi.invMass * bi.force.x + gi.x) + j.vBy * (bi.invMass * bi.force.y + gi.y) + j.vBz * (bi.invMass * bi.force.z + gi.z))
         + (j.vBx * (bj.invMass * bj.force.x + gj.x) + j.vBy * (bj.invMass * bj.force.y + gj.y) + j.vBz * (bj.invMass * bj.force.z + gj.z))
         + (j.wAx * (I.m11 * bi.torque.x + I.m12 * bi.torque.y + I.m13 * bi.torque.z))
         + (j.wAy * (I.m21 * bi.torque.x + I.m22 * bi.torque.y + I.m23 * bi.torque.z))
         + (j.wAz * (I.m31 * bi.torque.x + I.m32 * bi.torque.y + I.m33 * bi.torque.z))
         + (j.wBx * (J.m11 * bj.torque.x + J.m12 * bj.torque.y + J.m13 * bj.torque.z))
         + (j.wBy * (J.m21 * bj.torque.x + J.m22 * bj.torque.y + J.m23 * bj.torque.z))
         + (j.wBz * (J.m31 * bj.torque.x + J.m32 * bj.torque.y + J.m33 * bj.torque.z));
}
struct Ctx { double dt; V3 gravity; };
static double computeSolverB(const Ctx& ctx, const OBody& bi, const OBody& bj, const Jac& j, double velocityB) {
  return velocityB - (ctx.dt * computeGiMf(ctx.gravity, bi, bj, j));
}
static double computeSolverInvC(double eps, const OBody& bi, const OBody& bj, const Jac& j) { return 1 / (computeGimgt(bi, bj, j) + eps); }
static double computeContactB(double spookA, double spookB, double bounciness, V3 pi, V3 pj, V3 ni, const OBody& bi, const OBody& bj, const Jac& j) {
  double g = ((pj.x - pi.x) * ni.x) + ((pj.y - pi.y) * ni.y) + ((pj.z - pi.z) * ni.z);
  double gW = (bounciness + 1) * (dot(bj.vel, ni) - dot(bi.vel, ni))
              + (bj.angvel.x * j.wBx + bj.angvel.y * j.wBy + bj.angvel.z * j.wBz)
              + (bi.angvel.x * j.wAx + bi.angvel.y * j.wAy + bi.angvel.z * j.wAz);
  return -g * spookA - gW * spookB;
}
static Jac contactJac(V3 n, V3 ri, V3 rj) {   // wA = n x ri, vB = n, wB = rj x n
  return Jac{n.y * ri.z - n.z * ri.y, n.z * ri.x - n.x * ri.z, n.x * ri.y - n.y * ri.x, n.x, n.y, n.z,
             rj.y * n.z - rj.z * n.y, rj.z * n.x - rj.x * n.z, rj.x * n.y - rj.y * n.x};
}
struct Spook { double a, b, eps; };
static Spook spook(double dt) {
  return Spook{4.0 / (dt * (1 + 4 * kDefaultRelaxation)), (4.0 * kDefaultRelaxation) / (1 + 4 * kDefaultRelaxation),
               4.0 / (dt * dt * kDefaultStiffness * (1 + 4 * kDefaultRelaxation))};
}
// Equation.elm:364-450
static ContactEq contactEquations(double seedLambda, V3 cachedT1, const Ctx& ctx, const OBody& b1, const OBody& b2, const SolverContact& sc) {
  Spook sp = spook(ctx.dt);
  const Contact& c = sc.contact;
  V3 ri = sub(c.pi, b1.tf.o), rj = sub(c.pj, b2.tf.o);
  V3 t1, t2; stableTangents(cachedT1, c.ni, t1, t2);
  ContactEq e{};
  e.normal = contactJac(c.ni, ri, rj); e.friction1 = contactJac(t1, ri, rj); e.friction2 = contactJac(t2, ri, rj);
  e.normalLambda = seedLambda; e.friction1Lambda = 0; e.friction2Lambda = 0;
  e.normalSolverB = computeSolverB(ctx, b1, b2, e.normal, computeContactB(sp.a, sp.b, sc.bounciness, c.pi, c.pj, c.ni, b1, b2, e.normal));
  e.normalSolverInvC = computeSolverInvC(sp.eps, b1, b2, e.normal);
  e.normalMinImpulse = 0; e.normalMaxImpulse = kDefaultMaxImpulse;
  e.friction1SolverB = computeSolverB(ctx, b1, b2, e.friction1, -computeGW(b1, b2, e.friction1) * sp.b);
  e.friction1SolverInvC = computeSolverInvC(sp.eps, b1, b2, e.friction1);
  e.friction2SolverB = computeSolverB(ctx, b1, b2, e.friction2, -computeGW(b1, b2, e.friction2) * sp.b);
  e.friction2SolverInvC = computeSolverInvC(sp.eps, b1, b2, e.friction2);
  e.spookEps = sp.eps; e.frictionCoefficient = sc.friction; e.shapeKey = c.shapeKey; e.featureKey = c.featureKey;
  return e;
}
static ConEq jointEq(const Ctx& ctx, const OBody& b1, const OBody& b2, const Jac& j, double velocityB, double eps) {
  return ConEq{j, computeSolverB(ctx, b1, b2, j, velocityB), computeSolverInvC(eps, b1, b2, j), eps, -kDefaultMaxImpulse, kDefaultMaxImpulse, 0};
}
// Equation.elm:308-361 ; `eqs` grows by push_back == cons
static void addPointToPoint(const Ctx& ctx, const OBody& b1, const OBody& b2, V3 pivot1, V3 pivot2, std::vector<ConEq>& eqs) {
  Spook sp = spook(ctx.dt);
  V3 ri = directionPlaceIn(b1.tf, pivot1), rj = directionPlaceIn(b2.tf, pivot2);
  V3 basis[3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (V3 ni : basis) {
    V3 pi = add(b1.tf.o, ri), pj = add(b2.tf.o, rj);
    Jac j = contactJac(ni, ri, rj);
    eqs.push_back(jointEq(ctx, b1, b2, j, computeContactB(sp.a, sp.b, 0, pi, pj, ni, b1, b2, j), sp.eps));
  }
}
// Equation.elm:272-305, 549-565
static void addRotational(const Ctx& ctx, const OBody& b1, const OBody& b2, V3 ni, V3 nj, std::vector<ConEq>& eqs) {
  Spook sp = spook(ctx.dt);
  Jac j{nj.y * ni.z - nj.z * ni.y, nj.z * ni.x - nj.x * ni.z, nj.x * ni.y - nj.y * ni.x, 0, 0, 0,
        ni.y * nj.z - ni.z * nj.y, ni.z * nj.x - ni.x * nj.z, ni.x * nj.y - ni.y * nj.x};
  double g = 0 - dot(ni, nj);
  double gW = computeGW(b1, b2, j);
  eqs.push_back(jointEq(ctx, b1, b2, j, -g * sp.a - gW * sp.b, sp.eps));
}
// Equation.elm:157-228
static void addConstraintEquations(const Ctx& ctx, const OBody& b1, const OBody& b2, const ConstraintIn& c, std::vector<ConEq>& eqs) {
  const double *s1 = c.side1, *s2 = c.side2;
  switch (c.kind) {
    case EP_POINT_TO_POINT: addPointToPoint(ctx, b1, b2, rd3(s1), rd3(s2), eqs); break;
    case EP_HINGE: {
      addPointToPoint(ctx, b1, b2, rd3(s1), rd3(s2), eqs);
      V3 worldAxis2 = directionPlaceIn(b2.tf, rd3(s2 + 3));
      V3 ni1, ni2; tangents(directionPlaceIn(b1.tf, rd3(s1 + 3)), ni1, ni2);
      addRotational(ctx, b1, b2, ni1, worldAxis2, eqs);
      addRotational(ctx, b1, b2, ni2, worldAxis2, eqs);
      break;
    }
    case EP_LOCK: {
      addPointToPoint(ctx, b1, b2, rd3(s1), rd3(s2), eqs);
      V3 x1 = directionPlaceIn(b1.tf, rd3(s1 + 3)), y1 = directionPlaceIn(b1.tf, rd3(s1 + 6)), z1 = directionPlaceIn(b1.tf, rd3(s1 + 9));
      V3 x2 = directionPlaceIn(b2.tf, rd3(s2 + 3)), y2 = directionPlaceIn(b2.tf, rd3(s2 + 6)), z2 = directionPlaceIn(b2.tf, rd3(s2 + 9));
      addRotational(ctx, b1, b2, x1, y2, eqs);
      addRotational(ctx, b1, b2, y1, z2, eqs);
      addRotational(ctx, b1, b2, z1, x2, eqs);
      break;
    }
    case EP_DISTANCE: {
      Spook sp = spook(ctx.dt);
      double halfDistance = s1[0] / 2;
      V3 ni = direction(b2.tf.o, b1.tf.o);
      V3 ri = scale(halfDistance, ni), rj = scale(-halfDistance, ni);
      V3 pi = add(ri, b1.tf.o), pj = add(rj, b2.tf.o);
      Jac j = contactJac(ni, ri, rj);
      eqs.push_back(jointEq(ctx, b1, b2, j, computeContactB(sp.a, sp.b, 0, pi, pj, ni, b1, b2, j), sp.eps));
      break;
    }
  }
}

// ---- Solver.elm -----------------------------------------------------------------------------------
// Solver.elm:39-85
static void applyEquationWarmStart(double lambda, const Jac& j, const OBody& b1, const OBody& b2, SB& s1, SB& s2) {
  if (lambda == 0) return;
  if (b1.kind == 2) {
    const M3& I = b1.iiw; double k1 = lambda * b1.invMass;
    s1 = SB{s1.vX - k1 * j.vBx, s1.vY - k1 * j.vBy, s1.vZ - k1 * j.vBz,
            s1.wX + (I.m11 * j.wAx + I.m12 * j.wAy + I.m13 * j.wAz) * lambda,
            s1.wY + (I.m21 * j.wAx + I.m22 * j.wAy + I.m23 * j.wAz) * lambda,
            s1.wZ + (I.m31 * j.wAx + I.m32 * j.wAy + I.m33 * j.wAz) * lambda};
  }
  if (b2.kind == 2) {
    const M3& I = b2.iiw; double k2 = lambda * b2.invMass;
    s2 = SB{s2.vX + k2 * j.vBx, s2.vY + k2 * j.vBy, s2.vZ + k2 * j.vBz,
            s2.wX + (I.m11 * j.wBx + I.m12 * j.wBy + I.m13 * j.wBz) * lambda,
            s2.wY + (I.m21 * j.wBx + I.m22 * j.wBy + I.m23 * j.wBz) * lambda,
            s2.wZ + (I.m31 * j.wBx + I.m32 * j.wBy + I.m33 * j.wBz) * lambda};
  }
}
// Solver.elm:505-550 (applyVelocityBody1/2): same arithmetic as the warm start, without the lambda==0 skip
static void applyVelocity(double dl, const Jac& j, const OBody& b1, const OBody& b2, SB& s1, SB& s2) {
  if (b1.kind == 2) {
    const M3& I = b1.iiw; double k = dl * b1.invMass;
    s1 = SB{s1.vX - k * j.vBx, s1.vY - k * j.vBy, s1.vZ - k * j.vBz,
            s1.wX + (I.m11 * j.wAx + I.m12 * j.wAy + I.m13 * j.wAz) * dl,
            s1.wY + (I.m21 * j.wAx + I.m22 * j.wAy + I.m23 * j.wAz) * dl,
            s1.wZ + (I.m31 * j.wAx + I.m32 * j.wAy + I.m33 * j.wAz) * dl};
  }
  if (b2.kind == 2) {
    const M3& I = b2.iiw; double k = dl * b2.invMass;
    s2 = SB{s2.vX + k * j.vBx, s2.vY + k * j.vBy, s2.vZ + k * j.vBz,
            s2.wX + (I.m11 * j.wBx + I.m12 * j.wBy + I.m13 * j.wBz) * dl,
            s2.wY + (I.m21 * j.wBx + I.m22 * j.wBy + I.m23 * j.wBz) * dl,
            s2.wZ + (I.m31 * j.wBx + I.m32 * j.wBy + I.m33 * j.wBz) * dl};
  }
}
static inline double gWlambda(const Jac& j, const SB& b1, const SB& b2) {
  return -(j.vBx * b1.vX + j.vBy * b1.vY + j.vBz * b1.vZ)
         + (j.wAx * b1.wX + j.wAy * b1.wY + j.wAz * b1.wZ)
         + (j.vBx * b2.vX + j.vBy * b2.vY + j.vBz * b2.vZ)
         + (j.wBx * b2.wX + j.wBy * b2.wY + j.wBz * b2.wZ);
}
// Solver.elm:850-864 = solveVelocityConstraints (572-619) then solveVelocityNormals (625-672)
static double velocityNonFrictionGroup(Group& g, SB& s1, SB& s2, double tot) {
  for (ConEq& c : g.jeq) {
    double lam = c.solverLambda;
    double prev = c.solverInvC * (c.solverB - gWlambda(c.jacobian, s1, s2) - c.spookEps * lam);
    double dl = (lam + prev - c.minImpulse < 0) ? c.minImpulse - lam : ((lam + prev - c.maxImpulse > 0) ? c.maxImpulse - lam : prev);
    applyVelocity(dl, c.jacobian, *g.b1, *g.b2, s1, s2);
    c.solverLambda = lam + dl;
    tot = tot + eabs(dl);
  }
  for (ContactEq& c : g.ceq) {
    double lam = c.normalLambda;
    double prev = c.normalSolverInvC * (c.normalSolverB - gWlambda(c.normal, s1, s2) - c.spookEps * lam);
    double dl = (lam + prev - c.normalMinImpulse < 0) ? c.normalMinImpulse - lam : ((lam + prev - c.normalMaxImpulse > 0) ? c.normalMaxImpulse - lam : prev);
    applyVelocity(dl, c.normal, *g.b1, *g.b2, s1, s2);
    c.normalLambda = lam + dl;
    tot = tot + eabs(dl);
  }
  return tot;
}
// Solver.elm:678-844
static double velocityFrictionGroup(Group& g, SB& s1, SB& s2, double tot) {
  const OBody &B1 = *g.b1, &B2 = *g.b2;
  const M3 &I1 = B1.iiw, &I2 = B2.iiw;
  for (ContactEq& c : g.ceq) {
    double normalLambda = c.normalLambda;
    const Jac& e1 = c.friction1;
    double cap1 = c.frictionCoefficient * normalLambda;
    double gW1 = gWlambda(e1, s1, s2);
    double dPrev1 = c.friction1SolverInvC * (c.friction1SolverB - gW1 - c.spookEps * c.friction1Lambda);
    double d1 = (c.friction1Lambda + dPrev1 + cap1 < 0) ? -cap1 - c.friction1Lambda : ((c.friction1Lambda + dPrev1 - cap1 > 0) ? cap1 - c.friction1Lambda : dPrev1);
    double k1a = d1 * B1.invMass;
    double b1vX = s1.vX - k1a * e1.vBx, b1vY = s1.vY - k1a * e1.vBy, b1vZ = s1.vZ - k1a * e1.vBz;
    double b1wX = s1.wX + (I1.m11 * e1.wAx + I1.m12 * e1.wAy + I1.m13 * e1.wAz) * d1;
    double b1wY = s1.wY + (I1.m21 * e1.wAx + I1.m22 * e1.wAy + I1.m23 * e1.wAz) * d1;
    double b1wZ = s1.wZ + (I1.m31 * e1.wAx + I1.m32 * e1.wAy + I1.m33 * e1.wAz) * d1;
    double k1b = d1 * B2.invMass;
    double b2vX = s2.vX + k1b * e1.vBx, b2vY = s2.vY + k1b * e1.vBy, b2vZ = s2.vZ + k1b * e1.vBz;
    double b2wX = s2.wX + (I2.m11 * e1.wBx + I2.m12 * e1.wBy + I2.m13 * e1.wBz) * d1;
    double b2wY = s2.wY + (I2.m21 * e1.wBx + I2.m22 * e1.wBy + I2.m23 * e1.wBz) * d1;
    double b2wZ = s2.wZ + (I2.m31 * e1.wBx + I2.m32 * e1.wBy + I2.m33 * e1.wBz) * d1;
    const Jac& e2 = c.friction2;
    double cap2 = c.frictionCoefficient * normalLambda;
    double gW2 = -(e2.vBx * b1vX + e2.vBy * b1vY + e2.vBz * b1vZ)
                 + (e2.wAx * b1wX + e2.wAy * b1wY + e2.wAz * b1wZ)
                 + (e2.vBx * b2vX + e2.vBy * b2vY + e2.vBz * b2vZ)
                 + (e2.wBx * b2wX + e2.wBy * b2wY + e2.wBz * b2wZ);
    double dPrev2 = c.friction2SolverInvC * (c.friction2SolverB - gW2 - c.spookEps * c.friction2Lambda);
    double d2 = (c.friction2Lambda + dPrev2 + cap2 < 0) ? -cap2 - c.friction2Lambda : ((c.friction2Lambda + dPrev2 - cap2 > 0) ? cap2 - c.friction2Lambda : dPrev2);
    if (B1.kind == 2) {
      double k2a = d2 * B1.invMass;
      s1 = SB{b1vX - k2a * e2.vBx, b1vY - k2a * e2.vBy, b1vZ - k2a * e2.vBz,
              b1wX + (I1.m11 * e2.wAx + I1.m12 * e2.wAy + I1.m13 * e2.wAz) * d2,
              b1wY + (I1.m21 * e2.wAx + I1.m22 * e2.wAy + I1.m23 * e2.wAz) * d2,
              b1wZ + (I1.m31 * e2.wAx + I1.m32 * e2.wAy + I1.m33 * e2.wAz) * d2};
    }
    if (B2.kind == 2) {
      double k2b = d2 * B2.invMass;
      s2 = SB{b2vX + k2b * e2.vBx, b2vY + k2b * e2.vBy, b2vZ + k2b * e2.vBz,
              b2wX + (I2.m11 * e2.wBx + I2.m12 * e2.wBy + I2.m13 * e2.wBz) * d2,
              b2wY + (I2.m21 * e2.wBx + I2.m22 * e2.wBy + I2.m23 * e2.wBz) * d2,
              b2wZ + (I2.m31 * e2.wBx + I2.m32 * e2.wBy + I2.m33 * e2.wBz) * d2};
    }
    c.friction1Lambda = c.friction1Lambda + d1;
    c.friction2Lambda = c.friction2Lambda + d2;
    tot = tot + eabs(d1) + eabs(d2);
  }
  return tot;
}

// Islands.elm:125-145 (path halving; observable result = root)
static int findRoot(int id, std::vector<int>& parents) {
  for (;;) {
    if (id < 0 || id >= (int)parents.size()) return id;
    int parent = parents[id];
    if (parent - id == 0) return id;
    int grand = parents[parent];
    if (grand - parent == 0) return parent;
    parents[id] = grand;
    id = grand;
  }
}

// Islands.init/connect/fold (Islands.elm:23-119) over groups given as (id, kind) pairs in CSR order.
// `connect` is applied in REVERSE group order, as buildAndWarmStart does (Solver.elm:201-206).
// Returns per-group root and the visiting order: islands by ascending root, groups inside in CSR order.
static void islandsOrder(int nGroups, const int* id1, const int* kind1, const int* id2, const int* kind2, int maxId,
                         std::vector<int>& root, std::vector<int>& gorder) {
  std::vector<int> parents(maxId + 1);
  for (int i = 0; i <= maxId; i++) parents[i] = i;
  for (int g = nGroups - 1; g >= 0; g--)
    if (kind1[g] == 2 && kind2[g] == 2) {   // Islands.connect, Islands.elm:30-46
      int ra = findRoot(id1[g], parents), rb = findRoot(id2[g], parents);
      if (ra - rb < 0) parents[rb] = ra; else if (ra - rb > 0) parents[ra] = rb;
    }
  root.resize(nGroups); gorder.resize(nGroups);
  for (int g = 0; g < nGroups; g++) { root[g] = findRoot(kind1[g] == 2 ? id1[g] : id2[g], parents); gorder[g] = g; }
  std::stable_sort(gorder.begin(), gorder.end(), [&](int a, int b) { return root[a] - root[b] < 0; });
}

// SolverBody.elm:97-225
static void integrateBody(OBody& b, const SB& s, double dt, V3 gravity) {
  if (b.kind == 1) return;
  if (b.kind == 3) {
    V3 v = b.vel, w = b.angvel;
    b.tf = normalizeTf(translateBy(V3{v.x * dt, v.y * dt, v.z * dt}, rotateBy(V3{w.x * dt, w.y * dt, w.z * dt}, b.tf)));
    b.force = V3{0, 0, 0}; b.torque = V3{0, 0, 0};
    return;
  }
  double ld = b.ldp, ad = b.adp;   // (1 - damping)^dt, computed by the caller (SolverBody.elm:149-153)
  V3 nv{((gravity.x + b.force.x * b.invMass) * dt + b.vel.x * ld + s.vX) * b.linLock.x,
        ((gravity.y + b.force.y * b.invMass) * dt + b.vel.y * ld + s.vY) * b.linLock.y,
        ((gravity.z + b.force.z * b.invMass) * dt + b.vel.z * ld + s.vZ) * b.linLock.z};
  double vlen = length(nv);
  double r = b.radius;
  V3 capped = ((vlen == 0) || (r == 0) || (vlen * dt - r < 0)) ? nv : scale(r / (vlen * dt), nv);
  const M3& I = b.iiw;
  V3 nw{((I.m11 * b.torque.x + I.m12 * b.torque.y + I.m13 * b.torque.z) * dt + b.angvel.x * ad + s.wX) * b.angLock.x,
        ((I.m21 * b.torque.x + I.m22 * b.torque.y + I.m23 * b.torque.z) * dt + b.angvel.y * ad + s.wY) * b.angLock.y,
        ((I.m31 * b.torque.x + I.m32 * b.torque.y + I.m33 * b.torque.z) * dt + b.angvel.z * ad + s.wZ) * b.angLock.z};
  b.tf = normalizeTf(translateBy(V3{capped.x * dt, capped.y * dt, capped.z * dt}, rotateBy(V3{nw.x * dt, nw.y * dt, nw.z * dt}, b.tf)));
  b.vel = nv; b.angvel = nw;
  b.iiw = invertedInertiaRotateIn(b.tf, b.invInertia);
  b.force = V3{0, 0, 0}; b.torque = V3{0, 0, 0};
}

struct GroupOut {
  int id1, id2, nConstraints, island; Tf pre1;
  std::vector<SolverContact> contacts; std::vector<double> lambda; std::vector<V3> t1;
};
struct WorldOut { std::vector<GroupOut> groups; int iterations = 0; int nIslands = 0; };

struct WorldIn {
  std::vector<OBody> bodies;      // user order
  V3 gravity; double dt; int iterations;
  const uint8_t* collide; int n;  // n x n or null
  std::vector<std::vector<std::pair<int, ConstraintIn>>> regs;   // per local body a: (local b, constraint with side1 = a-side)
};

// One Physics.simulate (src/Physics.elm:522-583) for one world.
static void stepWorld(const ep_shapes* S, const ep_bodies* B, WorldIn& W, const WarmCache& warm, WorldOut& out) {
  out = WorldOut{};
  int n = (int)W.bodies.size();
  if (n == 0) { out.iterations = 0; return; }
  Ctx ctx{W.dt, W.gravity};
  // worldShapesWithMaterials = placeIn transform3d of every geometry shape (Body.elm:253, SolverBody.elm:212)
  for (OBody& b : W.bodies) {
    b.worldShapes.clear();
    for (int k = B->inst_off[b.row]; k < B->inst_off[b.row + 1]; k++)
      b.worldShapes.push_back(placeShape(S, B->inst_shape[k], b.tf, B->inst_friction[k], B->inst_bounciness[k]));
  }
  // gravity sort, Physics.elm:548-573 (stable; comparator a-b<0 / a-b>0)
  std::vector<int> order(n);
  std::vector<double> key(n);
  for (int i = 0; i < n; i++) {
    order[i] = i;
    V3 p = W.bodies[i].tf.o;
    key[i] = -(p.x * W.gravity.x + p.y * W.gravity.y + p.z * W.gravity.z);
  }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] - key[b] < 0; });
  // BroadPhase.getPairs, BroadPhase.elm:18-215
  std::vector<Group> groups;   // enumeration (CSR) order
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) {
      OBody &b1 = W.bodies[order[i]], &b2 = W.bodies[order[j]];
      Group g{&b1, &b2, {}, {}, {}, {}, 0};
      double rr = b1.radius + b2.radius;
      double dx = b2.tf.o.x - b1.tf.o.x, dy = b2.tf.o.y - b1.tf.o.y, dz = b2.tf.o.z - b1.tf.o.z;
      double distSq = dx * dx + dy * dy + dz * dz;
      bool may = (rr * rr - distSq > 0) && (b1.kind == 2 || b2.kind == 2) && (W.collide == nullptr || W.collide[order[i] * n + order[j]] != 0);
      if (may) getContactsSolverOrder(b1.worldShapes, b2.worldShapes, g.contacts);
      if (!W.regs.empty() && (b1.kind == 2 || b2.kind == 2)) {   // constraintsBetween, BroadPhase.elm:144-186
        for (auto& r : W.regs[order[i]]) if (r.first == order[j]) g.constraints.push_back(r.second);
        if (g.constraints.empty())
          for (auto& r : W.regs[order[j]]) if (r.first == order[i]) {
            ConstraintIn c = r.second; ConstraintIn f = c;   // relativeToCenterOfMassFlipped, Constraint.elm:64-95
            if (c.kind != EP_DISTANCE) { memcpy(f.side1, c.side2, sizeof f.side1); memcpy(f.side2, c.side1, sizeof f.side2); }   // Distance length -> Distance length
            g.constraints.push_back(f);
          }
      }
      if (!g.contacts.empty() || !g.constraints.empty()) groups.push_back(std::move(g));
    }
  // Solver.solve, Solver.elm:223-290
  int maxId = -1;
  for (OBody& b : W.bodies) maxId = b.id > maxId ? b.id : maxId;
  std::vector<SB> sb(maxId + 1, SB{0, 0, 0, 0, 0, 0});
  // buildAndWarmStart walks pairGroups (reverse enumeration order), Solver.elm:122-208
  for (int gi = (int)groups.size() - 1; gi >= 0; gi--) {
    Group& g = groups[gi];
    const std::vector<WarmEntry>* wl = nullptr;
    if (!g.contacts.empty()) { auto it = warm.find(bodyKey(g.b1->id, g.b2->id)); if (it != warm.end()) wl = &it->second; }
    for (const SolverContact& sc : g.contacts) {   // buildContactEquations result == solver order
      double lambda = 0; V3 t1{0, 0, 0};
      if (wl) for (const WarmEntry& e : *wl) if ((e.featureKey - sc.contact.featureKey == 0) && (e.shapeKey - sc.contact.shapeKey == 0)) { lambda = e.lambda; t1 = e.t1; break; }
      g.ceq.push_back(contactEquations(lambda * kWarmStartFactor, t1, ctx, *g.b1, *g.b2, sc));
    }
    std::vector<ConEq> emitted;
    for (const ConstraintIn& c : g.constraints) addConstraintEquations(ctx, *g.b1, *g.b2, c, emitted);
    g.jeq.assign(emitted.rbegin(), emitted.rend());
    SB &s1 = sb[g.b1->id], &s2 = sb[g.b2->id];
    for (ConEq& c : g.jeq) applyEquationWarmStart(c.solverLambda, c.jacobian, *g.b1, *g.b2, s1, s2);
    for (ContactEq& c : g.ceq) {
      applyEquationWarmStart(c.normalLambda, c.normal, *g.b1, *g.b2, s1, s2);
      applyEquationWarmStart(c.friction1Lambda, c.friction1, *g.b1, *g.b2, s1, s2);
      applyEquationWarmStart(c.friction2Lambda, c.friction2, *g.b1, *g.b2, s1, s2);
    }
  }
  // Islands.connect (inside buildAndWarmStart) + Islands.fold, Islands.elm:30-119
  std::vector<int> gorder, groot;
  {
    std::vector<int> i1, k1, i2, k2;
    for (Group& g : groups) { i1.push_back(g.b1->id); k1.push_back(g.b1->kind); i2.push_back(g.b2->id); k2.push_back(g.b2->kind); }
    islandsOrder((int)groups.size(), i1.data(), k1.data(), i2.data(), k2.data(), maxId, groot, gorder);
    for (size_t i = 0; i < groups.size(); i++) groups[i].root = groot[i];
  }
  int minRem = W.iterations;
  size_t pos = 0;
  while (pos < gorder.size()) {
    size_t end = pos;
    while (end < gorder.size() && groups[gorder[end]].root == groups[gorder[pos]].root) end++;
    out.nIslands++;
    int remaining = W.iterations, rem;
    if (end - pos == 1) {   // solve2Body, Solver.elm:473-489
      Group& g = groups[gorder[pos]];
      SB &s1 = sb[g.b1->id], &s2 = sb[g.b2->id];
      for (;;) {
        double tot = velocityNonFrictionGroup(g, s1, s2, 0);
        tot = velocityFrictionGroup(g, s1, s2, tot);
        if (remaining == 1) { rem = 0; break; }
        if (tot - kSolverTolerance < 0) { rem = remaining - 1; break; }
        remaining--;
      }
    } else {                // step, Solver.elm:345-367
      for (;;) {
        double p1 = 0, p2 = 0;
        for (size_t k = pos; k < end; k++) { Group& g = groups[gorder[k]]; p1 = velocityNonFrictionGroup(g, sb[g.b1->id], sb[g.b2->id], p1); }
        for (size_t k = pos; k < end; k++) { Group& g = groups[gorder[k]]; p2 = velocityFrictionGroup(g, sb[g.b1->id], sb[g.b2->id], p2); }
        double deltaTot = p1 + p2;
        if (remaining == 1) { rem = 0; break; }
        if (deltaTot - kSolverTolerance < 0) { rem = remaining - 1; break; }
        remaining--;
      }
    }
    minRem = (minRem - rem < 0) ? minRem : rem;
    pos = end;
  }
  int used = W.iterations - minRem;
  out.iterations = (1 - used > 0) ? 1 : used;
  // pairGroups + next-frame warm start (Solver.elm:297-335)
  for (Group& g : groups) {
    GroupOut go{g.b1->id, g.b2->id, (int)g.constraints.size(), g.root, g.b1->tf, g.contacts, {}, {}};
    for (ContactEq& c : g.ceq) { go.lambda.push_back(c.normalLambda); go.t1.push_back(V3{c.friction1.vBx, c.friction1.vBy, c.friction1.vBz}); }
    out.groups.push_back(std::move(go));
  }
  // Array.map (SolverBody.solved dt gravity), Solver.elm:283-284
  for (OBody& b : W.bodies) integrateBody(b, sb[b.id], W.dt, W.gravity);
}

static void warmFromOut(const WorldOut& o, WarmCache& warm) {
  warm.clear();
  for (const GroupOut& g : o.groups) {
    if (g.contacts.empty()) continue;
    std::vector<WarmEntry> l;   // warmStartEntries conses over solver order => L order = reverse
    for (int k = (int)g.contacts.size() - 1; k >= 0; k--)
      l.push_back(WarmEntry{g.contacts[k].contact.shapeKey, g.contacts[k].contact.featureKey, g.lambda[k], g.t1[k]});
    warm[bodyKey(g.id1, g.id2)] = std::move(l);
  }
}

static M3 rdM3(const double* p) { return M3{p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]}; }
static void wr3(double* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

}  // namespace epo

using namespace epo;

extern "C" int ep_oracle_step(const ep_shapes* S, ep_bodies* B, const ep_params* P, const ep_contacts* warmIn, ep_contacts* out, int32_t n_steps) {
  int nw = B->n_worlds;
  std::vector<WorldOut> outs(nw);
  for (int w = 0; w < nw; w++) {
    WorldIn W;
    int r0 = B->world_body_off[w], r1 = B->world_body_off[w + 1];
    for (int r = r0; r < r1; r++) {
      OBody b{};
      b.row = r; b.id = B->body_id[r]; b.kind = B->kind[r];
      b.tf = Tf{rd3(B->pos + 3 * r), Quat{B->quat[4 * r], B->quat[4 * r + 1], B->quat[4 * r + 2], B->quat[4 * r + 3]}};
      b.vel = rd3(B->vel + 3 * r); b.angvel = rd3(B->angvel + 3 * r); b.force = rd3(B->force + 3 * r); b.torque = rd3(B->torque + 3 * r);
      b.invMass = B->inv_mass[r]; b.invInertia = rd3(B->inv_inertia + 3 * r); b.iiw = rdM3(B->inv_inertia_world + 9 * r);
      b.ldp = B->lin_damp_pow[r]; b.adp = B->ang_damp_pow[r]; b.linLock = rd3(B->lin_lock + 3 * r); b.angLock = rd3(B->ang_lock + 3 * r);
      b.radius = B->bounding_radius[r];
      W.bodies.push_back(b);
    }
    W.n = r1 - r0;
    W.gravity = rd3(P->gravity + 3 * w); W.dt = P->dt[w]; W.iterations = P->iterations[w];
    W.collide = P->collide ? P->collide + P->collide_off[w] : nullptr;
    for (int c = 0; c < P->n_constraints; c++) {
      int a = P->con_body_a[c], b = P->con_body_b[c];
      if (a < r0 || a >= r1) continue;
      if (W.regs.empty()) W.regs.resize(W.n);
      ConstraintIn ci; ci.kind = P->con_kind[c];
      memcpy(ci.side1, P->con_data + 24 * c, sizeof ci.side1); memcpy(ci.side2, P->con_data + 24 * c + 12, sizeof ci.side2);
      W.regs[a - r0].push_back({b - r0, ci});
    }
    WarmCache warm;
    if (warmIn && warmIn->world_group_off) {
      WorldOut prev;
      for (int g = warmIn->world_group_off[w]; g < warmIn->world_group_off[w + 1]; g++) {
        GroupOut go{}; go.id1 = warmIn->group_id1[g]; go.id2 = warmIn->group_id2[g];
        for (int c = warmIn->group_contact_off[g]; c < warmIn->group_contact_off[g + 1]; c++) {
          SolverContact sc{}; sc.contact.shapeKey = warmIn->shape_key[c]; sc.contact.featureKey = warmIn->feature_key[c];
          go.contacts.push_back(sc); go.lambda.push_back(warmIn->lambda[c]); go.t1.push_back(rd3(warmIn->t1 + 3 * c));
        }
        prev.groups.push_back(std::move(go));
      }
      warmFromOut(prev, warm);
    }
    for (int s = 0; s < n_steps; s++) {
      stepWorld(S, B, W, warm, outs[w]);
      warmFromOut(outs[w], warm);
    }
    for (OBody& b : W.bodies) {
      int r = b.row;
      wr3(B->pos + 3 * r, b.tf.o);
      B->quat[4 * r] = b.tf.q.x; B->quat[4 * r + 1] = b.tf.q.y; B->quat[4 * r + 2] = b.tf.q.z; B->quat[4 * r + 3] = b.tf.q.w;
      wr3(B->vel + 3 * r, b.vel); wr3(B->angvel + 3 * r, b.angvel); wr3(B->force + 3 * r, b.force); wr3(B->torque + 3 * r, b.torque);
      const M3& m = b.iiw; double* q = B->inv_inertia_world + 9 * r;
      q[0] = m.m11; q[1] = m.m21; q[2] = m.m31; q[3] = m.m12; q[4] = m.m22; q[5] = m.m32; q[6] = m.m13; q[7] = m.m23; q[8] = m.m33;
    }
  }
  if (!out) return EP_OK;
  int64_t ng = 0, nc = 0;
  for (auto& o : outs) { ng += (int64_t)o.groups.size(); for (auto& g : o.groups) nc += (int64_t)g.contacts.size(); }
  if (ng > out->group_cap || nc > out->contact_cap) return EP_ECAPACITY;
  int g = 0, c = 0;
  out->n_worlds = nw;
  for (int w = 0; w < nw; w++) {
    out->world_group_off[w] = g;
    out->iterations[w] = outs[w].iterations;
    if (out->world_n_islands) out->world_n_islands[w] = outs[w].nIslands;
    for (auto& go : outs[w].groups) {
      out->group_id1[g] = go.id1; out->group_id2[g] = go.id2;
      if (out->group_n_constraints) out->group_n_constraints[g] = go.nConstraints;
      if (out->group_island) out->group_island[g] = go.island;
      if (out->group_pre_pos1) { wr3(out->group_pre_pos1 + 3 * g, go.pre1.o); double* q = out->group_pre_quat1 + 4 * g; q[0] = go.pre1.q.x; q[1] = go.pre1.q.y; q[2] = go.pre1.q.z; q[3] = go.pre1.q.w; }
      out->group_contact_off[g] = c;
      for (size_t k = 0; k < go.contacts.size(); k++, c++) {
        const SolverContact& sc = go.contacts[k];
        out->shape_key[c] = sc.contact.shapeKey; out->feature_key[c] = sc.contact.featureKey;
        wr3(out->ni + 3 * c, sc.contact.ni); wr3(out->pi + 3 * c, sc.contact.pi); wr3(out->pj + 3 * c, sc.contact.pj);
        out->friction[c] = sc.friction; out->bounciness[c] = sc.bounciness; out->lambda[c] = go.lambda[k]; wr3(out->t1 + 3 * c, go.t1[k]);
      }
      g++;
    }
  }
  out->world_group_off[nw] = g;
  out->group_contact_off[g] = c;
  return EP_OK;
}

// Raw contacts of ONE shape pair in the list order the reference's `addContacts` returns (head first =
// last emitted first), for the narrowphase known-answer tests (tests/Collision/*.elm).
extern "C" int ep_oracle_shape_contacts(const ep_shapes* S, int shape1, const double* pos1, const double* quat1, int shape2,
                                        const double* pos2, const double* quat2, int32_t shape_key, int32_t cap, int64_t* feature_key,
                                        double* ni, double* pi, double* pj) {
  Tf t1{rd3(pos1), Quat{quat1[0], quat1[1], quat1[2], quat1[3]}}, t2{rd3(pos2), Quat{quat2[0], quat2[1], quat2[2], quat2[3]}};
  WShape a = placeShape(S, shape1, t1, 0, 0), b = placeShape(S, shape2, t2, 0, 0);
  std::vector<Contact> raw;
  addRawShapeContacts(shape_key, a, b, raw);
  if ((int)raw.size() > cap) return EP_ECAPACITY;
  int k = 0;
  for (int i = (int)raw.size() - 1; i >= 0; i--, k++) {
    feature_key[k] = raw[i].featureKey;
    wr3(ni + 3 * k, raw[i].ni); wr3(pi + 3 * k, raw[i].pi); wr3(pj + 3 * k, raw[i].pj);
  }
  return k;
}

// ConvexConvex.testSeparatingAxis / findSeparatingAxis / CapsuleConvex.supportFeature hooks for the known-answer tests.
extern "C" int ep_oracle_test_separating_axis(const ep_shapes* S, int shape1, const double* pos1, const double* quat1, int shape2,
                                              const double* pos2, const double* quat2, const double* axis, double* depth) {
  Tf t1{rd3(pos1), Quat{quat1[0], quat1[1], quat1[2], quat1[3]}}, t2{rd3(pos2), Quat{quat2[0], quat2[1], quat2[2], quat2[3]}};
  WShape a = placeShape(S, shape1, t1, 0, 0), b = placeShape(S, shape2, t2, 0, 0);
  return testSeparatingAxis(a.convex, b.convex, rd3(axis), *depth) ? 1 : 0;
}
extern "C" int ep_oracle_find_separating_axis(const ep_shapes* S, int shape1, const double* pos1, const double* quat1, int shape2,
                                              const double* pos2, const double* quat2, double* axis) {
  Tf t1{rd3(pos1), Quat{quat1[0], quat1[1], quat1[2], quat1[3]}}, t2{rd3(pos2), Quat{quat2[0], quat2[1], quat2[2], quat2[3]}};
  WShape a = placeShape(S, shape1, t1, 0, 0), b = placeShape(S, shape2, t2, 0, 0);
  V3 ax;
  if (!findSeparatingAxis(a.convex, b.convex, ax)) return 0;
  wr3(axis, ax);
  return 1;
}
extern "C" int ep_oracle_support_feature(const ep_shapes* S, int shape, const double* pos, const double* quat, const double* axis, double* out6) {
  Tf t{rd3(pos), Quat{quat[0], quat[1], quat[2], quat[3]}};
  WShape a = placeShape(S, shape, t, 0, 0);
  std::vector<V3> v = supportFeature(rd3(axis), a.convex);
  for (size_t i = 0; i < v.size(); i++) wr3(out6 + 3 * i, v[i]);
  return (int)v.size();
}

// Collision.ConvexConvex.project (ConvexConvex.elm:689-710) of a placed hull's convexVertices: out2 = {min, max}
extern "C" int ep_oracle_project(const ep_shapes* S, int shape, const double* pos, const double* quat, const double* axis, double* out2) {
  Tf t{rd3(pos), Quat{quat[0], quat[1], quat[2], quat[3]}};
  WShape a = placeShape(S, shape, t, 0, 0);
  projectVertices(rd3(axis), a.convex, out2);
  return EP_OK;
}

// Islands.fold visiting order for the known-answer test tests/SolverIslandsTest.elm:136-163.
extern "C" int ep_oracle_islands(int32_t n_groups, const int32_t* id1, const int32_t* kind1, const int32_t* id2, const int32_t* kind2,
                                 int32_t max_id, int32_t* out_root, int32_t* out_order) {
  std::vector<int> root, order;
  islandsOrder(n_groups, id1, kind1, id2, kind2, max_id, root, order);
  for (int g = 0; g < n_groups; g++) { out_root[g] = root[g]; out_order[g] = order[g]; }
  return EP_OK;
}
