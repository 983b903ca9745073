// ORACLE — TEST INFRASTRUCTURE ONLY (see ep_oracle_math.h).
#pragma once
#include "ep_oracle_shapes.h"

namespace epo {

// src/Internal/Contact.elm:22-33
struct Contact { int shapeKey; int64_t featureKey; V3 ni, pi, pj; };
// src/Internal/Contact.elm:16-20
struct SolverContact { double friction, bounciness; Contact contact; };

struct MinMax;
void addRawShapeContacts(int shapeKey, const WShape& s1, const WShape& s2, std::vector<Contact>& out);
void getContactsSolverOrder(const std::vector<WShape>& shapes1, const std::vector<WShape>& shapes2, std::vector<SolverContact>& out);
bool testSeparatingAxis(const WConvex& c1, const WConvex& c2, V3 axis, double& depth);
bool findSeparatingAxis(const WConvex& c1, const WConvex& c2, V3& axis);
std::vector<V3> supportFeature(V3 axis, const WConvex& cv);
void projectVertices(V3 axis, const WConvex& c, double* out2);

}  // namespace epo
