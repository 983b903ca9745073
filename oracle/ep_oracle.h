/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * CPU restatement (C++17, -O2 -ffp-contract=off) of the reference's algorithm for the path behind
 * Physics.simulate, on the SAME packed buffers as the product ABI (include/elm_physics_b200.h).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * Parity status: narrowphase / islands / id assignment are pinned by the reference's own known-answer
 * tests (tests/test_oracle_golden.py ports them); whole-step fp64 trajectories are PARITY UNPINNED —
 * the reference ships no golden trajectories and cannot be built here (no elm / node), see DESIGN.md. */
#ifndef EP_ORACLE_H
#define EP_ORACLE_H
#include "../include/elm_physics_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* Same contract as ep_step (getPairs + solve + integrate for every world, n_steps times). */
int ep_oracle_step(const ep_shapes* shapes, ep_bodies* bodies, const ep_params* params, const ep_contacts* warm_in,
                   ep_contacts* out, int32_t n_steps);
/* One Collision.<A><B>.addContacts call; returns the contact count (list order of the reference). */
int ep_oracle_shape_contacts(const ep_shapes* shapes, int shape1, const double* pos1, const double* quat1, int shape2,
                             const double* pos2, const double* quat2, int32_t shape_key, int32_t cap, int64_t* feature_key,
                             double* ni, double* pi, double* pj);
int ep_oracle_test_separating_axis(const ep_shapes* shapes, int shape1, const double* pos1, const double* quat1, int shape2,
                                   const double* pos2, const double* quat2, const double* axis, double* depth);
int ep_oracle_find_separating_axis(const ep_shapes* shapes, int shape1, const double* pos1, const double* quat1, int shape2,
                                   const double* pos2, const double* quat2, double* axis);
int ep_oracle_support_feature(const ep_shapes* shapes, int shape, const double* pos, const double* quat, const double* axis,
                              double* out6);
int ep_oracle_project(const ep_shapes* shapes, int shape, const double* pos, const double* quat, const double* axis, double* out2);
int ep_oracle_islands(int32_t n_groups, const int32_t* id1, const int32_t* kind1, const int32_t* id2, const int32_t* kind2,
                      int32_t max_id, int32_t* out_root, int32_t* out_order);
/* Physics.raycast (src/Physics.elm:802-841) for a batch of rays against the CURRENT state of `bodies`; same contract as ep_raycast. */
int ep_oracle_raycast(const ep_shapes* shapes, const ep_bodies* bodies, int32_t n_rays, const int32_t* ray_world, const double* from,
                      const double* direction, int32_t* hit_body, double* hit_distance, double* hit_point, double* hit_normal);
/* Body mutators on host buffers, in list order; same contract as ep_apply. */
int ep_oracle_apply(ep_bodies* bodies, int32_t n_ops, const int32_t* kind, const int32_t* body_row, const double* a, const double* b);
#ifdef __cplusplus
}
#endif
#endif
