// Device-resident data of one context (all worlds concatenated) — see DESIGN.md "Data layout in HBM".
#pragma once
#include <stdint.h>

namespace ep {

struct RawContact;

// One solved contact's geometry + identity: `Contact`/`SolverContact` (src/Internal/Contact.elm:16-33).
struct ContactRec {
  int64_t fk;            // featureKey < 2^48 (ContactId.elm:94-96)
  int32_t sk;            // shapeKey = s1*256+s2 (ContactId.elm:75-77)
  int32_t slot;          // owning pair slot
  double ni[3], pi[3], pj[3];
  double friction, bounciness;
  double lam;            // normal lambda: the 0.85-scaled seed after k_equations, the solved value after the solver
  double t1[3];          // friction-1 direction of this frame (= vB of the friction-1 row): with lam, all that the NEXT frame's
                         // warm start reads of this contact's solution (Solver.elm:320-335)
};

// `ContactEquations` + `ContactData` (src/Internal/Equation.elm:63-93): three Jacobian rows (wA, vB, wB;
// vA = -vB) and their Spook scalars; min/max normal impulse are the constants 0 / 1e6.
// Spook eps is the same for every row of a world (it depends on dt only) and is recomputed where needed.
// xIa / xIb = invInertiaWorld(body1) * wA and invInertiaWorld(body2) * wB of each row (zero for a non-dynamic body): the
// reference forms them on every row of every iteration (Solver.elm:505-550); they are constant during the solve and
// computed once, with the same operations in the same order, when the row is built.
struct RowRec {
  double nJ[9], nB, nInvC, lamN;                                        // sweep 1 reads these 12 doubles (+ nIa nIb)
  double f1J[9], f2J[9], f1B, f1InvC, f2B, f2InvC, mu, lamF1, lamF2;   // sweep 2 reads these 25 (+ lamN, f*Ia f*Ib)
  double dN, dF1, dF2;   // |deltaLambda| of the last iteration, for the canonical-order sum of the team solver
  double nIa[3], nIb[3], f1Ia[3], f1Ib[3], f2Ia[3], f2Ib[3];
};                       // 58 doubles = 464 B (16 B aligned)

// The same row in the layout of the QUAD solver (ep_solve_quads.cuh), which solves a row with four lanes; lane role q holds
// one half of one body of the row's pair:
//   q = 0  body1 linear   a = -vB   c = a        m = invMass1        (vA = -vB, Equation.elm:22-35)
//   q = 1  body1 angular  a =  wA   c = I1 wA    m = 1
//   q = 2  body2 linear   a =  vB   c = a        m = invMass2
//   q = 3  body2 angular  a =  wB   c = I2 wB    m = 1
// The row's G.W is ((p0 + p1) + p2) + p3 with p_q = (a0 s0 + a1 s1) + a2 s2 over the lane's half s of the solver body --
// bit for bit the association of Solver.elm:642-646 (negating every product of the first dot negates its rounded sum) --
// and the lane's update is s += (dlambda m) c, the same operations as applyVelocityBody1/2 (Solver.elm:505-550).
// Contact c owns rows 3c (normal), 3c+1 (friction 1), 3c+2 (friction 2); joint rows live in their own array.
struct __align__(16) QRow {
  double a[3][4];        // [component][role]
  double c[3][2];        // [component][0: I1 wA, 1: I2 wB]
  double B, invC;
  double lam, mu;        // mu: friction rows only (bounds +-mu * lambda of the contact's normal row)
  int32_t i0, i1, i2, i3;   // free in global memory; the solver's staged copy keeps its row descriptor here
};
static_assert(sizeof(QRow) == 192, "QRow is 12 chunks of 16 B");

// `ConstraintEquation` (src/Internal/Equation.elm:503-511); bounds are the constants -1e6 / 1e6.
struct JointRow { double J[9]; double B, invC, eps, lam; double dl; double Ia[3], Ib[3]; double pad[4]; };   // 24 doubles; dl = |deltaLambda| of the last iteration

constexpr int kIterKeys = 128, kGroupKeys = 4;     // (16 size sub-classes x 8 iteration classes) x 4 group-count classes per size class
constexpr int kBins = 10;    // island size classes for the solver's work list (ep_solve.cuh)
constexpr int kKeys = kBins * kIterKeys * kGroupKeys;
constexpr int kClasses = 36; // narrowphase work-item classes: unordered pair of shape sub-kinds, lo * 6 + hi (ep_kernels.cuh)
constexpr int kRawStride = 8;// raw contacts reserved per work item (plane-box emits up to 8)
enum { CNT_CAND = 0, CNT_CONTACTS = 1, CNT_ISLANDS = 2, CNT_ISL_GROUPS = 3, CNT_JROWS = 4, CNT_GROUPS = 5, CNT_WORK = 6, CNT_ITEMS = 7,
       CNT_BIN0 = 8, CNT_CLASS0 = CNT_BIN0 + kBins, CNT_COUNT = CNT_CLASS0 + kClasses * 8 /* x depth bins, ep_kernels.cuh */ };
enum { FLAG_POOL_OVERFLOW = 0, FLAG_PAIR_OVERFLOW = 1, FLAG_POLY_OVERFLOW = 2, FLAG_JOINT_OVERFLOW = 3, FLAG_COUNT = 4 };

// The last frame's contacts in the layout of ep_contacts, compacted on the device (ep_output.cuh)
struct OutBuf {
  int32_t group_cap, contact_cap;
  int32_t *world_group_off, *group_id1, *group_id2, *group_ncon, *group_contact_off, *group_island, *iterations, *world_n_islands;
  double *group_pre_pos1, *group_pre_quat1;
  int32_t* shape_key; int64_t* feature_key; double *ni, *pi, *pj, *friction, *bounciness, *lambda, *t1;
  int32_t *wg_cnt, *wc_cnt, *wc_off;   // [n_worlds] valid groups / contacts of each world, first contact of each world
  int32_t* totals;                     // [2] groups, contacts of the frame
};

struct FrameBuf {        // double-buffered: the previous frame is the warm-start cache
  ContactRec* contacts;
  int32_t* slot_row1;    // pair slots: body rows (body1 = lower gravity-sort position)
  int32_t* slot_row2;
  int32_t* slot_cstart;  // contacts of the slot, solver order
  int32_t* slot_ccount;
  int32_t* slot_jstart;  // joint rows of the slot, solver order
  int32_t* slot_jcount;
  int32_t* slot_ncon;    // user constraints carried by the pair
  int32_t* slot_root;    // island root body row (-1 when the slot is not a pair group)
  int32_t* slot_group;   // index of the slot's pair group in the compacted output (ep_output.cuh), -1 when it is none
  int32_t* slot_item0;   // first narrowphase work item (shape pair) of the slot; items are numbered shape1-major
  int32_t* slot_nitems;  // shapes1 x shapes2 when the pair passed the bounding-sphere test, else 0
  int32_t* pair_slot;    // [slot_cap] dense (world, i<j body pair) -> slot of this frame, -1 = no contacts (the warm-start cache index)
  int32_t* cand_base;    // [n_worlds] first slot of the world (enumeration order inside)
  int32_t* cand_n;       // [n_worlds]
  int32_t* counters;     // [CNT_COUNT]
};

struct Dev {
  // shape templates (ep_shapes)
  const int32_t* sh_kind; const int32_t* sh_f64_off; const int32_t* sh_i32_off; const double* sh_f64; const int32_t* sh_i32;
  // bodies (ep_bodies)
  int32_t n_worlds, n_bodies, n_insts;
  const int32_t* world_pair_off;   // [n_worlds] first entry of the world in the pair tables (sum of n(n-1)/2 of the worlds before)
  const int32_t* world_body_off; const int32_t* body_id; const int32_t* kind; const int32_t* body_world;
  double *pos, *quat, *vel, *angvel, *force, *torque, *iiw;
  const double *inv_mass, *inv_inertia, *ldp, *adp, *lin_lock, *ang_lock, *radius;
  const int32_t* inst_off; const int32_t* inst_shape; const int32_t* inst_body;
  const int32_t* inst_sub;    // [n_insts] narrowphase sub-kind of the instance's shape (box 0, other hull 1, plane 2, sphere 3, capsule 4, particle 5): fixed per batch
  const double* inst_friction; const double* inst_bounciness;
  const int64_t* inst_woff;   // world-placed f64 block of each instance (same layout as its template)
  double* wf64;
  double *pre_pos, *pre_quat; // transform3d before the last integration (contactPoints, Physics.elm:1180-1186)
  // placement work items: (instance, rel offset*2 + isDirection)
  const int32_t* pl_inst; const int32_t* pl_code; int32_t n_pl_moving; int32_t n_pl_all;
  // broadphase chunks: consecutive pair-index ranges of a world (ep_kernels.cuh k_bp_*)
  int32_t n_chunks; const int32_t* chunk_world; const int32_t* chunk_p0; const int32_t* world_chunk_off;
  int32_t *chunk_cnt, *chunk_icnt, *chunk_base, *chunk_ibase;
  // params (ep_params)
  const double* gravity; const double* dt; const int32_t* iterations;
  const uint8_t* collide; const int64_t* collide_off;
  // constraint registrations sorted by (a row, b row): CSR over a
  int32_t n_constraints; const int32_t* con_a_off; const int32_t* con_b; const int32_t* con_kind; const double* con_data;
  // scratch
  int32_t* sorted;            // [n_bodies] world base + gravity-sort position -> body row
  double* sort_key;           // [n_bodies]
  double* sbv;                // [n_bodies][6] solver-body delta velocities (SolverBody.elm:16-25)
  int32_t* uf_parent;         // [n_bodies] union-find over world-local body indices
  int32_t* uf_count;          // [n_bodies]
  int32_t* uf_fill;           // [n_bodies]
  int32_t* isl_off; int32_t* isl_cnt; int32_t* isl_world; int32_t* isl_groups;
  int4* sched_rec;            // [2 * slot_cap] TeamRec of every group of a team island, schedule order (ep_solve.cuh)
  int32_t *sched, *sched_lvl, *lvl_off, *lvl_cur;   // [slot_cap] level schedule of each island's groups (k_solve_team), parallel to isl_groups
  int32_t* world_min_rem;     // [n_worlds] min over islands of remaining iterations
  int32_t* world_n_islands;   // [n_worlds]
  // rows of this frame (scratch of the step: the next frame reads only ContactRec.lam / t1).  Two solver families, chosen per
  // batch at ep_upload_worlds (Dev::quads): lane tiles read RowRec / JointRow, the quad solver QRow.
  int32_t quads;
  int32_t sched_serial;       // EP_SCHED_SERIAL=1 (test / A-B aid): k_solve_cluster builds its level schedule with the one-thread walk
  RowRec* rows;               // [contact_cap]                         (lane solver)
  JointRow* jrows;            // [jrow_cap]
  QRow* qrows;                // [3 * contact_cap] rows 3c, 3c+1, 3c+2  (quad solver)
  QRow* jqrows;               // [jrow_cap]
  double* row_dl;             // [3 * contact_cap + jrow_cap] |deltaLambda| of the last iteration (team of quads: canonical re-sum)
  double* team_body;          // [(n_bodies + n_worlds + 1)][8] body table of the giant-island quad team when it does not fit shared memory
  int32_t* body_local;        // [n_bodies] 1-based index of a dynamic body inside its island (0: in no island)
  int32_t* flags;             // [FLAG_COUNT]
  unsigned long long* work;   // [4] cumulative contact-iterations, joint-row-iterations of the solver (roofline denominators), canonical re-sums
  int64_t slot_cap; int32_t contact_cap; int32_t jrow_cap;
  // narrowphase work items, binned by class so that the lanes of a warp run the same Collision.* module
  int32_t item_cap;
  int4* items;                // [item_cap + 32 kClasses] (slot, instance a, instance b, item id), by class (32-aligned) and depth bin
  int32_t *cls_off, *cls_cur; // [kClasses * 8] first item / fill cursor of every (class, depth bin)
  int32_t *np_first, *np_end; // [kClasses + 1] / [kClasses] item range of every class in kClassOrder order
  struct RawContact* raw;     // [item_cap][kRawStride] contacts of each item, solver order
  int32_t* raw_cnt;           // [item_cap]
  int32_t* sep_hint;          // [slot_cap] dense (world, i<j body pair) -> separating axis of last frame (ep_kernels.cuh), 0 = none; persistent
  int2* contact_src;          // [contact_cap] (slot, raw contact index) of every contact of the frame (k_slot_counts -> k_equations)
  // solver work list: islands sorted by (size class, predicted iterations, group count), largest / longest first, so that the
  // 32 islands of a lane-per-island tile run about the same number of rows for about the same number of iterations
  int32_t* isl_key;           // [n_bodies] sort key of each island
  int32_t* isl_nc;            // [n_bodies] contacts of the island
  int32_t* isl_nb;            // [n_bodies] dynamic bodies of the island
  int32_t* isl_nj;            // [n_bodies] joint rows of the island
  int32_t* isl_root;          // [n_bodies] root body row (Islands.elm:39-46)
  int32_t* isl_sorted;        // [n_bodies] island indices in work-list order
  int32_t* key_cnt;           // [kKeys] histogram, [kKeys..2*kKeys) scatter cursors
  int32_t* body_iters;        // [n_bodies] iterations the island rooted at this body used last frame (the predictor)
  // tile images of the lane solver that do not fit shared memory (rows of the large lane classes; ep_solve.cuh)
  double* tile_arena; int64_t tile_arena_cap;   // in doubles
  unsigned long long* tile_top;                 // bump pointer (doubles)
  int32_t tile_smem[6];       // bytes of dynamic shared memory of the launch that serves each lane class
  int32_t nb_max;             // largest world (bodies)
  const int32_t* sh_pl_off;   // [n_shapes+1] placement program of each shape template
  const int32_t* sh_pl_code;  // rel offset * 2 + isDirection
  FrameBuf cur, prev;
  OutBuf out;
};

}  // namespace ep
