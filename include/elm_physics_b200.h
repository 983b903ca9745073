/* elm_physics_b200.h — C ABI of the B200 stepping engine that replaces the hot path behind
 * `Physics.simulate` (reference: src/Physics.elm:522-583).
 *
 * The reference has no FFI; the seam this ABI replaces is the pair of internal calls
 *     BroadPhase.getPairs collide constrain sortedBodies        (src/Physics.elm:575-576)
 *     Solver.solve dt gravity iterations pairGroups maxId ...   (src/Physics.elm:578-579)
 * plus the gravity sort in front of them (src/Physics.elm:548-573).  Id assignment
 * (src/Internal/AssignIds.elm), unit unwrapping, user callbacks and `contactPoints` stay on
 * the caller's (Elm) side: the shim only marshals packed SoA arrays.
 *
 * Plain C, caller-allocated buffers, little-endian, fp64 / int32 / int64.  Every function
 * returns 0 on success or a negative EP_E* code; `ep_last_error` gives the text.  Nothing
 * throws or aborts.  One context per process/GPU, not re-entrant (the reference is
 * single-threaded).  The same structs are accepted by the CPU oracle (oracle/ep_oracle.h),
 * which is test infrastructure and is NOT linked into this library.
 */
#ifndef ELM_PHYSICS_B200_H
#define ELM_PHYSICS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { EP_OK = 0, EP_EINVAL = -1, EP_ECUDA = -2, EP_ECAPACITY = -3, EP_ENODEVICE = -4, EP_ESTATE = -5 };

/* Shape kinds: src/Internal/Shape.elm:27-32 */
enum { EP_CONVEX = 0, EP_PLANE = 1, EP_SPHERE = 2, EP_CAPSULE = 3, EP_PARTICLE = 4 };
/* Body kinds: src/Internal/Body.elm:22-34 */
enum { EP_STATIC = 1, EP_DYNAMIC = 2, EP_KINEMATIC = 3 };
/* Constraint kinds: src/Internal/Constraint.elm:9-13 */
enum { EP_POINT_TO_POINT = 0, EP_HINGE = 1, EP_LOCK = 2, EP_DISTANCE = 3 };

/* ---- Shape templates (immutable after construction, src/Internal/Body.elm:64-67) -----------
 * Centre-of-mass-frame shapes = `geometry.shapesWithMaterials` without the material
 * (src/Internal/Body.elm:68-72).  Identical shapes may be shared by many bodies.
 * Shape s owns f64[f64_off[s] .. f64_off[s+1]) and i32[i32_off[s] .. i32_off[s+1]).
 *
 *   sphere   f64: pos[3], radius                                   (src/Shapes/Sphere.elm:14-19)
 *   capsule  f64: pos[3], axis[3], radius, halfLength              (src/Shapes/Capsule.elm:22-29)
 *   plane    f64: pos[3], normal[3]                                (src/Shapes/Plane.elm:7-10)
 *   particle f64: pos[3]
 *   convex   f64: pos[3], ax[3], ay[3], az[3], he[3]   (Obb.Box payload, zeros when NotBox;
 *                                                       src/Shapes/Convex.elm:110-112)
 *                 verts[3*nv]                           (vertex buffer, key order; VertexBuffer.elm)
 *                 groups[8*ng]: normal[3], faceDist, partnerNormal[3], partnerDist
 *                                                      (CoM-frame list order; the step walks it
 *                                                       REVERSED, src/Shapes/Convex.elm:265-287)
 *            i32: is_box, nv, ng, ne,
 *                 ng x { two_sided, n1, off1, n2, off2 }   face vertex-index lists
 *                 ne x { n, off }                          uniqueEdges groups, flat endpoint pairs,
 *                                                          stored order (src/Shapes/Convex.elm:173)
 *                 index data (offsets are relative to the shape's i32 block)
 */
typedef struct ep_shapes {
  int32_t n_shapes;
  const int32_t* kind;     /* [n_shapes] */
  const int32_t* f64_off;  /* [n_shapes+1] */
  const int32_t* i32_off;  /* [n_shapes+1] */
  const double* f64;
  const int32_t* i32;
} ep_shapes;

enum { EP_CONVEX_F64_HEADER = 15, EP_CONVEX_I32_HEADER = 4, EP_GROUP_F64 = 8, EP_GROUP_I32 = 5, EP_EDGE_I32 = 2 };

/* ---- Bodies: SoA over all worlds, in the caller's (user list) order -------------------------
 * One row per `Internal.Body.Body` (src/Internal/Body.elm:36-61).  World w owns rows
 * [world_body_off[w], world_body_off[w+1]).  In/out fields are overwritten by ep_step /
 * ep_download_bodies with the integrated state (src/Internal/SolverBody.elm:97-225).
 */
typedef struct ep_bodies {
  int32_t n_worlds;
  int32_t n_bodies;
  const int32_t* world_body_off;   /* [n_worlds+1] */
  const int32_t* body_id;          /* body.id, assigned caller-side, unique within a world, < 2^18 */
  const int32_t* kind;             /* EP_STATIC / EP_DYNAMIC / EP_KINEMATIC */
  double* pos;                     /* [n][3]  transform3d origin = centre of mass, world   (in/out) */
  double* quat;                    /* [n][4]  transform3d orientation x,y,z,w              (in/out) */
  double* vel;                     /* [n][3]                                              (in/out) */
  double* angvel;                  /* [n][3]                                              (in/out) */
  double* force;                   /* [n][3]  cleared for dynamic+kinematic                (in/out) */
  double* torque;                  /* [n][3]                                              (in/out) */
  double* inv_inertia_world;       /* [n][9]  m11,m21,m31,m12,m22,m32,m13,m23,m33          (in/out) */
  const double* inv_mass;          /* [n] */
  const double* inv_inertia;       /* [n][3]  diagonal, CoM frame */
  const double* lin_damp_pow;      /* [n]  (1-linearDamping)^dt, computed caller-side with the host
                                           language's pow (src/Internal/SolverBody.elm:149-153) */
  const double* ang_damp_pow;      /* [n] */
  const double* lin_lock;          /* [n][3] 0/1 masks */
  const double* ang_lock;          /* [n][3] */
  const double* bounding_radius;   /* [n]  geometry.boundingSphereRadius */
  const int32_t* inst_off;         /* [n+1] body b owns shape instances [inst_off[b], inst_off[b+1])
                                            in geometry.shapesWithMaterials order (1-based shape ids) */
  int32_t n_insts;
  const int32_t* inst_shape;       /* [n_insts] index into ep_shapes */
  const double* inst_friction;     /* [n_insts] material of the (shape, material) pair */
  const double* inst_bounciness;   /* [n_insts] */
} ep_bodies;

/* ---- Per-world step parameters (src/Physics.elm:625-632) and callback results ---------------
 * `collide`: NULL == always true.  Otherwise for world w an n_w x n_w byte matrix at
 * collide[collide_off[w] + i*n_w + j] = collide id_i id_j, i/j = positions in the user list
 * (evaluated caller-side for every ordered pair; src/Internal/BroadPhase.elm:189-215).
 * Constraints: one row per `constrain idA` result entry toward body idB
 * (src/Internal/BroadPhase.elm:144-186), already converted to centre-of-mass frames
 * (src/Internal/Constraint.elm:16-47; each pivot/axis with its own body's frame).  Rows of one
 * (a -> b) registration are contiguous and in list order; the engine picks forward or flipped
 * use from the gravity order exactly as constraintsBetween does.
 */
typedef struct ep_params {
  int32_t n_worlds;
  const double* gravity;           /* [n_worlds][3] */
  const double* dt;                /* [n_worlds] */
  const int32_t* iterations;       /* [n_worlds] solverIterations */
  const uint8_t* collide;          /* optional */
  const int64_t* collide_off;      /* [n_worlds] when collide != NULL */
  int32_t n_constraints;
  const int32_t* con_body_a;       /* [n_constraints] global body row of the registering body */
  const int32_t* con_body_b;       /* [n_constraints] global body row of the partner */
  const int32_t* con_kind;         /* EP_POINT_TO_POINT.. */
  const double* con_data;          /* [n_constraints][24]: a-side pivot,x/axis,y,z then b-side
                                      pivot,x/axis,y,z (unused slots 0); distance in [0] */
} ep_params;

/* ---- Contacts: pair groups + solved contacts = `pairGroups` and the warm-start cache --------
 * (src/Solver.elm:215-220, src/Physics/Types.elm:19-25, src/Internal/ContactCache.elm).
 * Groups of world w: [world_group_off[w], world_group_off[w+1]) in pair-enumeration (CSR)
 * order, i.e. the REVERSE of the reference's `pairGroups` list (src/BroadPhase.elm:96-109).
 * Contacts of group g: [group_contact_off[g], group_contact_off[g+1]) in SOLVER order
 * (src/Internal/Equation.elm:135-154), i.e. the reverse of `pairGroup.contacts`.
 * `lambda` = solved normal lambda, `t1` = friction-1 direction (src/Solver.elm:320-335).
 * As INPUT only world_group_off, group_id1/2, group_contact_off, shape_key, feature_key, lambda,
 * t1 are read.  Capacities are in *_cap; overflow -> EP_ECAPACITY, nothing usable written.
 * As an OUTPUT every array pointer may be NULL: that array is then not copied back.  A simulate loop that keeps the batch
 * resident (ep_step_frame) needs on the host only what `Physics.contactPoints` reads (src/Physics.elm:1158-1207:
 * group_id1/2, group_contact_off, group_pre_pos1/quat1, pi) -- the rest of `Contacts` (warm-start lambdas, keys, normals)
 * is consumed by the next step only, and that step reads the resident frame.
 */
typedef struct ep_contacts {
  int32_t n_worlds;
  int32_t group_cap;
  int32_t contact_cap;
  int32_t* world_group_off;        /* [n_worlds+1] */
  int32_t* group_id1;              /* body ids (body1 = lower gravity-sort position) */
  int32_t* group_id2;
  int32_t* group_n_constraints;    /* joint constraints carried by the group */
  int32_t* group_contact_off;      /* [n_groups+1] */
  double* group_pre_pos1;          /* [n_groups][3] pre-step transform3d of body1 (for contactPoints,
                                      src/Physics.elm:1180-1186) */
  double* group_pre_quat1;         /* [n_groups][4] */
  int32_t* shape_key;              /* s1*256+s2 (src/Internal/ContactId.elm:75-77) */
  int64_t* feature_key;            /* < 2^48 (src/Internal/ContactId.elm:94-96) */
  double* ni;                      /* [n][3] */
  double* pi;                      /* [n][3] */
  double* pj;                      /* [n][3] */
  double* friction;                /* [n] */
  double* bounciness;              /* [n] */
  double* lambda;                  /* [n] */
  double* t1;                      /* [n][3] */
  int32_t* iterations;             /* [n_worlds] Contacts.iterations (src/Solver.elm:275-276) */
  int32_t* world_n_islands;        /* [n_worlds] */
  int32_t* group_island;           /* [n_groups] island root body id (src/Internal/Islands.elm) */
} ep_contacts;

typedef struct ep_ctx ep_ctx;

/* Create a context on CUDA device `device`.  Fails with EP_ENODEVICE when no B200-class GPU /
 * CUDA runtime is usable: there is no CPU fallback. */
int ep_create(ep_ctx** out, int device);
void ep_destroy(ep_ctx* ctx);
const char* ep_last_error(const ep_ctx* ctx);

/* Tuning knobs, read by the next ep_upload_worlds.  EP_OPT_PATH: reserved (one device path).
 * EP_OPT_CONTACT_CAP: device contact pool size (default max(16384, 16 x bodies)).  EP_OPT_RAW_CAP: reserved.
 * EP_OPT_ITEM_CAP: narrowphase work items (shape pairs of candidate body pairs) per step (default max(65536, 8 x shape
 * instances)).  Exceeding a capacity is reported as EP_ECAPACITY by ep_sync / the downloads, never silently. */
enum { EP_OPT_PATH = 0, EP_OPT_CONTACT_CAP = 1, EP_OPT_RAW_CAP = 2, EP_OPT_ITEM_CAP = 3, EP_OPT_COUNT = 4 };
int ep_configure(ep_ctx* ctx, int32_t option, int64_t value);

/* Upload the immutable shape table (replaces nothing in the reference: shapes are Elm values
 * captured in Body.geometry, src/Internal/Body.elm:68-72). */
int ep_upload_shapes(ep_ctx* ctx, const ep_shapes* shapes);

/* Make a batch of worlds resident: bodies, params and (optionally, may be NULL = emptyContacts,
 * src/Physics.elm:658-665) the previous frame's contacts as warm start. */
int ep_upload_worlds(ep_ctx* ctx, const ep_bodies* bodies, const ep_params* params, const ep_contacts* warm);

/* Advance every resident world by n_steps simulate calls, feeding each frame's contacts back as
 * the next frame's warm start (README.md:21-28).  Forces/torques are cleared by the first step
 * exactly as SolverBody.solved does.  Asynchronous on the context's stream. */
int ep_step_resident(ep_ctx* ctx, int32_t n_steps);

/* Block until the stream is idle. */
int ep_sync(ep_ctx* ctx);

/* Copy the integrated state (in/out fields) / the last frame's contacts back. */
int ep_download_bodies(ep_ctx* ctx, ep_bodies* bodies);
int ep_download_contacts(ep_ctx* ctx, ep_contacts* contacts);

/* One-call form = getPairs + solve for every world: upload, n_steps, download. */
int ep_step(ep_ctx* ctx, ep_bodies* bodies, const ep_params* params, const ep_contacts* warm_in,
            ep_contacts* out, int32_t n_steps);

/* One frame of the RESIDENT batch through host buffers = one `simulate` call of a loop that feeds the returned contacts
 * back (README.md:21-28): H2D of every per-body f64 array of `bodies` (the caller may have applied forces / impulses /
 * moved bodies between frames, src/Physics.elm:857-1017), one step with the resident contacts of the previous frame as
 * the warm start, D2H of the integrated state and (if `out` is not NULL) of the frame's contacts.  The topology (worlds,
 * body ids and kinds, shape instances, params, constraints) must be the one of the last ep_upload_worlds; use ep_step
 * when it changes or when the warm start is not the previous frame.
 * Pass pinned buffers (ep_host_alloc): the call is a pipeline over three streams -- the bulk of the input is copied in beside
 * the front kernels, and when `out` asks for no solver result (lambda, t1, iterations, world_n_islands NULL) the frame's pair
 * groups and contact points are marshalled and copied back beside the solver.  Synchronous: every buffer is complete on
 * return.  Contexts of one device may be driven from different host threads concurrently (one thread per context). */
int ep_step_frame(ep_ctx* ctx, ep_bodies* bodies, ep_contacts* out);

/* The same frame moving only the arrays that changed hands.  `upload`: the per-body arrays the caller modified since the
 * last frame (the resident copy of the others IS what the last frame returned; constants -- inverse mass and inertia, damping
 * powers, lock masks, bounding radius -- came with ep_upload_worlds).  A simulate loop that only reads the bodies sends nothing
 * or the dynamic state (EP_F_STATE); one that pushed bodies (src/Physics.elm:857-1017) adds EP_F_FORCE | EP_F_TORQUE on those
 * frames; one that moved / rotated bodies adds EP_F_INV_INERTIA_WORLD (Physics.elm:291-303 recompute it caller-side).
 * `download`: the arrays to copy back.  force / torque of every non-static body are zero after a step
 * (src/Internal/SolverBody.elm:139-141, 221-223) and invInertiaWorld is a function of the returned orientation
 * (Transform3d.elm:181-205), so EP_F_STATE is everything `Physics.simulate`'s caller cannot derive itself.
 * ep_step_frame == ep_step_frame_fields(.., EP_F_ALL, EP_F_ALL). */
enum { EP_F_POS = 1, EP_F_QUAT = 2, EP_F_VEL = 4, EP_F_ANGVEL = 8, EP_F_FORCE = 16, EP_F_TORQUE = 32, EP_F_INV_INERTIA_WORLD = 64,
       EP_F_CONSTANTS = 128, EP_F_STATE = 15, EP_F_ALL = 255 };
int ep_step_frame_fields(ep_ctx* ctx, ep_bodies* bodies, ep_contacts* out, uint32_t upload, uint32_t download);

/* `Physics.raycast` (src/Physics.elm:802-841) for a batch of rays against the RESIDENT worlds in their current state (after
 * the last step / upload): ray r runs against the bodies of world ray_world[r] in list order; closest hit over every
 * shape of every body (src/Internal/Body.elm:465-489, src/Internal/Shape.elm:132-148), first hit wins ties, point masses
 * are never hit, planes only from their front side.  hit_body[r] = global body row (index into ep_bodies) or -1;
 * hit_normal is normalised (Vec3.normalize).  Host buffers, synchronous. */
int ep_raycast(ep_ctx* ctx, int32_t n_rays, const int32_t* ray_world, const double* from /*[n][3]*/, const double* direction /*[n][3]*/,
               int32_t* hit_body, double* hit_distance, double* hit_point /*[n][3]*/, double* hit_normal /*[n][3]*/);

/* The same for rays ATTACHED to bodies -- the RaycastCar pattern (examples/src/RaycastCar/Car.elm:116-128: a ray through a point of
 * the chassis along its down direction, `Axis3d.placeIn frame`, cast against the bodies WITHOUT the car, four per car and frame):
 * ray r is given in the centre-of-mass frame of global body row ray_body[r] and placed in the world on the device with that
 * body's current transform3d (Transform3d.pointPlaceIn for the origin, directionPlaceIn for the direction,
 * src/Internal/Transform3d.elm:134-153, 218-220), so a resident batch needs no state download between steps; it runs against
 * the other bodies of that body's world (its own body is skipped). */
int ep_raycast_attached(ep_ctx* ctx, int32_t n_rays, const int32_t* ray_body, const double* local_from /*[n][3]*/, const double* local_direction /*[n][3]*/,
                        int32_t* hit_body, double* hit_distance, double* hit_point /*[n][3]*/, double* hit_normal /*[n][3]*/);

/* The 22 doubles per body a step writes (src/Internal/SolverBody.elm:97-225: transform3d 7, velocity 3, angularVelocity 3,
 * invInertiaWorld 9) packed [n_bodies][22] into DEVICE memory of the caller, on `stream` (a cudaStream_t; NULL = the context's
 * stream), ordered behind the steps already enqueued: the send buffer of the one collective of a batch sharded over several
 * GPUs -- the final state gather (ncclAllGather over NVLink) -- with no host staging.  Asynchronous. */
int ep_pack_state_device(ep_ctx* ctx, void* dst_device, void* stream);

/* Between-step body mutators on the RESIDENT state, so that a multi-step batch can push bodies without a host round trip
 * (src/Physics.elm:857-1017 over src/Internal/Body.elm:305-430).  Operation k acts on global body row body_row[k]; operations
 * on the same body are applied in list order.  a / b per kind:
 *   EP_APPLY_FORCE (force a at world point b), EP_APPLY_IMPULSE (impulse a at world point b), EP_APPLY_TORQUE (a),
 *   EP_APPLY_ANGULAR_IMPULSE (a): dynamic bodies only, others unchanged;
 *   EP_SET_VELOCITY (a), EP_SET_ANGULAR_VELOCITY (a): every body that is not static. */
enum { EP_APPLY_FORCE = 0, EP_APPLY_IMPULSE = 1, EP_APPLY_TORQUE = 2, EP_APPLY_ANGULAR_IMPULSE = 3, EP_SET_VELOCITY = 4,
       EP_SET_ANGULAR_VELOCITY = 5 };
int ep_apply(ep_ctx* ctx, int32_t n_ops, const int32_t* kind, const int32_t* body_row, const double* a /*[n][3]*/, const double* b /*[n][3]*/);

/* Page-locked host memory for the buffers of ep_bodies / ep_contacts (any host memory works; pinned buffers copy at
 * full PCIe rate and let the copies run asynchronously).  NULL when the allocation fails. */
void* ep_host_alloc(size_t bytes);
void ep_host_free(void* p);

/* Introspection used by bench.py / tests. */
int ep_stats(ep_ctx* ctx, int64_t* out, int32_t n);   /* see EP_STAT_* */
enum { EP_STAT_KERNEL_LAUNCHES = 0, EP_STAT_CANDIDATE_PAIRS = 1, EP_STAT_GROUPS = 2, EP_STAT_CONTACTS = 3,
       EP_STAT_ISLANDS = 4,
       /* cumulative since ep_upload_worlds: sum over islands of contacts x PGS iterations run, joint rows x iterations run */
       EP_STAT_CONTACT_ITERATIONS = 5, EP_STAT_JOINT_ROW_ITERATIONS = 6,
       /* cumulative: early-exit decisions of the multi-lane island solvers that were taken from the re-summation of
          |deltaLambda| in the reference's association (src/Internal/Solver.elm:345-367) instead of the parallel sum */
       EP_STAT_RESUMS = 7, EP_STAT_COUNT = 8 };
/* Per-kernel device timing.  ep_profile(ctx, 1) makes ep_step_resident bracket every kernel launch with
 * CUDA events on the context's stream and resets the accumulators; ep_kernel_times synchronises and returns
 * the accumulated milliseconds and launch counts per kernel kind (EP_K_*).  Launch counts are always kept. */
enum { EP_K_PLACE = 0, EP_K_SORT = 1, EP_K_BROADPHASE = 2, EP_K_NARROWPHASE = 3, EP_K_EQUATIONS = 4, EP_K_ISLANDS = 5,
       EP_K_SOLVE = 6, EP_K_INTEGRATE = 7, EP_K_SOLVE_CLUSTER = 8,   /* giant islands on a thread-block cluster (slot of the former k_tile_build) */
       /* the solver's concurrent launches by island size class (they overlap; EP_K_SOLVE brackets all of them) */
       EP_K_SOLVE_LANES_B0 = 9, EP_K_SOLVE_LANES_B1 = 10, EP_K_SOLVE_LANES_B2 = 11, EP_K_SOLVE_LANES_B3 = 12, EP_K_SOLVE_LANES_B4 = 13,
       EP_K_SOLVE_LANES_B5 = 14, EP_K_SOLVE_LANES_INPLACE = 15, EP_K_SOLVE_TEAM_WARP = 16, EP_K_SOLVE_TEAM_CTA = 17, EP_K_COUNT = 18 };
int ep_profile(ep_ctx* ctx, int enable);
int ep_kernel_times(ep_ctx* ctx, double* ms_out, int64_t* launches_out, int32_t n);
const char* ep_kernel_name(int32_t kernel);
/* Raw CUDA stream handle (cudaStream_t) so callers can record CUDA events on it. */
void* ep_stream(ep_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
