// ep_napi.cc — the Node-API addon between the compiled Elm (`Physics.simulate`, /root/reference/src/Physics.elm:522-583) and the
// C ABI of include/elm_physics_b200.h.  PURE MARSHALLING: every function below turns JS typed arrays into the plain pointers of
// ep_shapes / ep_bodies / ep_params / ep_contacts and calls one entry point; no arithmetic on physics data happens here.  (The
// one exception the header asks of the caller, `(1 - damping) ^ dt` with the HOST language's pow, is done by the JS side of the
// shim -- patch_simulate.js -- with Math.pow, exactly as src/Internal/SolverBody.elm:149-153 does.)
//
// JS objects carry one typed array per field of the C structs, under the field's own name (see shim/README.md for the wire
// format): the shim reads `obj.pos` as Float64Array -> ep_bodies.pos and so on.  Arrays are best allocated with hostAlloc()
// (pinned memory wrapped as external ArrayBuffers) so that ep_step_frame_fields copies at full PCIe rate.
//
// Build: node-gyp with this file, -Iinclude, -lelm_physics_b200.  Compile check without Node: see shim/stub/node_api.h.
#include <node_api.h>

#include <cstdio>
#include <cstring>

#include "elm_physics_b200.h"

namespace {

ep_ctx* g_ctx = nullptr;                    // one context per process: Elm / Node is single-threaded

struct Fail { const char* what; };          // thrown by the getters below, turned into a JS exception at the entry point

napi_value prop(napi_env env, napi_value obj, const char* name) {
  napi_value v = nullptr;
  if (napi_get_named_property(env, obj, name, &v) != napi_ok) throw Fail{name};
  return v;
}
bool absent(napi_env env, napi_value v) {
  napi_valuetype t = napi_undefined;
  return napi_typeof(env, v, &t) != napi_ok || t == napi_undefined || t == napi_null;
}
// typed array `name` of `obj` as a pointer; `optional` arrays may be undefined / null -> nullptr (the C ABI's "not wanted")
template <typename T>
T* typed(napi_env env, napi_value obj, const char* name, napi_typedarray_type want, size_t min_len, bool optional = false) {
  napi_value v = prop(env, obj, name);
  if (absent(env, v)) { if (optional) return nullptr; throw Fail{name}; }
  napi_typedarray_type type; size_t len = 0; void* data = nullptr; napi_value ab; size_t off = 0;
  if (napi_get_typedarray_info(env, v, &type, &len, &data, &ab, &off) != napi_ok || type != want || len < min_len) throw Fail{name};
  return static_cast<T*>(data);
}
double* f64(napi_env env, napi_value o, const char* n, size_t len, bool opt = false) { return typed<double>(env, o, n, napi_float64_array, len, opt); }
int32_t* i32(napi_env env, napi_value o, const char* n, size_t len, bool opt = false) { return typed<int32_t>(env, o, n, napi_int32_array, len, opt); }
int32_t int_prop(napi_env env, napi_value obj, const char* name) {
  int32_t x = 0;
  if (napi_get_value_int32(env, prop(env, obj, name), &x) != napi_ok) throw Fail{name};
  return x;
}

// ---- JS object -> C struct (pointers only) -------------------------------------------------------------------------------
ep_shapes shapes_of(napi_env env, napi_value o) {
  ep_shapes s;
  s.n_shapes = int_prop(env, o, "n_shapes");
  const size_t n = (size_t)s.n_shapes;
  s.kind = i32(env, o, "kind", n); s.f64_off = i32(env, o, "f64_off", n + 1); s.i32_off = i32(env, o, "i32_off", n + 1);
  s.f64 = f64(env, o, "f64", (size_t)s.f64_off[n]); s.i32 = i32(env, o, "i32", (size_t)s.i32_off[n]);
  return s;
}
ep_bodies bodies_of(napi_env env, napi_value o) {
  ep_bodies b;
  b.n_worlds = int_prop(env, o, "n_worlds"); b.n_bodies = int_prop(env, o, "n_bodies"); b.n_insts = int_prop(env, o, "n_insts");
  const size_t n = (size_t)b.n_bodies, ni = (size_t)b.n_insts;
  b.world_body_off = i32(env, o, "world_body_off", (size_t)b.n_worlds + 1);
  b.body_id = i32(env, o, "body_id", n); b.kind = i32(env, o, "kind", n);
  b.pos = f64(env, o, "pos", 3 * n); b.quat = f64(env, o, "quat", 4 * n); b.vel = f64(env, o, "vel", 3 * n); b.angvel = f64(env, o, "angvel", 3 * n);
  b.force = f64(env, o, "force", 3 * n); b.torque = f64(env, o, "torque", 3 * n); b.inv_inertia_world = f64(env, o, "inv_inertia_world", 9 * n);
  b.inv_mass = f64(env, o, "inv_mass", n); b.inv_inertia = f64(env, o, "inv_inertia", 3 * n);
  b.lin_damp_pow = f64(env, o, "lin_damp_pow", n); b.ang_damp_pow = f64(env, o, "ang_damp_pow", n);
  b.lin_lock = f64(env, o, "lin_lock", 3 * n); b.ang_lock = f64(env, o, "ang_lock", 3 * n); b.bounding_radius = f64(env, o, "bounding_radius", n);
  b.inst_off = i32(env, o, "inst_off", n + 1); b.inst_shape = i32(env, o, "inst_shape", ni);
  b.inst_friction = f64(env, o, "inst_friction", ni); b.inst_bounciness = f64(env, o, "inst_bounciness", ni);
  return b;
}
ep_params params_of(napi_env env, napi_value o, const ep_bodies& b) {
  ep_params p;
  p.n_worlds = b.n_worlds;
  const size_t nw = (size_t)p.n_worlds;
  p.gravity = f64(env, o, "gravity", 3 * nw); p.dt = f64(env, o, "dt", nw); p.iterations = i32(env, o, "iterations", nw);
  p.collide = typed<uint8_t>(env, o, "collide", napi_uint8_array, 0, true);          // null == `\_ _ -> True`
  p.collide_off = p.collide ? typed<int64_t>(env, o, "collide_off", napi_bigint64_array, nw) : nullptr;
  p.n_constraints = int_prop(env, o, "n_constraints");
  const size_t nc = (size_t)p.n_constraints;
  p.con_body_a = i32(env, o, "con_body_a", nc, nc == 0); p.con_body_b = i32(env, o, "con_body_b", nc, nc == 0);
  p.con_kind = i32(env, o, "con_kind", nc, nc == 0); p.con_data = f64(env, o, "con_data", 24 * nc, nc == 0);
  return p;
}
// every array of ep_contacts is optional on the way OUT (a missing one is not copied back); as warm-start INPUT the keys,
// lambda and t1 must be there
ep_contacts contacts_of(napi_env env, napi_value o, int n_worlds) {
  ep_contacts c;
  c.n_worlds = n_worlds; c.group_cap = int_prop(env, o, "group_cap"); c.contact_cap = int_prop(env, o, "contact_cap");
  const size_t g = (size_t)c.group_cap, k = (size_t)c.contact_cap, nw = (size_t)n_worlds;
  c.world_group_off = i32(env, o, "world_group_off", nw + 1, true);
  c.group_id1 = i32(env, o, "group_id1", g, true); c.group_id2 = i32(env, o, "group_id2", g, true);
  c.group_n_constraints = i32(env, o, "group_n_constraints", g, true); c.group_contact_off = i32(env, o, "group_contact_off", g + 1, true);
  c.group_pre_pos1 = f64(env, o, "group_pre_pos1", 3 * g, true); c.group_pre_quat1 = f64(env, o, "group_pre_quat1", 4 * g, true);
  c.shape_key = i32(env, o, "shape_key", k, true);
  c.feature_key = typed<int64_t>(env, o, "feature_key", napi_bigint64_array, k, true);
  c.ni = f64(env, o, "ni", 3 * k, true); c.pi = f64(env, o, "pi", 3 * k, true); c.pj = f64(env, o, "pj", 3 * k, true);
  c.friction = f64(env, o, "friction", k, true); c.bounciness = f64(env, o, "bounciness", k, true);
  c.lambda = f64(env, o, "lambda", k, true); c.t1 = f64(env, o, "t1", 3 * k, true);
  c.iterations = i32(env, o, "iterations", nw, true); c.world_n_islands = i32(env, o, "world_n_islands", nw, true);
  c.group_island = i32(env, o, "group_island", g, true);
  return c;
}

napi_value undefined(napi_env env) { napi_value u; napi_get_undefined(env, &u); return u; }
// `simulate` is total on the Elm side: a non-zero return code is a programmer error (capacity, CUDA failure) -> JS exception
napi_value check(napi_env env, int rc) {
  if (rc != EP_OK) { napi_throw_error(env, "EP_ERROR", g_ctx ? ep_last_error(g_ctx) : "no context"); }
  return undefined(env);
}
size_t args(napi_env env, napi_callback_info info, napi_value* argv, size_t want) {
  size_t argc = want;
  if (napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr) != napi_ok) throw Fail{"arguments"};
  return argc;
}
#define EP_ENTRY(...)                                                                          \
  try { __VA_ARGS__ } catch (const Fail& f) {                                                         \
    char msg[160]; snprintf(msg, sizeof msg, "ep_napi: missing or mistyped field `%s`", f.what); \
    napi_throw_error(env, "EP_MARSHAL", msg); return undefined(env);                           \
  }

// uploadShapes(shapes)
napi_value UploadShapes(napi_env env, napi_callback_info info) {
  EP_ENTRY(napi_value a[1]; args(env, info, a, 1); ep_shapes s = shapes_of(env, a[0]); return check(env, ep_upload_shapes(g_ctx, &s));)
}
// step(bodies, params, warmOrNull, out, nSteps): one-call form = BroadPhase.getPairs + Solver.solve for every world
// (Physics.elm:575-579); used when the topology changed or the warm start is not the previous frame
napi_value Step(napi_env env, napi_callback_info info) {
  EP_ENTRY(
    napi_value a[5]; args(env, info, a, 5);
    ep_bodies b = bodies_of(env, a[0]); ep_params p = params_of(env, a[1], b);
    ep_contacts warm, out; const bool has_warm = !absent(env, a[2]);
    if (has_warm) warm = contacts_of(env, a[2], b.n_worlds);
    out = contacts_of(env, a[3], b.n_worlds);
    int32_t n = 1; napi_get_value_int32(env, a[4], &n);
    return check(env, ep_step(g_ctx, &b, &p, has_warm ? &warm : nullptr, &out, n));)
}
// uploadWorlds(bodies, params, warmOrNull): make a batch resident (start of a simulate loop)
napi_value UploadWorlds(napi_env env, napi_callback_info info) {
  EP_ENTRY(
    napi_value a[3]; args(env, info, a, 3);
    ep_bodies b = bodies_of(env, a[0]); ep_params p = params_of(env, a[1], b);
    ep_contacts warm; const bool has_warm = !absent(env, a[2]);
    if (has_warm) warm = contacts_of(env, a[2], b.n_worlds);
    return check(env, ep_upload_worlds(g_ctx, &b, &p, has_warm ? &warm : nullptr));)
}
// stepFrame(bodies, out, uploadMask, downloadMask): one `simulate` call of a loop that feeds the returned Contacts back
// (README.md:21-28): the resident previous frame is the warm start
napi_value StepFrame(napi_env env, napi_callback_info info) {
  EP_ENTRY(
    napi_value a[4]; args(env, info, a, 4);
    ep_bodies b = bodies_of(env, a[0]); ep_contacts out = contacts_of(env, a[1], b.n_worlds);
    uint32_t up = EP_F_ALL, down = EP_F_ALL; napi_get_value_uint32(env, a[2], &up); napi_get_value_uint32(env, a[3], &down);
    return check(env, ep_step_frame_fields(g_ctx, &b, &out, up, down));)
}
// raycast(rays): Physics.raycast (Physics.elm:802-841) for a batch of rays; rays.attached -> rays in body frames
napi_value Raycast(napi_env env, napi_callback_info info) {
  EP_ENTRY(
    napi_value a[1]; args(env, info, a, 1);
    const int32_t n = int_prop(env, a[0], "n_rays"); const size_t k = (size_t)n;
    int32_t attached = 0; napi_get_value_int32(env, prop(env, a[0], "attached"), &attached);
    int32_t* ref = i32(env, a[0], "ray_ref", k); double* from = f64(env, a[0], "from", 3 * k); double* dir = f64(env, a[0], "direction", 3 * k);
    int32_t* hit = i32(env, a[0], "hit_body", k); double* dist = f64(env, a[0], "hit_distance", k);
    double* pt = f64(env, a[0], "hit_point", 3 * k); double* nr = f64(env, a[0], "hit_normal", 3 * k);
    return check(env, attached ? ep_raycast_attached(g_ctx, n, ref, from, dir, hit, dist, pt, nr) : ep_raycast(g_ctx, n, ref, from, dir, hit, dist, pt, nr));)
}
// hostAlloc(bytes) -> ArrayBuffer on page-locked memory (ep_host_alloc), freed when the ArrayBuffer is collected
void free_pinned(napi_env, void* data, void*) { ep_host_free(data); }
napi_value HostAlloc(napi_env env, napi_callback_info info) {
  EP_ENTRY(
    napi_value a[1]; args(env, info, a, 1);
    double bytes = 0; napi_get_value_double(env, a[0], &bytes);
    void* p = ep_host_alloc((size_t)bytes);
    if (!p) { napi_throw_error(env, "EP_ERROR", "ep_host_alloc failed"); return undefined(env); }
    napi_value ab; napi_create_external_arraybuffer(env, p, (size_t)bytes, free_pinned, nullptr, &ab);
    return ab;)
}

}  // namespace

NAPI_MODULE_INIT() {
  if (ep_create(&g_ctx, 0) != EP_OK) {           // there is no CPU fallback: without a usable GPU the addon refuses to load
    napi_throw_error(env, "EP_ENODEVICE", "elm_physics_b200: no usable CUDA device (no CPU fallback)");
    return exports;
  }
  const napi_property_descriptor d[] = {
      {"uploadShapes", nullptr, UploadShapes, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"uploadWorlds", nullptr, UploadWorlds, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"step", nullptr, Step, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"stepFrame", nullptr, StepFrame, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"raycast", nullptr, Raycast, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"hostAlloc", nullptr, HostAlloc, nullptr, nullptr, nullptr, napi_default, nullptr},
  };
  napi_define_properties(env, exports, sizeof d / sizeof d[0], d);
  return exports;
}
