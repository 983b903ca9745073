#!/usr/bin/env node
// patch_simulate.js — routes `Physics.simulate` of a COMPILED Elm program through the N-API addon (shim/ep_napi.cc) and the
// C ABI of include/elm_physics_b200.h.
//
// Why a patch of the generated JS: `simulate` (/root/reference/src/Physics.elm:522-583) is a synchronous pure function, Elm
// ports are asynchronous and packages cannot declare them, so the seam is the compiled function itself.  `elm make` emits
//     var $w0rm$elm_physics$Physics$simulate = F2(function (config, bodiesWithIds) { ... });
// (the name is stable: author$project$Module$function).  This script replaces its body by `__ep_simulate(config, bodiesWithIds)`
// and prepends the runtime below.  Build WITHOUT --optimize, or with it plus the field map of your build: --optimize renames
// record fields (a.id -> a.a ...), and the runtime reads Elm records by field name.
//
//     node shim/patch_simulate.js elm.js > elm.b200.js
//
// Nothing here could be run in the build image (no node, no elm).  The marshalling it performs -- Elm list of bodies -> one
// typed array per field of ep_bodies, Math.pow for the damping factors (src/Internal/SolverBody.elm:149-153), and back -- is
// mirrored line by line by elm_physics_b200/wire.py, which tests/test_wire_format.py checks against the packer the GPU
// parity tests use.
'use strict';

const RUNTIME = String.raw`
var __ep = require('./ep_napi.node');
var __ep_state = { token: 0, nBodies: -1, ids: null, bodies: null, out: null, shapes: new Map(), shapeRows: [] };

function __ep_list(l) { var a = []; for (; l.b; l = l.b) a.push(l.a); return a; }          // Elm List -> JS array ($: 1 = cons)
function __ep_f64(n) { return new Float64Array(__ep.hostAlloc(8 * Math.max(n, 1)), 0, n); } // pinned (ep_host_alloc)
function __ep_i32(n) { return new Int32Array(__ep.hostAlloc(4 * Math.max(n, 1)), 0, n); }

// ep_bodies for ONE world: field names == the C struct's (shim/README.md)
function __ep_alloc_bodies(n, nInsts) {
  return { n_worlds: 1, n_bodies: n, n_insts: nInsts, world_body_off: Int32Array.of(0, n), body_id: __ep_i32(n), kind: __ep_i32(n),
    pos: __ep_f64(3 * n), quat: __ep_f64(4 * n), vel: __ep_f64(3 * n), angvel: __ep_f64(3 * n), force: __ep_f64(3 * n), torque: __ep_f64(3 * n),
    inv_inertia_world: __ep_f64(9 * n), inv_mass: __ep_f64(n), inv_inertia: __ep_f64(3 * n), lin_damp_pow: __ep_f64(n), ang_damp_pow: __ep_f64(n),
    lin_lock: __ep_f64(3 * n), ang_lock: __ep_f64(3 * n), bounding_radius: __ep_f64(n),
    inst_off: __ep_i32(n + 1), inst_shape: __ep_i32(nInsts), inst_friction: __ep_f64(nInsts), inst_bounciness: __ep_f64(nInsts) };
}
function __ep_put3(a, i, v) { a[3 * i] = v.x; a[3 * i + 1] = v.y; a[3 * i + 2] = v.z; }
// Internal.Body.Body (src/Internal/Body.elm:36-61) -> row i.  transform3d = Transform3d origin (Orientation3d x y z w)
// (src/Internal/Transform3d.elm:30-31, 379-380): a = origin, b = { a: x, b: y, c: z, d: w } in the compiled representation.
function __ep_put_body(B, i, body, dt) {
  var t = body.transform3d, o = t.a, q = t.b, m = body.invInertiaWorld;
  B.body_id[i] = body.id; B.kind[i] = body.kindInt;
  __ep_put3(B.pos, i, o); B.quat[4 * i] = q.a; B.quat[4 * i + 1] = q.b; B.quat[4 * i + 2] = q.c; B.quat[4 * i + 3] = q.d;
  __ep_put3(B.vel, i, body.velocity); __ep_put3(B.angvel, i, body.angularVelocity); __ep_put3(B.force, i, body.force); __ep_put3(B.torque, i, body.torque);
  B.inv_inertia_world.set([m.m11, m.m21, m.m31, m.m12, m.m22, m.m32, m.m13, m.m23, m.m33], 9 * i);
  B.inv_mass[i] = body.invMass; __ep_put3(B.inv_inertia, i, body.invInertia);
  B.lin_damp_pow[i] = Math.pow(1 - body.linearDamping, dt);          // the HOST language's pow, as SolverBody.elm:149-153
  B.ang_damp_pow[i] = Math.pow(1 - body.angularDamping, dt);
  __ep_put3(B.lin_lock, i, body.linearLock); __ep_put3(B.ang_lock, i, body.angularLock);
  B.bounding_radius[i] = body.geometry.boundingSphereRadius;
}

// ep_contacts with only what Physics.contactPoints reads (src/Physics.elm:1158-1207); the other arrays stay undefined = NULL =
// not copied back (the rest of the frame is the next call's warm start and stays resident on the device)
function __ep_alloc_contacts(n) {
  var g = n * (n - 1) / 2 + 1, c = Math.max(16384, 16 * n);
  return { group_cap: g, contact_cap: c, world_group_off: __ep_i32(2), group_id1: __ep_i32(g), group_id2: __ep_i32(g),
    group_contact_off: __ep_i32(g + 1), group_pre_pos1: __ep_f64(3 * g), group_pre_quat1: __ep_f64(4 * g), pi: __ep_f64(3 * c) };
}
// Shape table (ep_shapes): one row per distinct CoM-frame shape OBJECT (shapes are immutable after construction,
// src/Internal/Body.elm:64-67, so identity is a sound cache key); rows are produced by the compiled Elm encoder
// Physics.B200.Encode.shapeTable so that the convex layout has a single definition.
function __ep_upload_new_shapes(list, S) {
  var B = S.bodies, k = 0, fresh = [];
  for (var i = 0; i < list.length; i++) {
    var sm = __ep_list(list[i].b.a.geometry.shapesWithMaterials);
    B.inst_off[i] = k;
    for (var j = 0; j < sm.length; j++, k++) {
      var shape = sm[j].a, mat = sm[j].b;
      if (!S.shapes.has(shape)) { S.shapes.set(shape, S.shapeRows.length); S.shapeRows.push(shape); fresh.push(shape); }
      B.inst_shape[k] = S.shapes.get(shape); B.inst_friction[k] = mat.friction; B.inst_bounciness[k] = mat.bounciness;
    }
  }
  B.inst_off[list.length] = k;
  if (fresh.length || !S.shapeTable) {
    var t = __ep_json($author$project$Physics$B200$Encode$shapeTable(__ep_to_list(S.shapeRows)));      // Json.Encode.Value is the plain JS value
    S.shapeTable = { n_shapes: t.n_shapes, kind: Int32Array.from(t.kind), f64_off: Int32Array.from(t.f64_off), i32_off: Int32Array.from(t.i32_off),
                     f64: Float64Array.from(t.f64), i32: Int32Array.from(t.i32) };
  }
}
function __ep_json(v) { return v.a !== undefined && v.$ !== undefined ? v.a : v; }         // _Json_wrap
function __ep_to_list(a) { var l = { $: 0 }; for (var i = a.length - 1; i >= 0; i--) l = { $: 1, a: a[i], b: l }; return l; }
// ep_params: the user callbacks are evaluated HERE, by the compiled Elm closures, for every ordered pair up front
// (src/Internal/BroadPhase.elm:144-215; purity makes the extra calls harmless); pivots / axes go through the compiled
// Constraint.relativeToCenterOfMass (src/Internal/Constraint.elm:16-47) -- no re-implementation in JS
function __ep_params(config, list) {
  var n = list.length, g = config.gravity, rows = [];
  var collide = new Uint8Array(n * n);
  for (var i = 0; i < n; i++) for (var j = 0; j < n; j++) collide[i * n + j] = A2(config.collide, list[i].a, list[j].a) ? 1 : 0;
  for (var i = 0; i < n; i++) {
    var f = config.constrain(list[i].a);                      // Maybe (id -> List Constraint)
    if (f.$ !== 0) continue;                                  // Nothing
    for (var j = 0; j < n; j++) {
      if (j === i) continue;
      var cs = __ep_list(f.a(list[j].a));
      for (var c = 0; c < cs.length; c++) rows.push(__ep_constraint_row(i, j, cs[c], list[i].b.a, list[j].b.a));
    }
  }
  return { gravity: Float64Array.of(g.x, g.y, g.z), dt: Float64Array.of(config.duration), iterations: Int32Array.of(config.solverIterations),
    collide: collide, collide_off: BigInt64Array.of(0n), n_constraints: rows.length,
    con_body_a: Int32Array.from(rows, function (r) { return r.a; }), con_body_b: Int32Array.from(rows, function (r) { return r.b; }),
    con_kind: Int32Array.from(rows, function (r) { return r.kind; }), con_data: Float64Array.from([].concat.apply([], rows.map(function (r) { return r.data; }))) };
}
// kind codes of include/elm_physics_b200.h; data = 12 doubles of the a side, 12 of the b side (distance in [0])
function __ep_constraint_row(a, b, c, bodyA, bodyB) {
  var rel = A3($w0rm$elm_physics$Internal$Constraint$relativeToCenterOfMass, bodyA.centerOfMassTransform3d, bodyB.centerOfMassTransform3d, c);
  var z = new Array(24).fill(0), v = function (o, p) { z[o] = p.x; z[o + 1] = p.y; z[o + 2] = p.z; };
  switch (rel.$) {          // PointToPoint v v | Hinge v v v v | Lock 8 x v | Distance f (src/Internal/Constraint.elm:9-13)
    case 0: v(0, rel.a); v(12, rel.b); return { a: a, b: b, kind: 0, data: z };
    case 1: v(0, rel.a); v(3, rel.b); v(12, rel.c); v(15, rel.d); return { a: a, b: b, kind: 1, data: z };
    case 2: v(0, rel.a); v(3, rel.b); v(6, rel.c); v(9, rel.d); v(12, rel.e); v(15, rel.f); v(18, rel.g); v(21, rel.h); return { a: a, b: b, kind: 2, data: z };
    default: z[0] = rel.a; return { a: a, b: b, kind: 3, data: z };
  }
}
// Result: bodies in the caller's order with the returned state written in by the compiled Physics.B200.Decode.applyState
// (it re-derives invInertiaWorld, world shapes and the cleared force / torque exactly as SolverBody.solved does), and a
// Contacts value that carries the frame token + the arrays contactPoints reads
function __ep_result(assigned, list, B, S) {
  var out = [];
  for (var i = 0; i < list.length; i++) {
    var st = [B.pos[3 * i], B.pos[3 * i + 1], B.pos[3 * i + 2], B.quat[4 * i], B.quat[4 * i + 1], B.quat[4 * i + 2], B.quat[4 * i + 3],
              B.vel[3 * i], B.vel[3 * i + 1], B.vel[3 * i + 2], B.angvel[3 * i], B.angvel[3 * i + 1], B.angvel[3 * i + 2]];
    var body = A2($author$project$Physics$B200$Decode$applyState, __ep_to_list(st), list[i].b.a);
    out.push({ a: list[i].a, b: { $: 0, a: body } });                    // ( id, Types.Body internal )
  }
  var contacts = { $: 0, a: { __ep_token: S.token, __ep_frame: S.out, iterations: 0, bodies: out } };
  return { a: __ep_to_list(out), b: contacts };
}

// The replacement of $w0rm$elm_physics$Physics$simulate: same arguments, same result shape ( List ( id, Body ), Contacts id ).
function __ep_simulate(config, bodiesWithIds) {
  // 1. id assignment stays in Elm (src/Internal/AssignIds.elm:7-27): call the compiled function
  var assigned = $w0rm$elm_physics$Internal$AssignIds$assignIds(bodiesWithIds);
  var list = __ep_list(assigned.bodies), n = list.length;
  var dt = config.duration, g = config.gravity;                 // already unwrapped Floats in the compiled record (Physics.elm:534-538)
  var S = __ep_state, sameLoop = S.nBodies === n && config.contacts.a && config.contacts.a.__ep_token === S.token;
  if (!sameLoop) {
    // new topology or a Contacts value that is not the previous frame: (re)build everything and use the one-call form
    var nInsts = 0; for (var i = 0; i < n; i++) nInsts += __ep_list(list[i].b.a.geometry.shapesWithMaterials).length;
    S.bodies = __ep_alloc_bodies(n, nInsts); S.nBodies = n;
    S.out = __ep_alloc_contacts(n);
    __ep_upload_new_shapes(list, S);                               // ep_upload_shapes once per distinct shape set
  }
  var B = S.bodies;
  for (var i = 0; i < n; i++) __ep_put_body(B, i, list[i].b.a, dt);   // ( id, Body internal ): tuple = { a, b }, Body = { $: 0, a: internal }
  var params = __ep_params(config, list);                          // collide / constrain callbacks evaluated here, in Elm's own code
  if (sameLoop) __ep.stepFrame(B, S.out, 15 | (config.__ep_dirty || 0), 15);          // EP_F_STATE up and down (+ force / torque when pushed)
  else { __ep.uploadShapes(S.shapeTable); __ep.step(B, params, null, S.out, 1); __ep.uploadWorlds(B, params, S.out); }
  S.token++;
  // 2. back: bodies in the caller's order (Physics.elm:581, 1210-1232), Contacts = opaque frame token + lazy views for contactPoints
  return __ep_result(assigned, list, B, S);
}
`;

function patch(source) {
  const re = /var \$w0rm\$elm_physics\$Physics\$simulate = F2\(\s*function \(([a-zA-Z_$][\w$]*), ([a-zA-Z_$][\w$]*)\) \{/;
  const m = re.exec(source);
  if (!m) throw new Error('compiled Physics.simulate not found: is this an elm make output that uses w0rm/elm-physics?');
  const head = m[0] + ` return __ep_simulate(${m[1]}, ${m[2]});`;
  return RUNTIME + source.slice(0, m.index) + head + source.slice(m.index + m[0].length);
}

if (require.main === module) {
  const fs = require('fs');
  process.stdout.write(patch(fs.readFileSync(process.argv[2], 'utf8')));
}
module.exports = { patch };
