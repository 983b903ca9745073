"""The reference-side shim (shim/): the N-API marshalling compiles against a stand-in for Node's header, names only entry
points the C ABI declares, and the wire format it and the Elm encoder document (shim/README.md) -- mirrored by
elm_physics_b200/wire.py -- produces, array by array, what the packer behind every GPU parity test produces."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

from elm_physics_b200 import abi, pack, scenes, wire
from elm_physics_b200 import physics as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_napi_shim_compiles_against_the_stub_header():
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    r = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "shim", "stub"),
                        "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "shim", "ep_napi.cc")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_shim_calls_only_declared_entry_points_and_uses_the_c_field_names():
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "elm_physics_b200.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(ep_[a-z0-9_]+)\s*\(", header))
    src = re.sub(r"//.*", "", open(os.path.join(ROOT, "shim", "ep_napi.cc")).read())
    used = set(re.findall(r"\b(ep_[a-z0-9_]+)\s*\(", src)) - {"ep_napi"}
    assert used and used <= declared, used - declared
    # every field the shim reads by name is a member of the C struct it fills (and of the ctypes mirror)
    for struct, mirror in (("ep_bodies", abi.EpBodies), ("ep_contacts", abi.EpContacts), ("ep_params", abi.EpParams), ("ep_shapes", abi.EpShapes)):
        members = {n.rstrip("_") for n, _ in mirror._fields_}
        fn = {"ep_bodies": "bodies_of", "ep_contacts": "contacts_of", "ep_params": "params_of", "ep_shapes": "shapes_of"}[struct]
        body = src[src.index(struct + " " + fn):]
        body = body[:body.index("\n}\n")]
        names = set(re.findall(r'"([a-z0-9_]+)"', body))
        assert names <= members | {"lambda"}, (struct, names - members)
        # the JS runtime and the Elm encoder use the same names
    js = open(os.path.join(ROOT, "shim", "patch_simulate.js")).read()
    elm = open(os.path.join(ROOT, "shim", "elm", "Physics", "B200", "Encode.elm")).read()
    for name, _ in abi.BODY_F64_FIELDS:
        assert name in js, name
        assert ('"%s"' % name) in elm or name in ("lin_damp_pow", "ang_damp_pow"), name       # the damping factors are made by Math.pow in JS


@pytest.mark.parametrize("scene", ["readme", "boxes", "mixed", "cars", "collide"])
def test_wire_format_equals_the_packer(scene):
    if scene == "readme":
        world = scenes.readme_scene()[0]
    elif scene == "boxes":
        world = scenes.boxes_scene()[0]
    elif scene == "mixed":
        world = scenes.small_mixed_world(seed=3)[0]
    elif scene == "cars":
        world = scenes.car_world(3)
    else:
        world = scenes.stack_scene(3, 10)[0]
        world[0].collide = lambda a, b: not ({a, b} == {"b0", "b1"})
    cfg, bodies = world
    s0, b0, p0 = pack.pack_worlds([(cfg, bodies)])
    s1, b1, p1 = wire.to_abi(*wire.encode_world(cfg, bodies))
    for name in ("kind", "f64_off", "i32_off", "f64", "i32"):
        assert getattr(s0, name).tobytes() == getattr(s1, name).tobytes(), name
    for name in ("world_body_off", "body_id", "kind", "inst_off", "inst_shape", "inst_friction", "inst_bounciness"):
        assert getattr(b0, name).tobytes() == getattr(b1, name).tobytes(), name
    for name, _ in abi.BODY_F64_FIELDS:
        assert getattr(b0, name).tobytes() == getattr(b1, name).tobytes(), name
    for name in ("gravity", "dt", "iterations", "con_body_a", "con_body_b", "con_kind", "con_data"):
        assert getattr(p0, name).tobytes() == getattr(p1, name).tobytes(), name
    assert (p0.collide is None) == (p1.collide is None)
    if p0.collide is not None:
        assert p0.collide.tobytes() == p1.collide.tobytes()
    if scene == "cars":
        assert p1.n_constraints == 5
