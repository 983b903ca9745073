#!/usr/bin/env python
"""Small scenes of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck): the smoke scenes, a C4 pile
(team solver, chunked broadphase), C5 car worlds (joint rows, raycast, apply) and a C3 batch (lane tiles).
Usage: compute-sanitizer --tool racecheck python tools/sanitize_scenes.py [steps]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from elm_physics_b200 import engine as eng, pack, scenes
import helpers as H

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
e = eng.Engine(0)
for name, (shapes, bodies, params) in [
        ("readme", pack.pack_worlds(scenes.readme_scene())),
        ("mixed64", scenes.mixed_worlds(64)),
        ("boxes", pack.pack_worlds(scenes.boxes_scene())),
        ("pile150", pack.pack_worlds(scenes.pile_scene(150))),
        ("cars24", pack.pack_worlds(scenes.car_worlds(24)))]:
    e.upload_shapes(shapes)
    e.upload_worlds(bodies, params, None)
    e.step_resident(steps)
    e.sync()
    out = H.contacts_for(bodies)
    e.download_contacts(out)
    if name == "cars24":
        nw = bodies.n_worlds
        rw = np.repeat(np.arange(nw, dtype=np.int32), 4)
        frm = np.tile(np.array([[0.0, 0.0, 5.0]]), (4 * nw, 1))
        dr = np.tile(np.array([[0.0, 0.0, -1.0]]), (4 * nw, 1))
        e.raycast(rw, frm, dr)
    print(name, "ok", e.stats()["contacts"], "contacts", flush=True)
e.close()
