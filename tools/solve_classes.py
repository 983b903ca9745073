#!/usr/bin/env python
"""Solver work per island size class on the settled C3 scene (islands, contacts, tiles) next to the stand-alone duration of
each class's launch (EP_SOLVE_SERIAL=1): tells which classes are bound by the latency of their longest tile and which by
throughput.  Usage: EP_SOLVE_SERIAL=1 python tools/solve_classes.py [worlds]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from elm_physics_b200 import engine as eng, scenes
import helpers as H
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
shapes, bodies, params = scenes.mixed_worlds(W)
e = eng.Engine(0); e.upload_shapes(shapes); e.upload_worlds(bodies, params, None)
e.step_resident(120); e.sync()
e.profile(True); e.step_resident(20); e.sync(); kt = e.kernel_times(); e.profile(False)
out = H.contacts_for(bodies); e.download_contacts(out)
ng = out.n_groups
gi = out.group_island[:ng].astype(np.int64)
world = np.repeat(np.arange(W), np.diff(out.world_group_off[:W + 1]))
cn = np.diff(out.group_contact_off[:ng + 1])
key = world * 64 + gi
ok = gi >= 0
uk, inv = np.unique(key[ok], return_inverse=True)
isl_c = np.bincount(inv, weights=cn[ok]).astype(int)
isl_g = np.bincount(inv)
rows = 3 * isl_c
bins = np.zeros(len(rows), int)
for b in range(1, 12): bins[rows > (3 << (b - 1))] = b
res = {}
for b in range(12):
    m = bins == b
    if m.any(): res[b] = {"islands": int(m.sum()), "contacts": int(isl_c[m].sum()), "tiles": int((m.sum() + 31) // 32), "mean_groups": float(isl_g[m].mean())}
print(json.dumps({"classes": res, "kernels_ms": {k: round(v[0] / max(v[1], 1), 4) for k, v in kt.items() if v[1] and "solve" in k}}))
e.close()
