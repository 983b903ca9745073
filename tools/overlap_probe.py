#!/usr/bin/env python
"""Experiment: K independent contexts (world sub-batches) on one GPU, stepped round-robin so that the latency-bound
phases of one batch overlap the throughput-bound phases of the others."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from elm_physics_b200 import engine as eng, scenes
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
for K in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]:
    es = []
    for k in range(K):
        shapes, bodies, params = scenes.mixed_worlds(W // K, first_world=k * (W // K))
        e = eng.Engine(0); e.upload_shapes(shapes); e.upload_worlds(bodies, params, None); es.append(e)
    for e in es: e.step_resident(120)
    for e in es: e.sync()
    t0 = time.perf_counter()
    steps = 40
    for s in range(steps):
        for e in es: e.step_resident(1)
    for e in es: e.sync()
    dt = time.perf_counter() - t0
    print(f"K={K}: {1e3*dt/steps:.3f} ms/step  {W*32*steps/dt/1e6:.1f} M body-steps/s", flush=True)
    for e in es: e.close()
