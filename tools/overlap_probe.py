#!/usr/bin/env python
"""Experiment: K independent contexts (world sub-batches) on one GPU, stepped round-robin (one host thread) or from one host
thread per context, so that the latency-bound phases of one sub-batch overlap the throughput-bound phases of the others.
Usage: python tools/overlap_probe.py <worlds> <K list> [threads]"""
import sys, os, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from elm_physics_b200 import engine as eng, scenes
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
threaded = len(sys.argv) > 3 and sys.argv[3] == "threads"
for K in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,2,4,8").split(",")]:
    es = []
    for k in range(K):
        shapes, bodies, params = scenes.mixed_worlds(W // K, first_world=k * (W // K))
        e = eng.Engine(0); e.upload_shapes(shapes); e.upload_worlds(bodies, params, None); es.append(e)
    for e in es: e.step_resident(120)
    for e in es: e.sync()
    steps = 40
    t0 = time.perf_counter()
    if threaded:
        def run(e):
            e.step_resident(steps); e.sync()
        th = [threading.Thread(target=run, args=(e,)) for e in es]
        for t in th: t.start()
        for t in th: t.join()
    else:
        for s in range(steps):
            for e in es: e.step_resident(1)
        for e in es: e.sync()
    dt = time.perf_counter() - t0
    print(f"W={W} K={K} {'threads' if threaded else 'round-robin'}: {1e3*dt/steps:.3f} ms/step  {W*32*steps/dt/1e6:.1f} M body-steps/s", flush=True)
    for e in es: e.close()
