#!/usr/bin/env python
"""k_solve_cluster's level schedule: the parallel build against the one-thread walk (EP_SCHED_SERIAL=1) on config 4 and config 2 --
the two must give the same bodies bit for bit after hundreds of steps (a schedule that broke a dependency would not), and the
step times side by side.  Usage: python tools/sched_check.py [pile bodies]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(shapes, bodies, params, serial, preroll, steps):
    from elm_physics_b200 import engine as eng
    os.environ["EP_SCHED_SERIAL"] = "1" if serial else "0"
    e = eng.Engine(0)
    e.upload_shapes(shapes)
    b = bodies.copy()
    e.upload_worlds(b, params, None)
    e.step_resident(preroll)
    e.sync()
    t0 = time.perf_counter()
    e.step_resident(steps)
    e.sync()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    e.download_bodies(b)
    e.close()
    return b, ms


def main():
    from elm_physics_b200 import abi, pack, scenes
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    ok = True
    for name, worlds, pre in (("pile:%d" % n, scenes.pile_scene(n), 600), ("boxes", scenes.boxes_scene(), 300)):
        shapes, bodies, params = pack.pack_worlds(worlds)
        bs, ms_s = run(shapes, bodies, params, True, pre, 60)
        bp, ms_p = run(shapes, bodies, params, False, pre, 60)
        same = all(getattr(bs, f).tobytes() == getattr(bp, f).tobytes() for f in abi.BODY_INOUT)
        ok = ok and same
        print("%s: serial walk %.3f ms/step, parallel build %.3f ms/step, bodies bit-identical after %d steps: %s" % (name, ms_s, ms_p, pre + 60, same))
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
