#!/usr/bin/env python
"""Device-time a resident scene step by step (per-kernel CUDA-event breakdown).  Usage:
    python tools/time_scene.py boxes|mixed:<worlds>|pile:<bodies> [--steps K] [--preroll P]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scene")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--preroll", type=int, default=60)
    a = ap.parse_args()
    from elm_physics_b200 import engine as eng, pack, scenes
    name, _, arg = a.scene.partition(":")
    if name == "boxes":
        shapes, bodies, params = pack.pack_worlds(scenes.boxes_scene())
    elif name == "stack":
        shapes, bodies, params = pack.pack_worlds(scenes.stack_scene(int(arg or 5), 20))
    elif name == "mixed":
        shapes, bodies, params = scenes.mixed_worlds(int(arg or 1024))
    elif name == "pile":
        shapes, bodies, params = pack.pack_worlds(scenes.pile_scene(int(arg or 2000), iterations=int(os.environ.get('PILE_ITERS', 20))))
    elif name == "cars":
        shapes, bodies, params = pack.pack_worlds(scenes.car_worlds(64))
        shapes, bodies, params = scenes.replicate_worlds(shapes, bodies, params, max(1, int(arg or 4096) // 64))
    else:
        raise SystemExit("unknown scene")
    e = eng.Engine(0)
    e.upload_shapes(shapes)
    e.upload_worlds(bodies, params, None)
    e.step_resident(a.preroll)
    e.sync()
    import time
    e.step_resident(3)
    e.sync()
    t0 = time.perf_counter()
    e.step_resident(a.steps)
    e.sync()
    wall_ms = (time.perf_counter() - t0) * 1e3 / a.steps      # the step time: kernels overlap on side streams, their brackets do not add up
    e.profile(True)
    e.step_resident(a.steps)
    e.sync()
    kt = e.kernel_times()
    e.profile(False)
    st = e.stats()
    import helpers as H
    from elm_physics_b200 import abi
    nb = bodies.n_bodies
    out = H.contacts_for(bodies) if nb < 5000 or bodies.n_worlds > 1 else abi.Contacts(bodies.n_worlds, 40 * nb, 160 * nb)
    e.download_contacts(out)
    st = e.stats()
    if bodies.n_worlds == 1:      # level schedule of the largest island (what bounds k_solve_team): levels, sum over levels of the longest group
        import numpy as np
        ng = out.n_groups
        gi, id1, id2 = out.group_island[:ng], out.group_id1[:ng], out.group_id2[:ng]
        cn = np.diff(out.group_contact_off[:ng + 1]) + out.group_n_constraints[:ng] * 3
        kind = {int(i): int(k) for i, k in zip(bodies.body_id, bodies.kind)}
        big = np.bincount(gi[gi >= 0]).argmax()
        last, lv_max = {}, {}
        for g in np.nonzero(gi == big)[0]:
            lv = max(last.get(int(b), 0) for b in (id1[g], id2[g]) if kind[int(b)] == 2)
            for b in (id1[g], id2[g]):
                if kind[int(b)] == 2:
                    last[int(b)] = lv + 1
            lv_max[lv] = max(lv_max.get(lv, 0), int(cn[g]))
        print(json.dumps({"largest_island_groups": int((gi == big).sum()), "levels": len(lv_max), "sum_over_levels_of_longest_group_rows": int(sum(lv_max.values())),
                          "contacts_in_island": int(cn[gi == big].sum())}))
    main_k = {k: v for k, v in kt.items() if not k.startswith("k_solve_")}
    tot = sum(v[0] for v in main_k.values())
    print(json.dumps({"scene": a.scene, "bodies": bodies.n_bodies, "ms_per_step": wall_ms, "sum_of_brackets_ms": tot / a.steps,
                      "kernels_ms": {k: round(v[0] / max(v[1], 1), 4) for k, v in kt.items() if v[1]},
                      "contacts": st["contacts"], "groups": st["groups"], "islands": st["islands"], "candidate_pairs": st["candidate_pairs"]}))
    e.close()


if __name__ == "__main__":
    main()
