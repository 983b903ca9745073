"""GPU parity for ONE GIANT ISLAND (BASELINE configs 2 and 4 in small): every variant and fallback of the thread-block-cluster
solver (csrc/ep_solve_cluster.cuh) against the CPU oracle, bit for bit.  The variants are chosen when a context is created
(EP_SOLO / EP_CLUSTER / EP_CLUSTER_WS / EP_CLUSTER_BODIES are read by ep_create), so every case builds its own Engine.

Reference behaviour under test: src/Internal/Solver.elm:39-208, 345-489 (warm start, sweeps, early exit) on islands of
hundreds of pair groups, with joint rows (src/Internal/Equation.elm:157-361) inside them.
"""
import os

import numpy as np
import pytest

from elm_physics_b200 import pack
from elm_physics_b200 import physics as P
from elm_physics_b200 import scenes

import helpers as H

pytestmark = pytest.mark.gpu

VARIANTS = {
    "solo": {},                                                   # one CTA, bodies in its shared memory
    "cluster16": {"EP_SOLO": "0"},                                # 16-CTA cluster, bodies in distributed shared memory
    "cluster2": {"EP_SOLO": "0", "EP_CLUSTER": "2"},
    "cluster_global_bodies": {"EP_SOLO": "0", "EP_CLUSTER_BODIES": "4"},      # the table is too small: global solver-body array
    "solo_in_place": {"EP_CLUSTER_WS": "0"},                      # no window: every group solved in place
    "cluster_small_window": {"EP_SOLO": "0", "EP_CLUSTER_WS": "6"},           # groups that do not fit: mixed window / in-place steps
    "team_kernel": {"EP_CLUSTER": "0"},                           # k_solve_team<256>, the single-SM solver
}


class _Env:
    def __init__(self, env):
        self.env, self.old = env, {}

    def __enter__(self):
        for k, v in self.env.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = v

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _engine(env):
    from elm_physics_b200 import engine as eng
    with _Env(env):
        return eng.Engine(0)


def _assert_variant_ran(e, variant):
    """The launch counters of the context: the cluster launches exist exactly when the variant asks for them."""
    launches = e.kernel_times()["k_solve_cluster"][1]
    if variant == "team_kernel":
        assert launches == 0
    else:
        assert launches > 0


def _chain_world(n=96):
    """A row of n boxes resting on the plane, neighbours overlapping by 5 mm (box-box contacts) and linked alternately by a
    hinge and a point-to-point constraint, a second shorter row stacked on top: one island of ~2 n pair groups whose rows are
    contacts AND joint rows."""
    bodies = [("floor", P.plane(P.wood))]
    cons = {}
    for i in range(n):
        b = P.move_to((0.995 * i, 0.0, 0.5), P.block((1.0, 1.0, 1.0), P.wood))
        bodies.append((f"a{i}", P.set_velocity_to((0.0, 0.05 * ((i % 3) - 1), 0.0), b)))
        if i:
            cons[(f"a{i - 1}", f"a{i}")] = ([P.Hinge((0.5, 0, 0), (0, 1, 0), (-0.495, 0, 0), (0, 1, 0))] if i % 2
                                            else [P.PointToPoint((0.5, 0.2, 0.1), (-0.495, 0.2, 0.1))])
    for i in range(n // 2):
        bodies.append((f"b{i}", P.move_to((0.995 * (2 * i) + 0.5, 0.0, 1.5), P.block((1.0, 1.0, 1.0), P.wood))))
    cfg = P.on_earth()

    def constrain(a):
        tab = {b: v for (x, b), v in cons.items() if x == a}
        return (lambda b: tab.get(b, [])) if tab else None
    cfg.constrain = constrain
    with_ids, _ = P.assign_ids(bodies)
    return [(cfg, with_ids)]


@pytest.mark.parametrize("variant", list(VARIANTS))
def test_boxes_pyramid_every_solver_variant(oracle, variant):
    """Config 2 (100 boxes, one island of ~240 groups / 700 contacts / ~69 levels), 60 resident steps then 12 lock-step frames."""
    shapes, bodies, params = pack.pack_worlds(scenes.boxes_scene())
    e = _engine(VARIANTS[variant])
    try:
        _, bg, _, og = H.resident_vs_oracle(oracle, e, shapes, bodies, params, 60, what=f"boxes-{variant}")
        assert og.n_contacts > 500 and og.world_n_islands[0] <= 3
        H.lockstep(oracle, e, shapes, bg, params, 12, what=f"boxes-{variant}-lockstep")
        _assert_variant_ran(e, variant)
    finally:
        e.close()


@pytest.mark.parametrize("variant", ["solo", "cluster16", "cluster_global_bodies", "solo_in_place", "team_kernel"])
def test_giant_island_with_joint_rows(oracle, variant):
    """Contacts and joint rows in one island of class 9: the groups with joint rows are solved in place by the quads (sweep 1)
    and from the window (sweep 2)."""
    shapes, bodies, params = pack.pack_worlds(_chain_world())
    assert params.n_constraints == 95
    e = _engine(VARIANTS[variant])
    try:
        _, bg, _, og = H.resident_vs_oracle(oracle, e, shapes, bodies, params, 40, what=f"chain-{variant}")
        ng = og.n_groups
        assert int(og.group_n_constraints[:ng].sum()) == 95
        biggest = np.bincount(og.group_island[:ng][og.group_island[:ng] >= 0]).max()
        assert biggest > 250, biggest                      # one island holds (nearly) everything
        assert og.n_contacts > 600
        H.lockstep(oracle, e, shapes, bg, params, 8, what=f"chain-{variant}-lockstep")
        assert np.isfinite(bg.pos).all()
        _assert_variant_ran(e, variant)
    finally:
        e.close()


@pytest.mark.parametrize("variant", ["cluster16", "cluster_small_window", "cluster_global_bodies"])
def test_convex_pile_cluster_variants(oracle, variant):
    """C4 at test size (150 hulls / compounds): hull-hull groups of 1..8 contacts, compounds with several groups per body pair."""
    shapes, bodies, params = pack.pack_worlds(scenes.pile_scene(150, layers=10))
    e = _engine(VARIANTS[variant])
    try:
        _, _, _, og = H.resident_vs_oracle(oracle, e, shapes, bodies, params, 260, what=f"pile150-{variant}")
        assert og.n_contacts > 400
    finally:
        e.close()
